set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_render_parity.py tests/test_gpu_fuzz.py tests/test_gpu_builder.py tests/test_gpu_textures.py tests/test_gpu_full_size.py tests/test_gpu_mesh_import.py -x -q -m gpu 2>&1 | tail -15
python tools/pool_stats.py 2>&1 | tail -16
tools/sweep_pool.sh "-DRTC_POOL_STACK=2 -DRTC_POOL_BLOCKS=5" "-DRTC_POOL_STACK=2 -DRTC_POOL_BLOCKS=4" "-DRTC_POOL_STACK=3 -DRTC_POOL_BLOCKS=4 -DRTC_POOL_FETCH=8" "-DRTC_POOL_STACK=3 -DRTC_POOL_BLOCKS=4 -DRTC_POOL_FETCH=24" "-DRTC_POOL_STACK=3 -DRTC_POOL_TWEIGHT=3" "-DRTC_POOL_STACK=3 -DRTC_POOL_TWEIGHT=1" 2>&1
# leave the default build in place and capture the traversal kernels of one small step
touch tweeker_raytracer_b200/csrc/kernels_trace.cu tweeker_raytracer_b200/csrc/kernels_shade.cu; make -s -j4 core host
ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_extend_primary' -c 4 -o gpurun_out/prof_pool_r2a python bench.py --steps 1 --warmup 0 --spp-per-step 8 --no-cpu-baseline > gpurun_out/ncu_pool_r2a.log 2>&1
tail -3 gpurun_out/ncu_pool_r2a.log
