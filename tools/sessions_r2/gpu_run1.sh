set -x
cd $GRAFT_REPO_ROOT
nvidia-smi -L
python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_render_parity.py tests/test_gpu_fuzz.py tests/test_gpu_builder.py -x -q -m gpu 2>&1 | tail -15
tools/sweep_pool.sh "" "-DRTC_TRACE_POOL=0" "-DRTC_POOL_BLOCKS=3" "-DRTC_POOL_K=3 -DRTC_POOL_BLOCKS=3 -DRTC_POOL_STACK=2" "-DRTC_POOL_STACK=2" 2>&1
