cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_render_parity.py tests/test_gpu_fuzz.py tests/test_gpu_textures.py -x -q -m gpu 2>&1 | tail -3
tools/sweep_pool.sh "" "-DRTC_FETCH_THRESHOLD=12" "-DRTC_FETCH_THRESHOLD=16" "-DRTC_FETCH_THRESHOLD=6" 2>&1
