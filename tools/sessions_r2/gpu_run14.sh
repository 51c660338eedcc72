cd $GRAFT_REPO_ROOT
# final knob sweep of the lane-owned driver (after the round's changes), then parity with the defaults
tools/sweep_pool.sh "" "-DRTC_TRI_MINMAX=0" "-DRTC_FETCH_THRESHOLD=12" "-DRTC_FETCH_THRESHOLD=16" "-DRTC_I2F_AXES=0" "-DRTC_I2F_AXES=2" "-DRTC_TRACE_MIN_BLOCKS=7" 2>&1
touch tweeker_raytracer_b200/csrc/kernels_trace.cu tweeker_raytracer_b200/csrc/kernels_shade.cu; make -s -j4 core host
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
