cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_wide_bvh.py tests/test_gpu_render_parity.py -x -q -m gpu 2>&1 | tail -6
nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/cond tools/probe_cond_graph.cu && /tmp/cond
tools/sweep_pool.sh "-DRTC_LEAF_THRESHOLD=6" "-DRTC_LEAF_THRESHOLD=8" "-DRTC_LEAF_THRESHOLD=12" "-DRTC_LEAF_THRESHOLD=8 -DRTC_FETCH_THRESHOLD=12" 2>&1
echo "== instances scene, leaf threshold 8/12"; python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline --no-ncu --no-probes 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), "connect", round(r["connect"]["mrays_per_s"], 1))'
python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_fuzz.py tests/test_gpu_textures.py -x -q -m gpu 2>&1 | tail -2
touch tweeker_raytracer_b200/csrc/kernels_trace.cu tweeker_raytracer_b200/csrc/kernels_shade.cu; make -s -j4 core host
echo "== textures scene"; python bench.py --scene rtigo3_textures --steps 3 --warmup 2 --no-cpu-baseline --no-ncu --no-probes 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; ms/step", round(d["ms_per_step"], 2), "extend", round(r["extend_mrays_per_s"], 1), {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()}, "launches", d["gpu_launches"], "e2e", round(d["e2e"]["value"], 1))'
echo "== c2 per-iteration calling pattern"; python bench.py --steps 4 --warmup 2 --calling-pattern per-iteration --no-cpu-baseline --no-ncu --no-probes 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read())
print(round(d["value"], 1), "Msamples/s; e2e", round(d["e2e"]["value"], 1), d["e2e"]["path"])'
