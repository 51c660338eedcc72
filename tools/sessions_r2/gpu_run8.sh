cd $GRAFT_REPO_ROOT
make -s -j4 core host
python -m pytest tests/test_gpu_render_parity.py tests/test_gpu_trace_parity.py -x -q -m gpu 2>&1 | tail -3
RTC_TRACE_DRIVER=pool python -m pytest tests/test_gpu_render_parity.py tests/test_gpu_trace_parity.py tests/test_gpu_textures.py -x -q -m gpu 2>&1 | tail -3
echo "== microbench lane"; python tools/ray_microbench.py --tris 1e6,1e7 --rays 3e7 --repeat 2 2>&1 | python -c '
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d["triangles"], d["kind"], d["mode"], round(d["mrays_per_s"]), "Mrays/s", round(d["nodes_per_ray"],1), round(d["tris_per_ray"],1))'
echo "== microbench pool"; RTC_TRACE_DRIVER=pool python tools/ray_microbench.py --tris 1e6,1e7 --rays 3e7 --repeat 2 2>&1 | python -c '
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: continue
    print(d["triangles"], d["kind"], d["mode"], round(d["mrays_per_s"]), "Mrays/s", round(d["nodes_per_ray"],1), round(d["tris_per_ray"],1))'
for drv in lane pool; do
echo "== instances scene $drv"; RTC_TRACE_DRIVER=$drv python bench.py --scene rtigo3_instances --steps 2 --warmup 1 --spp-per-step 16 --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), "connect", round(r["connect"]["mrays_per_s"], 1), {k: round(v, 2) for k, v in r["per_ray"].items()}, {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()})'
echo "== cornell $drv"; RTC_TRACE_DRIVER=$drv python bench.py --scene rtigo3_cornell_box --resolution "512 512" --steps 4 --warmup 2 --spp-per-step 16 --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), "connect", round(r["connect"]["mrays_per_s"], 1), {k: round(v, 2) for k, v in r["per_ray"].items()}, {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()}, "launches", d["gpu_launches"], "ms", d["ms_per_step"])'
done
