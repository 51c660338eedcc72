cd $GRAFT_REPO_ROOT
export RTC_PRIMARY_PACKETS=1
timeout 600 python -m pytest tests/test_gpu_render_parity.py tests/test_gpu_full_size.py tests/test_gpu_fuzz.py tests/test_gpu_edge_cases.py tests/test_gpu_mesh_import.py tests/test_gpu_instances_multigpu.py -x -q -m gpu 2>&1 | tail -4
for p in 1 0; do
echo "== packets=$p"
RTC_PRIMARY_PACKETS=$p timeout 300 tools/sweep_pool.sh "" 2>&1 | tail -1
echo "== c1 packets=$p"; RTC_PRIMARY_PACKETS=$p timeout 300 python bench.py --config c1 --steps 4 --warmup 2 --no-cpu-baseline --no-ncu --no-probes 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()})'
echo "== c4 packets=$p"; RTC_PRIMARY_PACKETS=$p timeout 600 python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline --no-ncu --no-probes 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()})'
done
for b in 3 5 6; do
echo "== packets, RTC_PACKET_BLOCKS=$b"; RTC_PRIMARY_PACKETS=1 timeout 300 tools/sweep_pool.sh "-DRTC_PACKET_BLOCKS=$b" 2>&1 | tail -1
done
unset RTC_PRIMARY_PACKETS
echo "== cutout graph"
RTC_CUTOUT_GRAPH=1 timeout 300 python -m pytest tests/test_gpu_textures.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -3
for g in 0 1; do
echo "== textures scene, RTC_CUTOUT_GRAPH=$g"; RTC_CUTOUT_GRAPH=$g timeout 300 python bench.py --scene rtigo3_textures --steps 3 --warmup 2 --no-cpu-baseline --no-ncu --no-probes 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; ms/step", round(d["ms_per_step"], 2), {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()}, "e2e", round(d["e2e"]["value"], 1))'
done
