cd $GRAFT_REPO_ROOT
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for b in tight box; do
echo "== instances scene bounds=$b"; RTC_INSTANCE_BOUNDS=$b python bench.py --scene rtigo3_instances --steps 2 --warmup 1 --spp-per-step 16 --no-cpu-baseline --no-ncu --no-probes 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), "connect", round(r["connect"]["mrays_per_s"], 1), {k: round(v, 2) for k, v in r["per_ray"].items()}, {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()})'
done
tools/sweep_pool.sh "" 2>&1
echo "== cornell"; python bench.py --scene rtigo3_cornell_box --resolution "512 512" --steps 4 --warmup 2 --spp-per-step 16 --no-cpu-baseline --no-ncu --no-probes 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), "connect", round(r["connect"]["mrays_per_s"], 1), {k: round(v, 2) for k, v in r["per_ray"].items()}, "ms", d["ms_per_step"])'
