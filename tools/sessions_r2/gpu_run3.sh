set -x
cd $GRAFT_REPO_ROOT
touch tweeker_raytracer_b200/csrc/kernels_trace.cu tweeker_raytracer_b200/csrc/kernels_shade.cu
make -s -j4 core host TRACE_DEFS="-DRTC_TRACE_POOL=0"
for cfg in "3 1.2" "3 0.6" "3 0.3" "2 1.2" "2 0.3" "1 1.2"; do
  set -- $cfg
  echo "== old kernel, leafMax $1 splitCost $2"
  RTC_LEAF_MAX=$1 RTC_LEAF_SPLIT_COST=$2 python bench.py --steps 4 --warmup 3 --spp-per-step 32 --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), "connect", round(r["connect"]["mrays_per_s"], 1), "per ray", {k: round(v, 3) for k, v in r["per_ray"].items()}, "nodes", d["config"]["bvh_nodes"], "ms/step", round(d["ms_per_step"], 2))'
done
touch tweeker_raytracer_b200/csrc/kernels_trace.cu tweeker_raytracer_b200/csrc/kernels_shade.cu
make -s -j4 core host TRACE_DEFS="-DRTC_POOL_STACK=2 -DRTC_POOL_BLOCKS=5"
for cfg in "3 1.2" "2 0.3" "1 1.2"; do
  set -- $cfg
  echo "== pool kernel, leafMax $1 splitCost $2"
  RTC_LEAF_MAX=$1 RTC_LEAF_SPLIT_COST=$2 python bench.py --steps 4 --warmup 3 --spp-per-step 32 --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), "connect", round(r["connect"]["mrays_per_s"], 1), "per ray", {k: round(v, 3) for k, v in r["per_ray"].items()}, "nodes", d["config"]["bvh_nodes"], "ms/step", round(d["ms_per_step"], 2))'
done
