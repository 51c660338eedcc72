cd $GRAFT_REPO_ROOT
# size of the SAH-rebuilt top of the device LBVH (cut of subtrees with <= n / divisor triangles)
for dv in 2048 8192 32768; do for c in c3-1M-incoh-closest c3-10M-incoh-closest c3-10M-coh-closest; do
  echo "== RTC_GPU_CUT_DIVISOR=$dv $c"; RTC_GPU_CUT_DIVISOR=$dv python bench.py --config $c --steps 3 --warmup 2 --rays 3e7 --no-cpu-baseline --no-probes 2>/dev/null | python -c '
import sys, json
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1]); r = d["roofline"]["per_ray"]
print(round(d["value"], 1), "Mrays/s nodes", round(r["nodes"], 1), "tris", round(r["tris"], 1), "build_s", round(d["scene_info"]["gas_build_s"], 3))'
done; done
