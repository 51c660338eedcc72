#!/bin/bash
# Runs the default bench with different environment settings: tools/sweep_env.sh "RTC_MAX_PATHS=8388608" "RTC_MAX_PATHS=33554432" ...
for cfg in "$@"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 6 --warmup 3 --spp-per-step 16 --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s", round(d["mrays_per_s"], 1), "Mrays/s; extend", round(r["extend_mrays_per_s"], 1), "frac", round(r["frac"], 3), "connect", round(r["connect"]["mrays_per_s"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()})'
done
