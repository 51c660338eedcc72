#!/bin/bash
# tools/sweep_spp.sh "<spp-per-step> <RTC_MAX_PATHS>" ...
for cfg in "$@"; do
  set -- $cfg
  echo "== spp/step $1 max paths $2"
  RTC_MAX_PATHS=$2 python bench.py --steps 4 --warmup 3 --spp-per-step $1 --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s", round(d["mrays_per_s"], 1), "Mrays/s; extend", round(r["extend_mrays_per_s"], 1), "frac", round(r["frac"], 3), "connect", round(r["connect"]["mrays_per_s"], 1), "e2e", round(d["e2e"]["value"], 1))'
done
