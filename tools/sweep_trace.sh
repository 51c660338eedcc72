#!/bin/bash
# Rebuilds the traversal kernels with different tuning knobs on the GPU box and runs a short bench for each.
# usage: tools/sweep_trace.sh "6 8" "5 8" ...   (pairs: RTC_TRACE_MIN_BLOCKS RTC_FETCH_THRESHOLD)
for cfg in "$@"; do
  set -- $cfg
  touch tweeker_raytracer_b200/csrc/kernels_trace.cu
  make -s core host TRACE_DEFS="-DRTC_TRACE_MIN_BLOCKS=$1 -DRTC_FETCH_THRESHOLD=$2 $3" > /dev/null 2>&1 || { echo "build failed for $cfg"; continue; }
  echo "== blocks $1 threshold $2 $3"
  python bench.py --steps 4 --warmup 3 --spp-per-step 32 --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s", round(d["mrays_per_s"], 1), "Mrays/s; extend", round(r["extend_mrays_per_s"], 1), "connect", round(r["connect"]["mrays_per_s"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()})'
done
