#!/bin/bash
# Rebuilds BOTH traversal translation units (kernels_trace.cu and the fused primary kernel in kernels_shade.cu) with
# RTC_I2F_AXES = 0, 1, 2 (how many axes decode their plane bytes through I2F.U8 on the XU pipe instead of PRMT) and benches.
for n in "$@"; do
  sed -i "s/^#define RTC_I2F_AXES [0-9]/#define RTC_I2F_AXES $n/" tweeker_raytracer_b200/csrc/trace.cuh
  make -s core host > /dev/null 2>&1 || { echo "build failed for $n"; continue; }
  echo "== I2F axes $n"
  python bench.py --steps 4 --warmup 3 --spp-per-step 32 --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), "connect", round(r["connect"]["mrays_per_s"], 1), {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()}, round(d["ms_per_step"], 2))'
done
