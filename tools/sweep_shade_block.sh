#!/bin/bash
# Rebuilds the shading kernels with a different CTA size / grid multiplier on the GPU box and benches each.
# usage: tools/sweep_shade_block.sh "128 16" "256 8" "512 4"      (pairs: threads per CTA, CTAs per SM in the grid)
for cfg in "$@"; do
  set -- $cfg
  sed -i "s/^constexpr int kBlock = [0-9]*;/constexpr int kBlock = $1;/; s/const int gridShade = ctx->numSMs \* [0-9]*;/const int gridShade = ctx->numSMs * $2;/" tweeker_raytracer_b200/csrc/kernels_shade.cu
  make -s core host > /dev/null 2>&1 || { echo "build failed for $cfg"; continue; }
  echo "== shade block $1 grid x$2"
  python bench.py --steps 4 --warmup 3 --spp-per-step 32 --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), "connect", round(r["connect"]["mrays_per_s"], 1), {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()}, round(d["ms_per_step"], 2))'
done
