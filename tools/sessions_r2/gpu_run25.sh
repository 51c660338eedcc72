#!/bin/bash
# shading kernels compiled for 4 / 5 CTAs per SM (register caps 64 / 48; variants linked beforehand into lib/variants/): bench each
mkdir -p gpurun_out
L=tweeker_raytracer_b200/lib
for v in base 4_4 4_5 5_5 base; do
  cp $L/variants/librtcore_$v.so $L/librtcore.so
  echo "== shade min blocks $v"
  timeout 100 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ncu --no-probes 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()}, round(d["ms_per_step"], 2))'
done 2>&1 | tee gpurun_out/run25.log
cp $L/variants/librtcore_base.so $L/librtcore.so
