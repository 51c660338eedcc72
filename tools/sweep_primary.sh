#!/bin/bash
for b in 6 7 8; do
  touch tweeker_raytracer_b200/csrc/kernels_shade.cu
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Iinclude -Itweeker_raytracer_b200/csrc -fmad=false -prec-div=true -prec-sqrt=true -DRTC_TRACE_MIN_BLOCKS=$b -c tweeker_raytracer_b200/csrc/kernels_shade.cu -o tweeker_raytracer_b200/lib/kernels_shade.o && make -s core host > /dev/null 2>&1
  echo "== primary kernel blocks $b"
  python bench.py --steps 4 --warmup 3 --spp-per-step 32 --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(round(d["value"], 1), "Msamples/s; extend", round(r["extend_mrays_per_s"], 1), "connect", round(r["connect"]["mrays_per_s"], 1), {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()})'
done
