#!/bin/bash
# ncu --set full of the binning + shading kernels of the first two depths of one default step (what bounds the 21 % that is not traversal)
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ncu --no-probes > /dev/null 2>&1 || exit 1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:'k_shade|k_bin' -c 16 -o gpurun_out/prof_shade_r2 -f \
  python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-ncu --no-probes > gpurun_out/ncu_shade_r2.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/prof_shade_r2.ncu-rep
