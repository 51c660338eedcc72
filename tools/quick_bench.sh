#!/bin/bash
# Short GPU check used while tuning kernels: the trace/render parity tests, then a 4-step bench; prints one summary line.
# usage (on the GPU box, through gpurun): tools/quick_bench.sh [label]
python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_render_parity.py -x -q -m gpu 2>&1 | tail -2
python bench.py --steps 4 --warmup 3 --spp-per-step 32 --no-cpu-baseline 2>/dev/null | tail -1 | python -c '
import sys, json
d = json.loads(sys.stdin.read()); r = d["roofline"]
print(sys.argv[1], round(d["value"], 1), "Msamples/s", round(d["mrays_per_s"], 1), "Mrays/s; extend", round(r["extend_mrays_per_s"], 1), "connect", round(r["connect"]["mrays_per_s"], 1), "e2e", round(d["e2e"]["value"], 1), {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()}, "ms/step", round(d["ms_per_step"], 2))' "${1:-run}"
