#!/bin/bash
# final-state check: the new config-4 full-count parity test, then the whole GPU suite, then the default bench line
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_full_size.py -x -q -m gpu -k config4 > gpurun_out/run22_c4test.log 2>&1
tail -5 gpurun_out/run22_c4test.log
timeout 240 python -m pytest tests -x -q -m gpu > gpurun_out/run22_pytest.log 2>&1
tail -3 gpurun_out/run22_pytest.log
timeout 120 python bench.py > gpurun_out/run22_bench_c2.json 2> gpurun_out/run22_bench_c2.err
cut -c1-400 gpurun_out/run22_bench_c2.json
