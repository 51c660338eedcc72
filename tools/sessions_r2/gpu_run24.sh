#!/bin/bash
# the device arithmetic probe against gcc's build of the same header, then the whole GPU suite
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_edge_cases.py -x -q -m gpu -k arithmetic > gpurun_out/run24_math.log 2>&1
tail -15 gpurun_out/run24_math.log
timeout 240 python -m pytest tests -x -q -m gpu > gpurun_out/run24_pytest.log 2>&1
tail -3 gpurun_out/run24_pytest.log
