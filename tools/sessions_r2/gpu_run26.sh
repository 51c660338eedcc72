#!/bin/bash
# final build: the whole GPU suite under the two alternative traversal drivers (results must stay bit-identical)
mkdir -p gpurun_out
RTC_TRACE_DRIVER=pool timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/run26_pool.log 2>&1; tail -2 gpurun_out/run26_pool.log
RTC_PRIMARY_PACKETS=1 timeout 150 python -m pytest tests -x -q -m gpu > gpurun_out/run26_packets.log 2>&1; tail -2 gpurun_out/run26_packets.log
