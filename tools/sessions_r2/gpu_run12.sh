cd $GRAFT_REPO_ROOT
for tool in memcheck racecheck; do
  for drv in lane pool; do
    echo "== compute-sanitizer --tool $tool, RTC_TRACE_DRIVER=$drv"
    RTC_TRACE_DRIVER=$drv timeout 900 compute-sanitizer --tool $tool --print-limit 5 python tools/sanitize_smoke.py > gpurun_out/sanitizer_${tool}_${drv}_r2.log 2>&1
    tail -6 gpurun_out/sanitizer_${tool}_${drv}_r2.log
  done
done
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ncu --no-probes > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ncu --no-probes > gpurun_out/ncu_launches_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_extend_primary' --launch-skip 12 -c 4 -o gpurun_out/prof_lane_r2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ncu --no-probes > gpurun_out/ncu_lane_r2.log 2>&1
tail -2 gpurun_out/ncu_lane_r2.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_extend_primary' --launch-skip 12 -c 3 -o gpurun_out/prof_c4_r2 python bench.py --config c4 --steps 1 --warmup 1 --no-cpu-baseline --no-ncu --no-probes > gpurun_out/ncu_c4_r2.log 2>&1
tail -2 gpurun_out/ncu_c4_r2.log | cut -c1-200
