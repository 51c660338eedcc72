cd $GRAFT_REPO_ROOT
# ncu evidence of the FINAL build: launch list of one default step, ncu --set full of the traversal kernels
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-ncu --no-probes > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ncu --no-probes > gpurun_out/ncu_launches_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_trace|k_extend_primary' --launch-skip 12 -c 4 -o gpurun_out/prof_lane_r2 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-ncu --no-probes > gpurun_out/ncu_lane_r2.log 2>&1
tail -1 gpurun_out/ncu_lane_r2.log | cut -c1-120
