# usage (under gpurun --gpus N): bash tools/sessions_r2/gpu_run_multi.sh N   -- multi-GPU tests, weak + strong scaling, 4K tile partition
cd $GRAFT_REPO_ROOT
N=${1:-2}
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_instances_multigpu.py tests/test_gpu_process_group.py -x -q -m gpu 2>&1 | tail -4
run() {  # run <n> <extra bench flags...>
  local n=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $n "$@" 2>gpurun_out/multi.err | grep '^{' | tail -1
}
for n in 2 4 8; do
  [ $n -le $N ] || continue
  echo "== weak N=$n";   run $n --steps 8 --warmup 3 --no-ncu --no-probes > gpurun_out/bench_r2_${n}gpu_weak.json;   python tools/show_bench.py gpurun_out/bench_r2_${n}gpu_weak.json
  echo "== strong N=$n"; run $n --steps 8 --warmup 3 --scaling strong --no-ncu --no-probes > gpurun_out/bench_r2_${n}gpu_strong.json; python tools/show_bench.py gpurun_out/bench_r2_${n}gpu_strong.json
done
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  echo "== strong, the whole config per step (256 spp split over N=$n)"
  if [ $n -eq 1 ]; then timeout 600 python bench.py --steps 2 --warmup 2 --spp-per-step 256 --scaling strong --no-ncu --no-probes --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 > gpurun_out/bench_r2_1gpu_strong256.json
  else run $n --steps 2 --warmup 2 --spp-per-step 256 --scaling strong --no-ncu --no-probes > gpurun_out/bench_r2_${n}gpu_strong256.json; fi
  python tools/show_bench.py gpurun_out/bench_r2_${n}gpu_strong256.json
done
echo "== c5 (4K) sample-range, N=$N"; run $N --config c5 --steps 8 --warmup 3 --spp-per-step 8 --no-ncu --no-probes > gpurun_out/bench_r2_${N}gpu_c5_weak.json; python tools/show_bench.py gpurun_out/bench_r2_${N}gpu_c5_weak.json
echo "== c5 (4K) strong, 64 spp per step split over N=$N"; run $N --config c5 --steps 4 --warmup 2 --spp-per-step 64 --scaling strong --no-ncu --no-probes > gpurun_out/bench_r2_${N}gpu_c5_strong.json; python tools/show_bench.py gpurun_out/bench_r2_${N}gpu_c5_strong.json
echo "== c5 tile partition (reference strategy 3 + NCCL composite), one process driving $N GPUs"
mask=$(( (1 << N) - 1 ))
sed "s/^devicesMask .*/devicesMask $mask/" scenes/system_rtigo3_geometry_4k_tiles.txt > /tmp/system_4k_tiles.txt
( cd /tmp && timeout 600 $GRAFT_REPO_ROOT/tweeker_raytracer_b200/lib/rtigo3_b200 -s /tmp/system_4k_tiles.txt -d $GRAFT_REPO_ROOT/scenes/scene_rtigo3_geometry.txt -m 1 2>&1 | grep -E "fps|ERROR" | tail -3 )
