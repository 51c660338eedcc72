# usage (under gpurun --gpus N): bash tools/sessions_r2/gpu_run_multi_b.sh N -- strong scaling with the WHOLE config as one step (256 spp split
# over the ranks), and one weak line per N with the parity step
cd $GRAFT_REPO_ROOT
N=${1:-2}
run() {
  local n=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $n "$@" 2>gpurun_out/multi_b.err | grep '^{' | tail -1
}
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  echo "== strong, the whole config per step (256 spp split over N=$n)"
  if [ $n -eq 1 ]; then timeout 600 python bench.py --steps 2 --warmup 2 --spp-per-step 256 --scaling strong --no-ncu --no-probes --no-cpu-baseline 2>/dev/null | grep '^{' | tail -1 > gpurun_out/bench_r2_1gpu_strong256.json
  else run $n --steps 2 --warmup 2 --spp-per-step 256 --scaling strong --no-ncu --no-probes > gpurun_out/bench_r2_${n}gpu_strong256.json; fi
  python tools/show_bench.py gpurun_out/bench_r2_${n}gpu_strong256.json
done
echo "== weak N=$N (parity step)"; run $N --steps 8 --warmup 3 --no-ncu --no-probes > gpurun_out/bench_r2_${N}gpu_weak.json; python tools/show_bench.py gpurun_out/bench_r2_${N}gpu_weak.json
echo "== strong N=$N, 32-spp steps (parity step)"; run $N --steps 8 --warmup 3 --scaling strong --no-ncu --no-probes > gpurun_out/bench_r2_${N}gpu_strong.json; python tools/show_bench.py gpurun_out/bench_r2_${N}gpu_strong.json
