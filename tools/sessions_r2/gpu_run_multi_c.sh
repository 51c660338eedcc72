# usage (under gpurun --gpus 8): bash tools/sessions_r2/gpu_run_multi_c.sh -- the 4- and 8-GPU lines of round 2 (every line with its parity step)
cd $GRAFT_REPO_ROOT
run() {
  local n=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 1000)) bench.py --gpus $n "$@" 2>gpurun_out/multi_c.err | grep '^{' | tail -1
}
show() { python tools/show_bench.py $1 | cut -c1-250; }
echo "== weak N=8";        run 8 --steps 8 --warmup 3 --no-ncu --no-probes > gpurun_out/bench_r2_8gpu_weak.json; show gpurun_out/bench_r2_8gpu_weak.json
echo "== weak N=4";        run 4 --steps 8 --warmup 3 --no-ncu --no-probes > gpurun_out/bench_r2_4gpu_weak.json; show gpurun_out/bench_r2_4gpu_weak.json
for n in 4 8; do
echo "== strong256 N=$n";  run $n --steps 2 --warmup 2 --spp-per-step 256 --scaling strong --no-ncu --no-probes > gpurun_out/bench_r2_${n}gpu_strong256.json; show gpurun_out/bench_r2_${n}gpu_strong256.json
echo "== strong32 N=$n";   run $n --steps 8 --warmup 3 --scaling strong --no-ncu --no-probes > gpurun_out/bench_r2_${n}gpu_strong.json; show gpurun_out/bench_r2_${n}gpu_strong.json
done
echo "== c5 weak N=8";     run 8 --config c5 --steps 8 --warmup 3 --no-ncu --no-probes > gpurun_out/bench_r2_8gpu_c5_weak.json; show gpurun_out/bench_r2_8gpu_c5_weak.json
echo "== c5 strong N=8 (64 spp per step split)"; run 8 --config c5 --steps 4 --warmup 2 --spp-per-step 64 --scaling strong --no-ncu --no-probes > gpurun_out/bench_r2_8gpu_c5_strong.json; show gpurun_out/bench_r2_8gpu_c5_strong.json
echo "== c5 tile partition (reference strategy 3 + NCCL composite), one process driving 8 GPUs"
( cd /tmp && timeout 600 $GRAFT_REPO_ROOT/tweeker_raytracer_b200/lib/rtigo3_b200 -s $GRAFT_REPO_ROOT/scenes/system_rtigo3_geometry_4k_tiles.txt -d $GRAFT_REPO_ROOT/scenes/scene_rtigo3_geometry.txt -m 1 2>&1 | grep -E "fps|ERROR" | tail -3 )
