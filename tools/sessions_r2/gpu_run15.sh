cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
python -c "import __graft_entry__ as e; e.smoke()" 2>&1 | tail -1
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
RTC_TRACE_DRIVER=pool python -m pytest tests/test_gpu_trace_parity.py tests/test_gpu_render_parity.py tests/test_gpu_textures.py tests/test_gpu_fuzz.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -1
RTC_PRIMARY_PACKETS=1 python -m pytest tests/test_gpu_render_parity.py tests/test_gpu_full_size.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -1
echo "== default bench (c2)"; ( time python bench.py > gpurun_out/bench_r2_c2.json 2> gpurun_out/bench_r2_c2.err ) 2>&1 | grep real
echo "== reference arm"; ( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_c2_reference.json 2> gpurun_out/bench_r2_ref.err ) 2>&1 | grep real
echo "== c2 per-iteration calling pattern"; python bench.py --calling-pattern per-iteration --no-cpu-baseline > gpurun_out/bench_r2_c2_per_iteration.json 2>/dev/null
echo "== c1"; python bench.py --config c1 > gpurun_out/bench_r2_c1.json 2> gpurun_out/bench_r2_c1.err
echo "== c4"; python bench.py --config c4 --steps 4 > gpurun_out/bench_r2_c4.json 2> gpurun_out/bench_r2_c4.err
echo "== textures"; python bench.py --scene rtigo3_textures --steps 4 --no-cpu-baseline > gpurun_out/bench_r2_textures.json 2>/dev/null
for c in c3-1M-coh-closest c3-1M-coh-any c3-1M-incoh-closest c3-1M-incoh-any c3-10M-coh-closest c3-10M-coh-any c3-10M-incoh-closest c3-10M-incoh-any; do
  python bench.py --config $c --steps 3 --warmup 3 --rays 1e8 > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err
done
for c in c3-100M-coh-closest c3-100M-incoh-closest c3-100M-incoh-any; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 --rays 1e8 --no-cpu-baseline > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err
done
for f in gpurun_out/bench_r2_c*.json gpurun_out/bench_r2_textures.json; do python tools/show_bench.py $f | cut -c1-400; done
