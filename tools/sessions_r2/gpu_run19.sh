cd $GRAFT_REPO_ROOT
# final lines after the leaf-size change of both builders
python -m pytest tests -x -q -m gpu 2>&1 | tail -2
python bench.py > gpurun_out/bench_r2_c2.json 2> gpurun_out/bench_r2_c2.err; python tools/show_bench.py gpurun_out/bench_r2_c2.json | cut -c1-330
python bench.py --config c1 > gpurun_out/bench_r2_c1.json 2>/dev/null; python tools/show_bench.py gpurun_out/bench_r2_c1.json | cut -c1-200
for c in c3-1M-coh-closest c3-1M-coh-any c3-1M-incoh-closest c3-1M-incoh-any c3-10M-coh-closest c3-10M-coh-any c3-10M-incoh-closest c3-10M-incoh-any; do
  python bench.py --config $c --steps 3 --warmup 3 --rays 1e8 > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err; python tools/show_bench.py gpurun_out/bench_r2_$c.json | cut -c1-150
done
for c in c3-100M-coh-closest c3-100M-incoh-closest c3-100M-incoh-any; do
  timeout 900 python bench.py --config $c --steps 3 --warmup 3 --rays 1e8 --no-cpu-baseline > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err; python tools/show_bench.py gpurun_out/bench_r2_$c.json | cut -c1-150
done
