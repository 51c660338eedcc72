cd $GRAFT_REPO_ROOT
python -m pytest tests/test_gpu_builder.py tests/test_gpu_full_size.py -x -q -m gpu 2>&1 | tail -2
for l in 3 2 1; do
  echo "== RTC_HOST_LEAF_MAX=$l c4"; RTC_HOST_LEAF_MAX=$l python bench.py --config c4 --steps 2 --warmup 1 --no-cpu-baseline --no-ncu --no-probes 2>/dev/null | python tools/show_bench.py | cut -c1-100
  echo "== RTC_HOST_LEAF_MAX=$l c2"; RTC_HOST_LEAF_MAX=$l python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-ncu --no-probes 2>/dev/null | python tools/show_bench.py | cut -c1-100
  echo "== RTC_HOST_LEAF_MAX=$l c1"; RTC_HOST_LEAF_MAX=$l python bench.py --config c1 --steps 4 --warmup 2 --no-cpu-baseline --no-ncu --no-probes 2>/dev/null | python tools/show_bench.py | cut -c1-100
done
