cd $GRAFT_REPO_ROOT
# leaf size of the device LBVH builder on the triangle soups
for l in 3 2 1; do for c in c3-1M-incoh-closest c3-1M-coh-closest c3-10M-incoh-closest c3-10M-incoh-any; do
  echo "== RTC_GPU_LEAF_MAX=$l $c"; RTC_GPU_LEAF_MAX=$l python bench.py --config $c --steps 3 --warmup 2 --rays 3e7 --no-cpu-baseline --no-probes 2>/dev/null | python tools/show_bench.py | cut -c1-120
done; done
