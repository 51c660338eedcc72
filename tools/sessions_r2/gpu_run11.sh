cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
echo "== default bench (c2)"; ( time python bench.py > gpurun_out/bench_r2_c2.json 2> gpurun_out/bench_r2_c2.err ) 2>&1 | grep real; tail -c 600 gpurun_out/bench_r2_c2.err | tail -3
echo "== reference arm"; ( time python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r2_c2_reference.json 2> gpurun_out/bench_r2_ref.err ) 2>&1 | grep real
echo "== c1"; ( time python bench.py --config c1 > gpurun_out/bench_r2_c1.json 2> gpurun_out/bench_r2_c1.err ) 2>&1 | grep real
echo "== c4"; ( time python bench.py --config c4 --steps 4 > gpurun_out/bench_r2_c4.json 2> gpurun_out/bench_r2_c4.err ) 2>&1 | grep real
for c in c3-1M-coh-closest c3-1M-incoh-closest c3-1M-incoh-any c3-10M-coh-closest c3-10M-incoh-closest c3-10M-incoh-any; do
echo "== $c"; ( time python bench.py --config $c --steps 3 --warmup 3 --rays 3e7 > gpurun_out/bench_r2_$c.json 2> gpurun_out/bench_r2_$c.err ) 2>&1 | grep real
done
for f in gpurun_out/bench_r2_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get("roofline", {})
    print(sys.argv[1].split("/")[-1], round(d["value"], 1), d["unit"], "e2e", round(d["e2e"]["value"], 1), "fractions", {k: round(v, 3) for k, v in r.get("fractions", {}).items()}, "bound", r.get("bound"), "traffic", r.get("traffic"), "cpu", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in d.get("cpu_baseline", {}).items() if k in ("value", "cores", "kind")}, "clocks", d.get("clocks", {}).get("sm_mhz"), d.get("clocks", {}).get("reasons"))
except Exception as e:
    print(sys.argv[1], "unreadable:", e)
PY
done
