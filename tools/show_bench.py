#!/usr/bin/env python
"""One-line summary of a bench.py JSON line (file argument or stdin)."""
import json
import sys

text = open(sys.argv[1]).read() if len(sys.argv) > 1 else sys.stdin.read()
lines = [l for l in text.splitlines() if l.startswith("{")]
if not lines:
    print("no bench line")
    sys.exit(0)
d = json.loads(lines[-1])
r = d.get("roofline", {})
out = [d.get("metric"), round(d["value"], 1), d["unit"], "n_gpus", d.get("n_gpus"), d.get("scaling"), "ms/step", round(d.get("ms_per_step", 0), 2),
       "e2e", round(d["e2e"]["value"], 1)]
if "parity_check" in d:
    out += ["parity", d["parity_check"], "max_abs_err", d["parity_detail"]["max_abs_err"]]
if r.get("fractions"):
    out += ["fractions", {k: round(v, 3) for k, v in r["fractions"].items()}, "bound", r.get("bound")]
if "extend_mrays_per_s" in r:
    out += ["extend", round(r["extend_mrays_per_s"]), "connect", round(r["connect"]["mrays_per_s"]), {k: round(v, 3) for k, v in r["kernel_share_of_step"].items()}]
if "trace_schedule" in d:
    t = d["trace_schedule"]
    out += ["schedule", t.get("schedule")]
    if t.get("measured"):
        out += ["group/one/two ms", [round(min(t["group_ms"]), 2), round(t["one_tri_ms"], 2), round(t["two_tri_ms"], 2)]]
if "clocks" in d:
    out += ["clocks", d["clocks"].get("sm_mhz"), d["clocks"].get("reasons")]
print(*out)
