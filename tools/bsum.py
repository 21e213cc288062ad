"""Compact one-line summary of a bench.py JSON line read from stdin (developer convenience)."""
import json
import sys

tag = " ".join(sys.argv[1:])
lines = [l for l in sys.stdin if l.startswith("{")]
if not lines:
    print(tag, "NO JSON")
    sys.exit(0)
d = json.loads(lines[-1])
r = d.get("roofline") or {}
e = d.get("e2e") or {}
print(tag, "gpus=%s ms/step=%.3f value=%.4g kernel_ms=%s frac=%s e2e_ms=%s launches=%s" % (
    d.get("n_gpus"), d.get("ms_per_step", 0), d.get("value", 0),
    None if r.get("kernel_ms_per_launch") is None else round(r["kernel_ms_per_launch"], 3),
    None if r.get("frac") is None else round(r["frac"], 3),
    None if e.get("ms_per_step") is None else round(e["ms_per_step"], 2), d.get("gpu_launches")))
