"""Reads a bench.py JSON line from stdin and prints the headline numbers (dev tool for A/B runs)."""
import json
import sys

line = [l for l in sys.stdin.read().strip().splitlines() if l.startswith("{")][-1]
d = json.loads(line)
print("value %.0f e2e %.0f ms/step %.3f n_gpus %d" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["n_gpus"]))
for k, v in (d.get("extra_configs") or {}).items():
    print("  ", k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if a != "note"})
