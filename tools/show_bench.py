#!/usr/bin/env python
"""Print the headline numbers and the per-kernel table of a bench.py JSON line (file argument or stdin)."""
import json
import signal
import sys

signal.signal(signal.SIGPIPE, signal.SIG_DFL)   # `| head` closes the pipe early

src = open(sys.argv[1]) if len(sys.argv) > 1 else sys.stdin
for line in src:
    line = line.strip()
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    e = d.get("e2e") or {}
    print("value", d.get("value"), d.get("unit"), "ms/step", d.get("ms_per_step"), "| e2e", e.get("value"), e.get("ms_per_step"),
          "| e2e float32", (e.get("float32_input") or {}).get("value"), "| launches", d.get("gpu_launches"))
    print("config", json.dumps(d.get("config", {}).get("band_decimation")), "clocks", d.get("clocks"))
    r = d.get("roofline") or {}
    print("roofline", r.get("stage"), r.get("achieved"), r.get("frac"), "traffic", r.get("traffic"))
    for k, v in (d.get("roofline_by_kernel") or {}).items():
        print(f"  {k:16s} {v['ms']:9.3f} ms  share {v['share_of_step']:.3f}  frac {v['frac']:.3f}")
    if d.get("cpu_baseline"):
        print("cpu", d["cpu_baseline"])
