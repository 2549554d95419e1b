#!/usr/bin/env python3
"""Print the headline fields of bench.py JSON lines read from stdin (development aid)."""
import json
import sys
for line in sys.stdin:
    if not line.startswith("{"):
        continue
    d = json.loads(line)
    c = d.get("config", {})
    print(sys.argv[1] if len(sys.argv) > 1 else "", "N=%d value=%.0f Mrays/s ms/step=%.4f host_submit=%.4f launches=%s e2e=%.0f" % (
        d["n_gpus"], d["value"], d["ms_per_step"], c.get("host_submit_ms_per_step", -1), d.get("gpu_launches"), (d.get("e2e") or {}).get("value", -1)),
        d.get("pass_ms", {}).get("primary"), d.get("pass_ms", {}).get("shadow"), d.get("pass_ms", {}).get("diffuse"))
