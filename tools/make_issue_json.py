#!/usr/bin/env python3
"""profiles/issue.json from an ncu launch list with instruction counts:
    ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum --clock-control none -c 60 --csv \
        --log-file gpurun_out/<run>_inst_gi.csv python tools/gi_probe.py 3
    python tools/make_issue_json.py gpurun_out/<run>_inst_gi.csv
Warp instructions per launch of the three trace passes on the bench frame (1080p plains; medians over the captured frames).  bench.py
divides them by the live pass time and the issue peak (148 SMs x 4 schedulers x 1 warp instruction per cycle at the sampled SM clock) ->
`roofline_all.<pass>.issue_frac`, the second roofline the traversal kernels are actually bound by."""
import csv
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PASS_OF = {"primary_kernel": "primary_kernel", "shadow_kernel": "shadow_kernel", "gi_gen_trace0": "diffuse_pass", "gi_continue": "diffuse_pass",
           "gi_finalize": "diffuse_pass"}


def main():
    rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iid = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per_launch = {}
    for r in rows[1:]:
        per_launch.setdefault((r[iid], r[ik]), {})[r[im]] = float(r[iv].replace(",", ""))
    by_kernel = {}
    for (_, name), m in per_launch.items():
        short = next((k for k in PASS_OF if k in name), None)
        if short and "smsp__inst_executed.sum" in m:
            by_kernel.setdefault(short, []).append((m["smsp__inst_executed.sum"], m.get("smsp__thread_inst_executed.sum", 0.0), m.get("gpu__time_duration.sum", 0.0)))
    out = {"source": os.path.basename(sys.argv[1]), "workload": "tools/gi_probe.py: 1920x1080 plains, the bench frame", "kernels": {}, "passes": {}}
    for k, v in by_kernel.items():
        a = np.array(v)
        out["kernels"][k] = {"launches": len(v), "warp_instructions": float(np.median(a[:, 0])), "thread_instructions": float(np.median(a[:, 1])),
                             "lanes_per_instruction": float(np.median(a[:, 1]) / max(np.median(a[:, 0]), 1.0)), "ncu_ns": float(np.median(a[:, 2]))}
        p = PASS_OF[k]
        out["passes"][p] = out["passes"].get(p, 0.0) + float(np.median(a[:, 0]))
    json.dump(out, open(os.path.join(ROOT, "profiles", "issue.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(out["passes"]), {k: round(v["lanes_per_instruction"], 1) for k, v in out["kernels"].items()})


if __name__ == "__main__":
    main()
