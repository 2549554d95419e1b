#!/usr/bin/env python3
"""Measurement aid: phase timeline of the two distance-field kernels from a trace build of the library
    python voxelpathtracer_b200/build.py -DVXPT_DF_TRACE --out=libvxpt_dftrace.so
    python tools/df_timeline.py            (on the GPU box)
Thread 0 of every CTA logs %globaltimer at its phase boundaries (df_build.cu, DF_TRACE); prints, per kernel, when the phases start and
end relative to the first CTA's start (min / median / max over CTAs, microseconds)."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voxelpathtracer_b200 import abi  # noqa: E402
abi.LIB_PATH = os.path.join(os.path.dirname(abi.LIB_PATH), "libvxpt_dftrace.so")
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import assets, world  # noqa: E402


def main():
    r = vx.Renderer(0)
    r.upload_world(world.generate_plains(assets.load_plains_columns()))
    lib = abi.load()
    buf = np.zeros((2, 2048, 16), dtype=np.uint64)
    for _ in range(3):
        r.build_distance_field()
    r.sync()
    lib.vxpt_debug_df_trace_clear()
    r.build_distance_field()
    r.sync()
    lib.vxpt_debug_df_trace.argtypes = [C.c_void_p]
    lib.vxpt_debug_df_trace(buf.ctypes.data)
    t0 = buf[0][buf[0] > 0].min()
    out = {}
    names = {0: ["start", "slice0 loaded", "slice0 x done", "slice0 y done", "slice0 store issued", "slice1 loaded", "slice1 x done", "slice1 y done",
                 "slice1 store issued", "", "", "", "", "", "", "end"],
             1: ["start", "after griddep wait", "loaded + local sweeps", "carries ready", "end"] + [""] * 11}
    for k in (0, 1):
        for slot, name in enumerate(names[k]):
            v = buf[k][:, slot]
            v = v[v > 0]
            if name and v.size:
                rel = (v.astype(np.int64) - int(t0)) / 1e3
                out[f"k{k} {name}"] = {"ctas": int(v.size), "min_us": round(float(rel.min()), 2), "median_us": round(float(np.median(rel)), 2), "max_us": round(float(rel.max()), 2)}
    for key, v in out.items():
        print(f"{key:32s} {v}")
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "df_timeline.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
