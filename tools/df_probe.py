#!/usr/bin/env python3
"""Measurement aid: the distance-field build on the plains world, the DPX build whose z sweep writes the step field too (VXPT_OPT_DF_ALGO = 1,
default) against the reference-shaped build + pack_steps (= 0), timed by the library's CUDA events (VxStats.df_build_ms + brick_pack_ms).
Prints one JSON line; algorithmic bytes per build + pack = read grid, write DF, write step field = 3 x 18,874,368 B."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import abi as _abi  # noqa: E402
if os.environ.get("VXPT_LIB"):  # development: an experiment build of the library (build.py --out=...), never the product path
    _abi.LIB_PATH = os.path.join(os.path.dirname(_abi.LIB_PATH), os.environ["VXPT_LIB"])
from voxelpathtracer_b200 import abi, assets, world  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    r = vx.Renderer(0)
    r.upload_world({'plains': lambda: world.generate_plains(assets.load_plains_columns()), 'gi_box': lambda: world.generate_gi_box(assets.load_plains_columns()), 'city': world.generate_city, 'superflat': world.generate_superflat}[os.environ.get('VXPT_PROBE_WORLD', 'plains')]())
    out = {"kernel": "df_xy_dpx + df_z_dpx (+ pack_steps)", "iters": iters, "algorithmic_bytes": 3 * abi.WORLD_VOXELS}
    peak = 6451.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    for algo in (1, 0):
        r.set_option(abi.OPT_DF_ALGO, algo)
        df, pack = [], []
        for k in range(iters + 3):
            r.build_distance_field()
            r.sync()
            st = r.stats()
            if k >= 3:
                df.append(st["df_build_ms"])
                pack.append(st["brick_pack_ms"])
        total = float(np.mean(df) + np.mean(pack))
        out["algo%d" % algo] = {"df_build_ms": float(np.mean(df)), "brick_pack_ms": float(np.mean(pack)), "total_ms": total,
                                "achieved_gbs": 3 * abi.WORLD_VOXELS / (total * 1e-3) / 1e9, "frac": 3 * abi.WORLD_VOXELS / (total * 1e-3) / 1e9 / peak}
    out["hbm_peak_gbs"] = peak
    print(json.dumps(out))


if __name__ == "__main__":
    main()
