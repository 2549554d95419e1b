#!/usr/bin/env python3
"""Measurement aid: the trace passes at 1080p on the plains world, planes resident in device memory, each pass timed by the library's CUDA
events (VxStats.last_ms), median over `iters` frames after 3 warm-up frames; three scene replicas are rotated so a frame does not find the
previous frame's lines in L2.  Prints one JSON line (the experiment knobs VXPT_GI_SORT / VXPT_GI_CTAS / ... of the environment are echoed).

  python tools/gi_probe.py [iters [width height]]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import abi as _abi  # noqa: E402
if os.environ.get("VXPT_LIB"):  # development: an experiment build of the library (build.py --out=...), never the product path
    _abi.LIB_PATH = os.path.join(os.path.dirname(_abi.LIB_PATH), os.environ["VXPT_LIB"])
from voxelpathtracer_b200 import abi, assets, camera, world  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
    r = vx.Renderer(0)
    r.upload_world(world.generate_plains(assets.load_plains_columns()))
    r.build_distance_field()
    r.set_option(abi.OPT_SCENE_REPLICAS, 3)
    r.set_option(abi.OPT_TEXEL_FORMAT, 1)
    sun, moon, stronger, sunvis = camera.sun_moon_direction(50.0)
    r.load_scene_tables(assets.load_materials(), assets.load_blue_noise(), assets.analytic_sky(16, sun), assets.load_shadow_noise())
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    g = r.alloc_gbuffer(W, H, device=True, texel=True)
    s = r.alloc_shadow(W, H, device=True, texel=True)
    d = r.alloc_diffuse(W, H, device=True, texel=True)
    ms = {"primary": [], "shadow": [], "diffuse": []}
    fetches = {}
    for f in range(iters + 3):
        for name, call in (("primary", lambda: r.trace_primary(cam, vx.primary_params(350, camera.taa_jitter(f)), g)),
                           ("shadow", lambda: r.trace_shadow(cam, g, vx.shadow_params(stronger, frame=f, soft=True), s)),
                           ("diffuse", lambda: r.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, sunvis, spp=1, frame=f), d))):
            r.reset_stats()
            call()
            st = r.stats()
            if f >= 3:
                ms[name].append(st["last_ms"])
                fetches[name] = (st["rays"], st["df_fetches"] + st["vox_fetches"])
    l2 = r.measure_l2_sector_peak()
    out = {"resolution": [W, H], "iters": iters, "l2_sector_peak_gbs": l2,
           "env": {k: v for k, v in os.environ.items() if k.startswith("VXPT_")}}
    for name in ms:
        m = float(np.median(ms[name]))
        out[name] = {"ms": m, "ms_min": float(np.min(ms[name])), "rays": fetches[name][0], "fetches": fetches[name][1],
                     "frac_l2": fetches[name][1] * 32.0 / (m * 1e-3) / 1e9 / l2}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
