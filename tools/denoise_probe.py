#!/usr/bin/env python3
"""Measurement aid for the passes that follow the trace passes at 1920x1080 on the plains world, device planes:
G-buffer material pass, SVGF temporal / variance / five a-trous passes, shadow temporal + spatial filter.  Each pass is timed by the
library's CUDA events around its launch (VxStats.last_ms), averaged over `iters` frames after 3 warm-up frames; prints one JSON line with
ms per pass and algorithmic GB/s (each input plane read once, each output plane written once) against the measured HBM peak.
Use under ncu for the launch list:  ncu --metrics gpu__time_duration.sum --clock-control none python tools/denoise_probe.py 3"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import assets, camera, denoise, world  # noqa: E402

# algorithmic bytes per pixel: inputs read once + outputs written once (fp32 planes; ids 1 B)
BYTES = {"svgf_initial": 4 + 1 + (16 + 8 + 4 + 8) * 2,
         "svgf_temporal": (4 + 1 + 1) * 2 + 16 + 8 + 4 + 8 + 16 + 8 + 12 + 8 + 16 + 8 + 12 + 8,
         "svgf_variance": 4 + 1 + 16 + 8 + 12 + 16 + 8 + 4,
         "svgf_spatial": 4 + 1 + 16 + 8 + 4 + 8 + 12 + 16 + 8 + 4 + 8,
         "shadow_temporal": 4 + 1 + 4 + 1 + 4 + 4 + 4 + 4 + 4,
         "shadow_filter": 4 + 1 + 4 + 4 + 4 + 4}


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    W, H = 1920, 1080
    r = vx.Renderer(0)
    r.upload_world(world.generate_plains(assets.load_plains_columns()))
    r.build_distance_field()
    mats = assets.load_materials()
    sun, moon, stronger, vis = camera.sun_moon_direction(50.0)
    r.load_scene_tables(mats, assets.load_blue_noise(), assets.analytic_sky(16, sun), assets.load_shadow_noise())
    ms = {k: [] for k in BYTES}
    prev_g = prev_fc = None
    prev_t = r.alloc_denoise(W, H, ("sh", "cocg", "utility", "ao_sky"), device=True)
    prev_s = r.alloc_denoise(W, H, ("shadow", "frames"), device=True)
    for v in list(prev_t.values()) + list(prev_s.values()):
        v.zero_()
    import torch
    torch.cuda.synchronize()
    pong = [r.alloc_denoise(W, H, ("sh", "cocg", "variance", "ao_sky"), device=True) for _ in range(2)]
    for f in range(iters + 3):
        fc = camera.FpsCamera(position=(192.0 + 0.05 * f, 75.0, 192.0 + 0.03 * f), pitch_deg=-20.0, yaw_deg=90.0 + 0.2 * f)
        cam = fc.vx_camera(W, H)
        g = r.trace_primary(cam, vx.primary_params(350, camera.taa_jitter(f)), r.alloc_gbuffer(W, H, device=True))
        s = r.trace_shadow(cam, g, vx.shadow_params(stronger, frame=f, soft=True), r.alloc_shadow(W, H, device=True))
        d = r.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=f), r.alloc_diffuse(W, H, device=True))
        pfc = prev_fc or fc
        view, proj = pfc.view().T.reshape(16), pfc.projection().T.reshape(16)
        rec = f >= 3

        def timed(name, fn):
            out = fn()
            if rec:
                ms[name].append(r.stats()["last_ms"])
            return out

        pre = timed("svgf_initial", lambda: r.svgf_initial(cam, g, d, r.alloc_denoise(W, H, ("sh", "cocg", "luma", "ao_sky"), device=True)))
        t = timed("svgf_temporal", lambda: r.svgf_temporal(cam, g, prev_g or g, pre, prev_t, denoise.temporal_params(view, proj),
                                                            r.alloc_denoise(W, H, ("sh", "cocg", "utility", "ao_sky"), device=True)))
        v = timed("svgf_variance", lambda: r.svgf_variance(cam, g, t, denoise.variance_params(), r.alloc_denoise(W, H, ("sh", "cocg", "variance"), device=True)))
        cur = {"sh": v["sh"], "cocg": v["cocg"], "variance": v["variance"], "ao_sky": t["ao_sky"]}
        for n, step in enumerate(denoise.ATROUS_STEPS):
            cur = timed("svgf_spatial", lambda: r.svgf_spatial(cam, g, cur, t["utility"], denoise.spatial_params(step, time=1.0 + f / 60.0), pong[n % 2]))
        st = timed("shadow_temporal", lambda: r.shadow_temporal(cam, g, prev_g or g, s, prev_s, denoise.shadow_temporal_params(view, proj),
                                                               r.alloc_denoise(W, H, ("shadow", "frames"), device=True)))
        timed("shadow_filter", lambda: r.shadow_filter(cam, g, st, s["transversal"], denoise.shadow_filter_params(1.0), r.alloc((H, W), np.float32, device=True)))
        prev_g, prev_t, prev_s, prev_fc = g, t, st, fc
    peak = 6451.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {"resolution": [W, H], "iters": iters, "hbm_peak_gbs": peak, "passes": {}}
    for k, v in ms.items():
        m = float(np.mean(v))
        out["passes"][k] = {"ms": m, "launches_per_frame": len(v) // iters, "algorithmic_bytes": W * H * BYTES[k],
                            "achieved_gbs": W * H * BYTES[k] / (m * 1e-3) / 1e9, "frac": W * H * BYTES[k] / (m * 1e-3) / 1e9 / peak}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
