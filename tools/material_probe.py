#!/usr/bin/env python3
"""Measurement aid: the G-buffer material pass (vxpt_generate_gbuffer) at 1920x1080 on the plains world, device planes, three rotating
output plane sets (3 x 91 MB > L2, so no launch finds its lines cached), timed by the library's CUDA events around each launch.
Prints one JSON line: ms per launch, algorithmic GB/s (6 B read + 44 B written per pixel) against the measured HBM peak."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import assets, camera, world  # noqa: E402
import material_cases as mc  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    W, H = 1920, 1080
    r = vx.Renderer(0)
    r.upload_world(world.generate_plains(assets.load_plains_columns()))
    r.build_distance_field()
    mats = assets.load_materials()
    sun, _, _, _ = camera.sun_moon_direction(50.0)
    r.load_scene_tables(mats, assets.load_blue_noise(), assets.analytic_sky(16, sun), assets.load_shadow_noise())
    r.set_gbuffer_textures(*mc.material_mips(mats["albedo_lod3"].shape[0]))
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    g = r.trace_primary(cam, vx.primary_params(350), r.alloc_gbuffer(W, H, device=True))
    sets = [r.alloc_material(W, H, device=True) for _ in range(3)]
    mp = vx.material_params(mats["grass_props"])
    def timed():
        ms = []
        for k in range(iters + 3):
            r.generate_gbuffer(cam, g, mp, sets[k % 3])
            st = r.stats()
            if k >= 3:
                ms.append(st["last_ms"])
        return np.array(ms)
    ms = timed()
    r.set_option(vx.abi.OPT_MATERIAL_QUAD_SHUFFLE, 1)   # opt-in variant: quad partners by warp shuffle (A/B, same planes)
    ms_shfl = timed()
    r.set_option(vx.abi.OPT_MATERIAL_QUAD_SHUFFLE, 0)
    peak = 6451.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    bytes_per_launch = W * H * 50
    out = {"kernel": "gbuffer_kernel", "resolution": [W, H], "iters": iters, "ms_mean": float(ms.mean()), "ms_min": float(ms.min()),
           "algorithmic_bytes": bytes_per_launch, "achieved_gbs": bytes_per_launch / (ms.mean() * 1e-3) / 1e9, "hbm_peak_gbs": peak,
           "frac": bytes_per_launch / (ms.mean() * 1e-3) / 1e9 / peak, "hit_fraction": float((g["t"] > 0).float().mean()),
           "quad_shuffle_ms_mean": float(ms_shfl.mean()), "quad_shuffle_ms_min": float(ms_shfl.min())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
