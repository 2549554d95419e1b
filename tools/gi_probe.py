#!/usr/bin/env python3
"""Development aid: run the diffuse-GI pass at 1080p in the given wavefront modes (for ncu launch lists)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import abi, assets, camera, world  # noqa: E402


def main():
    modes = [int(m) for m in sys.argv[1:]] or [3, 2]
    W, H = 1920, 1080
    r = vx.Renderer(0)
    w = world.generate_plains(assets.load_plains_columns())
    r.upload_world(w)
    r.build_distance_field()
    sun, moon, stronger, sunvis = camera.sun_moon_direction(50.0)
    r.load_scene_tables(assets.load_materials(), assets.load_blue_noise(), assets.analytic_sky(16, sun), assets.load_shadow_noise())
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    g = r.alloc_gbuffer(W, H, device=True)
    r.trace_primary(cam, vx.primary_params(350, camera.taa_jitter(3)), g)
    d = r.alloc_diffuse(W, H, device=True)
    for m in modes:
        r.set_option(abi.OPT_GI_WAVEFRONT, m)
        for f in range(3):
            r.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, sunvis, spp=1, frame=7 + f), d)
            r.reset_stats()
            st = r.stats()
        print("mode", m, "last_ms", st["last_ms"])


if __name__ == "__main__":
    main()
