#!/usr/bin/env python3
"""Development aid: time vxpt_render_frame with pinned host planes (texel formats) on rows [0, R)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import abi, assets, camera, world  # noqa: E402


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 1080
    W, H = 1920, 1080
    r = vx.Renderer(0)
    r.upload_world(world.generate_plains(assets.load_plains_columns()))
    r.build_distance_field()
    sun, moon, stronger, vis = camera.sun_moon_direction(50.0)
    r.load_scene_tables(assets.load_materials(), assets.load_blue_noise(), assets.analytic_sky(16, sun), assets.load_shadow_noise())
    r.set_option(abi.OPT_TEXEL_FORMAT, 1)
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H, 0, R)
    g, s, d = (r.alloc_gbuffer(W, H, texel=True, pinned=True), r.alloc_shadow(W, H, texel=True, pinned=True), r.alloc_diffuse(W, H, texel=True, pinned=True))
    ps = [(vx.primary_params(350, camera.taa_jitter(k)), vx.shadow_params(stronger, frame=k), vx.diffuse_params(sun, moon, vis, spp=1, frame=k)) for k in range(64)]
    for k in range(10):
        r.render_frame(cam, *ps[k], g, s, d)
    t0 = time.perf_counter()
    n = 200
    for k in range(n):
        r.render_frame(cam, *ps[k % 64], g, s, d)
    dt = (time.perf_counter() - t0) / n
    print(f"rows {R} mode {os.environ.get('VXPT_FRAME_MODE', '0')}: {dt * 1e3:.3f} ms/frame, {R * W * 27 / dt / 1e9:.1f} GB/s of planes")


if __name__ == "__main__":
    main()
