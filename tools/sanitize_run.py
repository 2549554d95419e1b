#!/usr/bin/env python3
"""Small workload for compute-sanitizer: one distance-field build with the DPX kernels (shared memory, mbarrier / TMA, cp.async),
one step-field repack, one 480x270 frame of primary + shadow + wavefront GI (shared-memory ray sort, warp-ballot queue), the re-queued
reflection pass (1 and 3 samples) and one frame over two handles through vxpt_mg_*.  Run as
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py
    compute-sanitizer --tool racecheck python tools/sanitize_run.py
The script itself checks nothing; the tool's summary line is the result."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import abi, assets, camera, world  # noqa: E402


def main():
    W, H = 480, 270
    r = vx.Renderer(0)
    r.upload_world(world.generate_plains(assets.load_plains_columns()))
    r.build_distance_field()
    r.set_option(abi.OPT_TRAVERSAL_LAYOUT, 0)   # repack (pack_steps<0>), then back (pack_steps<1>)
    r.set_option(abi.OPT_TRAVERSAL_LAYOUT, 1)
    r.set_block(190, 60, 200, 5)
    r.build_distance_field()
    sun, moon, stronger, vis = camera.sun_moon_direction(50.0)
    r.load_scene_tables(assets.load_materials(), assets.load_blue_noise(), assets.analytic_sky(16, sun), assets.load_shadow_noise())
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    g = r.trace_primary(cam, vx.primary_params(350, camera.taa_jitter(3)), r.alloc_gbuffer(W, H))
    r.trace_shadow(cam, g, vx.shadow_params(stronger, frame=3, soft=True), r.alloc_shadow(W, H))
    for spp in (1, 2):
        r.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=spp, frame=3), r.alloc_diffuse(W, H))
    # reflection pass, re-queued form: sorted rays in shared memory, hit queue, per-pixel sums of a multi-sample pass
    mats = assets.load_materials()
    fc = camera.FpsCamera(pitch_deg=-20.0)
    d1 = r.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=3), r.alloc_diffuse(W, H))
    for spp, checker in ((1, False), (3, True)):
        rp = vx.reflection_params(sun, moon, stronger, fc.position, mats["grass_props"], spp=spp, rough=True, checkerboard=checker, frame=3,
                                  halton=camera.taa_jitter_secondary(3))
        r.trace_reflection(cam, g, d1, rp, r.alloc_reflection(W, H))
    r.sync()
    # one frame over two handles (vxpt_mg_*), host planes, reflection halo rows
    mg = vx.MultiRenderer([0, 0])
    mg.load_scene_tables(mats, assets.load_blue_noise(), assets.analytic_sky(16, sun), assets.load_shadow_noise())
    mg.upload_world(world.generate_plains(assets.load_plains_columns()))
    mg.build_distance_field()
    rp = vx.reflection_params(sun, moon, stronger, fc.position, mats["grass_props"], spp=1, rough=True, frame=3, halton=camera.taa_jitter_secondary(3))
    mg.render_frame(cam, vx.primary_params(350), shadow=vx.shadow_params(stronger, frame=3, soft=True), diffuse=vx.diffuse_params(sun, moon, vis, spp=1, frame=3),
                    gbuf=r.alloc_gbuffer(W, H), shadow_out=r.alloc_shadow(W, H), diffuse_out=r.alloc_diffuse(W, H), reflection=rp, reflection_out=r.alloc_reflection(W, H))
    mg.close()
    r.sync()
    print("sanitize_run: done", r.stats()["rays"])


if __name__ == "__main__":
    main()
