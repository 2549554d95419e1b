#!/usr/bin/env python3
"""Quick on-GPU sanity run (development aid): CUDA passes vs the CPU oracle, with timings."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import abi, assets, camera, world  # noqa: E402
from oracle import vxo  # noqa: E402


def cmp(name, a, b):
    a, b = np.asarray(a), np.asarray(b)
    if a.dtype.kind == "f":
        same = (a == b) | (np.isnan(a) & np.isnan(b))
        bad = int((~same).sum())
        err = float(np.nanmax(np.abs(a.astype(np.float64) - b.astype(np.float64)))) if bad else 0.0
        print(f"  {name:14s} mismatches {bad:9d} / {a.size}  max|d| {err:.3e}")
    else:
        bad = int((a != b).sum())
        print(f"  {name:14s} mismatches {bad:9d} / {a.size}")
    return bad


def main():
    W, H = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (640, 360)
    r = vx.Renderer(0)
    w = world.generate_plains(assets.load_plains_columns())
    t0 = time.time(); df_ref = vxo.df_build(w.data); print("oracle df s", time.time() - t0)
    r.upload_world(w)
    for algo in (0, 1):
        r.set_option(abi.OPT_DF_ALGO, algo)
        for _ in range(3):
            r.build_distance_field()
        st = r.stats()
        df = r.download_distance_field()
        print(f"DF algo {algo}: build {st['df_build_ms']*1e3:.1f} us, pack {st['brick_pack_ms']*1e3:.1f} us")
        bad = cmp("df", df, df_ref)
        if bad:
            idx = np.nonzero(df != df_ref)[0][:10]
            for k in idx:
                print("   ", k % 384, (k // 384) % 128, k // 49152, df[k], df_ref[k])
    orc = vxo.Oracle(w.data, df_ref)
    mats = assets.load_materials()
    bn = assets.load_blue_noise()
    sun, moon, stronger, sunvis = camera.sun_moon_direction(50.0)
    sky = assets.analytic_sky(16, sun)
    sn = assets.load_shadow_noise()
    orc.set_tables(mats, bn, sky, sn)
    r.load_scene_tables(mats, bn, sky, sn)
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    pp = vx.primary_params(350, camera.taa_jitter(3))
    t0 = time.time(); g_ref, st_ref = orc.trace_primary(cam, pp); print("oracle primary s", time.time() - t0, st_ref)
    for layout in (0, 1):
        r.set_option(abi.OPT_TRAVERSAL_LAYOUT, layout)
        r.reset_stats()
        g = r.alloc_gbuffer(W, H, hit_voxel=True)
        for _ in range(3):
            r.trace_primary(cam, pp, g)
        r.reset_stats()
        r.trace_primary(cam, pp, g)
        st = r.stats()
        print(f"primary layout {layout}: {st['last_ms']:.3f} ms  {W*H/st['last_ms']/1e3:.1f} Mrays/s stats {st['rays']} {st['df_fetches']} {st['vox_fetches']}")
        for k in ("t", "normal_id", "block_id", "inv_t", "hit_voxel"):
            cmp(k, g[k], g_ref[k])
    sp = vx.shadow_params(stronger, frame=5, soft=True)
    t0 = time.time(); s_ref, sst_ref = orc.trace_shadow(cam, g_ref, sp); print("oracle shadow s", time.time() - t0, sst_ref)
    r.reset_stats()
    s = r.alloc_shadow(W, H)
    r.trace_shadow(cam, g_ref, sp, s)
    r.reset_stats()
    r.trace_shadow(cam, g_ref, sp, s)
    st = r.stats()
    print(f"shadow: {st['last_ms']:.3f} ms stats {st['rays']} {st['df_fetches']} {st['vox_fetches']}")
    cmp("shadow", s["shadow"], s_ref["shadow"]); cmp("transversal", s["transversal"], s_ref["transversal"])
    dp = vx.diffuse_params(sun, moon, sunvis, spp=1, frame=7)
    t0 = time.time(); d_ref, dst_ref = orc.trace_diffuse(cam, g_ref, dp); print("oracle diffuse s", time.time() - t0, dst_ref)
    gdev = {k: __import__("torch").from_numpy(v).cuda() for k, v in g_ref.items()}
    for wf in (0, 1):
        for layout in (0, 1):
            r.set_option(abi.OPT_TRAVERSAL_LAYOUT, layout)
            r.set_option(abi.OPT_GI_WAVEFRONT, wf)
            d = r.alloc_diffuse(W, H, device=True)
            for _ in range(3):
                r.trace_diffuse(cam, gdev, dp, d)
            r.reset_stats()
            import torch
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ext = torch.cuda.ExternalStream(r.cuda_stream())
            e0.record(ext)
            r.trace_diffuse(cam, gdev, dp, d)
            e1.record(ext)
            st = r.stats()
            print(f"diffuse wavefront={wf} layout={layout}: {e0.elapsed_time(e1):.3f} ms stats {st['rays']} {st['df_fetches']} {st['vox_fetches']}")
            for k in ("sh", "cocg", "luma", "ao_sky"):
                cmp(k, d[k].cpu().numpy(), d_ref[k])
    print("L2 sector peak GB/s", r.measure_l2_sector_peak())
    print("launches", r.launch_count())


if __name__ == "__main__":
    main()
