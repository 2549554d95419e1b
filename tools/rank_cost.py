#!/usr/bin/env python3
"""Development aid: cost of ONE rank's share of a 1080p frame on one GPU, for several sharding shapes."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import abi, assets, camera, world
import bench

W, H = 1920, 1080
tables = bench.load_tables()
r = vx.Renderer(0)
r.load_scene_tables(tables["materials"], tables["blue_noise"], tables["sky"], tables["shadow_noise"])
r.upload_world(world.generate_plains(assets.load_plains_columns()))
r.build_distance_field()
r.set_option(abi.OPT_SCENE_REPLICAS, 3)
fc = camera.FpsCamera(pitch_deg=-20.0)
ext = torch.cuda.ExternalStream(r.cuda_stream())


def cost(n, rank, band, frames=20):
    rows = H // n
    cam = fc.vx_camera(W, H, 0, rows, n if n > 1 else 0, rank, band)
    g = r.alloc_gbuffer(W, rows, device=True); s = r.alloc_shadow(W, rows, device=True); d = r.alloc_diffuse(W, rows, device=True)
    t = np.zeros(4)
    for f in range(frames + 3):
        pp, sp, dp = bench.frame_params(vx, camera, tables, f)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        e[0].record(ext); r.trace_primary(cam, pp, g); e[1].record(ext); r.trace_shadow(cam, g, sp, s); e[2].record(ext); r.trace_diffuse(cam, g, dp, d); e[3].record(ext)
        r.sync()
        if f >= 3:
            t[:3] += [e[i].elapsed_time(e[i + 1]) for i in range(3)]
            t[3] += e[0].elapsed_time(e[3])
    return t / frames


print("n rank band  primary shadow diffuse total(ms)")
print(1, 0, 0, cost(1, 0, 0).round(4))
for n, bands in ((2, (6, 4, 12, 540)), (4, (6, 2, 270)), (8, (5, 1, 3, 9, 15, 27, 135))):
    for band in bands:
        ts = np.array([cost(n, rk, band, 10) for rk in (range(n) if band * n >= H // 1 or band == H // n else (0, n // 2))])
        print(n, "max-over-ranks", band, ts.max(0).round(4), " mean", ts.mean(0).round(4))
