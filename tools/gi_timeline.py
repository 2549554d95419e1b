#!/usr/bin/env python3
"""Measurement aid: phase timeline of the two GI kernels from a trace build of the library
    python voxelpathtracer_b200/build.py -DVXPT_GI_TRACE --out=libvxpt_gitrace.so
    python tools/gi_timeline.py            (on the GPU box)
Thread 0 of every CTA logs %globaltimer at its phase boundaries (trace_gi.cu, GI_TRACE); prints, per kernel, when the phases end relative
to the first CTA's start (min / median / p90 / max over the CTAs that reached the phase, microseconds) on the 1080p plains bench frame."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voxelpathtracer_b200 import abi  # noqa: E402
abi.LIB_PATH = os.path.join(os.path.dirname(abi.LIB_PATH), "libvxpt_gitrace.so")
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import assets, camera, world  # noqa: E402


def main():
    W, H = 1920, 1080
    r = vx.Renderer(0)
    r.upload_world(world.generate_plains(assets.load_plains_columns()))
    r.build_distance_field()
    r.set_option(abi.OPT_TEXEL_FORMAT, 1)
    sun, moon, stronger, vis = camera.sun_moon_direction(50.0)
    r.load_scene_tables(assets.load_materials(), assets.load_blue_noise(), assets.analytic_sky(16, sun), assets.load_shadow_noise())
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    g = r.alloc_gbuffer(W, H, device=True, texel=True)
    d = r.alloc_diffuse(W, H, device=True, texel=True)
    lib = abi.load()
    for f in range(4):
        r.trace_primary(cam, vx.primary_params(350, camera.taa_jitter(f)), g)
        if f == 3:
            r.sync()
            lib.vxpt_debug_gi_trace_clear()
        r.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=f), d)
    r.sync()
    print("diffuse pass ms (library events):", r.stats()["last_ms"])
    buf = np.zeros((2, 4096, 16), dtype=np.uint64)
    lib.vxpt_debug_gi_trace.argtypes = [C.c_void_p]
    lib.vxpt_debug_gi_trace(buf.ctypes.data)
    t0 = buf[0][buf[0] > 0].min()
    names = {0: ["start", "rays generated", "sorted", "warp 0 out of groups"],
             1: ["chunk0 start", "chunk0 A shade", "chunk0 sorted", "chunk0 B trace bounce + shadow0", "chunk0 C shade", "chunk0 D trace shadow1", "chunk0 E finish",
                 "chunk1 start", "chunk1 A", "chunk1 sorted", "chunk1 B", "chunk1 C", "chunk1 D", "chunk1 E"]}
    out = {}
    for k in (0, 1):
        for slot, name in enumerate(names[k]):
            v = buf[k][:, slot]
            v = v[v > 0]
            if v.size:
                rel = (v.astype(np.int64) - int(t0)) / 1e3
                out[f"k{k} {name}"] = {"ctas": int(v.size), "min_us": round(float(rel.min()), 1), "median_us": round(float(np.median(rel)), 1),
                                       "p90_us": round(float(np.percentile(rel, 90)), 1), "max_us": round(float(rel.max()), 1)}
    for key, v in out.items():
        print(f"{key:40s} {v}")
    # per-CTA phase durations of gi_continue's first chunk
    c = buf[1].astype(np.int64)
    ok = (c[:, 0] > 0) & (c[:, 6] > 0)
    if ok.any():
        dur = np.diff(c[ok, :7], axis=1) / 1e3
        for j, nm in enumerate(("A shade", "sort", "B trace", "C shade", "D trace", "E finish")):
            print(f"gi_continue chunk0 {nm:10s} duration us: median {np.median(dur[:, j]):6.1f}  p90 {np.percentile(dur[:, j], 90):6.1f}  max {dur[:, j].max():6.1f}")
    g0 = buf[0].astype(np.int64)
    ok = (g0[:, 0] > 0) & (g0[:, 3] > 0)
    dur = np.diff(g0[ok, :4], axis=1) / 1e3
    for j, nm in enumerate(("generate", "sort", "trace")):
        print(f"gi_gen_trace0 {nm:10s} duration us: median {np.median(dur[:, j]):6.1f}  p90 {np.percentile(dur[:, j], 90):6.1f}  max {dur[:, j].max():6.1f}")
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "gi_timeline.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
