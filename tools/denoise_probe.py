#!/usr/bin/env python3
"""Measurement aid for the passes that follow the trace passes on the plains world, planes resident in device
memory: G-buffer material pass, SVGF pre-pass / temporal / variance / five a-trous passes, shadow temporal + spatial filter, and the two
frame-level calls.  Each pass is timed by the library's CUDA events around its launch (VxStats.last_ms), averaged over `iters` frames
after `warm` warm-up frames (16: steady-state history); prints one JSON line with ms per pass and algorithmic GB/s (each input plane read once, each output plane written
once) against the measured HBM peak.

  python tools/denoise_probe.py [iters [width height]]          default 20 frames at 1920 x 1080
  ncu --metrics gpu__time_duration.sum --clock-control none python tools/denoise_probe.py 3      launch list

Torch-free: device planes are allocated and filled through the ABI itself (vxpt_shared_alloc / vxpt_copy_async), so the script also runs
against the emulated ABI of the CPU suite (tools check: VXPT_PROBE_EMULATED=1, small sizes)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("VXPT_PROBE_EMULATED"):     # development check without a GPU: never set on the GPU box
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from host_shadow import hostemu
    from voxelpathtracer_b200 import abi as _abi
    _abi.LIB_PATH = hostemu.build()
if os.environ.get("VXPT_LIB"):  # development: an experiment build of the library (build.py --out=...), never the product path
    from voxelpathtracer_b200 import abi as _abi2
    _abi2.LIB_PATH = os.path.join(os.path.dirname(_abi2.LIB_PATH), os.environ["VXPT_LIB"])
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import assets, camera, denoise, world  # noqa: E402

# algorithmic bytes per pixel: inputs read once + outputs written once (fp32 planes; ids 1 B)
BYTES = {"material": 6 + 44,
         "svgf_initial": 4 + 1 + (16 + 8 + 4 + 8) * 2,
         "svgf_temporal": (4 + 1 + 1) * 2 + (16 + 8 + 4 + 8) + (16 + 8 + 12 + 8) * 2,
         "svgf_variance": 4 + 1 + 16 + 8 + 12 + 16 + 8 + 4,
         "svgf_spatial": 4 + 1 + 16 + 8 + 4 + 8 + 12 + 16 + 8 + 4 + 8,
         "shadow_temporal": 4 + 1 + 4 + 1 + 4 + 4 + 4 + 4 + 4,
         "shadow_filter": 4 + 1 + 4 + 4 + 4 + 4,
         # reflection pass (a14): G-buffer t / normal id / block id, material normal + PBR, GI SH + CoCg in; colour, hit distance, mask out
         "reflection": 4 + 1 + 1 + 12 + 16 + 16 + 8 + 16 + 4 + 1}
BYTES["svgf_frame"] = BYTES["svgf_initial"] + BYTES["svgf_temporal"] + BYTES["svgf_variance"] + 5 * BYTES["svgf_spatial"] + 6
BYTES["shadow_filter_frame"] = BYTES["shadow_temporal"] + BYTES["shadow_filter"] + 4


class DevicePlanes:
    def __init__(self, r):
        self.r, self.ptrs = r, []

    def new(self, shape, dtype=np.float32):
        ptr, _ = self.r.shared_alloc(max(int(np.prod(shape)) * np.dtype(dtype).itemsize, 256))
        self.ptrs.append(ptr)
        return ptr

    def close(self):
        self.r.sync()
        for p in self.ptrs:
            self.r.shared_close(p)


def material_mips(n_layers):
    """(albedo, normal, pbr) RGBA8 mip chains of the synthetic level-0 textures (the ones tests/material_cases.py uses)."""
    a, n, p = assets.synthetic_material_lod0(n_layers)
    return assets.rgba_mip_chain(a, srgb=True), assets.rgba_mip_chain(n), assets.rgba_mip_chain(p)


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080)
    r = vx.Renderer(0)
    r.upload_world(world.generate_plains(assets.load_plains_columns()))
    r.build_distance_field()
    print(json.dumps(measure(r, iters, W, H)))


def measure(r, iters=20, W=1920, H=1080, warm=16):
    """ms per pass on renderer `r` (world uploaded, distance field built), planes resident in device memory; bench.py calls this too.
    warm: frames before the timed ones.  16 by default = a denoiser in steady state: VarianceEstimate.glsl filters 9 x 9 taps around a pixel
    only while its history is shorter than 12 frames (:101-104) and SpatialFilter.glsl's 'strong' mode ends after 8, so the first dozen
    frames after a history reset cost about twice a later frame (warm=3 times those)."""
    mats = assets.load_materials()
    sun, moon, stronger, vis = camera.sun_moon_direction(50.0)
    r.load_scene_tables(mats, assets.load_blue_noise(), assets.analytic_sky(16, sun), assets.load_shadow_noise())
    r.set_gbuffer_textures(*material_mips(mats["albedo_lod3"].shape[0]))
    dev = DevicePlanes(r)
    sh = denoise.plane_shapes(W, H)

    def planes(names):
        return {k: dev.new(sh[k]) for k in names}

    def gbuffer():
        return {"t": dev.new((H, W)), "normal_id": dev.new((H, W), np.uint8), "block_id": dev.new((H, W), np.uint8), "inv_t": dev.new((H, W))}

    ms = {k: [] for k in BYTES}
    g2 = [gbuffer(), gbuffer()]                         # this frame's and the previous frame's G-buffer
    temporal = [planes(("sh", "cocg", "utility", "ao_sky")) for _ in range(2)]
    shadow_t = [planes(("shadow", "frames")) for _ in range(2)]
    s = {"shadow": dev.new((H, W), np.uint8), "transversal": dev.new((H, W))}
    d = planes(("sh", "cocg", "luma", "ao_sky"))
    pre = planes(("sh", "cocg", "luma", "ao_sky"))
    var = planes(("sh", "cocg", "variance"))
    pong = [planes(("sh", "cocg", "variance", "ao_sky")) for _ in range(2)]
    mat = {"albedo": dev.new((H, W, 3)), "normal": dev.new((H, W, 3)), "pbr": dev.new((H, W, 4)), "texture_ao": dev.new((H, W))}
    filtered, frame_out, shadow_frame_out = dev.new((H, W)), planes(("sh", "cocg", "variance", "ao_sky")), dev.new((H, W))
    refl = {"color": dev.new((H, W, 4)), "hit_distance": dev.new((H, W)), "emissive_mask": dev.new((H, W), np.uint8)}
    prev_fc = None
    try:
        for f in range(iters + warm):
            fc = camera.FpsCamera(position=(192.0 + 0.05 * f, 75.0, 192.0 + 0.03 * f), pitch_deg=-20.0, yaw_deg=90.0 + 0.2 * f, aspect=W / H)
            cam = fc.vx_camera(W, H)
            g, pg = g2[f & 1], (g2[(f & 1) ^ 1] if f else g2[0])
            r.trace_primary(cam, vx.primary_params(350, camera.taa_jitter(f)), g)
            r.trace_shadow(cam, g, vx.shadow_params(stronger, frame=f, soft=True), s)
            r.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=f), d)
            view, proj = (prev_fc or fc).view_projection_f32()
            cview, cproj = fc.view_projection_f32()
            rec = f >= warm

            def timed(name, fn):
                out = fn()
                if rec:
                    ms[name].append(r.stats()["last_ms"])
                return out

            timed("material", lambda: r.generate_gbuffer(cam, g, vx.material_params(mats["grass_props"]), mat))
            # config 3's reflection pass: 1 spp, rough, u_Halton = GetTAAJitterSecondary(frame), normals / PBR from the material pass
            rp = vx.reflection_params(sun, moon, stronger, fc.position, mats["grass_props"], spp=1, rough=True, frame=f, halton=camera.taa_jitter_secondary(f))
            timed("reflection", lambda: r.trace_reflection(cam, g, d, rp, refl, g_normal=mat["normal"], g_pbr=mat["pbr"]))
            timed("svgf_initial", lambda: r.svgf_initial(cam, g, d, pre))
            t, pt = temporal[f & 1], temporal[(f & 1) ^ 1]
            timed("svgf_temporal", lambda: r.svgf_temporal(cam, g, pg, pre, pt, denoise.temporal_params(view, proj), t))
            timed("svgf_variance", lambda: r.svgf_variance(cam, g, t, denoise.variance_params(), var))
            cur = {"sh": var["sh"], "cocg": var["cocg"], "variance": var["variance"], "ao_sky": t["ao_sky"]}
            for n, step in enumerate(denoise.ATROUS_STEPS):
                cur = timed("svgf_spatial", lambda: r.svgf_spatial(cam, g, cur, t["utility"], denoise.spatial_params(step, time=1.0 + f / 60.0), pong[n & 1]))
            st, pst = shadow_t[f & 1], shadow_t[(f & 1) ^ 1]
            timed("shadow_temporal", lambda: r.shadow_temporal(cam, g, pg, s, pst, denoise.shadow_temporal_params(view, proj), st))
            timed("shadow_filter", lambda: r.shadow_filter(cam, g, st, s["transversal"], denoise.shadow_filter_params(1.0), filtered))
            timed("svgf_frame", lambda: r.svgf_frame(cam, g, d, denoise.frame_params(cview, cproj, time=1.0 + f / 60.0, reset_history=(f == 0)), frame_out))
            timed("shadow_filter_frame", lambda: r.shadow_filter_frame(cam, g, s, denoise.shadow_frame_params(cview, cproj, reset_history=(f == 0)), shadow_frame_out))
            prev_fc = fc
    finally:
        dev.close()
    peak = 6451.5
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    out = {"resolution": [W, H], "iters": iters, "hbm_peak_gbs": peak, "passes": {}}
    for k, v in ms.items():
        m = float(np.mean(v))
        out["passes"][k] = {"ms": m, "calls_per_frame": len(v) // iters, "algorithmic_bytes": W * H * BYTES[k],
                            "achieved_gbs": W * H * BYTES[k] / (m * 1e-3) / 1e9, "frac": W * H * BYTES[k] / (m * 1e-3) / 1e9 / peak}
    return out


if __name__ == "__main__":
    main()
