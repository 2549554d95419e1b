#!/usr/bin/env python3
"""Write tests/golden/ref_shader_alpha_digests.json: sha256 digests of what the REFERENCE'S OWN InitialRayTraceFrag.glsl and
ShadowRayTraceFrag.glsl produce with u_ShouldAlphaTest = true (VoxelTraversalDF_AlphaTest + StopRay), compiled as C++ against the
reference's vendored glm (oracle/_ref/libref_shaders.so), on the orchard world (plains + trees with Transparent leaves) and the
synthetic cut-out alpha pyramid of voxelpathtracer_b200.assets.  Needs /root/reference (this container only); the JSON travels."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import assets, camera, world  # noqa: E402
from oracle import ref_shaders  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def alpha_cases():
    """(name, width, height, camera kwargs, jitter frame, shadow frame)"""
    return [("orchard_640x360_low", 640, 360, dict(position=(192.0, 66.0, 192.0), pitch_deg=-8.0), None, 5),
            ("orchard_640x360_canopy_j3", 640, 360, dict(position=(150.0, 72.0, 210.0), pitch_deg=-30.0, yaw_deg=120.0), 3, 11),
            ("orchard_1920x1080", 1920, 1080, dict(position=(192.0, 75.0, 192.0), pitch_deg=-20.0), None, 7)]


def alpha_inputs(materials):
    return assets.alpha_mip_pyramid(assets.synthetic_alpha_lod0(materials["albedo_lod3"].shape[0], [int(materials["table"][world.LEAVES])]))


def main():
    w = world.generate_orchard(assets.load_plains_columns())
    mats, sn = assets.load_materials(), assets.load_shadow_noise()
    alpha = alpha_inputs(mats)
    _, _, stronger, _ = camera.sun_moon_direction(50.0)
    df = ref_shaders.df_build(w.data)
    out = {"source": "Core/Shaders/InitialRayTraceFrag.glsl + ShadowRayTraceFrag.glsl with u_ShouldAlphaTest = true, compiled as C++ (oracle/_ref/libref_shaders.so)",
           "world": sha(w.data), "df": sha(df), "alpha_mips": sha(alpha), "primary": {}, "shadow": {}}
    for name, W, H, cam_kw, jf, sframe in alpha_cases():
        t0 = time.time()
        cam = camera.FpsCamera(**cam_kw).vx_camera(W, H)
        pp = vx.primary_params(350, None if jf is None else camera.taa_jitter(jf), alpha_test=True, fov_degrees=60.0)
        g = ref_shaders.trace_primary(w.data, df, cam, pp, mats["table"], alpha)
        plain = ref_shaders.trace_primary(w.data, df, cam, vx.primary_params(350, None if jf is None else camera.taa_jitter(jf)))
        out["primary"][name] = {k: sha(g[k]) for k in ("t", "normal_id", "block_id", "inv_t")}
        out["primary"][name]["pixels_changed_by_the_alpha_test"] = int((plain["block_id"] != g["block_id"]).sum())
        sp = vx.shadow_params(stronger, frame=sframe, soft=True, alpha_test=True, fov_degrees=60.0)
        s = ref_shaders.trace_shadow(w.data, df, cam, g, sp, sn, mats["table"], alpha)
        out["shadow"][name] = {"shadow": sha(s["shadow"]), "transversal": sha(s["transversal"]), "shadowed_fraction": float(s["shadow"].mean())}
        print(f"{name}: {time.time() - t0:.1f} s, {out['primary'][name]['pixels_changed_by_the_alpha_test']} pixels changed", flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "ref_shader_alpha_digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
