#!/usr/bin/env python3
"""Write tests/golden/ref_shader_digests.json: sha256 digests of what the REFERENCE'S OWN shaders produce — ManhattanDistance{X,Y,Z}.comp,
InitialRayTraceFrag.glsl, ShadowRayTraceFrag.glsl and DiffuseRayTraceFrag.glsl compiled as C++ against the reference's vendored glm (oracle/_ref/libref_shaders.so, built by oracle/Makefile
from /root/reference) — on the same worlds and frames as tests/golden/oracle_digests.json.  Needs /root/reference (this container only);
the committed JSON travels.  These are the golden vectors that pin the oracle: tests/test_oracle_vs_reference_shaders.py."""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import assets, camera, world  # noqa: E402
from oracle import ref_shaders  # noqa: E402
from make_golden_outputs import frame_cases  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def synthetic_material_planes(g, W, H):
    """Deterministic stand-ins for the G-buffer material pass's outputs (u_GBufferNormals, u_GBufferPBR): the face normal tilted by a
    fixed integer pattern and renormalised in float64, roughness / metalness from integer patterns.  No RNG, so tests can rebuild them."""
    j, i = np.mgrid[0:H, 0:W]
    faces = np.array([[0, 0, 1], [0, 0, -1], [0, 1, 0], [0, -1, 0], [-1, 0, 0], [1, 0, 0], [0, 1, 0]], dtype=np.float64)
    n = faces[np.minimum(g["normal_id"], 6)] + 0.05 * np.stack([((i * 3 + j) % 7) - 3.0, ((i + j * 5) % 5) - 2.0, ((i * 2 + j * 3) % 9) - 4.0], -1)
    g_normal = np.ascontiguousarray(n / np.linalg.norm(n, axis=-1, keepdims=True), dtype=np.float32)
    rough = (0.05 + 0.9 * ((i * 7 + j * 13) % 97) / 96.0)
    metal = ((i + j) % 5 == 0) * 0.9
    g_pbr = np.ascontiguousarray(np.stack([rough, metal, np.zeros((H, W)), np.zeros((H, W))], -1), dtype=np.float32)
    return g_normal, g_pbr


def reflection_cases():
    """(name, world, width, height, camera kwargs, spp, rough, checkerboard, frame).  u_Halton = GetTAAJitterSecondary(frame) as
    Core/Pipeline.cpp:3032 sets it every frame (frame < 0: entry 0); the first case is BASELINE configs[2] at its own resolution."""
    return [("plains_1920x1080_spp1_rough", "plains", 1920, 1080, dict(pitch_deg=-20.0), 1, True, False, 5),
            ("plains_960x540_spp2_rough", "plains", 960, 540, dict(pitch_deg=-20.0), 2, True, False, 7),
            ("city_640x360_spp4_checker", "city", 640, 360, dict(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0), 4, True, True, 12),
            ("gi_box_640x360_mirror", "gi_box", 640, 360, dict(pitch_deg=-20.0), 2, False, False, -1)]


def reflection_digests(out, worlds, dfs, sun, moon, stronger, sunvis, mats, bn, sky):
    for cname, wname, W, H, cam_kw, spp, rough, checker, frame in reflection_cases():
        t0 = time.time()
        fc = camera.FpsCamera(**cam_kw)
        cam = fc.vx_camera(W, H)
        wd, df = worlds[wname].data, dfs[wname]
        g = ref_shaders.trace_primary(wd, df, cam, vx.primary_params(350))
        d = ref_shaders.trace_diffuse(wd, df, cam, g, vx.diffuse_params(sun, moon, sunvis, spp=1, frame=max(frame, 0)), mats, bn, sky)
        g_normal, g_pbr = synthetic_material_planes(g, W, H)
        rp = vx.reflection_params(sun, moon, stronger, fc.position, mats["grass_props"], spp=spp, rough=rough, checkerboard=checker, frame=frame,
                                  halton=camera.taa_jitter_secondary(max(frame, 0)))
        r = ref_shaders.trace_reflection(wd, df, cam, g, d, rp, g_normal, g_pbr, mats, bn, sky)
        out["reflection"][cname] = {"color": sha(r["color"]), "hit_distance": sha(r["hit_distance"]), "emissive_mask": sha(r["emissive_mask"]),
                                    "hit_fraction": float((r["hit_distance"] > 0).mean())}
        print(f"reflection {cname}: {time.time() - t0:.1f} s, hit fraction {out['reflection'][cname]['hit_fraction']:.4f}", flush=True)


def main():
    cols = assets.load_plains_columns()
    rng = np.random.RandomState(5)
    sparse = world.World()
    idx = rng.randint(0, sparse.data.size, size=400)
    sparse.data[idx] = rng.randint(1, 100, size=400)
    worlds = {"superflat": world.generate_superflat(), "plains": world.generate_plains(cols), "gi_box": world.generate_gi_box(cols),
              "city": world.generate_city(), "sparse": sparse}
    out = {"source": "Core/Shaders/ManhattanDistance{X,Y,Z}.comp, InitialRayTraceFrag.glsl, ShadowRayTraceFrag.glsl, DiffuseRayTraceFrag.glsl, ReflectionTraceFrag.glsl compiled as C++ (oracle/_ref/libref_shaders.so)",
           "world": {}, "df": {}, "primary": {}, "shadow": {}, "diffuse": {}, "reflection": {}}
    sun, moon, stronger, sunvis = camera.sun_moon_direction(50.0)
    mats, bn, sky, sn = assets.load_materials(), assets.load_blue_noise(), assets.analytic_sky(16, sun), assets.load_shadow_noise()
    dfs = {}
    only_reflection = "--only-reflection" in sys.argv   # refresh the reflection digests, keep everything else in the committed file
    path = os.path.join(ROOT, "tests", "golden", "ref_shader_digests.json")
    if only_reflection:
        keep = json.load(open(path))
        keep["reflection"] = {}
        for name in {c[1] for c in reflection_cases()}:
            dfs[name] = ref_shaders.df_build(worlds[name].data)
        reflection_digests(keep, worlds, dfs, sun, moon, stronger, sunvis, mats, bn, sky)
        with open(path, "w") as f:
            json.dump(keep, f, indent=1, sort_keys=True)
        return
    for name, w in worlds.items():
        t0 = time.time()
        dfs[name] = ref_shaders.df_build(w.data)
        out["world"][name] = sha(w.data)
        out["df"][name] = sha(dfs[name])
        print(f"df {name}: {time.time() - t0:.1f} s, max {int(dfs[name].max())}", flush=True)
    for cname, wname, W, H, pitch, jf in frame_cases():
        t0 = time.time()
        cam = camera.FpsCamera(pitch_deg=pitch, aspect=W / H).vx_camera(W, H)
        pp = vx.primary_params(350, None if jf is None else camera.taa_jitter(jf))
        g = ref_shaders.trace_primary(worlds[wname].data, dfs[wname], cam, pp)
        out["primary"][cname] = {"t": sha(g["t"]), "normal_id": sha(g["normal_id"]), "block_id": sha(g["block_id"]), "inv_t": sha(g["inv_t"]),
                                 "hit_fraction": float((g["t"] > 0).mean())}
        print(f"primary {cname}: {time.time() - t0:.1f} s, hit fraction {out['primary'][cname]['hit_fraction']:.4f}", flush=True)
        if wname in ("plains", "city") and jf is None:   # the same secondary frames as tools/make_golden_outputs.py
            t0 = time.time()
            s = ref_shaders.trace_shadow(worlds[wname].data, dfs[wname], cam, g, vx.shadow_params(stronger, frame=5, soft=True), sn)
            out["shadow"][cname] = {"shadow": sha(s["shadow"]), "transversal": sha(s["transversal"]), "shadowed_fraction": float(s["shadow"].mean())}
            d = ref_shaders.trace_diffuse(worlds[wname].data, dfs[wname], cam, g, vx.diffuse_params(sun, moon, sunvis, spp=1, frame=7), mats, bn, sky)
            out["diffuse"][cname] = {"sh": sha(d["sh"]), "cocg": sha(d["cocg"]), "luma": sha(d["luma"]), "ao_sky": sha(d["ao_sky"]),
                                     "mean_luma": float(d["luma"].mean())}
            print(f"  shadow + diffuse {cname}: {time.time() - t0:.1f} s", flush=True)
    # BASELINE config 4: 3840x2160, 4-spp GI (and the soft shadow pass) on the gi-box stand-in scene
    t0 = time.time()
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(3840, 2160)
    g = ref_shaders.trace_primary(worlds["gi_box"].data, dfs["gi_box"], cam, vx.primary_params(350))
    s4 = ref_shaders.trace_shadow(worlds["gi_box"].data, dfs["gi_box"], cam, g, vx.shadow_params(stronger, frame=9, soft=True), sn)
    d4 = ref_shaders.trace_diffuse(worlds["gi_box"].data, dfs["gi_box"], cam, g, vx.diffuse_params(sun, moon, sunvis, spp=4, frame=9), mats, bn, sky)
    out["shadow"]["gi_box_3840x2160_f9"] = {"shadow": sha(s4["shadow"]), "transversal": sha(s4["transversal"]), "shadowed_fraction": float(s4["shadow"].mean())}
    out["diffuse"]["gi_box_3840x2160_spp4_f9"] = {"sh": sha(d4["sh"]), "cocg": sha(d4["cocg"]), "luma": sha(d4["luma"]), "ao_sky": sha(d4["ao_sky"]),
                                                  "mean_luma": float(d4["luma"].mean())}
    print(f"config 4 (4K, 4 spp): {time.time() - t0:.1f} s", flush=True)
    reflection_digests(out, worlds, dfs, sun, moon, stronger, sunvis, mats, bn, sky)
    with open(os.path.join(ROOT, "tests", "golden", "ref_shader_digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
