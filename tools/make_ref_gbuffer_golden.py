#!/usr/bin/env python3
"""Golden digests of the G-buffer material pass from the reference's OWN shader: Core/Shaders/GenerateGBuffer.glsl compiled as C++
(oracle/_ref/libref_shaders.so, oracle/ref_gbuffer_driver.cpp) on the frames of tests/material_cases.py, fed by the reference's
InitialRayTraceFrag.glsl for the primary hits.  Run in the build container (needs /root/reference); writes
tests/golden/ref_gbuffer_digests.json, which is committed."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import assets, world  # noqa: E402
from oracle import ref_shaders  # noqa: E402
import material_cases as mc  # noqa: E402


def main():
    cols = assets.load_plains_columns()
    mats = assets.load_materials()
    mips = mc.material_mips(mats["albedo_lod3"].shape[0])
    make = {"plains": lambda: world.generate_plains(cols), "gi_box": lambda: world.generate_gi_box(cols), "city": world.generate_city,
            "superflat": world.generate_superflat}
    worlds, dfs, out = {}, {}, {"mips": {k: mc.sha(m) for k, m in zip(("albedo", "normal", "pbr"), mips)}, "cases": {}}
    for case in mc.CASES:
        name, wname = case[0], case[1]
        if wname not in worlds:
            worlds[wname] = make[wname]()
            dfs[wname] = ref_shaders.df_build(worlds[wname].data)
        t0 = time.time()
        cam = mc.case_camera(case)
        g = ref_shaders.trace_primary(worlds[wname].data, dfs[wname], cam, vx.primary_params(350))
        m = ref_shaders.generate_gbuffer(cam, g, vx.material_params(mats["grass_props"]), mats, mips)
        out["cases"][name] = {k: mc.sha(m[k]) for k in mc.PLANES}
        out["cases"][name].update(hit_fraction=float((g["t"] > 0).mean()), mean_albedo=float(m["albedo"].mean()),
                                  emissive_pixels=int((m["pbr"][..., 3] > 0).sum()))
        print(f"{name}: {time.time() - t0:.1f} s {out['cases'][name]}", flush=True)
    out["pom_cases"] = {}
    for name, idx, kw in mc.POM_CASES:
        case = mc.CASES[idx]
        wname = case[1]
        cam = mc.case_camera(case)
        g = ref_shaders.trace_primary(worlds[wname].data, dfs[wname], cam, vx.primary_params(350))
        m = ref_shaders.generate_gbuffer(cam, g, vx.material_params(mats["grass_props"], pom=True, **kw), mats, mips)
        out["pom_cases"][name] = {k: mc.sha(m[k]) for k in mc.PLANES}
        print(f"{name}: {out['pom_cases'][name]['albedo'][:16]}", flush=True)
    out["lava_cases"], lava = {}, mc.lava_textures()
    out["lava_textures"] = [mc.sha(t) for t in lava]
    for name, idx, block, kw in mc.LAVA_CASES:
        case = mc.CASES[idx]
        wname = case[1]
        cam = mc.case_camera(case)
        g = ref_shaders.trace_primary(worlds[wname].data, dfs[wname], cam, vx.primary_params(350))
        m = ref_shaders.generate_gbuffer(cam, g, vx.material_params(mats["grass_props"], lava_block_id=block, **kw), mats, mips,
                                         mc.seeded_planes(cam.width, cam.height), lava=lava)
        out["lava_cases"][name] = {k: mc.sha(m[k]) for k in mc.PLANES}
        out["lava_cases"][name]["lava_pixels"] = int((g["block_id"] == block).sum())
        print(f"{name}: {out['lava_cases'][name]['albedo'][:16]} lava pixels {out['lava_cases'][name]['lava_pixels']}", flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "ref_gbuffer_digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
