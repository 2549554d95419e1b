#!/usr/bin/env python3
"""Write tests/golden/oracle_digests.json: sha256 digests of the ORACLE's outputs on the BASELINE.json configs.

The reference ships no golden vectors and cannot run headless here, so these digests pin the
oracle against drift and give the GPU tests a full-size target that needs no CPU tracing at test time.
Regenerate only when the oracle's pinned semantics change deliberately."""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import assets, camera, world  # noqa: E402
from oracle import vxo  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def frame_cases():
    """(name, world, width, height, pitch, jitter_frame or None) — the primary-ray golden frames."""
    cases = []
    for pitch in (0.0, -20.0):
        for jf in (None, 0, 1, 17, 63):
            cases.append((f"superflat_640x360_p{int(pitch)}_j{jf}", "superflat", 640, 360, pitch, jf))
    cases.append(("plains_1920x1080_p-20_jNone", "plains", 1920, 1080, -20.0, None))
    cases.append(("plains_1920x1080_p0_j7", "plains", 1920, 1080, 0.0, 7))
    cases.append(("city_1920x1080_p-20_jNone", "city", 1920, 1080, -20.0, None))
    cases.append(("gi_box_3840x2160_p-20_jNone", "gi_box", 3840, 2160, -20.0, None))
    return cases


def main():
    cols = assets.load_plains_columns()
    rng = np.random.RandomState(5)
    sparse = world.World()
    idx = rng.randint(0, sparse.data.size, size=400)
    sparse.data[idx] = rng.randint(1, 100, size=400)
    worlds = {"superflat": world.generate_superflat(), "plains": world.generate_plains(cols), "gi_box": world.generate_gi_box(cols),
              "city": world.generate_city(), "sparse": sparse}
    out = {"world": {}, "df": {}, "primary": {}, "shadow": {}, "diffuse": {}}
    sun, moon, stronger, sunvis = camera.sun_moon_direction(50.0)
    tables = (assets.load_materials(), assets.load_blue_noise(), assets.analytic_sky(16, sun), assets.load_shadow_noise())
    oracles = {}
    for name, w in worlds.items():
        out["world"][name] = sha(w.data)
        df = vxo.df_build(w.data)
        out["df"][name] = sha(df)
        oracles[name] = vxo.Oracle(w.data, df)
        oracles[name].set_tables(*tables)
        print(name, "fill", float((w.data > 0).mean()), "df max", int(df.max()))
    for cname, wname, W, H, pitch, jf in frame_cases():
        cam = camera.FpsCamera(pitch_deg=pitch, aspect=W / H).vx_camera(W, H)
        pp = vx.primary_params(350, None if jf is None else camera.taa_jitter(jf))
        g, st = oracles[wname].trace_primary(cam, pp)
        out["primary"][cname] = {"t": sha(g["t"]), "normal_id": sha(g["normal_id"]), "block_id": sha(g["block_id"]),
                                 "hit_voxel": sha(g["hit_voxel"]), "stats": st, "hit_fraction": float((g["t"] > 0).mean())}
        print(cname, st, out["primary"][cname]["hit_fraction"])
        if wname in ("plains", "city") and jf is None:
            s, sst = oracles[wname].trace_shadow(cam, g, vx.shadow_params(stronger, frame=5, soft=True))
            out["shadow"][cname] = {"shadow": sha(s["shadow"]), "transversal": sha(s["transversal"]), "stats": sst,
                                    "shadowed_fraction": float(s["shadow"].mean())}
            d, dst = oracles[wname].trace_diffuse(cam, g, vx.diffuse_params(sun, moon, sunvis, spp=1, frame=7))
            out["diffuse"][cname] = {"stats": dst, "mean_luma": float(d["luma"].mean()), "sh": sha(d["sh"])}
            print("  shadow", sst, "diffuse", dst)
    with open(os.path.join(ROOT, "tests", "golden", "oracle_digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
