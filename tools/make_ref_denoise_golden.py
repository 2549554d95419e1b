#!/usr/bin/env python3
"""Golden digests of the SVGF denoiser from the reference's OWN shaders: Core/Shaders/SVGF/{TemporalFilter,VarianceEstimate,SpatialFilter}.glsl
and Core/Shaders/{ShadowTemporalFilter,ShadowFilter}.glsl
compiled as C++ (oracle/_ref/libref_shaders.so, oracle/ref_denoise_driver.cpp), fed by the reference's InitialRayTraceFrag / DiffuseRayTraceFrag
for the G-buffer and the 1-spp GI planes, on the frame sequences of tests/denoise_cases.py.  Run in the build container (needs
/root/reference); writes tests/golden/ref_denoise_digests.json, which is committed."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import assets, camera, world  # noqa: E402
from oracle import ref_shaders  # noqa: E402
import denoise_cases as dc  # noqa: E402


def main():
    cols = assets.load_plains_columns()
    mats, bn = assets.load_materials(), assets.load_blue_noise()
    sun, moon, stronger, vis = camera.sun_moon_direction(50.0)
    sn = assets.load_shadow_noise()
    sky = assets.analytic_sky(16, sun)
    make = {"plains": lambda: world.generate_plains(cols), "gi_box": lambda: world.generate_gi_box(cols), "city": world.generate_city}
    out = {}
    for name, (wname, W, H, cams) in dc.SEQUENCES.items():
        t0 = time.time()
        w = make[wname]()
        df = ref_shaders.df_build(w.data)

        def trace(cam, f):
            g = ref_shaders.trace_primary(w.data, df, cam, vx.primary_params(350))
            d = ref_shaders.trace_diffuse(w.data, df, cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=f), mats, bn, sky)
            return g, d

        out[name] = [dc.frame_digest(fr) for fr in dc.run_sequence(name, trace, ref_shaders, None)]

        def trace_shadow(cam, f):
            g = ref_shaders.trace_primary(w.data, df, cam, vx.primary_params(350))
            s = ref_shaders.trace_shadow(w.data, df, cam, g, vx.shadow_params(stronger, frame=f, soft=True), sn)
            return g, s

        out["shadow:" + name] = [dc.shadow_frame_digest(fr) for fr in dc.run_shadow_sequence(name, trace_shadow, ref_shaders)]
        print(f"{name}: {len(out[name])} frames, {time.time() - t0:.1f} s", flush=True)
    with open(os.path.join(ROOT, "tests", "golden", "ref_denoise_digests.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
