"""The per-pixel CUDA kernels' own source (csrc/trace.cu, csrc/trace_reflection.cu), compiled by g++ and run on the CPU
(tests/host_shadow), against the oracle: the CPU suite's check that a kernel computes what the reference computes before a GPU
ever sees it.  Everything here is bit-exact: the two sides evaluate the same IEEE operations in the same order."""
import numpy as np
import pytest

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import camera

from host_shadow import kernels_on_host as koh

pytestmark = pytest.mark.skipif(not koh.available(), reason="CUDA toolkit headers not present")


@pytest.fixture(scope="module")
def host_kernels(oracles):
    made = {}

    def get(name, layout=1):
        if (name, layout) not in made:
            made[(name, layout)] = koh.HostKernels(oracles[name], layout)
        return made[(name, layout)]

    yield get
    for k in made.values():
        k.close()


@pytest.mark.parametrize("name,layout,pitch,jitter", [("plains", 1, -20.0, 1), ("plains", 0, 0.0, None), ("city", 1, -10.0, 5)])
def test_primary_shadow_gi_kernels_equal_the_oracle(oracles, host_kernels, scene_tables, name, layout, pitch, jitter):
    o, k = oracles[name], host_kernels(name, layout)
    W, H = 256, 144
    fc = camera.FpsCamera(pitch_deg=pitch) if name != "city" else camera.FpsCamera(position=(100.0, 60.0, 100.0), pitch_deg=pitch, yaw_deg=45.0)
    cam = fc.vx_camera(W, H)
    pp = vx.primary_params(350, camera.taa_jitter(jitter) if jitter is not None else None)
    g_ref, st_ref = o.trace_primary(cam, pp)
    g, st = k.trace_primary(cam, pp)
    for key in g_ref:
        assert np.array_equal(g[key], g_ref[key], equal_nan=True), key
    assert st == st_ref
    sp = vx.shadow_params(scene_tables["stronger"], frame=3, soft=True)
    s_ref, st_ref = o.trace_shadow(cam, g_ref, sp)
    s, st = k.trace_shadow(cam, g_ref, sp)
    assert np.array_equal(s["shadow"], s_ref["shadow"]) and np.array_equal(s["transversal"], s_ref["transversal"])
    assert st == st_ref
    dp = vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=2, frame=9)
    d_ref, st_ref = o.trace_diffuse(cam, g_ref, dp)
    d, st = k.trace_diffuse(cam, g_ref, dp)
    for key in d_ref:
        assert np.array_equal(d[key], d_ref[key]), key
    assert st == st_ref


def test_reflection_kernel_equals_the_oracle(oracles, host_kernels, scene_tables):
    o, k = oracles["gi_box"], host_kernels("gi_box", 1)
    W, H = 192, 108
    fc = camera.FpsCamera(pitch_deg=-20.0)
    cam = fc.vx_camera(W, H)
    sun, moon, stronger, vis = scene_tables["sun"], scene_tables["moon"], scene_tables["stronger"], scene_tables["sun_visibility"]
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    d, _ = o.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=4))
    for halton in (camera.taa_jitter_secondary(4), (0.0, 0.0), (-1.25, 1.75)):   # u_Halton as Pipeline.cpp:3032 sets it; none; both signs, beyond a texel
        rp = vx.reflection_params(sun, moon, stronger, fc.position, scene_tables["materials"]["grass_props"], spp=2, rough=True, frame=4, halton=halton)
        ref, st_ref = o.trace_reflection(cam, g, d, rp)
        out, st = k.trace_reflection(cam, g, d, rp)
        for key in ref:
            assert np.array_equal(out[key], ref[key]), (halton, key)
        assert st == st_ref


def test_alpha_tested_kernels_equal_the_oracle(worlds, oracle_dfs, scene_tables):
    """u_ShouldAlphaTest: primary and shadow kernels through VoxelTraversalDF_AlphaTest + StopRay on the orchard world."""
    from oracle import vxo
    from voxelpathtracer_b200 import assets
    mats = scene_tables["materials"]
    o = vxo.Oracle(worlds["orchard"].data, oracle_dfs["orchard"])
    o.set_tables(mats, scene_tables["blue_noise"], scene_tables["sky"], scene_tables["shadow_noise"])
    o.set_alpha_mips(assets.alpha_mip_pyramid(assets.synthetic_alpha_lod0(mats["albedo_lod3"].shape[0], [int(mats["table"][7])])))
    for layout in (1, 0):
        k = koh.HostKernels(o, layout)
        W, H = 256, 144
        cam = camera.FpsCamera(position=(192.0, 66.0, 192.0), pitch_deg=-8.0, yaw_deg=30.0 * layout).vx_camera(W, H)
        pp = vx.primary_params(350, camera.taa_jitter(2), alpha_test=True, fov_degrees=60.0)
        g_ref, st_ref = o.trace_primary(cam, pp)
        plain, _ = o.trace_primary(cam, vx.primary_params(350, camera.taa_jitter(2)))
        assert (plain["block_id"] != g_ref["block_id"]).sum() > 1000      # the alpha test does change what is hit
        g, st = k.trace_primary(cam, pp)
        for key in g_ref:
            assert np.array_equal(g[key], g_ref[key], equal_nan=True), key
        assert st == st_ref
        sp = vx.shadow_params(scene_tables["stronger"], frame=5, soft=True, alpha_test=True, fov_degrees=60.0)
        s_ref, st_ref = o.trace_shadow(cam, g_ref, sp)
        s, st = k.trace_shadow(cam, g_ref, sp)
        assert np.array_equal(s["shadow"], s_ref["shadow"]) and np.array_equal(s["transversal"], s_ref["transversal"])
        assert st == st_ref
        k.close()
