"""Alpha-tested traversal (VoxelTraversalDF_AlphaTest + StopRay + CalculateUV, InitialRayTraceFrag.glsl:189-305,
ShadowRayTraceFrag.glsl:105-220; u_ShouldAlphaTest, off by default in the reference).

CPU: the oracle against the committed digests of the reference's own shaders (tests/golden/ref_shader_alpha_digests.json, made by
tools/make_ref_alpha_golden.py) and live against the compiled shaders on edge cases.  GPU: the CUDA kernels through the C ABI against
the oracle and the same digests.  Everything is bit-exact."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import abi, assets, camera, world
from oracle import ref_shaders, vxo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
from make_ref_alpha_golden import alpha_cases, alpha_inputs  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_shaders.available(), reason="oracle/_ref/libref_shaders.so not built (needs /root/reference at build time)")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def digests():
    with open(os.path.join(ROOT, "tests", "golden", "ref_shader_alpha_digests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def alpha_mips(scene_tables):
    return alpha_inputs(scene_tables["materials"])


@pytest.fixture(scope="module")
def orchard_oracle(worlds, oracle_dfs, scene_tables, alpha_mips):
    o = vxo.Oracle(worlds["orchard"].data, oracle_dfs["orchard"])
    o.set_tables(scene_tables["materials"], scene_tables["blue_noise"], scene_tables["sky"], scene_tables["shadow_noise"])
    o.set_alpha_mips(alpha_mips)
    return o


def case_params(case, scene_tables):
    name, W, H, cam_kw, jf, sframe = case
    cam = camera.FpsCamera(**cam_kw).vx_camera(W, H)
    pp = vx.primary_params(350, None if jf is None else camera.taa_jitter(jf), alpha_test=True, fov_degrees=60.0)
    sp = vx.shadow_params(scene_tables["stronger"], frame=sframe, soft=True, alpha_test=True, fov_degrees=60.0)
    return name, cam, pp, sp


def test_alpha_pyramid_layout():
    a0 = assets.synthetic_alpha_lod0(3, [1])
    pyr = assets.alpha_mip_pyramid(a0)
    assert pyr.shape == (3, abi.ALPHA_MIP_TEXELS) and abi.ALPHA_MIP_TEXELS == 349524
    assert np.array_equal(pyr[:, :512 * 512].reshape(3, 512, 512), a0)
    off = 0
    for k in range(9):
        assert off == (4 ** 10 - 4 ** (10 - k)) // 3          # the closed form the kernels use
        off += (512 >> k) ** 2
    assert (pyr[0] == 255).all() and (pyr[2] == 255).all() and 0.3 < (pyr[1] < 249).mean() < 0.6


@pytest.mark.parametrize("case", alpha_cases(), ids=lambda c: c[0])
def test_oracle_equals_the_reference_shader_digests(orchard_oracle, worlds, scene_tables, alpha_mips, digests, case):
    assert sha(worlds["orchard"].data) == digests["world"] and sha(orchard_oracle.df) == digests["df"] and sha(alpha_mips) == digests["alpha_mips"]
    name, cam, pp, sp = case_params(case, scene_tables)
    g, st = orchard_oracle.trace_primary(cam, pp)
    for k in ("t", "normal_id", "block_id", "inv_t"):
        assert sha(g[k]) == digests["primary"][name][k], k
    plain, pst = orchard_oracle.trace_primary(cam, vx.primary_params(350, (pp.jitter[0], pp.jitter[1]) if pp.jitter_enable else None))
    assert int((plain["block_id"] != g["block_id"]).sum()) == digests["primary"][name]["pixels_changed_by_the_alpha_test"] > 10000
    assert st["vox_fetches"] > pst["vox_fetches"]                      # StopRay's extra GetVoxel calls are counted
    s, _ = orchard_oracle.trace_shadow(cam, g, sp)
    assert sha(s["shadow"]) == digests["shadow"][name]["shadow"] and sha(s["transversal"]) == digests["shadow"][name]["transversal"]


def test_alpha_test_is_a_no_op_without_transparent_blocks(oracles, scene_tables, alpha_mips):
    """Every StopRay returns true on the first line when no block is Transparent -> identical planes (but one more block fetch per
    ray that ends on E == 0 inside the loop)."""
    o = oracles["plains"]
    o.set_alpha_mips(alpha_mips)
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(320, 180)
    a, _ = o.trace_primary(cam, vx.primary_params(350, alpha_test=True))
    b, _ = o.trace_primary(cam, vx.primary_params(350))
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


def test_opaque_alpha_stops_at_the_leaves(orchard_oracle, worlds, scene_tables):
    """With alpha = 255 everywhere the leaves stop every ray: same hits as the plain traversal."""
    o = vxo.Oracle(worlds["orchard"].data, orchard_oracle.df)
    o.set_tables(scene_tables["materials"], scene_tables["blue_noise"], scene_tables["sky"], scene_tables["shadow_noise"])
    o.set_alpha_mips(np.full((scene_tables["materials"]["albedo_lod3"].shape[0], abi.ALPHA_MIP_TEXELS), 255, np.uint8))
    cam = camera.FpsCamera(position=(192.0, 66.0, 192.0), pitch_deg=-8.0).vx_camera(320, 180)
    a, _ = o.trace_primary(cam, vx.primary_params(350, alpha_test=True))
    b, _ = o.trace_primary(cam, vx.primary_params(350))
    for k in a:
        assert np.array_equal(a[k], b[k], equal_nan=True), k


@needs_ref
@pytest.mark.parametrize("cam_kw,fov,max_it", [
    (dict(position=(192.0, 66.0, 192.0), pitch_deg=-8.0), 60.0, 350),
    (dict(position=(150.0, 72.0, 210.0), pitch_deg=-60.0, yaw_deg=200.0), 90.0, 475),        # other FOV -> other g_K / LODs
    (dict(position=(192.0, 100.0, 192.0), pitch_deg=-89.0), 45.0, 350),                       # straight down through canopies
    (dict(position=(-30.0, 90.0, -30.0), pitch_deg=-15.0, yaw_deg=45.0), 60.0, 350),          # camera outside the volume
    (dict(position=(192.0, 66.0, 192.0), pitch_deg=0.0, yaw_deg=90.0), 60.0, 40),             # axis-aligned rays, tight iteration cap
])
def test_oracle_equals_the_reference_shaders_live(orchard_oracle, worlds, scene_tables, alpha_mips, cam_kw, fov, max_it):
    W, H = 160, 90
    cam = camera.FpsCamera(fov_deg=fov, **cam_kw).vx_camera(W, H)
    pp = vx.primary_params(max_it, camera.taa_jitter(4), alpha_test=True, fov_degrees=fov)
    table = scene_tables["materials"]["table"]
    g, _ = orchard_oracle.trace_primary(cam, pp)
    r = ref_shaders.trace_primary(worlds["orchard"].data, orchard_oracle.df, cam, pp, table, alpha_mips)
    for k in ("t", "normal_id", "block_id", "inv_t"):
        assert np.array_equal(g[k], r[k], equal_nan=True), k
    for soft in (True, False):
        sp = vx.shadow_params(scene_tables["stronger"], frame=2, soft=soft, alpha_test=True, fov_degrees=fov)
        s, _ = orchard_oracle.trace_shadow(cam, g, sp)
        rs = ref_shaders.trace_shadow(worlds["orchard"].data, orchard_oracle.df, cam, g, sp, scene_tables["shadow_noise"], table, alpha_mips)
        assert np.array_equal(s["shadow"], rs["shadow"]) and np.array_equal(s["transversal"], rs["transversal"])


# ------------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("case", alpha_cases(), ids=lambda c: c[0])
def test_cuda_alpha_traversal_equals_oracle_and_reference_digests(renderer, orchard_oracle, worlds, scene_tables, alpha_mips, digests, case):
    renderer.upload_world(worlds["orchard"])
    renderer.build_distance_field()
    renderer.set_albedo_alpha_mips(alpha_mips)
    name, cam, pp, sp = case_params(case, scene_tables)
    g_ref, st_ref = orchard_oracle.trace_primary(cam, pp)
    s_ref, sst_ref = orchard_oracle.trace_shadow(cam, g_ref, sp)
    for layout in (1, 0):
        renderer.set_option(abi.OPT_TRAVERSAL_LAYOUT, layout)
        renderer.reset_stats()
        g = renderer.trace_primary(cam, pp, renderer.alloc_gbuffer(cam.width, cam.height, hit_voxel=True))
        st = renderer.stats()
        for k in ("t", "normal_id", "block_id", "inv_t", "hit_voxel"):
            assert np.array_equal(g[k], g_ref[k], equal_nan=True), (layout, k)
        for k in ("t", "normal_id", "block_id", "inv_t"):
            assert sha(g[k]) == digests["primary"][name][k], (layout, k)
        assert (st["rays"], st["df_fetches"], st["vox_fetches"]) == (st_ref["rays"], st_ref["df_fetches"], st_ref["vox_fetches"])
        renderer.reset_stats()
        s = renderer.trace_shadow(cam, g, sp, renderer.alloc_shadow(cam.width, cam.height))
        st = renderer.stats()
        assert np.array_equal(s["shadow"], s_ref["shadow"]) and np.array_equal(s["transversal"], s_ref["transversal"])
        assert sha(s["shadow"]) == digests["shadow"][name]["shadow"] and sha(s["transversal"]) == digests["shadow"][name]["transversal"]
        assert (st["rays"], st["df_fetches"], st["vox_fetches"]) == (sst_ref["rays"], sst_ref["df_fetches"], sst_ref["vox_fetches"])
    renderer.set_option(abi.OPT_TRAVERSAL_LAYOUT, 1)


@pytest.mark.gpu
def test_cuda_alpha_test_needs_its_inputs(worlds):
    r = vx.Renderer(0)
    try:
        r.upload_world(worlds["superflat"])
        r.build_distance_field()
        cam = camera.FpsCamera().vx_camera(64, 36)
        with pytest.raises(abi.VxptError) as e:
            r.trace_primary(cam, vx.primary_params(350, alpha_test=True), r.alloc_gbuffer(64, 36))
        assert e.value.code == abi.E_STATE                     # no material table / alpha pyramid on this handle
    finally:
        r.close()
