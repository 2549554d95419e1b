"""G-buffer material pass (Core/Shaders/GenerateGBuffer.glsl, Core/Pipeline.cpp:2066-2136).

Pins, in this order: the reference's own shader compiled as C++ (oracle/_ref/libref_shaders.so; committed digests in
tests/golden/ref_gbuffer_digests.json, made by tools/make_ref_gbuffer_golden.py) == the oracle restatement (oracle/vxo_oracle.cpp,
vxo_generate_gbuffer) == the CUDA kernel's source run on the host (tests/host_shadow) == the CUDA kernel on the GPU through the C ABI.
Everything is fp32 arithmetic in a fixed order, so the CPU comparisons are bit for bit; the GPU comparison allows the one pinned
transcendental (log2 of the mip scale factor, evaluated in double by two different libms) to differ by an ulp on isolated texels."""
import json
import os

import numpy as np
import pytest

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import abi, assets, camera
from oracle import ref_shaders, vxo

import material_cases as mc
from host_shadow import kernels_on_host as koh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not ref_shaders.available(), reason="oracle/_ref/libref_shaders.so not built (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def ref_digests():
    with open(os.path.join(ROOT, "tests", "golden", "ref_gbuffer_digests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def mips(scene_tables):
    return mc.material_mips(scene_tables["materials"]["albedo_lod3"].shape[0])


@pytest.fixture(scope="module")
def gb_oracles(oracles, mips):
    class Lazy(dict):
        def __missing__(self, name):
            o = oracles[name]
            o.set_gbuffer_textures(*mips)
            self[name] = o
            return o

    return Lazy()


def _params(scene_tables, **kw):
    return vx.material_params(scene_tables["materials"]["grass_props"], **kw)


# ------------------------------------------------------------------------------------------------------ oracle vs the reference's shader
@pytest.mark.parametrize("case", mc.CASES, ids=[c[0] for c in mc.CASES])
def test_oracle_equals_the_reference_shader_digests(gb_oracles, scene_tables, mips, ref_digests, case):
    if any(mc.sha(m) != ref_digests["mips"][k] for k, m in zip(("albedo", "normal", "pbr"), mips)):
        pytest.skip("the synthetic textures differ from the ones the digests were made with (other numpy / libm)")
    o = gb_oracles[case[1]]
    cam = mc.case_camera(case)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    m = o.generate_gbuffer(cam, g, _params(scene_tables))
    want = ref_digests["cases"][case[0]]
    for k in mc.PLANES:
        assert mc.sha(m[k]) == want[k], k
    assert want["hit_fraction"] > 0.3
    # the digests cover what they should: lamps are emissive, the grazing view reaches the small mip levels
    if case[0] == "gi_box_512x288_lamps":
        assert want["emissive_pixels"] > 1000


@pytest.mark.parametrize("pom", mc.POM_CASES, ids=[c[0] for c in mc.POM_CASES])
def test_relief_parallax_oracle_equals_the_reference_shader_digests(gb_oracles, scene_tables, mips, ref_digests, pom):
    """u_POM (ReliefParallax :153-199): dithered and fixed step counts, high quality, non-default depth / exponent, and u_POM keeping the
    pass alive when u_UpdateGBufferThisFrame is off."""
    if any(mc.sha(m) != ref_digests["mips"][k] for k, m in zip(("albedo", "normal", "pbr"), mips)):
        pytest.skip("the synthetic textures differ from the ones the digests were made with (other numpy / libm)")
    name, idx, kw = pom
    case = mc.CASES[idx]
    o = gb_oracles[case[1]]
    cam = mc.case_camera(case)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    m = o.generate_gbuffer(cam, g, _params(scene_tables, pom=True, **kw))
    for k in mc.PLANES:
        assert mc.sha(m[k]) == ref_digests["pom_cases"][name][k], k
    flat = o.generate_gbuffer(cam, g, _params(scene_tables))
    assert not np.array_equal(m["albedo"], flat["albedo"])        # the march does move the texture coordinates


@pytest.mark.parametrize("lava", mc.LAVA_CASES, ids=[c[0] for c in mc.LAVA_CASES])
def test_lava_path_oracle_equals_the_reference_shader_digests(gb_oracles, scene_tables, mips, ref_digests, lava):
    """u_LavaBlockID: time-driven UV distortion, the two animated 3-D textures, no parallax / bloom fix on liquid pixels — alone, with
    u_POM on the other pixels, and with u_UpdateGBufferThisFrame off (only the lava pixels are shaded, everything else keeps its texels)."""
    tex = mc.lava_textures()
    if any(mc.sha(m) != ref_digests["mips"][k] for k, m in zip(("albedo", "normal", "pbr"), mips)) or [mc.sha(t) for t in tex] != ref_digests["lava_textures"]:
        pytest.skip("the synthetic textures differ from the ones the digests were made with (other numpy / libm)")
    name, idx, block, kw = lava
    case = mc.CASES[idx]
    o = gb_oracles[case[1]]
    o.set_lava_textures(*tex)
    cam = mc.case_camera(case)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    seed = mc.seeded_planes(cam.width, cam.height)
    m = o.generate_gbuffer(cam, g, _params(scene_tables, lava_block_id=block, **kw), {k: v.copy() for k, v in seed.items()})
    for k in mc.PLANES:
        assert mc.sha(m[k]) == ref_digests["lava_cases"][name][k], k
    is_lava = g["block_id"] == block
    assert is_lava.sum() == ref_digests["lava_cases"][name]["lava_pixels"] > 100
    if not kw.get("update_this_frame", True):
        assert np.array_equal(m["pbr"][~is_lava], seed["pbr"][~is_lava]) and not np.array_equal(m["pbr"][is_lava], seed["pbr"][is_lava])
    else:   # the animation moves with u_Time
        m2 = o.generate_gbuffer(cam, g, _params(scene_tables, lava_block_id=block, **dict(kw, time=kw["time"] + 0.4)), {k: v.copy() for k, v in seed.items()})
        assert not np.array_equal(m2["albedo"][is_lava], m["albedo"][is_lava]) and np.array_equal(m2["albedo"][~is_lava], m["albedo"][~is_lava], equal_nan=True)


@needs_ref
def test_oracle_equals_the_reference_shader_live(gb_oracles, scene_tables, mips):
    """Frames not in the committed set: odd sizes (quads cut by the frame edge), a row slab, a roll of the camera, a frame that is
    mostly sky, and u_UpdateGBufferThisFrame = false (every fragment discards: the planes keep their contents)."""
    mats = scene_tables["materials"]
    o = gb_oracles["gi_box"]
    rng = np.random.RandomState(3)
    for W, H, pos, pitch, yaw, rows in [(133, 77, (192.0, 75.0, 192.0), -25.0, 33.0, None), (96, 54, (192.0, 75.0, 192.0), 35.0, 120.0, None),
                                        (160, 90, (140.0, 62.0, 16.0), 10.0, 60.0, (20, 64)), (64, 36, (192.0, 90.0, 192.0), -89.0, 0.0, None)]:
        cam = camera.FpsCamera(position=pos, pitch_deg=pitch, yaw_deg=yaw, aspect=W / H).vx_camera(W, H)
        g, _ = o.trace_primary(cam, vx.primary_params(350))
        if rows:
            cam.row_begin, cam.row_end = rows
        seed = {k: rng.rand(*s).astype(np.float32) for k, s in (("albedo", (H, W, 3)), ("normal", (H, W, 3)), ("pbr", (H, W, 4)), ("texture_ao", (H, W)))}
        a = o.generate_gbuffer(cam, g, _params(scene_tables), {k: v.copy() for k, v in seed.items()})
        b = ref_shaders.generate_gbuffer(cam, g, _params(scene_tables), mats, mips, {k: v.copy() for k, v in seed.items()})
        for k in mc.PLANES:
            assert np.array_equal(a[k], b[k]), (W, H, k)
        if rows:   # rows outside the slab are untouched
            assert np.array_equal(a["pbr"][:rows[0]], seed["pbr"][:rows[0]]) and np.array_equal(a["pbr"][rows[1]:], seed["pbr"][rows[1]:])
        off = _params(scene_tables, update_this_frame=False)
        a = o.generate_gbuffer(cam, g, off, {k: v.copy() for k, v in seed.items()})
        b = ref_shaders.generate_gbuffer(cam, g, off, mats, mips, {k: v.copy() for k, v in seed.items()})
        for k in mc.PLANES:
            assert np.array_equal(a[k], seed[k]) and np.array_equal(b[k], seed[k]), k


# ------------------------------------------------------------------------------------------------------ known answers / properties
def test_material_pass_known_answers(gb_oracles, scene_tables, mips):
    mats = scene_tables["materials"]
    table = mats["table"].reshape(6, 128)
    o = gb_oracles["superflat"]
    W, H = 64, 36
    # straight down onto the flat grass plain from 6 blocks up: every pixel hits the top face of Grass (block 1)
    cam = camera.FpsCamera(position=(192.3, 56.0, 192.6), pitch_deg=-89.9, aspect=W / H).vx_camera(W, H)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    assert (g["block_id"] == 1).all() and (g["normal_id"] == 2).all()
    m = o.generate_gbuffer(cam, g, _params(scene_tables))
    assert (m["pbr"] >= 0).all() and (m["pbr"] <= 1).all() and (m["texture_ao"] > 0).all() and (m["texture_ao"] <= 1).all()
    # the grass block's top face uses u_GrassBlockProps[1..3], not the table's (front-face) layers: swapping the table entries of block 1
    # changes nothing, swapping the top-face props does
    alt = dict(mats)
    t2 = table.copy()
    t2[0:3, 1] = (t2[0:3, 1] + 1) % mips[0].shape[0]
    o2 = vxo.Oracle(o.grid, o.df)
    alt["table"] = t2.reshape(768)
    o2.set_tables(alt, scene_tables["blue_noise"], scene_tables["sky"], scene_tables["shadow_noise"])
    o2.set_gbuffer_textures(*mips)
    m2 = o2.generate_gbuffer(cam, g, _params(scene_tables))
    assert all(np.array_equal(m[k], m2[k]) for k in mc.PLANES)
    gp = mats["grass_props"].copy()
    gp[1:4] = (gp[1:4] + 1) % mips[0].shape[0]
    m3 = o.generate_gbuffer(cam, g, vx.material_params(gp))
    assert not np.array_equal(m["albedo"], m3["albedo"])
    # texel known answer: pixel footprint here is 6 * tan(30 deg) * 2 / 36 blocks = 98.5 texels -> lambda in (6, 7): the albedo is a blend
    # of the nearest texels of mip levels 6 and 7 of the top-face layer, so it lies between their extremes
    layer = int(mats["grass_props"][1])
    off6 = sum((512 >> k) ** 2 for k in range(6))
    lv = mips[0][layer, off6:off6 + 64 + 16, :3].astype(np.float64) / 255.0
    lin = np.where(lv <= 0.04045, lv / 12.92, ((lv + 0.055) / 1.055) ** 2.4)
    assert (m["albedo"].reshape(-1, 3).min(0) >= lin.min(0) - 1e-6).all() and (m["albedo"].reshape(-1, 3).max(0) <= lin.max(0) + 1e-6).all()
    # a miss: black albedo, (1,1,1) normal, zero PBR / AO (:360-366)
    cam = camera.FpsCamera(position=(192.0, 100.0, 192.0), pitch_deg=60.0, aspect=W / H).vx_camera(W, H)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    assert (g["t"] < 0).all()
    m = o.generate_gbuffer(cam, g, _params(scene_tables))
    assert (m["albedo"] == 0).all() and (m["normal"] == 1).all() and (m["pbr"] == 0).all() and (m["texture_ao"] == 0).all()


def test_shading_normal_follows_the_face_basis(gb_oracles, scene_tables):
    """tbn * (2 n - 1): with the synthetic normal maps (z dominant) the mapped normal stays within 60 degrees of the face normal on every
    hit pixel; its length is at most 1 (mip levels average unit normals, which shortens them)."""
    o = gb_oracles["city"]
    case = mc.CASES[2]
    cam = mc.case_camera(case)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    m = o.generate_gbuffer(cam, g, _params(scene_tables))
    faces = np.array([[0, 0, 1], [0, 0, -1], [0, 1, 0], [0, -1, 0], [-1, 0, 0], [1, 0, 0]], np.float32)
    hit = g["t"] > 0
    n = m["normal"][hit]
    f = faces[g["normal_id"][hit]]
    ln = np.linalg.norm(n, axis=1)
    assert (ln > 0.3).all() and (ln < 1.02).all()
    assert ((n * f).sum(1) / ln > 0.5).all()
    assert len(np.unique(g["normal_id"][hit])) >= 3


def test_row_slabs_compose_to_the_whole_frame(gb_oracles, scene_tables):
    """The multi-GPU contract: even row slabs shade whole quads, so slab by slab equals the frame in one call."""
    o = gb_oracles["plains"]
    W, H = 160, 90
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    whole = o.generate_gbuffer(cam, g, _params(scene_tables))
    parts = None
    for rb, re in [(0, 24), (24, 58), (58, 90)]:
        cam.row_begin, cam.row_end = rb, re
        parts = o.generate_gbuffer(cam, g, _params(scene_tables), parts)
    for k in mc.PLANES:
        assert np.array_equal(whole[k], parts[k]), k


# ------------------------------------------------------------------------------------------------------ the kernel's source on the host
@pytest.mark.skipif(not koh.available(), reason="CUDA toolkit headers not present")
@pytest.mark.parametrize("case", mc.CASES[1:4], ids=[c[0] for c in mc.CASES[1:4]])
def test_kernel_source_on_host_equals_the_oracle(gb_oracles, scene_tables, case):
    o = gb_oracles[case[1]]
    k = koh.HostKernels(o, 1)
    cam = mc.case_camera(case)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    want = o.generate_gbuffer(cam, g, _params(scene_tables))
    got = k.generate_gbuffer(cam, g, _params(scene_tables))
    for key in mc.PLANES:
        assert np.array_equal(got[key], want[key]), key
    pom = _params(scene_tables, pom=True, frame=5, high_quality_pom=(case[0] == "city_480x270_street"))   # the relief-parallax instantiation
    want_pom, got_pom = o.generate_gbuffer(cam, g, pom), k.generate_gbuffer(cam, g, pom)
    for key in mc.PLANES:
        assert np.array_equal(got_pom[key], want_pom[key], equal_nan=True), ("pom", key)   # (NaN where the camera stands inside a block: t ~ 1e-4)
    o.set_lava_textures(*mc.lava_textures())                       # the lava instantiations (with and without u_POM, lava-only updates)
    k2 = koh.HostKernels(o, 1)
    block = int(np.bincount(g["block_id"][g["block_id"] > 0]).argmax())
    for kw in (dict(time=7.3), dict(time=0.25, pom=True, frame=2), dict(time=41.0, update_this_frame=False)):
        lp = _params(scene_tables, lava_block_id=block, **kw)
        seed = mc.seeded_planes(cam.width, cam.height)
        want_l = o.generate_gbuffer(cam, g, lp, {q: v.copy() for q, v in seed.items()})
        got_l = k2.generate_gbuffer(cam, g, lp, {q: v.copy() for q, v in seed.items()})
        for key in mc.PLANES:
            assert np.array_equal(got_l[key], want_l[key], equal_nan=True), ("lava", kw, key)
    k2.close()
    cam.row_begin, cam.row_end = 10, 52   # a slab: rows outside stay as they were
    seed = {key: np.full_like(want[key], 7.0) for key in mc.PLANES}
    got = k.generate_gbuffer(cam, g, _params(scene_tables), seed)
    for key in mc.PLANES:
        assert np.array_equal(got[key][10:52], want[key][10:52]) and (got[key][:10] == 7.0).all() and (got[key][52:] == 7.0).all(), key
    k.close()


# ------------------------------------------------------------------------------------------------------ GPU, through the C ABI
def _close(got, want, what):
    """bit-equal but for isolated texels whose mip blend weight differs by an ulp (CUDA's and glibc's double log2)"""
    diff = got != want
    assert diff.mean() <= 1e-4, (what, float(diff.mean()))
    assert float(np.abs(got.astype(np.float64) - want).max()) <= 2e-6, what


@pytest.fixture(scope="module")
def gb_renderer(renderer, mips):
    renderer.set_gbuffer_textures(*mips)
    return renderer


@pytest.mark.gpu
@pytest.mark.parametrize("case", mc.CASES, ids=[c[0] for c in mc.CASES])
def test_gpu_material_pass_equals_the_oracle(gb_renderer, worlds, gb_oracles, scene_tables, case):
    r, o = gb_renderer, gb_oracles[case[1]]
    r.upload_world(worlds[case[1]])
    r.build_distance_field()
    cam = mc.case_camera(case)
    W, H = cam.width, cam.height
    g = r.trace_primary(cam, vx.primary_params(350), r.alloc_gbuffer(W, H))
    g_ref, _ = o.trace_primary(cam, vx.primary_params(350))
    assert np.array_equal(g["inv_t"], g_ref["inv_t"], equal_nan=True) and np.array_equal(g["block_id"], g_ref["block_id"])
    want = o.generate_gbuffer(cam, g_ref, _params(scene_tables))
    got = r.generate_gbuffer(cam, g, _params(scene_tables), r.alloc_material(W, H))
    for k in mc.PLANES:
        _close(got[k], want[k], (case[0], k))
    assert r.launch_count() > 0


@pytest.mark.gpu
def test_gpu_material_pass_device_planes_slabs_and_reflection_chain(gb_renderer, worlds, gb_oracles, scene_tables):
    """Device-resident planes (torch tensors), the frame in two row slabs, and the pass feeding the reflection pass's g_normal / g_pbr."""
    import torch
    r, o = gb_renderer, gb_oracles["gi_box"]
    r.upload_world(worlds["gi_box"])
    r.build_distance_field()
    W, H = 192, 108
    fc = camera.FpsCamera(pitch_deg=-20.0)
    cam = fc.vx_camera(W, H)
    sun, moon, stronger, vis = (scene_tables[k] for k in ("sun", "moon", "stronger", "sun_visibility"))
    g = r.trace_primary(cam, vx.primary_params(350), r.alloc_gbuffer(W, H, device=True))
    m = r.alloc_material(W, H, device=True)
    for rb, re in [(0, 50), (50, 108)]:
        cam.row_begin, cam.row_end = rb, re
        r.generate_gbuffer(cam, g, _params(scene_tables), m)
    cam.row_begin, cam.row_end = 0, H
    d = r.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=4), r.alloc_diffuse(W, H, device=True))
    rp = vx.reflection_params(sun, moon, stronger, fc.position, scene_tables["materials"]["grass_props"], spp=2, rough=True, frame=4)
    refl = r.trace_reflection(cam, g, d, rp, r.alloc_reflection(W, H, device=True), g_normal=m["normal"], g_pbr=m["pbr"])
    r.sync()
    g_ref, _ = o.trace_primary(cam, vx.primary_params(350))
    m_ref = o.generate_gbuffer(cam, g_ref, _params(scene_tables))
    for k in mc.PLANES:
        _close(m[k].cpu().numpy(), m_ref[k], k)
    d_ref, _ = o.trace_diffuse(cam, g_ref, vx.diffuse_params(sun, moon, vis, spp=1, frame=4))
    refl_ref, _ = o.trace_reflection(cam, g_ref, d_ref, rp, g_normal=m_ref["normal"], g_pbr=m_ref["pbr"])
    col, col_ref = refl["color"].cpu().numpy().astype(np.float64), refl_ref["color"].astype(np.float64)
    assert float(np.mean(np.abs(col - col_ref))) <= 1e-3          # north_star radiance tolerance
    assert torch.isfinite(refl["color"]).all()


@pytest.mark.gpu
def test_gpu_material_pass_argument_checks(gb_renderer, worlds, scene_tables):
    r = gb_renderer
    r.upload_world(worlds["superflat"])
    r.build_distance_field()
    W, H = 64, 36
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    g = r.trace_primary(cam, vx.primary_params(350), r.alloc_gbuffer(W, H))
    out = r.alloc_material(W, H)
    for kw, code in [(dict(pom=True, pom_height=-1.0), abi.E_INVALID), (dict(lava_block_id=11), abi.E_STATE)]:   # no lava textures set
        with pytest.raises(abi.VxptError) as e:
            r.generate_gbuffer(cam, g, _params(scene_tables, **kw), out)
        assert e.value.code == code
    cam.row_begin = 3
    with pytest.raises(abi.VxptError) as e:
        r.generate_gbuffer(cam, g, _params(scene_tables), out)
    assert e.value.code == abi.E_INVALID
    cam.row_begin = 0
    seed = {k: np.full_like(v, 3.0) for k, v in out.items()}
    r.generate_gbuffer(cam, g, _params(scene_tables, update_this_frame=False), seed)   # discard: planes untouched
    assert all((v == 3.0).all() for v in seed.values())
    fresh = vx.Renderer(0)
    try:
        fresh.set_materials(scene_tables["materials"]["table"])
        fresh.upload_world(worlds["superflat"])
        fresh.build_distance_field()
        with pytest.raises(abi.VxptError) as e:
            fresh.generate_gbuffer(cam, g, _params(scene_tables), out)
        assert e.value.code == abi.E_STATE
    finally:
        fresh.close()
