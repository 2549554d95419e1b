"""The oracle pinned by the reference ITSELF: the reference's own shader source — Core/Shaders/ManhattanDistance{X,Y,Z}.comp (distance
field), InitialRayTraceFrag.glsl (ray set-up, VoxelTraversalDF, G-buffer outputs), ShadowRayTraceFrag.glsl (soft sun shadows) and
DiffuseRayTraceFrag.glsl (1-bounce diffuse GI, blue-noise sampler, SH projection) — is compiled as C++ against the reference's
vendored glm (oracle/Makefile -> oracle/_ref/libref_shaders.so; oracle/glsl2cpp.py rewrites declarations only) and run on
the CPU.  Its outputs on the BASELINE worlds / frames are committed as digests (tests/golden/ref_shader_digests.json, made by
tools/make_ref_shader_golden.py) and must equal the oracle's, bit for bit; when the library is present the two are also compared
live on edge cases."""
import json
import os

import numpy as np
import pytest

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import abi, camera, world
from oracle import ref_shaders, vxo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not ref_shaders.available(), reason="oracle/_ref/libref_shaders.so not built (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def ref_digests():
    with open(os.path.join(ROOT, "tests", "golden", "ref_shader_digests.json")) as f:
        return json.load(f)


def test_committed_reference_shader_digests_equal_the_oracle_digests(ref_digests, golden_digests):
    """Golden vectors produced by the reference's shaders == what the oracle produced for the same inputs (5 distance fields,
    14 primary frames up to 3840x2160: t, face id and block id planes)."""
    assert set(ref_digests["df"]) == {"superflat", "plains", "gi_box", "city", "sparse"}
    for name, digest in ref_digests["df"].items():
        assert ref_digests["world"][name] == golden_digests["world"][name]
        assert digest == golden_digests["df"][name], name
    assert len(ref_digests["primary"]) >= 14
    for case, planes in ref_digests["primary"].items():
        for k in ("t", "normal_id", "block_id"):
            assert planes[k] == golden_digests["primary"][case][k], (case, k)
    cases_1080p = {"plains_1920x1080_p-20_jNone", "city_1920x1080_p-20_jNone"}
    assert cases_1080p <= set(ref_digests["shadow"]) and cases_1080p <= set(ref_digests["diffuse"])
    for case in cases_1080p:                                   # 1080p soft sun shadows, 1-spp diffuse GI (the SH plane), bit for bit
        for k in ("shadow", "transversal"):
            assert ref_digests["shadow"][case][k] == golden_digests["shadow"][case][k], (case, k)
        assert ref_digests["diffuse"][case]["sh"] == golden_digests["diffuse"][case]["sh"], case


@pytest.mark.parametrize("name", ["superflat", "sparse"])
def test_oracle_distance_field_reproduces_the_committed_reference_digest(worlds, oracle_dfs, ref_digests, name):
    import hashlib
    assert hashlib.sha256(oracle_dfs[name].tobytes()).hexdigest() == ref_digests["df"][name]


@needs_ref
def test_distance_field_edge_worlds_live():
    cases = {"empty": world.World(), "full": world.World(np.full(abi.WORLD_VOXELS, 3, np.uint8))}
    for corner in [(0, 0, 0), (383, 127, 383), (191, 64, 200)]:
        w = world.World()
        w.set_block(*corner, 9)
        cases[f"voxel{corner}"] = w
    rng = np.random.RandomState(11)
    w = world.World()
    w.data[rng.randint(0, w.data.size, size=3000)] = rng.randint(1, 128, size=3000)
    w.zyx[100:140, 30:90, 37:300] = 7          # a slab, so that clamped and unclamped regions both exist
    cases["random+slab"] = w
    for name, w in cases.items():
        assert np.array_equal(ref_shaders.df_build(w.data), vxo.df_build(w.data)), name


def _same(a, b):
    if a.dtype.kind == "f":
        return bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))
    return bool(np.array_equal(a, b))


@needs_ref
@pytest.mark.parametrize("name,cam_kw,max_it,jf", [
    ("plains", dict(pitch_deg=-20.0), 350, 3),                                         # config 2 view, TAA jitter
    ("plains", dict(pitch_deg=-89.9), 350, None),                                      # nearly straight down
    ("plains", dict(position=(192.0, 75.0, 192.0), pitch_deg=0.0, yaw_deg=90.0), 350, None),   # axis-aligned rays: zero direction components
    ("plains", dict(position=(-40.0, 140.0, 500.0), pitch_deg=-25.0, yaw_deg=-60.0), 475, 9),  # camera outside the volume
    ("city", dict(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0), 350, None),  # dense geometry
    ("city", dict(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0), 12, None),   # iteration cap cuts most rays short
    ("gi_box", dict(pitch_deg=10.0), 64, 40),                                          # looking up past the rooms, short cap
    ("superflat", dict(position=(192.0, 20.0, 192.0), pitch_deg=30.0), 350, None),     # camera inside solid rock
])
def test_primary_pass_live(worlds, oracle_dfs, oracles, name, cam_kw, max_it, jf):
    """Oracle G-buffer == the reference shader's, bit for bit: hit distance, 1/t, face id, block id."""
    W, H = 224, 126
    cam = camera.FpsCamera(aspect=W / H, **cam_kw).vx_camera(W, H)
    pp = vx.primary_params(max_it, None if jf is None else camera.taa_jitter(jf))
    ref = ref_shaders.trace_primary(worlds[name].data, oracle_dfs[name], cam, pp)
    got, _ = oracles[name].trace_primary(cam, pp)
    for k in ("t", "inv_t", "normal_id", "block_id"):
        assert _same(got[k], ref[k]), (k, int(np.sum(got[k] != ref[k])))


@needs_ref
@pytest.mark.parametrize("name,frame,soft", [("plains", 5, True), ("plains", 1023, True), ("city", 40, False), ("gi_box", 2048, True)])
def test_shadow_pass_live(worlds, oracle_dfs, oracles, scene_tables, name, frame, soft):
    """Oracle shadow planes == ShadowRayTraceFrag.glsl's, bit for bit (cone jitter from the blue-noise texture, N.L cull, start-voxel
    test, 350-iteration traversal, transversal output)."""
    W, H = 224, 126
    fc = camera.FpsCamera(pitch_deg=-20.0, aspect=W / H) if name != "city" else camera.FpsCamera(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0, aspect=W / H)
    cam = fc.vx_camera(W, H)
    g, _ = oracles[name].trace_primary(cam, vx.primary_params(350, camera.taa_jitter(frame)))
    sp = vx.shadow_params(scene_tables["stronger"], frame=frame, soft=soft)
    ref = ref_shaders.trace_shadow(worlds[name].data, oracle_dfs[name], cam, g, sp, scene_tables["shadow_noise"])
    got, _ = oracles[name].trace_shadow(cam, g, sp)
    assert np.array_equal(got["shadow"], ref["shadow"])
    assert _same(got["transversal"], ref["transversal"])


@needs_ref
@pytest.mark.parametrize("name,spp,frame,checker,tick", [
    ("plains", 1, 7, False, 50.0),       # config 3: 1 spp
    ("plains", 4, 3, False, 50.0),       # config 4's sample count: the blue-noise dimension counter runs on across samples
    ("gi_box", 2, 130, True, 50.0),      # checkerboard SPP, frame > 128 (u_CurrentFrameMod128), emissive lamps
    ("city", 3, 9, False, 140.0),        # night: the moon is the stronger light (no shadow sub-rays, doubled sample count)
    ("city", 16, 1, False, 50.0),        # many samples: the sampler indexes past rankingTile (pinned by clamping)
])
def test_diffuse_gi_pass_live(worlds, oracle_dfs, oracles, scene_tables, name, spp, frame, checker, tick):
    """Oracle GI planes == DiffuseRayTraceFrag.glsl's, bit for bit: SH, CoCg, luminance, AO / sky visibility."""
    W, H = 160, 90
    sun, moon, stronger, vis = camera.sun_moon_direction(tick)
    fc = camera.FpsCamera(pitch_deg=-20.0, aspect=W / H) if name != "city" else camera.FpsCamera(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0, aspect=W / H)
    cam = fc.vx_camera(W, H)
    g, _ = oracles[name].trace_primary(cam, vx.primary_params(350))
    dp = vx.diffuse_params(sun, moon, vis, spp=spp, frame=frame, checkerboard=checker)
    ref = ref_shaders.trace_diffuse(worlds[name].data, oracle_dfs[name], cam, g, dp, scene_tables["materials"], scene_tables["blue_noise"], scene_tables["sky"])
    got, _ = oracles[name].trace_diffuse(cam, g, dp)
    for k in ("sh", "cocg", "luma", "ao_sky"):
        assert _same(got[k], ref[k]), (k, int(np.sum(got[k] != ref[k])), float(np.max(np.abs(got[k].astype(np.float64) - ref[k]))))


def _reflection_inputs(oracle, scene_tables, case):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_ref_shader_golden import synthetic_material_planes
    cname, wname, W, H, cam_kw, spp, rough, checker, frame = case
    sun, moon, stronger, vis = scene_tables["sun"], scene_tables["moon"], scene_tables["stronger"], scene_tables["sun_visibility"]
    fc = camera.FpsCamera(**cam_kw)
    cam = fc.vx_camera(W, H)
    g, _ = oracle.trace_primary(cam, vx.primary_params(350))
    d, _ = oracle.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=max(frame, 0)))
    g_normal, g_pbr = synthetic_material_planes(g, W, H)
    rp = vx.reflection_params(sun, moon, stronger, fc.position, scene_tables["materials"]["grass_props"], spp=spp, rough=rough, checkerboard=checker, frame=frame,
                              halton=camera.taa_jitter_secondary(max(frame, 0)))
    return cam, g, d, rp, g_normal, g_pbr


def test_oracle_reflections_reproduce_the_committed_reference_digests(oracles, scene_tables, ref_digests):
    """ReflectionTraceFrag.glsl (v1 parity profile, u_Halton = GetTAAJitterSecondary(frame): the primary distance filtered GL_LINEAR and the
    normal id GL_NEAREST at the jittered coordinate) as compiled from the reference vs the oracle: colour, hit distance and emissive mask
    planes, bit for bit, on four frames (config 3 at 1920x1080, rough GGX-sampled, checkerboard SPP on dense geometry, mirror with lamps)."""
    import hashlib
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_ref_shader_golden import reflection_cases
    for case in reflection_cases():
        cam, g, d, rp, g_normal, g_pbr = _reflection_inputs(oracles[case[1]], scene_tables, case)
        got, _ = oracles[case[1]].trace_reflection(cam, g, d, rp, g_normal, g_pbr)
        want = ref_digests["reflection"][case[0]]
        for k in ("color", "hit_distance", "emissive_mask"):
            assert hashlib.sha256(np.ascontiguousarray(got[k]).tobytes()).hexdigest() == want[k], (case[0], k)


@needs_ref
def test_reflection_pass_live_night(worlds, oracle_dfs, oracles, scene_tables):
    """Moon as the stronger light (SampleMoonColor path), 1 spp, small frame."""
    W, H = 160, 90
    sun, moon, stronger, vis = camera.sun_moon_direction(140.0)
    fc = camera.FpsCamera(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0, aspect=W / H)
    cam = fc.vx_camera(W, H)
    o = oracles["city"]
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    d, _ = o.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=3))
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from make_ref_shader_golden import synthetic_material_planes
    g_normal, g_pbr = synthetic_material_planes(g, W, H)
    rp = vx.reflection_params(sun, moon, stronger, fc.position, scene_tables["materials"]["grass_props"], spp=1, rough=True, frame=3,
                              halton=camera.taa_jitter_secondary(3))
    ref = ref_shaders.trace_reflection(worlds["city"].data, oracle_dfs["city"], cam, g, d, rp, g_normal, g_pbr, scene_tables["materials"],
                                       scene_tables["blue_noise"], scene_tables["sky"])
    got, _ = o.trace_reflection(cam, g, d, rp, g_normal, g_pbr)
    for k in ("color", "hit_distance", "emissive_mask"):
        assert _same(got[k], ref[k]), (k, int(np.sum(got[k] != ref[k])))


def test_oracle_config4_frame_reproduces_the_committed_reference_digests(oracles, scene_tables, ref_digests):
    """BASELINE config 4 (3840x2160, 4-spp GI on the gi-box scene): the oracle's shadow and GI planes == digests of the reference's shaders."""
    import hashlib
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()  # noqa: E731
    o = oracles["gi_box"]
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(3840, 2160)
    g, _ = o.trace_primary(cam, vx.primary_params(350), hit_voxel=False)
    s, _ = o.trace_shadow(cam, g, vx.shadow_params(scene_tables["stronger"], frame=9, soft=True))
    d, _ = o.trace_diffuse(cam, g, vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=4, frame=9))
    for k in ("shadow", "transversal"):
        assert sha(s[k]) == ref_digests["shadow"]["gi_box_3840x2160_f9"][k], k
    for k in ("sh", "cocg", "luma", "ao_sky"):
        assert sha(d[k]) == ref_digests["diffuse"]["gi_box_3840x2160_spp4_f9"][k], k
