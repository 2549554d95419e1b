"""Pins for the traversal oracle: analytic known answers for VoxelTraversalDF (InitialRayTraceFrag.glsl:307-374),
the reference's documented quirks (SURVEY.md A.3 notes), an independent plain-DDA cross-check
(after Core/Shaders/Implementations/DDA/DDA.glsl) and the survey's independent probe statistics."""
import math

import numpy as np
import pytest

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import camera


def _norm(v):
    v = np.asarray(v, dtype=np.float64)
    return (v / np.linalg.norm(v)).astype(np.float32)


def test_ray_down_onto_superflat_hits_top_face(oracles):
    o = oracles["superflat"]
    origin = (192.3, 75.2, 192.4)
    d = _norm((0.05, -1.0, 0.02))
    r = o.traverse(origin, d, 350)
    # ground top is y = 50; the reported point is nudged 1e-4 past the face (InitialRayTraceFrag.glsl:352)
    expect_t = (75.2 - (50.0 - 1e-4)) / abs(float(d[1]))
    assert r["min_idx"] == 1 and r["sgn"] == -1           # +Y face
    assert r["block"] == 1                                 # grass
    assert r["voxel"][1] == 49
    assert abs(r["t"] - expect_t) < 2e-3
    hx, hz = origin[0] + d[0] * r["t"], origin[2] + d[2] * r["t"]
    assert r["voxel"][0] == int(math.floor(hx)) and r["voxel"][2] == int(math.floor(hz))
    assert r["rays"] == 1 and r["vox_fetches"] == 1 and 3 <= r["df_fetches"] <= 40


def test_exactly_vertical_ray_works_but_zero_y_component_is_a_miss(oracles):
    o = oracles["superflat"]
    r = o.traverse((100.5, 90.5, 100.5), (0.0, -1.0, 0.0), 350)
    assert r["t"] == pytest.approx(40.5001, abs=1e-3) and r["voxel"] == (100, 49, 100)
    # A.3 note 2: sgn.y == 0 is not guarded in the min-axis select -> NaN position -> miss, even with a wall ahead
    wall = oracles["city"]
    r = wall.traverse((1.5, 60.5, 1.5), _norm((1.0, 0.0, 0.7)), 350)
    assert r["t"] == -1.0


def test_start_inside_solid_returns_miss(oracles):
    r = oracles["superflat"].traverse((10.5, 20.5, 10.5), _norm((0.3, 1.0, 0.2)), 350)
    assert r["t"] == -1.0 and r["df_fetches"] == 1  # E == 0 on iteration 0 with no intersection recorded (A.3 note 1)


def test_camera_outside_volume_sees_nothing(oracles):
    r = oracles["superflat"].traverse((192.0, 200.0, 192.0), _norm((0.0, -1.0, 0.01)), 350)
    assert r["t"] == -1.0 and r["df_fetches"] == 0  # no entry clipping (A.2)


def test_iteration_cap(oracles):
    o = oracles["superflat"]
    full = o.traverse((192.3, 120.2, 192.4), _norm((0.6, -0.25, 0.5)), 350)
    assert full["t"] > 0
    capped = o.traverse((192.3, 120.2, 192.4), _norm((0.6, -0.25, 0.5)), full["df_fetches"] - 2)
    assert capped["t"] == -1.0


def test_ray_leaving_through_the_top_is_a_miss(oracles):
    r = oracles["plains"].traverse((192.0, 75.0, 192.0), _norm((0.1, 1.0, 0.1)), 350)
    assert r["t"] == -1.0 and r["df_fetches"] < 10


@pytest.mark.parametrize("name", ["plains", "city", "gi_box"])
def test_hit_voxels_agree_with_plain_dda(oracles, worlds, name):
    """The DF-accelerated traversal must find the same first solid voxel as a plain voxel-by-voxel DDA
    (an independent formulation); grazing rays may differ by the 1e-4 nudges, so allow a small fraction."""
    o = oracles[name]
    grid = worlds[name].zyx
    rng = np.random.RandomState(3)
    n, agree, both = 4000, 0, 0
    for _ in range(n):
        while True:
            p = (rng.uniform(1, 383), rng.uniform(41, 127), rng.uniform(1, 383))
            if grid[int(p[2]), int(p[1]), int(p[0])] == 0:
                break
        d = _norm(rng.normal(size=3))
        r = o.traverse(p, d, 1000)
        hit, vox, axis = o.plain_dda(p, d, 4000)
        if r["t"] > 0:
            assert grid[r["voxel"][2], r["voxel"][1], r["voxel"][0]] == r["block"] > 0
        if r["t"] > 0 and hit:
            both += 1
            agree += int(tuple(vox) == tuple(r["voxel"]) and axis == r["min_idx"])
        elif (r["t"] > 0) == hit:
            agree += 0  # both miss: nothing to compare
    assert both > 500
    assert agree / both > 0.995, (agree, both)


def test_primary_statistics_match_the_survey_probe(oracles):
    """SURVEY.md Appendix C (an independent numba probe): superflat 640x360 from (192,75,192)."""
    o = oracles["superflat"]
    for pitch, hit_frac, mean_fetch in ((0.0, 0.386, 23.7), (-20.0, 0.694, 26.7)):
        cam = camera.FpsCamera(pitch_deg=pitch).vx_camera(640, 360)
        g, st = o.trace_primary(cam, vx.primary_params(350))
        assert abs((g["t"] > 0).mean() - hit_frac) < 0.002
        assert abs(st["df_fetches"] / st["rays"] - mean_fetch) < 0.15
        assert st["rays"] == 640 * 360


def test_primary_gbuffer_known_answers(oracles):
    o = oracles["superflat"]
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(640, 360)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    t = g["t"][180, 320]  # just off the image centre: ray ~ (0, -sin20, cos20) hits the y = 50 plane
    assert abs(t - 25.0 / math.sin(math.radians(20.0))) < 0.5
    assert g["normal_id"][180, 320] == 2 and g["block_id"][180, 320] == 1
    assert tuple(g["hit_voxel"][180, 320])[1] == 49
    # top rows look at the sky: miss encoding
    assert g["t"][359, 0] == -1.0 and g["normal_id"][359, 0] == vx.abi.NORMAL_MISS and g["block_id"][359, 0] == 0
    assert g["inv_t"][359, 0] == -1.0
    hit = g["t"] > 0
    assert np.all(g["normal_id"][hit] == 2) and np.all(g["block_id"][hit] == 1)
    assert np.allclose(g["inv_t"][hit], 1.0 / g["t"][hit], rtol=1e-6)


def test_jitter_moves_rays_by_less_than_a_pixel(oracles):
    o = oracles["superflat"]
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(640, 360)
    g0, _ = o.trace_primary(cam, vx.primary_params(350))
    g1, _ = o.trace_primary(cam, vx.primary_params(350, camera.taa_jitter(5)))
    both = (g0["t"] > 0) & (g1["t"] > 0)
    assert not np.array_equal(g0["t"], g1["t"])
    # one pixel of v changes t by at most a few percent at this pitch; a half-pixel jitter must stay well inside
    assert np.max(np.abs(g0["t"][both] - g1["t"][both]) / g0["t"][both]) < 0.05


def test_shadow_on_flat_ground_is_unshadowed(oracles, scene_tables):
    o = oracles["superflat"]
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(320, 180)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    s, st = o.trace_shadow(cam, g, vx.shadow_params(scene_tables["stronger"], frame=3, soft=True))
    hit = g["t"] > 0
    assert np.all(s["shadow"][hit] == 0) and np.all(s["shadow"][~hit] == 0)
    assert np.all(s["transversal"][~hit] == 64.0)                       # sky pixels (ShadowRayTraceFrag.glsl:426-430)
    assert np.allclose(s["transversal"][hit], 4.25 / 100.0)              # T < 0 (:509-511)
    assert st["rays"] == int(hit.sum()) and st["vox_fetches"] == int(hit.sum())  # one start-voxel test per traced pixel
    # light from below the horizon of the face: N.L <= 0.01 -> shadow 1 without tracing (:483-489)
    s2, st2 = o.trace_shadow(cam, g, vx.shadow_params((0.0, -1.0, 0.0), frame=3, soft=False))
    assert np.all(s2["shadow"][hit] == 1) and st2["rays"] == 0
    assert np.allclose(s2["transversal"][hit], 0.01)


def test_gi_on_flat_ground_under_a_uniform_sky(oracles, scene_tables):
    """Every hemisphere ray from flat ground escapes: radiance = sky * clamp(mix(1,1.05,vis)*GISky, 0, 5),
    AO = 1, sky-hit = 1 (DiffuseRayTraceFrag.glsl:625-634), and the SH projection follows :766-784."""
    from oracle import vxo
    o = vxo.Oracle(oracles["superflat"].grid, oracles["superflat"].df)
    sky_rgb = np.array([0.3, 0.5, 0.9], np.float32)
    sky = np.broadcast_to(sky_rgb, (6, 16, 16, 3)).copy()
    o.set_tables(scene_tables["materials"], scene_tables["blue_noise"], sky, scene_tables["shadow_noise"])
    cam = camera.FpsCamera(pitch_deg=-30.0).vx_camera(160, 90)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    vis = scene_tables["sun_visibility"]
    d, st = o.trace_diffuse(cam, g, vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], vis, spp=2, frame=9))
    hit = g["t"] > 0
    x = min(max((1.0 * (1 - vis) + 1.05 * vis) * 1.125, 0.0), 5.0)
    rad = sky_rgb.astype(np.float64) * x
    lum = 0.299 * rad[0] + 0.587 * rad[1] + 0.114 * rad[2]
    assert np.allclose(d["luma"][hit], lum, rtol=1e-5)
    assert np.all(d["ao_sky"][hit] == 1.0)
    co = rad[0] - rad[2]
    tt = rad[2] + co * 0.5
    cg = rad[1] - tt
    y = tt + cg * 0.5
    assert np.allclose(d["cocg"][hit], [co, cg], rtol=1e-5)
    assert np.allclose(d["sh"][hit][:, 3], 0.282095 * y, rtol=1e-5)
    # band-1 coefficients are 0.488603 * dir * Y with |dir| = 1 and dir.y > 0 for an upward hemisphere
    b1 = d["sh"][hit][:, :3] / (0.488603 * y)
    assert np.all(b1[:, 1] > 0) and np.all(np.linalg.norm(b1, axis=1) <= 1.0 + 1e-5)
    assert st["rays"] == 2 * int(hit.sum())
    # sky pixels: SH of 2.66 * sky projected on the (0.5,0.5,0.5) "normal" (:866-872), utility 0, ao/sky = (1, 0)
    assert np.all(d["luma"][~hit] == 0.0) and np.all(d["ao_sky"][~hit] == [1.0, 0.0])
    rad_s = sky_rgb.astype(np.float64) * 2.66
    co_s = rad_s[0] - rad_s[2]
    t_s = rad_s[2] + co_s * 0.5
    y_s = t_s + (rad_s[1] - t_s) * 0.5
    assert np.allclose(d["sh"][~hit], [0.488603 * 0.5 * y_s] * 3 + [0.282095 * y_s], rtol=1e-5)


def test_blue_noise_sampler_matches_a_direct_restatement(oracles, scene_tables):
    """(0.5 + (sobol[d + (i ^ rank[d + p*8]) * 256] ^ scramble[d % 8 + p*8])) / 256 — SURVEY.md A.5 — observed through the
    first GI direction of a pixel: with a uniform sky the band-1 SH gives back the sampled direction."""
    from oracle import vxo
    sobol, scramble, rank = scene_tables["blue_noise"]
    o = vxo.Oracle(oracles["superflat"].grid, oracles["superflat"].df)
    sky = np.full((6, 16, 16, 3), 0.5, np.float32)
    o.set_tables(scene_tables["materials"], scene_tables["blue_noise"], sky, scene_tables["shadow_noise"])
    cam = camera.FpsCamera(pitch_deg=-30.0).vx_camera(160, 90)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    frame = 21
    d, _ = o.trace_diffuse(cam, g, vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], 0.0, spp=1, frame=frame))
    jj, ii = np.nonzero(g["t"] > 0)
    for k in range(0, jj.size, max(jj.size // 40, 1)):
        i, j = int(ii[k]), int(jj[k])
        p = (i & 127) + (j & 127) * 128
        r = []
        for dim in (1, 2):
            ranked = (frame % 128) ^ int(rank[dim + p * 8])
            v = int(sobol[dim + ranked * 256]) ^ int(scramble[dim % 8 + p * 8])
            r.append((0.5 + v) / 256.0)
        # cosWeightedRandomHemisphereDirection on n = (0,1,0): uu = normalize(cross(n,(0,1,1))) = (1,0,0), vv = cross(uu,n) = (0,0,1)
        ra = math.sqrt(r[1])
        expect = np.array([ra * math.cos(2 * math.pi * r[0]), math.sqrt(1 - r[1]), ra * math.sin(2 * math.pi * r[0])])
        sh = d["sh"][j, i]
        got = sh[[0, 1, 2]] / 0.488603 / (sh[3] / 0.282095)  # (dir.x, dir.y, dir.z)
        assert np.allclose(got, expect, atol=2e-5), (i, j, got, expect)


def test_reflection_on_flat_ground_mirrors_the_sky(oracles, scene_tables):
    """Mirror reflections (u_RoughReflections = false) off flat ground leave the scene: colour = sky sample along
    reflect(I, N), alpha 1, hit distance -1 (ReflectionTraceFrag.glsl:1013-1037); sky pixels give (0, -1, 0) (:758-764)."""
    from oracle import vxo
    o = vxo.Oracle(oracles["superflat"].grid, oracles["superflat"].df)
    sky_rgb = np.array([0.3, 0.5, 0.9], np.float32)
    sky = np.broadcast_to(sky_rgb, (6, 16, 16, 3)).copy()
    mats = scene_tables["materials"]
    o.set_tables(mats, scene_tables["blue_noise"], sky, scene_tables["shadow_noise"])
    fc = camera.FpsCamera(pitch_deg=-30.0)
    cam = fc.vx_camera(160, 90)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    d, _ = o.trace_diffuse(cam, g, vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=1, frame=2))
    rp = vx.reflection_params(scene_tables["sun"], scene_tables["moon"], scene_tables["stronger"], fc.position, mats["grass_props"], spp=3, rough=False, frame=2)
    r, st = o.trace_reflection(cam, g, d, rp)
    hit = g["t"] > 0
    assert np.allclose(r["color"][hit], [0.3, 0.5, 0.9, 1.0], rtol=1e-6)       # grass metalness 0 -> no 1.175 boost
    assert np.all(r["hit_distance"][hit] == -1.0) and not r["emissive_mask"].any()
    assert np.all(r["color"][~hit] == 0.0) and np.all(r["hit_distance"][~hit] == -1.0)
    assert st["rays"] == 3 * int(hit.sum())                                      # one traversal per sample, no hits -> no shadow rays
    # rough reflections: the GGX-perturbed normal tips a few grazing rays back into the ground (those are shaded, and get
    # one shadow ray); everything else still sees the uniform sky
    rp2 = vx.reflection_params(scene_tables["sun"], scene_tables["moon"], scene_tables["stronger"], fc.position, mats["grass_props"], spp=2, rough=True, frame=2)
    r2, st2 = o.trace_reflection(cam, g, d, rp2)
    sky_only = r2["hit_distance"] == -1.0
    assert np.allclose(r2["color"][hit & sky_only], [0.3, 0.5, 0.9, 1.0], rtol=1e-6) and (hit & sky_only).sum() > 0.5 * hit.sum()
    n_ground = int((hit & ~sky_only).sum())
    assert 2 * int(hit.sum()) + 1 <= st2["rays"] <= 2 * int(hit.sum()) + n_ground and n_ground > 0


def test_reflection_hits_shade_with_gi_ambient_and_sun(oracles, scene_tables):
    """In the city nearly every reflection ray hits: hit distance is the mean T of the hitting samples, alpha is 1, emissive
    lamps raise the mask, and shadow rays are cast for the first max(SPP/4, 1) hitting samples only."""
    o = oracles["city"]
    mats = scene_tables["materials"]
    fc = camera.FpsCamera(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0)
    cam = fc.vx_camera(160, 90)
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    d, _ = o.trace_diffuse(cam, g, vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=1, frame=2))
    rp = vx.reflection_params(scene_tables["sun"], scene_tables["moon"], scene_tables["stronger"], fc.position, mats["grass_props"], spp=4, frame=2)
    r, st = o.trace_reflection(cam, g, d, rp)
    hit = g["t"] > 0
    assert hit.mean() > 0.5
    refl_hit = r["hit_distance"] > 0
    assert refl_hit[hit].mean() > 0.3 and np.all(r["hit_distance"][refl_hit] <= 200.0)
    assert np.all(r["color"][hit][:, 3] == 1.0) and np.all(r["color"][hit] >= 1e-7) and np.all(r["color"] <= 100.0)
    assert not np.isnan(r["color"]).any()
    n_px = int(hit.sum())
    assert 4 * n_px <= st["rays"] <= 5 * n_px        # 4 reflection rays + at most max(4/4, 1) = 1 shadow ray per pixel
