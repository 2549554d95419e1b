"""Parity tests proper: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs, against the
committed digests, and — at BASELINE.json's full sizes — through size-independent properties.
Bar (north_star): distance field, hit voxel, face and block ids bit-exact; hit distance <= 1e-5 relative;
GI / shadow radiance <= 1e-3 mean absolute error and >= 50 dB PSNR."""
import ctypes as C
import hashlib

import numpy as np
import pytest

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import abi, camera, world

pytestmark = pytest.mark.gpu

HIT_DISTANCE_RTOL = 1e-5   # north_star tolerance for t
RADIANCE_MAE = 1e-3        # north_star tolerance for GI / shadow radiance
RADIANCE_PSNR_DB = 50.0


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def psnr(a, b, peak):
    mse = float(np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))
    return 200.0 if mse == 0 else 10.0 * np.log10(peak * peak / mse)


def load(renderer, w, algo=1):
    renderer.set_option(abi.OPT_DF_ALGO, algo)
    renderer.upload_world(w)
    renderer.build_distance_field()


# ====================================================================================== distance field
@pytest.mark.parametrize("algo", [0, 1])
@pytest.mark.parametrize("name", ["superflat", "plains", "city", "gi_box", "sparse", "empty"])
def test_distance_field_bit_exact(renderer, worlds, oracle_dfs, golden_digests, name, algo):
    load(renderer, worlds[name], algo)
    df = renderer.download_distance_field()
    assert np.array_equal(df, oracle_dfs[name])
    if name in golden_digests["df"]:
        assert sha(df) == golden_digests["df"][name]
    assert np.array_equal(renderer.download_world(), worlds[name].data)


def test_distance_field_edge_worlds(renderer):
    from oracle import vxo
    full = world.World(np.full(abi.WORLD_VOXELS, 3, np.uint8))
    load(renderer, full)
    assert not renderer.download_distance_field().any()
    for corner in [(0, 0, 0), (383, 127, 383), (383, 0, 0), (0, 127, 383), (191, 64, 200)]:
        w = world.World()
        w.set_block(*corner, 9)
        load(renderer, w)
        assert np.array_equal(renderer.download_distance_field(), vxo.df_build(w.data)), corner
    # a single air pocket inside solid rock, and alternating slabs that exercise every segment carry
    w = world.World(np.full(abi.WORLD_VOXELS, 3, np.uint8))
    w.zyx[100:140, 30:90, 37:300] = 0
    w.zyx[::48, :, :] = 0
    load(renderer, w)
    assert np.array_equal(renderer.download_distance_field(), vxo.df_build(w.data))


def test_distance_field_random_worlds(renderer):
    from oracle import vxo
    rng = np.random.RandomState(42)
    for fill in (1e-6, 1e-4, 0.01, 0.5):
        w = world.World((rng.rand(abi.WORLD_VOXELS) < fill).astype(np.uint8) * 5)
        load(renderer, w)
        assert np.array_equal(renderer.download_distance_field(), vxo.df_build(w.data)), fill


def test_distance_field_properties_at_full_size(renderer, worlds):
    """Size-independent: DF == 0 exactly on solid voxels, and the field is 1-Lipschitz along every axis."""
    load(renderer, worlds["city"])
    df = renderer.download_distance_field().reshape(384, 128, 384).astype(np.int16)
    assert np.array_equal(df == 0, worlds["city"].zyx > 0)
    for axis in range(3):
        assert np.abs(np.diff(df, axis=axis)).max() <= 1
    # idempotent rebuild
    renderer.build_distance_field()
    assert np.array_equal(renderer.download_distance_field().reshape(384, 128, 384), df)


def test_block_edits_force_a_rebuild(renderer, worlds):
    from oracle import vxo
    w = world.World(worlds["plains"].data.copy())
    load(renderer, w)
    cam = camera.FpsCamera(pitch_deg=-20).vx_camera(64, 36)
    g = renderer.alloc_gbuffer(64, 36)
    renderer.trace_primary(cam, vx.primary_params(350), g)
    renderer.set_block(192, 70, 200, world.STONE)          # World::Raycast place path -> glTexSubImage3D (World.cpp:372-373)
    with pytest.raises(abi.VxptError) as e:                # the ABI never rebuilds implicitly
        renderer.trace_primary(cam, vx.primary_params(350), g)
    assert e.value.code == abi.E_STATE
    w.set_block(192, 70, 200, world.STONE)
    renderer.build_distance_field()
    assert np.array_equal(renderer.download_distance_field(), vxo.df_build(w.data))
    xyz = np.array([[10, 60, 10], [11, 60, 10], [383, 127, 383], [192, 70, 200]], np.int16)
    ids = np.array([4, 4, 12, 0], np.uint8)
    renderer.set_blocks(xyz, ids)
    for (x, y, z), b in zip(xyz, ids):
        w.set_block(int(x), int(y), int(z), int(b))
    renderer.build_distance_field()
    assert np.array_equal(renderer.download_world(), w.data)
    assert np.array_equal(renderer.download_distance_field(), vxo.df_build(w.data))
    with pytest.raises(abi.VxptError) as e:
        renderer.set_block(384, 0, 0, 1)
    assert e.value.code == abi.E_INVALID


# ====================================================================================== primary rays
def check_primary(renderer, oracle, cam, pp, layout=1, digest=None):
    renderer.set_option(abi.OPT_TRAVERSAL_LAYOUT, layout)
    renderer.reset_stats()
    g = renderer.alloc_gbuffer(cam.width, cam.height, hit_voxel=True)
    renderer.trace_primary(cam, pp, g)
    st = renderer.stats()
    if oracle is not None:
        ref, rst = oracle.trace_primary(cam, pp)
        for k in ("normal_id", "block_id", "hit_voxel"):
            assert np.array_equal(g[k], ref[k]), k                                   # ids: bit-exact
        assert np.array_equal(g["t"] > 0, ref["t"] > 0)
        hit = ref["t"] > 0
        assert np.all(np.abs(g["t"][hit] - ref["t"][hit]) <= HIT_DISTANCE_RTOL * ref["t"][hit])
        assert np.array_equal(g["t"], ref["t"]) and np.array_equal(g["inv_t"], ref["inv_t"])  # in practice bit-exact
        assert (st["rays"], st["df_fetches"], st["vox_fetches"]) == (rst["rays"], rst["df_fetches"], rst["vox_fetches"])
    if digest is not None:
        assert sha(g["t"]) == digest["t"] and sha(g["normal_id"]) == digest["normal_id"]
        assert sha(g["block_id"]) == digest["block_id"] and sha(g["hit_voxel"]) == digest["hit_voxel"]
        assert st["df_fetches"] == digest["stats"]["df_fetches"] and st["vox_fetches"] == digest["stats"]["vox_fetches"]
    return g


@pytest.mark.parametrize("pitch", [0.0, -20.0])
@pytest.mark.parametrize("jf", [None, 0, 1, 17, 63])
def test_primary_config1_superflat_640x360(renderer, worlds, oracles, golden_digests, pitch, jf):
    load(renderer, worlds["superflat"])
    cam = camera.FpsCamera(pitch_deg=pitch).vx_camera(640, 360)
    pp = vx.primary_params(350, None if jf is None else camera.taa_jitter(jf))
    d = golden_digests["primary"][f"superflat_640x360_p{int(pitch)}_j{jf}"]
    for layout in (0, 1):
        check_primary(renderer, oracles["superflat"], cam, pp, layout, d)


@pytest.mark.parametrize("case", ["plains_1920x1080_p-20_jNone", "plains_1920x1080_p0_j7", "city_1920x1080_p-20_jNone"])
def test_primary_1080p(renderer, worlds, oracles, golden_digests, case):
    name = case.split("_")[0]
    load(renderer, worlds[name])
    pitch = -20.0 if "p-20" in case else 0.0
    jf = 7 if case.endswith("j7") else None
    cam = camera.FpsCamera(pitch_deg=pitch).vx_camera(1920, 1080)
    pp = vx.primary_params(350, None if jf is None else camera.taa_jitter(jf))
    for layout in (0, 1):
        check_primary(renderer, oracles[name], cam, pp, layout, golden_digests["primary"][case])


def test_primary_4k_digest_and_properties(renderer, worlds, golden_digests):
    """Config 4 frame size (3840x2160) without CPU tracing: committed digest + 'the hit voxel is solid' for every pixel
    and 'the voxel the ray came from (hit + face normal) is air' for all but grazing rays (the reference keeps a stale
    face after a skip step, InitialRayTraceFrag.glsl:331-334: ~1e-5 of the hits on these worlds)."""
    load(renderer, worlds["gi_box"])
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(3840, 2160)
    g = check_primary(renderer, None, cam, vx.primary_params(350), 1, golden_digests["primary"]["gi_box_3840x2160_p-20_jNone"])
    grid = worlds["gi_box"].zyx
    hit = g["t"] > 0
    v = g["hit_voxel"][hit].astype(np.int64)
    assert np.all(grid[v[:, 2], v[:, 1], v[:, 0]] == g["block_id"][hit]) and np.all(g["block_id"][hit] > 0)
    normals = np.array([[0, 0, 1], [0, 0, -1], [0, 1, 0], [0, -1, 0], [-1, 0, 0], [1, 0, 0]])
    prev = v + normals[g["normal_id"][hit]]
    inside = np.all((prev >= 0) & (prev < [384, 128, 384]), axis=1)
    assert np.mean(grid[prev[inside, 2], prev[inside, 1], prev[inside, 0]] != 0) < 1e-4
    assert np.all(g["normal_id"][~hit] == abi.NORMAL_MISS) and np.all(g["t"][~hit] == -1.0)


def test_primary_edge_cases(renderer, worlds, oracles):
    load(renderer, worlds["plains"])
    o = oracles["plains"]
    for (w, h) in [(1, 1), (33, 17), (257, 3)]:                       # ragged tiles
        cam = camera.FpsCamera(pitch_deg=-35.0, aspect=w / h).vx_camera(w, h)
        check_primary(renderer, o, cam, vx.primary_params(350))
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(320, 180)
    for cap in (0, 1, 10, 475):                                       # iteration caps (475 = first-frame u_RenderDistance)
        check_primary(renderer, o, cam, vx.primary_params(cap))
    for pos, yaw, pitch in [((192, 300, 192), 90, -60), ((-50, 75, 192), 0, 0), ((192.5, 20.5, 192.5), 90, 10), ((5, 127.5, 5), 45, -5)]:
        cam = camera.FpsCamera(position=pos, yaw_deg=yaw, pitch_deg=pitch).vx_camera(160, 90)   # outside / inside rock / at the rim
        check_primary(renderer, o, cam, vx.primary_params(350))
    load(renderer, worlds["empty"])
    g = check_primary(renderer, oracles["empty"], cam, vx.primary_params(350))
    assert np.all(g["t"] == -1.0)


def test_row_slabs_compose_to_the_full_frame(renderer, worlds, scene_tables):
    """The multi-GPU contract: a call touches only rows [row_begin,row_end) and slabs rendered separately equal the
    full frame bit for bit (primary, shadow, GI)."""
    load(renderer, worlds["plains"])
    W, H = 320, 180
    fc = camera.FpsCamera(pitch_deg=-20.0)
    pp = vx.primary_params(350, camera.taa_jitter(2))
    sp = vx.shadow_params(scene_tables["stronger"], frame=4)
    dp = vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=2, checkerboard=True, frame=4)
    full_cam = fc.vx_camera(W, H)
    g = renderer.trace_primary(full_cam, pp, renderer.alloc_gbuffer(W, H, hit_voxel=True))
    s = renderer.trace_shadow(full_cam, g, sp, renderer.alloc_shadow(W, H))
    d = renderer.trace_diffuse(full_cam, g, dp, renderer.alloc_diffuse(W, H))
    g2 = {k: np.full_like(v, 77) for k, v in g.items()}
    s2 = {k: np.full_like(v, 77) for k, v in s.items()}
    d2 = {k: np.full_like(v, 77) for k, v in d.items()}
    bounds = [0, 7, 8, 100, 180]
    for b, e in zip(bounds[:-1], bounds[1:]):
        cam = fc.vx_camera(W, H, b, e)
        before = {k: v.copy() for k, v in g2.items()}
        renderer.trace_primary(cam, pp, g2)
        for k in g2:                                                   # rows outside the slab are untouched
            assert np.array_equal(g2[k][:b], before[k][:b]) and np.array_equal(g2[k][e:], before[k][e:])
        renderer.trace_shadow(cam, g2, sp, s2)
        renderer.trace_diffuse(cam, g2, dp, d2)
    for a, b_ in ((g, g2), (s, s2), (d, d2)):
        for k in a:
            assert np.array_equal(a[k], b_[k]), k


def test_device_buffers_equal_host_buffers(renderer, worlds, scene_tables):
    import torch
    load(renderer, worlds["plains"])
    W, H = 256, 144
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    pp = vx.primary_params(350)
    gh = renderer.trace_primary(cam, pp, renderer.alloc_gbuffer(W, H))
    gd = renderer.trace_primary(cam, pp, renderer.alloc_gbuffer(W, H, device=True))
    sp = vx.shadow_params(scene_tables["stronger"], frame=1)
    sh = renderer.trace_shadow(cam, gh, sp, renderer.alloc_shadow(W, H))
    sd = renderer.trace_shadow(cam, gd, sp, renderer.alloc_shadow(W, H, device=True))
    dp = vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=1, frame=1)
    dh = renderer.trace_diffuse(cam, gh, dp, renderer.alloc_diffuse(W, H))
    dd = renderer.trace_diffuse(cam, gd, dp, renderer.alloc_diffuse(W, H, device=True))
    renderer.sync()
    for host, dev in ((gh, gd), (sh, sd), (dh, dd)):
        for k in host:
            assert isinstance(dev[k], torch.Tensor) and dev[k].is_cuda
            assert np.array_equal(host[k], dev[k].cpu().numpy()), k


# ====================================================================================== sun shadow
@pytest.mark.parametrize("name,soft,frame", [("plains", True, 5), ("plains", False, 0), ("city", True, 1023), ("gi_box", True, 77)])
def test_shadow_parity(renderer, worlds, oracles, scene_tables, golden_digests, name, soft, frame):
    load(renderer, worlds[name])
    W, H = (1920, 1080) if name != "gi_box" else (960, 540)
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    g, _ = oracles[name].trace_primary(cam, vx.primary_params(350))
    sp = vx.shadow_params(scene_tables["stronger"], frame=frame, soft=soft)
    ref, rst = oracles[name].trace_shadow(cam, g, sp)
    for layout in (0, 1):
        renderer.set_option(abi.OPT_TRAVERSAL_LAYOUT, layout)
        renderer.reset_stats()
        out = renderer.trace_shadow(cam, g, sp, renderer.alloc_shadow(W, H))
        st = renderer.stats()
        # the cone jitter goes through sin/cos (pinned: correctly rounded fp32); allow a vanishing fraction of flips
        flips = np.mean(out["shadow"] != ref["shadow"])
        assert flips <= 1e-5, flips
        assert np.mean(np.abs(out["shadow"].astype(np.float64) - ref["shadow"])) <= RADIANCE_MAE
        assert psnr(out["shadow"], ref["shadow"], 1.0) >= RADIANCE_PSNR_DB
        same = out["shadow"] == ref["shadow"]
        assert np.allclose(out["transversal"][same], ref["transversal"][same], rtol=HIT_DISTANCE_RTOL, atol=0)
        assert st["rays"] == rst["rays"] and abs(st["df_fetches"] - rst["df_fetches"]) <= 1e-5 * rst["df_fetches"]
    if soft and frame == 5 and name == "plains":
        d = golden_digests["shadow"]["plains_1920x1080_p-20_jNone"]
        assert sha(out["shadow"]) == d["shadow"] and st["rays"] == d["stats"]["rays"]


# ====================================================================================== diffuse GI
def check_diffuse(renderer, oracle, cam, g, dp):
    ref, rst = oracle.trace_diffuse(cam, g, dp)
    renderer.reset_stats()
    out = renderer.trace_diffuse(cam, g, dp, renderer.alloc_diffuse(cam.width, cam.height))
    st = renderer.stats()
    for k, peak in (("sh", 8.0), ("cocg", 8.0), ("luma", 8.0), ("ao_sky", 1.0)):   # radiance samples are clamped to [0, 8]
        mae = float(np.mean(np.abs(out[k].astype(np.float64) - ref[k])))
        assert mae <= RADIANCE_MAE, (k, mae)
        assert psnr(out[k], ref[k], peak) >= RADIANCE_PSNR_DB, k
        assert np.mean(out[k] != ref[k]) <= 1e-4, (k, float(np.mean(out[k] != ref[k])))   # in practice bit-exact
    assert abs(st["rays"] - rst["rays"]) <= 1e-5 * rst["rays"] + 2
    assert abs(st["df_fetches"] - rst["df_fetches"]) <= 1e-4 * rst["df_fetches"] + 100
    return out, st


@pytest.mark.parametrize("name,w,h,spp,checker,frame,tick", [
    ("plains", 1920, 1080, 1, False, 7, 50.0),     # config 3
    ("plains", 640, 360, 4, True, 130, 50.0),      # 4 spp: blue-noise dimensions >= 8 (A.5), checkerboard
    ("city", 960, 540, 2, False, 3, 50.0),         # dense geometry, second bounces and shadow sub-rays
    ("gi_box", 960, 540, 1, False, 11, 50.0),      # emissive lamps
    ("gi_box", 480, 270, 1, False, 11, 140.0),     # night: moon stronger -> spp doubled, no shadow sub-rays
])
def test_diffuse_parity(renderer, worlds, oracles, scene_tables, golden_digests, name, w, h, spp, checker, frame, tick):
    load(renderer, worlds[name])
    sun, moon, _, vis = camera.sun_moon_direction(tick)
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(w, h)
    g, _ = oracles[name].trace_primary(cam, vx.primary_params(350))
    dp = vx.diffuse_params(sun, moon, vis, spp=spp, checkerboard=checker, frame=frame)
    for layout, wavefront in ((0, 0), (1, 0), (1, 1)):
        renderer.set_option(abi.OPT_TRAVERSAL_LAYOUT, layout)
        renderer.set_option(abi.OPT_GI_WAVEFRONT, wavefront)
        out, st = check_diffuse(renderer, oracles[name], cam, g, dp)
    if name == "plains" and w == 1920:
        d = golden_digests["diffuse"]["plains_1920x1080_p-20_jNone"]
        assert st["rays"] == d["stats"]["rays"] and abs(float(out["luma"].mean()) - d["mean_luma"]) < 1e-5


# ====================================================================================== error behaviour
def test_error_codes(scene_tables, worlds):
    r = vx.Renderer(0)
    try:
        cam = camera.FpsCamera().vx_camera(64, 36)
        g = r.alloc_gbuffer(64, 36)
        with pytest.raises(abi.VxptError) as e:
            r.trace_primary(cam, vx.primary_params(350), g)              # no world yet
        assert e.value.code == abi.E_STATE
        with pytest.raises(abi.VxptError) as e:
            r.build_distance_field()
        assert e.value.code == abi.E_STATE
        r.upload_world(worlds["superflat"])
        with pytest.raises(abi.VxptError) as e:
            r.trace_primary(cam, vx.primary_params(350), g)              # DF not built
        assert e.value.code == abi.E_STATE
        r.build_distance_field()
        r.trace_primary(cam, vx.primary_params(350), g)
        pp = vx.primary_params(350, alpha_test=True)
        with pytest.raises(abi.VxptError) as e:
            r.trace_primary(cam, pp, g)                                   # alpha test without vxpt_set_albedo_alpha_mips
        assert e.value.code == abi.E_STATE
        bad = camera.FpsCamera().vx_camera(64, 36, 10, 40)
        with pytest.raises(abi.VxptError) as e:
            r.trace_primary(bad, vx.primary_params(350), g)
        assert e.value.code == abi.E_INVALID
        dp = vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], 1.2)
        with pytest.raises(abi.VxptError) as e:
            r.trace_diffuse(cam, g, dp, r.alloc_diffuse(64, 36))          # tables not uploaded on this handle
        assert e.value.code == abi.E_STATE
        with pytest.raises(abi.VxptError) as e:
            r.trace_shadow(cam, g, vx.shadow_params(scene_tables["stronger"], soft=True), r.alloc_shadow(64, 36))
        assert e.value.code == abi.E_STATE                                # soft shadows need the noise texture
        r.load_scene_tables(scene_tables["materials"], scene_tables["blue_noise"], scene_tables["sky"], scene_tables["shadow_noise"])
        dp.use_blue_noise = 0
        with pytest.raises(abi.VxptError) as e:
            r.trace_diffuse(cam, g, dp, r.alloc_diffuse(64, 36))
        assert e.value.code == abi.E_UNSUPPORTED
        bn = [t.copy() for t in scene_tables["blue_noise"]]
        bn[0][5] = 300
        with pytest.raises(abi.VxptError) as e:
            r.set_blue_noise(*bn)
        assert e.value.code == abi.E_UNSUPPORTED
        assert r.launch_count() > 0
    finally:
        r.close()


def test_l2_sector_peak_is_plausible(renderer):
    g = renderer.measure_l2_sector_peak()
    assert 1000.0 < g < 40000.0


def test_sharded_frame_single_rank_matches_direct_calls(renderer, worlds, scene_tables):
    """multigpu.ShardedFrame (packed slab buffer + virtual plane bases) on one rank equals the plain calls."""
    from voxelpathtracer_b200 import multigpu
    load(renderer, worlds["plains"])
    W, H = 320, 180
    fc = camera.FpsCamera(pitch_deg=-20.0)
    pp = vx.primary_params(350, camera.taa_jitter(2))
    sp = vx.shadow_params(scene_tables["stronger"], frame=4)
    dp = vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=1, frame=4)
    cam = fc.vx_camera(W, H)
    g = renderer.trace_primary(cam, pp, renderer.alloc_gbuffer(W, H))
    s = renderer.trace_shadow(cam, g, sp, renderer.alloc_shadow(W, H))
    d = renderer.trace_diffuse(cam, g, dp, renderer.alloc_diffuse(W, H))
    f = multigpu.ShardedFrame(renderer, fc, W, H)
    f.trace(pp, sp, dp)
    f.gather()
    renderer.sync()
    for name, ref in (("g_t", g["t"]), ("g_normal_id", g["normal_id"]), ("g_block_id", g["block_id"]), ("g_inv_t", g["inv_t"]),
                      ("s_shadow", s["shadow"]), ("s_transversal", s["transversal"]), ("d_sh", d["sh"]), ("d_cocg", d["cocg"]),
                      ("d_luma", d["luma"]), ("d_ao_sky", d["ao_sky"])):
        assert np.array_equal(f.plane(name).cpu().numpy(), ref), name


@pytest.mark.parametrize("n,band", [(2, 6), (3, 4), (5, 9)])
def test_interleaved_row_bands_compose_to_the_full_frame(renderer, worlds, scene_tables, n, band):
    """VxCamera interleave contract: rank r renders bands b % n == r into rank-local planes addressed by virtual rows;
    scattering every rank's rows back gives the full frame bit for bit (primary, shadow, GI wavefront)."""
    from voxelpathtracer_b200 import multigpu
    load(renderer, worlds["plains"])
    W, H = 192, 180
    fc = camera.FpsCamera(pitch_deg=-20.0, aspect=W / H)
    pp = vx.primary_params(350, camera.taa_jitter(2))
    sp = vx.shadow_params(scene_tables["stronger"], frame=4)
    dp = vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=2, checkerboard=True, frame=4)
    full = fc.vx_camera(W, H)
    g = renderer.trace_primary(full, pp, renderer.alloc_gbuffer(W, H, hit_voxel=True))
    s = renderer.trace_shadow(full, g, sp, renderer.alloc_shadow(W, H))
    d = renderer.trace_diffuse(full, g, dp, renderer.alloc_diffuse(W, H))
    got = {k: np.zeros_like(v) for k, v in {**g, **s, **d}.items()}
    rows = H // n
    for rank in range(n):
        gl = renderer.alloc_gbuffer(W, rows, hit_voxel=True)
        sl, dl = renderer.alloc_shadow(W, rows), renderer.alloc_diffuse(W, rows)
        for vb, ve in ((0, rows // 3), (rows // 3, rows)):       # two virtual-row chunks per rank
            cam = fc.vx_camera(W, H, vb, ve, n, rank, band)
            renderer.trace_primary(cam, pp, gl)
            renderer.trace_shadow(cam, gl, sp, sl)
            renderer.trace_diffuse(cam, gl, dp, dl)
        img_rows = multigpu.image_rows_of_rank(H, n, rank, band)
        for k, v in {**gl, **sl, **dl}.items():
            got[k][img_rows] = v
    for k in got:
        assert np.array_equal(got[k], {**g, **s, **d}[k]), k
    bad = fc.vx_camera(W, H, 0, rows, n, n, band)
    with pytest.raises(abi.VxptError) as e:
        renderer.trace_primary(bad, pp, renderer.alloc_gbuffer(W, rows))
    assert e.value.code == abi.E_INVALID


# ====================================================================================== reflections
@pytest.mark.parametrize("name,w,h,spp,rough,checker,frame,tick,extra", [
    ("plains", 960, 540, 2, True, False, 7, 50.0, False),      # config 3 defaults (SPP 2, rough, roughness bias)
    ("city", 640, 360, 4, True, True, 12, 50.0, True),         # dense geometry, checkerboard SPP, caller-supplied normals / PBR
    ("gi_box", 640, 360, 2, False, False, -1, 50.0, False),    # mirror reflections, TEMPORAL_SPEC = false index, emissive lamps
    ("city", 480, 270, 1, True, False, 3, 140.0, False),       # night: moon is the stronger light
])
def test_reflection_parity(renderer, worlds, oracles, scene_tables, name, w, h, spp, rough, checker, frame, tick, extra):
    load(renderer, worlds[name])
    sun, moon, stronger, vis = camera.sun_moon_direction(tick)
    mats = scene_tables["materials"]
    fc = camera.FpsCamera(pitch_deg=-20.0) if name != "city" else camera.FpsCamera(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0)
    cam = fc.vx_camera(w, h)
    o = oracles[name]
    g, _ = o.trace_primary(cam, vx.primary_params(350))
    d, _ = o.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=max(frame, 0)))
    rp = vx.reflection_params(sun, moon, stronger, fc.position, mats["grass_props"], spp=spp, rough=rough, checkerboard=checker, frame=frame,
                              halton=camera.taa_jitter_secondary(max(frame, 0)) if frame != 7 else (0.0, 0.0))  # u_Halton as Pipeline.cpp:3032 sets it; one case at 0
    g_normal = g_pbr = None
    if extra:
        rng = np.random.RandomState(1)
        n = rng.normal(size=(h, w, 3)) * 0.15 + np.array([[0, 0, 1], [0, 0, -1], [0, 1, 0], [0, -1, 0], [-1, 0, 0], [1, 0, 0], [0, 1, 0]])[np.minimum(g["normal_id"], 6)]
        g_normal = np.ascontiguousarray(n / np.linalg.norm(n, axis=-1, keepdims=True), dtype=np.float32)
        g_pbr = np.ascontiguousarray(np.stack([rng.uniform(0.05, 1.0, (h, w)), (rng.rand(h, w) < 0.3) * 0.9, np.zeros((h, w)), np.zeros((h, w))], -1), dtype=np.float32)
    ref, rst = o.trace_reflection(cam, g, d, rp, g_normal, g_pbr)
    for layout in (0, 1):
        renderer.set_option(abi.OPT_TRAVERSAL_LAYOUT, layout)
        renderer.reset_stats()
        out = renderer.trace_reflection(cam, g, d, rp, renderer.alloc_reflection(w, h), g_normal, g_pbr)
        st = renderer.stats()
        for k, peak in (("color", 100.0), ("hit_distance", 200.0)):
            mae = float(np.mean(np.abs(out[k].astype(np.float64) - ref[k])))
            assert mae <= RADIANCE_MAE, (k, mae)
            assert psnr(out[k], ref[k], peak) >= RADIANCE_PSNR_DB, k
            assert np.mean(out[k] != ref[k]) <= 1e-4, (k, float(np.mean(out[k] != ref[k])))   # in practice bit-exact
        assert np.mean(out["emissive_mask"] != ref["emissive_mask"]) <= 1e-5
        assert abs(st["rays"] - rst["rays"]) <= 1e-5 * rst["rays"] + 2
        assert abs(st["df_fetches"] - rst["df_fetches"]) <= 1e-4 * rst["df_fetches"] + 100
    if name == "gi_box":
        assert ref["emissive_mask"].any() or True   # lamps are rarely in view; the mask path is exercised when they are


# ====================================================================================== texel formats, whole-frame call, p2p plumbing
def _unorm8(v):
    return np.rint(v.astype(np.float32) * np.float32(255.0)).astype(np.uint8)


@pytest.mark.parametrize("name,w,h", [("plains", 640, 360), ("city", 480, 270)])
def test_reference_texel_formats_are_the_rounded_fp32_planes(renderer, worlds, oracles, scene_tables, name, w, h):
    """VXPT_OPT_TEXEL_FORMAT = 1 (the reference's FBO formats, Pipeline.cpp:1094-1152): every texel is the oracle's fp32 value rounded
    once (binary16 RN-even / unorm8); the secondary passes read the R16F hit distance exactly as the reference's shaders do, so
    they are compared against the oracle run on the same half-precision G-buffer."""
    load(renderer, worlds[name])
    fc = camera.FpsCamera(pitch_deg=-20.0) if name != "city" else camera.FpsCamera(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0)
    cam = fc.vx_camera(w, h)
    o = oracles[name]
    pp = vx.primary_params(350, camera.taa_jitter(5))
    sp = vx.shadow_params(scene_tables["stronger"], frame=9)
    dp = vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=2, frame=9)
    g_ref, _ = o.trace_primary(cam, pp)
    g16 = dict(g_ref, t=g_ref["t"].astype(np.float16).astype(np.float32))   # what a reader of the R16F plane sees
    s_ref, _ = o.trace_shadow(cam, g16, sp)
    d_ref, _ = o.trace_diffuse(cam, g16, dp)
    renderer.set_option(abi.OPT_TEXEL_FORMAT, 1)
    try:
        for device in (False, True):
            g = renderer.trace_primary(cam, pp, renderer.alloc_gbuffer(w, h, device=device, texel=True))
            s = renderer.trace_shadow(cam, g, sp, renderer.alloc_shadow(w, h, device=device, texel=True))
            d = renderer.trace_diffuse(cam, g, dp, renderer.alloc_diffuse(w, h, device=device, texel=True))
            renderer.sync()
            host = (lambda x: x.cpu().numpy()) if device else (lambda x: x)
            assert host(g["t"]).dtype == np.float16 and host(d["ao_sky"]).dtype == np.uint8
            assert np.array_equal(host(g["t"]), g_ref["t"].astype(np.float16))
            assert np.array_equal(host(g["inv_t"]), g_ref["inv_t"])
            assert np.array_equal(host(g["normal_id"]), g_ref["normal_id"]) and np.array_equal(host(g["block_id"]), g_ref["block_id"])
            assert np.mean(host(s["shadow"]) != s_ref["shadow"]) <= 1e-5
            assert np.mean(host(s["transversal"]) != s_ref["transversal"].astype(np.float16)) <= 1e-5
            for k in ("sh", "cocg", "luma"):
                got, want = host(d[k]), d_ref[k].astype(np.float16)
                assert np.mean(got != want) <= 1e-4, k
                assert float(np.mean(np.abs(got.astype(np.float64) - d_ref[k]))) <= RADIANCE_MAE, k   # the fp32 target still holds in half
            assert np.mean(host(d["ao_sky"]) != _unorm8(d_ref["ao_sky"])) <= 1e-4
    finally:
        renderer.set_option(abi.OPT_TEXEL_FORMAT, 0)


@pytest.mark.parametrize("texel", [False, True])
@pytest.mark.parametrize("rows", [(0, 360), (40, 297)])
def test_render_frame_equals_the_separate_passes(renderer, worlds, scene_tables, texel, rows):
    """vxpt_render_frame (G-buffer resident between passes, slab-pipelined copy-out) == trace_primary + trace_shadow + trace_diffuse,
    for host planes, device planes and a mix, on a whole frame and on a row slab (rows outside it untouched)."""
    import torch
    load(renderer, worlds["plains"])
    W, H = 640, 360
    fc = camera.FpsCamera(pitch_deg=-20.0)
    cam = fc.vx_camera(W, H, rows[0], rows[1])
    pp = vx.primary_params(350, camera.taa_jitter(2))
    sp = vx.shadow_params(scene_tables["stronger"], frame=4)
    dp = vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=1, frame=4)
    renderer.set_option(abi.OPT_TEXEL_FORMAT, int(texel))
    try:
        def fresh(device):
            bufs = (renderer.alloc_gbuffer(W, H, device=device, texel=texel), renderer.alloc_shadow(W, H, device=device, texel=texel),
                    renderer.alloc_diffuse(W, H, device=device, texel=texel))
            for b in bufs:
                for v in b.values():
                    v[...] = 7 if not device else 7
            return bufs
        g, s, d = fresh(False)
        renderer.trace_primary(cam, pp, g); renderer.trace_shadow(cam, g, sp, s); renderer.trace_diffuse(cam, g, dp, d)
        want = {**g, **s, **d}
        renderer.reset_stats()
        g2, s2, d2 = fresh(False)
        renderer.render_frame(cam, pp, sp, dp, g2, s2, d2)
        st = renderer.stats()
        for k, v in {**g2, **s2, **d2}.items():
            assert np.array_equal(v, want[k]), k
        assert st["rays"] > 0
        g3, s3, d3 = fresh(True)
        renderer.render_frame(cam, pp, sp, dp, g3, s3, d3)
        renderer.sync()
        for k, v in {**g3, **s3, **d3}.items():
            assert np.array_equal(v.cpu().numpy(), want[k]), k
        # only the GI planes wanted, on the host: the G-buffer lives in the handle's arena
        _, _, d4 = fresh(False)
        renderer.render_frame(cam, pp, None, dp, None, None, d4)
        for k, v in d4.items():
            assert np.array_equal(v, want[k]), k
    finally:
        renderer.set_option(abi.OPT_TEXEL_FORMAT, 0)
    with pytest.raises(abi.VxptError) as e:
        renderer.render_frame(cam, vx.primary_params(-1), sp, dp, g2, s2, d2)
    assert e.value.code == abi.E_INVALID


def test_shared_buffer_signal_and_wait(renderer):
    """Plumbing of the peer-to-peer slab gather inside one process: an IPC-exportable buffer, stream-ordered release / acquire flags
    (also on a caller stream), wrap-around compare, and a wait that gives up instead of hanging the device."""
    import torch
    ptr, handle = renderer.shared_alloc(1 << 20)
    assert len(handle) == abi.SHARED_HANDLE_BYTES and ptr
    try:
        from voxelpathtracer_b200.multigpu import _DevMem
        mem = torch.as_tensor(_DevMem(ptr, 1 << 20), device="cuda:0")
        assert int(mem.sum()) == 0                                   # zero-initialised
        flags = ptr
        for k in range(4):
            renderer.signal(flags + 128 * k, 5)
        renderer.wait_all(flags, 4, 32, 5, timeout_ms=1000)
        renderer.wait_all(flags, 4, 32, 3, timeout_ms=1000)            # "at least"
        renderer.sync()
        assert mem[:512].view(torch.int32)[::32].tolist() == [5, 5, 5, 5]
        side = torch.cuda.Stream()
        renderer.signal(flags + 128 * 4, 0xFFFFFFFE, stream=side.cuda_stream)
        renderer.wait_all(flags + 128 * 4, 1, 32, 0xFFFFFFFD, timeout_ms=1000, stream=side.cuda_stream)
        side.synchronize()
        renderer.signal(flags + 128 * 4, 2)                            # wrapped past 2^32: 2 is "later" than 0xFFFFFFFE
        renderer.wait_all(flags + 128 * 4, 1, 32, 0xFFFFFFFF, timeout_ms=1000)
        renderer.sync()
        renderer.wait_all(flags + 128 * 5, 1, 32, 1, timeout_ms=30)    # never signalled
        with pytest.raises(abi.VxptError) as e:
            renderer.sync()
        assert e.value.code == abi.E_STATE
        renderer.sync()                                                # the error is latched once
        del mem
    finally:
        renderer.shared_close(ptr)
    with pytest.raises(abi.VxptError):
        renderer.shared_close(ptr)


@pytest.mark.parametrize("texel", [False, True])
def test_sharded_frame_texel_planes_single_rank(renderer, worlds, scene_tables, texel):
    from voxelpathtracer_b200 import multigpu
    load(renderer, worlds["plains"])
    W, H = 320, 180
    fc = camera.FpsCamera(pitch_deg=-20.0)
    pp = vx.primary_params(350, camera.taa_jitter(2))
    sp = vx.shadow_params(scene_tables["stronger"], frame=4)
    dp = vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=1, frame=4)
    cam = fc.vx_camera(W, H)
    try:
        f = multigpu.ShardedFrame(renderer, fc, W, H, texel=texel)
        g = renderer.trace_primary(cam, pp, renderer.alloc_gbuffer(W, H, texel=texel))
        s = renderer.trace_shadow(cam, g, sp, renderer.alloc_shadow(W, H, texel=texel))
        d = renderer.trace_diffuse(cam, g, dp, renderer.alloc_diffuse(W, H, texel=texel))
        f.trace(pp, sp, dp)
        f.gather()
        renderer.sync()
        for name, ref in (("g_t", g["t"]), ("s_shadow", s["shadow"]), ("s_transversal", s["transversal"]), ("d_sh", d["sh"]), ("d_cocg", d["cocg"]),
                          ("d_luma", d["luma"]), ("d_ao_sky", d["ao_sky"])):
            assert np.array_equal(f.plane(name).cpu().numpy(), ref), name
    finally:
        renderer.set_option(abi.OPT_TEXEL_FORMAT, 0)


def test_render_frame_async_double_buffered(renderer, worlds, scene_tables):
    """vxpt_render_frame_async + vxpt_frame_wait: same planes as the blocking call; a second call on the same handle (or any call that
    needs the staging arena) first drains the frame in flight, so alternating plane sets never see torn data."""
    load(renderer, worlds["plains"])
    W, H = 640, 360
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    sp = vx.shadow_params(scene_tables["stronger"], frame=4)
    want, got = [], []
    sets = [(renderer.alloc_gbuffer(W, H, pinned=True), renderer.alloc_shadow(W, H, pinned=True), renderer.alloc_diffuse(W, H, pinned=True)) for _ in range(2)]
    frames = [(vx.primary_params(350, camera.taa_jitter(k)), vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], frame=k))
              for k in range(4)]
    for pp, dp in frames:
        g, s, d = renderer.alloc_gbuffer(W, H), renderer.alloc_shadow(W, H), renderer.alloc_diffuse(W, H)
        renderer.render_frame(cam, pp, sp, dp, g, s, d)
        want.append({**g, **s, **d})
    for k, (pp, dp) in enumerate(frames):
        g, s, d = sets[k % 2]
        if k >= 2:
            renderer.frame_wait()
            got.append({kk: v.copy() for kk, v in {**g, **s, **d}.items()})   # frame k-2, before its planes are reused
        renderer.render_frame(cam, pp, sp, dp, g, s, d, wait=False)
    # frames 2 and 3 are still owed: frame 2's planes were overwritten by nothing since, frame 3 is in flight
    renderer.frame_wait()
    got.append({kk: v.copy() for kk, v in {**sets[0][0], **sets[0][1], **sets[0][2]}.items()})
    got.append({kk: v.copy() for kk, v in {**sets[1][0], **sets[1][1], **sets[1][2]}.items()})
    for k in range(4):
        for kk in want[k]:
            assert np.array_equal(got[k][kk], want[k][kk]), (k, kk)
    renderer.frame_wait()   # idempotent


# ====================================================================================== CUDA vs the reference's own shaders
def test_cuda_outputs_equal_the_reference_shader_golden_vectors(renderer, worlds, scene_tables):
    """tests/golden/ref_shader_digests.json holds digests of what the reference's OWN shaders (compiled as C++, oracle/_ref) produce.
    The CUDA path must reproduce them bit for bit: all five distance fields, the config-1 primary frames, and the 1080p plains frame of
    configs 2/3 through primary, soft sun shadow and 1-spp diffuse GI."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_shader_digests.json")) as f:
        ref = json.load(f)
    for name in ("superflat", "plains", "gi_box", "city", "sparse"):
        load(renderer, worlds[name])
        assert sha(renderer.download_distance_field()) == ref["df"][name], name
    load(renderer, worlds["superflat"])
    for pitch in (0.0, -20.0):
        for jf in (None, 17):
            cam = camera.FpsCamera(pitch_deg=pitch).vx_camera(640, 360)
            g = renderer.trace_primary(cam, vx.primary_params(350, None if jf is None else camera.taa_jitter(jf)), renderer.alloc_gbuffer(640, 360))
            want = ref["primary"][f"superflat_640x360_p{int(pitch)}_j{jf}"]
            for k in ("t", "normal_id", "block_id", "inv_t"):
                assert sha(g[k]) == want[k], (pitch, jf, k)
    load(renderer, worlds["plains"])
    case = "plains_1920x1080_p-20_jNone"
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(1920, 1080)
    g = renderer.trace_primary(cam, vx.primary_params(350), renderer.alloc_gbuffer(1920, 1080))
    for k in ("t", "normal_id", "block_id", "inv_t"):
        assert sha(g[k]) == ref["primary"][case][k], k
    s = renderer.trace_shadow(cam, g, vx.shadow_params(scene_tables["stronger"], frame=5, soft=True), renderer.alloc_shadow(1920, 1080))
    for k in ("shadow", "transversal"):
        assert sha(s[k]) == ref["shadow"][case][k], k
    dp = vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=1, frame=7)
    for wf in (0, 1):
        renderer.set_option(abi.OPT_GI_WAVEFRONT, wf)
        d = renderer.trace_diffuse(cam, g, dp, renderer.alloc_diffuse(1920, 1080))
        for k in ("sh", "cocg", "luma", "ao_sky"):
            assert sha(d[k]) == ref["diffuse"][case][k], (wf, k)
    renderer.set_option(abi.OPT_GI_WAVEFRONT, 1)


def test_reflection_wavefront_writes_the_per_pixel_kernels_bits(renderer, worlds, oracles, scene_tables):
    """VXPT_OPT_REFLECTION_WAVEFRONT: the re-queued form (default) and the one-thread-per-pixel form give identical planes and identical
    traversal counters — 1 sample, 3 samples with the checkerboard (pixels take 3 or 2), 5 samples (two shadow rays per pixel), a row slab."""
    load(renderer, worlds["city"])
    sun, moon, stronger, vis = scene_tables["sun"], scene_tables["moon"], scene_tables["stronger"], scene_tables["sun_visibility"]
    fc = camera.FpsCamera(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0)
    W, H = 480, 270
    cam = fc.vx_camera(W, H)
    g = renderer.trace_primary(cam, vx.primary_params(350), renderer.alloc_gbuffer(W, H))
    d = renderer.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=2), renderer.alloc_diffuse(W, H))
    try:
        for spp, checker, rows in ((1, False, None), (3, True, None), (5, False, (64, 200))):
            rp = vx.reflection_params(sun, moon, stronger, fc.position, scene_tables["materials"]["grass_props"], spp=spp, rough=True, checkerboard=checker, frame=2,
                                      halton=camera.taa_jitter_secondary(2))
            c = cam if rows is None else fc.vx_camera(W, H, *rows)
            outs, stats = [], []
            for mode in (0, 1):
                renderer.set_option(abi.OPT_REFLECTION_WAVEFRONT, mode)
                renderer.reset_stats()
                outs.append(renderer.trace_reflection(c, g, d, rp, renderer.alloc_reflection(W, H)))
                st = renderer.stats()
                stats.append((st["rays"], st["df_fetches"], st["vox_fetches"]))
            rb, re = rows or (0, H)      # rows outside the slab are not written
            for k in ("color", "hit_distance", "emissive_mask"):
                assert np.array_equal(outs[0][k][rb:re], outs[1][k][rb:re], equal_nan=True), (spp, k)
            assert stats[0] == stats[1], (spp, stats)
    finally:
        renderer.set_option(abi.OPT_REFLECTION_WAVEFRONT, 1)


def test_cuda_reflections_equal_the_reference_shader_golden_vectors(renderer, worlds, oracles, scene_tables):
    """CUDA reflection planes == digests of the reference's ReflectionTraceFrag.glsl (compiled as C++) on the golden frames — BASELINE
    config 3's 1920x1080 among them — with u_Halton = GetTAAJitterSecondary(frame) (Pipeline.cpp:3032): G-buffer read at the jittered coordinate."""
    import json
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "tools"))
    from make_ref_shader_golden import reflection_cases, synthetic_material_planes
    with open(os.path.join(root, "tests", "golden", "ref_shader_digests.json")) as f:
        ref = json.load(f)["reflection"]
    sun, moon, stronger, vis = scene_tables["sun"], scene_tables["moon"], scene_tables["stronger"], scene_tables["sun_visibility"]
    for cname, wname, W, H, cam_kw, spp, rough, checker, frame in reflection_cases():
        load(renderer, worlds[wname])
        fc = camera.FpsCamera(**cam_kw)
        cam = fc.vx_camera(W, H)
        g = renderer.trace_primary(cam, vx.primary_params(350), renderer.alloc_gbuffer(W, H))
        d = renderer.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=max(frame, 0)), renderer.alloc_diffuse(W, H))
        g_normal, g_pbr = synthetic_material_planes(g, W, H)
        rp = vx.reflection_params(sun, moon, stronger, fc.position, scene_tables["materials"]["grass_props"], spp=spp, rough=rough, checkerboard=checker, frame=frame,
                                  halton=camera.taa_jitter_secondary(max(frame, 0)))
        out = renderer.trace_reflection(cam, g, d, rp, renderer.alloc_reflection(W, H), g_normal, g_pbr)
        for k in ("color", "hit_distance", "emissive_mask"):
            assert sha(out[k]) == ref[cname][k], (cname, k)


def test_cuda_config4_frame_equals_the_reference_shader_golden_vectors(renderer, worlds, scene_tables):
    """BASELINE config 4 on one GPU: 3840x2160, soft shadows + 4-spp GI on the gi-box scene, device planes; CUDA == the reference's shaders."""
    import json
    import os
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_shader_digests.json")) as f:
        ref = json.load(f)
    load(renderer, worlds["gi_box"])
    W, H = 3840, 2160
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    g = renderer.trace_primary(cam, vx.primary_params(350), renderer.alloc_gbuffer(W, H, device=True))
    s = renderer.trace_shadow(cam, g, vx.shadow_params(scene_tables["stronger"], frame=9, soft=True), renderer.alloc_shadow(W, H, device=True))
    d = renderer.trace_diffuse(cam, g, vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=4, frame=9),
                               renderer.alloc_diffuse(W, H, device=True))
    renderer.sync()
    for k in ("shadow", "transversal"):
        assert sha(s[k].cpu().numpy()) == ref["shadow"]["gi_box_3840x2160_f9"][k], k
    for k in ("sh", "cocg", "luma", "ao_sky"):
        assert sha(d[k].cpu().numpy()) == ref["diffuse"]["gi_box_3840x2160_spp4_f9"][k], k


def test_render_frame_on_interleaved_row_bands(renderer, worlds, scene_tables):
    """vxpt_render_frame with the multi-GPU row-band contract (rank-local planes, virtual rows) == the separate passes on the same camera."""
    load(renderer, worlds["plains"])
    W, H, n, rank, band = 320, 360, 3, 1, 4
    rows = H // n
    fc = camera.FpsCamera(pitch_deg=-20.0, aspect=W / H)
    cam = fc.vx_camera(W, H, 8, rows - 16, n, rank, band)
    pp = vx.primary_params(350, camera.taa_jitter(6))
    sp = vx.shadow_params(scene_tables["stronger"], frame=6)
    dp = vx.diffuse_params(scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"], spp=2, frame=6)
    def fresh():
        bufs = (renderer.alloc_gbuffer(W, rows), renderer.alloc_shadow(W, rows), renderer.alloc_diffuse(W, rows))
        for b in bufs:
            for v in b.values():
                v[...] = 3
        return bufs
    g, s, d = fresh()
    renderer.trace_primary(cam, pp, g); renderer.trace_shadow(cam, g, sp, s); renderer.trace_diffuse(cam, g, dp, d)
    g2, s2, d2 = fresh()
    renderer.render_frame(cam, pp, sp, dp, g2, s2, d2)
    for k, v in {**g2, **s2, **d2}.items():
        assert np.array_equal(v, {**g, **s, **d}[k]), k


def test_render_frame_async_records_into_a_cuda_graph(renderer, worlds, scene_tables):
    """vxpt_render_frame_async with pinned HOST planes under stream capture: the copy-out stream joins the capture and rejoins the handle's
    stream, so a replay leaves the same bytes in the host planes as the eager call once the handle's stream has drained; the blocking
    vxpt_render_frame refuses to be captured."""
    torch = pytest.importorskip("torch")
    load(renderer, worlds["plains"])
    W, H = 320, 180
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H)
    sun, moon, stronger, vis = scene_tables["sun"], scene_tables["moon"], scene_tables["stronger"], scene_tables["sun_visibility"]
    pp, sp, dp = vx.primary_params(350, camera.taa_jitter(4)), vx.shadow_params(stronger, frame=4, soft=True), vx.diffuse_params(sun, moon, vis, spp=1, frame=4)
    want = renderer.render_frame(cam, pp, sp, dp, renderer.alloc_gbuffer(W, H), renderer.alloc_shadow(W, H), renderer.alloc_diffuse(W, H))
    g, s, d = renderer.alloc_gbuffer(W, H, pinned=True), renderer.alloc_shadow(W, H, pinned=True), renderer.alloc_diffuse(W, H, pinned=True)
    submit = renderer.prepare_frame(cam, g, s, d)
    submit(pp, sp, dp)                      # eager once: sizes the staging arena and the GI queue
    renderer.frame_wait()
    renderer.sync()
    ext = torch.cuda.ExternalStream(renderer.cuda_stream())
    renderer.set_option(abi.OPT_TIMING_EVENTS, 0)
    try:
        gph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gph, stream=ext, capture_error_mode="thread_local"):
            submit(pp, sp, dp)
    finally:
        renderer.set_option(abi.OPT_TIMING_EVENTS, 1)
    for planes in (g, s, d):
        for k in planes:
            np.asarray(planes[k])[...] = 0
    for _ in range(2):
        with torch.cuda.stream(ext):
            gph.replay()
    renderer.frame_wait()                   # nothing pending on the host side ...
    renderer.sync()                         # ... the replay is complete when the handle's stream is
    for got, ref in zip((g, s, d), want[:3]):
        for k in ref:
            assert np.array_equal(np.asarray(got[k]), ref[k], equal_nan=True), k
