"""Host-side mirror of the reference interfaces: world layout/generators/save files, camera, jitter, sun."""
import math
import os

import numpy as np
import pytest

from voxelpathtracer_b200 import abi, camera, world

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_world_layout_is_x_fastest():
    w = world.World()
    w.set_block(3, 5, 7, 42)                                  # Core/World.h:46-70
    assert w.data[3 + 5 * 384 + 7 * 384 * 128] == 42 and w.get_block(3, 5, 7) == 42 and w.zyx[7, 5, 3] == 42
    assert w.GetBlock(3, 5, 7) == 42
    with pytest.raises(IndexError):
        w.set_block(384, 0, 0, 1)
    with pytest.raises(ValueError):
        world.World(np.zeros(10, np.uint8))


def test_superflat_columns():
    w = world.generate_superflat().zyx                       # WorldGenerator.cpp:28-46,109-121
    col = w[10, :, 20]
    assert np.all(col[:45] == world.STONE) and np.all(col[45:49] == world.DIRT) and col[49] == world.GRASS and np.all(col[50:] == 0)
    assert np.all(w == w[0:1, :, 0:1])


def test_plains_follow_the_reference_column_rules(plains_columns, worlds):
    cols = plains_columns.reshape(384, 384, 2)
    w = worlds["plains"].zyx
    rng = np.random.RandomState(0)
    for _ in range(200):
        x, z = int(rng.randint(384)), int(rng.randint(384))
        h, biome = int(cols[x, z, 0]), int(cols[x, z, 1])
        col = w[z, :, x]
        assert np.all(col[h:] == 0) and np.all(col[:h] > 0)
        if biome == 1:
            assert col[h - 1] == world.GRASS and np.all(col[h - 5:h - 1] == world.DIRT) and col[h - 6] == world.STONE
        else:
            assert np.all(col[h - 8:h] == world.SAND) and col[h - 9] == world.STONE
    assert set(np.unique(cols[:, :, 1])) <= {0, 1} and cols[:, :, 0].min() == 43 and cols[:, :, 0].max() == 58


def test_stand_in_worlds_are_deterministic_and_dense(plains_columns):
    a, b = world.generate_city(), world.generate_city()
    assert np.array_equal(a.data, b.data)
    assert (a.data > 0).mean() >= 0.25 and a.zyx[:, 100:, :].any()      # BASELINE.md config 5 stand-in
    g = world.generate_gi_box(plains_columns)
    assert (g.data == world.LAMP).sum() >= 20                            # emissive blocks for config 4


def test_save_load_roundtrip(tmp_path, worlds):
    p = os.path.join(tmp_path, "Saves", "gi")
    world.save_world(worlds["sparse"], p)
    assert os.path.getsize(p) == abi.WORLD_VOXELS                        # WorldFileHandler.cpp:27
    assert np.array_equal(world.load_world(p).data, worlds["sparse"].data)
    open(p, "ab").write(b"x")
    with pytest.raises(ValueError):
        world.load_world(p)


def test_camera_matrices_invert_and_point_down_z():
    c = camera.FpsCamera()
    assert np.allclose(c.front, [0, 0, 1], atol=1e-12)
    cam = c.vx_camera(640, 360)
    inv_view = np.array(cam.inv_view[:]).reshape(4, 4).T
    inv_proj = np.array(cam.inv_proj[:]).reshape(4, 4).T
    assert np.allclose(inv_view @ c.view(), np.eye(4), atol=1e-5)
    assert np.allclose(inv_proj @ c.projection(), np.eye(4), atol=1e-4)
    assert np.allclose(inv_view[:3, 3], [192, 75, 192])                  # u_InverseView[3].xyz = camera position
    # centre of the screen maps to the view direction (GetRayStuff, InitialRayTraceFrag.glsl:410-413)
    eye = inv_proj @ np.array([0.0, 0.0, -1.0, 1.0])
    d = inv_view @ np.array([eye[0], eye[1], -1.0, 0.0])
    assert np.allclose(d[:3] / np.linalg.norm(d[:3]), [0, 0, 1], atol=1e-6)
    assert (cam.width, cam.height, cam.row_begin, cam.row_end) == (640, 360, 0, 360)
    # top-right pixel looks up and to +x... for a camera facing +z with up = +y, right is -x
    eye = inv_proj @ np.array([1.0, 1.0, -1.0, 1.0])
    d = inv_view @ np.array([eye[0], eye[1], -1.0, 0.0])
    assert d[1] > 0 and abs(d[1] / d[2] - math.tan(math.radians(30))) < 1e-5 and abs(abs(d[0] / d[1]) - 16 / 9) < 1e-4


def test_halton_table_matches_taajitter():
    # Core/TAAJitter.cpp:6-28 with primes 2 and 3, indices 1..64
    assert camera.HALTON_TABLE[0] == (0.5, pytest.approx(1 / 3, abs=1e-7))
    assert camera.HALTON_TABLE[1] == (0.25, pytest.approx(2 / 3, abs=1e-7))
    assert camera.HALTON_TABLE[2][0] == 0.75 and camera.HALTON_TABLE[3][0] == 0.125
    assert len(camera.HALTON_TABLE) == 64 and camera.taa_jitter(64 + 5) == camera.HALTON_TABLE[5]
    assert camera.taa_jitter_secondary(32 + 5) == camera.HALTON_TABLE[5]
    assert all(0 < x < 1 and 0 < y < 1 for x, y in camera.HALTON_TABLE)


def test_sun_direction_at_suntick_50():
    sun, moon, stronger, vis = camera.sun_moon_direction(50.0)
    assert np.allclose(sun, [-0.669, 0.468, 0.577], atol=1e-3)
    assert np.allclose(moon, [-sun[0], -sun[1], sun[2]], atol=1e-6) and np.array_equal(stronger, sun)
    assert abs(np.linalg.norm(sun) - 1) < 1e-6 and vis == pytest.approx(1.2, abs=1e-6)
    _, _, stronger_night, vis_n = camera.sun_moon_direction(140.0)
    assert stronger_night[1] > 0 and vis_n == 0.0                         # moon is the stronger light below the horizon


def test_brick_offset_is_injective_and_matches_its_bit_layout():
    """Python restatement of brick_offset() in csrc/vxpt_internal.h: three independent bit-deposits (LOP3 + IMAD each)
    that place every voxel at a distinct byte of the 25,165,824-byte tiled step field."""
    u = np.uint32
    x, y, z = np.meshgrid(np.arange(384, dtype=u), np.arange(128, dtype=u), np.arange(384, dtype=u), indexing="ij")
    off = (x + (x & ~u(3)) * u(15)) + (y * u(16) + (y & ~u(3)) * u(2032)) + (z * u(4) + (z & ~u(3)) * u(65532))
    bits = (x & 3) | ((z & 3) << 2) | ((y & 3) << 4) | (((x >> 2) & 1) << 6) | ((x >> 3) << 7) | ((y >> 2) << 13) | ((z >> 2) << 18)
    assert np.array_equal(off, bits)
    assert off.max() < (96 << 18) and np.unique(off.reshape(-1)).size == abi.WORLD_VOXELS
    # one 32-byte sector = 4 x * 4 z * 2 y voxels, one 128-byte line = 8 x * 4 y * 4 z
    assert np.unique(off[:4, :2, :4] >> 5).size == 1 and np.unique(off[:8, :4, :4] >> 7).size == 1
    # what the step-field stores of df_z_dpx<1> rely on: the four z-neighbours of an x-word are 16 contiguous, 16-byte aligned bytes
    w = off[::4, :, ::4]
    assert (w % 16 == 0).all() and all(np.array_equal(off[::4, :, k::4], w + u(4 * k)) for k in range(4))
    # and the closed form of the step value that DESIGN.md quotes for M != 1
    m = np.arange(256)
    lut = np.where(m == 1, 1, np.floor(m.astype(np.float32) * np.float32(0.57735026918)).astype(np.int64))
    assert np.array_equal(lut[m != 1], ((m * 9459) >> 14)[m != 1])


def test_renderer_wrappers_call_through_with_a_stub_library():
    """No GPU here: run every newer Renderer wrapper against a stub of the C library (records the call, returns VXPT_OK), so that a
    misspelt attribute or a bad argument list in the Python layer fails in the CPU suite rather than on the GPU box."""
    import ctypes as C
    import voxelpathtracer_b200 as vx
    from voxelpathtracer_b200 import abi

    calls = []

    class Stub:
        def __getattr__(self, name):
            if name not in abi.EXPORTS:
                raise AttributeError(name)

            def fn(*args):
                assert len(args) == len(abi.EXPORTS[name][1]), (name, len(args))
                calls.append(name)
                return abi.OK
            return fn

    r = vx.Renderer.__new__(vx.Renderer)
    r.lib, r.handle, r.device, r._keep = Stub(), C.c_void_p(1), 0, []
    r.set_albedo_alpha_mips(np.zeros((2, abi.ALPHA_MIP_TEXELS), np.uint8))
    out = r.trace_rays(np.zeros((5, 3), np.float32), np.ones((5, 3), np.float32), 64)
    assert out["t"].shape == (5,) and out["hit_voxel"].shape == (5, 3)
    assert r.player_shadowed((1.0, 2.0, 3.0), (0.0, 1.0, 0.0)) is False
    agg, per = r.estimate_ambient_sound((1.0, 2.0, 3.0), 7)
    assert agg == 0 and per.shape == (32,)
    assert calls == ["vxpt_set_albedo_alpha_mips", "vxpt_trace_rays", "vxpt_player_shadowed", "vxpt_estimate_ambient_sound"]
    # the passes on either side of the path
    from voxelpathtracer_b200 import denoise
    del calls[:]
    W, H = 32, 18
    fc = camera.FpsCamera(aspect=W / H)
    cam = fc.vx_camera(W, H)
    r.set_gbuffer_textures(*[np.zeros((2, abi.MIP_CHAIN_TEXELS, 4), np.uint8)] * 3)
    g = r.alloc_gbuffer(W, H)
    m = r.generate_gbuffer(cam, g, vx.material_params(np.zeros(10, np.int32)), r.alloc_material(W, H))
    assert m["albedo"].shape == (H, W, 3) and m["pbr"].shape == (H, W, 4)
    d = r.alloc_diffuse(W, H)
    prev_t = r.alloc_denoise(W, H, ("sh", "cocg", "utility", "ao_sky"))
    tp = denoise.temporal_params(*fc.view_projection_f32())
    out, temporal = r.svgf_denoise(cam, g, g, d, prev_t, tp, time=1.0)
    assert out["sh"].shape == (H, W, 4) and out["variance"].shape == (H, W) and temporal["utility"].shape == (H, W, 3)
    s = r.alloc_shadow(W, H)
    st = r.shadow_temporal(cam, g, g, s, r.alloc_denoise(W, H, ("shadow", "frames")), denoise.shadow_temporal_params(*fc.view_projection_f32()),
                           r.alloc_denoise(W, H, ("shadow", "frames")))
    f = r.shadow_filter(cam, g, st, s["transversal"], denoise.shadow_filter_params(1.0), np.zeros((H, W), np.float32))
    assert f.shape == (H, W)
    assert calls == ["vxpt_set_gbuffer_textures", "vxpt_generate_gbuffer", "vxpt_svgf_initial", "vxpt_svgf_temporal", "vxpt_svgf_variance"] + \
        ["vxpt_svgf_spatial"] * 5 + ["vxpt_shadow_temporal", "vxpt_shadow_filter"]
    r.handle = C.c_void_p()      # nothing to destroy


# ---------------------------------------------------------------------------------------------------- CPU picking (World::Raycast)
def test_picking_known_answers():
    """World::Raycast / RaycastDetect (Core/World.cpp:215-546) on the superflat world (top layer y = 49, grass)."""
    w = world.generate_superflat()
    d = np.array([0.6, -0.7, 0.39])
    d /= np.linalg.norm(d)
    eye = (192.5, 53.5, 192.5)
    assert w.raycast_detect(eye, d) == (195, 49, 194, world.GRASS)          # the ray drops 3.5 blocks: x + 3.0, z + 1.95
    assert w.raycast_detect((192.5, 100.0, 192.5), (0.1, 1.0, 0.1)) is None    # looking up: nothing within 48 steps
    r = w.raycast(1, eye, d, held_block=world.LAMP)                            # place on the top face
    assert r == {"changed": True, "voxel": (195, 50, 194), "block": world.LAMP} and w.get_block(195, 50, 194) == world.LAMP
    assert w.raycast(2, eye, d)["block"] == world.LAMP                         # pick: the lamp is now what the ray meets first
    r = w.raycast(0, eye, d)                                                   # break it again
    assert r == {"changed": True, "voxel": (195, 50, 194), "block": world.LAMP} and w.get_block(195, 50, 194) == 0
    # a block that would intersect the player is refused: standing on the ground, looking almost straight down
    d2 = np.array([0.01, -1.0, 0.01])
    d2 /= np.linalg.norm(d2)
    before = w.data.copy()
    assert w.raycast(1, (192.5, 51.6, 192.5), d2)["changed"] is False and np.array_equal(w.data, before)


def test_picking_python_and_cpp_mirrors_agree(tmp_path):
    """The same random edit script through the Python mirror (world.World.raycast) and the C++ mirror (host/VoxelRT.h, compiled here)."""
    import subprocess
    abi.load()
    cc = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    exe = str(tmp_path / "picking")
    pkg = os.path.join(ROOT, "voxelpathtracer_b200")
    subprocess.run([cc, "-std=c++17", "-O2", "-Wall", "-Werror", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "cpp", "picking_main.cpp"),
                    "-L" + pkg, "-lvxpt", "-Wl,-rpath," + pkg], check=True)
    rng = np.random.RandomState(17)
    w = world.generate_superflat()
    lines, want = [], []
    for k in range(300):
        op = int(rng.choice([0, 1, 1, 2, 3]))
        pos = np.array([192.5 + rng.uniform(-6, 6), 50.2 + rng.uniform(0.0, 6.0), 192.5 + rng.uniform(-6, 6)], np.float32)
        d = rng.normal(size=3)
        d[1] = -abs(d[1]) - 0.2
        d = (d / np.linalg.norm(d)).astype(np.float32)
        held = int(rng.choice([world.STONE, world.LAMP, world.PLANKS]))
        lines.append("%d %.9g %.9g %.9g %.9g %.9g %.9g %d" % (op, *pos, *d, held))
        if op == 3:
            r = w.raycast_detect(pos, d)
            want.append((1,) + r if r else (0, -1, -1, -1, -1))
        else:
            r = w.raycast(op, pos, d, held)
            v = r["voxel"] or (-1, -1, -1)
            want.append((int(r["changed"]), v[0], v[1], v[2], r["block"]))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    got = [tuple(int(v) for v in ln.split()) for ln in out if ln.strip()]
    assert got == want
    assert sum(1 for g in got if g[0] == 1) > 100     # the script does edit the world
