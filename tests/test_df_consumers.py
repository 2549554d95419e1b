"""The other consumers of the distance field: the ray batch behind PostProcessingVert.glsl's player-shadow test
(vxpt_trace_rays / vxpt_player_shadowed) and EstimateAmbientSoundLevel.comp (vxpt_estimate_ambient_sound).

CPU: the oracle against the reference's compiled compute shader (live, when oracle/_ref is built) and against committed golden values;
the kernels' own source on the host against the oracle.  GPU: the CUDA kernels through the C ABI against the oracle."""
import json
import os
import re

import numpy as np
import pytest

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import abi, camera
from oracle import ref_shaders

from host_shadow import kernels_on_host as koh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SHADERS = "/root/reference/Core/Shaders"
needs_ref = pytest.mark.skipif(not ref_shaders.available(), reason="oracle/_ref/libref_shaders.so not built (needs /root/reference at build time)")


def listener_positions(world, n, seed):
    """Air voxels with something solid above them (rooms, arcades, overhangs): where the ambience estimate is not trivially 'open sky'."""
    v = world.zyx
    roof = np.flip(np.maximum.accumulate(np.flip(v > 0, axis=1), axis=1), axis=1)
    cand = np.argwhere((v == 0) & roof & (np.arange(128)[None, :, None] > 41))
    if len(cand) == 0:                                                   # open terrain: nothing overhangs
        return []
    rng = np.random.RandomState(seed)
    pick = cand[rng.choice(len(cand), n, replace=False)]
    return [(float(x) + 0.37, float(y) + 0.52, float(z) + 0.61) for z, y, x in pick]


def ray_batch(n, seed):
    rng = np.random.RandomState(seed)
    o = np.stack([rng.uniform(-20, 404, n), rng.uniform(30, 140, n), rng.uniform(-20, 404, n)], -1).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d[: n // 8, rng.randint(0, 3)] = 0.0                                  # some axis-aligned components
    d = (d / np.linalg.norm(d, axis=-1, keepdims=True)).astype(np.float32)
    return o, d


@pytest.fixture(scope="module")
def golden():
    with open(os.path.join(ROOT, "tests", "golden", "df_consumer_golden.json")) as f:
        return json.load(f)


@pytest.mark.skipif(not os.path.isdir(REF_SHADERS), reason="reference tree not present")
def test_post_processing_traversal_is_the_primary_traversal():
    """PostProcessingVert.glsl:103-170 == InitialRayTraceFrag.glsl:307-374 word for word, apart from the literal iteration cap."""
    def fn(path):
        s = open(os.path.join(REF_SHADERS, path), encoding="utf-8", errors="replace").read()
        i = s.index("float VoxelTraversalDF(vec3 origin, vec3 direction, inout vec3 normal, inout float blockType) \n{")
        depth, k = 0, s.index("{", i)
        for k in range(k, len(s)):
            depth += {"{": 1, "}": -1}.get(s[k], 0)
            if depth == 0:
                break
        return re.sub(r"\s+", " ", s[i:k + 1])
    assert fn("PostProcessingVert.glsl") == fn("InitialRayTraceFrag.glsl").replace("itr < u_RenderDistance", "itr < 350")


@needs_ref
@pytest.mark.parametrize("name", ["gi_box", "city", "plains"])
def test_oracle_ambient_sound_equals_the_reference_shader_live(oracles, worlds, name):
    o = oracles[name]
    spots = listener_positions(worlds[name], 12, 3) + [(192.0, 75.0, 192.0), (500.0, 80.0, 100.0), (192.0, 20.0, 192.0)]
    seen = set()
    for pos in spots:
        for frame in (0, 1, 6, 333, 511, 512):
            agg, per, _ = o.ambient_sound(pos, frame)
            ragg, rper = ref_shaders.ambient_sound(o.df, pos, frame)
            assert agg == ragg and np.array_equal(per, rper), (pos, frame)
            assert agg == int(per.sum()) and set(np.unique(per)) <= {0, 256, 512}
            seen.add(agg)
    assert name == "plains" or len(seen) > 5          # the interior spots give a spread of values, not just "open sky"


def test_oracle_ambient_sound_golden_values(oracles, worlds, golden):
    """Committed values produced by the reference's compiled shader (tools/make_df_consumer_golden.py)."""
    for name, cases in golden["ambient"].items():
        o = oracles[name]
        for c in cases:
            agg, per, _ = o.ambient_sound(c["pos"], c["frame"])
            assert agg == c["aggregate"] and per.tolist() == c["per_invocation"], (name, c["pos"], c["frame"])


def test_ambient_sound_analytic(oracles):
    """Under an open sky every path escapes: 32 invocations x 512.  Inside solid rock the very first traversal ends on E == 0 without a
    DDA step, returns -1 and the sample counts as escaped too (the shader's behaviour, not a physical one)."""
    o = oracles["superflat"]
    assert o.ambient_sound((192.0, 90.0, 192.0), 4)[0] == 32 * 512
    assert o.ambient_sound((192.0, 20.0, 192.0), 5)[0] == 32 * 512
    assert o.ambient_sound((192.0, 300.0, 192.0), 5)[0] == 32 * 512     # outside the volume


def test_oracle_ray_batch_is_the_primary_traversal(oracles):
    """vxo_trace_rays on the rays the primary pass forms == the primary pass."""
    o = oracles["plains"]
    cam = camera.FpsCamera(pitch_deg=-20.0).vx_camera(64, 36)
    g, st = o.trace_primary(cam, vx.primary_params(350))
    inv_view = np.frombuffer(cam.inv_view, dtype=np.float32).reshape(4, 4).T     # column-major
    inv_proj = np.frombuffer(cam.inv_proj, dtype=np.float32).reshape(4, 4).T
    origins, dirs = [], []
    f = np.float32
    for j in range(36):
        for i in range(64):
            u, v = (f(i) + f(0.5)) / f(64), (f(j) + f(0.5)) / f(36)
            clip = np.array([u * f(2) - f(1), v * f(2) - f(1), -1, 1], np.float32)
            e = [(inv_proj[r, 0] * clip[0] + inv_proj[r, 1] * clip[1]) + (inv_proj[r, 2] * clip[2] + inv_proj[r, 3] * clip[3]) for r in range(2)]
            eye = np.array([e[0], e[1], -1, 0], np.float32)
            rd = np.array([(inv_view[r, 0] * eye[0] + inv_view[r, 1] * eye[1]) + (inv_view[r, 2] * eye[2] + inv_view[r, 3] * eye[3]) for r in range(3)], np.float32)
            rd = rd * (f(1) / np.sqrt((rd[0] * rd[0] + rd[1] * rd[1]) + rd[2] * rd[2], dtype=np.float32))
            origins.append(inv_view[:3, 3]); dirs.append(rd)
    r, rst = o.trace_rays(np.array(origins), np.array(dirs), 350)
    assert np.array_equal(r["t"].reshape(36, 64), g["t"]) and np.array_equal(r["normal_id"].reshape(36, 64), g["normal_id"])
    assert np.array_equal(r["block_id"].reshape(36, 64), g["block_id"]) and np.array_equal(r["hit_voxel"].reshape(36, 64, 3), g["hit_voxel"])
    assert rst == st


def test_player_shadow(oracles, scene_tables):
    city = oracles["city"]
    sun = scene_tables["sun"]
    assert city.player_shadowed((9.5, 43.5, 9.5), sun) is True          # inside the first tower's ground floor
    assert city.player_shadowed((9.5, 125.0, 9.5), sun) is False        # above every roof
    assert oracles["superflat"].player_shadowed((192.0, 75.0, 192.0), sun) is False
    assert oracles["superflat"].player_shadowed((192.0, 75.0, 192.0), [-s for s in sun]) is True   # sun below the ground plane
    # the direction is normalised by a division (u_VertSunDir / L), any positive scale of it gives the same verdict
    assert city.player_shadowed((9.5, 43.5, 9.5), [3.0 * s for s in sun]) is True


@pytest.mark.skipif(not koh.available(), reason="CUDA toolkit headers not present")
@pytest.mark.parametrize("layout", [1, 0])
def test_consumer_kernels_on_host_equal_the_oracle(oracles, worlds, layout):
    o = oracles["city"]
    k = koh.HostKernels(o, layout)
    origins, dirs = ray_batch(3000, 9)
    ref, rst = o.trace_rays(origins, dirs, 200)
    out, st = k.trace_rays(origins, dirs, 200)
    for key in ref:
        assert np.array_equal(out[key], ref[key]), key
    assert st == rst and (ref["t"] > 0).mean() > 0.2
    for pos in listener_positions(worlds["city"], 6, 8) + [(192.0, 125.0, 192.0)]:
        for frame in (0, 3, 512 + 77):
            agg, per, rst = o.ambient_sound(pos, frame)
            hagg, hper, st = k.ambient_sound(pos, frame)
            assert hagg == agg and np.array_equal(hper, per) and st == rst, (pos, frame)
    k.close()


# ------------------------------------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_cuda_consumers_equal_the_oracle(renderer, oracles, worlds, scene_tables, golden):
    renderer.upload_world(worlds["city"])
    renderer.build_distance_field()
    o = oracles["city"]
    origins, dirs = ray_batch(100000, 21)
    ref, rst = o.trace_rays(origins, dirs, 350)
    for layout in (1, 0):
        renderer.set_option(abi.OPT_TRAVERSAL_LAYOUT, layout)
        renderer.reset_stats()
        out = renderer.trace_rays(origins, dirs, 350)
        st = renderer.stats()
        for key in ref:
            assert np.array_equal(out[key], ref[key]), (layout, key)
        assert (st["rays"], st["df_fetches"], st["vox_fetches"]) == (rst["rays"], rst["df_fetches"], rst["vox_fetches"])
        for pos in listener_positions(worlds["city"], 10, 8) + [(192.0, 125.0, 192.0), (500.0, 80.0, 100.0)]:
            for frame in (0, 3, 512 + 77):
                agg, per, _ = o.ambient_sound(pos, frame)
                gagg, gper = renderer.estimate_ambient_sound(pos, frame)
                assert gagg == agg and np.array_equal(gper, per), (layout, pos, frame)
        for c in golden["ambient"]["city"]:
            gagg, gper = renderer.estimate_ambient_sound(c["pos"], c["frame"])
            assert gagg == c["aggregate"] and gper.tolist() == c["per_invocation"]
    renderer.set_option(abi.OPT_TRAVERSAL_LAYOUT, 1)
    sun = scene_tables["sun"]
    for pos in [(9.5, 43.5, 9.5), (9.5, 125.0, 9.5), (100.0, 60.0, 100.0)]:
        assert renderer.player_shadowed(pos, sun) == o.player_shadowed(pos, sun)
    # device-resident ray buffers take the zero-copy path
    import torch
    dev = torch.device("cuda:0")
    to, td = torch.from_numpy(origins).to(dev), torch.from_numpy(dirs).to(dev)
    t = torch.empty(origins.shape[0], dtype=torch.float32, device=dev)
    abi.check(renderer.lib.vxpt_trace_rays(renderer.handle, to.data_ptr(), td.data_ptr(), origins.shape[0], 350, t.data_ptr(), None, None, None))
    renderer.sync()
    assert np.array_equal(t.cpu().numpy(), ref["t"])
