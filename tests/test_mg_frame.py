"""vxpt_mg_* (csrc/mg.cu): one frame sharded over N devices from one host thread.  The C-level surface SURVEY.md §8(b) lists: scene state
replicated by the same call on every device, rows cut into contiguous slabs, planes gathered without an exchange step.  The tests name
device 0 several times, so they run wherever one GPU exists (and, host planes only, against the emulated ABI in the CPU suite); with two or
more GPUs they also use distinct devices."""
import numpy as np
import pytest

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import abi, camera

pytestmark = pytest.mark.gpu


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 1
    except Exception:
        return 1


def _params(scene_tables, frame):
    sun, moon, stronger, vis = (scene_tables[k] for k in ("sun", "moon", "stronger", "sun_visibility"))
    return (vx.primary_params(350, camera.taa_jitter(frame)), vx.shadow_params(stronger, frame=frame, soft=True),
            vx.diffuse_params(sun, moon, vis, spp=1, frame=frame))


@pytest.mark.parametrize("n", [2, 3])
def test_mg_frame_in_host_planes_equals_one_handle(worlds, scene_tables, n):
    W, H = 256, 136   # 136 rows: slabs of 72 + 64 (n = 2), 48 + 48 + 40 (n = 3)
    ids = [k % _n_gpus() for k in range(n)]
    mg = vx.MultiRenderer(ids)
    one = vx.Renderer(0)
    for r in (mg, one):
        r.load_scene_tables(scene_tables["materials"], scene_tables["blue_noise"], scene_tables["sky"], scene_tables["shadow_noise"])
        r.upload_world(worlds["gi_box"])
        r.build_distance_field()
    fc = camera.FpsCamera(pitch_deg=-20.0, aspect=W / H)
    cam = fc.vx_camera(W, H)
    slabs = [mg.slab(cam, k) for k in range(n)]
    assert slabs[0][0] == 0 and slabs[-1][1] == H and all(a[1] == b[0] for a, b in zip(slabs, slabs[1:])) and all(rb % 8 == 0 for rb, _ in slabs)
    pp, sp, dp = _params(scene_tables, 3)
    mats = scene_tables["materials"]
    rp = vx.reflection_params(scene_tables["sun"], scene_tables["moon"], scene_tables["stronger"], fc.position, mats["grass_props"], spp=1, rough=True, frame=3,
                              halton=camera.taa_jitter_secondary(3))
    want = one.render_frame(cam, pp, shadow=sp, diffuse=dp, gbuf=one.alloc_gbuffer(W, H), shadow_out=one.alloc_shadow(W, H), diffuse_out=one.alloc_diffuse(W, H),
                            reflection=rp, reflection_out=one.alloc_reflection(W, H))
    got = mg.render_frame(cam, pp, shadow=sp, diffuse=dp, gbuf=one.alloc_gbuffer(W, H), shadow_out=one.alloc_shadow(W, H), diffuse_out=one.alloc_diffuse(W, H),
                          reflection=rp, reflection_out=one.alloc_reflection(W, H))
    for a, b in zip(want, got):
        for k in a:
            assert np.array_equal(a[k], b[k], equal_nan=True), k
    # an edit goes to every device; the rebuilt frame differs from the first and again equals the single handle's
    for r in (mg, one):
        r.set_block(192, 70, 200, 12)
        r.build_distance_field()
    want2 = one.render_frame(cam, pp, diffuse=dp, gbuf=one.alloc_gbuffer(W, H), diffuse_out=one.alloc_diffuse(W, H))
    got2 = mg.render_frame(cam, pp, diffuse=dp, gbuf=one.alloc_gbuffer(W, H), diffuse_out=one.alloc_diffuse(W, H))
    assert np.array_equal(want2[0]["t"], got2[0]["t"]) and np.array_equal(want2[2]["sh"], got2[2]["sh"], equal_nan=True)
    assert not np.array_equal(want2[0]["t"], want[0]["t"])
    st = mg.stats()
    assert st["rays"] > 0 and st["rays"] == sum(d.stats()["rays"] for d in mg.devices)
    # part of a frame: rows 40..104 only, the rest of the planes untouched
    part = one.alloc_gbuffer(W, H)
    part["t"][:] = 7.0
    mg.render_frame(fc.vx_camera(W, H, 40, 104), pp, gbuf=part)
    assert np.array_equal(part["t"][40:104], got2[0]["t"][40:104]) and (part["t"][:40] == 7.0).all() and (part["t"][104:] == 7.0).all()
    with pytest.raises(abi.VxptError):
        cam_il = fc.vx_camera(W, H)
        cam_il.interleave_n, cam_il.interleave_rank, cam_il.band_rows, cam_il.row_end = 2, 0, 4, H // 2
        mg.render_frame(cam_il, pp, gbuf=one.alloc_gbuffer(W, H))
    mg.close()
    one.close()


def test_mg_frame_in_device_planes_of_device_0(worlds, scene_tables):
    """Device planes: the kernels of every device store their rows straight into planes that live on device 0 (peer access over NVLink
    when the devices differ)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("device planes need torch's CUDA allocator")
    W, H = 256, 136
    n = 2
    ids = [k % _n_gpus() for k in range(n)]
    mg = vx.MultiRenderer(ids)
    one = vx.Renderer(0)
    for r in (mg, one):
        r.load_scene_tables(scene_tables["materials"], scene_tables["blue_noise"], scene_tables["sky"], scene_tables["shadow_noise"])
        r.upload_world(worlds["plains"])
        r.build_distance_field()
    cam = camera.FpsCamera(pitch_deg=-20.0, aspect=W / H).vx_camera(W, H)
    pp, sp, dp = _params(scene_tables, 5)
    want = one.render_frame(cam, pp, shadow=sp, diffuse=dp, gbuf=one.alloc_gbuffer(W, H, device=True), shadow_out=one.alloc_shadow(W, H, device=True),
                            diffuse_out=one.alloc_diffuse(W, H, device=True))
    one.sync()
    got = mg.render_frame(cam, pp, shadow=sp, diffuse=dp, gbuf=one.alloc_gbuffer(W, H, device=True), shadow_out=one.alloc_shadow(W, H, device=True),
                          diffuse_out=one.alloc_diffuse(W, H, device=True))
    for a, b in zip(want[:3], got[:3]):
        for k in a:
            assert torch.equal(a[k], b[k]), k
    mg.close()
    one.close()
