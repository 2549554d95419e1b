"""The distance-field build also writes the traversal's step field (df_z_dpx converts the finished words two voxels per table read,
csrc/df_build.cu).  GPU: the DPX build (VXPT_OPT_DF_ALGO = 1, default) against the reference-shaped build + pack_steps (= 0) — the distance
field equals the oracle's bit for bit and every pass that walks the step field writes the same planes, in both layouts, also after a
re-layout.  CPU: the property the two-voxel table rests on."""
import numpy as np
import pytest

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import abi, camera


def test_x_neighbours_of_a_distance_field_differ_by_at_most_one(oracle_dfs):
    """The pair table of df_z_dpx is indexed by 2 * M0 + M1 + 1 = 3 * M0 + (M1 - M0 + 1): unique iff |M1 - M0| <= 1, which holds because the
    field is 1-Lipschitz in the L1 metric (also under the clamp at 254)."""
    for name, df in oracle_dfs.items():
        d = df.reshape(abi.WORLD_SIZE_Z, abi.WORLD_SIZE_Y, abi.WORLD_SIZE_X).astype(np.int16)
        assert np.abs(np.diff(d, axis=2)).max() <= 1, name
    # the table as init_df_kernels builds it equals the per-voxel definition on every reachable pair
    m = np.arange(256)
    lut = np.where(m == 1, 1, np.floor(m.astype(np.float32) * np.float32(0.57735026918)).astype(np.int64))
    for m0 in range(255):
        for m1 in (m0 - 1, m0, m0 + 1):
            if 0 <= m1 <= 254:
                i = 2 * m0 + m1 + 1
                assert i < 768 and i // 3 == m0 and m0 + i % 3 - 1 == m1
                assert (lut[i // 3] | (lut[min(max(i // 3 + i % 3 - 1, 0), 255)] << 8)) == (lut[m0] | (lut[m1] << 8))


def _frame(r, scene_tables, cam, W, H):
    g = r.trace_primary(cam, vx.primary_params(350), r.alloc_gbuffer(W, H))
    s = r.trace_shadow(cam, g, vx.shadow_params(scene_tables["stronger"], frame=5, soft=True), r.alloc_shadow(W, H))
    sun, moon, vis = scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"]
    d = r.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=4), r.alloc_diffuse(W, H))
    out = {}
    for name, planes in (("g", g), ("s", s), ("d", d)):
        for k, v in planes.items():
            if v is not None:
                out[name + "." + k] = np.array(v, copy=True)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["plains", "city", "sparse"])
def test_gpu_fused_step_field_build_equals_the_reference_shaped_build(renderer, worlds, oracle_dfs, scene_tables, name):
    r = renderer
    W, H = 960, 540
    cams = [camera.FpsCamera(pitch_deg=-20.0).vx_camera(W, H), camera.FpsCamera(position=(100.0, 70.0, 100.0), pitch_deg=-35.0, yaw_deg=45.0).vx_camera(W, H)]
    try:
        for layout in (1, 0):
            r.set_option(abi.OPT_TRAVERSAL_LAYOUT, layout)
            r.set_option(abi.OPT_DF_ALGO, 0)
            r.upload_world(worlds[name])
            n0 = r.launch_count()
            r.build_distance_field()
            assert r.launch_count() - n0 == 4                        # three line sweeps + pack_steps
            want = [_frame(r, scene_tables, cam, W, H) for cam in cams]
            r.set_option(abi.OPT_DF_ALGO, 1)
            r.upload_world(worlds[name])
            n0 = r.launch_count()
            r.build_distance_field()
            assert r.launch_count() - n0 == 2                        # xy sweep, z sweep + step field
            assert np.array_equal(r.download_distance_field(), oracle_dfs[name])
            got = [_frame(r, scene_tables, cam, W, H) for cam in cams]
            for a, b in zip(got, want):
                assert a.keys() == b.keys()
                for k in a:
                    assert np.array_equal(a[k], b[k], equal_nan=True), (name, layout, k)
            # switching the layout after a build re-packs from the distance field
            r.set_option(abi.OPT_TRAVERSAL_LAYOUT, 1 - layout)
            r.set_option(abi.OPT_TRAVERSAL_LAYOUT, layout)
            again = _frame(r, scene_tables, cams[0], W, H)
            for k in again:
                assert np.array_equal(again[k], want[0][k], equal_nan=True), (name, layout, "re-layout", k)
    finally:
        r.set_option(abi.OPT_DF_ALGO, 1)
        r.set_option(abi.OPT_TRAVERSAL_LAYOUT, 1)
