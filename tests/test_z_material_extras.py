"""GPU tests of the material pass that were written after the round's GPU budget was spent (vxpt_render_frame with the material pass,
relief parallax mapping): kept in a file that sorts after the GPU-verified ones, so that `pytest -x` reaches those first.  They do run
against the emulated ABI in every CPU run (tests/test_host_emulation.py)."""
import numpy as np
import pytest

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import abi, camera

import material_cases as mc
from test_material_pass import _params, gb_oracles, gb_renderer, mips  # noqa: F401  (fixtures)


@pytest.mark.gpu
def test_gpu_render_frame_runs_the_material_pass_and_feeds_the_reflections(gb_renderer, worlds, scene_tables):
    """vxpt_render_frame with VxFrameParams.material: primary -> material pass -> shadow -> GI -> reflections in one call, the reflection pass
    reading the material pass's normal / pbr planes from device memory — equal to the separate calls chained by hand.  Host planes."""
    r = gb_renderer
    r.upload_world(worlds["gi_box"])
    r.build_distance_field()
    W, H = 160, 90
    fc = camera.FpsCamera(pitch_deg=-20.0, aspect=W / H)
    cam = fc.vx_camera(W, H)
    sun, moon, stronger, vis = (scene_tables[k] for k in ("sun", "moon", "stronger", "sun_visibility"))
    pp, mp = vx.primary_params(350), _params(scene_tables)
    sp, dp = vx.shadow_params(stronger, frame=2, soft=True), vx.diffuse_params(sun, moon, vis, spp=1, frame=2)
    rp = vx.reflection_params(sun, moon, stronger, fc.position, scene_tables["materials"]["grass_props"], spp=2, rough=True, frame=2)
    g = r.trace_primary(cam, pp, r.alloc_gbuffer(W, H))
    m = r.generate_gbuffer(cam, g, mp, r.alloc_material(W, H))
    d = r.trace_diffuse(cam, g, dp, r.alloc_diffuse(W, H))
    want = r.trace_reflection(cam, g, d, rp, r.alloc_reflection(W, H), g_normal=m["normal"], g_pbr=m["pbr"])
    g2, m2, d2, s2, got = r.alloc_gbuffer(W, H), r.alloc_material(W, H), r.alloc_diffuse(W, H), r.alloc_shadow(W, H), r.alloc_reflection(W, H)
    r.render_frame(cam, pp, shadow=sp, diffuse=dp, gbuf=g2, shadow_out=s2, diffuse_out=d2, reflection=rp, reflection_out=got, material=mp, material_out=m2)
    for k in mc.PLANES:
        assert np.array_equal(m2[k], m[k]), k
    for k in ("color", "hit_distance", "emissive_mask"):
        assert np.array_equal(got[k], want[k], equal_nan=True), k
    # the caller need not take the material planes back for the reflections to use them
    got3 = r.alloc_reflection(W, H)
    r.render_frame(cam, pp, diffuse=dp, reflection=rp, reflection_out=got3, material=mp)
    assert np.array_equal(got3["color"], want["color"], equal_nan=True)
    # without the material pass (and without caller planes) the reflections fall back to face normals: a different picture
    got4 = r.alloc_reflection(W, H)
    r.render_frame(cam, pp, diffuse=dp, reflection=rp, reflection_out=got4)
    assert not np.array_equal(got4["color"], want["color"], equal_nan=True)
    odd = fc.vx_camera(W, H, 1, H)
    with pytest.raises(abi.VxptError) as e:
        r.render_frame(odd, pp, material=mp, material_out=r.alloc_material(W, H))
    assert e.value.code == abi.E_INVALID


@pytest.mark.gpu
def test_gpu_render_frame_in_row_slabs_traces_the_reflection_halo_rows(gb_renderer, worlds, scene_tables):
    """The reflection pass reads the primary distance / normal id at the Halton-jittered coordinate (ReflectionTraceFrag.glsl:754-777), i.e.
    rows next to its slab: vxpt_render_frame traces those halo rows itself, so a frame rendered in row slabs — host planes,
    one call per slab, each into fresh arena contents — equals the frame rendered in one piece.  Jitter of both signs, wrap at the frame edge."""
    r = gb_renderer
    r.upload_world(worlds["gi_box"])
    r.build_distance_field()
    W, H = 160, 96
    fc = camera.FpsCamera(pitch_deg=-20.0, aspect=W / H)
    sun, moon, stronger, vis = (scene_tables[k] for k in ("sun", "moon", "stronger", "sun_visibility"))
    pp, dp = vx.primary_params(350), vx.diffuse_params(sun, moon, vis, spp=1, frame=3)
    for halton in (camera.taa_jitter_secondary(3), (-0.75, -1.5)):
        rp = vx.reflection_params(sun, moon, stronger, fc.position, scene_tables["materials"]["grass_props"], spp=2, rough=True, frame=3, halton=halton)
        want = r.alloc_reflection(W, H)
        r.render_frame(fc.vx_camera(W, H), pp, diffuse=dp, reflection=rp, reflection_out=want)
        got = r.alloc_reflection(W, H)
        for rb, re in ((0, 40), (40, 41), (41, 96)):
            # poison the arena's G-buffer between the slabs: whatever the halo needs must be traced by this call
            r.render_frame(fc.vx_camera(W, H), vx.primary_params(1), diffuse=dp)
            r.render_frame(fc.vx_camera(W, H, rb, re), pp, diffuse=dp, reflection=rp, reflection_out=got)
        for k in ("color", "hit_distance", "emissive_mask"):
            assert np.array_equal(got[k], want[k], equal_nan=True), (halton, k)
    with pytest.raises(abi.VxptError) as e:   # interleaved bands have no contiguous neighbours to read
        cam_il = fc.vx_camera(W, H)
        cam_il.interleave_n, cam_il.interleave_rank, cam_il.band_rows, cam_il.row_end = 2, 0, 8, H // 2
        r.render_frame(cam_il, pp, diffuse=dp, reflection=rp, reflection_out=r.alloc_reflection(W, H))
    assert e.value.code == abi.E_INVALID


@pytest.mark.gpu
def test_gpu_relief_parallax_equals_the_oracle(gb_renderer, worlds, gb_oracles, scene_tables):
    name, idx, kw = mc.POM_CASES[0]
    case = mc.CASES[idx]
    r, o = gb_renderer, gb_oracles[case[1]]
    r.upload_world(worlds[case[1]])
    r.build_distance_field()
    cam = mc.case_camera(case)
    W, H = cam.width, cam.height
    g = r.trace_primary(cam, vx.primary_params(350), r.alloc_gbuffer(W, H))
    g_ref, _ = o.trace_primary(cam, vx.primary_params(350))
    for kw2 in (kw, dict(high_quality_pom=True, dither_pom=False, pom_height=2.0)):
        mp = _params(scene_tables, pom=True, **kw2)
        want = o.generate_gbuffer(cam, g_ref, mp)
        got = r.generate_gbuffer(cam, g, mp, r.alloc_material(W, H))
        for k in mc.PLANES:     # the march compares pow() results against a depth: an ulp can flip a step, so a few more texels may differ
            diff = got[k] != want[k]
            assert diff.mean() <= 2e-3, (k, float(diff.mean()))


@pytest.mark.gpu
def test_gpu_lava_path_equals_the_oracle(gb_renderer, worlds, gb_oracles, scene_tables):
    name, idx, block, kw = mc.LAVA_CASES[0]
    case = mc.CASES[idx]
    r, o = gb_renderer, gb_oracles[case[1]]
    tex = mc.lava_textures()
    r.set_lava_textures(*tex)
    o.set_lava_textures(*tex)
    r.upload_world(worlds[case[1]])
    r.build_distance_field()
    cam = mc.case_camera(case)
    W, H = cam.width, cam.height
    g = r.trace_primary(cam, vx.primary_params(350), r.alloc_gbuffer(W, H))
    g_ref, _ = o.trace_primary(cam, vx.primary_params(350))
    for kw2 in (kw, dict(time=55.5, update_this_frame=False)):
        mp = _params(scene_tables, lava_block_id=block, **kw2)
        seed = mc.seeded_planes(W, H)
        want = o.generate_gbuffer(cam, g_ref, mp, {k: v.copy() for k, v in seed.items()})
        got = r.generate_gbuffer(cam, g, mp, {k: v.copy() for k, v in seed.items()})
        for k in mc.PLANES:     # sin / cos of the distortion are pinned double evaluations: an ulp may differ on isolated texels
            diff = got[k] != want[k]
            assert diff.mean() <= 1e-3 and float(np.abs(got[k].astype(np.float64) - want[k]).max()) <= 1e-4, (k, float(diff.mean()))
