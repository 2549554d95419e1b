"""GPU test of VXPT_OPT_MATERIAL_QUAD_SHUFFLE (the G-buffer pass's opt-in instantiation that takes the quad partners' UV by warp shuffle).
Written after the round's GPU budget was spent: sorts behind every other GPU test."""
import numpy as np
import pytest

import voxelpathtracer_b200 as vx

import material_cases as mc
from test_material_pass import _params, gb_oracles, gb_renderer, mips  # noqa: F401  (fixtures)


@pytest.mark.gpu
def test_gpu_quad_shuffle_variant_writes_the_same_planes(gb_renderer, worlds, scene_tables):
    """VXPT_OPT_MATERIAL_QUAD_SHUFFLE: the quad partners' UV by warp shuffle instead of two more ray set-ups - same operands, so the planes
    must be the same BITS as the default kernel's (odd width / height, a row slab, sky quads included)."""
    from voxelpathtracer_b200 import abi
    r = gb_renderer
    for case in (mc.CASES[0], mc.CASES[3]):
        r.upload_world(worlds[case[1]])
        r.build_distance_field()
        cam = mc.case_camera(case)
        W, H = cam.width, cam.height
        g = r.trace_primary(cam, vx.primary_params(350), r.alloc_gbuffer(W, H))
        for rb, re in ((0, H), (2, min(H, 38))):
            cam.row_begin, cam.row_end = rb, re
            seed = mc.seeded_planes(W, H)
            want = r.generate_gbuffer(cam, g, _params(scene_tables), {k: v.copy() for k, v in seed.items()})
            r.set_option(abi.OPT_MATERIAL_QUAD_SHUFFLE, 1)
            try:
                got = r.generate_gbuffer(cam, g, _params(scene_tables), {k: v.copy() for k, v in seed.items()})
            finally:
                r.set_option(abi.OPT_MATERIAL_QUAD_SHUFFLE, 0)
            for k in mc.PLANES:
                assert np.array_equal(got[k].view(np.uint32), want[k].view(np.uint32)), (case[0], rb, re, k)
