"""Frames of the G-buffer material pass shared by tools/make_ref_gbuffer_golden.py, which runs the reference's own
GenerateGBuffer.glsl on them and commits the digests, and tests/test_material_pass.py, which checks the oracle, the kernel source on
the host and the GPU against those digests."""
import hashlib

import numpy as np

from voxelpathtracer_b200 import assets, camera

# name, world, width, height, camera keywords
CASES = [
    ("plains_640x360_p-20", "plains", 640, 360, dict(pitch_deg=-20.0)),
    ("gi_box_512x288_lamps", "gi_box", 512, 288, dict(position=(136.0, 60.5, 12.0), pitch_deg=30.0, yaw_deg=45.0)),
    ("city_480x270_street", "city", 480, 270, dict(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0)),
    ("plains_250x141_close", "plains", 250, 141, dict(position=(192.0, 60.3, 192.0), pitch_deg=-60.0, yaw_deg=10.0)),   # odd height, magnification
    ("superflat_320x180_horizon", "superflat", 320, 180, dict(position=(192.0, 51.0, 192.0), pitch_deg=-2.0)),           # grazing: high mip levels
]
# relief parallax mapping (u_POM): name, index into CASES, material_params keywords
POM_CASES = [
    ("pom_gi_box_lamps_dither_f3", 1, dict(frame=3)),
    ("pom_city_hq_deep", 2, dict(high_quality_pom=True, frame=777, pom_height=2.5, pom_exp=0.5)),
    ("pom_close_nodither_noupdate", 3, dict(dither_pom=False, update_this_frame=False)),   # u_POM keeps the pass alive without u_UpdateGBufferThisFrame
]
# animated lava path (u_LavaBlockID): the test scenes hold no lava, so a block that is in view plays its part.  name, index into CASES,
# block id treated as lava, material_params keywords
LAVA_CASES = [
    ("lava_gi_box_cobble_t12.75", 1, 4, dict(time=12.75)),
    ("lava_city_lamps_pom_t3.1", 2, 12, dict(time=3.1, pom=True, frame=9)),       # the lamps flow, the planks around them get the parallax march
    ("lava_gi_box_only_lava_updates_t100.5", 1, 4, dict(time=100.5, update_this_frame=False)),   # every other fragment discards
]
PLANES = ("albedo", "normal", "pbr", "texture_ao")
_lava = []


def lava_textures():
    if not _lava:
        _lava.append(assets.synthetic_lava_textures())
    return _lava[0]


def seeded_planes(W, H, seed=1):
    """planes with recognisable contents, to see which texels a pass leaves alone"""
    rng = np.random.RandomState(seed)
    return {"albedo": rng.rand(H, W, 3).astype(np.float32), "normal": rng.rand(H, W, 3).astype(np.float32), "pbr": rng.rand(H, W, 4).astype(np.float32),
            "texture_ao": rng.rand(H, W).astype(np.float32)}
_mips = {}


def material_mips(n_layers):
    """(albedo, normal, pbr) RGBA8 mip chains of the synthetic level-0 textures (assets.synthetic_material_lod0, seed 23)."""
    if n_layers not in _mips:
        a, n, p = assets.synthetic_material_lod0(n_layers)
        _mips[n_layers] = (assets.rgba_mip_chain(a, srgb=True), assets.rgba_mip_chain(n), assets.rgba_mip_chain(p))
    return _mips[n_layers]


def case_camera(case):
    _, _, W, H, kw = case
    return camera.FpsCamera(aspect=W / H, **kw).vx_camera(W, H)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
