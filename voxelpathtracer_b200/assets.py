"""Input tables of the trace passes: material table, blue-noise tables, baked material texels, sky, shadow noise.

These are the boundary INPUTS a host application supplies (BlockDataSSBO, BlueNoiseDataSSBO, texture binds of
Core/Pipeline.cpp:2236-2270, 2825-2841).  The package ships its own copies under voxelpathtracer_b200/data/ (generated from the
reference tree by tools/make_fixtures.py: the reference's blue-noise tables, block database, FastNoise plains columns), so worlds and
scene tables can be built wherever the package is installed — tests, the smoke run and the benchmark use the same files.
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")  # (the name predates the move out of tests/golden/)


def load_plains_columns(golden=GOLDEN):
    return np.fromfile(os.path.join(golden, "plains_columns.u8"), dtype=np.uint8)


def load_blue_noise(golden=GOLDEN):
    """(sobol[65536], scramble[131072], rank[131072]) as int32, the SSBO element type."""
    t = np.fromfile(os.path.join(golden, "bluenoise_tables.u8"), dtype=np.uint8).astype(np.int32)
    assert t.size == 65536 + 2 * 131072
    return np.ascontiguousarray(t[:65536]), np.ascontiguousarray(t[65536:65536 + 131072]), np.ascontiguousarray(t[65536 + 131072:])


def load_shadow_noise(golden=GOLDEN):
    return np.fromfile(os.path.join(golden, "shadow_blue_noise.rgba8"), dtype=np.uint8).reshape(256, 256, 4)


def load_minecraft_id_lut(golden=GOLDEN):
    """uint8[256]: Minecraft block id -> engine block id (blockdb.minecraft_id_lut over the reference's blockdb.txt)."""
    return np.fromfile(os.path.join(golden, "mcid_lut.u8"), dtype=np.uint8)


def load_materials(golden=GOLDEN):
    """dict(table int32[768], albedo_lod3 f32[L,64,64,4], pbr_lod2 f32[L,128,128,4], emissive_lod0 f32[E,512,512])."""
    z = np.load(os.path.join(golden, "materials.npz"))

    def rgba(a):
        out = np.ones(a.shape[:-1] + (4,), dtype=np.float32)
        out[..., :3] = a.astype(np.float32) / np.float32(255.0)
        return np.ascontiguousarray(out)

    return {
        "table": np.ascontiguousarray(z["table"].astype(np.int32).reshape(768)),
        "albedo_lod3": rgba(z["albedo_lod3"]),
        "pbr_lod2": rgba(z["pbr_lod2"]),
        "emissive_lod0": np.ascontiguousarray(z["emissive_lod0"].astype(np.float32) / np.float32(255.0)),
        "grass_props": z["grass_props"].astype(np.int32),
        "normal_lod3": rgba(z["normal_lod3"]),
        "emissive_lod2": np.ascontiguousarray(z["emissive_lod2"].astype(np.float32) / np.float32(255.0)),
    }


def constant_materials(n_blocks=128):
    """One flat-coloured layer per block id (BASELINE.md config 3's fallback when no textures are loaded)."""
    rng = np.random.RandomState(7)
    cols = (0.2 + 0.6 * rng.rand(n_blocks, 3)).astype(np.float32)
    albedo = np.ones((n_blocks, 64, 64, 4), np.float32)
    albedo[..., :3] = cols[:, None, None, :]
    pbr = np.ones((n_blocks, 128, 128, 4), np.float32)
    pbr[..., 0] = 0.8
    pbr[..., 1] = 0.0
    table = np.zeros((6, 128), np.int32)
    table[0] = np.arange(128)
    table[1] = np.arange(128)
    table[2] = np.arange(128)
    table[3] = -1
    normal = np.ones((n_blocks, 64, 64, 4), np.float32)
    normal[..., :3] = (0.5, 0.5, 1.0)
    return {"table": table.reshape(768), "albedo_lod3": albedo, "pbr_lod2": pbr, "emissive_lod0": np.zeros((0, 512, 512), np.float32),
            "grass_props": np.zeros(10, np.int32), "normal_lod3": normal, "emissive_lod2": np.zeros((0, 128, 128), np.float32)}


def alpha_mip_pyramid(alpha_lod0):
    """uint8 [L][512][512] level-0 alpha -> uint8 [L][ALPHA_MIP_TEXELS]: levels 0..8 back to back, each level the 2x2 box filter
    of the one above rounded to nearest (ties away from zero, like most GL drivers' glGenerateMipmap on unorm8).  GL leaves the
    filter to the driver; whatever pyramid the host's driver produced is what it hands to vxpt_set_albedo_alpha_mips."""
    a = np.ascontiguousarray(alpha_lod0, dtype=np.uint8)
    assert a.ndim == 3 and a.shape[1:] == (512, 512), a.shape
    levels, cur = [a.reshape(a.shape[0], -1)], a.astype(np.uint16)
    for _ in range(8):
        cur = (cur[:, 0::2, 0::2] + cur[:, 1::2, 0::2] + cur[:, 0::2, 1::2] + cur[:, 1::2, 1::2] + 2) >> 2
        levels.append(cur.astype(np.uint8).reshape(a.shape[0], -1))
    return np.ascontiguousarray(np.concatenate(levels, axis=1))


def synthetic_alpha_lod0(n_layers, cutout_layers, seed=11):
    """Level-0 alpha for tests: opaque (255) everywhere except `cutout_layers`, which get a leaf-like cut-out pattern
    (blobs of alpha 0 covering about 40 % of the tile) so that the alpha test passes some texels and stops others."""
    a = np.full((n_layers, 512, 512), 255, np.uint8)
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:512, 0:512]
    for layer in cutout_layers:
        hole = np.zeros((512, 512), bool)
        for _ in range(60):
            cx, cy, r = rng.randint(0, 512), rng.randint(0, 512), rng.randint(12, 44)
            dx = np.minimum(np.abs(xx - cx), 512 - np.abs(xx - cx))   # toroidal: the array wraps with GL_REPEAT
            dy = np.minimum(np.abs(yy - cy), 512 - np.abs(yy - cy))
            hole |= dx * dx + dy * dy < r * r
        a[layer][hole] = 0
    return a


def _srgb_decode(u8):
    c = u8.astype(np.float64) / 255.0
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


def _srgb_encode(lin):
    lin = np.clip(lin, 0.0, 1.0)
    return np.where(lin <= 0.0031308, lin * 12.92, 1.055 * lin ** (1.0 / 2.4) - 0.055)


def rgba_mip_chain(lod0, srgb=False):
    """uint8 [L][512][512][4] level-0 texels -> uint8 [L][MIP_CHAIN_TEXELS][4]: mip levels 0..9 back to back, the layout
    vxpt_set_gbuffer_textures takes.  Each level is the 2x2 box filter of the one above, rounded to nearest; for an sRGB array the rgb
    is filtered in linear space and re-encoded (alpha stays linear).  GL leaves glGenerateMipmap's filter to the driver: a host
    application hands over whatever pyramid its driver built (glGetTexImage per level); this helper builds one for tests and benches."""
    a = np.ascontiguousarray(lod0, dtype=np.uint8)
    assert a.ndim == 4 and a.shape[1:] == (512, 512, 4), a.shape
    levels = [a.reshape(a.shape[0], -1, 4)]
    cur = a.astype(np.float64) / 255.0
    if srgb:
        cur[..., :3] = _srgb_decode(a[..., :3])
    for _ in range(9):
        cur = 0.25 * (cur[:, 0::2, 0::2] + cur[:, 1::2, 0::2] + cur[:, 0::2, 1::2] + cur[:, 1::2, 1::2])
        enc = cur.copy()
        if srgb:
            enc[..., :3] = _srgb_encode(cur[..., :3])
        levels.append(np.floor(enc * 255.0 + 0.5).astype(np.uint8).reshape(a.shape[0], -1, 4))
    return np.ascontiguousarray(np.concatenate(levels, axis=1))


def synthetic_material_lod0(n_layers, seed=23):
    """Level-0 RGBA8 texels (albedo sRGB, normal map, PBR) for the G-buffer material pass, [n_layers][512][512][4] each: per layer a
    tiling pattern with detail at several scales (so that every mip level differs from its neighbours), deterministic in `seed`.
    Stand-in for the reference's Res/Block textures, which are too large to commit at full resolution."""
    rng = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:512, 0:512].astype(np.float64) / 512.0
    albedo = np.zeros((n_layers, 512, 512, 4), np.uint8)
    normal = np.zeros_like(albedo)
    pbr = np.zeros_like(albedo)
    for layer in range(n_layers):
        h = np.zeros((512, 512))
        for octave in range(1, 7):
            f = 2 ** octave
            px, py = rng.rand(2) * 2.0 * np.pi
            h += (np.sin(2 * np.pi * f * xx + px) * np.cos(2 * np.pi * f * yy + py) + 0.5 * np.sin(2 * np.pi * f * (xx + yy) + px * py)) / octave
        h = (h - h.min()) / (h.max() - h.min())
        noise = rng.rand(512, 512)
        base = 0.25 + 0.6 * rng.rand(3)
        for k in range(3):
            albedo[layer, ..., k] = np.clip((base[k] * (0.55 + 0.45 * h) + 0.12 * (noise - 0.5)) * 255.0, 0, 255).astype(np.uint8)
        albedo[layer, ..., 3] = 255
        gy, gx = np.gradient(h)
        n = np.stack([-gx * 40.0, -gy * 40.0, np.ones_like(h)], -1)
        n /= np.linalg.norm(n, axis=-1, keepdims=True)
        normal[layer, ..., :3] = np.clip((n * 0.5 + 0.5) * 255.0 + 0.5, 0, 255).astype(np.uint8)
        normal[layer, ..., 3] = 255
        pbr[layer, ..., 0] = np.clip((0.35 + 0.5 * h + 0.1 * (noise - 0.5)) * 255.0, 0, 255).astype(np.uint8)  # roughness
        pbr[layer, ..., 1] = np.clip((rng.rand() < 0.3) * (0.6 + 0.4 * noise) * 255.0, 0, 255).astype(np.uint8)  # metalness
        pbr[layer, ..., 2] = np.clip(h * 255.0, 0, 255).astype(np.uint8)                                          # displacement
        pbr[layer, ..., 3] = np.clip((0.6 + 0.4 * h) * 255.0, 0, 255).astype(np.uint8)                             # texture AO
    return albedo, normal, pbr


def synthetic_lava_textures(seed=31):
    """(albedo, normal) uint8 [8][256][256][4]: stand-ins for the reference's animated lava frames (Res/Block/Lava/Frames), smooth in all
    three axes and periodic (the textures wrap with GL_REPEAT, the frame axis too)."""
    rng = np.random.RandomState(seed)
    k, y, x = np.mgrid[0:8, 0:256, 0:256].astype(np.float64)
    h = np.zeros((8, 256, 256))
    for octave in (1, 2, 4, 8):
        ph = rng.rand(3) * 2 * np.pi
        h += np.sin(2 * np.pi * octave * x / 256 + ph[0] + 2 * np.pi * k / 8) * np.cos(2 * np.pi * octave * y / 256 + ph[1]) / octave
        h += 0.5 * np.sin(2 * np.pi * (octave * (x + y) / 256 + k / 8) + ph[2]) / octave
    h = (h - h.min()) / (h.max() - h.min())
    albedo = np.zeros((8, 256, 256, 4), np.uint8)
    albedo[..., 0] = np.clip(180 + 75 * h, 0, 255)
    albedo[..., 1] = np.clip(40 + 160 * h ** 2, 0, 255)
    albedo[..., 2] = np.clip(10 + 40 * h ** 4, 0, 255)
    albedo[..., 3] = 255
    gy, gx = np.gradient(h, axis=(1, 2))
    n = np.stack([-gx * 30.0, -gy * 30.0, np.ones_like(h)], -1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    normal = np.zeros_like(albedo)
    normal[..., :3] = np.clip((n * 0.5 + 0.5) * 255.0 + 0.5, 0, 255).astype(np.uint8)
    normal[..., 3] = 255
    return albedo, normal


def analytic_sky(n=16, sun_dir=(-0.669, 0.468, 0.577)):
    """Documented stand-in for the reference's rendered atmosphere cubemap (RGB16F, 16^2 for GI; Pipeline.cpp:1392-1394):
    a horizon-to-zenith gradient plus a broad sun lobe.  Faces +X,-X,+Y,-Y,+Z,-Z, GL cube-map face orientation,
    array [6][n][n][3] float32 with row index = t, column index = s."""
    sun = np.asarray(sun_dir, dtype=np.float64)
    sun = sun / np.linalg.norm(sun)
    out = np.zeros((6, n, n, 3), dtype=np.float32)
    c = (np.arange(n) + 0.5) / n * 2.0 - 1.0
    sc, tc = np.meshgrid(c, c)  # tc varies along rows
    one = np.ones_like(sc)
    dirs = [
        (one, -tc, -sc),   # +X: sc = -z, tc = -y
        (-one, -tc, sc),   # -X: sc = +z, tc = -y
        (sc, one, tc),     # +Y: sc = +x, tc = +z
        (sc, -one, -tc),   # -Y: sc = +x, tc = -z
        (sc, -tc, one),    # +Z: sc = +x, tc = -y
        (-sc, -tc, -one),  # -Z: sc = -x, tc = -y
    ]
    zenith = np.array([0.18, 0.36, 0.85])
    horizon = np.array([0.75, 0.82, 0.95])
    ground = np.array([0.12, 0.11, 0.10])
    for f, (x, y, z) in enumerate(dirs):
        d = np.stack([x, y, z], -1)
        d = d / np.linalg.norm(d, axis=-1, keepdims=True)
        up = np.clip(d[..., 1], 0.0, 1.0) ** 0.5
        col = horizon[None, None, :] * (1.0 - up[..., None]) + zenith[None, None, :] * up[..., None]
        down = np.clip(-d[..., 1] * 4.0, 0.0, 1.0)
        col = col * (1.0 - down[..., None]) + ground[None, None, :] * down[..., None]
        lobe = np.clip((d * sun[None, None, :]).sum(-1), 0.0, 1.0) ** 8
        col = col + lobe[..., None] * np.array([1.6, 1.4, 1.1])[None, None, :]
        out[f] = col.astype(np.float32)
    return out
