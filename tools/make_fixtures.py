#!/usr/bin/env python3
"""Generate the golden INPUT fixtures under tests/golden/ from the read-only reference tree.

Run once in the build container (the GPU box has no /root/reference); the outputs are committed.

  plains_columns.u8      per-column (height, biome) of the reference's "plains" terrain, produced by
                         oracle/_ref/ref_worldgen = the reference's vendored FastNoise + the call sequence of
                         Core/WorldGenerator.cpp:71-107 (glibc rand() seeds 9383 / 6886).
  bluenoise_tables.u8    sobol_256spp_256d | scramblingTile | rankingTile of Core/BlueNoiseDataSSBO.cpp:4,9,14
                         (all values < 256, stored as bytes; 327,680 B).
  shadow_blue_noise.rgba8  Res/Misc/blue_noise.png, 256x256 RGBA8 rows top-to-bottom as stored in the PNG.
  materials.npz          BlockDataSSBO-style table (6 x 128 int32, Core/BlockDataSSBO.cpp:15-35, Front face)
                         for a subset of blockdb.txt, with layer indices re-based onto compact baked arrays:
                         albedo level 3 (64^2, sRGB-decoded, 2x2 box filter in linear space applied 3 times),
                         PBR level 2 (128^2, box filter twice), emissive level 0 red channel (512^2); for the reflection pass
                         also normal maps level 3 (64^2) and emissive level 2 (128^2).
                         Quantised to u8 to keep the fixture small; both sides of every parity check read the
                         same bytes.  `block_names` maps id -> name for all 99 blocks of blockdb.txt.
"""
import os
import re
import subprocess
import sys

import numpy as np
from PIL import Image

REF = os.environ.get("VXPT_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "voxelpathtracer_b200", "data")  # input assets ship with the package
sys.path.insert(0, ROOT)

# blocks whose textures are baked (ids follow blockdb.txt order, Core/BlockDatabaseParser.cpp:31-42)
SUBSET = ["Grass", "Dirt", "Stone", "Cobblestone", "Sand", "Lamp", "Glowstone", "Bricks", "Planks", "oak_leaves"]


from voxelpathtracer_b200.blockdb import minecraft_id_lut, parse_blockdb  # noqa: E402  (the parser is product code now)


def srgb_to_linear(c):
    c = c.astype(np.float64) / 255.0
    return np.where(c <= 0.04045, c / 12.92, ((c + 0.055) / 1.055) ** 2.4)


def box(img):
    return 0.25 * (img[0::2, 0::2] + img[1::2, 0::2] + img[0::2, 1::2] + img[1::2, 1::2])


def load_rgba(path):
    im = Image.open(os.path.join(REF, path)).convert("RGBA")
    if im.size != (512, 512):
        im = im.resize((512, 512), Image.BILINEAR)
    return np.asarray(im, dtype=np.uint8)


def main():
    os.makedirs(OUT, exist_ok=True)
    # --- plains terrain through the reference's own noise library
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    subprocess.check_call([os.path.join(ROOT, "oracle", "_ref", "ref_worldgen"), os.path.join(OUT, "plains_columns.u8")])

    # --- blue-noise SSBO tables
    src = open(os.path.join(REF, "Core", "BlueNoiseDataSSBO.cpp")).read()
    tabs = []
    for name, n in (("sobol_256spp_256d", 65536), ("scramblingTile", 131072), ("rankingTile", 131072)):
        m = re.search(name + r"\s*=\s*\{([^}]*)\}", src)
        vals = np.array([int(v) for v in m.group(1).split(",") if v.strip()], dtype=np.int64)
        assert vals.size == n and vals.min() >= 0 and vals.max() < 256, (name, vals.size, vals.min(), vals.max())
        tabs.append(vals.astype(np.uint8))
    np.concatenate(tabs).tofile(os.path.join(OUT, "bluenoise_tables.u8"))

    # --- shadow jitter texture
    bn = np.asarray(Image.open(os.path.join(REF, "Res", "Misc", "blue_noise.png")).convert("RGBA"), dtype=np.uint8)
    assert bn.shape == (256, 256, 4)
    bn.tofile(os.path.join(OUT, "shadow_blue_noise.rgba8"))
    # Minecraft id -> engine block id (256 bytes), for the .mca importer
    minecraft_id_lut(parse_blockdb(os.path.join(REF, "blockdb.txt"))).tofile(os.path.join(OUT, "mcid_lut.u8"))

    # --- material table + baked texel arrays
    blocks = parse_blockdb(os.path.join(REF, "blockdb.txt"))
    names = {b["Name"]: b for b in blocks}
    subset = [names[n] for n in SUBSET if n in names]
    albedo_paths = sorted({b["faces"]["Albedo"][f] for b in subset for f in ("front", "top", "bottom")})
    pbr_paths = sorted({b["faces"]["PBR"][f] for b in subset for f in ("front", "top", "bottom")})
    normal_paths = sorted({b["faces"]["Normal"][f] for b in subset for f in ("front", "top", "bottom")})
    emissive_paths = sorted({b["Emissive"] for b in subset if b["Emissive"]})
    # The reference indexes the PBR array with the ALBEDO layer (DiffuseRayTraceFrag.glsl:578).  In its full
    # arrays albedo and PBR paths sort identically (one of each per texture directory), so both layers agree;
    # keep that property in the compact arrays by baking PBR layers in albedo-path order.
    pbr_for_albedo = [p.replace("Albedo.png", "PBR.png") for p in albedo_paths]
    table = np.zeros((6, 128), dtype=np.int32)
    table[0:3, :] = -1
    table[3, :] = -1
    for b in blocks:
        i = b["ID"]
        if b in subset:
            table[0, i] = albedo_paths.index(b["faces"]["Albedo"]["front"])
            table[1, i] = normal_paths.index(b["faces"]["Normal"]["front"])
            table[2, i] = pbr_for_albedo.index(b["faces"]["PBR"]["front"]) if b["faces"]["PBR"]["front"] in pbr_for_albedo else 0
            table[3, i] = emissive_paths.index(b["Emissive"]) if b["Emissive"] else -1
        else:
            table[0:3, i] = 0
        table[4, i] = 1 if b["Transparent"] else 0
        table[5, i] = 1 if b["SSS"] else 0
    alb = []
    for p in albedo_paths:
        lin = srgb_to_linear(load_rgba(p)[..., :3])
        for _ in range(3):
            lin = box(lin)
        alb.append(np.clip(np.rint(lin * 255.0), 0, 255).astype(np.uint8))
    pbr = []
    for p in pbr_for_albedo:
        if not os.path.exists(os.path.join(REF, p)):
            p = pbr_paths[0]
        v = load_rgba(p)[..., :3].astype(np.float64) / 255.0
        for _ in range(2):
            v = box(v)
        pbr.append(np.clip(np.rint(v * 255.0), 0, 255).astype(np.uint8))
    emi = [load_rgba(p)[..., 0].copy() for p in emissive_paths]
    nrm = []
    for p in normal_paths:  # normal maps: level 3 (ReflectionTraceFrag.glsl:961), plain box filter of the stored values
        v = load_rgba(p)[..., :3].astype(np.float64) / 255.0
        for _ in range(3):
            v = box(v)
        nrm.append(np.clip(np.rint(v * 255.0), 0, 255).astype(np.uint8))
    emi2 = []
    for p in emissive_paths:  # emissive: level 2 red channel (ReflectionTraceFrag.glsl:973)
        v = load_rgba(p)[..., 0].astype(np.float64) / 255.0
        for _ in range(2):
            v = box(v)
        emi2.append(np.clip(np.rint(v * 255.0), 0, 255).astype(np.uint8))
    grass = names["Grass"]
    # u_GrassBlockProps, Core/Pipeline.cpp:3040-3049: id, top(a,n,p), side/front(a,n,p), bottom(a,n,p)
    grass_props = np.array([
        grass["ID"],
        albedo_paths.index(grass["faces"]["Albedo"]["top"]), normal_paths.index(grass["faces"]["Normal"]["top"]),
        pbr_for_albedo.index(grass["faces"]["PBR"]["top"]),
        albedo_paths.index(grass["faces"]["Albedo"]["front"]), normal_paths.index(grass["faces"]["Normal"]["front"]),
        pbr_for_albedo.index(grass["faces"]["PBR"]["front"]),
        albedo_paths.index(grass["faces"]["Albedo"]["bottom"]), normal_paths.index(grass["faces"]["Normal"]["bottom"]),
        pbr_for_albedo.index(grass["faces"]["PBR"]["bottom"]),
    ], dtype=np.int32)
    np.savez_compressed(
        os.path.join(OUT, "materials.npz"),
        table=table,
        albedo_lod3=np.stack(alb), pbr_lod2=np.stack(pbr),
        emissive_lod0=np.stack(emi) if emi else np.zeros((0, 512, 512), np.uint8),
        normal_lod3=np.stack(nrm), emissive_lod2=np.stack(emi2) if emi2 else np.zeros((0, 128, 128), np.uint8),
        albedo_paths=np.array(albedo_paths), emissive_paths=np.array(emissive_paths),
        block_names=np.array([""] + [b["Name"] for b in blocks]),
        grass_props=grass_props,
    )
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
    print("ids:", {n: names[n]["ID"] for n in SUBSET if n in names})


if __name__ == "__main__":
    sys.exit(main())
