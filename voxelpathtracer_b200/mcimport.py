"""Minecraft (Anvil, pre-1.13) world import: region files -> the 384 x 128 x 384 uint8 grid (Core/NBT/Importer.cpp:66-166 over the
vendored enkiMI reader).  Host-side ingestion next to the hot path (BASELINE config 5).

Semantics kept from the reference: every chunk of every r.X.Z.mca file in the directory; a voxel is imported only when its data nibble
is 0 (Importer.cpp:124-127) and its Minecraft id maps to a non-zero engine id (blockdb.minecraft_id_lut); world position = voxel -
ivec3(origin) + (192, 0, 192), dropped outside the grid (WriteVoxel, :66-84)."""
import os
import struct
import zlib

import numpy as np

from .abi import WORLD_SIZE_X, WORLD_SIZE_Y, WORLD_SIZE_Z
from .world import World

SECTOR = 4096


def _parse_nbt(buf):
    """Minimal NBT reader (big-endian, uncompressed payload) -> nested dicts / lists; byte arrays come back as bytes."""
    pos = 0

    def rd(fmt):
        nonlocal pos
        v = struct.unpack_from(fmt, buf, pos)
        pos += struct.calcsize(fmt)
        return v[0]

    def name():
        nonlocal pos
        n = rd(">H")
        s = buf[pos:pos + n].decode("utf-8", "replace")
        pos += n
        return s

    def payload(t):
        nonlocal pos
        if t == 1:
            return rd(">b")
        if t == 2:
            return rd(">h")
        if t == 3:
            return rd(">i")
        if t == 4:
            return rd(">q")
        if t == 5:
            return rd(">f")
        if t == 6:
            return rd(">d")
        if t == 7:
            n = rd(">i")
            v = buf[pos:pos + n]
            pos += n
            return v
        if t == 8:
            return name()
        if t == 9:
            et, n = rd(">b"), rd(">i")
            return [payload(et) for _ in range(n)]
        if t == 10:
            d = {}
            while True:
                tt = rd(">b")
                if tt == 0:
                    return d
                k = name()
                d[k] = payload(tt)
        if t == 11:
            n = rd(">i")
            v = struct.unpack_from(f">{n}i", buf, pos)
            pos += 4 * n
            return v
        if t == 12:
            n = rd(">i")
            v = struct.unpack_from(f">{n}q", buf, pos)
            pos += 8 * n
            return v
        raise ValueError(f"unknown NBT tag {t}")

    t = rd(">b")
    name()
    return payload(t)


def read_region_chunks(path):
    """Yield the decoded NBT root of every chunk present in a region file (location table: 1024 x (3-byte sector offset, 1-byte count))."""
    data = open(path, "rb").read()
    if len(data) < 2 * SECTOR:
        return
    for loc in struct.unpack(">1024I", data[:SECTOR]):
        off, count = (loc >> 8) * SECTOR, loc & 0xFF
        if off == 0 or count == 0 or off + 5 > len(data):
            continue
        length, comp = struct.unpack_from(">IB", data, off)
        raw = data[off + 5:off + 4 + length]
        try:
            yield _parse_nbt(zlib.decompress(raw) if comp == 2 else (zlib.decompress(raw, 31) if comp == 1 else raw))
        except (zlib.error, struct.error, ValueError):
            continue  # enkiMI skips chunks it cannot read


def import_region_file(path, grid_zyx, origin, lut):
    """ImportRegionFile (Importer.cpp:86-147) into a [z][y][x] view of the grid."""
    ox, oy, oz = (int(v) for v in origin)  # glm::ivec3(ImportOrigin): truncation
    for root in read_region_chunks(path):
        level = root.get("Level", root)
        cx, cz = level.get("xPos"), level.get("zPos")
        if cx is None or cz is None:
            continue
        for sec in level.get("Sections", []):
            blocks, data, sy = sec.get("Blocks"), sec.get("Data"), sec.get("Y")
            if blocks is None or sy is None or len(blocks) != 4096 or not 0 <= sy < 16:
                continue
            ids = np.frombuffer(blocks, dtype=np.uint8).reshape(16, 16, 16)           # [y][z][x]
            if data is not None and len(data) == 2048:
                nib = np.frombuffer(data, dtype=np.uint8)
                dv = np.empty(4096, dtype=np.uint8)
                dv[0::2], dv[1::2] = nib & 0x0F, nib >> 4
                dv = dv.reshape(16, 16, 16)
            else:
                dv = np.zeros((16, 16, 16), dtype=np.uint8)
            vox = np.where(dv == 0, lut[ids], 0).astype(np.uint8)
            # world coordinates of the section's voxels
            wx0, wy0, wz0 = cx * 16 - ox + WORLD_SIZE_X // 2, sy * 16 - oy, cz * 16 - oz + WORLD_SIZE_Z // 2
            xs, ys, zs = max(0, -wx0), max(0, -wy0), max(0, -wz0)
            xe, ye, ze = min(16, WORLD_SIZE_X - wx0), min(16, WORLD_SIZE_Y - wy0), min(16, WORLD_SIZE_Z - wz0)
            if xs >= xe or ys >= ye or zs >= ze:
                continue
            src = vox[ys:ye, zs:ze, xs:xe].transpose(1, 0, 2)                        # -> [z][y][x]
            dst = grid_zyx[wz0 + zs:wz0 + ze, wy0 + ys:wy0 + ye, wx0 + xs:wx0 + xe]
            np.copyto(dst, src, where=src != 0)                                      # WriteVoxel ignores voxel == 0


def import_world(directory, origin, lut):
    """ImportWorld (Importer.cpp:149-166): every .mca file of the directory into an empty world."""
    w = World()
    for name in sorted(os.listdir(directory)):
        if name.rsplit(".", 1)[-1] == "mca":
            import_region_file(os.path.join(directory, name), w.zyx, origin, lut)
    return w


# ---- writer (tests and tools: build small region files without any Minecraft asset) ------------------------------------------------
def _nbt_named(tag, name, payload):
    n = name.encode("utf-8")
    return struct.pack(">bH", tag, len(n)) + n + payload


def write_region_file(path, chunks):
    """chunks: {(cx, cz): {section_y: (ids uint8[16][16][16] in [y][z][x] order, data nibbles uint8[16][16][16] or None)}} with absolute
    chunk coordinates inside one region.  Writes a zlib-compressed Anvil region file."""
    table = bytearray(SECTOR)
    body = bytearray()
    sector = 2
    for (cx, cz), sections in sorted(chunks.items()):
        secs = b""
        for sy, (ids, dv) in sorted(sections.items()):
            ids = np.ascontiguousarray(ids, dtype=np.uint8).reshape(4096)
            dvf = np.zeros(4096, np.uint8) if dv is None else np.ascontiguousarray(dv, dtype=np.uint8).reshape(4096)
            nib = (dvf[0::2] & 0x0F) | (dvf[1::2] << 4)
            secs += (_nbt_named(1, "Y", struct.pack(">b", sy)) + _nbt_named(7, "Blocks", struct.pack(">i", 4096) + ids.tobytes()) +
                     _nbt_named(7, "Data", struct.pack(">i", 2048) + nib.astype(np.uint8).tobytes()) + b"\x00")
        level = (_nbt_named(3, "xPos", struct.pack(">i", cx)) + _nbt_named(3, "zPos", struct.pack(">i", cz)) +
                 _nbt_named(9, "Sections", struct.pack(">bi", 10, len(sections)) + secs) + b"\x00")
        root = _nbt_named(10, "", _nbt_named(10, "Level", level) + b"\x00")
        comp = zlib.compress(root)
        blob = struct.pack(">IB", len(comp) + 1, 2) + comp
        n_sec = (len(blob) + SECTOR - 1) // SECTOR
        blob += b"\x00" * (n_sec * SECTOR - len(blob))
        idx = (cx & 31) + (cz & 31) * 32
        struct.pack_into(">I", table, idx * 4, (sector << 8) | n_sec)
        body += blob
        sector += n_sec
    with open(path, "wb") as f:
        f.write(bytes(table) + bytes(SECTOR) + bytes(body))
