"""Block database: blockdb.txt -> block ids, BlockDataSSBO rows and the Minecraft-id table (Core/BlockDatabaseParser.cpp:44-375,
Core/BlockDatabase.cpp:93-104, 599-612).  Host-side ingestion next to the hot path: it decides which uint8 lands in
the voxel grid when a Minecraft world is imported (mcimport.py) and which texture layers the material table points at."""
import numpy as np


def parse_blockdb(path):
    """Restates Core/BlockDatabaseParser.cpp:44-375 for the fields the path needs.  Block ids follow the order of the file, starting at 1
    (GenerateBlockID, :28-39).  Returns a list of dicts: Name, ID, faces{Albedo,Normal,PBR}{front..bottom}, Emissive, Transparent, SSS,
    MC_IDs (ParseIDsFromCommaSeparatedString, :8-27)."""
    blocks = []
    cur = None
    for raw in open(path, encoding="utf-8", errors="replace"):
        line = raw.strip()
        if line == "{":
            cur = {"faces": {k: {} for k in ("Albedo", "Normal", "PBR")}, "Emissive": "", "Transparent": False, "SSS": False, "MC_IDs": []}
            continue
        if line == "}":
            if cur is not None and "Name" in cur:
                blocks.append(cur)
            cur = None
            continue
        if cur is None or not line:
            continue
        key, _, val = line.partition(":")
        key, val = key.strip().rstrip(";"), val.strip()
        if key == "Name":
            cur["Name"] = val
        elif key.split("_")[0] in ("Albedo", "Normal", "PBR"):
            kind, _, face = key.partition("_")
            face = face or "default"
            cur["faces"][kind][face] = val
        elif key.startswith("Transparent"):
            cur["Transparent"] = True
        elif key.upper().startswith("SSS") or key.startswith("SUBSURFACE"):
            cur["SSS"] = True
        elif key == "Emissive":
            cur["Emissive"] = val
        elif key.upper().replace("_", "") == "MCID":
            cur["MC_IDs"] = [int(v) for v in val.split(",") if v.strip()]
    for i, b in enumerate(blocks):
        b["ID"] = i + 1
        for kind in ("Albedo", "Normal", "PBR"):
            f = b["faces"][kind]
            d = f.get("default", "")
            for face in ("front", "back", "left", "right", "top", "bottom"):
                f.setdefault(face, d)
    return blocks


def minecraft_id_lut(blocks):
    """uint8[256]: Minecraft block id -> engine block id (BlockDatabase::GetIDFromMCID, BlockDatabase.cpp:599-612): 0 -> 0, an id no block
    claims -> the id of INVALID_BLOCK.  Several blocks of the shipped blockdb.txt claim the same Minecraft id (35, 129, 249, 250); the
    reference fills its table while iterating an unordered_map (BlockDatabase.cpp:96-104), so the winner depends on the standard library's
    hash order.  Pinned here: the block that comes LATER in the file wins."""
    invalid = next((b["ID"] for b in blocks if b["Name"] == "INVALID_BLOCK"), 0)
    lut = np.full(256, invalid, dtype=np.uint8)
    lut[0] = 0
    for b in blocks:
        for mc in b["MC_IDs"]:
            lut[mc & 0xFF] = b["ID"]
    return lut
