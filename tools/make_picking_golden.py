#!/usr/bin/env python3
"""tests/golden/picking_cases.json: what the REFERENCE'S OWN World::Raycast / RaycastDetect (Core/World.cpp:215-546, compiled from the
reference tree into oracle/_ref/libref_picking.so by oracle/Makefile) return on seeded rays in the city and superflat worlds.  Needs
/root/reference (this container); the committed JSON travels and pins voxelpathtracer_b200.world.World.raycast / raycast_detect."""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from voxelpathtracer_b200 import assets, world  # noqa: E402

LIB = os.path.join(ROOT, "oracle", "_ref", "libref_picking.so")


def load():
    lib = C.CDLL(LIB)
    lib.ref_world_raycast.restype = C.c_int
    lib.ref_world_raycast.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.ref_world_raycast_detect.restype = None
    lib.ref_world_raycast_detect.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def cases(seed=11, n=160):
    """(world name, op, position, direction, held block): eyes a little above ground / inside the city, looking mostly down and sideways."""
    rng = np.random.RandomState(seed)
    out = []
    for k in range(n):
        wname = "city" if k % 2 else "superflat"
        pos = [float(np.float32(v)) for v in (rng.uniform(40, 340), rng.uniform(51, 70) if wname == "superflat" else rng.uniform(30, 100), rng.uniform(40, 340))]
        d = rng.normal(size=3)
        d[1] = -abs(d[1]) if k % 3 else d[1]
        d = d / np.linalg.norm(d)
        out.append((wname, k % 3, pos, [float(np.float32(v)) for v in d], int(rng.choice([world.STONE, world.LAMP, world.GRASS]))))
    return out


def reference_result(lib, blocks, emissive, op, pos, d, held):
    """The reference on a COPY of the grid: return value, held block afterwards, the voxels it changed, RaycastDetect's answer (hits only)."""
    grid = np.array(blocks, copy=True)
    p, dd = np.array(pos, np.float32), np.array(d, np.float32)
    held_out = C.c_int(0)
    ret = lib.ref_world_raycast(grid.ctypes.data, emissive.ctypes.data, op, p.ctypes.data, dd.ctypes.data, held, C.byref(held_out))
    changed = np.nonzero(grid != blocks)[0]
    edits = [[int(i % 384), int((i // 384) % 128), int(i // (384 * 128)), int(grid[i])] for i in changed]
    return {"ret": int(ret), "held": int(held_out.value), "edits": edits}


def main():
    lib = load()
    mats = assets.load_materials()
    emissive = np.ascontiguousarray(mats["table"][384:512], dtype=np.int32)
    worlds = {"superflat": world.generate_superflat(), "city": world.generate_city()}
    out = {"source": "Core/World.cpp:215-546 compiled as oracle/_ref/libref_picking.so", "cases": []}
    for wname, op, pos, d, held in cases():
        w = worlds[wname]
        r = reference_result(lib, w.data, emissive, op, pos, d, held)
        port_hit = world.World(w.data.copy()).raycast_detect(pos, d)
        if port_hit is not None:  # RaycastDetect is only defined where something is hit
            det = (C.c_int * 4)()
            p, dd = np.array(pos, np.float32), np.array(d, np.float32)
            grid = np.array(w.data, copy=True)
            lib.ref_world_raycast_detect(grid.ctypes.data, p.ctypes.data, dd.ctypes.data, det)
            r["detect"] = [int(v) for v in det]
        out["cases"].append({"world": wname, "op": op, "pos": pos, "dir": d, "held": held, **r})
    with open(os.path.join(ROOT, "tests", "golden", "picking_cases.json"), "w") as f:
        json.dump(out, f)
    print(len(out["cases"]), "cases;", sum(1 for c in out["cases"] if c["edits"]), "with edits;", sum(1 for c in out["cases"] if "detect" in c), "with RaycastDetect hits")


if __name__ == "__main__":
    main()
