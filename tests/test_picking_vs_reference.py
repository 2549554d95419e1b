"""CPU picking pinned to the reference: World::Raycast (break / place / pick) and World::RaycastDetect of Core/World.cpp:215-546, compiled from
the reference tree behind stand-ins for everything but the block grid (oracle/ref_picking_driver.cpp -> oracle/_ref/libref_picking.so),
against voxelpathtracer_b200.world.World.raycast / raycast_detect on seeded rays in the superflat and city worlds.
Committed golden vectors (tests/golden/picking_cases.json, tools/make_picking_golden.py) always; the library itself when it is present."""
import json
import os
import sys

import numpy as np
import pytest

from voxelpathtracer_b200 import world

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _port_result(w, op, pos, d, held):
    """The port on a copy of the grid, in the reference's terms: return value, held block afterwards, voxels changed."""
    cp = world.World(w.data.copy())
    det = cp.raycast_detect(pos, d)
    r = cp.raycast(op, pos, d, held_block=held)
    changed = np.nonzero(cp.data != w.data)[0]
    edits = [[int(i % 384), int((i // 384) % 128), int(i // (384 * 128)), int(cp.data[i])] for i in changed]
    held_after = held
    if op == 2 and r["voxel"] is not None and r["block"] > 0:
        held_after = r["block"]         # m_CurrentlyHeldBlock = block (:470-475)
    return {"ret": int(bool(r["changed"])), "held": int(held_after), "edits": edits, "detect": None if det is None else [int(v) for v in det]}


def _check(case, got):
    tag = (case["world"], case["op"], case["pos"], case["dir"])
    assert got["ret"] == case["ret"], tag
    assert got["edits"] == case["edits"], tag
    assert got["held"] == case["held"], tag
    if "detect" in case:
        assert got["detect"] == case["detect"], tag


def test_picking_reproduces_the_committed_reference_vectors(worlds):
    with open(os.path.join(ROOT, "tests", "golden", "picking_cases.json")) as f:
        cases = json.load(f)["cases"]
    assert len(cases) >= 100 and sum(1 for c in cases if c["edits"]) >= 30
    for c in cases:
        _check(c, _port_result(worlds[c["world"]], c["op"], c["pos"], c["dir"], c["held"]))


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libref_picking.so")), reason="reference picking library not built")
def test_picking_live_against_the_compiled_reference(worlds, scene_tables):
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import make_picking_golden as mpg
    lib = mpg.load()
    emissive = np.ascontiguousarray(scene_tables["materials"]["table"][384:512], dtype=np.int32)
    for wname, op, pos, d, held in mpg.cases(seed=29, n=90):      # other rays than the committed ones
        w = worlds[wname]
        ref = mpg.reference_result(lib, w.data, emissive, op, pos, d, held)
        _check({"world": wname, "op": op, "pos": pos, "dir": d, **ref}, _port_result(w, op, pos, d, held))
