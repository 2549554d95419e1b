"""World ingestion next to the hot path: blockdb.txt -> block ids / Minecraft-id table, Anvil region files -> voxel grid.
The importer is checked against the REFERENCE'S OWN importer — Core/NBT/Importer.cpp over its vendored enkiMI + miniz, compiled by
oracle/Makefile into oracle/_ref/ref_mc_import — on synthetic region files written by this repo and, when the reference tree is present,
on the Minecraft regions it ships; self-contained round-trip tests cover the semantics without it."""
import os
import subprocess

import numpy as np
import pytest

from voxelpathtracer_b200 import assets, blockdb, mcimport, world

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_IMPORTER = os.path.join(ROOT, "oracle", "_ref", "ref_mc_import")
REF_TREE = "/root/reference"
needs_ref_importer = pytest.mark.skipif(not os.path.exists(REF_IMPORTER), reason="oracle/_ref/ref_mc_import not built (needs /root/reference at build time)")


def reference_import(directory, origin, lut, tmp_path):
    lut_path, out_path = os.path.join(tmp_path, "lut.u8"), os.path.join(tmp_path, "ref_world.u8")
    np.asarray(lut, dtype=np.uint8).tofile(lut_path)
    subprocess.run([REF_IMPORTER, str(directory), *[str(v) for v in origin], lut_path, out_path], check=True, capture_output=True)
    return np.fromfile(out_path, dtype=np.uint8)


def synthetic_regions(directory, seed=3):
    """Two region files (r.0.0 and r.-1.-1) with a few chunks of random ids and data nibbles; returns {(x, y, z): (id, data)} of every voxel."""
    rng = np.random.RandomState(seed)
    truth = {}
    for (rx, rz), chunk_list in {(0, 0): [(0, 0), (1, 0), (5, 7), (31, 31)], (-1, -1): [(-1, -1), (-3, -2), (-32, -32)]}.items():
        chunks = {}
        for cx, cz in chunk_list:
            sections = {}
            for sy in rng.choice(8, size=2, replace=False):
                ids = (rng.rand(16, 16, 16) < 0.3) * rng.randint(1, 60, size=(16, 16, 16))
                dv = (rng.rand(16, 16, 16) < 0.2) * rng.randint(1, 16, size=(16, 16, 16))
                sections[int(sy)] = (ids.astype(np.uint8), dv.astype(np.uint8))
                for y, z, x in zip(*np.nonzero(ids)):
                    truth[(cx * 16 + x, sy * 16 + y, cz * 16 + z)] = (int(ids[y, z, x]), int(dv[y, z, x]))
            chunks[(cx, cz)] = sections
        mcimport.write_region_file(os.path.join(directory, f"r.{rx}.{rz}.mca"), chunks)
    return truth


def test_import_semantics_on_synthetic_regions(tmp_path):
    """Importer.cpp:66-147: data nibble 0 only, Minecraft id through the table (unmapped -> INVALID_BLOCK), position = voxel - origin +
    (192, 0, 192), voxels outside the grid dropped."""
    truth = synthetic_regions(str(tmp_path))
    lut = assets.load_minecraft_id_lut()
    for origin in [(0, 0, 0), (100.7, 3.2, -40.9), (-300, 0, -300)]:
        w = mcimport.import_world(str(tmp_path), origin, lut)
        want = np.zeros_like(w.data).reshape(384, 128, 384)
        ox, oy, oz = (int(v) for v in origin)
        for (x, y, z), (mc, dv) in truth.items():
            px, py, pz = x - ox + 192, y - oy, z - oz + 192
            if dv == 0 and lut[mc] != 0 and 0 <= px < 384 and 0 <= py < 128 and 0 <= pz < 384:
                want[pz, py, px] = lut[mc]
        assert np.array_equal(w.zyx, want), origin
        if origin == (0, 0, 0):
            assert w.data.any()
    assert lut[0] == 0 and lut[1] == world.STONE and lut[2] == world.GRASS and lut[3] == world.DIRT   # Minecraft stone / grass / dirt


@needs_ref_importer
def test_synthetic_regions_match_the_reference_importer(tmp_path):
    d = tmp_path / "regions"
    d.mkdir()
    synthetic_regions(str(d), seed=11)
    lut = assets.load_minecraft_id_lut()
    for origin in [(0, 0, 0), (37, 5, -60), (-100, 0, -100), (-400.5, 2.9, -420)]:
        ours = mcimport.import_world(str(d), origin, lut)
        ref = reference_import(d, origin, lut, str(tmp_path))
        assert np.array_equal(ours.data, ref), origin
        assert ref.any()


@needs_ref_importer
@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_TREE, "Test MC Worlds", "Medival")), reason="reference tree absent")
@pytest.mark.parametrize("origin", [(700, 0, -300), (-700, 0, -300), (150, 0, -700)])
def test_shipped_minecraft_regions_match_the_reference_importer(tmp_path, origin):
    """BASELINE config 5's loader on the regions the reference ships ('Test MC Worlds/Medival'), at origins inside the present files (the regions
    around the shipped Origin.txt are missing blobs): bit-equal to Core/NBT/Importer.cpp, and dense enough to be a real scene."""
    d = os.path.join(REF_TREE, "Test MC Worlds", "Medival")
    lut = assets.load_minecraft_id_lut()
    ours = mcimport.import_world(d, origin, lut)
    ref = reference_import(d, origin, lut, str(tmp_path))
    assert np.array_equal(ours.data, ref)
    assert (ref > 0).mean() > 0.01


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_TREE, "blockdb.txt")), reason="reference tree absent")
def test_block_database_parser_against_the_shipped_blockdb():
    blocks = blockdb.parse_blockdb(os.path.join(REF_TREE, "blockdb.txt"))
    assert [b["Name"] for b in blocks[:5]] == ["Grass", "Dirt", "Stone", "Cobblestone", "Sand"] and blocks[0]["ID"] == 1
    assert len(blocks) == 99 and all(b["ID"] == k + 1 for k, b in enumerate(blocks))
    names = np.load(os.path.join(ROOT, "voxelpathtracer_b200", "data", "materials.npz"))["block_names"]
    assert [str(n) for n in names[1:]] == [b["Name"] for b in blocks]            # the committed material fixture used the same ids
    lut = blockdb.minecraft_id_lut(blocks)
    assert np.array_equal(lut, assets.load_minecraft_id_lut())                    # committed fixture
    invalid = next(b["ID"] for b in blocks if b["Name"] == "INVALID_BLOCK")
    assert lut[0] == 0 and lut[200] == invalid and lut[2] == 1 and lut[50] == lut[51] == lut[76]   # torches share a block
