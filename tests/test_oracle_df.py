"""Pins for the distance-field oracle (the reference ships none — SURVEY.md §4): analytic known answers,
the brute-force definition on small grids, sampled definition checks at full size, committed digests."""
import hashlib

import numpy as np
import pytest

from oracle import vxo
from voxelpathtracer_b200 import abi


def test_empty_world_is_all_254():
    df = vxo.df_build(np.zeros(8 * 8 * 8, np.uint8), (8, 8, 8))
    assert df.min() == 24 and df.max() == 24  # min(254, wx+wy+wz) on an 8^3 grid (ManhattanDistanceX.comp:49)
    df = vxo.df_build(np.zeros(abi.WORLD_VOXELS, np.uint8))
    assert df.min() == 254 and df.max() == 254


def test_single_voxel_is_l1_distance():
    dims = (40, 24, 36)
    g = np.zeros(dims[2] * dims[1] * dims[0], np.uint8)
    sx, sy, sz = 7, 20, 30
    g[sx + dims[0] * (sy + dims[1] * sz)] = 9
    df = vxo.df_build(g, dims).reshape(dims[2], dims[1], dims[0])
    z, y, x = np.meshgrid(np.arange(dims[2]), np.arange(dims[1]), np.arange(dims[0]), indexing="ij")
    expect = np.minimum(np.abs(x - sx) + np.abs(y - sy) + np.abs(z - sz), min(254, sum(dims)))
    assert np.array_equal(df, expect.astype(np.uint8))


def test_clamp_at_254_full_size():
    g = np.zeros(abi.WORLD_VOXELS, np.uint8)
    g[0] = 1  # solid voxel at the origin: far corner is 383+127+383 away
    df = vxo.df_build(g).reshape(384, 128, 384)
    assert df[0, 0, 0] == 0 and df[0, 0, 100] == 100 and df[0, 0, 253] == 253 and df[0, 0, 254] == 254 and df[0, 0, 300] == 254
    assert df[100, 100, 53] == 253 and df[100, 100, 54] == 254 and df[383, 127, 383] == 254


@pytest.mark.parametrize("seed,fill", [(0, 0.002), (1, 0.02), (2, 0.3), (3, 0.9)])
def test_matches_bruteforce_definition_on_small_grids(seed, fill):
    rng = np.random.RandomState(seed)
    dims = (int(rng.randint(5, 40)), int(rng.randint(3, 20)), int(rng.randint(5, 40)))
    g = (rng.rand(dims[0] * dims[1] * dims[2]) < fill).astype(np.uint8) * rng.randint(1, 128, dims[0] * dims[1] * dims[2]).astype(np.uint8)
    assert np.array_equal(vxo.df_build(g, dims), vxo.df_bruteforce(g, dims))


def test_transparent_blocks_are_solid_for_the_df():
    # "solid" is block byte > 0 (ManhattanDistanceX.comp:40-43): leaves (id 7, Transparent) seed distance 0
    g = np.zeros(16 * 16 * 16, np.uint8)
    g[5 + 16 * (5 + 16 * 5)] = 7
    df = vxo.df_build(g, (16, 16, 16))
    assert df[5 + 16 * (5 + 16 * 5)] == 0 and df[6 + 16 * (5 + 16 * 5)] == 1


def test_sampled_definition_on_plains(worlds, oracle_dfs):
    """At random voxels the oracle value equals min(254, L1 distance to the nearest solid voxel), found by an
    expanding-shell search that is independent of the sweep algorithm."""
    w = worlds["plains"].zyx
    df = oracle_dfs["plains"].reshape(384, 128, 384)
    rng = np.random.RandomState(11)
    for _ in range(300):
        x, y, z = int(rng.randint(384)), int(rng.randint(128)), int(rng.randint(384))
        m = int(df[z, y, x])
        r = m
        x0, x1, y0, y1, z0, z1 = max(x - r, 0), min(x + r + 1, 384), max(y - r, 0), min(y + r + 1, 128), max(z - r, 0), min(z + r + 1, 384)
        sub = w[z0:z1, y0:y1, x0:x1] > 0
        zz, yy, xx = np.nonzero(sub)
        d = np.abs(xx + x0 - x) + np.abs(yy + y0 - y) + np.abs(zz + z0 - z)
        assert d.size > 0 and d.min() == m


def test_committed_digests(worlds, oracle_dfs, golden_digests):
    for name in ("superflat", "plains", "city", "gi_box", "sparse"):
        assert hashlib.sha256(worlds[name].data.tobytes()).hexdigest() == golden_digests["world"][name], name
        assert hashlib.sha256(oracle_dfs[name].tobytes()).hexdigest() == golden_digests["df"][name], name


def test_plains_statistics_match_the_survey_probe(worlds, oracle_dfs):
    # heights 43..58, max M = 85, ~41-42 % of voxels with M <= 3
    h = (worlds["plains"].zyx > 0).sum(1)
    assert h.min() == 43 and h.max() == 58
    df = oracle_dfs["plains"]
    assert df.max() == 85 and 0.41 < (df <= 3).mean() < 0.43
    assert oracle_dfs["superflat"].max() == 78
