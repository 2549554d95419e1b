"""N > 1 host logic on CPU: slab partition and the in-place slab gather, world_size 2 and 3 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from voxelpathtracer_b200 import multigpu


def test_slabs_partition_the_frame():
    for h in (1, 7, 360, 1080, 2160, 1081):
        for ws in (1, 2, 3, 4, 8):
            slabs = multigpu.all_slabs(h, ws)
            assert slabs[0][0] == 0 and slabs[-1][1] == h
            rows = [j for b, e in slabs for j in range(b, e)]
            assert rows == list(range(h))
            assert max(e - b for b, e in slabs) - min(e - b for b, e in slabs) <= -(-h // ws)
    assert multigpu.slab_rows(1080, 8, 3) == (405, 540)
    with pytest.raises(ValueError):
        multigpu.slab_rows(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, height, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        W = 16
        b, e = multigpu.slab_rows(height, ws, rank)
        # every rank fills only its own rows of the full-frame planes, like a slab trace does
        expect_t = torch.arange(height * W, dtype=torch.float32).reshape(height, W)
        expect_n = (torch.arange(height * W) % 7).to(torch.uint8).reshape(height, W)
        expect_sh = torch.arange(height * W * 4, dtype=torch.float32).reshape(height, W, 4)
        t = torch.full((height, W), -7.0)
        n = torch.full((height, W), 255, dtype=torch.uint8)
        sh = torch.full((height, W, 4), -7.0)
        t[b:e], n[b:e], sh[b:e] = expect_t[b:e], expect_n[b:e], expect_sh[b:e]
        multigpu.gather_planes({"t": t, "n": n, "sh": sh}, height)
        ok = bool(torch.equal(t, expect_t) and torch.equal(n, expect_n) and torch.equal(sh, expect_sh))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("ws,height", [(2, 36), (2, 37), (3, 30), (3, 31)])
def test_slab_gather_over_gloo(ws, height):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, height, q)) for r in range(ws)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(ws)]


def test_packed_layout_regions_do_not_overlap():
    table, total = multigpu.packed_layout(1920, 135)
    spans = sorted((off, off + 135 * 1920 * elem) for (name, elem, _, _), off in zip(multigpu.PLANES, [table[p[0]] for p in multigpu.PLANES]))
    assert spans[0][0] == 0 and spans[-1][1] <= total
    for (a0, a1), (b0, b1) in zip(spans[:-1], spans[1:]):
        assert a1 <= b0 and b0 % 256 == 0
    assert sum(p[1] for p in multigpu.PLANES) == 51  # bytes per pixel written per frame (fp32 planes)
    assert sum(p[1] for p in multigpu.plane_table(True)) == 27  # ... in the reference's FBO texel formats
    assert sum(p[1] for p in multigpu.plane_table(True) if not p[0].startswith('g_')) == 19  # ... of which cross the link
    assert [p[0] for p in multigpu.plane_table(True)] == [p[0] for p in multigpu.PLANES]
    # virtual plane bases stay inside the packed buffer for every rank: region_bytes >= any single plane's slab
    assert all(total >= 135 * 1920 * p[1] for p in multigpu.PLANES)
