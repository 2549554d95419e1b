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


def _edit_worker(rank, ws, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from voxelpathtracer_b200 import world

        class Recorder:            # stands in for the Renderer: the ABI calls a rank would make
            def __init__(self):
                self.calls = []

            def set_blocks(self, xyz, ids):
                self.calls.append(("set_blocks", xyz.copy(), ids.copy()))

            def build_distance_field(self):
                self.calls.append(("build",))

        w = world.World()
        r = Recorder()
        script = [[(10, 20, 30, 3), (383, 127, 383, 12), (0, 0, 0, 255)], [], [(10, 20, 30, 0)]]
        for edits in script:
            xyz, ids = multigpu.broadcast_edits(edits if rank == 0 else None)
            multigpu.apply_edits(r, w, xyz, ids)
        ok = w.get_block(10, 20, 30) == 0 and w.get_block(383, 127, 383) == 12 and w.get_block(0, 0, 0) == 255 and int((w.data != 0).sum()) == 2
        ok = ok and [c[0] for c in r.calls] == ["set_blocks", "build", "set_blocks", "build"] and r.calls[0][1].dtype == np.int16
        ok = ok and r.calls[0][1].tolist() == [[10, 20, 30], [383, 127, 383], [0, 0, 0]] and r.calls[0][2].tolist() == [3, 12, 255]
        bad = False
        if rank == 0:
            try:
                multigpu.broadcast_edits([(384, 0, 0, 1)])
            except ValueError:
                bad = True
        q.put((rank, bool(ok and (bad or rank != 0))))
    finally:
        dist.destroy_process_group()


def test_edit_list_broadcast_over_gloo():
    """Config 5 on N ranks: the edit list travels (7 bytes per edit), every rank edits its replica and rebuilds its own distance field."""
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_edit_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(results) == [(r, True) for r in range(ws)]


def test_stencil_rows_cover_the_slab_plus_halo():
    for h, ws, halo in [(1080, 8, 28), (360, 3, 17), (90, 2, 100)]:
        for r in range(ws):
            b, e = multigpu.slab_rows(h, ws, r)
            sb, se = multigpu.stencil_rows(h, ws, r, halo)
            assert sb == max(b - halo, 0) and se == min(e + halo, h) and sb <= b and se >= e
