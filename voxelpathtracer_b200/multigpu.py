"""Row-slab sharding of a frame across GPUs: one process per GPU, the voxel grid + distance field
replicated (37.7 MB), every rank traces rows [row_begin, row_end) of every plane, and the finished slabs are
exchanged with ONE collective per plane group (NCCL all-gather over NVLink; gloo on CPU for tests).

Secondary passes read the G-buffer at their own pixel only, so no halo is exchanged.  There is no data-path
collective inside a pass; the only exchange step is the gather of the output slabs.
"""
import numpy as np

try:
    import torch
    import torch.distributed as dist
except Exception:  # pragma: no cover
    torch = None
    dist = None


def slab_rows(height, world_size, rank):
    """Contiguous slab of rank `rank`: ceil-sized slabs, the last ranks may get fewer (or zero) rows."""
    if not (0 <= rank < world_size):
        raise ValueError("rank outside the world")
    per = -(-height // world_size)
    b = min(rank * per, height)
    e = min(b + per, height)
    return b, e


def all_slabs(height, world_size):
    return [slab_rows(height, world_size, r) for r in range(world_size)]


def gather_planes(planes, height, group=None):
    """All-gather row slabs in place.  `planes`: dict name -> tensor [H, W, ...] (full-frame sized on every rank, only
    this rank's rows valid on entry, all rows valid on return).  Equal slabs use all_gather_into_tensor directly on
    the plane (the send buffer is the plane's own slab = in-place all-gather); ragged slabs fall back to one
    broadcast per rank."""
    ws = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if ws == 1:
        return planes
    slabs = all_slabs(height, ws)
    equal = height % ws == 0
    works = []
    for name, full in planes.items():
        b, e = slabs[rank]
        if equal:
            flat = full.view(-1)
            per = flat.numel() // ws
            works.append(dist.all_gather_into_tensor(flat, flat[rank * per:(rank + 1) * per], group=group, async_op=True))
        else:
            for r, (rb, re_) in enumerate(slabs):
                if re_ > rb:
                    works.append(dist.broadcast(full[rb:re_], src=dist.get_global_rank(group, r) if group is not None else r, group=group, async_op=True))
    for w in works:
        w.wait()
    return planes


PLANES = (  # name, bytes per pixel, torch dtype name, trailing shape
    ("g_t", 4, "float32", ()), ("g_normal_id", 1, "uint8", ()), ("g_block_id", 1, "uint8", ()), ("g_inv_t", 4, "float32", ()),
    ("s_shadow", 1, "uint8", ()), ("s_transversal", 4, "float32", ()),
    ("d_sh", 16, "float32", (4,)), ("d_cocg", 8, "float32", (2,)), ("d_luma", 4, "float32", ()), ("d_ao_sky", 8, "float32", (2,)),
)


def packed_layout(width, rows_per_rank, planes=PLANES):
    """Byte offsets of every plane's slab inside one rank's region of the packed exchange buffer (256-byte aligned)."""
    off, table = 0, {}
    for name, elem, _, _ in planes:
        table[name] = off
        off += (rows_per_rank * width * elem + 255) & ~255
    return table, off


class ShardedFrame:
    """Per-rank driver of one frame: traces this rank's row slab of the primary, shadow and diffuse passes and exchanges the
    finished slabs with ONE collective.

    Every plane's slab of a rank lives in one contiguous region of a packed device buffer `[world_size][region_bytes]`; the
    kernels get *virtual* plane bases (region start - slab offset) so that the C ABI's `row * width + i` indexing lands
    inside the region, and the exchange is a single in-place `all_gather_into_tensor` over NVLink.  `plane(name)` returns
    the gathered full-frame view `[H, W, ...]` assembled from the per-rank regions (rows of rank r are region r)."""

    def __init__(self, renderer, fps_camera, width, height, group=None):
        self.r = renderer
        self.width, self.height = width, height
        self.group = group
        inited = dist is not None and dist.is_initialized()
        self.rank = dist.get_rank(group) if inited else 0
        self.world_size = dist.get_world_size(group) if inited else 1
        if height % self.world_size:
            raise ValueError("packed slab exchange needs height % world_size == 0")
        self.rows = height // self.world_size
        rb, re_ = self.rank * self.rows, (self.rank + 1) * self.rows
        self.cam = fps_camera.vx_camera(width, height, rb, re_)
        self.offsets, self.region_bytes = packed_layout(width, self.rows)
        dev = f"cuda:{renderer.device}"
        self.buf = torch.zeros((self.world_size, self.region_bytes), dtype=torch.uint8, device=dev)
        base = self.buf.data_ptr() + self.rank * self.region_bytes
        elem = {n: e for n, e, _, _ in PLANES}
        vb = {n: base + self.offsets[n] - rb * width * elem[n] for n in elem}  # virtual bases (never dereferenced outside the slab)
        self.gbuf = {"t": vb["g_t"], "normal_id": vb["g_normal_id"], "block_id": vb["g_block_id"], "inv_t": vb["g_inv_t"]}
        self.shadow = {"shadow": vb["s_shadow"], "transversal": vb["s_transversal"]}
        self.diffuse = {"sh": vb["d_sh"], "cocg": vb["d_cocg"], "luma": vb["d_luma"], "ao_sky": vb["d_ao_sky"]}
        self._ext_stream = torch.cuda.ExternalStream(renderer.cuda_stream(), device=dev)

    def trace(self, primary, shadow, diffuse):
        """Enqueue the three passes for this rank's slab (asynchronous on the renderer's stream)."""
        self.r.trace_primary(self.cam, primary, self.gbuf)
        if shadow is not None:
            self.r.trace_shadow(self.cam, self.gbuf, shadow, self.shadow)
        if diffuse is not None:
            self.r.trace_diffuse(self.cam, self.gbuf, diffuse, self.diffuse)

    def gather(self):
        """Exchange the finished slabs (one NCCL all-gather); afterwards every rank holds every region."""
        if self.world_size == 1:
            return
        torch.cuda.current_stream().wait_stream(self._ext_stream)  # NCCL runs after the trace kernels, no host sync
        dist.all_gather_into_tensor(self.buf.view(-1), self.buf[self.rank], group=self.group)
        self._ext_stream.wait_stream(torch.cuda.current_stream())

    def plane(self, name):
        """Full-frame tensor [H, W, ...] of a plane, assembled from the per-rank regions (a copy; for consumers and tests)."""
        _, elem, dtype, tail = next(p for p in PLANES if p[0] == name)
        n = self.rows * self.width * elem
        parts = [self.buf[r, self.offsets[name]:self.offsets[name] + n].view(getattr(torch, dtype)).view((self.rows, self.width) + tail)
                 for r in range(self.world_size)]
        return torch.cat(parts, 0)
