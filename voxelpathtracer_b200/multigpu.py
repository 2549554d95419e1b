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


class ShardedFrame:
    """Per-rank driver: traces this rank's slab of the primary, shadow and diffuse passes into full-frame device planes
    and gathers them.  `renderer` is a voxelpathtracer_b200.Renderer bound to this rank's GPU."""

    def __init__(self, renderer, fps_camera, width, height, group=None):
        self.r = renderer
        self.width, self.height = width, height
        self.group = group
        self.rank = dist.get_rank(group) if dist is not None and dist.is_initialized() else 0
        self.world_size = dist.get_world_size(group) if dist is not None and dist.is_initialized() else 1
        rb, re_ = slab_rows(height, self.world_size, self.rank)
        self.cam = fps_camera.vx_camera(width, height, rb, re_)
        self.gbuf = renderer.alloc_gbuffer(width, height, device=True)
        self.shadow = renderer.alloc_shadow(width, height, device=True)
        self.diffuse = renderer.alloc_diffuse(width, height, device=True)
        self._ext_stream = torch.cuda.ExternalStream(renderer.cuda_stream(), device=f"cuda:{renderer.device}")

    def trace(self, primary, shadow, diffuse):
        """Enqueue the three passes for this rank's slab (asynchronous on the renderer's stream)."""
        self.r.trace_primary(self.cam, primary, self.gbuf)
        if shadow is not None:
            self.r.trace_shadow(self.cam, self.gbuf, shadow, self.shadow)
        if diffuse is not None:
            self.r.trace_diffuse(self.cam, self.gbuf, diffuse, self.diffuse)

    def gather(self, with_gbuffer=True):
        """Exchange the finished slabs; afterwards every rank holds the full frame."""
        if self.world_size == 1:
            return
        torch.cuda.current_stream().wait_stream(self._ext_stream)  # NCCL runs after the trace kernels, no host sync
        planes = {}
        if with_gbuffer:
            planes.update({"g_" + k: v for k, v in self.gbuf.items()})
        planes.update({"s_" + k: v for k, v in self.shadow.items()})
        planes.update({"d_" + k: v for k, v in self.diffuse.items()})
        gather_planes(planes, self.height, self.group)
        self._ext_stream.wait_stream(torch.cuda.current_stream())
