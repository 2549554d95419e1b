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


def packed_layout(width, rows, planes=PLANES):
    """Byte offsets of every plane's slab inside one packed region (256-byte aligned) and the region size."""
    off, table = 0, {}
    for name, elem, _, _ in planes:
        table[name] = off
        off += (rows * width * elem + 255) & ~255
    return table, off


def pick_band_rows(rows_per_rank, limit=8):
    """Largest band height <= limit that divides a rank's row count (bands are interleaved across ranks)."""
    return max(b for b in range(1, limit + 1) if rows_per_rank % b == 0)


def pick_chunks(rows_per_rank, world_size):
    """Sub-slabs per rank and frame (exchange of one overlaps tracing of the next).  Frames are double-buffered, so the
    exchange of frame k already overlaps the tracing of frame k+1; extra chunks only shrink the kernels (worse wave
    quantisation on 148 SMs, more launches), hence 1 by default."""
    return 1


def image_rows_of_rank(height, world_size, rank, band_rows):
    """Image row of every virtual row of `rank` (VxCamera interleave contract)."""
    rows = height // world_size
    v = np.arange(rows)
    if world_size == 1:
        return v
    return ((v // band_rows) * world_size + rank) * band_rows + v % band_rows


class ShardedFrame:
    """Per-rank driver of one frame on N GPUs.

    * Load balance: rows are dealt to the ranks in interleaved bands (`band_rows` image rows each), because sky rows cost a
      few loop iterations and horizon rows a hundred; a rank addresses its rows densely as virtual rows (VxCamera).
    * Overlap: a rank's rows are cut into `chunks` sub-slabs; the exchange of chunk c (NCCL, own stream) runs while chunk
      c+1 is traced (the library's stream), ordered with stream events only.
    * One collective per chunk: every exchanged plane of a (rank, chunk) lives in one contiguous region of the packed
      buffer `[chunks][world_size][region]`; kernels get *virtual* plane bases so the ABI's `v * width + i` indexing lands
      in the region, and the exchange is an in-place `all_gather_into_tensor`.
    * What is exchanged: the "radiance slabs" — the outputs of the secondary passes (`s_*`, `d_*`).  The G-buffer stays
      sharded unless `exchange_gbuffer=True` (its only consumers inside the path are the secondary passes of the same rows).
    """

    def __init__(self, renderer, fps_camera, width, height, group=None, chunks=None, band_rows=None, exchange_gbuffer=False, slots=2):
        self.r = renderer
        self.fc = fps_camera
        self.width, self.height = width, height
        self.group = group
        inited = dist is not None and dist.is_initialized()
        self.rank = dist.get_rank(group) if inited else 0
        self.world_size = N = dist.get_world_size(group) if inited else 1
        if height % N:
            raise ValueError("row sharding needs height % world_size == 0")
        self.rows = height // N
        self.band_rows = (band_rows or pick_band_rows(self.rows)) if N > 1 else 0
        self.chunks = chunks or pick_chunks(self.rows, N)
        if self.rows % self.chunks:
            raise ValueError("chunks must divide the rows of a rank")
        self.rpc = self.rows // self.chunks
        self.slots = slots if N > 1 else 1
        dev = f"cuda:{renderer.device}"
        self.exchanged = [p for p in PLANES if exchange_gbuffer or not p[0].startswith("g_")]
        self.local = [p for p in PLANES if p not in self.exchanged]
        self.offsets, self.region_bytes = packed_layout(width, self.rpc, self.exchanged)
        self.buf = torch.zeros((self.slots, self.chunks, N, self.region_bytes), dtype=torch.uint8, device=dev)
        self.loc_offsets, loc_bytes = packed_layout(width, self.rows, self.local)
        self.loc = torch.zeros((max(loc_bytes, 256),), dtype=torch.uint8, device=dev)
        self.cams, self.bases = [], []
        for c in range(self.chunks):
            vb, ve = c * self.rpc, (c + 1) * self.rpc
            self.cams.append(fps_camera.vx_camera(width, height, vb, ve, N if N > 1 else 0, self.rank, self.band_rows))
        for slot in range(self.slots):
            per_chunk = []
            for c in range(self.chunks):
                vb = c * self.rpc
                region = self.buf[slot, c, self.rank].data_ptr()
                base = {n: region + self.offsets[n] - vb * width * e for n, e, _, _ in self.exchanged}  # virtual plane bases
                base.update({n: self.loc.data_ptr() + self.loc_offsets[n] for n, e, _, _ in self.local})
                per_chunk.append(base)
            self.bases.append(per_chunk)
        self._ext = torch.cuda.ExternalStream(renderer.cuda_stream(), device=dev)
        self._comm = torch.cuda.Stream(device=dev) if N > 1 else None
        self._traced = [[torch.cuda.Event() for _ in range(self.chunks)] for _ in range(self.slots)]
        self._exchanged = [None] * self.slots  # event of the last exchange that read/wrote a slot
        self._slot = 0
        self.last_slot = 0

    def _planes(self, c, slot=0):
        b = self.bases[slot][c]
        return ({"t": b["g_t"], "normal_id": b["g_normal_id"], "block_id": b["g_block_id"], "inv_t": b["g_inv_t"]},
                {"shadow": b["s_shadow"], "transversal": b["s_transversal"]},
                {"sh": b["d_sh"], "cocg": b["d_cocg"], "luma": b["d_luma"], "ao_sky": b["d_ao_sky"]})

    def next_slot(self):
        slot = self._slot
        self._slot = (slot + 1) % self.slots
        self.last_slot = slot
        return slot

    def trace_into(self, slot, primary, shadow, diffuse):
        """Only the library calls of one frame (this rank's rows -> frame slot `slot`); capturable into a CUDA graph."""
        for c in range(self.chunks):
            g, s, d = self._planes(c, slot)
            self.r.trace_primary(self.cams[c], primary, g)
            if shadow is not None:
                self.r.trace_shadow(self.cams[c], g, shadow, s)
            if diffuse is not None:
                self.r.trace_diffuse(self.cams[c], g, diffuse, d)

    def before_trace(self, slot):
        """The slot's previous exchange must have drained before it is overwritten (stream-ordered, no host sync)."""
        if self._exchanged[slot] is not None:
            self._ext.wait_event(self._exchanged[slot])

    def exchange(self, slot):
        """Start the exchange of a traced slot on the communication stream: one in-place NCCL all-gather per chunk."""
        if self.world_size == 1:
            return
        self._traced[slot][0].record(self._ext)
        with torch.cuda.stream(self._comm):
            self._comm.wait_event(self._traced[slot][0])
            for c in range(self.chunks):
                dist.all_gather_into_tensor(self.buf[slot, c].view(-1), self.buf[slot, c, self.rank], group=self.group)
        ev = torch.cuda.Event()
        ev.record(self._comm)
        self._exchanged[slot] = ev

    def render(self, primary, shadow, diffuse):
        """Trace this rank's rows of one frame into the next frame slot and start its exchange.  Everything is enqueued
        asynchronously: the exchange of this frame (NCCL, own stream) overlaps the tracing of the next frame (library
        stream, other slot).  Call finish() before reading planes or stopping a timer."""
        slot = self.next_slot()
        self.before_trace(slot)
        self.trace_into(slot, primary, shadow, diffuse)
        self.exchange(slot)

    def finish(self):
        """Make the library's stream wait for every outstanding exchange (no host sync)."""
        if self.world_size > 1:
            self._ext.wait_stream(self._comm)

    # first interface
    def trace(self, primary, shadow, diffuse):
        self.render(primary, shadow, diffuse)

    def gather(self):
        self.finish()

    def plane(self, name, slot=None):
        """Full-frame tensor [H, W, ...] of an exchanged plane (or of this rank's rows for a sharded-only plane) of the last
        rendered frame, in IMAGE row order — a copy for consumers and tests."""
        slot = self.last_slot if slot is None else slot
        spec = next(p for p in PLANES if p[0] == name)
        _, elem, dtype, tail = spec
        tdt = getattr(torch, dtype)
        if spec in self.local:
            n = self.rows * self.width * elem
            local = self.loc[self.loc_offsets[name]:self.loc_offsets[name] + n].view(tdt).view((self.rows, self.width) + tail)
            return local
        n = self.rpc * self.width * elem
        out = torch.empty((self.height, self.width) + tail, dtype=tdt, device=self.buf.device)
        for r in range(self.world_size):
            rows = torch.from_numpy(image_rows_of_rank(self.height, self.world_size, r, self.band_rows)).to(self.buf.device)
            part = torch.cat([self.buf[slot, c, r, self.offsets[name]:self.offsets[name] + n].view(tdt).view((self.rpc, self.width) + tail)
                              for c in range(self.chunks)], 0)
            out[rows] = part
        return out
