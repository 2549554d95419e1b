"""Row-slab sharding of a frame across GPUs: one process per GPU, the voxel grid + distance field
replicated (37.7 MB), every rank traces rows [row_begin, row_end) of every plane, and the finished slabs are
exchanged with ONE collective per plane group (NCCL all-gather over NVLink; gloo on CPU for tests).

Secondary passes read the G-buffer at their own pixel only, so no halo is exchanged.  There is no data-path
collective inside a pass; the only exchange step is the gather of the output slabs.
"""
import numpy as np

from . import abi

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


def broadcast_edits(edits, src=0, group=None, device=None):
    """Block edits on a replicated world (BASELINE config 5 on N GPUs): the rank that took the player's input holds the
    edit list [(x, y, z, block), ...]; it is broadcast as 7 bytes per edit (3 x int16 + 1 x uint8) and every rank applies it to its own
    copy and rebuilds its own distance field (tens of microseconds) — instead of broadcasting the 18.9 MB field.  Returns
    (xyz int16 [n, 3], ids uint8 [n]) on every rank, ready for Renderer.set_blocks; other ranks pass edits=None."""
    ws = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if rank == src:
        arr = np.asarray(edits, dtype=np.int64).reshape(-1, 4)
        if arr.size and (arr[:, :3].min() < 0 or (arr[:, :3] >= np.array([abi.WORLD_SIZE_X, abi.WORLD_SIZE_Y, abi.WORLD_SIZE_Z])).any()
                         or arr[:, 3].min() < 0 or arr[:, 3].max() > 255):
            raise ValueError("edit outside the world / block id outside 0..255")
        xyz, ids = arr[:, :3].astype(np.int16), arr[:, 3].astype(np.uint8)
    else:
        xyz, ids = np.zeros((0, 3), np.int16), np.zeros(0, np.uint8)
    if ws == 1:
        return xyz, ids
    n = torch.tensor([xyz.shape[0]], dtype=torch.int64, device=device)
    dist.broadcast(n, src=src, group=group)
    count = int(n.item())
    payload = torch.zeros(count * 7, dtype=torch.uint8, device=device)
    if rank == src and count:
        packed = np.concatenate([xyz.view(np.uint8).reshape(count, 6), ids.reshape(count, 1)], axis=1).reshape(-1)
        payload.copy_(torch.from_numpy(packed))
    if count:
        dist.broadcast(payload, src=src, group=group)
    raw = payload.cpu().numpy().reshape(count, 7)
    return np.ascontiguousarray(raw[:, :6]).view(np.int16).reshape(count, 3).copy(), raw[:, 6].copy()


def apply_edits(renderer, world, xyz, ids):
    """Every rank: the edits into the host grid, the device grid (vxpt_set_blocks) and a distance-field rebuild."""
    for (x, y, z), b in zip(xyz, ids):
        world.set_block(int(x), int(y), int(z), int(b))
    if len(ids):
        renderer.set_blocks(xyz, ids)
        renderer.build_distance_field()


def stencil_rows(height, world_size, rank, halo):
    """Rows [b - halo, e + halo) clipped to the frame: what a rank must hold of the input planes to run a stencil pass (the denoisers of
    SURVEY.md §8 f2) on its slab.  The a-trous chain reaches 16 * (1 + 2.4 * resolution_scale) + 1 rows per pass at its widest step."""
    b, e = slab_rows(height, world_size, rank)
    return max(b - halo, 0), min(e + halo, height)


PLANES = (  # name, bytes per pixel, torch dtype name, trailing shape — fp32 planes (VXPT_OPT_TEXEL_FORMAT = 0)
    ("g_t", 4, "float32", ()), ("g_normal_id", 1, "uint8", ()), ("g_block_id", 1, "uint8", ()), ("g_inv_t", 4, "float32", ()),
    ("s_shadow", 1, "uint8", ()), ("s_transversal", 4, "float32", ()),
    ("d_sh", 16, "float32", (4,)), ("d_cocg", 8, "float32", (2,)), ("d_luma", 4, "float32", ()), ("d_ao_sky", 8, "float32", (2,)),
)
PLANES_TEXEL = (  # the reference's FBO texel formats (VXPT_OPT_TEXEL_FORMAT = 1): 19 B/pixel of shadow + GI instead of 41
    ("g_t", 2, "float16", ()), ("g_normal_id", 1, "uint8", ()), ("g_block_id", 1, "uint8", ()), ("g_inv_t", 4, "float32", ()),
    ("s_shadow", 1, "uint8", ()), ("s_transversal", 2, "float16", ()),
    ("d_sh", 8, "float16", (4,)), ("d_cocg", 4, "float16", (2,)), ("d_luma", 2, "float16", ()), ("d_ao_sky", 2, "uint8", (2,)),
)


def plane_table(texel=False):
    return PLANES_TEXEL if texel else PLANES


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


FLAG_STRIDE_WORDS = 32   # one 128-byte line per flag word
FLAG_BYTES = 8192        # [0, 4096): arrived[rank]; [4096, ...): consumed


class P2PUnavailable(RuntimeError):
    """Raised on EVERY rank when some rank cannot map the gather root's buffer (fall back to exchange='nccl')."""


class _DevMem:
    """Raw device memory as a __cuda_array_interface__ object (so torch can view a vxpt_shared_alloc buffer)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


class ShardedFrame:
    """Per-rank driver of one frame on N GPUs.

    * Load balance: rows are dealt to the ranks in interleaved bands (`band_rows` image rows each), because sky rows cost a
      few loop iterations and horizon rows a hundred; a rank addresses its rows densely as virtual rows (VxCamera).
    * Packed slabs: every exchanged plane of a (rank, chunk) lives in one contiguous region of the packed buffer
      `[slots][chunks][world_size][region]`; kernels get *virtual* plane bases so the ABI's `v * width + i` indexing lands in the
      region.
    * What is exchanged: the "radiance slabs" — the outputs of the secondary passes (`s_*`, `d_*`).  The G-buffer stays
      sharded unless `exchange_gbuffer=True` (its only consumers inside the path are the secondary passes of the same rows).
    * exchange="nccl": one in-place `all_gather_into_tensor` per frame on its own stream (every rank ends up with every slab),
      double-buffered so it overlaps the tracing of the next frame.
    * exchange="p2p": the packed buffer lives on the gather root only (`vxpt_shared_alloc`), every other rank maps it
      (`vxpt_shared_open`, CUDA IPC) and hands addresses inside the mapping to its trace calls, so the trace kernels store their
      rows straight into the root's memory over NVLink while they trace — there is no exchange step and no copy.  Frames are
      ordered by flag words in the same buffer: a rank releases `arrived[rank] = k` after the passes of frame k
      (`vxpt_signal`), the root's gather stream waits for all N (`vxpt_wait_all`) and then publishes `consumed = k`, and a rank
      waits for `consumed >= k - slots` before it overwrites a slot.  Nothing blocks the host.
    * exchange="p2pcopy" (what bench.py uses): as p2p, but a rank traces into local memory and pushes its packed region to the root
      with ONE copy-engine transfer per frame (`vxpt_copy_async`, stream-ordered between the passes and the arrival flag): full-line
      NVLink writes instead of the kernels' scattered 2-8-byte stores.  Measured on 8 B200 (r01g, 1080p, 19 B/pixel): direct stores
      0.148 ms/frame (the root's NVLink ingress saturates on partial-sector writes), copy-engine push 0.071 ms/frame = 7.65x one GPU.
    * texel=True selects VXPT_OPT_TEXEL_FORMAT = 1 on the renderer: 19 instead of 41 bytes per pixel cross the link.
    """

    def __init__(self, renderer, fps_camera, width, height, group=None, chunks=None, band_rows=None, exchange_gbuffer=False, slots=2,
                 exchange="nccl", texel=False, root=0, timeout_ms=2000, emulate=None, max_gi_spp=1):
        self.r = renderer
        self.fc = fps_camera
        self.width, self.height = width, height
        self.group = group
        inited = dist is not None and dist.is_initialized()
        self.rank = dist.get_rank(group) if inited else 0
        self.world_size = N = dist.get_world_size(group) if inited else 1
        if emulate is not None:  # (world_size, rank): this process traces that rank's share of the frame and exchanges nothing
            N, self.rank = emulate
            self.world_size = N
        if height % N:
            raise ValueError("row sharding needs height % world_size == 0")
        if exchange not in ("nccl", "p2p", "p2pcopy"):
            raise ValueError("exchange must be 'nccl', 'p2p' or 'p2pcopy'")
        self.exchange_mode = "none" if (N == 1 or emulate is not None) else exchange
        self.texel = bool(texel)
        self.root = root
        self.timeout_ms = timeout_ms
        self.planes = plane_table(self.texel)
        self.rows = height // N
        self.band_rows = (band_rows or pick_band_rows(self.rows)) if N > 1 else 0
        self.chunks = chunks or pick_chunks(self.rows, N)
        if self.rows % self.chunks:
            raise ValueError("chunks must divide the rows of a rank")
        self.rpc = self.rows // self.chunks
        self.slots = slots if N > 1 else 1
        dev = f"cuda:{renderer.device}"
        self.dev = dev
        self.exchanged = [p for p in self.planes if exchange_gbuffer or not p[0].startswith("g_")]
        self.local = [p for p in self.planes if p not in self.exchanged]
        self.offsets, self.region_bytes = packed_layout(width, self.rpc, self.exchanged)
        payload = self.slots * self.chunks * N * self.region_bytes
        self._shared_ptr = None
        if self.exchange_mode in ("p2p", "p2pcopy"):
            # the packed buffer (+ flag words in front of it) lives on the root; everybody else maps it
            if self.rank == root:
                ptr, handle = renderer.shared_alloc(FLAG_BYTES + payload)
                box = [handle]
            else:
                ptr, box = None, [None]
            dist.broadcast_object_list(box, src=dist.get_global_rank(group, root) if group is not None else root, group=group)
            ok = torch.ones(1, dtype=torch.int32, device=dev)
            if self.rank != root:
                try:
                    ptr = renderer.shared_open(box[0])
                except abi.VxptError as e:  # no CUDA IPC / peer access between these processes
                    ok.zero_()
                    self._why = str(e)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)  # the verdict is collective: all ranks map it or none does
            if int(ok.item()) == 0:
                if ptr is not None:
                    renderer.shared_close(ptr)
                raise P2PUnavailable(getattr(self, "_why", "a peer rank could not map the root's slab buffer"))
            self._shared_ptr = ptr
            self._base = self._remote = ptr + FLAG_BYTES
            self.buf = (torch.as_tensor(_DevMem(self._base, payload), device=dev).view(self.slots, self.chunks, N, self.region_bytes)
                        if self.rank == root else None)
            if self.exchange_mode == "p2pcopy" and self.rank != root:
                # trace into a local copy of the layout; one DMA per frame pushes this rank's region to the root
                self._stage = torch.zeros((self.slots, self.chunks, N, self.region_bytes), dtype=torch.uint8, device=dev)
                self._base = self._stage.data_ptr()
        else:
            self.buf = torch.zeros((self.slots, self.chunks, N, self.region_bytes), dtype=torch.uint8, device=dev)
            self._base = self.buf.data_ptr()
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
                region = self._base + ((slot * self.chunks + c) * N + self.rank) * self.region_bytes
                base = {n: region + self.offsets[n] - vb * width * e for n, e, _, _ in self.exchanged}  # virtual plane bases
                base.update({n: self.loc.data_ptr() + self.loc_offsets[n] for n, e, _, _ in self.local})
                per_chunk.append(base)
            self.bases.append(per_chunk)
        renderer.set_option(abi.OPT_TEXEL_FORMAT, 1 if self.texel else 0)
        for cam in self.cams:  # scratch sized now: a frame may be captured into a CUDA graph before it ever ran eagerly
            renderer.reserve(cam, max_gi_spp=max_gi_spp)
        self._ext = torch.cuda.ExternalStream(renderer.cuda_stream(), device=dev)
        self._comm = torch.cuda.Stream(device=dev) if N > 1 else None
        self._traced = [[torch.cuda.Event() for _ in range(self.chunks)] for _ in range(self.slots)]
        self._exchanged = [None] * self.slots  # event of the last exchange that read/wrote a slot
        self._slot = 0
        # p2p / p2pcopy: device-resident sequence counters, so that the flag operations are argument-free and a whole frame — wait for the
        # slot, the passes, arrival flag — replays as one CUDA graph: [0] frames started, [1] frames signalled, [2] frames
        # gathered (root), [3] frames released (root).  One 128-byte line each.
        self._counters = torch.zeros(4 * FLAG_STRIDE_WORDS, dtype=torch.int32, device=dev)
        self.last_slot = 0

    def close(self):
        """Unmap / free the shared slab buffer (collective in p2p mode: peers unmap before the root frees)."""
        if self._shared_ptr is None:
            return
        self.finish()
        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        if self.rank != self.root:
            self.r.shared_close(self._shared_ptr)
        dist.barrier(group=self.group)
        if self.rank == self.root:
            self.buf = None
            self.r.shared_close(self._shared_ptr)
        self._shared_ptr = None

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

    def _flag(self, word):
        return self._shared_ptr + 4 * FLAG_STRIDE_WORDS * word

    def _counter(self, k):
        return self._counters.data_ptr() + 4 * FLAG_STRIDE_WORDS * k

    def frame_into(self, slot, primary, shadow, diffuse):
        """Everything a rank enqueues on the library stream for one frame; capturable into ONE CUDA graph in p2p mode (the flag
        operations take their sequence numbers from device-resident counters)."""
        if self.exchange_mode in ("p2p", "p2pcopy"):
            # frame k may overwrite its slot once the root has released frame k - slots
            self.r.wait_next(self._shared_ptr + FLAG_BYTES // 2, 1, FLAG_STRIDE_WORDS, self._counter(0), lag=self.slots, timeout_ms=self.timeout_ms)
            self.trace_into(slot, primary, shadow, diffuse)
            if self.exchange_mode == "p2pcopy" and self.rank != self.root:
                # copy engine, same stream: this rank's packed region -> the same region of the root's buffer.  The pipe's next frame
                # waits for the push (a few microseconds of DMA); the other pipes keep the SMs busy meanwhile.
                for c in range(self.chunks):
                    off = ((slot * self.chunks + c) * self.world_size + self.rank) * self.region_bytes
                    self.r.copy_async(self._remote + off, self._base + off, self.region_bytes)
            self.r.signal_next(self._flag(self.rank), self._counter(1))
        else:
            self.before_trace(slot)
            self.trace_into(slot, primary, shadow, diffuse)

    def gather_into(self, slot=None):
        """p2p, root only: what the gather stream does per frame (capturable): wait until every rank's rows of the next frame have
        arrived, (consumer work would go here), release the slot."""
        cs = self._comm.cuda_stream
        self.r.wait_next(self._flag(0), self.world_size, FLAG_STRIDE_WORDS, self._counter(2), lag=0, timeout_ms=self.timeout_ms, stream=cs)
        self.r.signal_next(self._shared_ptr + FLAG_BYTES // 2, self._counter(3), stream=cs)

    def before_trace(self, slot):
        """The slot's previous contents must have been consumed before it is overwritten (stream-ordered, no host sync)."""
        if self.exchange_mode in ("p2p", "p2pcopy"):
            return  # part of frame_into
        if self._exchanged[slot] is not None:
            self._ext.wait_event(self._exchanged[slot])

    def exchange(self, slot):
        """nccl: start the all-gather of a traced slot on the communication stream.  p2p: the rows are already in the root's
        memory; release this rank's arrival flag, and on the root let the gather stream wait for all ranks."""
        if self.exchange_mode == "none":
            return
        if self.exchange_mode in ("p2p", "p2pcopy"):
            if self.rank == self.root:
                self.gather_into(slot)
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
        asynchronously: the exchange of this frame overlaps the tracing of the next frame (library stream, other slot).
        Call finish() before reading planes or stopping a timer."""
        slot = self.next_slot()
        self.frame_into(slot, primary, shadow, diffuse)
        self.exchange(slot)

    def finish(self):
        """Make the library's stream wait for every outstanding exchange (no host sync).  In p2p mode the non-root ranks have
        nothing to wait for: their stores are complete when their own stream is."""
        if self.world_size > 1:
            self._ext.wait_stream(self._comm)

    # first interface
    def trace(self, primary, shadow, diffuse):
        self.render(primary, shadow, diffuse)

    def gather(self):
        self.finish()

    def plane(self, name, slot=None):
        """Full-frame tensor [H, W, ...] of an exchanged plane (or of this rank's rows for a sharded-only plane) of the last
        rendered frame, in IMAGE row order — a copy for consumers and tests.  In p2p mode exchanged planes exist on the root only."""
        slot = self.last_slot if slot is None else slot
        spec = next(p for p in self.planes if p[0] == name)
        _, elem, dtype, tail = spec
        tdt = getattr(torch, dtype)
        if spec in self.local:
            n = self.rows * self.width * elem
            local = self.loc[self.loc_offsets[name]:self.loc_offsets[name] + n].view(tdt).view((self.rows, self.width) + tail)
            return local
        if self.buf is None:
            raise RuntimeError("p2p exchange: gathered planes live on the root rank only")
        n = self.rpc * self.width * elem
        out = torch.empty((self.height, self.width) + tail, dtype=tdt, device=self.buf.device)
        for r in range(self.world_size):
            rows = torch.from_numpy(image_rows_of_rank(self.height, self.world_size, r, self.band_rows)).to(self.buf.device)
            part = torch.cat([self.buf[slot, c, r, self.offsets[name]:self.offsets[name] + n].view(tdt).view((self.rpc, self.width) + tail)
                              for c in range(self.chunks)], 0)
            out[rows] = part
        return out
