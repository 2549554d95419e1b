#!/usr/bin/env python3
"""bench.py — headline benchmark of the hot path (BASELINE.json): Mrays/s of primary + sun-shadow + diffuse-GI rays
at 1920x1080 on the procedural plains world, on N B200s (image row slabs, grid replicated, NCCL gather).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
  N > 1:  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

One "step" = one frame: primary rays (cap 350) -> soft sun shadow rays (cap 350) -> 1-spp 1-bounce diffuse GI (cap 48,
shadow sub-rays cap 128) over resident grid + distance field; the frame index (TAA jitter, blue-noise seeds) advances
every step.  Rays are counted as VoxelTraversalDF calls.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

# frame pipes + gather streams of the multi-GPU path each want their own hardware queue: a flag-wait at the head of a shared
# queue would stall unrelated streams behind it (must be set before the CUDA context exists)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTH, HEIGHT = 1920, 1080
WORKLOAD = ("plains world (FastNoise seeds 9383/6886), 1920x1080, camera (192,75,192) pitch -20: primary + soft sun shadow + "
            "1-spp 1-bounce diffuse GI per frame (BASELINE configs[2] without the reflection pass); distance field resident")
METRIC = "Mrays/s (primary+shadow+GI) at 1080p"

# --config: 3 = the headline workload above (BASELINE configs[2], what the driver runs); 4 and 5 = BASELINE configs[3] / configs[4] on their
# stand-in scenes, measured the same way and kept under profiles/ — not driver-run lines.
CONFIGS = {
    3: dict(width=1920, height=1080, world="plains", spp=1, camera=dict(pitch_deg=-20.0), edits=False, workload=WORKLOAD, metric=METRIC),
    4: dict(width=3840, height=2160, world="gi_box", spp=4, camera=dict(pitch_deg=-20.0), edits=False, metric="Mrays/s (primary+shadow+GI) at 2160p, 4 spp",
            workload="gi-box scene (stand-in for the missing 'Test Worlds/gi' blob: plains + hollow rooms + emissive lamps, fixed seed), 3840x2160, camera "
                     "(192,75,192) pitch -20: primary + soft sun shadow + 4-spp 1-bounce diffuse GI per frame (BASELINE configs[3]); distance field resident"),
    5: dict(width=1920, height=1080, world="city", spp=1, camera=dict(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0), edits=True,
            metric="Mrays/s (edit + DF rebuild + primary+shadow+GI) at 1080p",
            workload="city scene (stand-in for the missing 'Test MC Worlds/Medival' regions: towers, arches, interiors), 1920x1080, camera (100,60,100) yaw 45 "
                     "pitch -10: EVERY frame one block in view is placed or broken (World::Raycast path, Core/World.cpp:372-374,458-460), the distance "
                     "field and step field are rebuilt in full, then primary + soft sun shadow + 1-spp GI (BASELINE configs[4])"),
}
CFG = CONFIGS[3]


def edit_of(k):
    """config 5: the k-th frame's edit — a pillar in front of the camera grows block by block and is knocked down again (period 16)."""
    j = k % 16
    return (104, 60 + (j if j < 8 else 15 - j), 104, 3 if j < 8 else 0)


def frame_params(vx, camera, tables, frame):
    pp = vx.primary_params(350, camera.taa_jitter(frame))
    sp = vx.shadow_params(tables["stronger"], frame=frame, soft=True)
    dp = vx.diffuse_params(tables["sun"], tables["moon"], tables["sun_visibility"], spp=CFG["spp"], frame=frame)
    return pp, sp, dp


def build_world(world, assets):
    name = CFG["world"]
    if name == "plains":
        return world.generate_plains(assets.load_plains_columns())
    if name == "gi_box":
        return world.generate_gi_box(assets.load_plains_columns())
    return world.generate_city()


def load_tables():
    from voxelpathtracer_b200 import assets, camera
    sun, moon, stronger, vis = camera.sun_moon_direction(50.0)
    return {"materials": assets.load_materials(), "blue_noise": assets.load_blue_noise(), "sky": assets.analytic_sky(16, sun),
            "shadow_noise": assets.load_shadow_noise(), "sun": sun, "moon": moon, "stronger": stronger, "sun_visibility": vis}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The sampler is started
    before the warm-up so it is already running when the timed region begins; samples are attributed by timestamp."""
    Q = ("timestamp,index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.err = tempfile.mktemp(suffix=".err")
        self.proc = None
        self.t_begin = self.t_end = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=open(self.path, "w"), stderr=open(self.err, "w"))
            t0 = time.time()
            while time.time() - t0 < 5.0 and os.path.getsize(self.path) == 0:  # wait for the first sample
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def begin(self):
        self.t_begin = time.time()

    def end(self):
        self.t_end = time.time()

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "window": "none"}
        if self.proc is None:
            return out
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 10:
                    continue
                try:
                    ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except Exception:
                    ts = None
                reasons = {name for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[6:10])
                           if val.lower().startswith("active")}
                rows.append((ts, float(f[2]), float(f[3]), float(f[4]), reasons))
            os.unlink(self.path)
        except Exception:
            pass
        if not rows:
            try:
                out["error"] = open(self.err).read()[-200:]
            except Exception:
                pass
            return out
        timed = [r for r in rows if r[0] is not None and self.t_begin is not None and self.t_begin - 0.02 <= r[0] <= self.t_end + 0.02]
        use, window = (timed, "timed region") if len(timed) >= 3 else (rows, "warm-up + timed region (timed region shorter than 3 samples)")
        out.update(sm_mhz=float(np.median([r[1] for r in use])), sm_max_mhz=use[0][2], power_w_max=max(r[3] for r in use),
                   reasons=sorted(set().union(*[r[4] for r in use])), samples=len(use), window=window)
        return out


class NvmlClockSampler:
    """SM clock + throttle reasons read straight from NVML on a thread, one sample every ~0.2 ms: the driver's runs time 20 steps (10 ms at
    N = 1, 1.5 ms at N = 8), too short for nvidia-smi's 20 ms loop to land three samples inside the timed region.  Same
    output as ClockSampler; falls back to it when NVML is not importable."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        import threading
        import pynvml
        pynvml.nvmlInit()
        self.nv = pynvml
        self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        self.rows, self.stop_flag = [], False
        self.t_begin = self.t_end = None
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def _loop(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(h)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                self.rows.append((time.time(), mhz, mask))
            except Exception:
                pass
            time.sleep(0.0002)

    def begin(self):
        self.t_begin = time.time()

    def end(self):
        self.t_end = time.time()

    def stop(self):
        self.stop_flag = True
        self.thread.join(timeout=2)
        rows = list(self.rows)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "window": "none", "source": "NVML, polled on a thread"}
        if not rows:
            return out
        timed = [r for r in rows if self.t_begin is not None and self.t_begin <= r[0] <= self.t_end]
        use, window = (timed, "timed region") if len(timed) >= 3 else (rows, "warm-up + timed region (timed region shorter than 3 samples)")
        mask = 0
        for r in use:
            mask |= r[2]
        out.update(sm_mhz=float(np.median([r[1] for r in use])), reasons=sorted(n for n, bit in self.REASONS if mask & bit), samples=len(use), window=window)
        return out


def make_clock_sampler(gpu_index):
    try:
        return NvmlClockSampler(gpu_index)
    except Exception:
        return ClockSampler(gpu_index)


# ------------------------------------------------------------------------------------------------- reference arm
class CpuReference:
    """The reference's CPU implementation of the path, timed on the host cores.

    kind "reference": the reference's OWN shaders — InitialRayTraceFrag / ShadowRayTraceFrag / DiffuseRayTraceFrag.glsl, translated and
    compiled as C++ against its vendored glm when this repo was built next to /root/reference (oracle/_ref/libref_shaders.so; the file
    travels with the repo) — one process per host core, each tracing 8-row slabs (oracle/ref_pool.py).
    kind "port": oracle/vxo_oracle.cpp, the C++ restatement (OpenMP over rows), when that library is absent.
    A step is one frame of the bench workload, or — when a frame is too slow for the requested number of steps — every m-th 8-row slab
    of it (rows interleaved over the whole image, so sky, horizon and ground are sampled in proportion)."""

    def __init__(self, tables, w):
        import voxelpathtracer_b200 as vx
        from voxelpathtracer_b200 import camera
        from oracle import ref_shaders, vxo
        self.vx, self.camera, self.vxo, self.tables = vx, camera, vxo, tables
        self.orc = vxo.Oracle(w.data)   # ray counting (and the "port" arm)
        self.orc.set_tables(tables["materials"], tables["blue_noise"], tables["sky"], tables["shadow_noise"])
        self.fc = camera.FpsCamera(pitch_deg=-20.0)
        self.pool = None
        self.kind = "port"
        self.cores = vxo.load().vxo_num_threads()
        if ref_shaders.available():
            from oracle import ref_pool
            self.pool = ref_pool.ReferenceFramePool(w.data, vxo.df_build(w.data), WIDTH, HEIGHT, dict(pitch_deg=-20.0))
            self.kind, self.cores = "reference", self.pool.procs
        self.stride = 1

    def slabs(self):
        return [(rb, min(rb + 8, HEIGHT)) for k, rb in enumerate(range(0, HEIGHT, 8)) if k % self.stride == 0]

    def step(self, frame):
        if self.pool is not None:
            tasks = [(frame, rb, re) for rb, re in self.slabs()]
            from oracle import ref_pool
            return sum(r[2] for r in self.pool.pool.imap_unordered(ref_pool._slab, tasks, chunksize=1))
        tot = 0.0
        for rb, re in ([(0, HEIGHT)] if self.stride == 1 else self.slabs()):
            tot += self._oracle_rows(frame, rb, re)[1]
        return tot

    def _oracle_rows(self, frame, rb, re):
        cam = self.fc.vx_camera(WIDTH, HEIGHT, rb, re)
        pp, sp, dp = frame_params(self.vx, self.camera, self.tables, frame)
        g, s0 = self.orc.trace_primary(cam, pp, hit_voxel=False)
        _, s1 = self.orc.trace_shadow(cam, g, sp)
        d, s2 = self.orc.trace_diffuse(cam, g, dp)
        return s0["rays"] + s1["rays"] + s2["rays"], float(d["luma"][rb:re].sum())

    def rays(self, frame):
        """VoxelTraversalDF calls of the rows a step covers (counted by the oracle: the counts are a property of the algorithm)."""
        if self.stride == 1:
            return self._oracle_rows(frame, 0, HEIGHT)[0]
        return sum(self._oracle_rows(frame, rb, re)[0] for rb, re in self.slabs())

    def calibrate(self, steps, budget_s):
        """Pick the slab stride so that `steps` steps fit the time budget; returns the seconds one full frame took."""
        t0 = time.perf_counter()
        self.step(0)
        t_frame = time.perf_counter() - t0
        self.stride = max(1, int(np.ceil(t_frame * steps / budget_s)))
        return t_frame

    def sample_text(self, steps):
        rows = sum(re - rb for rb, re in self.slabs())
        what = ("the reference's own GLSL shaders compiled as C++ (oracle/_ref/libref_shaders.so), one process per core" if self.kind == "reference"
                else "oracle C++ restatement, OpenMP over image rows")
        return f"{steps} steps of {rows} of {HEIGHT} rows of a 1080p frame (every {self.stride}. 8-row slab; primary+shadow+GI), {what}"

    def close(self):
        if self.pool is not None:
            self.pool.close()


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path on every host core, same config / metric.  Rank 0 alone works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from voxelpathtracer_b200 import assets, world
    tables = load_tables()
    if args.config != 3:
        raise SystemExit("the reference arm is measured on the headline workload (--config 3)")
    w = world.generate_plains(assets.load_plains_columns())
    ref = CpuReference(tables, w)
    ref.calibrate(args.steps + args.warmup, 150.0)
    for f in range(args.warmup):
        ref.step(f)
    t0 = time.perf_counter()
    for f in range(args.steps):
        ref.step(args.warmup + f)
    dt = time.perf_counter() - t0
    rays = sum(ref.rays(args.warmup + f) for f in range(args.steps))
    ref.close()
    v = rays / dt / 1e6
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "Mrays/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD},
        "cpu_baseline": {"value": v, "unit": "Mrays/s", "cores": ref.cores, "kind": ref.kind, "sample": ref.sample_text(args.steps)},
        "e2e": {"value": v, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))
    return 0


def bind_near_gpu(gpu_index):
    """Pin this process to the CPUs of the GPU's NUMA node (NVML's ideal affinity) while the pinned host planes are allocated and
    the end-to-end loop runs: a device->host copy into memory of the far socket is 20-30 % slower.  Returns the previous affinity
    (restored before the CPU baseline, which wants every core), or None if nothing was changed."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        ideal = {64 * w + b for w, word in enumerate(words) for b in range(64) if (int(word) >> b) & 1}
        before = os.sched_getaffinity(0)
        want = ideal & before
        if not want or want == before:
            return None
        os.sched_setaffinity(0, want)
        return before
    except Exception:
        return None


def check_gathered_frame(vx, fr, r, fc, frame_params_, texel, barrier):
    """Multi-GPU correctness inside the bench (un-timed): every rank traces its rows of one more frame into slot 0 and exchanges them as
    in the timed loop; the rank that holds the gathered planes renders the WHOLE frame by itself and compares the exchanged planes
    (shadow + GI) byte for byte.  Returns {"ok": bool, "planes": {...}} on that rank, None elsewhere."""
    import torch
    pp, sp, dp = frame_params_
    fr.finish()
    torch.cuda.synchronize()
    barrier()
    fr.last_slot = 0
    fr.frame_into(0, pp, sp, dp)
    fr.exchange(0)
    fr.finish()
    torch.cuda.synchronize()
    barrier()   # every rank's push and arrival flag have completed; the root's gather stream has seen all of them
    out = None
    if fr.buf is not None and fr.rank == fr.root:
        cam = fc.vx_camera(fr.width, fr.height)
        g = r.alloc_gbuffer(fr.width, fr.height, device=True, texel=texel)
        s = r.alloc_shadow(fr.width, fr.height, device=True, texel=texel)
        d = r.alloc_diffuse(fr.width, fr.height, device=True, texel=texel)
        r.trace_primary(cam, pp, g)
        r.trace_shadow(cam, g, sp, s)
        r.trace_diffuse(cam, g, dp, d)
        r.sync()
        whole = {"s_shadow": s["shadow"], "s_transversal": s["transversal"], "d_sh": d["sh"], "d_cocg": d["cocg"], "d_luma": d["luma"], "d_ao_sky": d["ao_sky"]}
        planes = {}
        for name, _, _, _ in fr.exchanged:
            if name not in whole:
                continue
            got, want = fr.plane(name, 0), whole[name]
            planes[name] = bool(torch.equal(got.contiguous().view(torch.uint8), want.contiguous().view(torch.uint8)))
        out = {"ok": all(planes.values()) and len(planes) > 0, "planes": planes,
               "digest": int(sum(int(whole[n].contiguous().view(torch.uint8).to(torch.int64).sum()) for n in planes) % (1 << 61))}
    barrier()
    return out


# ------------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    import voxelpathtracer_b200 as vx
    from voxelpathtracer_b200 import abi, assets, camera, multigpu, world

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={ws}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if ws > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    dev = torch.device(f"cuda:{local_rank}")

    tables = load_tables()
    WIDTH, HEIGHT, WORKLOAD = CFG["width"], CFG["height"], CFG["workload"]
    w = build_world(world, assets)
    fc = camera.FpsCamera(**CFG["camera"])
    texel = args.planes == "texel"
    px_out = sum(e for n_, e, _, _ in multigpu.plane_table(texel))                                # bytes per pixel written per frame
    px_xchg = sum(e for n_, e, _, _ in multigpu.plane_table(texel) if not n_.startswith("g_"))    # ... of which cross the link
    exchange = args.exchange
    # P independent handles ("pipes") per GPU, each with its own CUDA stream and double-buffered frame slots: consecutive
    # frames go to alternating pipes, so the latency-bound tail of one frame's kernels overlaps the next frame's kernels.
    P = max(1, args.pipes if args.pipes > 0 else (2 if ws == 1 else (4 if ws == 2 else 8)))
    if CFG["edits"]:
        # the frames of config 5 depend on each other through the world (frame k sees edits 0..k): one pipe, eager submission — every
        # step is vxpt_set_block + vxpt_build_distance_field + the three passes, in stream order
        P, args.no_graph = 1, True
    renderers, frames, exts = [], [], []
    for _ in range(P):
        rr = vx.Renderer(local_rank)
        rr.load_scene_tables(tables["materials"], tables["blue_noise"], tables["sky"], tables["shadow_noise"])
        rr.upload_world(w)
        rr.build_distance_field()
        # Timing rule "inputs larger than L2": three copies of grid + step field (132 MB > 126 MB L2) are rotated frame by
        # frame and every frame streams its output planes through the cache; no flush kernel inside the timed region.
        rr.set_option(abi.OPT_SCENE_REPLICAS, 3)
        rr.set_option(abi.OPT_GI_WAVEFRONT, args.gi_mode)
        renderers.append(rr)
        try:
            frames.append(multigpu.ShardedFrame(rr, fc, WIDTH, HEIGHT, exchange=exchange, texel=texel, slots=args.slots,
                                                emulate=(args.emulate, 0) if args.emulate else None, max_gi_spp=CFG["spp"]))
        except multigpu.P2PUnavailable as e:  # raised on every rank together: the NCCL all-gather still works
            sys.stderr.write(f"[bench] rank {rank}: p2p slab gather unavailable ({e}); falling back to the NCCL all-gather\n")
            exchange = "nccl"
            frames.append(multigpu.ShardedFrame(rr, fc, WIDTH, HEIGHT, exchange=exchange, texel=texel, slots=args.slots, max_gi_spp=CFG["spp"]))
        exts.append(torch.cuda.ExternalStream(rr.cuda_stream(), device=dev))
    r, frame, ext = renderers[0], frames[0], exts[0]
    for _ in range(3):
        r.build_distance_field()
    df_times = []
    for _ in range(10):
        r.build_distance_field()
        st = r.stats()
        df_times.append((st["df_build_ms"], st["brick_pack_ms"]))
    df_ms = float(np.median([a for a, _ in df_times]))
    pack_ms = float(np.median([b for _, b in df_times]))
    l2_peak = r.measure_l2_sector_peak()

    def barrier():
        if ws > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n_frames = args.warmup + args.steps
    params = [frame_params(vx, camera, tables, f) for f in range(n_frames)]  # tiny host structs, one per frame index
    slots = frame.slots
    M = P * slots * max(1, -(-16 // (P * slots)))  # >= 16 frame indices per cycle (a multiple of P * slots); step k -> pipe k % P, slot (k // P) % slots

    def eager_step(k, f):
        fr = frames[k % P]
        slot = (k // P) % slots
        fr.last_slot = slot
        if CFG["edits"]:  # every rank patches its replica of the grid (the edit list is a function of the frame index: nothing to broadcast)
            rr_ = renderers[k % P]
            rr_.set_block(*edit_of(f))
            rr_.build_distance_field()
        fr.frame_into(slot, *params[f])
        fr.exchange(slot)

    def finish_all():
        """Stream 0 waits for every pipe (and its exchanges): the end event recorded on it closes the timed region."""
        for fr in frames:
            fr.finish()
        for e in exts[1:]:
            ext.wait_stream(e)

    # ---- device-resident timing: exactly K steps between two CUDA events, bracketed by barrier + synchronize.
    # Submission: the library calls of one frame (this rank's rows) are captured into a CUDA graph per frame index (M indices,
    # cycled), so a step costs the host one graph launch + the eager NCCL exchange instead of ~0.34 ms of Python/ctypes per
    # frame.  --no-graph submits every call eagerly.
    sampler = make_clock_sampler(local_rank) if rank == 0 else None
    # every pipe x slot runs eagerly at least once before anything is captured (first-use work — module loading, scratch growth — is
    # not capturable; the handles' scratch is also sized up front by ShardedFrame through vxpt_reserve)
    for k in range(max(args.warmup, P * slots)):
        eager_step(k, k % n_frames)
    finish_all()
    barrier()
    graphs, submit = None, "eager"
    launches_per_graph = None
    # p2p: the flag operations take their sequence numbers from device counters, so the whole frame (wait for the slot, passes,
    # arrival flag) is one graph, and the root's gather stream replays a two-node graph of its own
    whole = frame.exchange_mode in ("p2p", "p2pcopy")
    is_root = whole and rank == frame.root
    gather_graphs = None
    if not args.no_graph:
        try:
            for rr in renderers:
                rr.set_option(abi.OPT_TIMING_EVENTS, 0)
            l0 = sum(rr.launch_count() for rr in renderers)
            graphs = []
            for m in range(M):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=exts[m % P], capture_error_mode="thread_local"):
                    fr_m = frames[m % P]
                    (fr_m.frame_into if whole else fr_m.trace_into)((m // P) % slots, *params[(args.warmup + m) % n_frames])
                graphs.append(gph)
            launches_per_graph = (sum(rr.launch_count() for rr in renderers) - l0) / M
            if is_root:
                gather_graphs = []
                for fr_p in frames:
                    gph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gph, stream=fr_p._comm, capture_error_mode="thread_local"):
                        fr_p.gather_into()
                    gather_graphs.append(gph)
                launches_per_graph += 2
            submit = (f"one CUDA graph per frame ({'wait for the slot, ' if whole else ''}trace passes of this rank's rows{', arrival flag' if whole else ''}; "
                      f"{M} frame indices cycled) on {P} pipe(s)" + ("; the root's gather stream replays a 2-node graph per frame" if whole else ", exchange eager"))
        except Exception as e:  # capture unsupported in this environment: fall back to eager submission
            sys.stderr.write(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); submitting eagerly\n")
            graphs, submit = None, "eager (graph capture failed)"
            torch.cuda.synchronize()
        finally:
            for rr in renderers:
                rr.set_option(abi.OPT_TIMING_EVENTS, 1)

    def graph_step(k):
        fr = frames[k % P]
        slot = (k // P) % slots
        fr.last_slot = slot
        if not whole:
            fr.before_trace(slot)
        with torch.cuda.stream(exts[k % P]):
            graphs[k % M].replay()
        if not whole:
            fr.exchange(slot)
        elif is_root:
            with torch.cuda.stream(fr._comm):
                gather_graphs[k % P].replay()

    step = graph_step if graphs is not None else (lambda k: eager_step(k, args.warmup + k))
    for fr in frames:
        fr._exchanged = [None] * fr.slots
    for k in range(M):
        step(k)  # untimed
    finish_all()
    barrier()
    for rr in renderers:
        rr.reset_stats()
    launches0 = sum(rr.launch_count() for rr in renderers)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if sampler:
        sampler.begin()
    t_submit = time.perf_counter()
    ev0.record(ext)
    for e in exts[1:]:
        e.wait_event(ev0)  # no pipe starts before the start event
    for k in range(args.steps):
        step(k)
    finish_all()  # the last frames' exchanges are inside the timed region
    ev1.record(ext)
    host_submit_ms = (time.perf_counter() - t_submit) * 1e3 / args.steps
    barrier()
    if sampler:
        sampler.end()
    clocks = sampler.stop() if sampler else None
    sts = [rr.stats() for rr in renderers]
    st = {k_: sum(s_[k_] for s_ in sts) for k_ in ("rays", "df_fetches", "vox_fetches")}
    launches = (launches_per_graph * args.steps) if graphs is not None else (sum(rr.launch_count() for rr in renderers) - launches0)
    step_ms = float(ev0.elapsed_time(ev1))
    tot = torch.tensor([step_ms, float(st["rays"]), float(st["df_fetches"]), float(st["vox_fetches"]), float(launches)], dtype=torch.float64, device=dev)
    mx = tot.clone()
    if ws > 1:
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    total_ms = float(mx[0])                      # max over ranks of the K-step time
    rays_all = float(tot[1])                     # all ranks, all K steps
    value = rays_all / (total_ms * 1e-3) / 1e6

    # ---- un-timed: is the gathered frame the frame?  One more frame goes through the same path (slot wait, passes, push, flags /
    # all-gather); the gather root then renders the whole frame alone and compares every exchanged plane bit for bit.
    gathered = None
    if ws > 1 and not args.emulate:
        gathered = check_gathered_frame(vx, frames[0], renderers[0], fc, params[args.warmup], texel, barrier)

    # per-pass rooflines on this rank (rank 0 reports): algorithmic bytes = (N_it + N_vox) * 32 B over the L2 sector peak
    # measured by the library's probe.  Counters are per pass, so re-run K frames with a stats read between passes.
    pass_stats = np.zeros((3, 3))
    per_pass = np.zeros((args.steps, 3))  # ms per pass on this rank's rows, from the library's own CUDA events (its stream)
    for k in range(args.steps):
        pp, sp, dp = params[args.warmup + k]
        for c in range(frame.chunks):
            g_, s_, d_ = frame._planes(c)
            cam_c = frame.cams[c]
            for i, call in enumerate((lambda: r.trace_primary(cam_c, pp, g_), lambda: r.trace_shadow(cam_c, g_, sp, s_),
                                      lambda: r.trace_diffuse(cam_c, g_, dp, d_))):
                r.reset_stats()
                call()
                st_ = r.stats()
                pass_stats[i] += (st_["rays"], st_["df_fetches"], st_["vox_fetches"])
                per_pass[k, i] += st_["last_ms"]
    names = ("primary_kernel", "shadow_kernel", "diffuse_pass")
    rooflines = {}
    for i, n in enumerate(names):
        ms = float(per_pass[:, i].sum())
        gbs = (pass_stats[i, 1] + pass_stats[i, 2]) * 32.0 / (ms * 1e-3) / 1e9
        rooflines[n] = {"bound": "l2", "achieved": gbs, "peak": l2_peak, "unit": "GB/s", "frac": gbs / l2_peak, "traffic": None,
                        "ms_per_launch": ms / args.steps, "rays_per_launch": pass_stats[i, 0] / args.steps,
                        "fetches_per_ray": (pass_stats[i, 1] + pass_stats[i, 2]) / max(pass_stats[i, 0], 1.0),
                        "share_of_step": ms / max(float(per_pass.sum()), 1e-9)}
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, hbm_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    # a rebuild is not usable by the trace passes before the step field exists, so the denominator is distance field + step field time;
    # the algorithmic bytes stay SURVEY.md §8(d)'s: read the grid once + write the distance field once (the step field is this design's own)
    df_gbs = 2 * abi.WORLD_VOXELS / ((df_ms + pack_ms) * 1e-3) / 1e9
    rooflines["df_build"] = {"bound": "hbm", "achieved": df_gbs, "peak": hbm_peak, "peak_source": hbm_src, "unit": "GB/s", "frac": df_gbs / hbm_peak,
                             "traffic": None, "ms_per_launch": df_ms + pack_ms, "distance_field_ms": df_ms, "step_field_ms": pack_ms,
                             "note": "37,748,736 algorithmic bytes over the whole rebuild (distance field + traversal step field)"}
    # second roofline of the traversal passes: warp instructions per launch (committed ncu count of the same frame, profiles/issue.json)
    # over the live pass time and the issue peak — 148 SMs x 4 schedulers x 1 warp instruction per cycle at the SM clock sampled in the
    # timed region.  The L2-sector figure above is SURVEY.md §8d's; the kernels are bound by instruction issue (ncu: L2 sector
    # throughput 1-10 % of peak, issue slots 64-75 % busy), which this fraction shows.
    issue_path = os.path.join(ROOT, "profiles", "issue.json")
    if os.path.exists(issue_path) and args.config == 3 and ws == 1:
        try:
            issue = json.load(open(issue_path))
            mhz = (clocks or {}).get("sm_mhz") or 1965.0
            for n in names:
                if n in issue["passes"]:
                    peak_ips = 148 * 4 * mhz * 1e6
                    rooflines[n]["issue_frac"] = issue["passes"][n] / (rooflines[n]["ms_per_launch"] * 1e-3) / peak_ips
                    rooflines[n]["warp_instructions"] = issue["passes"][n]
        except Exception:
            pass
    traffic_path = os.path.join(ROOT, "profiles", "traffic.json")  # dram bytes per launch from the committed ncu --set full capture
    if os.path.exists(traffic_path):
        for k, v in json.load(open(traffic_path)).items():
            if k in rooflines:
                rooflines[k]["traffic"] = v
    dominant = max(names, key=lambda n: rooflines[n]["ms_per_launch"])

    # ---- end to end through the C ABI with HOST buffers (pinned): camera/params in, every output plane out --------------
    # One vxpt_render_frame call per step on this rank's rows: the G-buffer stays on the device between the passes (as the
    # reference's FBO textures do) and the planes are copied to pinned host memory slab by slab while later slabs trace.
    for fr in frames:
        fr.finish()
    torch.cuda.synchronize()
    affinity_before = bind_near_gpu(local_rank)
    slab_b, slab_e = multigpu.slab_rows(HEIGHT, ws, rank)
    cam_rank = fc.vx_camera(WIDTH, HEIGHT, slab_b, slab_e)
    rows = slab_e - slab_b
    # double-buffered like a swap chain: two handles and two sets of pinned planes alternate (vxpt_render_frame_async), so the copy-out
    # of frame k overlaps the tracing of frame k+1; a frame's result is read when its handle comes round again (vxpt_frame_wait)
    hr = renderers[:max(2, min(len(renderers), args.e2e_handles))]
    for h in hr:
        h.set_option(abi.OPT_TEXEL_FORMAT, 1 if texel else 0)
    bufs = [(h.alloc_gbuffer(WIDTH, HEIGHT, texel=texel, pinned=True), h.alloc_shadow(WIDTH, HEIGHT, texel=texel, pinned=True),
             h.alloc_diffuse(WIDTH, HEIGHT, texel=texel, pinned=True)) for h in hr]
    d2h = rows * WIDTH * px_out
    h2d = 144 + 24 + 32 + 72  # VxCamera + VxPrimaryParams + VxShadowParams + VxDiffuseParams: the only per-frame inputs of the path
    checksum = 0.0

    # the call of a handle with its camera and output planes bound once (Renderer.prepare_frame): per frame the host passes three pointers
    submits = [h.prepare_frame(cam_rank, *b) for h, b in zip(hr, bufs)]
    waits = [h.frame_wait for h in hr]
    lumas = [b[2]["luma"] for b in bufs]
    row0 = cam_rank.row_begin

    # The frame call itself — kernels of the three passes and the pinned device->host copies of every plane — recorded into one CUDA graph
    # per (handle, frame index), G of them cycled: per step the host then issues ONE graph launch instead of ~25 runtime calls (at 8 ranks
    # on a 16-core host those calls, not the PCIe link, bounded the loop: 0.6 ms per frame for 7 MB, r02u).  A replay is complete when the
    # handle's stream is, so the wait is vxpt_sync.  --no-graph (or a failed capture) keeps the eager calls.
    H_e2e = len(hr)
    G_e2e = H_e2e * 8
    e2e_graphs, e2e_submit = None, "eager vxpt_render_frame_async calls"
    if not args.no_graph:
        try:
            for h in hr:
                h.frame_wait()
                h.sync()
                h.set_option(abi.OPT_TIMING_EVENTS, 0)
            for gi_ in range(H_e2e):     # one eager frame per handle: scratch and staging sized before anything is captured
                submits[gi_](*params[gi_ % len(params)])
            for h in hr:
                h.frame_wait()
                h.sync()
            e2e_graphs = []
            for gidx in range(G_e2e):
                gph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gph, stream=exts[gidx % H_e2e], capture_error_mode="thread_local"):
                    submits[gidx % H_e2e](*params[(args.warmup + gidx) % len(params)])
                e2e_graphs.append(gph)
            e2e_submit = f"one CUDA graph launch per step: vxpt_render_frame_async (three passes + pinned device->host copies) recorded per handle and frame index, {G_e2e} cycled"
        except Exception as e:
            sys.stderr.write(f"[bench] e2e graph capture failed ({type(e).__name__}: {e}); calling vxpt_render_frame_async eagerly\n")
            e2e_graphs = None
            torch.cuda.synchronize()
        finally:
            for h in hr:
                h.set_option(abi.OPT_TIMING_EVENTS, 1)
    syncs = [h.sync for h in hr]

    def host_step(k, f):
        i = k % H_e2e
        if e2e_graphs is not None:
            syncs[i]()                                            # the frame this handle rendered last: kernels done, host planes landed
            val = float(lumas[i][row0, 0])                        # ... read (part of) its result
            with torch.cuda.stream(exts[i]):
                e2e_graphs[k % G_e2e].replay()
            return val
        waits[i]()                                                # the frame this handle rendered last has landed in its host planes
        val = float(lumas[i][row0, 0])                            # ... read (part of) its result
        pp, sp, dp = params[f % len(params)]  # the per-frame uniforms (camera jitter, frame seeds): built once, passed by pointer per call
        submits[i](pp, sp, dp)
        return val

    for k in range(args.warmup):
        host_step(k, k)
    for h in hr:
        h.frame_wait()
        if e2e_graphs is not None:
            h.sync()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        checksum += host_step(k, args.warmup + k)
    for h in hr:
        h.frame_wait()
        if e2e_graphs is not None:
            h.sync()
    checksum += sum(float(b[2]["luma"][cam_rank.row_begin, 0]) for b in bufs)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if ws > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_step_s = float(e2e_s[0]) / args.steps
    e2e = {"value": rays_all / float(e2e_s[0]) / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": e2e_step_s * 1e3, "pcie_gbs_per_gpu": d2h / e2e_step_s / 1e9,
           "pcie_note": "device->host bytes of this rank's slab per second of the end-to-end loop; a PCIe 5.0 x16 link sustains about 50-55 GB/s to pinned memory, which is what bounds N = 1",
           "submit": e2e_submit,
           "note": (f"one vxpt_render_frame_async call per step with pinned HOST output planes, {len(hr)} handles / plane sets in flight (" + ("the reference's FBO texel formats, 27 B/pixel" if texel else "fp32 planes, 51 B/pixel") +
                    "): G-buffer resident on the device between passes, planes copied device->host slab by slab while later slabs trace; "
                    "the per-frame inputs are the camera and parameter structs" + ("; process bound to the GPU's NUMA node" if affinity_before is not None else ""))}

    if affinity_before is not None:
        os.sched_setaffinity(0, affinity_before)

    # ---- the passes either side of the headline path at the same 1080p frame, planes resident in device memory:
    # reflection pass (config 3: 1 spp, rough, Halton-jittered G-buffer read), G-buffer material pass, SVGF chain, shadow filters.  Timed by
    # the library's events around each call (tools/denoise_probe.py), N = 1 only; not part of the headline metric.
    aux_ms = None
    if rank == 0 and ws == 1 and not args.no_aux and args.config == 3:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import denoise_probe
            ra = vx.Renderer(local_rank)
            ra.upload_world(w)
            ra.build_distance_field()
            first = denoise_probe.measure(ra, iters=8, W=WIDTH, H=HEIGHT, warm=3)["passes"]       # the frames right after a history reset
            aux = denoise_probe.measure(ra, iters=min(args.steps, 10), W=WIDTH, H=HEIGHT)["passes"]  # steady state (16 frames of history)
            aux_ms = {"reflection": aux["reflection"]["ms"], "gbuffer": aux["material"]["ms"], "svgf_chain": aux["svgf_frame"]["ms"],
                      "shadow_filters": aux["shadow_filter_frame"]["ms"],
                      "svgf_kernels": {k: aux[k]["ms"] * aux[k]["calls_per_frame"] for k in ("svgf_initial", "svgf_temporal", "svgf_variance", "svgf_spatial")},
                      "frac_of_hbm": {k: aux[k]["frac"] for k in ("reflection", "material", "svgf_frame", "shadow_filter_frame")},
                      "svgf_chain_first_frames_after_reset": first["svgf_frame"]["ms"],
                      "note": "svgf_chain = a denoiser in steady state (16 frames of history before the timed ones); while a pixel's history is shorter than 12 frames "
                              "VarianceEstimate.glsl filters 9 x 9 taps around it, so the first frames after a reset cost svgf_chain_first_frames_after_reset"}
            # the same passes at the plane sizes the reference's defaults give a 1080p window (Core/Pipeline.cpp:72,101,129: GI and its
            # denoiser and the reflections at 0.25 of the window, the shadow pass and its filters at 0.5)
            q = denoise_probe.measure(ra, iters=min(args.steps, 10), W=WIDTH // 4, H=HEIGHT // 4)["passes"]
            hlf = denoise_probe.measure(ra, iters=min(args.steps, 10), W=WIDTH // 2, H=HEIGHT // 2)["passes"]
            aux_ms["at_reference_default_scales"] = {"svgf_chain_480x270": q["svgf_frame"]["ms"], "reflection_480x270": q["reflection"]["ms"],
                                                     "shadow_filters_960x540": hlf["shadow_filter_frame"]["ms"]}
            del ra
        except Exception as e:  # the headline line must not depend on the auxiliary passes
            aux_ms = {"error": f"{type(e).__name__}: {e}"[:200]}

    # ---- CPU baseline beside it (rank 0, N = 1 only): the oracle port on the host cores, bounded sample -------------
    cpu_baseline = None
    if rank == 0 and ws == 1 and not args.no_cpu_baseline and args.config == 3:
        ref = CpuReference(tables, w)
        t_frame = ref.calibrate(4, 30.0)   # even four frames must fit; then size the sample to about 12 s of CPU work
        n_cpu = max(4, min(32, int(12.0 / (t_frame / ref.stride))))
        t0 = time.perf_counter()
        for f in range(n_cpu):
            ref.step(f)
        dt = time.perf_counter() - t0
        rays_cpu = sum(ref.rays(f) for f in range(n_cpu))
        cpu_baseline = {"value": rays_cpu / dt / 1e6, "unit": "Mrays/s", "cores": ref.cores, "kind": ref.kind, "sample": ref.sample_text(n_cpu) + f", {dt:.1f} s"}
        ref.close()

    xchg_text = {"none": "single GPU, no exchange",
                 "p2pcopy": "radiance slabs (shadow + GI planes) traced locally and pushed into the gather root's memory with one copy-engine transfer per frame over NVLink (CUDA IPC mapping), "
                            "frames ordered by release/acquire flag words, no collective",
                 "p2p": "radiance slabs (shadow + GI planes) stored by the trace kernels straight into the gather root's memory over NVLink (CUDA IPC mapping), "
                        "frames ordered by release/acquire flag words, no exchange step",
                 "nccl": "one packed NCCL all-gather of the shadow + GI planes per frame on its own stream, overlapped with the following frames' tracing"}[frame.exchange_mode]
    if rank == 0:
        line = {
            "metric": CFG["metric"], "value": value, "unit": "Mrays/s", "n_gpus": ws, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "resolution": [WIDTH, HEIGHT], "rays_per_step": rays_all / args.steps,
                       "sharding": (f"{ws} ranks, interleaved {frame.band_rows}-row bands, {P} frame pipe(s) per GPU, grid replicated; " + xchg_text),
                       "planes": ("reference FBO texel formats (R16F/RGBA16F/RG16F/RG8/R8; VXPT_OPT_TEXEL_FORMAT=1)" if texel else "fp32 planes") + f", {px_out} B/pixel written, {px_xchg} B/pixel gathered",
                       "timing": f"two CUDA events on the library stream around exactly K steps (barrier + synchronize on both sides), max over ranks; inputs larger than L2: 3 scene replicas (132 MB) rotated per frame + {WIDTH * HEIGHT * px_out / 1e6:.0f} MB of planes written per frame, no flush kernel",
                       "submit": submit, "host_submit_ms_per_step": host_submit_ms, "traversal_layout": "8x4x4-voxel tiles of pre-converted step values", "gi": {0: "one thread per pixel", 1: "wavefront (sorted first-bounce rays, warp-ballot compaction of the hits, the rest of a sample in CTA-wide stages)"}[args.gi_mode]},
            "e2e": e2e, "gpu_launches": int(tot[4]),
            "roofline": dict(rooflines[dominant], kernel=dominant,
                             note="traversal roofline = (DF fetches + block fetches) x 32 B per launch over the measured random-sector L2 peak (SURVEY.md §8d)"),
            "roofline_all": rooflines,
            "df_build_ms": df_ms + pack_ms, "l2_sector_peak_gbs": l2_peak,
            "pass_ms": {"primary": float(per_pass[:, 0].mean()), "shadow": float(per_pass[:, 1].mean()), "diffuse": float(per_pass[:, 2].mean()),
                        **({k: v for k, v in aux_ms.items() if k in ("reflection", "gbuffer", "svgf_chain", "shadow_filters", "error")} if aux_ms else {}),
                        "note": "rank 0's rows, each pass timed alone after the timed region (library events); the step time above includes the overlapped exchange"
                                + ("; reflection (config 3: 1 spp, rough, jittered G-buffer read), gbuffer (material pass), svgf_chain and shadow_filters are the passes "
                                   "either side of the headline path, device planes, outside the metric" if aux_ms else "")},
            "aux_passes": aux_ms,
            "cpu_baseline": cpu_baseline, "clocks": clocks,
        }
        if CFG["edits"]:
            line["rebuild"] = {"ms": df_ms + pack_ms, "share_of_step": (df_ms + pack_ms) / (total_ms / args.steps),
                               "note": "one vxpt_set_block + one full distance-field and step-field rebuild inside EVERY timed step, in stream order before the frame's passes; "
                                       "ms = median of 10 rebuilds timed by the library's events before the timed region"}
        if gathered is not None:
            line["gathered_ok"] = gathered["ok"]
            line["gathered"] = gathered
        print(json.dumps(line))
    if ws > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=3, choices=[3, 4, 5],
                    help="3 = the headline workload (BASELINE configs[2], default); 4 = 3840x2160 4-spp GI on the gi-box scene; 5 = per-frame block edit + distance-field rebuild + 1080p GI on the city scene")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipes", type=int, default=0, help="independent handles/streams per GPU (frames in flight); 0 = 2 on one GPU, 4 on two, 8 on more")
    ap.add_argument("--slots", type=int, default=2, help="frame slots per pipe in the slab buffer")
    ap.add_argument("--exchange", default="p2pcopy", choices=["p2p", "p2pcopy", "nccl"],
                    help="multi-GPU slab gather: the trace kernels store into the root's memory (p2p), one copy-engine push per frame (p2pcopy), or NCCL all-gather")
    ap.add_argument("--gi-mode", type=int, default=1, choices=[0, 1], help="VXPT_OPT_GI_WAVEFRONT: 0 one thread per pixel, 1 wavefront (default)")
    ap.add_argument("--emulate", type=int, default=0, help="development: trace rank 0's share of an N-way sharded frame on one GPU, no exchange")
    ap.add_argument("--planes", default="texel", choices=["texel", "f32"], help="plane encoding: the reference's FBO texel formats (default) or fp32")
    ap.add_argument("--e2e-handles", type=int, default=4, help="handles / pinned plane sets in flight in the end-to-end loop (at most --pipes)")
    ap.add_argument("--no-aux", action="store_true", help="skip timing the reflection / G-buffer / denoiser passes (N = 1 only)")
    ap.add_argument("--no-graph", action="store_true", help="submit every pass eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    global CFG
    CFG = CONFIGS[args.config]
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
