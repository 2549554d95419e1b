#!/usr/bin/env python3
"""Multi-process check of the row-slab gather (run under torch.distributed.run, one rank per GPU): every rank traces its interleaved
row bands of a few frames through multigpu.ShardedFrame; the gathered planes of the last frame must equal a full-frame render on
one GPU, bit for bit.  Usage: python -m torch.distributed.run --nproc-per-node N tools/p2p_check.py [p2p|nccl] [texel|f32]"""
import os
import sys

import numpy as np

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import voxelpathtracer_b200 as vx  # noqa: E402
from voxelpathtracer_b200 import abi, assets, camera, multigpu, world  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "p2p"
    texel = (sys.argv[2] if len(sys.argv) > 2 else "texel") == "texel"
    rank, lr, ws = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{lr}"))
    W, H = 640, 360 if 360 % ws == 0 else 45 * ws
    sun, moon, stronger, vis = camera.sun_moon_direction(50.0)
    r = vx.Renderer(lr)
    r.load_scene_tables(assets.load_materials(), assets.load_blue_noise(), assets.analytic_sky(16, sun), assets.load_shadow_noise())
    r.upload_world(world.generate_plains(assets.load_plains_columns()))
    r.build_distance_field()
    fc = camera.FpsCamera(pitch_deg=-20.0)
    f = multigpu.ShardedFrame(r, fc, W, H, exchange=mode, texel=texel, slots=2, timeout_ms=5000)
    n_frames = 5
    for k in range(n_frames):
        pp = vx.primary_params(350, camera.taa_jitter(k))
        sp = vx.shadow_params(stronger, frame=k)
        dp = vx.diffuse_params(sun, moon, vis, spp=1, frame=k)
        f.render(pp, sp, dp)
    f.finish()
    r.sync()
    dist.barrier()
    bad = 0
    if rank == f.root or mode == "nccl":
        cam = fc.vx_camera(W, H)
        g = r.trace_primary(cam, pp, r.alloc_gbuffer(W, H, device=True, texel=texel))
        s = r.trace_shadow(cam, g, sp, r.alloc_shadow(W, H, device=True, texel=texel))
        d = r.trace_diffuse(cam, g, dp, r.alloc_diffuse(W, H, device=True, texel=texel))
        r.sync()
        for name, ref in (("s_shadow", s["shadow"]), ("s_transversal", s["transversal"]), ("d_sh", d["sh"]), ("d_cocg", d["cocg"]),
                          ("d_luma", d["luma"]), ("d_ao_sky", d["ao_sky"])):
            got = f.plane(name)
            n = int((got.view(torch.uint8) != ref.view(torch.uint8)).sum())
            bad += n
            if n:
                print(f"rank {rank}: plane {name}: {n} differing bytes", flush=True)
    t = torch.tensor([bad], device=f"cuda:{lr}")
    dist.all_reduce(t)
    f.close()
    if rank == 0:
        print(f"p2p_check mode={mode} texel={texel} ranks={ws}: {'OK' if int(t) == 0 else 'MISMATCH'}", flush=True)
    dist.destroy_process_group()
    return 0 if int(t) == 0 else 1


if __name__ == "__main__":
    sys.exit(main())
