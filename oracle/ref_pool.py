"""Process pool that runs the reference's own shaders (oracle/_ref/libref_shaders.so, see ref_shaders.py) on every host core.

The compiled shaders keep GLSL's global variables as process globals, so one process traces one row slab at a time; a frame is cut into
row slabs and each worker runs primary -> shadow -> diffuse GI for its slab (secondary passes read the G-buffer at their own pixel
only, so slabs are independent).  Used by bench.py's `--impl reference` arm and `cpu_baseline` leg.  Test / measurement infrastructure."""
import multiprocessing as mp
import os

import numpy as np

_W = {}


def _init(world_bytes, df_bytes, width, height, cam_kw):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from voxelpathtracer_b200 import assets, camera
    from oracle import ref_shaders
    sun, moon, stronger, vis = camera.sun_moon_direction(50.0)
    _W.update(blocks=np.frombuffer(world_bytes, dtype=np.uint8), df=np.frombuffer(df_bytes, dtype=np.uint8), W=width, H=height,
              fc=camera.FpsCamera(**cam_kw), ref=ref_shaders, mats=assets.load_materials(), bn=assets.load_blue_noise(),
              sky=assets.analytic_sky(16, sun), noise=assets.load_shadow_noise(), sun=sun, moon=moon, stronger=stronger, vis=vis)
    ref_shaders.load()


def _slab(args):
    frame, rb, re = args
    import voxelpathtracer_b200 as vx
    from voxelpathtracer_b200 import camera
    w = _W
    cam = w["fc"].vx_camera(w["W"], w["H"], rb, re)
    ref = w["ref"]
    g = ref.trace_primary(w["blocks"], w["df"], cam, vx.primary_params(350, camera.taa_jitter(frame)))
    s = ref.trace_shadow(w["blocks"], w["df"], cam, g, vx.shadow_params(w["stronger"], frame=frame, soft=True), w["noise"])
    d = ref.trace_diffuse(w["blocks"], w["df"], cam, g, vx.diffuse_params(w["sun"], w["moon"], w["vis"], spp=1, frame=frame), w["mats"], w["bn"], w["sky"])
    return rb, re, float(d["luma"][rb:re].sum()), int((g["t"][rb:re] > 0).sum()), int(s["shadow"][rb:re].sum())


class ReferenceFramePool:
    def __init__(self, world_data, df, width, height, cam_kw, procs=None, slab_rows=8):
        self.procs = procs or len(os.sched_getaffinity(0))
        self.height, self.slab_rows = height, slab_rows
        ctx = mp.get_context("fork")
        self.pool = ctx.Pool(self.procs, initializer=_init, initargs=(bytes(world_data), bytes(df), width, height, cam_kw))

    def frame(self, frame_index):
        """One frame (primary + soft sun shadow + 1-spp diffuse GI) through the reference's shaders.  Returns a checksum triple."""
        tasks = [(frame_index, rb, min(rb + self.slab_rows, self.height)) for rb in range(0, self.height, self.slab_rows)]
        luma = hits = shadowed = 0
        for _, _, a, b, c in self.pool.imap_unordered(_slab, tasks, chunksize=1):
            luma += a
            hits += b
            shadowed += c
        return luma, hits, shadowed

    def close(self):
        self.pool.close()
        self.pool.join()
