"""Frame sequences of the SVGF denoiser tests, shared by tools/make_ref_denoise_golden.py (runs the reference's own
shaders and commits digests) and tests/test_svgf_denoise.py.  A sequence is a camera path; every frame runs primary + 1-spp GI, then
the pre-temporal 3x3 pass -> temporal -> variance -> five a-trous passes, and hands its temporal planes and G-buffer to the next frame (Core/Pipeline.cpp:2335-2596)."""
import hashlib

import numpy as np

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import camera, denoise

# name -> (world, width, height, [camera keywords per frame])
SEQUENCES = {
    "gi_box_192x108_walk": ("gi_box", 192, 108, [dict(pitch_deg=-20.0), dict(position=(192.4, 75.1, 192.3), pitch_deg=-21.0, yaw_deg=92.0),
                                                dict(position=(192.9, 75.1, 192.7), pitch_deg=-21.5, yaw_deg=95.0)]),
    "city_160x90_still": ("city", 160, 90, [dict(position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0)] * 3),
    "plains_133x75_turn": ("plains", 133, 75, [dict(pitch_deg=5.0, yaw_deg=80.0), dict(pitch_deg=5.0, yaw_deg=100.0)]),   # odd size, half sky
}
TIME0 = 3.25   # u_Time of the first frame; 1/60 s per frame


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def zero_temporal(W, H):
    shapes = denoise.plane_shapes(W, H)
    return {k: np.zeros(shapes[k], np.float32) for k in ("sh", "cocg", "utility", "ao_sky")}


def run_sequence(name, trace, passes, scene_tables):
    """trace(cam, frame) -> (gbuf, diffuse planes); passes: object with svgf_temporal / svgf_variance / svgf_spatial (oracle.vxo, the
    reference shaders, the host-compiled kernels or a Renderer adapter).  Yields per frame a dict of every pass's output planes."""
    _, W, H, cams = SEQUENCES[name]
    prev_g, prev_t, prev_fc = None, zero_temporal(W, H), None
    for f, kw in enumerate(cams):
        fc = camera.FpsCamera(aspect=W / H, **kw)
        cam = fc.vx_camera(W, H)
        g, d = trace(cam, f)
        pfc = prev_fc or fc                     # frame 0: PreviousView = CurrentView (Pipeline.cpp initialises both from the camera)
        tp = denoise.temporal_params(pfc.view().T.reshape(16), pfc.projection().T.reshape(16))
        pre = passes.svgf_initial(cam, g, d)                # PreTemporalSpatialPass (default on): the temporal pass reads its outputs
        t = passes.svgf_temporal(cam, g, prev_g or g, pre, prev_t, tp)
        v = passes.svgf_variance(cam, g, t, denoise.variance_params())
        cur = {"sh": v["sh"], "cocg": v["cocg"], "variance": v["variance"], "ao_sky": t["ao_sky"]}
        spatial = []
        for step in denoise.ATROUS_STEPS:
            cur = passes.svgf_spatial(cam, g, cur, t["utility"], denoise.spatial_params(step, time=TIME0 + f / 60.0))
            spatial.append(cur)
        yield {"cam": cam, "gbuf": g, "diffuse": d, "initial": pre, "temporal": t, "variance": v, "spatial": spatial}
        prev_g, prev_t, prev_fc = g, t, fc


def frame_digest(fr):
    out = {"initial": {k: sha(a) for k, a in fr["initial"].items()}, "temporal": {k: sha(a) for k, a in fr["temporal"].items()},
           "variance": {k: sha(a) for k, a in fr["variance"].items()}}
    out["spatial"] = [{k: sha(a) for k, a in s.items()} for s in fr["spatial"]]
    return out


def oracle_tracer(o, scene_tables):
    sun, moon, vis = scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"]

    def trace(cam, f):
        g, _ = o.trace_primary(cam, vx.primary_params(350))
        d, _ = o.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=f))
        return g, d

    return trace


def run_shadow_sequence(name, trace, passes, filter_scale=1.0):
    """trace(cam, frame) -> (gbuf, shadow planes); passes: object with shadow_temporal / shadow_filter.  Yields per frame the temporal
    planes and the spatially filtered plane (Core/Pipeline.cpp:2854-2944)."""
    _, W, H, cams = SEQUENCES[name]
    prev_g, prev_fc = None, None
    prev_t = {"shadow": np.zeros((H, W), np.float32), "frames": np.zeros((H, W), np.float32)}
    for f, kw in enumerate(cams):
        fc = camera.FpsCamera(aspect=W / H, **kw)
        cam = fc.vx_camera(W, H)
        g, s = trace(cam, f)
        pfc = prev_fc or fc
        tp = denoise.shadow_temporal_params(pfc.view().T.reshape(16), pfc.projection().T.reshape(16))
        t = passes.shadow_temporal(cam, g, prev_g or g, s, prev_t, tp)
        filtered = passes.shadow_filter(cam, g, t, s["transversal"], denoise.shadow_filter_params(filter_scale))
        yield {"cam": cam, "gbuf": g, "shadow": s, "temporal": t, "filtered": filtered, "prev_gbuf": prev_g or g, "prev_temporal": prev_t, "params": tp}
        prev_g, prev_t, prev_fc = g, t, fc


def shadow_frame_digest(fr):
    return {"shadow": sha(fr["temporal"]["shadow"]), "frames": sha(fr["temporal"]["frames"]), "filtered": sha(fr["filtered"])}


def oracle_shadow_tracer(o, scene_tables):
    def trace(cam, f):
        g, _ = o.trace_primary(cam, vx.primary_params(350))
        s, _ = o.trace_shadow(cam, g, vx.shadow_params(scene_tables["stronger"], frame=f, soft=True))
        return g, s

    return trace
