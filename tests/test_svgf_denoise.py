"""SVGF diffuse denoiser (Core/Shaders/SVGF/{TemporalFilter,VarianceEstimate,SpatialFilter}.glsl, Core/Pipeline.cpp:2335-2596).

Same chain of pins as the trace passes: the reference's own shaders compiled as C++ (committed digests, tests/golden/ref_denoise_digests.json,
made by tools/make_ref_denoise_golden.py) == the oracle restatement (oracle/vxo_denoise.cpp) == the CUDA kernels' source on the host
(tests/host_shadow) == the CUDA kernels on the GPU through the C ABI.  CPU comparisons are bit for bit; the GPU comparison allows the pinned
exp / pow (double evaluation by two different libms) to differ by an ulp on isolated pixels."""
import json
import os

import numpy as np
import pytest

import voxelpathtracer_b200 as vx
from voxelpathtracer_b200 import abi, camera, denoise
from oracle import ref_shaders, vxo

import denoise_cases as dc
from host_shadow import kernels_on_host as koh

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not ref_shaders.available(), reason="oracle/_ref/libref_shaders.so not built (needs /root/reference at build time)")


@pytest.fixture(scope="module")
def ref_digests():
    with open(os.path.join(ROOT, "tests", "golden", "ref_denoise_digests.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def oracle_sequences(oracles, scene_tables):
    """name -> list of per-frame oracle results (cached: the sequences feed several tests)"""
    class Lazy(dict):
        def __missing__(self, name):
            self[name] = list(dc.run_sequence(name, dc.oracle_tracer(oracles[dc.SEQUENCES[name][0]], scene_tables), vxo, scene_tables))
            return self[name]

    return Lazy()


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


# ------------------------------------------------------------------------------------------------------ oracle vs the reference's shaders
@pytest.mark.parametrize("name", list(dc.SEQUENCES))
def test_oracle_equals_the_reference_shader_digests(oracle_sequences, ref_digests, name):
    frames = oracle_sequences[name]
    assert len(frames) == len(ref_digests[name])
    for f, (fr, want) in enumerate(zip(frames, ref_digests[name])):
        assert dc.frame_digest(fr) == want, (name, f)


@needs_ref
def test_oracle_equals_the_reference_shaders_live_on_option_and_edge_cases(oracle_sequences):
    """Switches the committed sequences leave at their defaults, and non-finite inputs: u_BeUseful = false, DO_SPATIAL = false,
    AGGRESSIVE_DISOCCLUSION_HANDLING = false, u_LargeKernel, the WiderSVGF steps, u_ResolutionScale = 1, a row slab, a zero frame count
    everywhere (0 * inf in the variance pass), NaN / inf planted in the input planes."""
    fr0, fr1 = oracle_sequences["gi_box_192x108_walk"][:2]
    cam, g, d, t = fr1["cam"], fr1["gbuf"], fr1["diffuse"], fr1["temporal"]
    fc0 = camera.FpsCamera(aspect=192 / 108, pitch_deg=-20.0)
    tp = denoise.temporal_params(fc0.view().T.reshape(16), fc0.projection().T.reshape(16), be_useful=False)
    a, b = vxo.svgf_temporal(cam, g, fr0["gbuf"], d, fr0["temporal"], tp), ref_shaders.svgf_temporal(cam, g, fr0["gbuf"], d, fr0["temporal"], tp)
    assert all(_same(a[k], b[k]) for k in a)
    for vp in (denoise.variance_params(do_spatial=False), denoise.variance_params(aggressive_disocclusion=False)):
        a, b = vxo.svgf_variance(cam, g, t, vp), ref_shaders.svgf_variance(cam, g, t, vp)
        assert all(_same(a[k], b[k]) for k in a)
    planes = {"sh": fr1["variance"]["sh"], "cocg": fr1["variance"]["cocg"], "variance": fr1["variance"]["variance"], "ao_sky": t["ao_sky"]}
    for sp in (denoise.spatial_params(8, time=1.5, do_spatial=False), denoise.spatial_params(4, time=7.75, aggressive_disocclusion=False),
               denoise.spatial_params(2, time=0.1, large_kernel=True), denoise.spatial_params(32, time=2.0), denoise.spatial_params(16, time=2.0, resolution_scale=1.0),
               denoise.spatial_params(1, time=99.0, color_phi_bias=0.05)):
        a, b = vxo.svgf_spatial(cam, g, planes, t["utility"], sp), ref_shaders.svgf_spatial(cam, g, planes, t["utility"], sp)
        assert all(_same(a[k], b[k]) for k in a), sp.step
    # a row slab writes its rows only
    cam2 = camera.FpsCamera(aspect=192 / 108, position=(192.4, 75.1, 192.3), pitch_deg=-21.0, yaw_deg=92.0).vx_camera(192, 108, 40, 71)
    seed = {k: np.full_like(v, 0.5) for k, v in fr1["spatial"][0].items()}
    a = vxo.svgf_spatial(cam2, g, planes, t["utility"], denoise.spatial_params(4, time=1.0), {k: v.copy() for k, v in seed.items()})
    b = ref_shaders.svgf_spatial(cam2, g, planes, t["utility"], denoise.spatial_params(4, time=1.0), {k: v.copy() for k, v in seed.items()})
    assert all(_same(a[k], b[k]) for k in a) and (a["sh"][:40] == 0.5).all() and (a["sh"][71:] == 0.5).all() and not (a["sh"][40:71] == 0.5).all()
    # zero frame counts and zero moments: thresh / 0 = inf, 0 * inf = NaN reach the clamps (IEEE minNum / maxNum: the non-NaN operand wins)
    t0 = {k: v.copy() for k, v in t.items()}
    t0["utility"][...] = 0.0
    t0["sh"][20:60, 30:90] = 0.0
    a, b = vxo.svgf_variance(cam, g, t0, denoise.variance_params()), ref_shaders.svgf_variance(cam, g, t0, denoise.variance_params())
    assert all(_same(a[k], b[k]) for k in a)
    # non-finite texels in the inputs propagate identically
    bad = {k: v.copy() for k, v in planes.items()}
    bad["variance"][50, 60] = np.nan
    bad["variance"][10, 10] = np.inf
    bad["sh"][70, 100, 3] = np.inf
    bad["sh"][30, 150] = np.nan
    for step in (16, 1):
        a, b = vxo.svgf_spatial(cam, g, bad, t["utility"], denoise.spatial_params(step, time=4.0)), ref_shaders.svgf_spatial(cam, g, bad, t["utility"], denoise.spatial_params(step, time=4.0))
        assert all(_same(a[k], b[k]) for k in a), step


# ------------------------------------------------------------------------------------------------------ known answers / properties
def test_still_camera_accumulates_and_the_filter_removes_noise(oracle_sequences):
    frames = oracle_sequences["city_160x90_still"]
    hit = frames[0]["gbuf"]["t"] > 0
    inner = np.zeros_like(hit)
    inner[4:-4, 4:-4] = True
    # accumulated frame count (o_Utility.x): 1, 2, 3 on every pixel the reprojection accepts — all of them for a camera that does not move
    for f, fr in enumerate(frames):
        spp = fr["temporal"]["utility"][..., 0]
        sel = spp[hit & inner]                   # silhouette pixels reject their history, and their neighbours average it in
        assert np.isclose(sel, f + 1.0, atol=1e-4).mean() > 0.7 and abs(float(np.median(sel)) - (f + 1.0)) < 1e-4 and sel.max() <= f + 1.0 + 1e-4, f
    # the blend factor is 1 / accumulated frames: frame 1's temporal SH = (history + this frame's pre-filtered SH) / 2, the history being a
    # convex combination of the previous temporal SH at the pixel and its four neighbours: 2 * out - current lies in their envelope
    prev, cur, got = frames[0]["temporal"]["sh"], frames[1]["initial"]["sh"], frames[1]["temporal"]["sh"]
    ok = np.isclose(frames[1]["temporal"]["utility"][..., 0], 2.0, atol=1e-4) & inner & (frames[1]["gbuf"]["t"] > 1.0)   # (the lowest rows look
    taps = np.stack([prev, np.roll(prev, 1, 0), np.roll(prev, -1, 0), np.roll(prev, 1, 1), np.roll(prev, -1, 1)])
    hist = 2.0 * got - cur                               # at the block the camera stands in, t = 1e-4: reprojection is ill-conditioned there)
    assert (hist[ok] >= taps.min(0)[ok] - 1e-4).all() and (hist[ok] <= taps.max(0)[ok] + 1e-4).all()
    # each a-trous pass leaves the mean radiance alone and lowers the pixel-to-pixel roughness of the luminance band
    def rough(sh):
        y = sh[..., 3]
        return float(np.abs(np.diff(y, axis=1))[hit[:, 1:] & hit[:, :-1]].mean())
    fr = frames[2]
    r_raw, r_t = rough(fr["diffuse"]["sh"]), rough(fr["temporal"]["sh"])
    r_s = [rough(s["sh"]) for s in fr["spatial"]]
    assert r_t < r_raw and r_s[-1] < 0.5 * r_raw and r_s[0] < r_t
    m_raw = float(np.mean([f_["diffuse"]["sh"][hit][:, 3].mean() for f_ in frames]))      # 1-spp frames: compare with the three-frame mean
    m_out = float(fr["spatial"][-1]["sh"][hit][:, 3].mean())
    assert 0.4 * m_raw < m_out < 1.5 * m_raw             # edge-stopping weights cut the heavy tail of 1-spp fireflies: not mean-preserving


def test_constant_planes_are_a_fixed_point():
    """Uniform inputs over a flat G-buffer: every weighted mean returns the constant, the variance of a constant signal is zero."""
    W, H = 64, 36
    cam = camera.FpsCamera(aspect=W / H).vx_camera(W, H)
    g = {"t": np.full((H, W), 12.5, np.float32), "normal_id": np.full((H, W), 2, np.uint8), "block_id": np.full((H, W), 3, np.uint8)}
    sh = np.tile(np.array([0.1, -0.2, 0.05, 0.3], np.float32), (H, W, 1))
    cocg = np.tile(np.array([0.02, -0.01], np.float32), (H, W, 1))
    ao = np.tile(np.array([0.7, 0.4], np.float32), (H, W, 1))
    y = np.float32(max(0.0, 3.544905 * 0.3))
    util = np.tile(np.array([20.0, y * y, y], np.float32), (H, W, 1))
    v = vxo.svgf_variance(cam, g, {"sh": sh, "cocg": cocg, "utility": util}, denoise.variance_params())
    assert np.allclose(v["sh"], sh, atol=1e-6) and np.allclose(v["variance"], 0.0, atol=1e-5)
    planes = {"sh": sh, "cocg": cocg, "variance": np.full((H, W), 0.02, np.float32), "ao_sky": ao}
    for step in denoise.ATROUS_STEPS:
        s = vxo.svgf_spatial(cam, g, planes, util, denoise.spatial_params(step, time=5.0))
        assert np.allclose(s["sh"], sh, atol=1e-6) and np.allclose(s["cocg"], cocg, atol=1e-6) and np.allclose(s["ao_sky"], ao, atol=1e-6)
        assert (s["variance"] <= 0.02 + 1e-7).all() and (s["variance"] > 0).all()    # sum w^2 v / (sum w)^2 < v
        planes = s


# ------------------------------------------------------------------------------------------------------ sun-shadow filters
@pytest.fixture(scope="module")
def oracle_shadow_sequences(oracles, scene_tables):
    class Lazy(dict):
        def __missing__(self, name):
            self[name] = list(dc.run_shadow_sequence(name, dc.oracle_shadow_tracer(oracles[dc.SEQUENCES[name][0]], scene_tables), vxo))
            return self[name]

    return Lazy()


@pytest.mark.parametrize("name", list(dc.SEQUENCES))
def test_shadow_filters_oracle_equals_the_reference_shader_digests(oracle_shadow_sequences, ref_digests, name):
    frames, want = oracle_shadow_sequences[name], ref_digests["shadow:" + name]
    assert len(frames) == len(want)
    for f, fr in enumerate(frames):
        assert dc.shadow_frame_digest(fr) == want[f], (name, f)


@needs_ref
def test_shadow_filters_oracle_equals_the_reference_shaders_live(oracle_shadow_sequences):
    """Other filter scales, a row slab, long frame histories (the luminance weight switches on above 7.5 frames) and non-finite inputs."""
    fr = oracle_shadow_sequences["city_160x90_still"][2]
    cam, g, s, t = fr["cam"], fr["gbuf"], fr["shadow"], fr["temporal"]
    for scale in (0.25, 3.0):
        assert _same(vxo.shadow_filter(cam, g, t, s["transversal"], denoise.shadow_filter_params(scale)),
                     ref_shaders.shadow_filter(cam, g, t, s["transversal"], denoise.shadow_filter_params(scale)))
    long_t = {"shadow": t["shadow"], "frames": np.full_like(t["frames"], 9.0)}
    assert _same(vxo.shadow_filter(cam, g, long_t, s["transversal"], denoise.shadow_filter_params(1.0)),
                 ref_shaders.shadow_filter(cam, g, long_t, s["transversal"], denoise.shadow_filter_params(1.0)))
    a = vxo.shadow_temporal(cam, g, fr["prev_gbuf"], s, long_t, fr["params"])
    b = ref_shaders.shadow_temporal(cam, g, fr["prev_gbuf"], s, long_t, fr["params"])
    assert _same(a["shadow"], b["shadow"]) and _same(a["frames"], b["frames"]) and a["frames"].max() > 9.0
    bad = {"shadow": t["shadow"].copy(), "frames": t["frames"].copy()}
    bad["shadow"][40, 50] = np.nan
    bad["shadow"][10, 100] = np.inf
    bad["frames"][60, 20] = np.nan
    a, b = vxo.shadow_temporal(cam, g, fr["prev_gbuf"], s, bad, fr["params"]), ref_shaders.shadow_temporal(cam, g, fr["prev_gbuf"], s, bad, fr["params"])
    assert _same(a["shadow"], b["shadow"]) and _same(a["frames"], b["frames"])
    assert _same(vxo.shadow_filter(cam, g, bad, s["transversal"], denoise.shadow_filter_params(1.0)),
                 ref_shaders.shadow_filter(cam, g, bad, s["transversal"], denoise.shadow_filter_params(1.0)))
    slab = camera.FpsCamera(aspect=160 / 90, position=(100.0, 60.0, 100.0), pitch_deg=-10.0, yaw_deg=45.0).vx_camera(160, 90, 30, 61)
    a = vxo.shadow_filter(slab, g, t, s["transversal"], denoise.shadow_filter_params(1.0), np.full((90, 160), 0.25, np.float32))
    b = ref_shaders.shadow_filter(slab, g, t, s["transversal"], denoise.shadow_filter_params(1.0), np.full((90, 160), 0.25, np.float32))
    assert _same(a, b) and (a[:30] == 0.25).all() and (a[61:] == 0.25).all()


def test_shadow_filters_properties(oracle_shadow_sequences):
    frames = oracle_shadow_sequences["city_160x90_still"]
    hit = frames[0]["gbuf"]["t"] > 1.0
    # a still camera accumulates: the frame counter grows by one per frame on most surface pixels, and never on the sky
    for f, fr in enumerate(frames):
        fc = fr["temporal"]["frames"]
        assert abs(float(np.median(fc[hit])) - (f + 1.0)) < 1e-3, (f, float(np.median(fc[hit])))
        assert (fc[frames[0]["gbuf"]["t"] < 0] == 0).all()
        assert (fr["temporal"]["shadow"] >= 0).all() and (fr["temporal"]["shadow"] <= 1).all()
        assert (fr["filtered"] >= -1e-6).all() and (fr["filtered"] <= 1 + 1e-6).all()
    # the filters do not leak light into a region that is shadowed throughout: where the 7x7 neighbourhood of the raw plane is all
    # shadow (1), three accumulated frames and the spatial filter leave the pixel shadowed
    raw = np.minimum.reduce([fr["shadow"]["shadow"] for fr in frames]).astype(np.float32)
    k = 3
    pad = np.pad(raw, k, mode="edge")
    win = np.stack([pad[dy:dy + raw.shape[0], dx:dx + raw.shape[1]] for dy in range(2 * k + 1) for dx in range(2 * k + 1)])
    dark = (win.min(0) == 1) & hit
    assert dark.sum() > 100 and (frames[2]["filtered"][dark] > 0.9).all()


@pytest.mark.skipif(not koh.available(), reason="CUDA toolkit headers not present")
@pytest.mark.parametrize("name", ["city_160x90_still", "plains_133x75_turn"])
def test_shadow_filter_kernel_source_on_host_equals_the_oracle(oracles, oracle_shadow_sequences, name):
    k = koh.HostKernels(oracles[dc.SEQUENCES[name][0]], 1)
    for f, fr in enumerate(oracle_shadow_sequences[name]):
        t = k.shadow_temporal(fr["cam"], fr["gbuf"], fr["prev_gbuf"], fr["shadow"], fr["prev_temporal"], fr["params"])
        assert _same(t["shadow"], fr["temporal"]["shadow"]) and _same(t["frames"], fr["temporal"]["frames"]), f
        assert _same(k.shadow_filter(fr["cam"], fr["gbuf"], fr["temporal"], fr["shadow"]["transversal"], denoise.shadow_filter_params(1.0)), fr["filtered"]), f
    k.close()


# ------------------------------------------------------------------------------------------------------ GPU, through the C ABI
def _close(got, want, what):
    """bit-equal but for isolated pixels: CUDA's and glibc's double exp / pow differ in the last place about once in 10^8 calls; such a
    pixel is off by an ulp — or, where a clamp follows a 0 * inf or a sign test, by the width of the clamp.  Three outliers per plane are
    tolerated outright, every other difference must be rounding-sized."""
    got, want = np.asarray(got), np.asarray(want)
    assert (np.isnan(got) != np.isnan(want)).sum() <= 3, what
    diff = ~((got == want) | (np.isnan(got) & np.isnan(want)))
    assert diff.mean() <= 2e-4, (what, float(diff.mean()))
    fin = np.isfinite(want) & np.isfinite(got)
    err = np.sort(np.abs(got[fin].astype(np.float64) - want[fin]).reshape(-1))
    assert err.size <= 3 or float(err[-4]) <= 2e-5, (what, err[-4:])


def _close_chain(got, want, what):
    """For results of several chained passes: a one-ulp difference in a pinned exp / pow (CUDA's libm against glibc's, about one call in
    10^8) is carried through the following passes' stencils and can flip a clamped variance, so isolated pixels may differ by more than
    an ulp; anything systematic (a wrong history plane, a wrong pass order) differs on most of the frame."""
    got, want = np.asarray(got), np.asarray(want)
    diff = ~((got == want) | (np.isnan(got) & np.isnan(want)))
    assert diff.mean() <= 1e-3, (what, float(diff.mean()))


class _RendererPasses:
    """adapts a Renderer to the call shapes of oracle.vxo.svgf_* (allocates the output planes)"""

    def __init__(self, r, device=False):
        self.r, self.device = r, device

    def _out(self, cam, names):
        return self.r.alloc_denoise(cam.width, cam.height, names, device=self.device)

    def _in(self, planes):
        if not self.device:
            return planes
        import torch
        out = {k: (v if isinstance(v, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(v)).cuda()) for k, v in planes.items() if v is not None}
        torch.cuda.synchronize()   # the uploads ran on torch's stream, the library launches on its own
        return out

    def _host(self, planes):
        if self.device:
            self.r.sync()              # device planes are complete after vxpt_sync()
        return {k: (v.cpu().numpy() if self.device else v) for k, v in planes.items()}

    def svgf_initial(self, cam, g, d):
        return self._host(self.r.svgf_initial(cam, self._in(g), self._in(d), self._out(cam, ("sh", "cocg", "luma", "ao_sky"))))

    def svgf_temporal(self, cam, g, pg, d, pt, params):
        return self._host(self.r.svgf_temporal(cam, self._in(g), self._in(pg), self._in(d), self._in(pt), params, self._out(cam, ("sh", "cocg", "utility", "ao_sky"))))

    def svgf_variance(self, cam, g, t, params):
        return self._host(self.r.svgf_variance(cam, self._in(g), self._in(t), params, self._out(cam, ("sh", "cocg", "variance"))))

    def svgf_spatial(self, cam, g, planes, util, params):
        util = self._in({"u": util})["u"]
        return self._host(self.r.svgf_spatial(cam, self._in(g), self._in(planes), util, params, self._out(cam, ("sh", "cocg", "variance", "ao_sky"))))


@pytest.mark.gpu
@pytest.mark.parametrize("name,device", [("gi_box_192x108_walk", False), ("plains_133x75_turn", True), ("city_160x90_still", True)])
def test_gpu_denoiser_equals_the_oracle(renderer, oracle_sequences, scene_tables, name, device):
    """Every pass of every frame, fed with the oracle's inputs of that pass, so a one-ulp difference cannot compound down the chain."""
    want = oracle_sequences[name]
    p = _RendererPasses(renderer, device)
    prev_g, prev_t = None, dc.zero_temporal(*dc.SEQUENCES[name][1:3])
    prev_fc = None
    for f, (fr, kw) in enumerate(zip(want, dc.SEQUENCES[name][3])):
        W, H = dc.SEQUENCES[name][1:3]
        fc = camera.FpsCamera(aspect=W / H, **kw)
        pfc = prev_fc or fc
        tp = denoise.temporal_params(pfc.view().T.reshape(16), pfc.projection().T.reshape(16))
        g = {k: fr["gbuf"][k] for k in ("t", "normal_id", "block_id")}
        pre = p.svgf_initial(fr["cam"], g, fr["diffuse"])
        for k in fr["initial"]:
            _close(pre[k], fr["initial"][k], (name, f, "initial", k))
        t = p.svgf_temporal(fr["cam"], g, prev_g or g, fr["initial"], prev_t, tp)
        for k in fr["temporal"]:
            _close(t[k], fr["temporal"][k], (name, f, "temporal", k))
        v = p.svgf_variance(fr["cam"], g, fr["temporal"], denoise.variance_params())
        for k in fr["variance"]:
            _close(v[k], fr["variance"][k], (name, f, "variance", k))
        cur = {"sh": fr["variance"]["sh"], "cocg": fr["variance"]["cocg"], "variance": fr["variance"]["variance"], "ao_sky": fr["temporal"]["ao_sky"]}
        for n, step in enumerate(denoise.ATROUS_STEPS):
            s = p.svgf_spatial(fr["cam"], g, cur, fr["temporal"]["utility"], denoise.spatial_params(step, time=dc.TIME0 + f / 60.0))
            for k in fr["spatial"][n]:
                _close(s[k], fr["spatial"][n][k], (name, f, step, k))
            cur = fr["spatial"][n]
        prev_g, prev_t, prev_fc = g, fr["temporal"], fc
    assert renderer.launch_count() > 0


@pytest.mark.gpu
def test_gpu_denoiser_whole_chain_on_the_traced_frame(renderer, worlds, oracles, scene_tables):
    """End to end on the device: trace (primary + GI) and denoise two frames with device-resident planes through Renderer.svgf_denoise;
    the result stays within the radiance tolerance of the oracle's chain (north_star: 1e-3 mean absolute error)."""
    name = "gi_box_192x108_walk"
    _, W, H, cams = dc.SEQUENCES[name]
    renderer.upload_world(worlds["gi_box"])
    renderer.build_distance_field()
    sun, moon, vis = scene_tables["sun"], scene_tables["moon"], scene_tables["sun_visibility"]
    want = list(dc.run_sequence(name, dc.oracle_tracer(oracles["gi_box"], scene_tables), vxo, scene_tables))[:2]
    prev_g, prev_fc = None, None
    prev_t = renderer.alloc_denoise(W, H, ("sh", "cocg", "utility", "ao_sky"), device=True)
    for v in prev_t.values():
        v.zero_()
    import torch
    torch.cuda.synchronize()       # zero_() ran on torch's stream, the library launches on its own
    for f, kw in enumerate(cams[:2]):
        fc = camera.FpsCamera(aspect=W / H, **kw)
        cam = fc.vx_camera(W, H)
        g = renderer.trace_primary(cam, vx.primary_params(350), renderer.alloc_gbuffer(W, H, device=True))
        d = renderer.trace_diffuse(cam, g, vx.diffuse_params(sun, moon, vis, spp=1, frame=f), renderer.alloc_diffuse(W, H, device=True))
        pfc = prev_fc or fc
        tp = denoise.temporal_params(pfc.view().T.reshape(16), pfc.projection().T.reshape(16))
        out, temporal = renderer.svgf_denoise(cam, g, prev_g or g, d, prev_t, tp, time=dc.TIME0 + f / 60.0, device=True)
        renderer.sync()
        ref = want[f]["spatial"][-1]
        for k in ("sh", "cocg", "ao_sky"):
            assert float(np.mean(np.abs(out[k].cpu().numpy().astype(np.float64) - ref[k]))) <= 1e-3, (f, k)
        prev_g, prev_t, prev_fc = g, temporal, fc


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["gi_box_192x108_walk", "city_160x90_still"])
def test_gpu_svgf_frame_keeps_the_history_on_the_device(renderer, oracle_sequences, name):
    """vxpt_svgf_frame == the separate passes chained by hand: per frame only the G-buffer and the GI planes go in; the pre-pass, temporal,
    variance and ping-pong planes and the previous frame stay in the handle.  Host planes; also a reset in mid-sequence and a resolution change."""
    _, W, H, cams = dc.SEQUENCES[name]
    want = oracle_sequences[name]
    for f, (fr, kw) in enumerate(zip(want, cams)):
        fc = camera.FpsCamera(aspect=W / H, **kw)
        view, proj = fc.view().T.reshape(16), fc.projection().T.reshape(16)       # the matrices the oracle sequence was run with
        out = renderer.svgf_frame(fr["cam"], fr["gbuf"], fr["diffuse"], denoise.frame_params(view, proj, time=dc.TIME0 + f / 60.0, reset_history=(f == 0)),
                                  renderer.alloc_denoise(W, H, ("sh", "cocg", "variance", "ao_sky")))
        for k in out:
            _close_chain(out[k], fr["spatial"][-1][k], (name, f, k))
    # reset_history on a later frame = that frame denoised as the first of a sequence
    fr, fc = want[1], camera.FpsCamera(aspect=W / H, **cams[1])
    view, proj = fc.view().T.reshape(16), fc.projection().T.reshape(16)
    a = renderer.svgf_frame(fr["cam"], fr["gbuf"], fr["diffuse"], denoise.frame_params(view, proj, time=1.0, reset_history=True),
                            renderer.alloc_denoise(W, H, ("sh", "cocg", "variance", "ao_sky")))
    # ... and so does a frame of another size in between (the history is per resolution)
    small = oracle_sequences["plains_133x75_turn"][0]
    fs = camera.FpsCamera(aspect=133 / 75, **dc.SEQUENCES["plains_133x75_turn"][3][0])
    renderer.svgf_frame(small["cam"], small["gbuf"], small["diffuse"], denoise.frame_params(fs.view().T.reshape(16), fs.projection().T.reshape(16), time=dc.TIME0),
                        renderer.alloc_denoise(133, 75, ("sh", "cocg", "variance", "ao_sky")))
    b = renderer.svgf_frame(fr["cam"], fr["gbuf"], fr["diffuse"], denoise.frame_params(view, proj, time=1.0),
                            renderer.alloc_denoise(W, H, ("sh", "cocg", "variance", "ao_sky")))
    for k in a:
        assert _same(a[k], b[k]), k             # the same kernels on the same inputs: bit for bit
    with pytest.raises(abi.VxptError) as e:     # whole frames only
        slab = camera.FpsCamera(aspect=W / H, **cams[0]).vx_camera(W, H, 0, H // 2)
        renderer.svgf_frame(slab, fr["gbuf"], fr["diffuse"], denoise.frame_params(view, proj), renderer.alloc_denoise(W, H, ("sh", "cocg", "variance", "ao_sky")))
    assert e.value.code == abi.E_INVALID


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["city_160x90_still", "gi_box_192x108_walk"])
def test_gpu_shadow_filter_frame_keeps_the_history_on_the_device(renderer, oracle_shadow_sequences, name):
    """vxpt_shadow_filter_frame == vxpt_shadow_temporal + vxpt_shadow_filter chained by hand over a sequence (host planes in and out)."""
    _, W, H, cams = dc.SEQUENCES[name]
    for f, (fr, kw) in enumerate(zip(oracle_shadow_sequences[name], cams)):
        fc = camera.FpsCamera(aspect=W / H, **kw)
        prm = denoise.shadow_frame_params(fc.view().T.reshape(16), fc.projection().T.reshape(16), reset_history=(f == 0))
        out = renderer.shadow_filter_frame(fr["cam"], fr["gbuf"], fr["shadow"], prm, np.zeros((H, W), np.float32))
        _close_chain(out, fr["filtered"], (name, f))
    fr, fc = oracle_shadow_sequences[name][0], camera.FpsCamera(aspect=W / H, **cams[0])
    prm = denoise.shadow_frame_params(fc.view().T.reshape(16), fc.projection().T.reshape(16), reset_history=True, spatial=False)
    out = renderer.shadow_filter_frame(fr["cam"], fr["gbuf"], fr["shadow"], prm, np.zeros((H, W), np.float32))
    _close(out, fr["temporal"]["shadow"], (name, "temporal only"))


class _DevicePlanes:
    """Device-resident planes without torch: vxpt_shared_alloc for the memory, vxpt_copy_async for the transfers.  The passes get raw
    device addresses, so the ABI takes its zero-copy path (no staging)."""

    def __init__(self, r):
        self.r, self.ptrs = r, []

    def up(self, a):
        a = np.ascontiguousarray(a)
        ptr, _ = self.r.shared_alloc(max(a.nbytes, 256))
        self.ptrs.append(ptr)
        self.r.copy_async(ptr, a.ctypes.data, a.nbytes)
        self.keep = getattr(self, "keep", []) + [a]
        return ptr

    def new(self, shape, dtype=np.float32):
        ptr, _ = self.r.shared_alloc(max(int(np.prod(shape)) * np.dtype(dtype).itemsize, 256))
        self.ptrs.append(ptr)
        return ptr

    def down(self, ptr, shape, dtype=np.float32):
        out = np.empty(shape, dtype)
        self.r.copy_async(out.ctypes.data, ptr, out.nbytes)
        self.r.sync()
        return out

    def close(self):
        self.r.sync()
        for p in self.ptrs:
            self.r.shared_close(p)


@pytest.mark.gpu
def test_gpu_resident_planes_take_the_zero_copy_path(renderer, worlds, oracle_sequences, oracle_shadow_sequences, scene_tables):
    """Raw device addresses in, raw device addresses out: vxpt_svgf_frame, the shadow filters and the material pass on planes that live in
    device memory (allocated through the ABI itself, so the test needs no torch and also runs against the emulated ABI)."""
    r = renderer
    dev = _DevicePlanes(r)
    try:
        name = "gi_box_192x108_walk"
        _, W, H, cams = dc.SEQUENCES[name]
        shapes = denoise.plane_shapes(W, H)
        for f, (fr, kw) in enumerate(zip(oracle_sequences[name][:2], cams)):
            fc = camera.FpsCamera(aspect=W / H, **kw)
            g = {k: dev.up(fr["gbuf"][k]) for k in ("t", "normal_id", "block_id")}
            d = {k: dev.up(fr["diffuse"][k]) for k in ("sh", "cocg", "luma", "ao_sky")}
            out = {k: dev.new(shapes[k]) for k in ("sh", "cocg", "variance", "ao_sky")}
            r.svgf_frame(fr["cam"], g, d, denoise.frame_params(fc.view().T.reshape(16), fc.projection().T.reshape(16), time=dc.TIME0 + f / 60.0,
                                                               reset_history=(f == 0)), out)
            for k in out:
                _close_chain(dev.down(out[k], shapes[k]), fr["spatial"][-1][k], (f, k))
        fr = oracle_shadow_sequences["city_160x90_still"][1]
        W, H = fr["cam"].width, fr["cam"].height
        g = {"t": dev.up(fr["gbuf"]["t"]), "normal_id": dev.up(fr["gbuf"]["normal_id"])}
        pg = {"t": dev.up(fr["prev_gbuf"]["t"])}
        s = {"shadow": dev.up(fr["shadow"]["shadow"]), "transversal": dev.up(fr["shadow"]["transversal"])}
        pt = {k: dev.up(fr["prev_temporal"][k]) for k in ("shadow", "frames")}
        t = {"shadow": dev.new((H, W)), "frames": dev.new((H, W))}
        r.shadow_temporal(fr["cam"], g, pg, s, pt, fr["params"], t)
        filt = r.shadow_filter(fr["cam"], g, t, s["transversal"], denoise.shadow_filter_params(1.0), dev.new((H, W)))
        _close(dev.down(t["shadow"], (H, W)), fr["temporal"]["shadow"], "temporal shadow")
        _close(dev.down(t["frames"], (H, W)), fr["temporal"]["frames"], "temporal frames")
        _close(dev.down(filt, (H, W)), fr["filtered"], "filtered")
    finally:
        dev.close()


@pytest.mark.gpu
def test_gpu_denoiser_argument_checks(renderer, oracle_sequences):
    fr = oracle_sequences["plains_133x75_turn"][0]
    cam, g, t = fr["cam"], fr["gbuf"], fr["temporal"]
    W, H = cam.width, cam.height
    out = renderer.alloc_denoise(W, H, ("sh", "cocg", "variance"))
    with pytest.raises(abi.VxptError) as e:          # a required plane is missing
        renderer.svgf_variance(cam, {"t": g["t"]}, t, denoise.variance_params(), out)
    assert e.value.code == abi.E_INVALID
    band = camera.FpsCamera(aspect=W / H).vx_camera(W, 72, interleave_n=2, interleave_rank=0, band_rows=4)
    with pytest.raises(abi.VxptError) as e:          # stencils cannot run on interleaved bands
        renderer.svgf_variance(band, g, t, denoise.variance_params(), out)
    assert e.value.code == abi.E_UNSUPPORTED
    renderer.set_option(abi.OPT_TEXEL_FORMAT, 1)
    try:
        with pytest.raises(abi.VxptError) as e:
            renderer.svgf_variance(cam, g, t, denoise.variance_params(), out)
        assert e.value.code == abi.E_UNSUPPORTED
    finally:
        renderer.set_option(abi.OPT_TEXEL_FORMAT, 0)
    with pytest.raises(abi.VxptError) as e:
        planes = {"sh": fr["variance"]["sh"], "cocg": fr["variance"]["cocg"], "variance": fr["variance"]["variance"], "ao_sky": t["ao_sky"]}
        renderer.svgf_spatial(cam, g, planes, t["utility"], denoise.spatial_params(0), renderer.alloc_denoise(W, H, ("sh", "cocg", "variance", "ao_sky")))
    assert e.value.code == abi.E_INVALID


@pytest.mark.gpu
@pytest.mark.parametrize("name,device", [("city_160x90_still", False), ("gi_box_192x108_walk", True)])
def test_gpu_shadow_filters_equal_the_oracle(renderer, oracle_shadow_sequences, name, device):
    r = renderer
    for f, fr in enumerate(oracle_shadow_sequences[name]):
        cam, W, H = fr["cam"], fr["cam"].width, fr["cam"].height
        g = {k: fr["gbuf"][k] for k in ("t", "normal_id")}
        pg = {"t": fr["prev_gbuf"]["t"]}
        ins = [g, pg, fr["shadow"], fr["prev_temporal"], fr["temporal"]]
        if device:
            import torch
            ins = [{k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in d.items()} for d in ins]
            torch.cuda.synchronize()
        g_, pg_, s_, pt_, t_ = ins
        out = r.alloc_denoise(W, H, ("shadow", "frames"), device=device)
        r.shadow_temporal(cam, g_, pg_, s_, pt_, fr["params"], out)
        filt = r.shadow_filter(cam, g_, t_, s_["transversal"], denoise.shadow_filter_params(1.0), r.alloc((H, W), np.float32, device=device))
        if device:
            r.sync()
            out, filt = {k: v.cpu().numpy() for k, v in out.items()}, filt.cpu().numpy()
        _close(out["shadow"], fr["temporal"]["shadow"], (name, f, "shadow"))
        _close(out["frames"], fr["temporal"]["frames"], (name, f, "frames"))
        _close(filt, fr["filtered"], (name, f, "filtered"))


@pytest.mark.gpu
def test_headless_cpp_shadow_filters_equal_python_driver():
    """The C++ host mirror drives vxpt_shadow_temporal / vxpt_shadow_filter on its last traced frame; the Python driver repeats the calls."""
    import subprocess
    from test_host_cpp import build_headless, fnv1a
    from voxelpathtracer_b200 import world
    exe = build_headless()
    W, H = 160, 90
    out = subprocess.run([exe, str(W), str(H)], capture_output=True, text=True, check=True, env=dict(os.environ, VXPT_HEADLESS_FILTERS="1")).stdout.split("\n")
    line = [ln.split() for ln in out if ln.startswith("shadow_filters")][0]
    assert line[1] != "failed", out
    got = {line[i]: int(line[i + 1], 16) for i in range(1, len(line), 2)}
    r = vx.Renderer(0)
    try:
        r.upload_world(world.generate_superflat())
        r.build_distance_field()
        fc = camera.FpsCamera(yaw_deg=90.0, pitch_deg=-20.0, aspect=W / H)
        cam = fc.vx_camera(W, H)
        sun = np.array([-0.66896474, 0.46841538, 0.57735026], np.float32)
        g = r.trace_primary(cam, vx.primary_params(350, camera.taa_jitter(2)), r.alloc_gbuffer(W, H))
        s = r.trace_shadow(cam, g, vx.shadow_params(sun, frame=2, soft=False), r.alloc_shadow(W, H))
        zero = {"shadow": np.zeros((H, W), np.float32), "frames": np.zeros((H, W), np.float32)}
        t = r.shadow_temporal(cam, g, g, s, zero, denoise.shadow_temporal_params(*fc.view_projection_f32()), r.alloc_denoise(W, H, ("shadow", "frames")))
        f = r.shadow_filter(cam, g, t, s["transversal"], denoise.shadow_filter_params(1.0), np.zeros((H, W), np.float32))
        assert fnv1a(t["shadow"]) == got["temporal"] and fnv1a(t["frames"]) == got["frames"] and fnv1a(f) == got["filtered"]
    finally:
        r.close()
