"""Host-side description of the SVGF denoiser calls (include/vxpt.h: vxpt_svgf_temporal / _variance / _spatial): builds the ABI structs
from dicts of planes.  The same structs drive the library (Renderer.svgf_*), the CPU oracle and the host-compiled kernels in tests, so
`address` is injected: it maps a numpy array or torch tensor to its raw address."""
import numpy as np

from .abi import (VxGBuffer, VxShadowFilterIn, VxShadowFrameParams, VxShadowFilterParams, VxShadowTemporalIn, VxShadowTemporalOut, VxShadowTemporalParams, VxSvgfFrameParams, VxSvgfInitialIn, VxSvgfInitialOut, VxSvgfSpatialIn, VxSvgfSpatialOut, VxSvgfSpatialParams, VxSvgfTemporalIn, VxSvgfTemporalOut, VxSvgfTemporalParams,
                  VxSvgfVarianceIn, VxSvgfVarianceOut, VxSvgfVarianceParams)

ATROUS_STEPS = (16, 8, 4, 2, 1)          # Core/Pipeline.cpp:2482-2487
ATROUS_STEPS_WIDE = (32, 16, 8, 4, 2)    # WiderSVGF, :2473-2478


def _gb(g, address):
    s = VxGBuffer()
    s.t, s.normal_id, s.block_id = address(g.get("t")), address(g.get("normal_id")), address(g.get("block_id"))
    return s


def temporal_params(prev_view, prev_projection, be_useful=True):
    """prev_view / prev_projection: 16 floats each, column-major (glm::value_ptr) — u_PrevView, u_PrevProjection."""
    p = VxSvgfTemporalParams()
    p.prev_view[:] = [float(v) for v in np.asarray(prev_view, dtype=np.float32).reshape(16)]
    p.prev_projection[:] = [float(v) for v in np.asarray(prev_projection, dtype=np.float32).reshape(16)]
    p.be_useful = int(bool(be_useful))
    return p


def variance_params(do_spatial=True, aggressive_disocclusion=True):
    p = VxSvgfVarianceParams()
    p.do_spatial, p.aggressive_disocclusion = int(bool(do_spatial)), int(bool(aggressive_disocclusion))
    return p


def spatial_params(step, time=0.0, large_kernel=False, do_spatial=True, aggressive_disocclusion=True, color_phi_bias=3.325, resolution_scale=0.25):
    """Defaults: ColorPhiBias 3.325, DiffuseIndirectSuperSampleRes 0.25 (Core/Pipeline.cpp:78,85)."""
    p = VxSvgfSpatialParams()
    p.step, p.large_kernel, p.do_spatial, p.aggressive_disocclusion = int(step), int(bool(large_kernel)), int(bool(do_spatial)), int(bool(aggressive_disocclusion))
    p.color_phi_bias, p.time, p.resolution_scale = float(color_phi_bias), float(time), float(resolution_scale)
    return p


def frame_params(view, projection, time=0.0, reset_history=False, pre_pass=True, wide=False, large_kernel=False, aggressive_disocclusion=True,
                 color_phi_bias=3.325, resolution_scale=0.25):
    """vxpt_svgf_frame: this frame's u_View / u_Projection (16 floats each, column-major) and the chain's switches at the reference's defaults."""
    p = VxSvgfFrameParams()
    p.view[:] = [float(v) for v in np.asarray(view, dtype=np.float32).reshape(16)]
    p.projection[:] = [float(v) for v in np.asarray(projection, dtype=np.float32).reshape(16)]
    p.reset_history, p.pre_pass, p.wide, p.large_kernel = int(bool(reset_history)), int(bool(pre_pass)), int(bool(wide)), int(bool(large_kernel))
    p.aggressive_disocclusion = int(bool(aggressive_disocclusion))
    p.color_phi_bias, p.time, p.resolution_scale = float(color_phi_bias), float(time), float(resolution_scale)
    return p


def shadow_frame_params(view, projection, reset_history=False, spatial=True, filter_scale=1.0):
    """vxpt_shadow_filter_frame: this frame's u_View / u_Projection and the switches of Core/Pipeline.cpp:132-133."""
    p = VxShadowFrameParams()
    p.view[:] = [float(v) for v in np.asarray(view, dtype=np.float32).reshape(16)]
    p.projection[:] = [float(v) for v in np.asarray(projection, dtype=np.float32).reshape(16)]
    p.reset_history, p.spatial, p.filter_scale = int(bool(reset_history)), int(bool(spatial)), float(filter_scale)
    return p


def initial_structs(gbuf, diffuse, out, address):
    i = VxSvgfInitialIn()
    i.current = _gb(gbuf, address)
    i.sh, i.cocg, i.luma, i.ao_sky = (address(diffuse[k]) for k in ("sh", "cocg", "luma", "ao_sky"))
    o = VxSvgfInitialOut()
    o.sh, o.cocg, o.luma, o.ao_sky = (address(out.get(k)) for k in ("sh", "cocg", "luma", "ao_sky"))
    return i, o


def temporal_structs(gbuf, prev_gbuf, diffuse, prev_temporal, out, address):
    i = VxSvgfTemporalIn()
    i.current, i.previous = _gb(gbuf, address), _gb(prev_gbuf, address)
    i.sh, i.cocg, i.luma, i.ao_sky = (address(diffuse[k]) for k in ("sh", "cocg", "luma", "ao_sky"))
    i.prev_sh, i.prev_cocg, i.prev_utility, i.prev_ao_sky = (address(prev_temporal[k]) for k in ("sh", "cocg", "utility", "ao_sky"))
    o = VxSvgfTemporalOut()
    o.sh, o.cocg, o.utility, o.ao_sky = (address(out.get(k)) for k in ("sh", "cocg", "utility", "ao_sky"))
    return i, o


def variance_structs(gbuf, temporal, out, address):
    i = VxSvgfVarianceIn()
    i.current = _gb(gbuf, address)
    i.sh, i.cocg, i.utility = (address(temporal[k]) for k in ("sh", "cocg", "utility"))
    o = VxSvgfVarianceOut()
    o.sh, o.cocg, o.variance = (address(out.get(k)) for k in ("sh", "cocg", "variance"))
    return i, o


def spatial_structs(gbuf, planes, temporal_utility, out, address):
    i = VxSvgfSpatialIn()
    i.current = _gb(gbuf, address)
    i.sh, i.cocg, i.variance, i.ao_sky = (address(planes[k]) for k in ("sh", "cocg", "variance", "ao_sky"))
    i.temporal_utility = address(temporal_utility)
    o = VxSvgfSpatialOut()
    o.sh, o.cocg, o.variance, o.ao_sky = (address(out.get(k)) for k in ("sh", "cocg", "variance", "ao_sky"))
    return i, o


def shadow_temporal_params(prev_view, prev_projection):
    p = VxShadowTemporalParams()
    p.prev_view[:] = [float(v) for v in np.asarray(prev_view, dtype=np.float32).reshape(16)]
    p.prev_projection[:] = [float(v) for v in np.asarray(prev_projection, dtype=np.float32).reshape(16)]
    return p


def shadow_filter_params(filter_scale=1.0):
    p = VxShadowFilterParams()
    p.filter_scale = float(filter_scale)
    return p


def shadow_temporal_structs(gbuf, prev_gbuf, shadow, prev_temporal, out, address):
    """shadow: the shadow pass's planes (shadow uint8, transversal); prev_temporal / out: dicts with "shadow" and "frames" (fp32)."""
    i = VxShadowTemporalIn()
    i.current, i.previous = _gb(gbuf, address), _gb(prev_gbuf, address)
    i.shadow, i.transversal = address(shadow["shadow"]), address(shadow["transversal"])
    i.prev_shadow, i.prev_frames = address(prev_temporal["shadow"]), address(prev_temporal["frames"])
    o = VxShadowTemporalOut()
    o.shadow, o.frames = address(out.get("shadow")), address(out.get("frames"))
    return i, o


def shadow_filter_struct(gbuf, temporal, transversal, address):
    i = VxShadowFilterIn()
    i.current = _gb(gbuf, address)
    i.shadow, i.transversal, i.frames = address(temporal["shadow"]), address(transversal), address(temporal["frames"])
    return i


def plane_shapes(width, height):
    """name -> shape of every fp32 plane the denoiser passes exchange."""
    return {"sh": (height, width, 4), "cocg": (height, width, 2), "utility": (height, width, 3), "ao_sky": (height, width, 2),
            "variance": (height, width), "luma": (height, width), "shadow": (height, width), "frames": (height, width)}
