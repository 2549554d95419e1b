"""ctypes loader of tests/host_shadow/libkernels_on_host.so: the per-pixel CUDA kernels' own source compiled by g++ and run on the
CPU (see kernels_on_host.cpp).  TEST INFRASTRUCTURE ONLY — never imported by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from voxelpathtracer_b200.abi import (VxCamera, VxDiffuseOut, VxDiffuseParams, VxGBuffer, VxMaterialOut, VxMaterialParams, VxPrimaryParams, VxReflectionIn, VxReflectionOut,
                                      VxReflectionParams, VxShadowOut, VxShadowParams)

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB_PATH = os.path.join(HERE, "libkernels_on_host.so")
CSRC = os.path.join(ROOT, "voxelpathtracer_b200", "csrc")
SOURCES = [os.path.join(HERE, "kernels_on_host.cpp"), os.path.join(HERE, "warp_emu.h")] + [os.path.join(CSRC, f) for f in
                                                         ("trace.cu", "trace_reflection.cu", "df_consumers.cu", "gbuffer.cu", "denoise.cu", "trace_device.cuh",
                                                          "gi_device.cuh", "vxpt_internal.h")]
CUDA_INCLUDE = "/usr/local/cuda/include"
_lib = None


def build():
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    # same floating-point contract as nvcc -fmad=false -prec-div=true -prec-sqrt=true
    cmd = [cxx, "-std=c++17", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fPIC", "-fopenmp", "-shared", "-fvisibility=hidden",
           "-Wno-attributes", "-Wno-unknown-pragmas", "-I" + CUDA_INCLUDE, "-x", "c++", SOURCES[0], "-o", LIB_PATH]
    subprocess.check_call(cmd, cwd=HERE)
    return LIB_PATH


def available():
    return os.path.exists(os.path.join(CUDA_INCLUDE, "cuda_runtime.h")) or os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is not None:
        return _lib
    have = [s for s in SOURCES if os.path.exists(s)]
    if not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in have):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.hs_create.restype = C.c_void_p
    lib.hs_create.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.hs_destroy.argtypes = [C.c_void_p]
    lib.hs_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
    lib.hs_trace_primary.argtypes = [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxPrimaryParams), C.POINTER(VxGBuffer)]
    lib.hs_trace_shadow.argtypes = [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxShadowParams), C.POINTER(VxShadowOut)]
    lib.hs_trace_diffuse.argtypes = [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxDiffuseParams), C.POINTER(VxDiffuseOut)]
    lib.hs_trace_reflection.argtypes = [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxReflectionIn),
                                        C.POINTER(VxReflectionParams), C.POINTER(VxReflectionOut)]
    lib.hs_generate_gbuffer.argtypes = [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxMaterialParams), C.POINTER(VxMaterialOut)]
    from voxelpathtracer_b200 import abi as _abi
    for name, kinds in (("temporal", ("TemporalIn", "TemporalParams", "TemporalOut")), ("variance", ("VarianceIn", "VarianceParams", "VarianceOut")),
                        ("spatial", ("SpatialIn", "SpatialParams", "SpatialOut"))):
        getattr(lib, "hs_svgf_" + name).argtypes = [C.c_void_p, C.POINTER(VxCamera)] + [C.POINTER(getattr(_abi, "VxSvgf" + k)) for k in kinds]
    lib.hs_svgf_initial.argtypes = [C.c_void_p, C.POINTER(VxCamera), C.POINTER(_abi.VxSvgfInitialIn), C.POINTER(_abi.VxSvgfInitialOut)]
    lib.hs_shadow_temporal.argtypes = [C.c_void_p, C.POINTER(VxCamera), C.POINTER(_abi.VxShadowTemporalIn), C.POINTER(_abi.VxShadowTemporalParams),
                                       C.POINTER(_abi.VxShadowTemporalOut)]
    lib.hs_shadow_filter.argtypes = [C.c_void_p, C.POINTER(VxCamera), C.POINTER(_abi.VxShadowFilterIn), C.POINTER(_abi.VxShadowFilterParams), C.c_void_p]
    lib.hs_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.hs_ambient_sound.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint32), C.c_void_p]
    lib.hs_exp_cr.argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    lib.hs_pow01_cr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_long]
    lib.hs_exp_short_refused.argtypes = [C.c_void_p, C.c_long]
    lib.hs_exp_short_refused.restype = C.c_long
    lib.hs_normal_weight.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int]
    lib.hs_normal_weight.restype = C.c_float
    _lib = lib
    return lib


def _addr(a):
    return None if a is None else a.ctypes.data


def _denoise_planes(cam, names):
    from voxelpathtracer_b200 import denoise
    shapes = denoise.plane_shapes(cam.width, cam.height)
    return {k: np.zeros(shapes[k], np.float32) for k in names}


class HostKernels:
    """Runs the kernels on the arrays of an oracle.vxo.Oracle (same scene struct), returning planes shaped like the oracle's."""

    def __init__(self, oracle, layout=1):
        self.lib = load()
        self.oracle = oracle  # keeps the arrays alive
        self.h = self.lib.hs_create(C.addressof(oracle.scene), layout, 0)

    def close(self):
        if self.h:
            self.lib.hs_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def stats(self, reset=True):
        out = (C.c_uint64 * 3)()
        self.lib.hs_stats(self.h, out, int(reset))
        return {"rays": int(out[0]), "df_fetches": int(out[1]), "vox_fetches": int(out[2])}

    def trace_primary(self, cam, params, hit_voxel=True):
        H, W = cam.height, cam.width
        g = {"t": np.zeros((H, W), np.float32), "normal_id": np.zeros((H, W), np.uint8), "block_id": np.zeros((H, W), np.uint8),
             "inv_t": np.zeros((H, W), np.float32)}
        if hit_voxel:
            g["hit_voxel"] = np.zeros((H, W, 3), np.int16)
        s = VxGBuffer()
        s.t, s.normal_id, s.block_id, s.inv_t = (g[k].ctypes.data for k in ("t", "normal_id", "block_id", "inv_t"))
        s.hit_voxel = g["hit_voxel"].ctypes.data if hit_voxel else None
        rc = self.lib.hs_trace_primary(self.h, C.byref(cam), C.byref(params), C.byref(s))
        assert rc == 0, rc
        return g, self.stats()

    def trace_shadow(self, cam, gbuf, params):
        H, W = cam.height, cam.width
        out = {"shadow": np.zeros((H, W), np.uint8), "transversal": np.zeros((H, W), np.float32)}
        g = VxGBuffer()
        g.t, g.normal_id = gbuf["t"].ctypes.data, gbuf["normal_id"].ctypes.data
        o = VxShadowOut()
        o.shadow, o.transversal = out["shadow"].ctypes.data, out["transversal"].ctypes.data
        rc = self.lib.hs_trace_shadow(self.h, C.byref(cam), C.byref(g), C.byref(params), C.byref(o))
        assert rc == 0, rc
        return out, self.stats()

    def trace_diffuse(self, cam, gbuf, params):
        H, W = cam.height, cam.width
        out = {"sh": np.zeros((H, W, 4), np.float32), "cocg": np.zeros((H, W, 2), np.float32), "luma": np.zeros((H, W), np.float32),
               "ao_sky": np.zeros((H, W, 2), np.float32)}
        g = VxGBuffer()
        g.t, g.normal_id = gbuf["t"].ctypes.data, gbuf["normal_id"].ctypes.data
        o = VxDiffuseOut()
        o.sh, o.cocg, o.luma, o.ao_sky = (out[k].ctypes.data for k in ("sh", "cocg", "luma", "ao_sky"))
        rc = self.lib.hs_trace_diffuse(self.h, C.byref(cam), C.byref(g), C.byref(params), C.byref(o))
        assert rc == 0, rc
        return out, self.stats()

    def trace_reflection(self, cam, gbuf, diffuse, params, g_normal=None, g_pbr=None):
        H, W = cam.height, cam.width
        out = {"color": np.zeros((H, W, 4), np.float32), "hit_distance": np.zeros((H, W), np.float32), "emissive_mask": np.zeros((H, W), np.uint8)}
        g = VxGBuffer()
        g.t, g.normal_id, g.block_id = gbuf["t"].ctypes.data, gbuf["normal_id"].ctypes.data, gbuf["block_id"].ctypes.data
        i = VxReflectionIn()
        i.sh, i.cocg = diffuse["sh"].ctypes.data, diffuse["cocg"].ctypes.data
        i.g_normal = g_normal.ctypes.data if g_normal is not None else None
        i.g_pbr = g_pbr.ctypes.data if g_pbr is not None else None
        o = VxReflectionOut()
        o.color, o.hit_distance, o.emissive_mask = out["color"].ctypes.data, out["hit_distance"].ctypes.data, out["emissive_mask"].ctypes.data
        rc = self.lib.hs_trace_reflection(self.h, C.byref(cam), C.byref(g), C.byref(i), C.byref(params), C.byref(o))
        assert rc == 0, rc
        return out, self.stats()

    def generate_gbuffer(self, cam, gbuf, params, out=None):
        H, W = cam.height, cam.width
        if out is None:
            out = {"albedo": np.zeros((H, W, 3), np.float32), "normal": np.zeros((H, W, 3), np.float32), "pbr": np.zeros((H, W, 4), np.float32),
                   "texture_ao": np.zeros((H, W), np.float32)}
        g = VxGBuffer()
        g.inv_t, g.normal_id, g.block_id = gbuf["inv_t"].ctypes.data, gbuf["normal_id"].ctypes.data, gbuf["block_id"].ctypes.data
        o = VxMaterialOut()
        o.albedo, o.normal, o.pbr, o.texture_ao = (out[k].ctypes.data for k in ("albedo", "normal", "pbr", "texture_ao"))
        rc = self.lib.hs_generate_gbuffer(self.h, C.byref(cam), C.byref(g), C.byref(params), C.byref(o))
        assert rc == 0, rc
        return out

    # SVGF denoiser passes (csrc/denoise.cu): same call shapes as oracle.vxo.svgf_*
    def svgf_initial(self, cam, gbuf, diffuse, out=None):
        from voxelpathtracer_b200 import denoise
        out = _denoise_planes(cam, ("sh", "cocg", "luma", "ao_sky")) if out is None else out
        i, o = denoise.initial_structs(gbuf, diffuse, out, _addr)
        assert self.lib.hs_svgf_initial(self.h, C.byref(cam), C.byref(i), C.byref(o)) == 0
        return out

    def svgf_temporal(self, cam, gbuf, prev_gbuf, diffuse, prev_temporal, params, out=None):
        from voxelpathtracer_b200 import denoise
        out = _denoise_planes(cam, ("sh", "cocg", "utility", "ao_sky")) if out is None else out
        i, o = denoise.temporal_structs(gbuf, prev_gbuf, diffuse, prev_temporal, out, _addr)
        assert self.lib.hs_svgf_temporal(self.h, C.byref(cam), C.byref(i), C.byref(params), C.byref(o)) == 0
        return out

    def svgf_variance(self, cam, gbuf, temporal, params, out=None):
        from voxelpathtracer_b200 import denoise
        out = _denoise_planes(cam, ("sh", "cocg", "variance")) if out is None else out
        i, o = denoise.variance_structs(gbuf, temporal, out, _addr)
        assert self.lib.hs_svgf_variance(self.h, C.byref(cam), C.byref(i), C.byref(params), C.byref(o)) == 0
        return out

    def svgf_spatial(self, cam, gbuf, planes, temporal_utility, params, out=None):
        from voxelpathtracer_b200 import denoise
        out = _denoise_planes(cam, ("sh", "cocg", "variance", "ao_sky")) if out is None else out
        i, o = denoise.spatial_structs(gbuf, planes, temporal_utility, out, _addr)
        assert self.lib.hs_svgf_spatial(self.h, C.byref(cam), C.byref(i), C.byref(params), C.byref(o)) == 0
        return out

    def shadow_temporal(self, cam, gbuf, prev_gbuf, shadow, prev_temporal, params, out=None):
        from voxelpathtracer_b200 import denoise
        out = _denoise_planes(cam, ("shadow", "frames")) if out is None else out
        i, o = denoise.shadow_temporal_structs(gbuf, prev_gbuf, shadow, prev_temporal, out, _addr)
        assert self.lib.hs_shadow_temporal(self.h, C.byref(cam), C.byref(i), C.byref(params), C.byref(o)) == 0
        return out

    def shadow_filter(self, cam, gbuf, temporal, transversal, params, out=None):
        from voxelpathtracer_b200 import denoise
        out = np.zeros((cam.height, cam.width), np.float32) if out is None else out
        i = denoise.shadow_filter_struct(gbuf, temporal, transversal, _addr)
        assert self.lib.hs_shadow_filter(self.h, C.byref(cam), C.byref(i), C.byref(params), out.ctypes.data) == 0
        return out

    def trace_rays(self, origins, directions, max_it):
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
        n = o.shape[0]
        out = {"t": np.zeros(n, np.float32), "normal_id": np.zeros(n, np.uint8), "block_id": np.zeros(n, np.uint8), "hit_voxel": np.zeros((n, 3), np.int16)}
        rc = self.lib.hs_trace_rays(self.h, o.ctypes.data, d.ctypes.data, n, int(max_it), out["t"].ctypes.data, out["normal_id"].ctypes.data,
                                    out["block_id"].ctypes.data, out["hit_voxel"].ctypes.data)
        assert rc == 0, rc
        return out, self.stats()

    def ambient_sound(self, player_pos, frame):
        p = np.array([float(v) for v in player_pos], dtype=np.float32)
        agg = C.c_uint32()
        per = np.zeros(32, np.uint32)
        rc = self.lib.hs_ambient_sound(self.h, p.ctypes.data, int(frame), C.byref(agg), per.ctypes.data)
        assert rc == 0, rc
        return int(agg.value), per, self.stats()
