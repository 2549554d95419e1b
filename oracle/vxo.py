"""ctypes loader of the CPU oracle (oracle/libvxo_oracle.so).  TEST INFRASTRUCTURE ONLY: imported by tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs, never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from voxelpathtracer_b200.abi import (VxCamera, VxDiffuseOut, VxDiffuseParams, VxGBuffer, VxMaterialOut, VxMaterialParams, VxPrimaryParams, VxReflectionIn, VxReflectionOut,
                                      VxReflectionParams, VxShadowOut, VxShadowParams)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvxo_oracle.so")


class VxoScene(C.Structure):
    _fields_ = [("wx", C.c_int32), ("wy", C.c_int32), ("wz", C.c_int32), ("grid", C.c_void_p), ("df", C.c_void_p),
                ("materials", C.c_void_p), ("sobol", C.c_void_p), ("scramble", C.c_void_p), ("rank", C.c_void_p),
                ("albedo_lod3", C.c_void_p), ("pbr_lod2", C.c_void_p), ("n_layers", C.c_int32),
                ("emissive_lod0", C.c_void_p), ("n_emissive_layers", C.c_int32), ("sky", C.c_void_p), ("sky_n", C.c_int32),
                ("shadow_noise", C.c_void_p), ("normal_lod3", C.c_void_p), ("n_normal_layers", C.c_int32), ("emissive_lod2", C.c_void_p),
                ("alpha_mips", C.c_void_p), ("n_alpha_layers", C.c_int32),
                ("albedo_mips", C.c_void_p), ("normal_mips", C.c_void_p), ("pbr_mips", C.c_void_p), ("n_mip_layers", C.c_int32),
                ("lava_albedo", C.c_void_p), ("lava_normal", C.c_void_p)]


class VxoStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("df_fetches", C.c_uint64), ("vox_fetches", C.c_uint64)]


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", HERE, "libvxo_oracle.so"])
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    srcs = [os.path.join(HERE, f) for f in ("vxo_oracle.cpp", "vxo_denoise.cpp")]
    if not os.path.exists(LIB_PATH) or any(os.path.getmtime(src) > os.path.getmtime(LIB_PATH) for src in srcs):
        build()
    lib = C.CDLL(LIB_PATH)
    lib.vxo_df_build.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.vxo_df_build.restype = None
    lib.vxo_df_bruteforce.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    lib.vxo_df_bruteforce.restype = None
    lib.vxo_traverse.argtypes = [C.POINTER(VxoScene), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int32), C.POINTER(VxoStats)]
    lib.vxo_traverse.restype = C.c_float
    lib.vxo_plain_dda.argtypes = [C.POINTER(VxoScene), C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.vxo_plain_dda.restype = C.c_int
    lib.vxo_trace_primary.argtypes = [C.POINTER(VxoScene), C.POINTER(VxCamera), C.POINTER(VxPrimaryParams), C.POINTER(VxGBuffer), C.POINTER(VxoStats)]
    lib.vxo_trace_shadow.argtypes = [C.POINTER(VxoScene), C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxShadowParams), C.POINTER(VxShadowOut), C.POINTER(VxoStats)]
    lib.vxo_trace_diffuse.argtypes = [C.POINTER(VxoScene), C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxDiffuseParams), C.POINTER(VxDiffuseOut), C.POINTER(VxoStats)]
    lib.vxo_trace_reflection.argtypes = [C.POINTER(VxoScene), C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxReflectionIn),
                                         C.POINTER(VxReflectionParams), C.POINTER(VxReflectionOut), C.POINTER(VxoStats)]
    for f in (lib.vxo_trace_primary, lib.vxo_trace_shadow, lib.vxo_trace_diffuse, lib.vxo_trace_reflection):
        f.restype = C.c_int
    lib.vxo_generate_gbuffer.argtypes = [C.POINTER(VxoScene), C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxMaterialParams), C.POINTER(VxMaterialOut)]
    lib.vxo_generate_gbuffer.restype = C.c_int
    from voxelpathtracer_b200 import abi as _abi
    for name, kinds in (("temporal", ("TemporalIn", "TemporalParams", "TemporalOut")), ("variance", ("VarianceIn", "VarianceParams", "VarianceOut")),
                        ("spatial", ("SpatialIn", "SpatialParams", "SpatialOut"))):
        fn = getattr(lib, "vxo_svgf_" + name)
        fn.argtypes = [C.POINTER(VxCamera)] + [C.POINTER(getattr(_abi, "VxSvgf" + k)) for k in kinds]
        fn.restype = C.c_int
    lib.vxo_svgf_initial.argtypes = [C.POINTER(VxCamera), C.POINTER(_abi.VxSvgfInitialIn), C.POINTER(_abi.VxSvgfInitialOut)]
    lib.vxo_svgf_initial.restype = C.c_int
    lib.vxo_shadow_temporal.argtypes = [C.POINTER(VxCamera), C.POINTER(_abi.VxShadowTemporalIn), C.POINTER(_abi.VxShadowTemporalParams),
                                        C.POINTER(_abi.VxShadowTemporalOut)]
    lib.vxo_shadow_filter.argtypes = [C.POINTER(VxCamera), C.POINTER(_abi.VxShadowFilterIn), C.POINTER(_abi.VxShadowFilterParams), C.c_void_p]
    lib.vxo_shadow_temporal.restype = lib.vxo_shadow_filter.restype = C.c_int
    lib.vxo_trace_rays.argtypes = [C.POINTER(VxoScene), C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(VxoStats)]
    lib.vxo_trace_rays.restype = C.c_int
    lib.vxo_player_shadowed.argtypes = [C.POINTER(VxoScene), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.vxo_player_shadowed.restype = C.c_int
    lib.vxo_ambient_sound.argtypes = [C.POINTER(VxoScene), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_uint32), C.c_void_p, C.POINTER(VxoStats)]
    lib.vxo_ambient_sound.restype = C.c_int
    lib.vxo_num_threads.restype = C.c_int
    lib.vxo_set_num_threads.argtypes = [C.c_int]
    _lib = lib
    return lib


def df_build(grid, dims=(384, 128, 384)):
    """Literal three-pass distance field (ManhattanDistance{X,Y,Z}.comp)."""
    grid = np.ascontiguousarray(grid, dtype=np.uint8).reshape(-1)
    assert grid.size == dims[0] * dims[1] * dims[2]
    out = np.empty_like(grid)
    load().vxo_df_build(grid.ctypes.data, out.ctypes.data, *dims)
    return out


def df_bruteforce(grid, dims):
    grid = np.ascontiguousarray(grid, dtype=np.uint8).reshape(-1)
    out = np.empty_like(grid)
    load().vxo_df_bruteforce(grid.ctypes.data, out.ctypes.data, *dims)
    return out


class Oracle:
    """Holds the scene arrays (keeps them alive) and runs the oracle passes on numpy buffers."""

    def __init__(self, grid, df=None, dims=(384, 128, 384)):
        self.lib = load()
        self.dims = dims
        self.grid = np.ascontiguousarray(grid, dtype=np.uint8).reshape(-1)
        self.df = df_build(self.grid, dims) if df is None else np.ascontiguousarray(df, dtype=np.uint8).reshape(-1)
        self.scene = VxoScene()
        self.scene.wx, self.scene.wy, self.scene.wz = dims
        self.scene.grid, self.scene.df = self.grid.ctypes.data, self.df.ctypes.data
        self._keep = {}

    def set_tables(self, materials, blue_noise, sky, shadow_noise):
        k = self._keep
        k["table"] = np.ascontiguousarray(materials["table"], dtype=np.int32)
        k["albedo"] = np.ascontiguousarray(materials["albedo_lod3"], dtype=np.float32)
        k["pbr"] = np.ascontiguousarray(materials["pbr_lod2"], dtype=np.float32)
        k["emissive"] = np.ascontiguousarray(materials["emissive_lod0"], dtype=np.float32)
        k["sobol"], k["scramble"], k["rank"] = (np.ascontiguousarray(x, dtype=np.int32) for x in blue_noise)
        k["sky"] = np.ascontiguousarray(sky, dtype=np.float32)
        k["shadow_noise"] = np.ascontiguousarray(shadow_noise, dtype=np.uint8)
        s = self.scene
        s.materials = k["table"].ctypes.data
        s.albedo_lod3, s.pbr_lod2, s.n_layers = k["albedo"].ctypes.data, k["pbr"].ctypes.data, k["albedo"].shape[0]
        s.emissive_lod0, s.n_emissive_layers = (k["emissive"].ctypes.data if k["emissive"].shape[0] else None), k["emissive"].shape[0]
        s.sobol, s.scramble, s.rank = k["sobol"].ctypes.data, k["scramble"].ctypes.data, k["rank"].ctypes.data
        s.sky, s.sky_n = k["sky"].ctypes.data, k["sky"].shape[1]
        s.shadow_noise = k["shadow_noise"].ctypes.data
        if "normal_lod3" in materials:
            k["normal"] = np.ascontiguousarray(materials["normal_lod3"], dtype=np.float32)
            k["emissive2"] = np.ascontiguousarray(materials["emissive_lod2"], dtype=np.float32)
            s.normal_lod3, s.n_normal_layers = k["normal"].ctypes.data, k["normal"].shape[0]
            s.emissive_lod2 = k["emissive2"].ctypes.data if k["emissive2"].shape[0] else None

    def set_alpha_mips(self, alpha_mips):
        """uint8 [layers][ALPHA_MIP_TEXELS] (assets.alpha_mip_pyramid): enables params.alpha_test in trace_primary / trace_shadow."""
        a = np.ascontiguousarray(alpha_mips, dtype=np.uint8)
        self._keep["alpha"] = a
        self.scene.alpha_mips, self.scene.n_alpha_layers = a.ctypes.data, a.shape[0]

    def set_gbuffer_textures(self, albedo_mips, normal_mips, pbr_mips):
        """uint8 [layers][MIP_CHAIN_TEXELS][4] each (assets.rgba_mip_chain): inputs of generate_gbuffer."""
        arrs = [np.ascontiguousarray(a, dtype=np.uint8) for a in (albedo_mips, normal_mips, pbr_mips)]
        self._keep["gb_mips"] = arrs
        s = self.scene
        s.albedo_mips, s.normal_mips, s.pbr_mips, s.n_mip_layers = arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, arrs[0].shape[0]

    def set_lava_textures(self, albedo_rgba8, normal_rgba8):
        arrs = [np.ascontiguousarray(a, dtype=np.uint8) for a in (albedo_rgba8, normal_rgba8)]
        self._keep["lava"] = arrs
        self.scene.lava_albedo, self.scene.lava_normal = arrs[0].ctypes.data, arrs[1].ctypes.data

    def generate_gbuffer(self, cam, gbuf, params, out=None):
        H, W = cam.height, cam.width
        if out is None:
            out = {"albedo": np.zeros((H, W, 3), np.float32), "normal": np.zeros((H, W, 3), np.float32), "pbr": np.zeros((H, W, 4), np.float32),
                   "texture_ao": np.zeros((H, W), np.float32)}
        g = VxGBuffer()
        g.inv_t, g.normal_id, g.block_id = gbuf["inv_t"].ctypes.data, gbuf["normal_id"].ctypes.data, gbuf["block_id"].ctypes.data
        o = VxMaterialOut()
        o.albedo, o.normal, o.pbr, o.texture_ao = (out[k].ctypes.data for k in ("albedo", "normal", "pbr", "texture_ao"))
        rc = self.lib.vxo_generate_gbuffer(C.byref(self.scene), C.byref(cam), C.byref(g), C.byref(params), C.byref(o))
        assert rc == 0, rc
        return out

    @staticmethod
    def _stats(st):
        return {"rays": int(st.rays), "df_fetches": int(st.df_fetches), "vox_fetches": int(st.vox_fetches)}

    def traverse(self, origin, direction, max_it):
        o = (C.c_float * 3)(*[float(v) for v in origin])
        d = (C.c_float * 3)(*[float(v) for v in direction])
        out = (C.c_int32 * 6)()
        st = VxoStats()
        t = self.lib.vxo_traverse(C.byref(self.scene), o, d, int(max_it), out, C.byref(st))
        return {"t": float(t), "min_idx": out[0], "sgn": out[1], "block": out[2], "voxel": (out[3], out[4], out[5]), **self._stats(st)}

    def plain_dda(self, origin, direction, max_steps=2000):
        o = (C.c_float * 3)(*[float(v) for v in origin])
        d = (C.c_float * 3)(*[float(v) for v in direction])
        vox = (C.c_int32 * 3)()
        axis = C.c_int32()
        hit = self.lib.vxo_plain_dda(C.byref(self.scene), o, d, int(max_steps), vox, C.byref(axis))
        return (bool(hit), (vox[0], vox[1], vox[2]), int(axis.value))

    def trace_primary(self, cam, params, hit_voxel=True):
        H, W = cam.height, cam.width
        g = {"t": np.zeros((H, W), np.float32), "normal_id": np.zeros((H, W), np.uint8), "block_id": np.zeros((H, W), np.uint8),
             "inv_t": np.zeros((H, W), np.float32)}
        if hit_voxel:
            g["hit_voxel"] = np.zeros((H, W, 3), np.int16)
        s = VxGBuffer()
        s.t, s.normal_id, s.block_id, s.inv_t = (g[k].ctypes.data for k in ("t", "normal_id", "block_id", "inv_t"))
        s.hit_voxel = g["hit_voxel"].ctypes.data if hit_voxel else None
        st = VxoStats()
        rc = self.lib.vxo_trace_primary(C.byref(self.scene), C.byref(cam), C.byref(params), C.byref(s), C.byref(st))
        assert rc == 0, rc
        return g, self._stats(st)

    def trace_shadow(self, cam, gbuf, params):
        H, W = cam.height, cam.width
        out = {"shadow": np.zeros((H, W), np.uint8), "transversal": np.zeros((H, W), np.float32)}
        g = VxGBuffer()
        g.t, g.normal_id = gbuf["t"].ctypes.data, gbuf["normal_id"].ctypes.data
        o = VxShadowOut()
        o.shadow, o.transversal = out["shadow"].ctypes.data, out["transversal"].ctypes.data
        st = VxoStats()
        rc = self.lib.vxo_trace_shadow(C.byref(self.scene), C.byref(cam), C.byref(g), C.byref(params), C.byref(o), C.byref(st))
        assert rc == 0, rc
        return out, self._stats(st)

    def trace_diffuse(self, cam, gbuf, params):
        H, W = cam.height, cam.width
        out = {"sh": np.zeros((H, W, 4), np.float32), "cocg": np.zeros((H, W, 2), np.float32), "luma": np.zeros((H, W), np.float32),
               "ao_sky": np.zeros((H, W, 2), np.float32)}
        g = VxGBuffer()
        g.t, g.normal_id = gbuf["t"].ctypes.data, gbuf["normal_id"].ctypes.data
        o = VxDiffuseOut()
        o.sh, o.cocg, o.luma, o.ao_sky = (out[k].ctypes.data for k in ("sh", "cocg", "luma", "ao_sky"))
        st = VxoStats()
        rc = self.lib.vxo_trace_diffuse(C.byref(self.scene), C.byref(cam), C.byref(g), C.byref(params), C.byref(o), C.byref(st))
        assert rc == 0, rc
        return out, self._stats(st)

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
        st = VxoStats()
        rc = self.lib.vxo_trace_reflection(C.byref(self.scene), C.byref(cam), C.byref(g), C.byref(i), C.byref(params), C.byref(o), C.byref(st))
        assert rc == 0, rc
        return out, self._stats(st)

    def trace_rays(self, origins, directions, max_it):
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
        n = o.shape[0]
        out = {"t": np.zeros(n, np.float32), "normal_id": np.zeros(n, np.uint8), "block_id": np.zeros(n, np.uint8), "hit_voxel": np.zeros((n, 3), np.int16)}
        st = VxoStats()
        rc = self.lib.vxo_trace_rays(C.byref(self.scene), o.ctypes.data, d.ctypes.data, n, int(max_it), out["t"].ctypes.data,
                                     out["normal_id"].ctypes.data, out["block_id"].ctypes.data, out["hit_voxel"].ctypes.data, C.byref(st))
        assert rc == 0, rc
        return out, self._stats(st)

    def player_shadowed(self, camera_pos, sun_dir):
        p = (C.c_float * 3)(*[float(v) for v in camera_pos])
        s = (C.c_float * 3)(*[float(v) for v in sun_dir])
        return bool(self.lib.vxo_player_shadowed(C.byref(self.scene), p, s))

    def ambient_sound(self, player_pos, frame):
        p = (C.c_float * 3)(*[float(v) for v in player_pos])
        agg = C.c_uint32()
        per = np.zeros(32, np.uint32)
        st = VxoStats()
        rc = self.lib.vxo_ambient_sound(C.byref(self.scene), p, int(frame), C.byref(agg), per.ctypes.data, C.byref(st))
        assert rc == 0, rc
        return int(agg.value), per, self._stats(st)


# ---- SVGF denoiser (oracle/vxo_denoise.cpp): plain functions of planes, no scene ------------------------------------------------------
def _addr(a):
    return None if a is None else a.ctypes.data


def _planes(cam, names):
    from voxelpathtracer_b200 import denoise
    shapes = denoise.plane_shapes(cam.width, cam.height)
    return {k: np.zeros(shapes[k], np.float32) for k in names}


def svgf_initial(cam, gbuf, diffuse, out=None):
    from voxelpathtracer_b200 import denoise
    out = _planes(cam, ("sh", "cocg", "luma", "ao_sky")) if out is None else out
    i, o = denoise.initial_structs(gbuf, diffuse, out, _addr)
    rc = load().vxo_svgf_initial(C.byref(cam), C.byref(i), C.byref(o))
    assert rc == 0, rc
    return out


def svgf_temporal(cam, gbuf, prev_gbuf, diffuse, prev_temporal, params, out=None):
    from voxelpathtracer_b200 import denoise
    out = _planes(cam, ("sh", "cocg", "utility", "ao_sky")) if out is None else out
    i, o = denoise.temporal_structs(gbuf, prev_gbuf, diffuse, prev_temporal, out, _addr)
    rc = load().vxo_svgf_temporal(C.byref(cam), C.byref(i), C.byref(params), C.byref(o))
    assert rc == 0, rc
    return out


def svgf_variance(cam, gbuf, temporal, params, out=None):
    from voxelpathtracer_b200 import denoise
    out = _planes(cam, ("sh", "cocg", "variance")) if out is None else out
    i, o = denoise.variance_structs(gbuf, temporal, out, _addr)
    rc = load().vxo_svgf_variance(C.byref(cam), C.byref(i), C.byref(params), C.byref(o))
    assert rc == 0, rc
    return out


def svgf_spatial(cam, gbuf, planes, temporal_utility, params, out=None):
    from voxelpathtracer_b200 import denoise
    out = _planes(cam, ("sh", "cocg", "variance", "ao_sky")) if out is None else out
    i, o = denoise.spatial_structs(gbuf, planes, temporal_utility, out, _addr)
    rc = load().vxo_svgf_spatial(C.byref(cam), C.byref(i), C.byref(params), C.byref(o))
    assert rc == 0, rc
    return out


def shadow_temporal(cam, gbuf, prev_gbuf, shadow, prev_temporal, params, out=None):
    from voxelpathtracer_b200 import denoise
    out = _planes(cam, ("shadow", "frames")) if out is None else out
    i, o = denoise.shadow_temporal_structs(gbuf, prev_gbuf, shadow, prev_temporal, out, _addr)
    rc = load().vxo_shadow_temporal(C.byref(cam), C.byref(i), C.byref(params), C.byref(o))
    assert rc == 0, rc
    return out


def shadow_filter(cam, gbuf, temporal, transversal, params, out=None):
    from voxelpathtracer_b200 import denoise
    out = np.zeros((cam.height, cam.width), np.float32) if out is None else out
    i = denoise.shadow_filter_struct(gbuf, temporal, transversal, _addr)
    rc = load().vxo_shadow_filter(C.byref(cam), C.byref(i), C.byref(params), out.ctypes.data)
    assert rc == 0, rc
    return out
