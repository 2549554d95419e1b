"""ctypes binding of oracle/_ref/libref_shaders.so: the reference's OWN GLSL (ManhattanDistance{X,Y,Z}.comp, InitialRayTraceFrag.glsl)
compiled as C++ against its vendored glm (recipe: oracle/Makefile, oracle/glsl2cpp.py, oracle/ref_shader_driver.cpp).
Test infrastructure only: pins oracle/vxo_oracle.cpp and generated tests/golden/ref_shader_digests.json."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libref_shaders.so")
_lib = None


class RefDiffuseArgs(C.Structure):
    _fields_ = [("blocks", C.c_void_p), ("df", C.c_void_p), ("inv_view", C.c_void_p), ("inv_proj", C.c_void_p),
                ("width", C.c_int32), ("height", C.c_int32), ("row_begin", C.c_int32), ("row_end", C.c_int32),
                ("g_t", C.c_void_p), ("g_normal_id", C.c_void_p), ("materials", C.c_void_p), ("sobol", C.c_void_p), ("scramble", C.c_void_p),
                ("rank", C.c_void_p), ("albedo_lod3", C.c_void_p), ("pbr_lod2", C.c_void_p), ("emissive_lod0", C.c_void_p), ("sky", C.c_void_p),
                ("sky_n", C.c_int32), ("spp", C.c_int32), ("checker_spp", C.c_int32), ("checkerboard", C.c_int32), ("trace_length", C.c_int32),
                ("frame", C.c_int32), ("supersample", C.c_int32), ("halton", C.c_float * 2), ("sun_dir", C.c_float * 3), ("moon_dir", C.c_float * 3),
                ("sun_visibility", C.c_float), ("gi_sun_strength", C.c_float), ("gi_sky_strength", C.c_float), ("light_intensity", C.c_float),
                ("o_sh", C.c_void_p), ("o_cocg", C.c_void_p), ("o_utility", C.c_void_p), ("o_ao_sky", C.c_void_p)]


class RefReflectionArgs(C.Structure):
    _fields_ = [("blocks", C.c_void_p), ("df", C.c_void_p), ("inv_view", C.c_void_p), ("inv_proj", C.c_void_p),
                ("width", C.c_int32), ("height", C.c_int32), ("row_begin", C.c_int32), ("row_end", C.c_int32),
                ("g_t", C.c_void_p), ("g_normal_id", C.c_void_p), ("g_block_id", C.c_void_p), ("g_normal", C.c_void_p), ("g_pbr", C.c_void_p),
                ("sh", C.c_void_p), ("cocg", C.c_void_p), ("materials", C.c_void_p), ("sobol", C.c_void_p), ("scramble", C.c_void_p), ("rank", C.c_void_p),
                ("albedo_lod3", C.c_void_p), ("normal_lod3", C.c_void_p), ("pbr_lod2", C.c_void_p), ("emissive_lod2", C.c_void_p), ("sky", C.c_void_p),
                ("sky_n", C.c_int32), ("spp", C.c_int32), ("trace_length", C.c_int32), ("frame", C.c_int32), ("rough", C.c_int32),
                ("roughness_bias", C.c_int32), ("checkerboard", C.c_int32), ("sun_dir", C.c_float * 3), ("moon_dir", C.c_float * 3),
                ("stronger_dir", C.c_float * 3), ("viewer_pos", C.c_float * 3), ("sun_strength", C.c_float), ("moon_strength", C.c_float),
                ("halton", C.c_float * 2), ("grass_props", C.c_int32 * 10), ("o_color", C.c_void_p), ("o_hit_distance", C.c_void_p),
                ("o_emissive_mask", C.c_void_p)]


class RefGBufferArgs(C.Structure):
    _fields_ = [("inv_view", C.c_void_p), ("inv_proj", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("row_begin", C.c_int32),
                ("row_end", C.c_int32), ("g_inv_t", C.c_void_p), ("g_normal_id", C.c_void_p), ("g_block_id", C.c_void_p), ("materials", C.c_void_p),
                ("albedo_mips", C.c_void_p), ("normal_mips", C.c_void_p), ("pbr_mips", C.c_void_p), ("n_mip_layers", C.c_int32),
                ("emissive_lod0", C.c_void_p), ("update_this_frame", C.c_int32), ("grass_props", C.c_int32 * 10),
                ("pom", C.c_int32), ("high_quality_pom", C.c_int32), ("dither_pom", C.c_int32), ("frame", C.c_int32), ("pom_height", C.c_float),
                ("pom_exp", C.c_float), ("lava_block_id", C.c_int32), ("time", C.c_float), ("lava_albedo", C.c_void_p), ("lava_normal", C.c_void_p),
                ("o_albedo", C.c_void_p), ("o_normal", C.c_void_p), ("o_pbr", C.c_void_p), ("o_texture_ao", C.c_void_p)]


class RefSvgfArgs(C.Structure):
    _fields_ = [("inv_view", C.c_void_p), ("inv_proj", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("row_begin", C.c_int32),
                ("row_end", C.c_int32), ("g_t", C.c_void_p), ("g_normal_id", C.c_void_p), ("g_block_id", C.c_void_p), ("prev_t", C.c_void_p),
                ("prev_normal_id", C.c_void_p), ("prev_block_id", C.c_void_p), ("sh", C.c_void_p), ("cocg", C.c_void_p), ("luma", C.c_void_p),
                ("ao_sky", C.c_void_p), ("prev_sh", C.c_void_p), ("prev_cocg", C.c_void_p), ("prev_utility", C.c_void_p), ("prev_ao_sky", C.c_void_p),
                ("utility", C.c_void_p), ("variance", C.c_void_p), ("temporal_utility", C.c_void_p), ("prev_view", C.c_void_p),
                ("prev_projection", C.c_void_p), ("be_useful", C.c_int32), ("do_spatial", C.c_int32), ("aggressive", C.c_int32), ("step", C.c_int32),
                ("large_kernel", C.c_int32), ("color_phi_bias", C.c_float), ("time", C.c_float), ("resolution_scale", C.c_float),
                ("o_sh", C.c_void_p), ("o_cocg", C.c_void_p), ("o_utility", C.c_void_p), ("o_variance", C.c_void_p), ("o_ao_sky", C.c_void_p)]


class RefShadowFilterArgs(C.Structure):
    _fields_ = [("inv_view", C.c_void_p), ("inv_proj", C.c_void_p), ("width", C.c_int32), ("height", C.c_int32), ("row_begin", C.c_int32),
                ("row_end", C.c_int32), ("g_t", C.c_void_p), ("g_normal_id", C.c_void_p), ("prev_t", C.c_void_p), ("shadow_u8", C.c_void_p),
                ("shadow", C.c_void_p), ("transversal", C.c_void_p), ("prev_shadow", C.c_void_p), ("frames", C.c_void_p), ("prev_view", C.c_void_p),
                ("prev_projection", C.c_void_p), ("filter_scale", C.c_float), ("o_shadow", C.c_void_p), ("o_frames", C.c_void_p)]


def available():
    return os.path.exists(LIB_PATH)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        lib.ref_df_build.restype = C.c_int
        lib.ref_df_build.argtypes = [C.c_void_p, C.c_void_p]
        lib.ref_trace_primary.restype = C.c_int
        lib.ref_trace_primary.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_trace_primary_alpha.restype = C.c_int
        lib.ref_trace_primary_alpha.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_trace_shadow_alpha.restype = C.c_int
        lib.ref_trace_shadow_alpha.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_float,
                                               C.c_void_p, C.c_void_p]
        lib.ref_ambient_sound.restype = C.c_int
        lib.ref_ambient_sound.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_uint32), C.c_void_p]
        lib.ref_trace_shadow.restype = C.c_int
        lib.ref_trace_shadow.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.ref_trace_diffuse.restype = C.c_int
        lib.ref_trace_diffuse.argtypes = [C.POINTER(RefDiffuseArgs)]
        lib.ref_trace_reflection.restype = C.c_int
        lib.ref_trace_reflection.argtypes = [C.POINTER(RefReflectionArgs)]
        for name in ("ref_svgf_initial", "ref_svgf_temporal", "ref_svgf_variance", "ref_svgf_spatial"):
            if hasattr(lib, name):
                getattr(lib, name).restype = C.c_int
                getattr(lib, name).argtypes = [C.POINTER(RefSvgfArgs)]
        for name in ("ref_shadow_temporal", "ref_shadow_filter"):
            if hasattr(lib, name):
                getattr(lib, name).restype = C.c_int
                getattr(lib, name).argtypes = [C.POINTER(RefShadowFilterArgs)]
        if hasattr(lib, "ref_generate_gbuffer"):
            lib.ref_generate_gbuffer.restype = C.c_int
            lib.ref_generate_gbuffer.argtypes = [C.POINTER(RefGBufferArgs)]
        _lib = lib
    return _lib


def df_build(blocks):
    """World::GenerateDistanceField through the three compute shaders (one invocation per grid line, X, Y, Z)."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
    out = np.empty_like(blocks)
    load().ref_df_build(blocks.ctypes.data, out.ctypes.data)
    return out


def trace_primary(blocks, df, cam, params, table=None, alpha_mips=None):
    """InitialRayTraceFrag.glsl main() per pixel of rows [cam.row_begin, cam.row_end).  Returns the G-buffer in the ABI's terms:
    t (o_HitDistance), normal_id = round(o_Normal * 10) (10 = miss), block_id = round(o_BlockID * 255), inv_t (o_DepthNonLinear).
    params.alpha_test sets u_ShouldAlphaTest and needs the BlockDataSSBO table and the albedo alpha mip chain."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
    df = np.ascontiguousarray(df, dtype=np.uint8).reshape(-1)
    W, H = cam.width, cam.height
    planes = [np.zeros((H, W), dtype=np.float32) for _ in range(4)]
    iv = np.ascontiguousarray(np.frombuffer(cam.inv_view, dtype=np.float32))
    ip = np.ascontiguousarray(np.frombuffer(cam.inv_proj, dtype=np.float32))
    jit = np.array([params.jitter[0], params.jitter[1]], dtype=np.float32)
    if params.alpha_test:
        table = np.ascontiguousarray(table, dtype=np.int32)
        alpha_mips = np.ascontiguousarray(alpha_mips, dtype=np.uint8)
        load().ref_trace_primary_alpha(blocks.ctypes.data, df.ctypes.data, iv.ctypes.data, ip.ctypes.data, W, H, cam.row_begin, cam.row_end,
                                       params.max_iterations, params.jitter_enable, jit.ctypes.data, table.ctypes.data, alpha_mips.ctypes.data,
                                       alpha_mips.shape[0], params.fov_degrees, *[p.ctypes.data for p in planes])
    else:
        load().ref_trace_primary(blocks.ctypes.data, df.ctypes.data, iv.ctypes.data, ip.ctypes.data, W, H, cam.row_begin, cam.row_end,
                                 params.max_iterations, params.jitter_enable, jit.ctypes.data, *[p.ctypes.data for p in planes])
    t, n, b, it = planes
    return {"t": t, "normal_id": np.rint(n * np.float32(10.0)).astype(np.uint8), "block_id": np.rint(b * np.float32(255.0)).astype(np.uint8), "inv_t": it}


def trace_shadow(blocks, df, cam, gbuf, params, noise_rgba8, table=None, alpha_mips=None):
    """ShadowRayTraceFrag.glsl main() per pixel.  gbuf: t (fp32 plane as the position texture holds it) and normal_id planes.
    Returns shadow (o_Shadow as 0/1 bytes) and transversal (o_IntersectionTransversal)."""
    blocks = np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1)
    df = np.ascontiguousarray(df, dtype=np.uint8).reshape(-1)
    W, H = cam.width, cam.height
    t = np.ascontiguousarray(gbuf["t"], dtype=np.float32)
    nid = np.ascontiguousarray(gbuf["normal_id"], dtype=np.uint8)
    noise = np.ascontiguousarray(noise_rgba8, dtype=np.uint8).reshape(256, 256, 4)
    iv = np.ascontiguousarray(np.frombuffer(cam.inv_view, dtype=np.float32))
    ip = np.ascontiguousarray(np.frombuffer(cam.inv_proj, dtype=np.float32))
    light = np.array(list(params.light_dir), dtype=np.float32)
    halton = np.array(list(params.halton), dtype=np.float32)
    o_s, o_t = np.zeros((H, W), dtype=np.float32), np.zeros((H, W), dtype=np.float32)
    if params.alpha_test:
        table = np.ascontiguousarray(table, dtype=np.int32)
        alpha_mips = np.ascontiguousarray(alpha_mips, dtype=np.uint8)
        load().ref_trace_shadow_alpha(blocks.ctypes.data, df.ctypes.data, iv.ctypes.data, ip.ctypes.data, W, H, cam.row_begin, cam.row_end, t.ctypes.data,
                                      nid.ctypes.data, light.ctypes.data, params.frame, params.soft, halton.ctypes.data, noise.ctypes.data,
                                      table.ctypes.data, alpha_mips.ctypes.data, alpha_mips.shape[0], params.fov_degrees, o_s.ctypes.data, o_t.ctypes.data)
    else:
        load().ref_trace_shadow(blocks.ctypes.data, df.ctypes.data, iv.ctypes.data, ip.ctypes.data, W, H, cam.row_begin, cam.row_end, t.ctypes.data,
                                nid.ctypes.data, light.ctypes.data, params.frame, params.soft, halton.ctypes.data, noise.ctypes.data, o_s.ctypes.data,
                                o_t.ctypes.data)
    return {"shadow": (o_s > 0.5).astype(np.uint8), "transversal": o_t}


def trace_diffuse(blocks, df, cam, gbuf, params, materials, blue_noise, sky):
    """DiffuseRayTraceFrag.glsl main() per pixel (uniforms as Core/Pipeline.cpp:2174-2281 sets them).  materials: the assets.load_materials()
    dict (table + baked albedo L3 / PBR L2 / emissive L0 texel arrays), blue_noise: (sobol, scramble, rank), sky: [6][n][n][3]."""
    keep = []

    def ptr(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    W, H = cam.width, cam.height
    out = {"sh": np.zeros((H, W, 4), np.float32), "cocg": np.zeros((H, W, 2), np.float32), "luma": np.zeros((H, W), np.float32),
           "ao_sky": np.zeros((H, W, 2), np.float32)}
    a = RefDiffuseArgs()
    a.blocks, a.df = ptr(np.asarray(blocks).reshape(-1), np.uint8), ptr(np.asarray(df).reshape(-1), np.uint8)
    a.inv_view, a.inv_proj = ptr(np.frombuffer(cam.inv_view, dtype=np.float32), np.float32), ptr(np.frombuffer(cam.inv_proj, dtype=np.float32), np.float32)
    a.width, a.height, a.row_begin, a.row_end = W, H, cam.row_begin, cam.row_end
    a.g_t, a.g_normal_id = ptr(gbuf["t"], np.float32), ptr(gbuf["normal_id"], np.uint8)
    a.materials = ptr(materials["table"], np.int32)
    a.sobol, a.scramble, a.rank = (ptr(t, np.int32) for t in blue_noise)
    a.albedo_lod3, a.pbr_lod2, a.emissive_lod0 = ptr(materials["albedo_lod3"], np.float32), ptr(materials["pbr_lod2"], np.float32), ptr(materials["emissive_lod0"], np.float32)
    sky = np.ascontiguousarray(sky, dtype=np.float32)
    a.sky, a.sky_n = ptr(sky, np.float32), sky.shape[1]
    a.spp, a.checker_spp, a.checkerboard, a.trace_length = params.spp, params.checker_spp, params.checkerboard, params.trace_length
    a.frame, a.supersample = params.frame, params.supersample
    a.halton[:] = list(params.halton)
    a.sun_dir[:] = list(params.sun_dir)
    a.moon_dir[:] = list(params.moon_dir)
    a.sun_visibility, a.gi_sun_strength, a.gi_sky_strength, a.light_intensity = (params.sun_visibility, params.gi_sun_strength, params.gi_sky_strength,
                                                                                 params.light_intensity)
    a.o_sh, a.o_cocg, a.o_utility, a.o_ao_sky = (out[k].ctypes.data for k in ("sh", "cocg", "luma", "ao_sky"))
    load().ref_trace_diffuse(C.byref(a))
    return out


def trace_reflection(blocks, df, cam, gbuf, diffuse, params, g_normal, g_pbr, materials, blue_noise, sky):
    """ReflectionTraceFrag.glsl main() per pixel in the v1 parity profile (Core/Pipeline.cpp:3003-3164).  g_normal [H][W][3] and g_pbr
    [H][W][4] stand for the G-buffer material pass's outputs; diffuse: the GI pass's sh / cocg planes."""
    keep = []

    def ptr(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    W, H = cam.width, cam.height
    out = {"color": np.zeros((H, W, 4), np.float32), "hit_distance": np.zeros((H, W), np.float32), "emissive_mask": np.zeros((H, W), np.float32)}
    a = RefReflectionArgs()
    a.blocks, a.df = ptr(np.asarray(blocks).reshape(-1), np.uint8), ptr(np.asarray(df).reshape(-1), np.uint8)
    a.inv_view, a.inv_proj = ptr(np.frombuffer(cam.inv_view, dtype=np.float32), np.float32), ptr(np.frombuffer(cam.inv_proj, dtype=np.float32), np.float32)
    a.width, a.height, a.row_begin, a.row_end = W, H, cam.row_begin, cam.row_end
    a.g_t, a.g_normal_id, a.g_block_id = ptr(gbuf["t"], np.float32), ptr(gbuf["normal_id"], np.uint8), ptr(gbuf["block_id"], np.uint8)
    a.g_normal, a.g_pbr = ptr(g_normal, np.float32), ptr(g_pbr, np.float32)
    a.sh, a.cocg = ptr(diffuse["sh"], np.float32), ptr(diffuse["cocg"], np.float32)
    a.materials = ptr(materials["table"], np.int32)
    a.sobol, a.scramble, a.rank = (ptr(t, np.int32) for t in blue_noise)
    a.albedo_lod3, a.normal_lod3 = ptr(materials["albedo_lod3"], np.float32), ptr(materials["normal_lod3"], np.float32)
    a.pbr_lod2, a.emissive_lod2 = ptr(materials["pbr_lod2"], np.float32), ptr(materials["emissive_lod2"], np.float32)
    sky = np.ascontiguousarray(sky, dtype=np.float32)
    a.sky, a.sky_n = ptr(sky, np.float32), sky.shape[1]
    a.spp, a.trace_length, a.frame, a.rough = params.spp, params.trace_length, params.frame, params.rough
    a.roughness_bias, a.checkerboard = params.roughness_bias, params.checkerboard
    for name in ("sun_dir", "moon_dir", "stronger_dir", "viewer_pos", "halton", "grass_props"):
        getattr(a, name)[:] = list(getattr(params, name))
    a.sun_strength, a.moon_strength = params.sun_strength, params.moon_strength
    a.o_color, a.o_hit_distance, a.o_emissive_mask = (out[k].ctypes.data for k in ("color", "hit_distance", "emissive_mask"))
    load().ref_trace_reflection(C.byref(a))
    out["emissive_mask"] = (out["emissive_mask"] > 0.5).astype(np.uint8)
    return out


def generate_gbuffer(cam, gbuf, params, materials, mips, out=None, lava=None):
    """GenerateGBuffer.glsl main() per 2x2 quad of rows [cam.row_begin, cam.row_end) in the v1 parity profile (Core/Pipeline.cpp:2066-2136).
    mips = (albedo, normal, pbr) uint8 [layers][349525][4]."""
    keep = []

    def ptr(a, dt):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    W, H = cam.width, cam.height
    if out is None:
        out = {"albedo": np.zeros((H, W, 3), np.float32), "normal": np.zeros((H, W, 3), np.float32), "pbr": np.zeros((H, W, 4), np.float32),
               "texture_ao": np.zeros((H, W), np.float32)}
    a = RefGBufferArgs()
    a.inv_view, a.inv_proj = ptr(np.frombuffer(cam.inv_view, dtype=np.float32), np.float32), ptr(np.frombuffer(cam.inv_proj, dtype=np.float32), np.float32)
    a.width, a.height, a.row_begin, a.row_end = W, H, cam.row_begin, cam.row_end
    a.g_inv_t, a.g_normal_id, a.g_block_id = ptr(gbuf["inv_t"], np.float32), ptr(gbuf["normal_id"], np.uint8), ptr(gbuf["block_id"], np.uint8)
    a.materials = ptr(materials["table"], np.int32)
    a.albedo_mips, a.normal_mips, a.pbr_mips = (ptr(m, np.uint8) for m in mips)
    a.n_mip_layers = mips[0].shape[0]
    a.emissive_lod0 = ptr(materials["emissive_lod0"], np.float32)
    a.update_this_frame = params.update_this_frame
    a.grass_props[:] = list(params.grass_props)
    a.pom, a.high_quality_pom, a.dither_pom, a.frame = params.pom, params.high_quality_pom, params.dither_pom, params.frame
    a.pom_height, a.pom_exp = params.pom_height, params.pom_exp
    a.lava_block_id, a.time = params.lava_block_id, params.time
    if lava is not None:      # (albedo, normal) uint8 [8][256][256][4]
        a.lava_albedo, a.lava_normal = ptr(lava[0], np.uint8), ptr(lava[1], np.uint8)
    a.o_albedo, a.o_normal, a.o_pbr, a.o_texture_ao = (out[k].ctypes.data for k in ("albedo", "normal", "pbr", "texture_ao"))
    load().ref_generate_gbuffer(C.byref(a))
    return out


def ambient_sound(df, player_pos, frame):
    """EstimateAmbientSoundLevel.comp over the 8 x 4 invocations of Core/Pipeline.cpp:1921: (SkyLevelAggregate, per-invocation addends[32])."""
    df = np.ascontiguousarray(df, dtype=np.uint8).reshape(-1)
    p = np.array([float(v) for v in player_pos], dtype=np.float32)
    agg = C.c_uint32()
    per = np.zeros(32, np.uint32)
    load().ref_ambient_sound(df.ctypes.data, p.ctypes.data, int(frame), C.byref(agg), per.ctypes.data)
    return int(agg.value), per


# ---- SVGF denoiser: Core/Shaders/SVGF/*.glsl (oracle/ref_denoise_driver.cpp) ---------------------------------------------------------
def _svgf_args(cam, keep):
    def ptr(a, dt=np.float32):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    a = RefSvgfArgs()
    a.inv_view, a.inv_proj = ptr(np.frombuffer(cam.inv_view, dtype=np.float32)), ptr(np.frombuffer(cam.inv_proj, dtype=np.float32))
    a.width, a.height, a.row_begin, a.row_end = cam.width, cam.height, cam.row_begin, cam.row_end
    return a, ptr


def _svgf_out(cam, names):
    W, H = cam.width, cam.height
    shapes = {"sh": (H, W, 4), "cocg": (H, W, 2), "utility": (H, W, 3), "ao_sky": (H, W, 2), "variance": (H, W)}
    return {k: np.zeros(shapes[k], np.float32) for k in names}


def svgf_initial(cam, gbuf, diffuse, out=None):
    keep = []
    a, ptr = _svgf_args(cam, keep)
    if out is None:
        out = _svgf_out(cam, ("sh", "cocg", "ao_sky"))
        out["luma"] = np.zeros((cam.height, cam.width), np.float32)
    a.g_t, a.g_normal_id = ptr(gbuf["t"]), ptr(gbuf["normal_id"], np.uint8)
    a.sh, a.cocg, a.luma, a.ao_sky = (ptr(diffuse[k]) for k in ("sh", "cocg", "luma", "ao_sky"))
    a.o_sh, a.o_cocg, a.o_utility, a.o_ao_sky = (out[k].ctypes.data for k in ("sh", "cocg", "luma", "ao_sky"))
    load().ref_svgf_initial(C.byref(a))
    return out


def svgf_temporal(cam, gbuf, prev_gbuf, diffuse, prev_temporal, params, out=None):
    keep = []
    a, ptr = _svgf_args(cam, keep)
    out = _svgf_out(cam, ("sh", "cocg", "utility", "ao_sky")) if out is None else out
    a.g_t, a.g_normal_id, a.g_block_id = ptr(gbuf["t"]), ptr(gbuf["normal_id"], np.uint8), ptr(gbuf["block_id"], np.uint8)
    a.prev_t, a.prev_normal_id, a.prev_block_id = ptr(prev_gbuf["t"]), ptr(prev_gbuf["normal_id"], np.uint8), ptr(prev_gbuf["block_id"], np.uint8)
    a.sh, a.cocg, a.luma, a.ao_sky = (ptr(diffuse[k]) for k in ("sh", "cocg", "luma", "ao_sky"))
    a.prev_sh, a.prev_cocg, a.prev_utility, a.prev_ao_sky = (ptr(prev_temporal[k]) for k in ("sh", "cocg", "utility", "ao_sky"))
    a.prev_view, a.prev_projection = ptr(np.array(list(params.prev_view), np.float32)), ptr(np.array(list(params.prev_projection), np.float32))
    a.be_useful = params.be_useful
    a.o_sh, a.o_cocg, a.o_utility, a.o_ao_sky = (out[k].ctypes.data for k in ("sh", "cocg", "utility", "ao_sky"))
    load().ref_svgf_temporal(C.byref(a))
    return out


def svgf_variance(cam, gbuf, temporal, params, out=None):
    keep = []
    a, ptr = _svgf_args(cam, keep)
    out = _svgf_out(cam, ("sh", "cocg", "variance")) if out is None else out
    a.g_t, a.g_normal_id = ptr(gbuf["t"]), ptr(gbuf["normal_id"], np.uint8)
    a.sh, a.cocg, a.utility = (ptr(temporal[k]) for k in ("sh", "cocg", "utility"))
    a.do_spatial, a.aggressive = params.do_spatial, params.aggressive_disocclusion
    a.o_sh, a.o_cocg, a.o_variance = (out[k].ctypes.data for k in ("sh", "cocg", "variance"))
    load().ref_svgf_variance(C.byref(a))
    return out


def svgf_spatial(cam, gbuf, planes, temporal_utility, params, out=None):
    keep = []
    a, ptr = _svgf_args(cam, keep)
    out = _svgf_out(cam, ("sh", "cocg", "variance", "ao_sky")) if out is None else out
    a.g_t, a.g_normal_id, a.g_block_id = ptr(gbuf["t"]), ptr(gbuf["normal_id"], np.uint8), ptr(gbuf["block_id"], np.uint8)
    a.sh, a.cocg, a.variance, a.ao_sky = (ptr(planes[k]) for k in ("sh", "cocg", "variance", "ao_sky"))
    a.temporal_utility = ptr(temporal_utility)
    a.step, a.large_kernel, a.do_spatial, a.aggressive = params.step, params.large_kernel, params.do_spatial, params.aggressive_disocclusion
    a.color_phi_bias, a.time, a.resolution_scale = params.color_phi_bias, params.time, params.resolution_scale
    a.o_sh, a.o_cocg, a.o_variance, a.o_ao_sky = (out[k].ctypes.data for k in ("sh", "cocg", "variance", "ao_sky"))
    load().ref_svgf_spatial(C.byref(a))
    return out


# ---- sun-shadow filters: Core/Shaders/ShadowTemporalFilter.glsl, ShadowFilter.glsl (oracle/ref_denoise_driver.cpp) -------------------------
def _shadow_args(cam, keep):
    def ptr(a, dt=np.float32):
        a = np.ascontiguousarray(a, dtype=dt)
        keep.append(a)
        return a.ctypes.data

    a = RefShadowFilterArgs()
    a.inv_view, a.inv_proj = ptr(np.frombuffer(cam.inv_view, dtype=np.float32)), ptr(np.frombuffer(cam.inv_proj, dtype=np.float32))
    a.width, a.height, a.row_begin, a.row_end = cam.width, cam.height, cam.row_begin, cam.row_end
    return a, ptr


def shadow_temporal(cam, gbuf, prev_gbuf, shadow, prev_temporal, params, out=None):
    keep = []
    a, ptr = _shadow_args(cam, keep)
    if out is None:
        out = {"shadow": np.zeros((cam.height, cam.width), np.float32), "frames": np.zeros((cam.height, cam.width), np.float32)}
    a.g_t, a.g_normal_id, a.prev_t = ptr(gbuf["t"]), ptr(gbuf["normal_id"], np.uint8), ptr(prev_gbuf["t"])
    a.shadow_u8, a.transversal = ptr(shadow["shadow"], np.uint8), ptr(shadow["transversal"])
    a.prev_shadow, a.frames = ptr(prev_temporal["shadow"]), ptr(prev_temporal["frames"])
    a.prev_view, a.prev_projection = ptr(np.array(list(params.prev_view), np.float32)), ptr(np.array(list(params.prev_projection), np.float32))
    a.o_shadow, a.o_frames = out["shadow"].ctypes.data, out["frames"].ctypes.data
    load().ref_shadow_temporal(C.byref(a))
    return out


def shadow_filter(cam, gbuf, temporal, transversal, params, out=None):
    keep = []
    a, ptr = _shadow_args(cam, keep)
    out = np.zeros((cam.height, cam.width), np.float32) if out is None else out
    a.g_t, a.g_normal_id = ptr(gbuf["t"]), ptr(gbuf["normal_id"], np.uint8)
    a.shadow, a.transversal, a.frames = ptr(temporal["shadow"]), ptr(transversal), ptr(temporal["frames"])
    a.filter_scale = params.filter_scale
    a.o_shadow = out.ctypes.data
    load().ref_shadow_filter(C.byref(a))
    return out
