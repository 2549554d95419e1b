"""ctypes binding of the C ABI in include/vxpt.h.

The structures below are the ABI's POD types; tests also hand them to the CPU oracle (oracle/vxo.py), whose
entry points take the same structs, so one description of the camera / parameters drives both sides.
There is no fallback: if libvxpt.so cannot be loaded, `load()` raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvxpt.so")

WORLD_SIZE_X, WORLD_SIZE_Y, WORLD_SIZE_Z = 384, 128, 384
WORLD_VOXELS = WORLD_SIZE_X * WORLD_SIZE_Y * WORLD_SIZE_Z
NORMAL_MISS = 10
ALPHA_MIP_TEXELS = sum((512 >> k) ** 2 for k in range(9))  # 349,524: levels 0..8 of a 512^2 layer (VXPT_ALPHA_MIP_TEXELS)

LAVA_SIZE, LAVA_FRAMES = 256, 8
MIP_CHAIN_TEXELS = sum((512 >> k) ** 2 for k in range(10))  # 349,525: levels 0..9 (VXPT_MIP_CHAIN_TEXELS)

OK, E_INVALID, E_CUDA, E_NOMEM, E_STATE, E_UNSUPPORTED = 0, -1, -2, -3, -4, -5
OPT_TRAVERSAL_LAYOUT, OPT_GI_WAVEFRONT, OPT_DF_ALGO, OPT_SCENE_REPLICAS, OPT_TIMING_EVENTS, OPT_TEXEL_FORMAT = 1, 2, 3, 4, 5, 6
OPT_MATERIAL_QUAD_SHUFFLE = 7
OPT_REFLECTION_WAVEFRONT = 8
SHARED_HANDLE_BYTES = 64

u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)
i16p = C.POINTER(C.c_int16)
i32p = C.POINTER(C.c_int32)


class VxCamera(C.Structure):
    _fields_ = [("inv_view", C.c_float * 16), ("inv_proj", C.c_float * 16), ("width", C.c_int32), ("height", C.c_int32),
                ("row_begin", C.c_int32), ("row_end", C.c_int32),
                ("interleave_n", C.c_int32), ("interleave_rank", C.c_int32), ("band_rows", C.c_int32), ("reserved", C.c_int32)]


class VxPrimaryParams(C.Structure):
    _fields_ = [("max_iterations", C.c_int32), ("jitter_enable", C.c_int32), ("jitter", C.c_float * 2), ("alpha_test", C.c_int32),
                ("fov_degrees", C.c_float)]


class VxGBuffer(C.Structure):
    _fields_ = [("t", C.c_void_p), ("normal_id", C.c_void_p), ("block_id", C.c_void_p), ("inv_t", C.c_void_p), ("hit_voxel", C.c_void_p)]


class VxShadowParams(C.Structure):
    _fields_ = [("light_dir", C.c_float * 3), ("frame", C.c_int32), ("soft", C.c_int32), ("halton", C.c_float * 2), ("alpha_test", C.c_int32),
                ("fov_degrees", C.c_float)]


class VxShadowOut(C.Structure):
    _fields_ = [("shadow", C.c_void_p), ("transversal", C.c_void_p)]


class VxDiffuseParams(C.Structure):
    _fields_ = [("spp", C.c_int32), ("checker_spp", C.c_int32), ("checkerboard", C.c_int32), ("trace_length", C.c_int32),
                ("frame", C.c_int32), ("use_blue_noise", C.c_int32), ("supersample", C.c_int32), ("direct_sampling", C.c_int32),
                ("halton", C.c_float * 2), ("sun_dir", C.c_float * 3), ("moon_dir", C.c_float * 3), ("sun_visibility", C.c_float),
                ("gi_sun_strength", C.c_float), ("gi_sky_strength", C.c_float), ("light_intensity", C.c_float)]


class VxDiffuseOut(C.Structure):
    _fields_ = [("sh", C.c_void_p), ("cocg", C.c_void_p), ("luma", C.c_void_p), ("ao_sky", C.c_void_p)]


class VxReflectionParams(C.Structure):
    _fields_ = [("spp", C.c_int32), ("trace_length", C.c_int32), ("frame", C.c_int32), ("rough", C.c_int32), ("roughness_bias", C.c_int32),
                ("checkerboard", C.c_int32), ("sun_dir", C.c_float * 3), ("moon_dir", C.c_float * 3), ("stronger_dir", C.c_float * 3),
                ("viewer_pos", C.c_float * 3), ("sun_strength", C.c_float), ("moon_strength", C.c_float), ("halton", C.c_float * 2),
                ("grass_props", C.c_int32 * 10)]


class VxReflectionIn(C.Structure):
    _fields_ = [("g_normal", C.c_void_p), ("g_pbr", C.c_void_p), ("sh", C.c_void_p), ("cocg", C.c_void_p)]


class VxReflectionOut(C.Structure):
    _fields_ = [("color", C.c_void_p), ("hit_distance", C.c_void_p), ("emissive_mask", C.c_void_p)]


class VxMaterialParams(C.Structure):
    _fields_ = [("update_this_frame", C.c_int32), ("pom", C.c_int32), ("lava_block_id", C.c_int32), ("grass_props", C.c_int32 * 10),
                ("pom_height", C.c_float), ("pom_exp", C.c_float), ("high_quality_pom", C.c_int32), ("dither_pom", C.c_int32), ("frame", C.c_int32), ("time", C.c_float)]


class VxMaterialOut(C.Structure):
    _fields_ = [("albedo", C.c_void_p), ("normal", C.c_void_p), ("pbr", C.c_void_p), ("texture_ao", C.c_void_p)]


class VxSvgfInitialIn(C.Structure):
    _fields_ = [("current", VxGBuffer), ("sh", C.c_void_p), ("cocg", C.c_void_p), ("luma", C.c_void_p), ("ao_sky", C.c_void_p)]


class VxSvgfInitialOut(C.Structure):
    _fields_ = [("sh", C.c_void_p), ("cocg", C.c_void_p), ("luma", C.c_void_p), ("ao_sky", C.c_void_p)]


class VxSvgfTemporalIn(C.Structure):
    _fields_ = [("current", VxGBuffer), ("previous", VxGBuffer), ("sh", C.c_void_p), ("cocg", C.c_void_p), ("luma", C.c_void_p),
                ("ao_sky", C.c_void_p), ("prev_sh", C.c_void_p), ("prev_cocg", C.c_void_p), ("prev_utility", C.c_void_p), ("prev_ao_sky", C.c_void_p)]


class VxSvgfTemporalParams(C.Structure):
    _fields_ = [("prev_view", C.c_float * 16), ("prev_projection", C.c_float * 16), ("be_useful", C.c_int32)]


class VxSvgfTemporalOut(C.Structure):
    _fields_ = [("sh", C.c_void_p), ("cocg", C.c_void_p), ("utility", C.c_void_p), ("ao_sky", C.c_void_p)]


class VxSvgfVarianceIn(C.Structure):
    _fields_ = [("current", VxGBuffer), ("sh", C.c_void_p), ("cocg", C.c_void_p), ("utility", C.c_void_p)]


class VxSvgfVarianceParams(C.Structure):
    _fields_ = [("do_spatial", C.c_int32), ("aggressive_disocclusion", C.c_int32)]


class VxSvgfVarianceOut(C.Structure):
    _fields_ = [("sh", C.c_void_p), ("cocg", C.c_void_p), ("variance", C.c_void_p)]


class VxSvgfSpatialIn(C.Structure):
    _fields_ = [("current", VxGBuffer), ("sh", C.c_void_p), ("cocg", C.c_void_p), ("variance", C.c_void_p), ("ao_sky", C.c_void_p),
                ("temporal_utility", C.c_void_p)]


class VxSvgfSpatialParams(C.Structure):
    _fields_ = [("step", C.c_int32), ("large_kernel", C.c_int32), ("do_spatial", C.c_int32), ("aggressive_disocclusion", C.c_int32),
                ("color_phi_bias", C.c_float), ("time", C.c_float), ("resolution_scale", C.c_float)]


class VxSvgfSpatialOut(C.Structure):
    _fields_ = [("sh", C.c_void_p), ("cocg", C.c_void_p), ("variance", C.c_void_p), ("ao_sky", C.c_void_p)]


class VxSvgfFrameParams(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("projection", C.c_float * 16), ("reset_history", C.c_int32), ("pre_pass", C.c_int32), ("wide", C.c_int32),
                ("large_kernel", C.c_int32), ("aggressive_disocclusion", C.c_int32), ("color_phi_bias", C.c_float), ("time", C.c_float),
                ("resolution_scale", C.c_float)]


class VxShadowFrameParams(C.Structure):
    _fields_ = [("view", C.c_float * 16), ("projection", C.c_float * 16), ("reset_history", C.c_int32), ("spatial", C.c_int32), ("filter_scale", C.c_float)]


class VxShadowTemporalIn(C.Structure):
    _fields_ = [("current", VxGBuffer), ("previous", VxGBuffer), ("shadow", C.c_void_p), ("transversal", C.c_void_p), ("prev_shadow", C.c_void_p),
                ("prev_frames", C.c_void_p)]


class VxShadowTemporalParams(C.Structure):
    _fields_ = [("prev_view", C.c_float * 16), ("prev_projection", C.c_float * 16)]


class VxShadowTemporalOut(C.Structure):
    _fields_ = [("shadow", C.c_void_p), ("frames", C.c_void_p)]


class VxShadowFilterIn(C.Structure):
    _fields_ = [("current", VxGBuffer), ("shadow", C.c_void_p), ("transversal", C.c_void_p), ("frames", C.c_void_p)]


class VxShadowFilterParams(C.Structure):
    _fields_ = [("filter_scale", C.c_float)]


class VxFrameParams(C.Structure):
    _fields_ = [("primary", C.POINTER(VxPrimaryParams)), ("shadow", C.POINTER(VxShadowParams)), ("diffuse", C.POINTER(VxDiffuseParams)),
                ("reflection", C.POINTER(VxReflectionParams)), ("g_normal", C.c_void_p), ("g_pbr", C.c_void_p), ("material", C.POINTER(VxMaterialParams))]


class VxFrameOut(C.Structure):
    _fields_ = [("gbuffer", VxGBuffer), ("shadow", VxShadowOut), ("diffuse", VxDiffuseOut), ("reflection", VxReflectionOut), ("material", VxMaterialOut)]


class VxStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("df_fetches", C.c_uint64), ("vox_fetches", C.c_uint64), ("last_ms", C.c_float),
                ("df_build_ms", C.c_float), ("brick_pack_ms", C.c_float)]


# every export of include/vxpt.h: name -> (restype, argtypes)
EXPORTS = {
    "vxpt_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "vxpt_destroy": (C.c_int, [C.c_void_p]),
    "vxpt_last_error": (C.c_char_p, []),
    "vxpt_version": (C.c_char_p, []),
    "vxpt_upload_world": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxpt_set_block": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint8]),
    "vxpt_set_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "vxpt_download_world": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxpt_build_distance_field": (C.c_int, [C.c_void_p]),
    "vxpt_download_distance_field": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxpt_device_pointers": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]),
    "vxpt_set_materials": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxpt_set_blue_noise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vxpt_set_material_textures": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "vxpt_set_reflection_textures": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "vxpt_set_albedo_alpha_mips": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "vxpt_set_sky_cubemap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "vxpt_set_shadow_noise": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxpt_trace_primary": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxPrimaryParams), C.POINTER(VxGBuffer)]),
    "vxpt_trace_shadow": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxShadowParams), C.POINTER(VxShadowOut)]),
    "vxpt_trace_diffuse": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxDiffuseParams), C.POINTER(VxDiffuseOut)]),
    "vxpt_trace_reflection": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxReflectionIn),
                                        C.POINTER(VxReflectionParams), C.POINTER(VxReflectionOut)]),
    "vxpt_set_lava_textures": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "vxpt_set_gbuffer_textures": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "vxpt_generate_gbuffer": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxMaterialParams),
                                        C.POINTER(VxMaterialOut)]),
    "vxpt_svgf_initial": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxSvgfInitialIn), C.POINTER(VxSvgfInitialOut)]),
    "vxpt_svgf_temporal": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxSvgfTemporalIn), C.POINTER(VxSvgfTemporalParams),
                                     C.POINTER(VxSvgfTemporalOut)]),
    "vxpt_svgf_variance": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxSvgfVarianceIn), C.POINTER(VxSvgfVarianceParams),
                                     C.POINTER(VxSvgfVarianceOut)]),
    "vxpt_svgf_spatial": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxSvgfSpatialIn), C.POINTER(VxSvgfSpatialParams),
                                    C.POINTER(VxSvgfSpatialOut)]),
    "vxpt_svgf_frame": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxDiffuseOut), C.POINTER(VxSvgfFrameParams),
                                  C.POINTER(VxSvgfSpatialOut)]),
    "vxpt_shadow_filter_frame": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxGBuffer), C.POINTER(VxShadowOut), C.POINTER(VxShadowFrameParams),
                                           C.c_void_p]),
    "vxpt_shadow_temporal": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxShadowTemporalIn), C.POINTER(VxShadowTemporalParams),
                                       C.POINTER(VxShadowTemporalOut)]),
    "vxpt_shadow_filter": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxShadowFilterIn), C.POINTER(VxShadowFilterParams), C.c_void_p]),
    "vxpt_trace_rays": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vxpt_player_shadowed": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int)]),
    "vxpt_estimate_ambient_sound": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_uint32), C.c_void_p]),
    "vxpt_render_frame": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxFrameParams), C.POINTER(VxFrameOut)]),
    "vxpt_render_frame_async": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxFrameParams), C.POINTER(VxFrameOut)]),
    "vxpt_frame_wait": (C.c_int, [C.c_void_p]),
    "vxpt_shared_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "vxpt_shared_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "vxpt_shared_close": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxpt_copy_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "vxpt_signal": (C.c_int, [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]),
    "vxpt_signal_next": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vxpt_wait_next": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "vxpt_wait_all": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_uint32, C.c_int, C.c_void_p]),
    "vxpt_sync": (C.c_int, [C.c_void_p]),
    "vxpt_get_stats": (C.c_int, [C.c_void_p, C.POINTER(VxStats)]),
    "vxpt_reset_stats": (C.c_int, [C.c_void_p]),
    "vxpt_launch_count": (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64)]),
    "vxpt_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "vxpt_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vxpt_reserve": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.c_int, C.c_size_t]),
    "vxpt_measure_l2_sector_peak": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    # one frame over N devices from one host thread (csrc/mg.cu)
    "vxpt_mg_create": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_void_p)]),
    "vxpt_mg_destroy": (C.c_int, [C.c_void_p]),
    "vxpt_mg_size": (C.c_int, [C.c_void_p]),
    "vxpt_mg_device": (C.c_void_p, [C.c_void_p, C.c_int]),
    "vxpt_mg_upload_world": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxpt_mg_set_block": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_uint8]),
    "vxpt_mg_set_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "vxpt_mg_build_distance_field": (C.c_int, [C.c_void_p]),
    "vxpt_mg_set_materials": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxpt_mg_set_blue_noise": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "vxpt_mg_set_material_textures": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "vxpt_mg_set_reflection_textures": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "vxpt_mg_set_sky_cubemap": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "vxpt_mg_set_shadow_noise": (C.c_int, [C.c_void_p, C.c_void_p]),
    "vxpt_mg_set_option": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "vxpt_mg_slab": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "vxpt_mg_render_frame": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxFrameParams), C.POINTER(VxFrameOut)]),
    "vxpt_mg_render_frame_async": (C.c_int, [C.c_void_p, C.POINTER(VxCamera), C.POINTER(VxFrameParams), C.POINTER(VxFrameOut)]),
    "vxpt_mg_frame_wait": (C.c_int, [C.c_void_p]),
    "vxpt_mg_sync": (C.c_int, [C.c_void_p]),
    "vxpt_mg_get_stats": (C.c_int, [C.c_void_p, C.POINTER(VxStats)]),
    "vxpt_mg_reset_stats": (C.c_int, [C.c_void_p]),
}

_lib = None


class VxptError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"vxpt error {code}: {message}")
        self.code = code


def load():
    """dlopen libvxpt.so and bind every export.  Raises if the library is missing — no CPU fallback exists."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not built: run `python -m voxelpathtracer_b200.build` (requires nvcc)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc != OK:
        raise VxptError(rc, load().vxpt_last_error().decode("utf-8", "replace"))
    return rc
