"""Renderer: thin Python driver over the C ABI (include/vxpt.h).  Plumbing only — every computation happens in
libvxpt.so on the GPU.  Buffers may be numpy arrays (host: staged through the library, results complete on
return) or torch CUDA tensors (device: zero-copy, complete after sync())."""
import ctypes as C

import numpy as np

from . import abi
from .abi import (VxCamera, VxDiffuseOut, VxDiffuseParams, VxFrameOut, VxFrameParams, VxGBuffer, VxMaterialOut, VxMaterialParams, VxPrimaryParams, VxReflectionIn,
                  VxReflectionOut, VxReflectionParams, VxShadowOut, VxShadowParams, VxStats, check)

try:  # torch is optional plumbing: device buffers, streams, torch.distributed
    import torch
except Exception:  # pragma: no cover
    torch = None


def _ptr(buf):
    """Raw address of a numpy array or torch tensor (None -> NULL)."""
    if buf is None:
        return None
    if isinstance(buf, int):  # raw address (e.g. a virtual plane base inside a packed slab buffer)
        return buf
    if isinstance(buf, np.ndarray):
        if not buf.flags["C_CONTIGUOUS"]:
            raise ValueError("numpy buffers handed to vxpt must be C-contiguous")
        return buf.ctypes.data
    if torch is not None and isinstance(buf, torch.Tensor):
        if not buf.is_contiguous():
            raise ValueError("torch buffers handed to vxpt must be contiguous")
        return buf.data_ptr()
    raise TypeError(f"unsupported buffer type {type(buf)}")


def primary_params(max_iterations=350, jitter=None, alpha_test=False, fov_degrees=60.0):
    p = VxPrimaryParams()
    p.max_iterations = max_iterations
    p.alpha_test, p.fov_degrees = int(bool(alpha_test)), float(fov_degrees)
    p.jitter_enable = 0 if jitter is None else 1
    if jitter is not None:
        p.jitter[0], p.jitter[1] = float(jitter[0]), float(jitter[1])
    return p


def shadow_params(light_dir, frame=0, soft=True, halton=(0.0, 0.0), alpha_test=False, fov_degrees=60.0):
    p = VxShadowParams()
    p.alpha_test, p.fov_degrees = int(bool(alpha_test)), float(fov_degrees)
    p.light_dir[:] = [float(v) for v in light_dir]
    p.frame, p.soft = int(frame), int(bool(soft))
    p.halton[0], p.halton[1] = float(halton[0]), float(halton[1])
    return p


def diffuse_params(sun_dir, moon_dir, sun_visibility, spp=1, checker_spp=None, checkerboard=False, trace_length=48, frame=0,
                   gi_sun_strength=1.0, gi_sky_strength=1.125, light_intensity=1.25):
    """Defaults of Core/Pipeline.cpp:73-75,280-281."""
    p = VxDiffuseParams()
    p.spp = int(spp)
    p.checker_spp = int((spp + spp % 2) // 2 if checker_spp is None else checker_spp)
    p.checkerboard = int(bool(checkerboard))
    p.trace_length, p.frame = int(trace_length), int(frame)
    p.use_blue_noise, p.supersample, p.direct_sampling = 1, 0, 0
    p.sun_dir[:] = [float(v) for v in sun_dir]
    p.moon_dir[:] = [float(v) for v in moon_dir]
    p.sun_visibility = float(sun_visibility)
    p.gi_sun_strength, p.gi_sky_strength, p.light_intensity = float(gi_sun_strength), float(gi_sky_strength), float(light_intensity)
    return p


def reflection_params(sun_dir, moon_dir, stronger_dir, viewer_pos, grass_props, spp=2, trace_length=64, frame=0, rough=True, roughness_bias=True,
                      checkerboard=False, sun_strength=0.85, moon_strength=1.0, halton=(0.0, 0.0)):
    """Defaults of Core/Pipeline.cpp:100-104,114,124,278-279."""
    p = VxReflectionParams()
    p.spp, p.trace_length, p.frame = int(spp), int(trace_length), int(frame)
    p.rough, p.roughness_bias, p.checkerboard = int(bool(rough)), int(bool(roughness_bias)), int(bool(checkerboard))
    p.sun_dir[:] = [float(v) for v in sun_dir]
    p.moon_dir[:] = [float(v) for v in moon_dir]
    p.stronger_dir[:] = [float(v) for v in stronger_dir]
    p.viewer_pos[:] = [float(v) for v in viewer_pos]
    p.sun_strength, p.moon_strength = float(sun_strength), float(moon_strength)
    p.halton[0], p.halton[1] = float(halton[0]), float(halton[1])
    p.grass_props[:] = [int(v) for v in grass_props]
    return p


def material_params(grass_props, update_this_frame=True, pom=False, lava_block_id=-1, pom_height=1.0, pom_exp=1.0, high_quality_pom=False,
                    dither_pom=True, frame=0, time=0.0):
    """GenerateGBuffer's uniforms (Core/Pipeline.cpp:2079-2104, defaults :264-268); the lava animation is outside the v1 parity profile."""
    p = VxMaterialParams()
    p.update_this_frame, p.pom, p.lava_block_id = int(bool(update_this_frame)), int(bool(pom)), int(lava_block_id)
    p.grass_props[:] = [int(v) for v in grass_props]
    p.pom_height, p.pom_exp = float(pom_height), float(pom_exp)
    p.high_quality_pom, p.dither_pom, p.frame = int(bool(high_quality_pom)), int(bool(dither_pom)), int(frame)
    p.time = float(time)
    return p


class Renderer:
    """One handle = one GPU = one caller thread (like the reference's GL context)."""

    def __init__(self, device=0):
        self.lib = abi.load()
        self.handle = C.c_void_p()
        check(self.lib.vxpt_create(int(device), C.byref(self.handle)))
        self.device = int(device)
        self._keep = []  # host arrays that must outlive an enqueued copy

    @classmethod
    def borrowed(cls, handle, device):
        """A Renderer over a handle somebody else owns (MultiRenderer.device(k)): every call works, close() does not destroy it."""
        r = cls.__new__(cls)
        r.lib, r.handle, r.device, r._keep, r._borrowed = abi.load(), C.c_void_p(handle), int(device), [], True
        return r

    def close(self):
        if self.handle and not getattr(self, "_borrowed", False):
            self.lib.vxpt_destroy(self.handle)
        self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- world / distance field ---------------------------------------------------------------------------
    def upload_world(self, blocks):
        """vxpt_upload_world copies exactly WORLD_VOXELS bytes from the address it is given: arrays and tensors are checked here (a raw
        address is the caller's promise)."""
        data = blocks.data if hasattr(blocks, "zyx") else blocks
        if isinstance(data, np.ndarray):
            if data.dtype != np.uint8 or data.size != abi.WORLD_VOXELS:
                raise ValueError(f"the world is {abi.WORLD_VOXELS} uint8 block ids (x + 384 * (y + 128 * z)); got {data.dtype} x {data.size}")
        elif torch is not None and isinstance(data, torch.Tensor):
            if data.dtype != torch.uint8 or data.numel() != abi.WORLD_VOXELS:
                raise ValueError(f"the world is {abi.WORLD_VOXELS} uint8 block ids; got {data.dtype} x {data.numel()}")
        check(self.lib.vxpt_upload_world(self.handle, _ptr(data)))

    def set_block(self, x, y, z, block):
        check(self.lib.vxpt_set_block(self.handle, int(x), int(y), int(z), int(block)))

    def set_blocks(self, xyz, ids):
        xyz = np.ascontiguousarray(xyz, dtype=np.int16).reshape(-1, 3)
        ids = np.ascontiguousarray(ids, dtype=np.uint8).reshape(-1)
        check(self.lib.vxpt_set_blocks(self.handle, _ptr(xyz), _ptr(ids), int(ids.size)))

    def build_distance_field(self):
        check(self.lib.vxpt_build_distance_field(self.handle))

    def download_distance_field(self):
        out = np.empty(abi.WORLD_VOXELS, dtype=np.uint8)
        check(self.lib.vxpt_download_distance_field(self.handle, _ptr(out)))
        return out

    def download_world(self):
        out = np.empty(abi.WORLD_VOXELS, dtype=np.uint8)
        check(self.lib.vxpt_download_world(self.handle, _ptr(out)))
        return out

    # ---- tables -------------------------------------------------------------------------------------------
    def set_materials(self, table):
        t = np.ascontiguousarray(table, dtype=np.int32).reshape(768)
        check(self.lib.vxpt_set_materials(self.handle, _ptr(t)))

    def set_blue_noise(self, sobol, scramble, rank):
        a, b, c = (np.ascontiguousarray(x, dtype=np.int32) for x in (sobol, scramble, rank))
        check(self.lib.vxpt_set_blue_noise(self.handle, _ptr(a), _ptr(b), _ptr(c)))

    def set_material_textures(self, albedo_lod3, pbr_lod2, emissive_lod0):
        a = np.ascontiguousarray(albedo_lod3, dtype=np.float32)
        p = np.ascontiguousarray(pbr_lod2, dtype=np.float32)
        e = np.ascontiguousarray(emissive_lod0, dtype=np.float32)
        assert a.shape[1:] == (64, 64, 4) and p.shape[1:] == (128, 128, 4) and a.shape[0] == p.shape[0]
        check(self.lib.vxpt_set_material_textures(self.handle, _ptr(a), _ptr(p), int(a.shape[0]), _ptr(e) if e.shape[0] else None, int(e.shape[0])))

    def set_reflection_textures(self, normal_lod3, emissive_lod2):
        n = np.ascontiguousarray(normal_lod3, dtype=np.float32)
        e = np.ascontiguousarray(emissive_lod2, dtype=np.float32)
        assert n.shape[1:] == (64, 64, 4) and (e.shape[0] == 0 or e.shape[1:] == (128, 128))
        check(self.lib.vxpt_set_reflection_textures(self.handle, _ptr(n), int(n.shape[0]), _ptr(e) if e.shape[0] else None, int(e.shape[0])))

    def set_albedo_alpha_mips(self, alpha_mips):
        """uint8 [n_layers][ALPHA_MIP_TEXELS]: alpha of the albedo array's mip levels 0..8 (see assets.alpha_mip_pyramid)."""
        a = np.ascontiguousarray(alpha_mips, dtype=np.uint8)
        assert a.ndim == 2 and a.shape[1] == abi.ALPHA_MIP_TEXELS, a.shape
        abi.check(self.lib.vxpt_set_albedo_alpha_mips(self.handle, a.ctypes.data, a.shape[0]))

    def set_lava_textures(self, albedo_rgba8, normal_rgba8):
        """uint8 [LAVA_FRAMES][LAVA_SIZE][LAVA_SIZE][4] each: the animated lava textures (Core/AnimatedTexture.cpp)."""
        a, n = (np.ascontiguousarray(x, dtype=np.uint8) for x in (albedo_rgba8, normal_rgba8))
        assert a.shape == n.shape == (abi.LAVA_FRAMES, abi.LAVA_SIZE, abi.LAVA_SIZE, 4), a.shape
        check(self.lib.vxpt_set_lava_textures(self.handle, _ptr(a), _ptr(n)))

    def set_gbuffer_textures(self, albedo_mips, normal_mips, pbr_mips):
        """uint8 [n_layers][MIP_CHAIN_TEXELS][4] each: the RGBA8 mip chains of the block arrays (see assets.rgba_mip_chain)."""
        arrs = [np.ascontiguousarray(a, dtype=np.uint8) for a in (albedo_mips, normal_mips, pbr_mips)]
        for a in arrs:
            assert a.ndim == 3 and a.shape[1:] == (abi.MIP_CHAIN_TEXELS, 4) and a.shape[0] == arrs[0].shape[0], a.shape
        check(self.lib.vxpt_set_gbuffer_textures(self.handle, _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]), int(arrs[0].shape[0])))

    def set_sky_cubemap(self, rgb):
        s = np.ascontiguousarray(rgb, dtype=np.float32)
        assert s.ndim == 4 and s.shape[0] == 6 and s.shape[1] == s.shape[2] and s.shape[3] == 3
        check(self.lib.vxpt_set_sky_cubemap(self.handle, _ptr(s), int(s.shape[1])))

    def set_shadow_noise(self, rgba8):
        n = np.ascontiguousarray(rgba8, dtype=np.uint8).reshape(256, 256, 4)
        check(self.lib.vxpt_set_shadow_noise(self.handle, _ptr(n)))

    def load_scene_tables(self, materials, blue_noise, sky, shadow_noise):
        self.set_materials(materials["table"])
        self.set_material_textures(materials["albedo_lod3"], materials["pbr_lod2"], materials["emissive_lod0"])
        if "normal_lod3" in materials:
            self.set_reflection_textures(materials["normal_lod3"], materials["emissive_lod2"])
        self.set_blue_noise(*blue_noise)
        self.set_sky_cubemap(sky)
        self.set_shadow_noise(shadow_noise)

    # ---- buffers ------------------------------------------------------------------------------------------
    def alloc(self, shape, dtype, device=False, pinned=False):
        if device or pinned:
            if torch is None:
                raise RuntimeError("device / pinned buffers need torch")
            tdt = {np.float32: torch.float32, np.float16: torch.float16, np.uint8: torch.uint8, np.int16: torch.int16}[dtype]
            if device:
                return torch.empty(shape, dtype=tdt, device=f"cuda:{self.device}")
            return torch.empty(shape, dtype=tdt).pin_memory().numpy()
        return np.empty(shape, dtype=dtype)

    # plane shapes / dtypes in the two texel formats of VXPT_OPT_TEXEL_FORMAT (False: fp32 planes, True: the reference's FBO formats)
    def alloc_gbuffer(self, width, height, device=False, hit_voxel=False, texel=False, pinned=False):
        f = np.float16 if texel else np.float32
        g = {"t": self.alloc((height, width), f, device, pinned), "normal_id": self.alloc((height, width), np.uint8, device, pinned),
             "block_id": self.alloc((height, width), np.uint8, device, pinned), "inv_t": self.alloc((height, width), np.float32, device, pinned)}
        if hit_voxel:
            g["hit_voxel"] = self.alloc((height, width, 3), np.int16, device, pinned)
        return g

    def alloc_shadow(self, width, height, device=False, texel=False, pinned=False):
        f = np.float16 if texel else np.float32
        return {"shadow": self.alloc((height, width), np.uint8, device, pinned), "transversal": self.alloc((height, width), f, device, pinned)}

    def alloc_reflection(self, width, height, device=False, texel=False, pinned=False):
        f = np.float16 if texel else np.float32
        return {"color": self.alloc((height, width, 4), f, device, pinned), "hit_distance": self.alloc((height, width), f, device, pinned),
                "emissive_mask": self.alloc((height, width), np.uint8, device, pinned)}

    def alloc_diffuse(self, width, height, device=False, texel=False, pinned=False):
        f = np.float16 if texel else np.float32
        return {"sh": self.alloc((height, width, 4), f, device, pinned), "cocg": self.alloc((height, width, 2), f, device, pinned),
                "luma": self.alloc((height, width), f, device, pinned),
                "ao_sky": self.alloc((height, width, 2), np.uint8 if texel else np.float32, device, pinned)}

    def alloc_material(self, width, height, device=False, pinned=False):
        return {"albedo": self.alloc((height, width, 3), np.float32, device, pinned), "normal": self.alloc((height, width, 3), np.float32, device, pinned),
                "pbr": self.alloc((height, width, 4), np.float32, device, pinned), "texture_ao": self.alloc((height, width), np.float32, device, pinned)}

    @staticmethod
    def gbuffer_struct(g):
        s = VxGBuffer()
        s.t, s.normal_id, s.block_id = _ptr(g.get("t")), _ptr(g.get("normal_id")), _ptr(g.get("block_id"))
        s.inv_t, s.hit_voxel = _ptr(g.get("inv_t")), _ptr(g.get("hit_voxel"))
        return s

    # ---- passes -------------------------------------------------------------------------------------------
    def trace_primary(self, cam, params, gbuf):
        s = self.gbuffer_struct(gbuf)
        check(self.lib.vxpt_trace_primary(self.handle, C.byref(cam), C.byref(params), C.byref(s)))
        return gbuf

    def trace_shadow(self, cam, gbuf, params, out):
        g = self.gbuffer_struct(gbuf)
        o = VxShadowOut()
        o.shadow, o.transversal = _ptr(out.get("shadow")), _ptr(out.get("transversal"))
        check(self.lib.vxpt_trace_shadow(self.handle, C.byref(cam), C.byref(g), C.byref(params), C.byref(o)))
        return out

    def trace_diffuse(self, cam, gbuf, params, out):
        g = self.gbuffer_struct(gbuf)
        o = VxDiffuseOut()
        o.sh, o.cocg, o.luma, o.ao_sky = _ptr(out.get("sh")), _ptr(out.get("cocg")), _ptr(out.get("luma")), _ptr(out.get("ao_sky"))
        check(self.lib.vxpt_trace_diffuse(self.handle, C.byref(cam), C.byref(g), C.byref(params), C.byref(o)))
        return out

    def trace_reflection(self, cam, gbuf, diffuse, params, out, g_normal=None, g_pbr=None):
        g = self.gbuffer_struct(gbuf)
        i = VxReflectionIn()
        i.g_normal, i.g_pbr, i.sh, i.cocg = _ptr(g_normal), _ptr(g_pbr), _ptr(diffuse.get("sh")), _ptr(diffuse.get("cocg"))
        o = VxReflectionOut()
        o.color, o.hit_distance, o.emissive_mask = _ptr(out.get("color")), _ptr(out.get("hit_distance")), _ptr(out.get("emissive_mask"))
        check(self.lib.vxpt_trace_reflection(self.handle, C.byref(cam), C.byref(g), C.byref(i), C.byref(params), C.byref(o)))
        return out

    def generate_gbuffer(self, cam, gbuf, params, out):
        """G-buffer material pass (vxpt_generate_gbuffer, GenerateGBuffer.glsl): albedo / normal / pbr / texture_ao planes; the normal
        and pbr planes are what trace_reflection takes as g_normal / g_pbr."""
        g = self.gbuffer_struct(gbuf)
        o = VxMaterialOut()
        o.albedo, o.normal, o.pbr, o.texture_ao = _ptr(out.get("albedo")), _ptr(out.get("normal")), _ptr(out.get("pbr")), _ptr(out.get("texture_ao"))
        check(self.lib.vxpt_generate_gbuffer(self.handle, C.byref(cam), C.byref(g), C.byref(params), C.byref(o)))
        return out

    # ---- SVGF denoiser (include/vxpt.h vxpt_svgf_*): full-frame fp32 planes, numpy or torch ----
    def alloc_denoise(self, width, height, names, device=False, pinned=False):
        from . import denoise
        shapes = denoise.plane_shapes(width, height)
        return {k: self.alloc(shapes[k], np.float32, device, pinned) for k in names}

    def svgf_initial(self, cam, gbuf, diffuse, out):
        """Spatial3x3Initial.glsl (vxpt_svgf_initial): the pre-temporal 3x3 pass over this frame's raw GI planes."""
        from . import denoise
        i, o = denoise.initial_structs(gbuf, diffuse, out, _ptr)
        check(self.lib.vxpt_svgf_initial(self.handle, C.byref(cam), C.byref(i), C.byref(o)))
        return out

    def svgf_temporal(self, cam, gbuf, prev_gbuf, diffuse, prev_temporal, params, out):
        from . import denoise
        i, o = denoise.temporal_structs(gbuf, prev_gbuf, diffuse, prev_temporal, out, _ptr)
        check(self.lib.vxpt_svgf_temporal(self.handle, C.byref(cam), C.byref(i), C.byref(params), C.byref(o)))
        return out

    def svgf_variance(self, cam, gbuf, temporal, params, out):
        from . import denoise
        i, o = denoise.variance_structs(gbuf, temporal, out, _ptr)
        check(self.lib.vxpt_svgf_variance(self.handle, C.byref(cam), C.byref(i), C.byref(params), C.byref(o)))
        return out

    def svgf_spatial(self, cam, gbuf, planes, temporal_utility, params, out):
        from . import denoise
        i, o = denoise.spatial_structs(gbuf, planes, temporal_utility, out, _ptr)
        check(self.lib.vxpt_svgf_spatial(self.handle, C.byref(cam), C.byref(i), C.byref(params), C.byref(o)))
        return out

    def svgf_frame(self, cam, gbuf, diffuse, params, out):
        """The whole SVGF chain of one frame with the intermediates and the history resident on the device (vxpt_svgf_frame):
        diffuse = the GI pass's planes, out = {"sh", "cocg", "variance", "ao_sky"}; params from denoise.frame_params."""
        from .abi import VxSvgfSpatialOut
        g = self.gbuffer_struct(gbuf)
        d = VxDiffuseOut()
        d.sh, d.cocg, d.luma, d.ao_sky = _ptr(diffuse.get("sh")), _ptr(diffuse.get("cocg")), _ptr(diffuse.get("luma")), _ptr(diffuse.get("ao_sky"))
        o = VxSvgfSpatialOut()
        o.sh, o.cocg, o.variance, o.ao_sky = _ptr(out.get("sh")), _ptr(out.get("cocg")), _ptr(out.get("variance")), _ptr(out.get("ao_sky"))
        check(self.lib.vxpt_svgf_frame(self.handle, C.byref(cam), C.byref(g), C.byref(d), C.byref(params), C.byref(o)))
        return out

    def shadow_filter_frame(self, cam, gbuf, shadow, params, out):
        """Both shadow filters of one frame with the temporal planes and history resident on the device (vxpt_shadow_filter_frame)."""
        g = self.gbuffer_struct(gbuf)
        s = VxShadowOut()
        s.shadow, s.transversal = _ptr(shadow.get("shadow")), _ptr(shadow.get("transversal"))
        check(self.lib.vxpt_shadow_filter_frame(self.handle, C.byref(cam), C.byref(g), C.byref(s), C.byref(params), _ptr(out)))
        return out

    def shadow_temporal(self, cam, gbuf, prev_gbuf, shadow, prev_temporal, params, out):
        """ShadowTemporalFilter.glsl (vxpt_shadow_temporal): shadow = the shadow pass's planes, prev_temporal / out = {"shadow", "frames"}."""
        from . import denoise
        i, o = denoise.shadow_temporal_structs(gbuf, prev_gbuf, shadow, prev_temporal, out, _ptr)
        check(self.lib.vxpt_shadow_temporal(self.handle, C.byref(cam), C.byref(i), C.byref(params), C.byref(o)))
        return out

    def shadow_filter(self, cam, gbuf, temporal, transversal, params, out):
        """ShadowFilter.glsl (vxpt_shadow_filter): out = one fp32 plane."""
        from . import denoise
        i = denoise.shadow_filter_struct(gbuf, temporal, transversal, _ptr)
        check(self.lib.vxpt_shadow_filter(self.handle, C.byref(cam), C.byref(i), C.byref(params), _ptr(out)))
        return out

    def svgf_denoise(self, cam, gbuf, prev_gbuf, diffuse, prev_temporal, temporal_params, time=0.0, steps=None, device=False, pre_pass=True):
        """The reference's whole SVGF chain for one frame (Core/Pipeline.cpp:2284-2596): pre-temporal 3x3 pass (PreTemporalSpatialPass,
        on by default) -> temporal -> variance -> five a-trous passes ping-ponging between two plane sets.  Returns (denoised planes,
        temporal planes to hand in as prev_temporal next frame)."""
        from . import denoise
        W, H = cam.width, cam.height
        if pre_pass:
            diffuse = self.svgf_initial(cam, gbuf, diffuse, self.alloc_denoise(W, H, ("sh", "cocg", "luma", "ao_sky"), device))
        temporal = self.svgf_temporal(cam, gbuf, prev_gbuf, diffuse, prev_temporal, temporal_params,
                                      self.alloc_denoise(W, H, ("sh", "cocg", "utility", "ao_sky"), device))
        var = self.svgf_variance(cam, gbuf, temporal, denoise.variance_params(), self.alloc_denoise(W, H, ("sh", "cocg", "variance"), device))
        cur = {"sh": var["sh"], "cocg": var["cocg"], "variance": var["variance"], "ao_sky": temporal["ao_sky"]}
        pong = [self.alloc_denoise(W, H, ("sh", "cocg", "variance", "ao_sky"), device) for _ in range(2)]
        for n, step in enumerate(steps or denoise.ATROUS_STEPS):
            cur = self.svgf_spatial(cam, gbuf, cur, temporal["utility"], denoise.spatial_params(step, time=time), pong[n % 2])
        return cur, temporal

    # ---- other consumers of the distance field ----
    def trace_rays(self, origins, directions, max_iterations=350, hit_voxel=True):
        """A batch of VoxelTraversalDF calls (vxpt_trace_rays): origins / directions [n][3] float32 -> dict of t, normal_id, block_id(, hit_voxel)."""
        o = np.ascontiguousarray(origins, dtype=np.float32).reshape(-1, 3)
        d = np.ascontiguousarray(directions, dtype=np.float32).reshape(-1, 3)
        assert o.shape == d.shape
        n = o.shape[0]
        out = {"t": np.zeros(n, np.float32), "normal_id": np.zeros(n, np.uint8), "block_id": np.zeros(n, np.uint8)}
        if hit_voxel:
            out["hit_voxel"] = np.zeros((n, 3), np.int16)
        abi.check(self.lib.vxpt_trace_rays(self.handle, o.ctypes.data, d.ctypes.data, n, int(max_iterations), out["t"].ctypes.data,
                                           out["normal_id"].ctypes.data, out["block_id"].ctypes.data,
                                           out["hit_voxel"].ctypes.data if hit_voxel else None))
        return out

    def player_shadowed(self, camera_pos, sun_dir):
        """PostProcessingVert.glsl:46-53: is the player's eye in the sun's shadow (v_PlayerShadowed)."""
        p = (C.c_float * 3)(*[float(v) for v in camera_pos])
        s = (C.c_float * 3)(*[float(v) for v in sun_dir])
        out = C.c_int(0)
        abi.check(self.lib.vxpt_player_shadowed(self.handle, p, s, C.byref(out)))
        return bool(out.value)

    def estimate_ambient_sound(self, player_pos, frame):
        """EstimateAmbientSoundLevel.comp (Pipeline.cpp:1908-1921): (SkyLevelAggregate, per-invocation addends uint32[32])."""
        p = (C.c_float * 3)(*[float(v) for v in player_pos])
        agg = C.c_uint32(0)
        per = np.zeros(32, np.uint32)
        abi.check(self.lib.vxpt_estimate_ambient_sound(self.handle, p, int(frame), C.byref(agg), per.ctypes.data))
        return int(agg.value), per

    def render_frame(self, cam, primary, shadow=None, diffuse=None, gbuf=None, shadow_out=None, diffuse_out=None, reflection=None,
                     reflection_out=None, g_normal=None, g_pbr=None, wait=True, material=None, material_out=None):
        """One frame of the path (vxpt_render_frame): primary -> shadow -> GI (-> reflections) with the G-buffer resident on
        the device; host planes are copied out pass by pass while later passes trace.  wait=False: vxpt_render_frame_async —
        returns once enqueued, host planes complete after frame_wait()."""
        fp = VxFrameParams()
        fp.primary = C.pointer(primary)
        if shadow is not None:
            fp.shadow = C.pointer(shadow)
        if diffuse is not None:
            fp.diffuse = C.pointer(diffuse)
        if reflection is not None:
            fp.reflection = C.pointer(reflection)
        fp.g_normal, fp.g_pbr = _ptr(g_normal), _ptr(g_pbr)
        if material is not None:     # the G-buffer material pass right after the primary pass; it feeds the reflection pass
            fp.material = C.pointer(material)
        fo = VxFrameOut()
        mo = material_out or {}
        fo.material.albedo, fo.material.normal, fo.material.pbr, fo.material.texture_ao = (_ptr(mo.get("albedo")), _ptr(mo.get("normal")), _ptr(mo.get("pbr")),
                                                                                            _ptr(mo.get("texture_ao")))
        fo.gbuffer = self.gbuffer_struct(gbuf or {})
        so, do, ro = shadow_out or {}, diffuse_out or {}, reflection_out or {}
        fo.shadow.shadow, fo.shadow.transversal = _ptr(so.get("shadow")), _ptr(so.get("transversal"))
        fo.diffuse.sh, fo.diffuse.cocg, fo.diffuse.luma, fo.diffuse.ao_sky = (_ptr(do.get("sh")), _ptr(do.get("cocg")), _ptr(do.get("luma")),
                                                                                _ptr(do.get("ao_sky")))
        fo.reflection.color, fo.reflection.hit_distance, fo.reflection.emissive_mask = (_ptr(ro.get("color")), _ptr(ro.get("hit_distance")),
                                                                                         _ptr(ro.get("emissive_mask")))
        fn = self.lib.vxpt_render_frame if wait else self.lib.vxpt_render_frame_async
        check(fn(self.handle, C.byref(cam), C.byref(fp), C.byref(fo)))
        return gbuf, shadow_out, diffuse_out, reflection_out

    def frame_wait(self):
        check(self.lib.vxpt_frame_wait(self.handle))

    def prepare_frame(self, cam, gbuf=None, shadow_out=None, diffuse_out=None, reflection_out=None):
        """A frame call with everything that does not change from frame to frame built ONCE: the camera, the output-plane struct and the
        ctypes argument references.  Returns submit(primary, shadow, diffuse, reflection=None) -> vxpt_render_frame_async; per frame the
        host then does three pointer stores and one foreign call instead of filling two structs field by field (bench.py e2e: about 0.3 ms
        of Python per frame before, r01)."""
        fo = VxFrameOut()
        fo.gbuffer = self.gbuffer_struct(gbuf or {})
        so, do, ro = shadow_out or {}, diffuse_out or {}, reflection_out or {}
        fo.shadow.shadow, fo.shadow.transversal = _ptr(so.get("shadow")), _ptr(so.get("transversal"))
        fo.diffuse.sh, fo.diffuse.cocg, fo.diffuse.luma, fo.diffuse.ao_sky = (_ptr(do.get("sh")), _ptr(do.get("cocg")), _ptr(do.get("luma")),
                                                                                _ptr(do.get("ao_sky")))
        fo.reflection.color, fo.reflection.hit_distance, fo.reflection.emissive_mask = (_ptr(ro.get("color")), _ptr(ro.get("hit_distance")),
                                                                                         _ptr(ro.get("emissive_mask")))
        fp = VxFrameParams()
        cam_ref, fp_ref, fo_ref = C.byref(cam), C.byref(fp), C.byref(fo)
        call, handle = self.lib.vxpt_render_frame_async, self.handle
        keep = (cam, fo, fp, gbuf, shadow_out, diffuse_out, reflection_out)

        def submit(primary, shadow=None, diffuse=None, reflection=None, _keep=keep):
            fp.primary = C.pointer(primary)
            fp.shadow = C.pointer(shadow) if shadow is not None else None
            fp.diffuse = C.pointer(diffuse) if diffuse is not None else None
            fp.reflection = C.pointer(reflection) if reflection is not None else None
            rc = call(handle, cam_ref, fp_ref, fo_ref)
            if rc:
                check(rc)
        return submit

    # ---- peer-to-peer slab gather (multi-GPU) --------------------------------------------------------------
    def shared_alloc(self, nbytes):
        """Device buffer other processes can map (returns address, 64-byte handle)."""
        ptr = C.c_void_p()
        handle = (C.c_uint8 * abi.SHARED_HANDLE_BYTES)()
        check(self.lib.vxpt_shared_alloc(self.handle, int(nbytes), C.byref(ptr), handle))
        return ptr.value, bytes(handle)

    def shared_open(self, handle):
        ptr = C.c_void_p()
        buf = (C.c_uint8 * abi.SHARED_HANDLE_BYTES).from_buffer_copy(handle)
        check(self.lib.vxpt_shared_open(self.handle, buf, C.byref(ptr)))
        return ptr.value

    def shared_close(self, ptr):
        check(self.lib.vxpt_shared_close(self.handle, C.c_void_p(ptr)))

    def copy_async(self, dst_ptr, src_ptr, nbytes, stream=None):
        check(self.lib.vxpt_copy_async(self.handle, C.c_void_p(dst_ptr), C.c_void_p(src_ptr), int(nbytes), C.c_void_p(stream)))

    def signal(self, flag_ptr, value, stream=None):
        """Stream-ordered system-scope release store of `value` to a (possibly peer-mapped) flag word."""
        check(self.lib.vxpt_signal(self.handle, C.c_void_p(flag_ptr), int(value) & 0xFFFFFFFF, C.c_void_p(stream)))

    def signal_next(self, flag_ptr, counter_ptr, stream=None):
        """Graph-capturable signal: the value is ++*counter (device-resident)."""
        check(self.lib.vxpt_signal_next(self.handle, C.c_void_p(flag_ptr), C.c_void_p(counter_ptr), C.c_void_p(stream)))

    def wait_next(self, flags_ptr, n, stride_words, counter_ptr, lag=0, timeout_ms=2000, stream=None):
        """Graph-capturable wait: target = ++*counter - lag; waits (if target >= 1) until all n flags have reached it."""
        check(self.lib.vxpt_wait_next(self.handle, C.c_void_p(flags_ptr), int(n), int(stride_words), C.c_void_p(counter_ptr), int(lag),
                                      int(timeout_ms), C.c_void_p(stream)))

    def wait_all(self, flags_ptr, n, stride_words, at_least, timeout_ms=2000, stream=None):
        """Stream-ordered wait until n flag words (stride_words apart) have all reached at_least."""
        check(self.lib.vxpt_wait_all(self.handle, C.c_void_p(flags_ptr), int(n), int(stride_words), int(at_least) & 0xFFFFFFFF, int(timeout_ms),
                                     C.c_void_p(stream)))

    # ---- sync / stats -------------------------------------------------------------------------------------
    def sync(self):
        check(self.lib.vxpt_sync(self.handle))

    def stats(self):
        s = VxStats()
        check(self.lib.vxpt_get_stats(self.handle, C.byref(s)))
        return {"rays": int(s.rays), "df_fetches": int(s.df_fetches), "vox_fetches": int(s.vox_fetches), "last_ms": float(s.last_ms),
                "df_build_ms": float(s.df_build_ms), "brick_pack_ms": float(s.brick_pack_ms)}

    def reset_stats(self):
        check(self.lib.vxpt_reset_stats(self.handle))

    def launch_count(self):
        n = C.c_uint64()
        check(self.lib.vxpt_launch_count(self.handle, C.byref(n)))
        return int(n.value)

    def cuda_stream(self):
        s = C.c_void_p()
        check(self.lib.vxpt_stream(self.handle, C.byref(s)))
        return s.value

    def set_option(self, option, value):
        check(self.lib.vxpt_set_option(self.handle, int(option), int(value)))

    def reserve(self, cam, max_gi_spp=1, staging_bytes=0):
        """Size the handle's scratch (GI wavefront queue, staging arena) for frames of `cam`'s size before capturing its stream into a CUDA
        graph (vxpt_reserve): a pass that had to grow its scratch during capture fails with VXPT_E_STATE."""
        check(self.lib.vxpt_reserve(self.handle, C.byref(cam), int(max_gi_spp), int(staging_bytes)))

    def measure_l2_sector_peak(self):
        g = C.c_double()
        check(self.lib.vxpt_measure_l2_sector_peak(self.handle, C.byref(g)))
        return float(g.value)


class MultiRenderer:
    """vxpt_mg_*: one frame over N devices from one host thread (csrc/mg.cu).  Scene state is replicated through device(k) — a Renderer
    over the k-th device's handle — or the broadcast helpers below; render_frame shards the frame's rows over the devices and gathers into
    the caller's planes (host planes, or device planes every device can reach).  device_ids may repeat (several slabs on one GPU)."""

    def __init__(self, device_ids):
        self.lib = abi.load()
        ids = [int(d) for d in device_ids]
        self.handle = C.c_void_p()
        check(self.lib.vxpt_mg_create(len(ids), (C.c_int * len(ids))(*ids), C.byref(self.handle)))
        self.devices = [Renderer.borrowed(self.lib.vxpt_mg_device(self.handle, k), ids[k]) for k in range(len(ids))]

    def close(self):
        if self.handle:
            self.lib.vxpt_mg_destroy(self.handle)
            self.handle = C.c_void_p()
            self.devices = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def device(self, k):
        return self.devices[k]

    def upload_world(self, blocks):
        data = blocks.data if hasattr(blocks, "zyx") else blocks
        if isinstance(data, np.ndarray) and (data.dtype != np.uint8 or data.size != abi.WORLD_VOXELS):
            raise ValueError(f"the world is {abi.WORLD_VOXELS} uint8 block ids")
        check(self.lib.vxpt_mg_upload_world(self.handle, _ptr(data)))

    def set_block(self, x, y, z, block_id):
        check(self.lib.vxpt_mg_set_block(self.handle, int(x), int(y), int(z), int(block_id)))

    def build_distance_field(self):
        check(self.lib.vxpt_mg_build_distance_field(self.handle))

    def load_scene_tables(self, *args, **kw):
        for d in self.devices:
            d.load_scene_tables(*args, **kw)

    def set_option(self, option, value):
        check(self.lib.vxpt_mg_set_option(self.handle, int(option), int(value)))

    def slab(self, cam, k):
        rb, re = C.c_int(), C.c_int()
        check(self.lib.vxpt_mg_slab(self.handle, C.byref(cam), int(k), C.byref(rb), C.byref(re)))
        return rb.value, re.value

    def render_frame(self, cam, primary, shadow=None, diffuse=None, gbuf=None, shadow_out=None, diffuse_out=None, reflection=None,
                     reflection_out=None, g_normal=None, g_pbr=None, wait=True):
        fp = VxFrameParams()
        fp.primary = C.pointer(primary)
        if shadow is not None:
            fp.shadow = C.pointer(shadow)
        if diffuse is not None:
            fp.diffuse = C.pointer(diffuse)
        if reflection is not None:
            fp.reflection = C.pointer(reflection)
        fp.g_normal, fp.g_pbr = _ptr(g_normal), _ptr(g_pbr)
        fo = VxFrameOut()
        fo.gbuffer = self.devices[0].gbuffer_struct(gbuf or {})
        so, do, ro = shadow_out or {}, diffuse_out or {}, reflection_out or {}
        fo.shadow.shadow, fo.shadow.transversal = _ptr(so.get("shadow")), _ptr(so.get("transversal"))
        fo.diffuse.sh, fo.diffuse.cocg, fo.diffuse.luma, fo.diffuse.ao_sky = (_ptr(do.get("sh")), _ptr(do.get("cocg")), _ptr(do.get("luma")),
                                                                                _ptr(do.get("ao_sky")))
        fo.reflection.color, fo.reflection.hit_distance, fo.reflection.emissive_mask = (_ptr(ro.get("color")), _ptr(ro.get("hit_distance")),
                                                                                         _ptr(ro.get("emissive_mask")))
        fn = self.lib.vxpt_mg_render_frame if wait else self.lib.vxpt_mg_render_frame_async
        check(fn(self.handle, C.byref(cam), C.byref(fp), C.byref(fo)))
        return gbuf, shadow_out, diffuse_out, reflection_out

    def frame_wait(self):
        check(self.lib.vxpt_mg_frame_wait(self.handle))

    def sync(self):
        check(self.lib.vxpt_mg_sync(self.handle))

    def stats(self):
        st = abi.VxStats()
        check(self.lib.vxpt_mg_get_stats(self.handle, C.byref(st)))
        return {"rays": int(st.rays), "df_fetches": int(st.df_fetches), "vox_fetches": int(st.vox_fetches), "last_ms": float(st.last_ms),
                "df_build_ms": float(st.df_build_ms), "brick_pack_ms": float(st.brick_pack_ms)}

    def reset_stats(self):
        check(self.lib.vxpt_mg_reset_stats(self.handle))
