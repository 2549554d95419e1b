// api.cu — the C ABI of include/vxpt.h: handle lifetime, uploads, host<->device staging, pass dispatch.
// Every export validates its arguments, never throws, and reports failures through vxpt_last_error().
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <string>
#include <vector>

#include "vxpt_internal.h"

namespace vxpt {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    std::snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s:%d in %s", (int)e, cudaGetErrorString(e), file, line, what);
    g_last_error = buf;
    cudaGetLastError();  // clear the sticky-less error state
    return VXPT_E_CUDA;
}
static int fail(int code, const char* msg) {
    g_last_error = msg;
    return code;
}

SceneDev make_scene(const vxpt_ctx* c) {
    SceneDev S;
    const int rep = c->opt_replicas > 1 ? (int)(c->frame_counter % (uint64_t)c->opt_replicas) : 0;
    S.grid = rep ? c->rep_grid[rep] : c->d_grid;
    S.df = c->d_df;
    S.steps = rep ? c->rep_steps[rep] : c->d_steps;
    S.materials = c->d_materials;
    S.sobol = c->d_bluenoise;
    S.scramble = c->d_bluenoise ? c->d_bluenoise + 65536 : nullptr;
    S.rank = c->d_bluenoise ? c->d_bluenoise + 65536 + 131072 : nullptr;
    S.albedo_lod3 = c->d_albedo;
    S.pbr_lod2 = c->d_pbr;
    S.emissive = c->d_emissive;
    S.normal_lod3 = c->d_normal;
    S.emissive_lod2 = c->d_emissive2;
    S.sky = c->d_sky;
    S.shadow_noise = c->d_shadow_noise;
    S.n_layers = c->n_layers;
    S.n_emissive = c->n_emissive;
    S.sky_n = c->sky_n;
    S.counters = c->d_counters;
    S.alpha_mips = c->d_alpha_mips;
    S.n_alpha_layers = c->n_alpha_layers;
    S.albedo_mips = c->d_albedo_mips;
    S.normal_mips = c->d_normal_mips;
    S.pbr_mips = c->d_pbr_mips;
    S.srgb_lut = c->d_srgb_lut;
    S.n_mip_layers = c->n_mip_layers;
    S.lava_albedo = c->d_lava_albedo;
    S.lava_normal = c->d_lava_normal;
    S.lut = c->d_lut;
    return S;
}

static bool is_device_pointer(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

// true while the handle's stream records a CUDA graph: nothing may synchronise or allocate then
bool stream_capturing(const vxpt_ctx* c) {
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(c->stream, &st) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return st != cudaStreamCaptureStatusNone;
}

// Grow-only device scratch of a handle (staging arena, wavefront queues).  Growing synchronises the stream and reallocates, which is
// illegal while the stream is being captured into a CUDA graph: there the call fails with VXPT_E_STATE instead and the caller is told to
// size the scratch first (vxpt_reserve, or one eager frame of the same size).
int grow_scratch(vxpt_ctx* c, void** buf, size_t* have, size_t need, const char* what) {
    if (need <= *have) return VXPT_OK;
    if (stream_capturing(c)) {
        set_error(std::string(what) + " would have to grow while the handle's stream is being captured: call vxpt_reserve (or run one frame of this size) before capturing");
        return VXPT_E_STATE;
    }
    VX_CUDA(cudaStreamSynchronize(c->stream));
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
    if (cudaMalloc(buf, need) != cudaSuccess) {
        cudaGetLastError();
        set_error(std::string(what) + ": device allocation failed");
        return VXPT_E_NOMEM;
    }
    *have = need;
    return VXPT_OK;
}

// bump allocator over a device arena for staged (host-pointer) planes; reset at the start of every pass
struct Arena {
    vxpt_ctx* c;
    size_t used = 0;
    explicit Arena(vxpt_ctx* ctx) : c(ctx) {}
    int reserve(size_t bytes) { return grow_scratch(c, &c->d_stage, &c->stage_bytes, bytes, "the staging arena"); }
    void* take(size_t bytes) {
        void* p = (char*)c->d_stage + used;
        used += (bytes + 255) & ~(size_t)255;
        return p;
    }
};

// one image plane of a pass: user pointer (host or device) -> device pointer the kernel uses
struct Plane {
    void* user = nullptr;
    void* dev = nullptr;
    size_t elem = 0;  // bytes per pixel
    bool staged = false;
    bool internal = false;  // no caller buffer: lives in the handle's arena only (vxpt_render_frame)
};

struct PassIO {
    vxpt_ctx* c;
    const VxCamera* cam;
    std::vector<Plane*> planes;
    Arena arena;
    PassIO(vxpt_ctx* ctx, const VxCamera* cm) : c(ctx), cam(cm), arena(ctx) {}
    void add(Plane& p, const void* user, size_t elem, bool required = false) {
        p.user = const_cast<void*>(user);
        p.elem = elem;
        p.internal = !user && required;
        if (user || required) planes.push_back(&p);
    }
    int resolve() {
        size_t need = 0;
        const size_t npx = (size_t)cam->width * cam->height;
        for (Plane* p : planes) {
            p->staged = p->internal || !is_device_pointer(p->user);
            if (p->staged) need += ((npx * p->elem) + 255) & ~(size_t)255;
        }
        if (need) {
            int rc = arena.reserve(need);
            if (rc) return rc;
        }
        for (Plane* p : planes) p->dev = p->staged ? arena.take(npx * p->elem) : p->user;
        return VXPT_OK;
    }
    size_t slab_offset(const Plane& p) const { return (size_t)cam->row_begin * cam->width * p.elem; }
    size_t slab_bytes(const Plane& p) const { return (size_t)(cam->row_end - cam->row_begin) * cam->width * p.elem; }
    int upload(const Plane& p) {  // host -> device, slab rows only
        if (!p.user || !p.staged) return VXPT_OK;
        VX_CUDA(cudaMemcpyAsync((char*)p.dev + slab_offset(p), (const char*)p.user + slab_offset(p), slab_bytes(p), cudaMemcpyHostToDevice,
                                c->stream));
        return VXPT_OK;
    }
    int upload_all(const Plane& p) {  // host -> device, the whole plane: inputs of stencil passes, which read rows outside the slab
        if (!p.user || !p.staged) return VXPT_OK;
        VX_CUDA(cudaMemcpyAsync(p.dev, p.user, (size_t)cam->width * cam->height * p.elem, cudaMemcpyHostToDevice, c->stream));
        return VXPT_OK;
    }
    int download(const Plane& p, bool& any) {
        if (!p.user || !p.staged) return VXPT_OK;
        VX_CUDA(cudaMemcpyAsync((char*)p.user + slab_offset(p), (const char*)p.dev + slab_offset(p), slab_bytes(p), cudaMemcpyDeviceToHost,
                                c->stream));
        any = true;
        return VXPT_OK;
    }
    // rows [rb, re) of a staged plane, on another stream (vxpt_render_frame's copy-out pipeline)
    int download_rows(const Plane& p, int rb, int re, cudaStream_t s) {
        if (!p.user || !p.staged || re <= rb) return VXPT_OK;
        const size_t off = (size_t)rb * cam->width * p.elem, bytes = (size_t)(re - rb) * cam->width * p.elem;
        VX_CUDA(cudaMemcpyAsync((char*)p.user + off, (const char*)p.dev + off, bytes, cudaMemcpyDeviceToHost, s));
        return VXPT_OK;
    }
};

// bytes per pixel of a plane: fp32 layout / the reference's texel format (VXPT_OPT_TEXEL_FORMAT)
static size_t px_bytes(const vxpt_ctx* c, size_t f32_bytes, size_t texel_bytes) { return c->opt_texel ? texel_bytes : f32_bytes; }

static int check_camera(const VxCamera* cam) {
    if (!cam) return fail(VXPT_E_INVALID, "camera is NULL");
    if (cam->width <= 0 || cam->height <= 0 || cam->width > 16384 || cam->height > 16384) return fail(VXPT_E_INVALID, "bad frame size");
    int rows = cam->height;
    if (cam->interleave_n > 1) {
        if (cam->band_rows <= 0 || cam->interleave_rank < 0 || cam->interleave_rank >= cam->interleave_n ||
            cam->height % (cam->interleave_n * cam->band_rows) != 0)
            return fail(VXPT_E_INVALID, "bad row-band interleave (need height % (interleave_n * band_rows) == 0 and 0 <= rank < n)");
        rows = cam->height / cam->interleave_n;  // virtual rows of this handle
    } else if (cam->interleave_n < 0) {
        return fail(VXPT_E_INVALID, "interleave_n < 0");
    }
    if (cam->row_begin < 0 || cam->row_end > rows || cam->row_begin > cam->row_end)
        return fail(VXPT_E_INVALID, "row slab outside the frame");
    return VXPT_OK;
}
// a frame enqueued by vxpt_render_frame_async owns the staging arena until its copies have landed
static int finish_pending_frame(vxpt_ctx* c) {
    if (!c->frame_pending) return VXPT_OK;
    VX_CUDA(cudaSetDevice(c->device));
    VX_CUDA(cudaStreamSynchronize(c->copy_stream));
    c->frame_pending = false;
    return VXPT_OK;
}
static int check_ready(vxpt_ctx* c) {
    if (!c) return fail(VXPT_E_INVALID, "handle is NULL");
    if (int rc = finish_pending_frame(c)) return rc;
    if (!c->world_uploaded) return fail(VXPT_E_STATE, "no world uploaded (vxpt_upload_world)");
    if (!c->df_valid) return fail(VXPT_E_STATE, "distance field is stale: call vxpt_build_distance_field after editing the world");
    return VXPT_OK;
}

// u_ShouldAlphaTest: StopRay reads BlockTransparentData / BlockAlbedoData and the albedo array's alpha mip chain
static int check_alpha(const vxpt_ctx* c, float fov_degrees) {
    if (!c->have_materials || !c->d_alpha_mips) return fail(VXPT_E_STATE, "alpha_test needs vxpt_set_materials and vxpt_set_albedo_alpha_mips");
    if (!(fov_degrees > 0.0f && fov_degrees < 180.0f)) return fail(VXPT_E_INVALID, "alpha_test needs fov_degrees in (0, 180)");
    return VXPT_OK;
}
static int check_primary(const vxpt_ctx* c, const VxPrimaryParams* p) {
    if (p->alpha_test)
        if (int rc = check_alpha(c, p->fov_degrees)) return rc;
    if (p->max_iterations < 0) return fail(VXPT_E_INVALID, "max_iterations < 0");
    return VXPT_OK;
}
static int check_shadow(const vxpt_ctx* c, const VxShadowParams* p) {
    if (p->alpha_test)
        if (int rc = check_alpha(c, p->fov_degrees)) return rc;
    if (p->soft && !c->have_shadow_noise) return fail(VXPT_E_STATE, "soft shadows need vxpt_set_shadow_noise");
    return VXPT_OK;
}
static int check_diffuse(const vxpt_ctx* c, const VxDiffuseParams* p) {
    if (!p->use_blue_noise) return fail(VXPT_E_UNSUPPORTED, "u_UseBlueNoise=false (fract(sin()) hash) is not reproducible; outside the parity profile");
    if (p->direct_sampling) return fail(VXPT_E_UNSUPPORTED, "u_UseDirectSampling (WIP light-chunk sampling) is outside the v1 parity profile");
    if (!c->have_materials || !c->have_bluenoise || !c->have_textures || !c->have_sky)
        return fail(VXPT_E_STATE, "diffuse GI needs materials, blue-noise tables, material textures and a sky cubemap");
    if (p->trace_length < 0 || p->spp < 0) return fail(VXPT_E_INVALID, "negative trace_length / spp");
    // every layer the table can reach must exist in the baked arrays
    for (int b = 0; b < 128; ++b) {
        if (c->h_materials[b] >= c->n_layers) return fail(VXPT_E_INVALID, "material table references an albedo layer that was not uploaded");
        if (c->h_materials[384 + b] >= c->n_emissive) return fail(VXPT_E_INVALID, "material table references an emissive layer that was not uploaded");
    }
    return VXPT_OK;
}
// Rows of the G-buffer's distance / normal-id planes the reflection pass reads beyond its slab: it samples them at uv + clamp(u_Halton) /
// dims (ReflectionTraceFrag.glsl:754-777), the distance bilinearly (rows floor(j + hy), + 1), the normal id at the nearest texel; one more
// row each way covers the rounding of the coordinate arithmetic.
static void reflection_halo(const VxReflectionParams* p, int* below, int* above) {
    const float hy = std::min(std::max(p->halton[1], -2.0f), 2.0f);
    *below = (int)std::ceil(std::max(-hy, 0.0f)) + 1;
    *above = (int)std::ceil(std::max(hy, 0.0f)) + 1;
}

static int check_reflection(const vxpt_ctx* c, const VxReflectionParams* p) {
    if (!c->have_materials || !c->have_bluenoise || !c->have_textures || !c->have_sky || !c->have_refl_textures)
        return fail(VXPT_E_STATE, "reflections need materials, blue-noise tables, material + reflection textures and a sky cubemap");
    if (p->trace_length < 0 || p->spp < 0) return fail(VXPT_E_INVALID, "negative trace_length / spp");
    for (int b = 0; b < 128; ++b) {
        if (c->h_materials[b] >= c->n_layers || c->h_materials[256 + b] >= c->n_layers)
            return fail(VXPT_E_INVALID, "material table references an albedo / PBR layer that was not uploaded");
        if (c->h_materials[128 + b] >= c->n_normal) return fail(VXPT_E_INVALID, "material table references a normal layer that was not uploaded");
        if (c->h_materials[384 + b] >= c->n_emissive2) return fail(VXPT_E_INVALID, "material table references an emissive layer that was not uploaded");
    }
    for (int k = 1; k < 10; ++k) {
        const int lim = (k % 3 == 2) ? c->n_normal : c->n_layers;  // props: id, then (albedo, normal, pbr) x {top, side, bottom}
        if (p->grass_props[k] < 0 || p->grass_props[k] >= lim) return fail(VXPT_E_INVALID, "u_GrassBlockProps layer out of range");
    }
    return VXPT_OK;
}

// VXPT_OPT_SCENE_REPLICAS: (re)make the extra copies of grid + step field
static int refresh_replicas(vxpt_ctx* c) {
    for (int r = 1; r < c->opt_replicas; ++r) {
        if (!c->rep_grid[r]) {
            if (cudaMalloc(&c->rep_grid[r], VOXELS) != cudaSuccess || cudaMalloc(&c->rep_steps[r], STEPS_TILED_BYTES) != cudaSuccess) {
                cudaGetLastError();
                return fail(VXPT_E_NOMEM, "scene replica allocation failed");
            }
        }
        VX_CUDA(cudaMemcpyAsync(c->rep_grid[r], c->d_grid, VOXELS, cudaMemcpyDeviceToDevice, c->stream));
        VX_CUDA(cudaMemcpyAsync(c->rep_steps[r], c->d_steps, STEPS_TILED_BYTES, cudaMemcpyDeviceToDevice, c->stream));
    }
    return VXPT_OK;
}

__global__ void scatter_blocks(uint8_t* grid, const int16_t* xyz, const uint8_t* ids, int n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int x = xyz[3 * k], y = xyz[3 * k + 1], z = xyz[3 * k + 2];
    grid[(size_t)x + (size_t)WX * ((size_t)y + (size_t)WY * (size_t)z)] = ids[k];
}

}  // namespace vxpt

using namespace vxpt;

extern "C" {

const char* vxpt_last_error(void) { return g_last_error.c_str(); }
const char* vxpt_version(void) { return "vxpt 0.1.0 (sm_100a)"; }

int vxpt_create(int device_id, vxpt_handle* out) {
    if (!out) return fail(VXPT_E_INVALID, "out handle is NULL");
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(VXPT_E_CUDA, "no CUDA device available (this library has no CPU fallback)");
    }
    if (device_id < 0 || device_id >= ndev) return fail(VXPT_E_INVALID, "device_id out of range");
    VX_CUDA(cudaSetDevice(device_id));
    vxpt_ctx* c = new (std::nothrow) vxpt_ctx();
    if (!c) return fail(VXPT_E_NOMEM, "host allocation failed");
    c->device = device_id;
    if (const char* env = std::getenv("VXPT_TRAVERSAL_LAYOUT")) c->opt_layout = (env[0] == '0') ? 0 : 1;  // experiment knob
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev1);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev2);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev3);
    if (e == cudaSuccess) e = cudaEventCreate(&c->ev4);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (int k = 0; k < 8 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&c->ev_slab[k], cudaEventDisableTiming);
    if (e == cudaSuccess) {
        int least = 0, greatest = 0;
        e = cudaDeviceGetStreamPriorityRange(&least, &greatest);
        if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&c->gi_stream, cudaStreamNonBlocking, greatest);
    }
    for (int k = 0; k < 9 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&c->ev_gi[k], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_wait_err, sizeof(unsigned));
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_wait_err, 0, sizeof(unsigned), c->stream);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_grid, VOXELS);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_df, VOXELS);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_tmp, VOXELS);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_steps, STEPS_TILED_BYTES);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_counters, sizeof(DeviceCounters));
    if (e == cudaSuccess) e = cudaMemsetAsync(c->d_counters, 0, sizeof(DeviceCounters), c->stream);
    if (e == cudaSuccess) e = cudaMalloc(&c->d_lut, LUT_FLOATS * sizeof(float));
    if (e == cudaSuccess) {
        float lut[LUT_FLOATS];
        fill_trace_lut(lut);
        e = cudaMemcpy(c->d_lut, lut, sizeof lut, cudaMemcpyHostToDevice);
    }
    if (e != cudaSuccess) {
        int rc = cuda_fail(e, "vxpt_create allocations", __FILE__, __LINE__);
        vxpt_destroy(c);
        return rc == VXPT_E_CUDA && e == cudaErrorMemoryAllocation ? VXPT_E_NOMEM : rc;
    }
    if (int rc = init_df_kernels(c)) {
        vxpt_destroy(c);
        return rc;
    }
    *out = c;
    return VXPT_OK;
}

int vxpt_destroy(vxpt_handle c) {
    if (!c) return VXPT_OK;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    c->frame_pending = false;
    if (c->copy_stream) {
        cudaStreamSynchronize(c->copy_stream);
        cudaStreamDestroy(c->copy_stream);
    }
    for (int k = 0; k < 8; ++k)
        if (c->ev_slab[k]) cudaEventDestroy(c->ev_slab[k]);
    if (c->gi_stream) {
        cudaStreamSynchronize(c->gi_stream);
        cudaStreamDestroy(c->gi_stream);
    }
    for (int k = 0; k < 9; ++k)
        if (c->ev_gi[k]) cudaEventDestroy(c->ev_gi[k]);
    for (const vxpt_ctx::SharedBuf& b : c->shared) {
        if (b.owned) cudaFree(b.ptr);
        else cudaIpcCloseMemHandle(b.ptr);
    }
    if (c->d_wait_err) cudaFree(c->d_wait_err);
    for (int r = 1; r < 8; ++r) {
        if (c->rep_grid[r]) cudaFree(c->rep_grid[r]);
        if (c->rep_steps[r]) cudaFree(c->rep_steps[r]);
    }
    void* bufs[] = {c->d_grid, c->d_df, c->d_tmp, c->d_steps, c->d_materials, c->d_bluenoise, c->d_albedo, c->d_pbr,
                    c->d_emissive, c->d_normal, c->d_emissive2, c->d_sky, c->d_shadow_noise, c->d_alpha_mips, c->d_counters, c->d_stage, c->d_queue,
                    c->d_albedo_mips, c->d_normal_mips, c->d_pbr_mips, c->d_srgb_lut, c->svgf.buf, c->shadow_hist.buf, c->d_lava_albedo, c->d_lava_normal, c->d_lut};
    for (void* b : bufs)
        if (b) cudaFree(b);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->ev2) cudaEventDestroy(c->ev2);
    if (c->ev3) cudaEventDestroy(c->ev3);
    if (c->ev4) cudaEventDestroy(c->ev4);
    if (c->stream) cudaStreamDestroy(c->stream);
    cudaGetLastError();
    delete c;
    return VXPT_OK;
}

// ---------------------------------------------------------------------------------------------------- world
int vxpt_upload_world(vxpt_handle c, const uint8_t* blocks) {
    if (!c || !blocks) return fail(VXPT_E_INVALID, "NULL argument");
    VX_CUDA(cudaSetDevice(c->device));
    VX_CUDA(cudaMemcpyAsync(c->d_grid, blocks, VOXELS, cudaMemcpyDefault, c->stream));
    // the copy must have consumed the caller's buffer before we return (World::Buffer semantics, World.h:167-171)
    if (!is_device_pointer(blocks)) VX_CUDA(cudaStreamSynchronize(c->stream));
    c->world_uploaded = true;
    c->df_valid = false;
    return VXPT_OK;
}

int vxpt_download_world(vxpt_handle c, uint8_t* blocks) {
    if (!c || !blocks) return fail(VXPT_E_INVALID, "NULL argument");
    if (!c->world_uploaded) return fail(VXPT_E_STATE, "no world uploaded");
    VX_CUDA(cudaSetDevice(c->device));
    VX_CUDA(cudaMemcpyAsync(blocks, c->d_grid, VOXELS, cudaMemcpyDefault, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXPT_OK;
}

int vxpt_set_block(vxpt_handle c, int x, int y, int z, uint8_t id) {
    if (!c) return fail(VXPT_E_INVALID, "handle is NULL");
    if (!c->world_uploaded) return fail(VXPT_E_STATE, "no world uploaded");
    if (x < 0 || y < 0 || z < 0 || x >= WX || y >= WY || z >= WZ) return fail(VXPT_E_INVALID, "voxel outside the world");
    VX_CUDA(cudaSetDevice(c->device));
    VX_CUDA(cudaMemsetAsync(c->d_grid + ((size_t)x + (size_t)WX * ((size_t)y + (size_t)WY * (size_t)z)), id, 1, c->stream));
    c->df_valid = false;
    return VXPT_OK;
}

int vxpt_set_blocks(vxpt_handle c, const int16_t* xyz, const uint8_t* ids, int n) {
    if (!c || (n > 0 && (!xyz || !ids)) || n < 0) return fail(VXPT_E_INVALID, "bad argument");
    if (!c->world_uploaded) return fail(VXPT_E_STATE, "no world uploaded");
    if (n == 0) return VXPT_OK;
    if (int rc = finish_pending_frame(c)) return rc;
    for (int k = 0; k < n; ++k)
        if (xyz[3 * k] < 0 || xyz[3 * k + 1] < 0 || xyz[3 * k + 2] < 0 || xyz[3 * k] >= WX || xyz[3 * k + 1] >= WY || xyz[3 * k + 2] >= WZ)
            return fail(VXPT_E_INVALID, "voxel outside the world");
    VX_CUDA(cudaSetDevice(c->device));
    Arena arena(c);
    const size_t b_xyz = (size_t)n * 3 * sizeof(int16_t), b_ids = (size_t)n;
    int rc = arena.reserve(((b_xyz + 255) & ~(size_t)255) + ((b_ids + 255) & ~(size_t)255));
    if (rc) return rc;
    int16_t* d_xyz = (int16_t*)arena.take(b_xyz);
    uint8_t* d_ids = (uint8_t*)arena.take(b_ids);
    VX_CUDA(cudaMemcpyAsync(d_xyz, xyz, b_xyz, cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaMemcpyAsync(d_ids, ids, b_ids, cudaMemcpyHostToDevice, c->stream));
    VX_LAUNCH(scatter_blocks, dim3((n + 127) / 128), 128, c->stream, c->d_grid, d_xyz, d_ids, n);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    VX_CUDA(cudaStreamSynchronize(c->stream));
    c->df_valid = false;
    return VXPT_OK;
}

// --------------------------------------------------------------------------------------------- distance field
int vxpt_build_distance_field(vxpt_handle c) {
    if (!c) return fail(VXPT_E_INVALID, "handle is NULL");
    if (!c->world_uploaded) return fail(VXPT_E_STATE, "no world uploaded");
    VX_CUDA(cudaSetDevice(c->device));
    VX_CUDA(cudaEventRecord(c->ev2, c->stream));
    int rc = launch_df_build(c);
    if (rc) return rc;
    VX_CUDA(cudaEventRecord(c->ev3, c->stream));
    c->pack_timed = c->steps_layout != c->opt_layout;  // the DPX build writes the step field itself: nothing follows, nothing to time
    if (c->pack_timed) {
        if ((rc = launch_pack_bricks(c))) return rc;
        VX_CUDA(cudaEventRecord(c->ev4, c->stream));
    }
    if ((rc = refresh_replicas(c))) return rc;
    c->df_timed = true;
    c->df_valid = true;
    return VXPT_OK;
}

int vxpt_download_distance_field(vxpt_handle c, uint8_t* out) {
    if (!c || !out) return fail(VXPT_E_INVALID, "NULL argument");
    if (!c->df_valid) return fail(VXPT_E_STATE, "distance field not built");
    VX_CUDA(cudaSetDevice(c->device));
    VX_CUDA(cudaMemcpyAsync(out, c->d_df, VOXELS, cudaMemcpyDefault, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXPT_OK;
}

int vxpt_device_pointers(vxpt_handle c, const uint8_t** grid, const uint8_t** df) {
    if (!c) return fail(VXPT_E_INVALID, "handle is NULL");
    if (grid) *grid = c->d_grid;
    if (df) *df = c->d_df;
    return VXPT_OK;
}

// ----------------------------------------------------------------------------------------------------- tables
}  // extern "C"
template <class T>
static int replace_buffer(vxpt_ctx* c, T** slot, const void* src, size_t bytes) {
    VX_CUDA(cudaStreamSynchronize(c->stream));
    if (*slot) cudaFree(*slot);
    *slot = nullptr;
    if (cudaMalloc((void**)slot, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(VXPT_E_NOMEM, "device allocation failed");
    }
    VX_CUDA(cudaMemcpyAsync(*slot, src, bytes, cudaMemcpyDefault, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXPT_OK;
}
extern "C" {

int vxpt_set_materials(vxpt_handle c, const int32_t table[768]) {
    if (!c || !table) return fail(VXPT_E_INVALID, "NULL argument");
    VX_CUDA(cudaSetDevice(c->device));
    int rc = replace_buffer(c, &c->d_materials, table, 768 * sizeof(int32_t));
    if (rc) return rc;
    std::memcpy(c->h_materials, table, sizeof c->h_materials);
    c->have_materials = true;
    return VXPT_OK;
}

int vxpt_set_blue_noise(vxpt_handle c, const int32_t* sobol, const int32_t* scramble, const int32_t* rank) {
    if (!c || !sobol || !scramble || !rank) return fail(VXPT_E_INVALID, "NULL argument");
    VX_CUDA(cudaSetDevice(c->device));
    // The Heitz et al. tables hold 8-bit values (Core/BlueNoiseDataSSBO.cpp:4,9,14); keep them as bytes on the
    // device (320 KB instead of 1.25 MB of int32, so they stay L1/L2 friendly).
    std::vector<uint8_t> packed(65536 + 131072 + 131072);
    const int32_t* src[3] = {sobol, scramble, rank};
    const size_t len[3] = {65536, 131072, 131072};
    size_t o = 0;
    for (int t = 0; t < 3; ++t)
        for (size_t k = 0; k < len[t]; ++k) {
            int32_t v = src[t][k];
            if (v < 0 || v > 255) return fail(VXPT_E_UNSUPPORTED, "blue-noise table value outside 0..255");
            packed[o++] = (uint8_t)v;
        }
    int rc = replace_buffer(c, &c->d_bluenoise, packed.data(), packed.size());
    if (rc) return rc;
    c->have_bluenoise = true;
    return VXPT_OK;
}

int vxpt_set_material_textures(vxpt_handle c, const float* albedo_lod3, const float* pbr_lod2, int n_layers, const float* emissive_lod0,
                               int n_emissive_layers) {
    if (!c || !albedo_lod3 || !pbr_lod2 || n_layers <= 0 || n_emissive_layers < 0 || (n_emissive_layers > 0 && !emissive_lod0))
        return fail(VXPT_E_INVALID, "bad argument");
    VX_CUDA(cudaSetDevice(c->device));
    int rc = replace_buffer(c, &c->d_albedo, albedo_lod3, (size_t)n_layers * 64 * 64 * 4 * sizeof(float));
    if (rc) return rc;
    rc = replace_buffer(c, &c->d_pbr, pbr_lod2, (size_t)n_layers * 128 * 128 * 4 * sizeof(float));
    if (rc) return rc;
    if (n_emissive_layers > 0) {
        rc = replace_buffer(c, &c->d_emissive, emissive_lod0, (size_t)n_emissive_layers * 512 * 512 * sizeof(float));
        if (rc) return rc;
    }
    c->n_layers = n_layers;
    c->n_emissive = n_emissive_layers;
    c->have_textures = true;
    return VXPT_OK;
}

int vxpt_set_reflection_textures(vxpt_handle c, const float* normal_lod3, int n_normal_layers, const float* emissive_lod2, int n_emissive_layers) {
    if (!c || !normal_lod3 || n_normal_layers <= 0 || n_emissive_layers < 0 || (n_emissive_layers > 0 && !emissive_lod2))
        return fail(VXPT_E_INVALID, "bad argument");
    VX_CUDA(cudaSetDevice(c->device));
    int rc = replace_buffer(c, &c->d_normal, normal_lod3, (size_t)n_normal_layers * 64 * 64 * 4 * sizeof(float));
    if (rc) return rc;
    if (n_emissive_layers > 0) {
        rc = replace_buffer(c, &c->d_emissive2, emissive_lod2, (size_t)n_emissive_layers * 128 * 128 * sizeof(float));
        if (rc) return rc;
    }
    c->n_normal = n_normal_layers;
    c->n_emissive2 = n_emissive_layers;
    c->have_refl_textures = true;
    return VXPT_OK;
}

int vxpt_set_sky_cubemap(vxpt_handle c, const float* rgb, int n) {
    if (!c || !rgb || n <= 0 || n > 4096) return fail(VXPT_E_INVALID, "bad argument");
    VX_CUDA(cudaSetDevice(c->device));
    int rc = replace_buffer(c, &c->d_sky, rgb, (size_t)6 * n * n * 3 * sizeof(float));
    if (rc) return rc;
    c->sky_n = n;
    c->h_sky.resize((size_t)6 * n * n * 3);
    VX_CUDA(cudaMemcpy(c->h_sky.data(), c->d_sky, c->h_sky.size() * sizeof(float), cudaMemcpyDeviceToHost));
    c->have_sky = true;
    return VXPT_OK;
}

int vxpt_set_shadow_noise(vxpt_handle c, const uint8_t* rgba8) {
    if (!c || !rgba8) return fail(VXPT_E_INVALID, "NULL argument");
    VX_CUDA(cudaSetDevice(c->device));
    int rc = replace_buffer(c, &c->d_shadow_noise, rgba8, (size_t)256 * 256 * 4);
    if (rc) return rc;
    c->have_shadow_noise = true;
    return VXPT_OK;
}

int vxpt_set_albedo_alpha_mips(vxpt_handle c, const uint8_t* alpha_mips, int n_layers) {
    if (!c || !alpha_mips) return fail(VXPT_E_INVALID, "NULL argument");
    if (n_layers <= 0 || n_layers > 4096) return fail(VXPT_E_INVALID, "bad layer count");
    VX_CUDA(cudaSetDevice(c->device));
    if (int rc = finish_pending_frame(c)) return rc;
    int rc = replace_buffer(c, &c->d_alpha_mips, alpha_mips, (size_t)n_layers * VXPT_ALPHA_MIP_TEXELS);
    if (rc) return rc;
    c->n_alpha_layers = n_layers;
    return VXPT_OK;
}

// ----------------------------------------------------------------------------------------------------- G-buffer material pass
int vxpt_set_gbuffer_textures(vxpt_handle c, const uint8_t* albedo_mips, const uint8_t* normal_mips, const uint8_t* pbr_mips, int n_layers) {
    if (!c || !albedo_mips || !normal_mips || !pbr_mips) return fail(VXPT_E_INVALID, "NULL argument");
    if (n_layers <= 0 || n_layers > 1024) return fail(VXPT_E_INVALID, "bad layer count");
    VX_CUDA(cudaSetDevice(c->device));
    if (int rc = finish_pending_frame(c)) return rc;
    const size_t bytes = (size_t)n_layers * VXPT_MIP_CHAIN_TEXELS * 4;
    int rc;
    if ((rc = replace_buffer(c, &c->d_albedo_mips, albedo_mips, bytes)) || (rc = replace_buffer(c, &c->d_normal_mips, normal_mips, bytes)) ||
        (rc = replace_buffer(c, &c->d_pbr_mips, pbr_mips, bytes)))
        return rc;
    // GL_SRGB_ALPHA decode (OpenGL 4.3 section 8.23): evaluated in double, rounded once to fp32
    // followed by the plain unorm8 decode c / 255 (fp32 division), so that a texel fetch is table reads only
    float lut[512];
    for (int k = 0; k < 256; ++k) {
        const double cs = (double)k / 255.0;
        lut[k] = (float)(cs <= 0.04045 ? cs / 12.92 : std::pow((cs + 0.055) / 1.055, 2.4));
        lut[256 + k] = (float)k / 255.0f;
    }
    if ((rc = replace_buffer(c, &c->d_srgb_lut, lut, sizeof lut))) return rc;
    c->n_mip_layers = n_layers;
    return VXPT_OK;
}

int vxpt_set_lava_textures(vxpt_handle c, const uint8_t* albedo_rgba8, const uint8_t* normal_rgba8) {
    if (!c || !albedo_rgba8 || !normal_rgba8) return fail(VXPT_E_INVALID, "NULL argument");
    VX_CUDA(cudaSetDevice(c->device));
    if (int rc = finish_pending_frame(c)) return rc;
    const size_t bytes = (size_t)VXPT_LAVA_FRAMES * VXPT_LAVA_SIZE * VXPT_LAVA_SIZE * 4;
    int rc;
    if ((rc = replace_buffer(c, &c->d_lava_albedo, albedo_rgba8, bytes)) || (rc = replace_buffer(c, &c->d_lava_normal, normal_rgba8, bytes))) return rc;
    return VXPT_OK;
}

}  // extern "C"
static int check_material(const vxpt_ctx* c, const VxCamera* cam, const VxMaterialParams* p);
extern "C" {

int vxpt_generate_gbuffer(vxpt_handle c, const VxCamera* cam, const VxGBuffer* g, const VxMaterialParams* p, const VxMaterialOut* out) {
    int rc = check_ready(c);
    if (rc) return rc;
    if ((rc = check_camera(cam))) return rc;
    if (!g || !p || !out || !g->inv_t || !g->normal_id || !g->block_id)
        return fail(VXPT_E_INVALID, "NULL argument (G-buffer inv_t, normal_id and block_id are required)");
    if ((rc = check_material(c, cam, p))) return rc;
    VX_CUDA(cudaSetDevice(c->device));
    PassIO io(c, cam);
    Plane it, nid, bid, al, nm, pb, ao;
    io.add(it, g->inv_t, 4); io.add(nid, g->normal_id, 1); io.add(bid, g->block_id, 1);
    io.add(al, out->albedo, 12); io.add(nm, out->normal, 12); io.add(pb, out->pbr, 16); io.add(ao, out->texture_ao, 4);
    if ((rc = io.resolve())) return rc;
    if (cam->row_end == cam->row_begin) return VXPT_OK;
    if ((rc = io.upload(it)) || (rc = io.upload(nid)) || (rc = io.upload(bid))) return rc;
    if (!p->update_this_frame && !p->pom) {  // invocations discard (all but lava pixels): staged output planes must keep what the caller's planes hold
        if ((rc = io.upload(al)) || (rc = io.upload(nm)) || (rc = io.upload(pb)) || (rc = io.upload(ao))) return rc;
    }
    VxGBuffer gd{nullptr, (uint8_t*)nid.dev, (uint8_t*)bid.dev, (float*)it.dev, nullptr};
    VxMaterialOut od{(float*)al.dev, (float*)nm.dev, (float*)pb.dev, (float*)ao.dev};
    if (c->opt_timing) VX_CUDA(cudaEventRecord(c->ev0, c->stream));
    if ((rc = launch_gbuffer(c, *cam, gd, *p, od))) return rc;
    if (c->opt_timing) {
        VX_CUDA(cudaEventRecord(c->ev1, c->stream));
        c->pass_timed = true;
    }
    bool any = false;
    if ((rc = io.download(al, any)) || (rc = io.download(nm, any)) || (rc = io.download(pb, any)) || (rc = io.download(ao, any))) return rc;
    if (any) VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXPT_OK;
}

}  // extern "C"
// parameter and state checks of the G-buffer material pass (vxpt_generate_gbuffer, vxpt_render_frame)
static int check_material(const vxpt_ctx* c, const VxCamera* cam, const VxMaterialParams* p) {
    if (p->pom && !(p->pom_height >= 0.0f && p->pom_height <= 16.0f && p->pom_exp >= 0.0f && p->pom_exp <= 16.0f))
        return fail(VXPT_E_INVALID, "u_POMHeight / u_POMExp outside 0..16");
    if (p->lava_block_id >= 0 && !c->d_lava_albedo) return fail(VXPT_E_STATE, "a lava block id needs vxpt_set_lava_textures");
    if (p->lava_block_id > 127) return fail(VXPT_E_INVALID, "lava_block_id outside the material table");
    if (!c->have_materials || !c->d_albedo_mips) return fail(VXPT_E_STATE, "the G-buffer pass needs vxpt_set_materials and vxpt_set_gbuffer_textures");
    const int rows = cam->interleave_n > 1 ? cam->height / cam->interleave_n : cam->height;
    if ((cam->row_begin & 1) || ((cam->row_end & 1) && cam->row_end != rows) || (cam->interleave_n > 1 && (cam->band_rows & 1)))
        return fail(VXPT_E_INVALID, "the G-buffer pass shades 2x2 quads: row_begin / row_end (and band_rows) must be even");
    for (int b = 0; b < 128; ++b) {
        for (int k = 0; k < 3; ++k)
            if (c->h_materials[128 * k + b] >= c->n_mip_layers)
                return fail(VXPT_E_INVALID, "material table references a texture layer that vxpt_set_gbuffer_textures did not receive");
        if (c->h_materials[384 + b] >= c->n_emissive) return fail(VXPT_E_INVALID, "material table references an emissive layer that was not uploaded");
    }
    for (int k = 1; k < 10; ++k)
        if (p->grass_props[k] < 0 || p->grass_props[k] >= c->n_mip_layers) return fail(VXPT_E_INVALID, "grass_props references a missing texture layer");
    return VXPT_OK;
}

// ----------------------------------------------------------------------------------------------------- SVGF denoiser
// shared by the three passes: a handle that is ready (no pending frame), fp32 planes, whole-frame inputs, plain row slabs
static int check_svgf(vxpt_ctx* c, const VxCamera* cam) {
    if (!c) return fail(VXPT_E_INVALID, "handle is NULL");
    if (int rc = finish_pending_frame(c)) return rc;
    if (int rc = check_camera(cam)) return rc;
    if (cam->interleave_n > 1) return fail(VXPT_E_UNSUPPORTED, "the denoiser passes read neighbouring rows: interleaved row bands are not supported");
    if (c->opt_texel) return fail(VXPT_E_UNSUPPORTED, "the denoiser passes take fp32 planes (VXPT_OPT_TEXEL_FORMAT must be 0)");
    return VXPT_OK;
}
namespace {
struct SvgfIO {  // input planes are staged whole, output planes by slab
    PassIO io;
    std::vector<Plane*> inputs, outputs;
    std::vector<Plane> store;
    SvgfIO(vxpt_ctx* c, const VxCamera* cam) : io(c, cam) { store.reserve(32); }
    Plane* in(const void* user, size_t elem) { store.emplace_back(); io.add(store.back(), user, elem); inputs.push_back(&store.back()); return &store.back(); }
    Plane* out(const void* user, size_t elem) { store.emplace_back(); io.add(store.back(), user, elem); outputs.push_back(&store.back()); return &store.back(); }
    int begin() {
        if (int rc = io.resolve()) return rc;
        for (Plane* p : inputs)
            if (int rc = io.upload_all(*p)) return rc;
        return VXPT_OK;
    }
    int end(vxpt_ctx* c) {
        bool any = false;
        for (Plane* p : outputs)
            if (int rc = io.download(*p, any)) return rc;
        if (any) VX_CUDA(cudaStreamSynchronize(c->stream));
        return VXPT_OK;
    }
};
template <class F> int timed_launch(vxpt_ctx* c, F&& launch) {
    if (c->opt_timing) VX_CUDA(cudaEventRecord(c->ev0, c->stream));
    if (int rc = launch()) return rc;
    if (c->opt_timing) {
        VX_CUDA(cudaEventRecord(c->ev1, c->stream));
        c->pass_timed = true;
    }
    return VXPT_OK;
}
}  // namespace

extern "C" {

int vxpt_svgf_initial(vxpt_handle c, const VxCamera* cam, const VxSvgfInitialIn* in, const VxSvgfInitialOut* out) {
    int rc = check_svgf(c, cam);
    if (rc) return rc;
    if (!in || !out || !in->current.t || !in->current.normal_id || !in->sh || !in->cocg || !in->luma || !in->ao_sky)
        return fail(VXPT_E_INVALID, "NULL argument (t, normal_id and the GI pass's sh / cocg / luma / ao_sky planes are required)");
    VX_CUDA(cudaSetDevice(c->device));
    SvgfIO s(c, cam);
    Plane *t = s.in(in->current.t, 4), *n = s.in(in->current.normal_id, 1);
    Plane *sh = s.in(in->sh, 16), *cc = s.in(in->cocg, 8), *lu = s.in(in->luma, 4), *ao = s.in(in->ao_sky, 8);
    Plane *osh = s.out(out->sh, 16), *occ = s.out(out->cocg, 8), *olu = s.out(out->luma, 4), *oao = s.out(out->ao_sky, 8);
    if ((rc = s.begin())) return rc;
    if (cam->row_end == cam->row_begin) return VXPT_OK;
    VxSvgfInitialIn id{};
    id.current = VxGBuffer{(float*)t->dev, (uint8_t*)n->dev, nullptr, nullptr, nullptr};
    id.sh = (const float*)sh->dev; id.cocg = (const float*)cc->dev; id.luma = (const float*)lu->dev; id.ao_sky = (const float*)ao->dev;
    const VxSvgfInitialOut od{(float*)osh->dev, (float*)occ->dev, (float*)olu->dev, (float*)oao->dev};
    if ((rc = timed_launch(c, [&] { return launch_svgf_initial(c, *cam, id, od); }))) return rc;
    return s.end(c);
}

int vxpt_svgf_temporal(vxpt_handle c, const VxCamera* cam, const VxSvgfTemporalIn* in, const VxSvgfTemporalParams* p, const VxSvgfTemporalOut* out) {
    int rc = check_svgf(c, cam);
    if (rc) return rc;
    if (!in || !p || !out) return fail(VXPT_E_INVALID, "NULL argument");
    const void* req[] = {in->current.t, in->current.normal_id, in->current.block_id, in->previous.t, in->previous.normal_id, in->previous.block_id,
                         in->sh, in->cocg, in->luma, in->ao_sky, in->prev_sh, in->prev_cocg, in->prev_utility, in->prev_ao_sky};
    for (const void* q : req)
        if (!q) return fail(VXPT_E_INVALID, "NULL input plane (current / previous t, normal_id, block_id, the GI planes and the previous temporal planes are required)");
    VX_CUDA(cudaSetDevice(c->device));
    SvgfIO s(c, cam);
    Plane *t = s.in(in->current.t, 4), *n = s.in(in->current.normal_id, 1), *b = s.in(in->current.block_id, 1);
    Plane *pt = s.in(in->previous.t, 4), *pn = s.in(in->previous.normal_id, 1), *pb = s.in(in->previous.block_id, 1);
    Plane *sh = s.in(in->sh, 16), *cc = s.in(in->cocg, 8), *lu = s.in(in->luma, 4), *ao = s.in(in->ao_sky, 8);
    Plane *psh = s.in(in->prev_sh, 16), *pcc = s.in(in->prev_cocg, 8), *put = s.in(in->prev_utility, 12), *pao = s.in(in->prev_ao_sky, 8);
    Plane *osh = s.out(out->sh, 16), *occ = s.out(out->cocg, 8), *out_ut = s.out(out->utility, 12), *oao = s.out(out->ao_sky, 8);
    if ((rc = s.begin())) return rc;
    if (cam->row_end == cam->row_begin) return VXPT_OK;
    VxSvgfTemporalIn id{};
    id.current = VxGBuffer{(float*)t->dev, (uint8_t*)n->dev, (uint8_t*)b->dev, nullptr, nullptr};
    id.previous = VxGBuffer{(float*)pt->dev, (uint8_t*)pn->dev, (uint8_t*)pb->dev, nullptr, nullptr};
    id.sh = (const float*)sh->dev; id.cocg = (const float*)cc->dev; id.luma = (const float*)lu->dev; id.ao_sky = (const float*)ao->dev;
    id.prev_sh = (const float*)psh->dev; id.prev_cocg = (const float*)pcc->dev; id.prev_utility = (const float*)put->dev; id.prev_ao_sky = (const float*)pao->dev;
    const VxSvgfTemporalOut od{(float*)osh->dev, (float*)occ->dev, (float*)out_ut->dev, (float*)oao->dev};
    if ((rc = timed_launch(c, [&] { return launch_svgf_temporal(c, *cam, id, *p, od); }))) return rc;
    return s.end(c);
}

int vxpt_svgf_variance(vxpt_handle c, const VxCamera* cam, const VxSvgfVarianceIn* in, const VxSvgfVarianceParams* p, const VxSvgfVarianceOut* out) {
    int rc = check_svgf(c, cam);
    if (rc) return rc;
    if (!in || !p || !out || !in->current.t || !in->current.normal_id || !in->sh || !in->cocg || !in->utility)
        return fail(VXPT_E_INVALID, "NULL argument (t, normal_id and the temporal pass's sh / cocg / utility planes are required)");
    VX_CUDA(cudaSetDevice(c->device));
    SvgfIO s(c, cam);
    Plane *t = s.in(in->current.t, 4), *n = s.in(in->current.normal_id, 1);
    Plane *sh = s.in(in->sh, 16), *cc = s.in(in->cocg, 8), *ut = s.in(in->utility, 12);
    Plane *osh = s.out(out->sh, 16), *occ = s.out(out->cocg, 8), *ov = s.out(out->variance, 4);
    if ((rc = s.begin())) return rc;
    if (cam->row_end == cam->row_begin) return VXPT_OK;
    VxSvgfVarianceIn id{};
    id.current = VxGBuffer{(float*)t->dev, (uint8_t*)n->dev, nullptr, nullptr, nullptr};
    id.sh = (const float*)sh->dev; id.cocg = (const float*)cc->dev; id.utility = (const float*)ut->dev;
    const VxSvgfVarianceOut od{(float*)osh->dev, (float*)occ->dev, (float*)ov->dev};
    if ((rc = timed_launch(c, [&] { return launch_svgf_variance(c, *cam, id, *p, od); }))) return rc;
    return s.end(c);
}

int vxpt_svgf_spatial(vxpt_handle c, const VxCamera* cam, const VxSvgfSpatialIn* in, const VxSvgfSpatialParams* p, const VxSvgfSpatialOut* out) {
    int rc = check_svgf(c, cam);
    if (rc) return rc;
    if (!in || !p || !out || !in->current.t || !in->current.normal_id || !in->sh || !in->cocg || !in->variance || !in->ao_sky || !in->temporal_utility)
        return fail(VXPT_E_INVALID, "NULL argument (t, normal_id, sh, cocg, variance, ao_sky and the temporal utility plane are required)");
    if (p->step < 1 || p->step > 1024) return fail(VXPT_E_INVALID, "bad a-trous step");
    VX_CUDA(cudaSetDevice(c->device));
    SvgfIO s(c, cam);
    Plane *t = s.in(in->current.t, 4), *n = s.in(in->current.normal_id, 1);
    Plane *sh = s.in(in->sh, 16), *cc = s.in(in->cocg, 8), *va = s.in(in->variance, 4), *ao = s.in(in->ao_sky, 8), *ut = s.in(in->temporal_utility, 12);
    Plane *osh = s.out(out->sh, 16), *occ = s.out(out->cocg, 8), *ov = s.out(out->variance, 4), *oao = s.out(out->ao_sky, 8);
    if ((rc = s.begin())) return rc;
    if (cam->row_end == cam->row_begin) return VXPT_OK;
    VxSvgfSpatialIn id{};
    id.current = VxGBuffer{(float*)t->dev, (uint8_t*)n->dev, nullptr, nullptr, nullptr};
    id.sh = (const float*)sh->dev; id.cocg = (const float*)cc->dev; id.variance = (const float*)va->dev; id.ao_sky = (const float*)ao->dev;
    id.temporal_utility = (const float*)ut->dev;
    const VxSvgfSpatialOut od{(float*)osh->dev, (float*)occ->dev, (float*)ov->dev, (float*)oao->dev};
    if ((rc = timed_launch(c, [&] { return launch_svgf_spatial(c, *cam, id, *p, od); }))) return rc;
    return s.end(c);
}

}  // extern "C"
// carve the history allocation into planes (256-byte aligned), (re)allocating when the frame size changes
static int svgf_history_for(vxpt_ctx* c, int W, int H) {
    vxpt_ctx::SvgfHistory& h = c->svgf;
    if (h.buf && h.width == W && h.height == H) return VXPT_OK;
    VX_CUDA(cudaStreamSynchronize(c->stream));
    if (h.buf) cudaFree(h.buf);
    h = vxpt_ctx::SvgfHistory();
    const size_t npx = (size_t)W * H;
    auto plane = [npx](size_t elem) { return (npx * elem + 255) & ~(size_t)255; };
    const size_t set4 = plane(16) + plane(8) + plane(12) + plane(8);  // sh, cocg, utility, ao (pre: luma and pong: variance fit in the utility slot)
    const size_t total = plane(4) + 2 * plane(1) + 2 * set4 + set4 + (plane(16) + plane(8) + plane(4)) + 2 * set4;
    if (cudaMalloc(&h.buf, total) != cudaSuccess) {
        cudaGetLastError();
        h.buf = nullptr;
        return fail(VXPT_E_NOMEM, "device allocation of the denoiser history failed");
    }
    char* p = (char*)h.buf;
    auto take = [&p, &plane](size_t elem) { char* q = p; p += plane(elem); return q; };
    h.prev_t = (float*)take(4); h.prev_nid = (uint8_t*)take(1); h.prev_bid = (uint8_t*)take(1);
    for (int k = 0; k < 2; ++k) { h.temporal[k][0] = (float*)take(16); h.temporal[k][1] = (float*)take(8); h.temporal[k][2] = (float*)take(12); h.temporal[k][3] = (float*)take(8); }
    h.pre[0] = (float*)take(16); h.pre[1] = (float*)take(8); h.pre[2] = (float*)take(12); h.pre[3] = (float*)take(8);
    h.var[0] = (float*)take(16); h.var[1] = (float*)take(8); h.var[2] = (float*)take(4);
    for (int k = 0; k < 2; ++k) { h.pong[k][0] = (float*)take(16); h.pong[k][1] = (float*)take(8); h.pong[k][2] = (float*)take(12); h.pong[k][3] = (float*)take(8); }
    h.width = W; h.height = H;
    return VXPT_OK;
}
extern "C" {

int vxpt_svgf_frame(vxpt_handle c, const VxCamera* cam, const VxGBuffer* g, const VxDiffuseOut* diffuse, const VxSvgfFrameParams* p, const VxSvgfSpatialOut* out) {
    int rc = check_svgf(c, cam);
    if (rc) return rc;
    if (!g || !diffuse || !p || !out || !g->t || !g->normal_id || !g->block_id || !diffuse->sh || !diffuse->cocg || !diffuse->luma || !diffuse->ao_sky)
        return fail(VXPT_E_INVALID, "NULL argument (G-buffer t / normal_id / block_id and the GI pass's sh / cocg / luma / ao_sky planes are required)");
    if (cam->row_begin != 0 || cam->row_end != cam->height) return fail(VXPT_E_INVALID, "vxpt_svgf_frame denoises whole frames (row_begin = 0, row_end = height)");
    VX_CUDA(cudaSetDevice(c->device));
    const int W = cam->width, H = cam->height;
    const size_t npx = (size_t)W * H;
    const bool had = c->svgf.buf && c->svgf.width == W && c->svgf.height == H && c->svgf.valid && !p->reset_history;
    if ((rc = svgf_history_for(c, W, H))) return rc;
    vxpt_ctx::SvgfHistory& h = c->svgf;
    SvgfIO s(c, cam);
    Plane *t = s.in(g->t, 4), *n = s.in(g->normal_id, 1), *b = s.in(g->block_id, 1);
    Plane *sh = s.in(diffuse->sh, 16), *cc = s.in(diffuse->cocg, 8), *lu = s.in(diffuse->luma, 4), *ao = s.in(diffuse->ao_sky, 8);
    Plane *osh = s.out(out->sh, 16), *occ = s.out(out->cocg, 8), *ov = s.out(out->variance, 4), *oao = s.out(out->ao_sky, 8);
    if ((rc = s.begin())) return rc;
    const VxGBuffer gd{(float*)t->dev, (uint8_t*)n->dev, (uint8_t*)b->dev, nullptr, nullptr};
    const int prev = h.cur, cur = h.cur ^ 1;
    VxSvgfTemporalParams tp{};
    if (!had) {  // a new history: previous planes zero, previous G-buffer and camera = this frame's
        for (int k = 0; k < 4; ++k) VX_CUDA(cudaMemsetAsync(h.temporal[prev][k], 0, npx * (k == 0 ? 16 : (k == 2 ? 12 : 8)), c->stream));
        std::memcpy(tp.prev_view, p->view, sizeof tp.prev_view);
        std::memcpy(tp.prev_projection, p->projection, sizeof tp.prev_projection);
    } else {
        std::memcpy(tp.prev_view, h.prev_view, sizeof tp.prev_view);
        std::memcpy(tp.prev_projection, h.prev_projection, sizeof tp.prev_projection);
    }
    tp.be_useful = 1;
    if (c->opt_timing) VX_CUDA(cudaEventRecord(c->ev0, c->stream));
    VxSvgfTemporalIn ti{};
    ti.current = gd;
    ti.previous = had ? VxGBuffer{h.prev_t, h.prev_nid, h.prev_bid, nullptr, nullptr} : gd;
    ti.sh = (const float*)sh->dev; ti.cocg = (const float*)cc->dev; ti.luma = (const float*)lu->dev; ti.ao_sky = (const float*)ao->dev;
    if (p->pre_pass) {
        VxSvgfInitialIn ii{};
        ii.current = gd; ii.sh = ti.sh; ii.cocg = ti.cocg; ii.luma = ti.luma; ii.ao_sky = ti.ao_sky;
        const VxSvgfInitialOut io{h.pre[0], h.pre[1], h.pre[2], h.pre[3]};
        if ((rc = launch_svgf_initial(c, *cam, ii, io))) return rc;
        ti.sh = h.pre[0]; ti.cocg = h.pre[1]; ti.luma = h.pre[2]; ti.ao_sky = h.pre[3];
    }
    ti.prev_sh = h.temporal[prev][0]; ti.prev_cocg = h.temporal[prev][1]; ti.prev_utility = h.temporal[prev][2]; ti.prev_ao_sky = h.temporal[prev][3];
    const VxSvgfTemporalOut to{h.temporal[cur][0], h.temporal[cur][1], h.temporal[cur][2], h.temporal[cur][3]};
    if ((rc = launch_svgf_temporal(c, *cam, ti, tp, to))) return rc;
    VxSvgfVarianceIn vi{};
    vi.current = gd; vi.sh = to.sh; vi.cocg = to.cocg; vi.utility = to.utility;
    const VxSvgfVarianceParams vp{1, p->aggressive_disocclusion};
    const VxSvgfVarianceOut vo{h.var[0], h.var[1], h.var[2]};
    if ((rc = launch_svgf_variance(c, *cam, vi, vp, vo))) return rc;
    VxSvgfSpatialIn si{};
    si.current = gd; si.sh = vo.sh; si.cocg = vo.cocg; si.variance = vo.variance; si.ao_sky = to.ao_sky; si.temporal_utility = to.utility;
    for (int k = 0; k < 5; ++k) {
        VxSvgfSpatialParams sp{};
        sp.step = (p->wide ? 32 : 16) >> k;
        sp.large_kernel = p->large_kernel; sp.do_spatial = 1; sp.aggressive_disocclusion = p->aggressive_disocclusion;
        sp.color_phi_bias = p->color_phi_bias; sp.time = p->time; sp.resolution_scale = p->resolution_scale;
        // the last pass writes the caller's planes (a plane the caller did not ask for still has to exist for the pass before it)
        const VxSvgfSpatialOut so = k == 4 ? VxSvgfSpatialOut{(float*)osh->dev, (float*)occ->dev, (float*)ov->dev, (float*)oao->dev}
                                           : VxSvgfSpatialOut{h.pong[k & 1][0], h.pong[k & 1][1], h.pong[k & 1][2], h.pong[k & 1][3]};
        if ((rc = launch_svgf_spatial(c, *cam, si, sp, so))) return rc;
        si.sh = so.sh; si.cocg = so.cocg; si.variance = so.variance; si.ao_sky = so.ao_sky;
    }
    // this frame becomes the history of the next
    VX_CUDA(cudaMemcpyAsync(h.prev_t, gd.t, npx * 4, cudaMemcpyDeviceToDevice, c->stream));
    VX_CUDA(cudaMemcpyAsync(h.prev_nid, gd.normal_id, npx, cudaMemcpyDeviceToDevice, c->stream));
    VX_CUDA(cudaMemcpyAsync(h.prev_bid, gd.block_id, npx, cudaMemcpyDeviceToDevice, c->stream));
    std::memcpy(h.prev_view, p->view, sizeof h.prev_view);
    std::memcpy(h.prev_projection, p->projection, sizeof h.prev_projection);
    h.cur = cur;
    h.valid = true;
    if (c->opt_timing) {
        VX_CUDA(cudaEventRecord(c->ev1, c->stream));
        c->pass_timed = true;
    }
    return s.end(c);
}

int vxpt_shadow_temporal(vxpt_handle c, const VxCamera* cam, const VxShadowTemporalIn* in, const VxShadowTemporalParams* p, const VxShadowTemporalOut* out) {
    int rc = check_svgf(c, cam);
    if (rc) return rc;
    if (!in || !p || !out || !in->current.t || !in->current.normal_id || !in->previous.t || !in->shadow || !in->transversal || !in->prev_shadow || !in->prev_frames)
        return fail(VXPT_E_INVALID, "NULL argument (current t / normal_id, previous t, the shadow pass's planes and the previous temporal planes are required)");
    VX_CUDA(cudaSetDevice(c->device));
    SvgfIO s(c, cam);
    Plane *t = s.in(in->current.t, 4), *n = s.in(in->current.normal_id, 1), *pt = s.in(in->previous.t, 4);
    Plane *sh = s.in(in->shadow, 1), *tr = s.in(in->transversal, 4), *ps = s.in(in->prev_shadow, 4), *pf = s.in(in->prev_frames, 4);
    Plane *os = s.out(out->shadow, 4), *of = s.out(out->frames, 4);
    if ((rc = s.begin())) return rc;
    if (cam->row_end == cam->row_begin) return VXPT_OK;
    VxShadowTemporalIn id{};
    id.current = VxGBuffer{(float*)t->dev, (uint8_t*)n->dev, nullptr, nullptr, nullptr};
    id.previous = VxGBuffer{(float*)pt->dev, nullptr, nullptr, nullptr, nullptr};
    id.shadow = (const uint8_t*)sh->dev; id.transversal = (const float*)tr->dev; id.prev_shadow = (const float*)ps->dev; id.prev_frames = (const float*)pf->dev;
    const VxShadowTemporalOut od{(float*)os->dev, (float*)of->dev};
    if ((rc = timed_launch(c, [&] { return launch_shadow_temporal(c, *cam, id, *p, od); }))) return rc;
    return s.end(c);
}

int vxpt_shadow_filter(vxpt_handle c, const VxCamera* cam, const VxShadowFilterIn* in, const VxShadowFilterParams* p, float* out) {
    int rc = check_svgf(c, cam);
    if (rc) return rc;
    if (!in || !p || !out || !in->current.t || !in->current.normal_id || !in->shadow || !in->transversal || !in->frames)
        return fail(VXPT_E_INVALID, "NULL argument (t, normal_id, the temporal shadow / frame planes, the transversal plane and the output are required)");
    VX_CUDA(cudaSetDevice(c->device));
    SvgfIO s(c, cam);
    Plane *t = s.in(in->current.t, 4), *n = s.in(in->current.normal_id, 1);
    Plane *sh = s.in(in->shadow, 4), *tr = s.in(in->transversal, 4), *fr = s.in(in->frames, 4);
    Plane* o = s.out(out, 4);
    if ((rc = s.begin())) return rc;
    if (cam->row_end == cam->row_begin) return VXPT_OK;
    VxShadowFilterIn id{};
    id.current = VxGBuffer{(float*)t->dev, (uint8_t*)n->dev, nullptr, nullptr, nullptr};
    id.shadow = (const float*)sh->dev; id.transversal = (const float*)tr->dev; id.frames = (const float*)fr->dev;
    if ((rc = timed_launch(c, [&] { return launch_shadow_filter(c, *cam, id, *p, (float*)o->dev); }))) return rc;
    return s.end(c);
}

int vxpt_shadow_filter_frame(vxpt_handle c, const VxCamera* cam, const VxGBuffer* g, const VxShadowOut* shadow, const VxShadowFrameParams* p, float* out) {
    int rc = check_svgf(c, cam);
    if (rc) return rc;
    if (!g || !shadow || !p || !out || !g->t || !g->normal_id || !shadow->shadow || !shadow->transversal)
        return fail(VXPT_E_INVALID, "NULL argument (G-buffer t / normal_id, the shadow pass's shadow / transversal planes and the output are required)");
    if (cam->row_begin != 0 || cam->row_end != cam->height) return fail(VXPT_E_INVALID, "vxpt_shadow_filter_frame filters whole frames (row_begin = 0, row_end = height)");
    VX_CUDA(cudaSetDevice(c->device));
    const int W = cam->width, H = cam->height;
    const size_t npx = (size_t)W * H, plane = (npx * 4 + 255) & ~(size_t)255;
    vxpt_ctx::ShadowHistory& h = c->shadow_hist;
    const bool had = h.buf && h.width == W && h.height == H && h.valid && !p->reset_history;
    if (!(h.buf && h.width == W && h.height == H)) {
        VX_CUDA(cudaStreamSynchronize(c->stream));
        if (h.buf) cudaFree(h.buf);
        h = vxpt_ctx::ShadowHistory();
        if (cudaMalloc(&h.buf, 5 * plane) != cudaSuccess) {
            cudaGetLastError();
            h.buf = nullptr;
            return fail(VXPT_E_NOMEM, "device allocation of the shadow-filter history failed");
        }
        char* q = (char*)h.buf;
        h.prev_t = (float*)q; h.shadow[0] = (float*)(q + plane); h.frames[0] = (float*)(q + 2 * plane); h.shadow[1] = (float*)(q + 3 * plane);
        h.frames[1] = (float*)(q + 4 * plane);
        h.width = W; h.height = H;
    }
    SvgfIO s(c, cam);
    Plane *t = s.in(g->t, 4), *n = s.in(g->normal_id, 1), *sh = s.in(shadow->shadow, 1), *tr = s.in(shadow->transversal, 4);
    Plane* o = s.out(out, 4);
    if ((rc = s.begin())) return rc;
    const VxGBuffer gd{(float*)t->dev, (uint8_t*)n->dev, nullptr, nullptr, nullptr};
    const int prev = h.cur, cur = h.cur ^ 1;
    VxShadowTemporalParams tp{};
    if (!had) {
        VX_CUDA(cudaMemsetAsync(h.shadow[prev], 0, npx * 4, c->stream));
        VX_CUDA(cudaMemsetAsync(h.frames[prev], 0, npx * 4, c->stream));
    }
    std::memcpy(tp.prev_view, had ? h.prev_view : p->view, sizeof tp.prev_view);
    std::memcpy(tp.prev_projection, had ? h.prev_projection : p->projection, sizeof tp.prev_projection);
    if (c->opt_timing) VX_CUDA(cudaEventRecord(c->ev0, c->stream));
    VxShadowTemporalIn ti{};
    ti.current = gd;
    ti.previous = VxGBuffer{had ? h.prev_t : gd.t, nullptr, nullptr, nullptr, nullptr};
    ti.shadow = (const uint8_t*)sh->dev; ti.transversal = (const float*)tr->dev; ti.prev_shadow = h.shadow[prev]; ti.prev_frames = h.frames[prev];
    const VxShadowTemporalOut to{h.shadow[cur], h.frames[cur]};
    if ((rc = launch_shadow_temporal(c, *cam, ti, tp, to))) return rc;
    if (p->spatial) {
        VxShadowFilterIn fi{};
        fi.current = gd; fi.shadow = to.shadow; fi.transversal = ti.transversal; fi.frames = to.frames;
        const VxShadowFilterParams fp{p->filter_scale};
        if ((rc = launch_shadow_filter(c, *cam, fi, fp, (float*)o->dev))) return rc;
    } else {
        VX_CUDA(cudaMemcpyAsync(o->dev, to.shadow, npx * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
    VX_CUDA(cudaMemcpyAsync(h.prev_t, gd.t, npx * 4, cudaMemcpyDeviceToDevice, c->stream));
    std::memcpy(h.prev_view, p->view, sizeof h.prev_view);
    std::memcpy(h.prev_projection, p->projection, sizeof h.prev_projection);
    h.cur = cur;
    h.valid = true;
    if (c->opt_timing) {
        VX_CUDA(cudaEventRecord(c->ev1, c->stream));
        c->pass_timed = true;
    }
    return s.end(c);
}

// ----------------------------------------------------------------------------------------------------- other DF consumers
int vxpt_trace_rays(vxpt_handle c, const float* origins, const float* directions, int n, int max_iterations, float* t, uint8_t* normal_id,
                    uint8_t* block_id, int16_t* hit_voxel) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (!origins || !directions) return fail(VXPT_E_INVALID, "NULL argument");
    if (n < 0 || n > (1 << 26) || max_iterations < 0) return fail(VXPT_E_INVALID, "bad ray count / max_iterations");
    if (n == 0) return VXPT_OK;
    VX_CUDA(cudaSetDevice(c->device));
    VxCamera batch{};  // the rays as one row of n "pixels", so the plane staging of the passes applies
    batch.width = n; batch.height = 1; batch.row_begin = 0; batch.row_end = 1;
    PassIO io(c, &batch);
    Plane o, d, pt, pn, pb, pv;
    io.add(o, origins, 12); io.add(d, directions, 12); io.add(pt, t, 4); io.add(pn, normal_id, 1); io.add(pb, block_id, 1); io.add(pv, hit_voxel, 6);
    if ((rc = io.resolve())) return rc;
    if ((rc = io.upload(o)) || (rc = io.upload(d))) return rc;
    if (c->opt_timing) VX_CUDA(cudaEventRecord(c->ev0, c->stream));
    if ((rc = launch_rays(c, (const float*)o.dev, (const float*)d.dev, n, max_iterations, (float*)pt.dev, (uint8_t*)pn.dev, (uint8_t*)pb.dev,
                          (int16_t*)pv.dev)))
        return rc;
    if (c->opt_timing) {
        VX_CUDA(cudaEventRecord(c->ev1, c->stream));
        c->pass_timed = true;
    }
    bool any = false;
    if ((rc = io.download(pt, any)) || (rc = io.download(pn, any)) || (rc = io.download(pb, any)) || (rc = io.download(pv, any))) return rc;
    if (any) VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXPT_OK;
}

int vxpt_player_shadowed(vxpt_handle c, const float camera_pos[3], const float sun_dir[3], int* shadowed) {
    if (!camera_pos || !sun_dir || !shadowed) return fail(VXPT_E_INVALID, "NULL argument");
    // PostProcessingVert.glsl:48-50: L = length(u_VertSunDir); D = u_VertSunDir / L   (fp32, no FMA: this file is built with -fmad=false)
    const float L = sqrtf((sun_dir[0] * sun_dir[0] + sun_dir[1] * sun_dir[1]) + sun_dir[2] * sun_dir[2]);
    const float D[3] = {sun_dir[0] / L, sun_dir[1] / L, sun_dir[2] / L};
    float T = -1.0f;
    const int rc = vxpt_trace_rays(c, camera_pos, D, 1, 350, &T, nullptr, nullptr, nullptr);
    if (rc) return rc;
    *shadowed = T > 0.0f ? 1 : 0;
    return VXPT_OK;
}

int vxpt_estimate_ambient_sound(vxpt_handle c, const float player_pos[3], int frame, uint32_t* sky_level_aggregate, uint32_t* per_invocation) {
    int rc = check_ready(c);
    if (rc) return rc;
    if (!player_pos || !sky_level_aggregate) return fail(VXPT_E_INVALID, "NULL argument");
    if (frame < 0) return fail(VXPT_E_INVALID, "frame < 0");
    VX_CUDA(cudaSetDevice(c->device));
    Arena arena(c);
    if ((rc = arena.reserve(256))) return rc;
    unsigned* d = (unsigned*)arena.take(33 * sizeof(unsigned));  // [0] = SkyLevelAggregate (cleared per dispatch, Pipeline.cpp:1931-1934), [1..32]
    VX_CUDA(cudaMemsetAsync(d, 0, 33 * sizeof(unsigned), c->stream));
    if ((rc = launch_ambient(c, player_pos, frame, d, d + 1))) return rc;
    unsigned h[33];
    VX_CUDA(cudaMemcpyAsync(h, d, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));  // glGetBufferSubData (Pipeline.cpp:1926-1929)
    *sky_level_aggregate = h[0];
    if (per_invocation)
        for (int k = 0; k < 32; ++k) per_invocation[k] = h[1 + k];
    return VXPT_OK;
}

// ----------------------------------------------------------------------------------------------------- passes
int vxpt_trace_primary(vxpt_handle c, const VxCamera* cam, const VxPrimaryParams* p, const VxGBuffer* out) {
    int rc = check_ready(c);
    if (rc) return rc;
    if ((rc = check_camera(cam))) return rc;
    if (!p || !out) return fail(VXPT_E_INVALID, "NULL argument");
    if ((rc = check_primary(c, p))) return rc;
    VX_CUDA(cudaSetDevice(c->device));
    PassIO io(c, cam);
    Plane t, nid, bid, it, hv;
    io.add(t, out->t, px_bytes(c, 4, 2)); io.add(nid, out->normal_id, 1); io.add(bid, out->block_id, 1); io.add(it, out->inv_t, 4); io.add(hv, out->hit_voxel, 6);
    if ((rc = io.resolve())) return rc;
    VxGBuffer dev{(float*)t.dev, (uint8_t*)nid.dev, (uint8_t*)bid.dev, (float*)it.dev, (int16_t*)hv.dev};
    if (cam->row_end == cam->row_begin) return VXPT_OK;
    c->frame_counter++;  // a primary pass opens a frame (scene replica rotation)
    if (c->opt_timing) VX_CUDA(cudaEventRecord(c->ev0, c->stream));
    if ((rc = launch_primary(c, *cam, *p, dev))) return rc;
    if (c->opt_timing) {
        VX_CUDA(cudaEventRecord(c->ev1, c->stream));
        c->pass_timed = true;
    }
    bool any = false;
    for (Plane* pl : io.planes)
        if ((rc = io.download(*pl, any))) return rc;
    if (any) VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXPT_OK;
}

int vxpt_trace_shadow(vxpt_handle c, const VxCamera* cam, const VxGBuffer* g, const VxShadowParams* p, const VxShadowOut* out) {
    int rc = check_ready(c);
    if (rc) return rc;
    if ((rc = check_camera(cam))) return rc;
    if (!g || !p || !out || !g->t || !g->normal_id) return fail(VXPT_E_INVALID, "NULL argument (G-buffer t and normal_id are required)");
    if ((rc = check_shadow(c, p))) return rc;
    VX_CUDA(cudaSetDevice(c->device));
    PassIO io(c, cam);
    Plane t, nid, sh, tr;
    io.add(t, g->t, px_bytes(c, 4, 2)); io.add(nid, g->normal_id, 1); io.add(sh, out->shadow, 1); io.add(tr, out->transversal, px_bytes(c, 4, 2));
    if ((rc = io.resolve())) return rc;
    if (cam->row_end == cam->row_begin) return VXPT_OK;
    if ((rc = io.upload(t)) || (rc = io.upload(nid))) return rc;
    VxGBuffer gd{(float*)t.dev, (uint8_t*)nid.dev, nullptr, nullptr, nullptr};
    VxShadowOut od{(uint8_t*)sh.dev, (float*)tr.dev};
    if (c->opt_timing) VX_CUDA(cudaEventRecord(c->ev0, c->stream));
    if ((rc = launch_shadow(c, *cam, gd, *p, od))) return rc;
    if (c->opt_timing) {
        VX_CUDA(cudaEventRecord(c->ev1, c->stream));
        c->pass_timed = true;
    }
    bool any = false;
    if ((rc = io.download(sh, any)) || (rc = io.download(tr, any))) return rc;
    if (any) VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXPT_OK;
}

int vxpt_trace_diffuse(vxpt_handle c, const VxCamera* cam, const VxGBuffer* g, const VxDiffuseParams* p, const VxDiffuseOut* out) {
    int rc = check_ready(c);
    if (rc) return rc;
    if ((rc = check_camera(cam))) return rc;
    if (!g || !p || !out || !g->t || !g->normal_id) return fail(VXPT_E_INVALID, "NULL argument (G-buffer t and normal_id are required)");
    if ((rc = check_diffuse(c, p))) return rc;
    VX_CUDA(cudaSetDevice(c->device));
    PassIO io(c, cam);
    Plane t, nid, sh, cg, lu, ao;
    io.add(t, g->t, px_bytes(c, 4, 2)); io.add(nid, g->normal_id, 1);
    io.add(sh, out->sh, px_bytes(c, 16, 8)); io.add(cg, out->cocg, px_bytes(c, 8, 4)); io.add(lu, out->luma, px_bytes(c, 4, 2));
    io.add(ao, out->ao_sky, px_bytes(c, 8, 2));
    if ((rc = io.resolve())) return rc;
    if (cam->row_end == cam->row_begin) return VXPT_OK;
    if ((rc = io.upload(t)) || (rc = io.upload(nid))) return rc;
    VxGBuffer gd{(float*)t.dev, (uint8_t*)nid.dev, nullptr, nullptr, nullptr};
    VxDiffuseOut od{(float*)sh.dev, (float*)cg.dev, (float*)lu.dev, (float*)ao.dev};
    if (c->opt_timing) VX_CUDA(cudaEventRecord(c->ev0, c->stream));
    if ((rc = launch_diffuse(c, *cam, gd, *p, od))) return rc;
    if (c->opt_timing) {
        VX_CUDA(cudaEventRecord(c->ev1, c->stream));
        c->pass_timed = true;
    }
    bool any = false;
    if ((rc = io.download(sh, any)) || (rc = io.download(cg, any)) || (rc = io.download(lu, any)) || (rc = io.download(ao, any))) return rc;
    if (any) VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXPT_OK;
}

int vxpt_trace_reflection(vxpt_handle c, const VxCamera* cam, const VxGBuffer* g, const VxReflectionIn* in, const VxReflectionParams* p,
                          const VxReflectionOut* out) {
    int rc = check_ready(c);
    if (rc) return rc;
    if ((rc = check_camera(cam))) return rc;
    if (!g || !in || !p || !out || !g->t || !g->normal_id || !in->sh || !in->cocg)
        return fail(VXPT_E_INVALID, "NULL argument (G-buffer t, normal_id and the GI sh / cocg planes are required)");
    if (!in->g_pbr && !g->block_id) return fail(VXPT_E_INVALID, "without g_pbr the G-buffer block_id plane is required");
    if ((rc = check_reflection(c, p))) return rc;
    if (cam->interleave_n > 1) return fail(VXPT_E_INVALID, "the reflection pass reads G-buffer rows next to its own (jittered coordinate): contiguous row slabs only");
    VX_CUDA(cudaSetDevice(c->device));
    PassIO io(c, cam);
    Plane t, nid, bid, gn, gp, sh, cg, col, hd, em;
    io.add(t, g->t, px_bytes(c, 4, 2)); io.add(nid, g->normal_id, 1); io.add(bid, g->block_id, 1);
    io.add(gn, in->g_normal, 12); io.add(gp, in->g_pbr, 16); io.add(sh, in->sh, px_bytes(c, 16, 8)); io.add(cg, in->cocg, px_bytes(c, 8, 4));
    io.add(col, out->color, px_bytes(c, 16, 8)); io.add(hd, out->hit_distance, px_bytes(c, 4, 2)); io.add(em, out->emissive_mask, 1);
    if ((rc = io.resolve())) return rc;
    if (cam->row_end == cam->row_begin) return VXPT_OK;
    // distance and normal id are read at the Halton-jittered coordinate (rows beyond the slab, GL_REPEAT at the frame edge): whole planes
    if ((rc = io.upload_all(t)) || (rc = io.upload_all(nid)) || (rc = io.upload(bid)) || (rc = io.upload(gn)) || (rc = io.upload(gp)) ||
        (rc = io.upload(sh)) || (rc = io.upload(cg)))
        return rc;
    VxGBuffer gd{(float*)t.dev, (uint8_t*)nid.dev, (uint8_t*)bid.dev, nullptr, nullptr};
    VxReflectionIn id{(const float*)gn.dev, (const float*)gp.dev, (const float*)sh.dev, (const float*)cg.dev};
    VxReflectionOut od{(float*)col.dev, (float*)hd.dev, (uint8_t*)em.dev};
    if (c->opt_timing) VX_CUDA(cudaEventRecord(c->ev0, c->stream));
    if ((rc = launch_reflection(c, *cam, gd, id, *p, od))) return rc;
    if (c->opt_timing) {
        VX_CUDA(cudaEventRecord(c->ev1, c->stream));
        c->pass_timed = true;
    }
    bool any = false;
    if ((rc = io.download(col, any)) || (rc = io.download(hd, any)) || (rc = io.download(em, any))) return rc;
    if (any) VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXPT_OK;
}

// --------------------------------------------------------------------------------------------- one whole frame
}  // extern "C"
static int render_frame_impl(vxpt_handle c, const VxCamera* cam, const VxFrameParams* p, const VxFrameOut* out, bool wait) {
    int rc = check_ready(c);
    if (rc) return rc;
    if ((rc = check_camera(cam))) return rc;
    if (!p || !out || !p->primary) return fail(VXPT_E_INVALID, "NULL argument (primary parameters are required)");
    if ((rc = check_primary(c, p->primary))) return rc;
    if (p->shadow && (rc = check_shadow(c, p->shadow))) return rc;
    if (p->diffuse && (rc = check_diffuse(c, p->diffuse))) return rc;
    if (p->reflection) {
        if (!p->diffuse) return fail(VXPT_E_INVALID, "the reflection pass reads the GI planes: diffuse parameters are required");
        if ((rc = check_reflection(c, p->reflection))) return rc;
        if (cam->interleave_n > 1) return fail(VXPT_E_INVALID, "the reflection pass reads G-buffer rows next to its own (jittered coordinate): contiguous row slabs only");
    }
    if (p->material && (rc = check_material(c, cam, p->material))) return rc;
    VX_CUDA(cudaSetDevice(c->device));
    const bool secondary = p->shadow || p->diffuse || p->reflection;
    PassIO io(c, cam);
    Plane t, nid, bid, it, hv, sh, tr, dsh, dcg, dlu, dao, gn, gp, col, hd, em, mal, mnm, mpb, mao;
    const bool mat = p->material != nullptr;
    const bool mat_feeds_reflection = mat && p->reflection && (p->material->update_this_frame || p->material->pom);
    // planes a later pass reads exist in the handle's arena even when the caller does not want them back
    io.add(t, out->gbuffer.t, px_bytes(c, 4, 2), secondary);
    io.add(nid, out->gbuffer.normal_id, 1, secondary);
    io.add(bid, out->gbuffer.block_id, 1, (p->reflection && !p->g_pbr) || mat);
    io.add(it, out->gbuffer.inv_t, 4, mat);
    io.add(hv, out->gbuffer.hit_voxel, 6);
    if (p->shadow) {
        io.add(sh, out->shadow.shadow, 1);
        io.add(tr, out->shadow.transversal, px_bytes(c, 4, 2));
    }
    if (p->diffuse) {
        io.add(dsh, out->diffuse.sh, px_bytes(c, 16, 8), p->reflection != nullptr);
        io.add(dcg, out->diffuse.cocg, px_bytes(c, 8, 4), p->reflection != nullptr);
        io.add(dlu, out->diffuse.luma, px_bytes(c, 4, 2));
        io.add(dao, out->diffuse.ao_sky, px_bytes(c, 8, 2));
    }
    if (mat) {  // the planes the reflection pass reads exist in the arena even when the caller does not want them back
        io.add(mal, out->material.albedo, 12);
        io.add(mnm, out->material.normal, 12, mat_feeds_reflection && !p->g_normal);
        io.add(mpb, out->material.pbr, 16, mat_feeds_reflection && !p->g_pbr);
        io.add(mao, out->material.texture_ao, 4);
    }
    if (p->reflection) {
        io.add(gn, p->g_normal, 12);
        io.add(gp, p->g_pbr, 16);
        io.add(col, out->reflection.color, px_bytes(c, 16, 8));
        io.add(hd, out->reflection.hit_distance, px_bytes(c, 4, 2));
        io.add(em, out->reflection.emissive_mask, 1);
    }
    if ((rc = io.resolve())) return rc;
    const int rows = cam->row_end - cam->row_begin;
    if (rows == 0) return VXPT_OK;
    if ((rc = io.upload(gn)) || (rc = io.upload(gp))) return rc;
    bool any_host = false;
    for (Plane* pl : io.planes) any_host = any_host || (pl->user && pl->staged && pl != &gn && pl != &gp);
    const VxGBuffer gd{(float*)t.dev, (uint8_t*)nid.dev, (uint8_t*)bid.dev, (float*)it.dev, (int16_t*)hv.dev};
    const VxShadowOut sd{(uint8_t*)sh.dev, (float*)tr.dev};
    const VxDiffuseOut dd{(float*)dsh.dev, (float*)dcg.dev, (float*)dlu.dev, (float*)dao.dev};
    const VxMaterialOut md{(float*)mal.dev, (float*)mnm.dev, (float*)mpb.dev, (float*)mao.dev};
    const VxReflectionIn ri{(const float*)(p->g_normal || !mat_feeds_reflection ? gn.dev : mnm.dev),
                            (const float*)(p->g_pbr || !mat_feeds_reflection ? gp.dev : mpb.dev), (const float*)dsh.dev, (const float*)dcg.dev};
    const VxReflectionOut rd{(float*)col.dev, (float*)hd.dev, (uint8_t*)em.dev};
    c->frame_counter++;
    if (c->opt_timing) VX_CUDA(cudaEventRecord(c->ev0, c->stream));
    // Host outputs leave through the copy stream while the next pass traces: the G-buffer planes during the shadow pass, the shadow
    // planes during GI, and the GI pass (the largest planes) in two row slabs so that the first half is on its way while the second
    // half traces.  Smaller pieces would starve the copy engine: a 135-row slab of the three passes takes longer to trace (latency-
    // bound kernels) than to copy (measured r01g, 1080 rows: this schedule 1.21 ms, 2 / 4 / 8 uniform slabs of all passes 1.45 / 1.38 / 1.48 ms).
    int n_ev = 0;
    // Once a copy to the caller's planes is queued, the frame owns the staging arena and those planes until the copy stream has drained —
    // also when a later launch fails and this function returns early: the guard then leaves the frame pending, so the next call on the
    // handle (check_ready -> finish_pending_frame) waits for the copies before anything reuses or frees the arena.
    struct PendingGuard {
        vxpt_ctx* c;
        bool queued = false, settled = false;
        ~PendingGuard() { if (queued && !settled) c->frame_pending = true; }
    } guard{c};
    auto copy_out = [&](std::initializer_list<Plane*> pls, int rb, int re) -> int {
        if (!any_host) return VXPT_OK;
        bool some = false;
        for (Plane* pl : pls) some = some || (pl->user && pl->staged);
        if (!some) return VXPT_OK;
        guard.queued = true;
        cudaEvent_t ev = c->ev_slab[n_ev++ & 7];
        VX_CUDA(cudaEventRecord(ev, c->stream));
        VX_CUDA(cudaStreamWaitEvent(c->copy_stream, ev, 0));
        for (Plane* pl : pls)
            if (int rc2 = io.download_rows(*pl, rb, re, c->copy_stream)) return rc2;
        return VXPT_OK;
    };
    if ((rc = launch_primary(c, *cam, *p->primary, gd))) return rc;
    if (p->reflection && rows < cam->height) {
        // the reflection pass reads distance / normal id at the jittered coordinate: trace the halo rows of the slab too (one
        // extra primary row per slab edge instead of a halo exchange); GL_REPEAT wraps at the frame edge.  They land in the G-buffer planes
        // outside [row_begin, row_end) (in the arena for host planes; for device planes the caller's rows there are overwritten with the
        // same values any other slab's primary pass writes) and are not copied out.
        int below = 0, above = 0;
        reflection_halo(p->reflection, &below, &above);
        const int H = cam->height;
        std::vector<int> halo;
        auto want = [&](int row) {
            row = (row % H + H) % H;
            if (row < cam->row_begin || row >= cam->row_end) halo.push_back(row);
        };
        for (int k = 1; k <= below; ++k) want(cam->row_begin - k);
        for (int k = 0; k < above; ++k) want(cam->row_end + k);
        std::sort(halo.begin(), halo.end());
        halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
        for (size_t a = 0; a < halo.size();) {  // one launch per run of consecutive rows
            size_t b = a;
            while (b + 1 < halo.size() && halo[b + 1] == halo[b] + 1) ++b;
            VxCamera hc = *cam;
            hc.row_begin = halo[a];
            hc.row_end = halo[b] + 1;
            a = b + 1;
            if ((rc = launch_primary(c, hc, *p->primary, gd))) return rc;
        }
    }
    if ((rc = copy_out({&t, &nid, &bid, &it, &hv}, cam->row_begin, cam->row_end))) return rc;
    if (mat) {
        if (!p->material->update_this_frame && !p->material->pom)  // fragments discard (all but lava pixels): staged planes must keep what the caller's hold
            if ((rc = io.upload(mal)) || (rc = io.upload(mnm)) || (rc = io.upload(mpb)) || (rc = io.upload(mao))) return rc;
        if ((rc = launch_gbuffer(c, *cam, gd, *p->material, md))) return rc;
        if ((rc = copy_out({&mal, &mnm, &mpb, &mao}, cam->row_begin, cam->row_end))) return rc;
    }
    if (p->shadow) {
        if ((rc = launch_shadow(c, *cam, gd, *p->shadow, sd))) return rc;
        if ((rc = copy_out({&sh, &tr}, cam->row_begin, cam->row_end))) return rc;
    }
    if (p->diffuse) {
        const int nslab = (any_host && rows >= 256) ? 2 : 1;
        for (int s = 0; s < nslab; ++s) {
            VxCamera sc = *cam;
            sc.row_begin = cam->row_begin + ((rows * s / nslab) & ~7);
            sc.row_end = (s + 1 == nslab) ? cam->row_end : cam->row_begin + ((rows * (s + 1) / nslab) & ~7);
            if ((rc = launch_diffuse(c, sc, gd, *p->diffuse, dd))) return rc;
            if ((rc = copy_out({&dsh, &dcg, &dlu, &dao}, sc.row_begin, sc.row_end))) return rc;
        }
    }
    if (p->reflection) {
        if ((rc = launch_reflection(c, *cam, gd, ri, *p->reflection, rd))) return rc;
        if ((rc = copy_out({&col, &hd, &em}, cam->row_begin, cam->row_end))) return rc;
    }

    if (c->opt_timing) {
        VX_CUDA(cudaEventRecord(c->ev1, c->stream));
        c->pass_timed = true;
    }
    if (any_host) {
        if (stream_capturing(c)) {
            // The frame is being recorded into a CUDA graph (vxpt_render_frame_async on a capturing stream): the copy stream joined the
            // capture at the first copy-out and must rejoin the handle's stream before the capture ends.  A replay is then complete —
            // kernels AND host planes — when the handle's stream is (vxpt_sync); nothing is pending on the host side.
            if (wait) return fail(VXPT_E_STATE, "vxpt_render_frame would block inside a stream capture: use vxpt_render_frame_async");
            if (guard.queued) {
                cudaEvent_t ev = c->ev_slab[n_ev++ & 7];
                VX_CUDA(cudaEventRecord(ev, c->copy_stream));
                VX_CUDA(cudaStreamWaitEvent(c->stream, ev, 0));
            }
            guard.settled = true;
        } else if (wait) {
            VX_CUDA(cudaStreamSynchronize(c->copy_stream));
            guard.settled = true;
        } else {
            c->frame_pending = true;
        }
    }
    return VXPT_OK;
}
extern "C" {

int vxpt_render_frame(vxpt_handle c, const VxCamera* cam, const VxFrameParams* p, const VxFrameOut* out) {
    return render_frame_impl(c, cam, p, out, true);
}
int vxpt_render_frame_async(vxpt_handle c, const VxCamera* cam, const VxFrameParams* p, const VxFrameOut* out) {
    return render_frame_impl(c, cam, p, out, false);
}
int vxpt_frame_wait(vxpt_handle c) {
    if (!c) return fail(VXPT_E_INVALID, "handle is NULL");
    return finish_pending_frame(c);
}

// ------------------------------------------------------------------------------------- peer-to-peer slab gather
}  // extern "C"
namespace vxpt {
#ifndef VXPT_HOST_SHADOW
// release at system scope: every store of the kernels that ran before on this stream is visible to a peer that observes the flag
__device__ __forceinline__ void store_release_sys(uint32_t* flag, uint32_t value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}
__device__ __forceinline__ uint32_t load_acquire_sys(const uint32_t* f) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#endif  // tests/host_shadow defines the three for the g++ build (plain loads / stores, a clock)
__global__ void signal_kernel(uint32_t* flag, uint32_t value) { store_release_sys(flag, value); }
__device__ __forceinline__ void spin_until(const uint32_t* f, uint32_t at_least, unsigned long long timeout_ns, unsigned* err) {
    const unsigned long long t0 = global_timer_ns();
    while (true) {
        const uint32_t v = load_acquire_sys(f);
        if ((int32_t)(v - at_least) >= 0) break;
        if (global_timer_ns() - t0 > timeout_ns) {
            atomicExch(err, 1u);
            break;
        }
        __nanosleep(100);
    }
}
__global__ void signal_next_kernel(uint32_t* flag, uint32_t* counter) {
    const uint32_t v = *counter + 1u;
    *counter = v;
    store_release_sys(flag, v);
}
__global__ void wait_next_kernel(const uint32_t* flags, int n, int stride_words, uint32_t* counter, int lag, unsigned long long timeout_ns, unsigned* err) {
    __shared__ uint32_t s_target;
    if (threadIdx.x == 0) {
        const uint32_t v = *counter + 1u;
        *counter = v;
        s_target = v - (uint32_t)lag;
    }
    __syncthreads();
    const uint32_t target = s_target;
    if ((int32_t)target < 1 || (int)threadIdx.x >= n) return;
    spin_until(flags + (size_t)threadIdx.x * stride_words, target, timeout_ns, err);
}
__global__ void wait_all_kernel(const uint32_t* flags, int n, int stride_words, uint32_t at_least, unsigned long long timeout_ns, unsigned* err) {
    if ((int)threadIdx.x >= n) return;
    spin_until(flags + (size_t)threadIdx.x * stride_words, at_least, timeout_ns, err);
}
}  // namespace vxpt
extern "C" {

int vxpt_shared_alloc(vxpt_handle c, size_t bytes, void** dptr, uint8_t handle_out[VXPT_SHARED_HANDLE_BYTES]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == VXPT_SHARED_HANDLE_BYTES, "IPC handle size");
    if (!c || !dptr || !handle_out || bytes == 0) return fail(VXPT_E_INVALID, "bad argument");
    VX_CUDA(cudaSetDevice(c->device));
    void* ptr = nullptr;
    if (cudaMalloc(&ptr, bytes) != cudaSuccess) {
        cudaGetLastError();
        return fail(VXPT_E_NOMEM, "shared slab buffer allocation failed");
    }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaMemset(ptr, 0, bytes);
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, ptr);
    if (e != cudaSuccess) {
        cudaFree(ptr);
        return cuda_fail(e, "cudaIpcGetMemHandle", __FILE__, __LINE__);
    }
    std::memcpy(handle_out, &h, sizeof h);
    c->shared.push_back({ptr, true});
    *dptr = ptr;
    return VXPT_OK;
}

int vxpt_shared_open(vxpt_handle c, const uint8_t handle[VXPT_SHARED_HANDLE_BYTES], void** dptr) {
    if (!c || !dptr || !handle) return fail(VXPT_E_INVALID, "bad argument");
    VX_CUDA(cudaSetDevice(c->device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    void* ptr = nullptr;
    VX_CUDA(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->shared.push_back({ptr, false});
    *dptr = ptr;
    return VXPT_OK;
}

int vxpt_shared_close(vxpt_handle c, void* dptr) {
    if (!c || !dptr) return fail(VXPT_E_INVALID, "bad argument");
    VX_CUDA(cudaSetDevice(c->device));
    for (size_t k = 0; k < c->shared.size(); ++k)
        if (c->shared[k].ptr == dptr) {
            VX_CUDA(cudaStreamSynchronize(c->stream));
            const bool owned = c->shared[k].owned;
            c->shared.erase(c->shared.begin() + k);
            if (owned) VX_CUDA(cudaFree(dptr));
            else VX_CUDA(cudaIpcCloseMemHandle(dptr));
            return VXPT_OK;
        }
    return fail(VXPT_E_INVALID, "pointer was not returned by vxpt_shared_alloc / vxpt_shared_open of this handle");
}

int vxpt_copy_async(vxpt_handle c, void* dst, const void* src, size_t bytes, void* cuda_stream) {
    if (!c || !dst || !src) return fail(VXPT_E_INVALID, "bad argument");
    if (bytes == 0) return VXPT_OK;
    VX_CUDA(cudaSetDevice(c->device));
    VX_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, cuda_stream ? (cudaStream_t)cuda_stream : c->stream));
    return VXPT_OK;
}

int vxpt_signal(vxpt_handle c, uint32_t* flag, uint32_t value, void* cuda_stream) {
    if (!c || !flag) return fail(VXPT_E_INVALID, "bad argument");
    VX_CUDA(cudaSetDevice(c->device));
    VX_LAUNCH(signal_kernel, dim3(1), 1, cuda_stream ? (cudaStream_t)cuda_stream : c->stream, flag, value);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

int vxpt_signal_next(vxpt_handle c, uint32_t* flag, uint32_t* counter, void* cuda_stream) {
    if (!c || !flag || !counter) return fail(VXPT_E_INVALID, "bad argument");
    VX_CUDA(cudaSetDevice(c->device));
    VX_LAUNCH(signal_next_kernel, dim3(1), 1, cuda_stream ? (cudaStream_t)cuda_stream : c->stream, flag, counter);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

int vxpt_wait_next(vxpt_handle c, const uint32_t* flags, int n, int stride_words, uint32_t* counter, int lag, int timeout_ms, void* cuda_stream) {
    if (!c || !flags || !counter || n <= 0 || n > 1024 || stride_words <= 0 || lag < 0 || timeout_ms <= 0) return fail(VXPT_E_INVALID, "bad argument");
    VX_CUDA(cudaSetDevice(c->device));
    VX_LAUNCH(wait_next_kernel, dim3(1), ((n + 31) / 32) * 32, cuda_stream ? (cudaStream_t)cuda_stream : c->stream, flags, n, stride_words, counter, lag,
              (unsigned long long)timeout_ms * 1000000ull, c->d_wait_err);
    c->launches += 1;
    c->waits_issued = true;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

int vxpt_wait_all(vxpt_handle c, const uint32_t* flags, int n, int stride_words, uint32_t at_least, int timeout_ms, void* cuda_stream) {
    if (!c || !flags || n <= 0 || n > 1024 || stride_words <= 0 || timeout_ms <= 0) return fail(VXPT_E_INVALID, "bad argument");
    VX_CUDA(cudaSetDevice(c->device));
    VX_LAUNCH(wait_all_kernel, dim3(1), ((n + 31) / 32) * 32, cuda_stream ? (cudaStream_t)cuda_stream : c->stream, flags, n, stride_words, at_least,
              (unsigned long long)timeout_ms * 1000000ull, c->d_wait_err);
    c->launches += 1;
    c->waits_issued = true;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

// ------------------------------------------------------------------------------------------------ sync / stats
static int check_wait_error(vxpt_ctx* c) {  // stream must be idle
    if (!c->waits_issued) return VXPT_OK;
    unsigned err = 0;
    VX_CUDA(cudaMemcpy(&err, c->d_wait_err, sizeof err, cudaMemcpyDeviceToHost));
    if (err) {
        VX_CUDA(cudaMemset(c->d_wait_err, 0, sizeof(unsigned)));
        return fail(VXPT_E_STATE, "vxpt_wait_all timed out: a peer never signalled its frame");
    }
    return VXPT_OK;
}

int vxpt_sync(vxpt_handle c) {
    if (!c) return fail(VXPT_E_INVALID, "handle is NULL");
    VX_CUDA(cudaSetDevice(c->device));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return check_wait_error(c);
}

int vxpt_get_stats(vxpt_handle c, VxStats* out) {
    if (!c || !out) return fail(VXPT_E_INVALID, "NULL argument");
    VX_CUDA(cudaSetDevice(c->device));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    if (int rc = check_wait_error(c)) return rc;
    DeviceCounters h;
    VX_CUDA(cudaMemcpy(&h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    if (c->pass_timed) {
        VX_CUDA(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
        c->pass_timed = false;
    }
    if (c->df_timed) {
        VX_CUDA(cudaEventElapsedTime(&c->df_build_ms, c->ev2, c->ev3));
        c->brick_pack_ms = 0.0f;  // (two back-to-back event records alone are 2.6 us apart)
        if (c->pack_timed) VX_CUDA(cudaEventElapsedTime(&c->brick_pack_ms, c->ev3, c->ev4));
        c->df_timed = false;
    }
    out->rays = h.rays;
    out->df_fetches = h.df_fetches;
    out->vox_fetches = h.vox_fetches;
    out->last_ms = c->last_ms;
    out->df_build_ms = c->df_build_ms;
    out->brick_pack_ms = c->brick_pack_ms;
    return VXPT_OK;
}

int vxpt_reset_stats(vxpt_handle c) {
    if (!c) return fail(VXPT_E_INVALID, "handle is NULL");
    VX_CUDA(cudaSetDevice(c->device));
    VX_CUDA(cudaMemsetAsync(c->d_counters, 0, sizeof(DeviceCounters), c->stream));
    return VXPT_OK;
}

int vxpt_launch_count(vxpt_handle c, uint64_t* n) {
    if (!c || !n) return fail(VXPT_E_INVALID, "NULL argument");
    *n = c->launches;
    return VXPT_OK;
}

int vxpt_stream(vxpt_handle c, void** s) {
    if (!c || !s) return fail(VXPT_E_INVALID, "NULL argument");
    *s = (void*)c->stream;
    return VXPT_OK;
}

int vxpt_reserve(vxpt_handle c, const VxCamera* cam, int max_gi_spp, size_t staging_bytes) {
    if (!c) return fail(VXPT_E_INVALID, "handle is NULL");
    if (int rc = check_camera(cam)) return rc;
    if (max_gi_spp < 0 || max_gi_spp > 64) return fail(VXPT_E_INVALID, "max_gi_spp outside 0..64");
    VX_CUDA(cudaSetDevice(c->device));
    if (int rc = finish_pending_frame(c)) return rc;
    if (max_gi_spp > 0) {
        const size_t slab_px = (size_t)(cam->row_end - cam->row_begin) * cam->width, frame_px = (size_t)cam->width * cam->height;
        if (int rc = grow_scratch(c, &c->d_queue, &c->queue_bytes, gi_scratch_bytes(slab_px, frame_px, max_gi_spp == 1), "the GI wavefront queue")) return rc;
    }
    if (staging_bytes)
        if (int rc = grow_scratch(c, &c->d_stage, &c->stage_bytes, staging_bytes, "the staging arena")) return rc;
    return VXPT_OK;
}

int vxpt_set_option(vxpt_handle c, int option, int value) {
    if (!c) return fail(VXPT_E_INVALID, "handle is NULL");
    switch (option) {
        case VXPT_OPT_TRAVERSAL_LAYOUT:
            if (value != 0 && value != 1) return fail(VXPT_E_INVALID, "layout must be 0 or 1");
            c->opt_layout = value;
            if (c->df_valid && c->steps_layout != value) {  // re-layout the step field for the new choice
                VX_CUDA(cudaSetDevice(c->device));
                int rc = launch_pack_bricks(c);
                return rc ? rc : refresh_replicas(c);
            }
            return VXPT_OK;
        case VXPT_OPT_GI_WAVEFRONT:
            if (value != 0 && value != 1) return fail(VXPT_E_INVALID, "wavefront must be 0 or 1");
            c->opt_wavefront = value;
            return VXPT_OK;
        case VXPT_OPT_REFLECTION_WAVEFRONT:
            if (value != 0 && value != 1) return fail(VXPT_E_INVALID, "reflection wavefront must be 0 or 1");
            c->opt_refl_wavefront = value;
            return VXPT_OK;
        case VXPT_OPT_SCENE_REPLICAS:
            if (value < 1 || value > 8) return fail(VXPT_E_INVALID, "replicas must be 1..8");
            c->opt_replicas = value;
            if (c->df_valid) {
                VX_CUDA(cudaSetDevice(c->device));
                return refresh_replicas(c);
            }
            return VXPT_OK;
        case VXPT_OPT_TIMING_EVENTS:
            c->opt_timing = value ? 1 : 0;
            if (!c->opt_timing) c->pass_timed = false;
            return VXPT_OK;
        case VXPT_OPT_TEXEL_FORMAT:
            if (value != 0 && value != 1) return fail(VXPT_E_INVALID, "texel format must be 0 or 1");
            c->opt_texel = value;
            return VXPT_OK;
        case VXPT_OPT_MATERIAL_QUAD_SHUFFLE:
            if (value != 0 && value != 1) return fail(VXPT_E_INVALID, "quad shuffle must be 0 or 1");
            c->opt_quad_shuffle = value;
            return VXPT_OK;
        case VXPT_OPT_DF_ALGO:
            if (value < 0 || value > 1) return fail(VXPT_E_INVALID, "df algo must be 0 or 1");
            c->opt_df_algo = value;
            return VXPT_OK;
        default:
            return fail(VXPT_E_INVALID, "unknown option");
    }
}

int vxpt_measure_l2_sector_peak(vxpt_handle c, double* gbps) {
    if (!c || !gbps) return fail(VXPT_E_INVALID, "NULL argument");
    VX_CUDA(cudaSetDevice(c->device));
    return run_l2_probe(c, gbps);
}

}  // extern "C"
