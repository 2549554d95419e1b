// vxpt_internal.h — private context of the vxpt C ABI (not installed; include/vxpt.h is the public surface).
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string>
#include <vector>

#include "../../include/vxpt.h"

namespace vxpt {

constexpr int WX = VXPT_WORLD_SIZE_X;  // 384
constexpr int WY = VXPT_WORLD_SIZE_Y;  // 128
constexpr int WZ = VXPT_WORLD_SIZE_Z;  // 384
constexpr size_t VOXELS = (size_t)WX * WY * WZ;
constexpr int SLICE_BYTES = WX * WY;  // one z-slice: 49,152 B

// ---- tiled step field ("bricks") --------------------------------------------------------------------------------
// Traversal never needs the Manhattan value M itself, only E(M) = (M==1) ? 1 : floor(M * 0.57735026918f)
// (ToConservativeEuclidean, InitialRayTraceFrag.glsl:90-93,329).  The step field stores E (<= 146) per voxel.
// Layout 1 tiles it into 8x4x4-voxel bricks of 128 B (one L1 line) whose low 5 address bits cover 4x x 4z x 2y voxels
// (one 32-B sector = a 3-D neighbourhood instead of a 32-voxel x-run).  The brick grid is padded to powers of two
// (64 x 32 x 96 bricks) so the address is three independent bit-deposits, one LOP3 + IMAD each:
//   bits 0-1 x&3 | 2-3 z&3 | 4-5 y&3 | 6 (x>>2)&1 | 7-12 x>>3 | 13-17 y>>2 | 18-24 z>>2
constexpr size_t STEPS_TILED_BYTES = (size_t)96 << 18;  // 25,165,824 B (x padded 48 -> 64 bricks)

__host__ __device__ __forceinline__ uint32_t brick_offset(int x, int y, int z) {
    const uint32_t ux = (uint32_t)x, uy = (uint32_t)y, uz = (uint32_t)z;
    return (ux + (ux & ~3u) * 15u) + (uy * 16u + (uy & ~3u) * 2032u) + (uz * 4u + (uz & ~3u) * 65532u);
}

struct DeviceCounters {
    unsigned long long rays, df_fetches, vox_fetches;
};

// device-side view of everything the trace kernels read
struct SceneDev {
    const uint8_t* grid;       // block ids, linear x + 384*(y + 128*z)
    const uint8_t* df;         // Manhattan distance field, linear
    const uint8_t* steps;      // step field E(M), bricks (layout 1) or linear (layout 0)
    const int32_t* materials;  // 768
    const uint8_t* sobol;      // 65536   (blue-noise tables hold values < 256: kept as bytes, 320 KB total)
    const uint8_t* scramble;   // 131072
    const uint8_t* rank;       // 131072
    const float4* albedo_lod3; // [n_layers][64][64]
    const float4* pbr_lod2;    // [n_layers][128][128]
    const float* emissive;     // [n_emissive][512][512]
    const float4* normal_lod3; // [n_normal][64][64]   (reflection pass)
    const float* emissive_lod2;  // [n_emissive][128][128] (reflection pass)
    const float* sky;          // [6][n][n][3]
    const uchar4* shadow_noise;  // [256][256]
    int n_layers, n_emissive, sky_n;
    DeviceCounters* counters;
    const uint8_t* alpha_mips;  // [n_alpha_layers][VXPT_ALPHA_MIP_TEXELS] albedo alpha, mip levels 0..8 (alpha-tested traversal)
    int n_alpha_layers;
    // G-buffer material pass (gbuffer.cu): RGBA8 mip chains [n_mip_layers][VXPT_MIP_CHAIN_TEXELS] and the sRGB decode table
    const uchar4* albedo_mips;
    const uchar4* normal_mips;
    const uchar4* pbr_mips;
    const float* srgb_lut;  // 256 sRGB-decoded values, then 256 plain unorm8 values c / 255
    int n_mip_layers;
    const uchar4* lava_albedo;  // [VXPT_LAVA_FRAMES][VXPT_LAVA_SIZE][VXPT_LAVA_SIZE] (vxpt_set_lava_textures)
    const uchar4* lava_normal;
    const float* lut;  // LUT_FLOATS floats of per-handle constants (fill_trace_lut below)
};

// ---- exact tables for the trace passes ---------------------------------------------------------------------------
// The hemisphere / cone samplers take sin and cos of an angle that is a function of ONE byte: a blue-noise value v in 0..255
// (cosWeightedRandomHemisphereDirection, DiffuseRayTraceFrag.glsl:945-967: 2*PI*r1 with r1 = (0.5 + v) / 256) or a noise texel
// channel (SampleCone, ShadowRayTraceFrag.glsl:388-398,474: phi = (v / 255) * PI * 2).  The parity contract pins sin / cos as the
// correctly rounded fp32 value (double evaluation, one rounding), so 256 (cos, sin) pairs per sampler, computed once on the host
// with exactly the fp32 argument arithmetic of the shader, replace two double-precision evaluations per ray.
// Likewise the tangent frame of the hemisphere sampler, uu = normalize(cross(n, (0,1,1))), vv = cross(uu, n), is a function of the
// face (6 axis normals): 6 x 6 floats, evaluated on the host with the device code's fp32 operation order.
constexpr int LUT_TRIG_GI = 0;         // [256] (cos, sin) of 2*PI * ((0.5 + v) / 256)
constexpr int LUT_TRIG_CONE = 512;     // [256] (cos, sin) of ((v / 255) * PI) * 2
constexpr int LUT_BASIS = 1024;        // [6] (uu.xyz, vv.xyz) per normal id 0..5 (GetNormalFromID order)
constexpr int LUT_TRIG_GGX = 1024 + 36;  // [256] (cos, sin) of 2*PI * (((0.5 + v) / 256) * 0.9): ImportanceSampleGGX's phi (ReflectionTraceFrag.glsl:345-365,
                                         // Xi.x = blue-noise sample * 0.9, :629-633)
constexpr int LUT_FLOATS = 1024 + 36 + 512;
inline void fill_trace_lut(float* lut) {  // host; fp32 expressions as in the device code (host objects are built with -ffp-contract=off)
    for (int v = 0; v < 256; ++v) {
        const float r1 = (0.5f + (float)v) / 256.0f;               // blue_noise_1d
        const float a = (2.0f * 3.14159265359f) * r1;               // PI2 * r1
        lut[LUT_TRIG_GI + 2 * v] = (float)cos((double)a);
        lut[LUT_TRIG_GI + 2 * v + 1] = (float)sin((double)a);
        const float phi = (((float)v / 255.0f) * 3.14159265359f) * 2.0f;  // xi_y * PI * 2
        lut[LUT_TRIG_CONE + 2 * v] = (float)cos((double)phi);
        lut[LUT_TRIG_CONE + 2 * v + 1] = (float)sin((double)phi);
        const float xi_x = r1 * 0.9f;                               // importance_sample_ggx(..., xx * 0.9f, ...)
        const float pg = (2.0f * 3.14159265359f) * xi_x;            // 2.0f * PI * Xi.x
        lut[LUT_TRIG_GGX + 2 * v] = (float)cos((double)pg);
        lut[LUT_TRIG_GGX + 2 * v + 1] = (float)sin((double)pg);
    }
    const float N[6][3] = {{0, 0, 1}, {0, 0, -1}, {0, 1, 0}, {0, -1, 0}, {-1, 0, 0}, {1, 0, 0}};
    for (int k = 0; k < 6; ++k) {
        const float nx = N[k][0], ny = N[k][1], nz = N[k][2], bx = 0.0f, by = 1.0f, bz = 1.0f;
        const float cx = ny * bz - by * nz, cy = nz * bx - bz * nx, cz = nx * by - bx * ny;  // cross3(n, (0,1,1))
        const float inv = 1.0f / sqrtf((cx * cx + cy * cy) + cz * cz);
        const float ux = cx * inv, uy = cy * inv, uz = cz * inv;                              // normalize3
        float* o = lut + LUT_BASIS + 6 * k;
        o[0] = ux; o[1] = uy; o[2] = uz;
        o[3] = uy * nz - ny * uz; o[4] = uz * nx - nz * ux; o[5] = ux * ny - nx * uy;         // cross3(uu, n)
    }
}

}  // namespace vxpt

struct vxpt_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;                  // around the last trace pass
    cudaEvent_t ev2 = nullptr, ev3 = nullptr, ev4 = nullptr;  // DF build start / end, brick pack end
    bool pass_timed = false, df_timed = false, pack_timed = false;

    uint8_t* d_grid = nullptr;
    uint8_t* d_df = nullptr;
    uint8_t* d_tmp = nullptr;    // XY-pass intermediate
    uint8_t* d_steps = nullptr;  // brick-swizzled E field
    bool world_uploaded = false;
    bool df_valid = false;

    int32_t* d_materials = nullptr;
    int32_t h_materials[768] = {0};  // host copy for argument validation
    uint8_t* d_bluenoise = nullptr;  // sobol | scramble | rank
    float4* d_albedo = nullptr;
    float4* d_pbr = nullptr;
    float* d_emissive = nullptr;
    float4* d_normal = nullptr;
    float* d_emissive2 = nullptr;
    int n_normal = 0, n_emissive2 = 0;
    bool have_refl_textures = false;
    std::vector<float> h_sky;  // host copy of the cubemap (per-frame sun / moon colours of the reflection pass)
    float* d_sky = nullptr;
    uchar4* d_shadow_noise = nullptr;
    uint8_t* d_alpha_mips = nullptr;
    int n_alpha_layers = 0;
    uchar4* d_albedo_mips = nullptr;  // vxpt_set_gbuffer_textures
    uchar4* d_normal_mips = nullptr;
    uchar4* d_pbr_mips = nullptr;
    float* d_srgb_lut = nullptr;
    int n_mip_layers = 0;
    uchar4* d_lava_albedo = nullptr;  // vxpt_set_lava_textures
    uchar4* d_lava_normal = nullptr;
    float* d_lut = nullptr;           // vxpt::fill_trace_lut
    int n_layers = 0, n_emissive = 0, sky_n = 0;
    bool have_materials = false, have_bluenoise = false, have_textures = false, have_sky = false, have_shadow_noise = false;

    vxpt::DeviceCounters* d_counters = nullptr;
    float last_ms = 0.f, df_build_ms = 0.f, brick_pack_ms = 0.f;
    uint64_t launches = 0;

    // options
    int opt_layout = 1;     // VXPT_OPT_TRAVERSAL_LAYOUT
    int steps_layout = -1;  // layout the step field currently holds (-1: stale, launch_pack_bricks must run)
    int opt_wavefront = 1;  // VXPT_OPT_GI_WAVEFRONT
    int opt_refl_wavefront = 1;  // VXPT_OPT_REFLECTION_WAVEFRONT
    int opt_df_algo = 1;    // 0 = one thread per line (reference-shaped) + pack_steps, 1 = DPX sweeps, step field written by the z sweep
    int opt_replicas = 1;   // VXPT_OPT_SCENE_REPLICAS
    int opt_timing = 1;     // VXPT_OPT_TIMING_EVENTS
    int opt_texel = 0;      // VXPT_OPT_TEXEL_FORMAT
    int opt_quad_shuffle = 1;  // VXPT_OPT_MATERIAL_QUAD_SHUFFLE
    uint8_t* rep_grid[8] = {nullptr};   // extra copies (index 1..replicas-1); index 0 unused (= d_grid / d_steps)
    uint8_t* rep_steps[8] = {nullptr};
    uint64_t frame_counter = 0;  // advanced by vxpt_trace_primary

    // device staging for host-pointer I/O (grown on demand)
    void* d_stage = nullptr;  // (void*: grow_scratch)
    size_t stage_bytes = 0;

    // vxpt_render_frame: copy-out stream + one event per row slab
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_slab[8] = {nullptr};
    bool frame_pending = false;  // vxpt_render_frame_async: copies to host planes still in flight

    // peer-to-peer gather (vxpt_shared_*, vxpt_signal, vxpt_wait_all)
    struct SharedBuf {
        void* ptr;
        bool owned;
    };
    std::vector<SharedBuf> shared;
    unsigned* d_wait_err = nullptr;  // latched by a wait that timed out
    bool waits_issued = false;

    // vxpt_svgf_frame: the denoiser's device-resident planes and history
    struct SvgfHistory {
        int width = 0, height = 0;
        bool valid = false;     // a previous frame is held
        int cur = 0;            // which temporal set the last frame wrote
        float prev_view[16] = {0}, prev_projection[16] = {0};
        void* buf = nullptr;    // one allocation carved into the planes below
        float *prev_t = nullptr, *temporal[2][4] = {{nullptr}}, *pre[4] = {nullptr}, *var[3] = {nullptr}, *pong[2][4] = {{nullptr}};
        uint8_t *prev_nid = nullptr, *prev_bid = nullptr;
    } svgf;

    // vxpt_shadow_filter_frame: temporal shadow planes and the previous frame's hit distances
    struct ShadowHistory {
        int width = 0, height = 0;
        bool valid = false;
        int cur = 0;
        float prev_view[16] = {0}, prev_projection[16] = {0};
        void* buf = nullptr;
        float *prev_t = nullptr, *shadow[2] = {nullptr, nullptr}, *frames[2] = {nullptr, nullptr};
    } shadow_hist;

    // wavefront queues (grown on demand)
    void* d_queue = nullptr;
    size_t queue_bytes = 0;
    // GI pass: second (higher-priority) stream for the latency-bound continuation kernels + fork / join events (trace_gi.cu)
    cudaStream_t gi_stream = nullptr;
    cudaEvent_t ev_gi[9] = {nullptr};
};

namespace vxpt {

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);
bool stream_capturing(const vxpt_ctx* c);
int grow_scratch(vxpt_ctx* c, void** buf, size_t* have, size_t need, const char* what);

#ifndef VXPT_HOST_SHADOW
#define VX_CUDA(expr)                                                         \
    do {                                                                      \
        cudaError_t _e = (expr);                                              \
        if (_e != cudaSuccess) return ::vxpt::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)
// every kernel launch of the per-pixel passes goes through this macro.  (tests/host_shadow compiles these translation units
// with g++ and defines it as a loop over blockIdx / threadIdx, so the CPU suite executes the kernels' own source against the
// oracle; the product library is only ever built by nvcc and has no host path.)
#define VX_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
// a kernel that exchanges values between the lanes of a warp (the host builds of the tests run each warp twice: record, then replay)
#define VX_LAUNCH_WARPSYNC(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#endif

// df_build.cu
int init_df_kernels(vxpt_ctx* c);
int launch_df_build(vxpt_ctx* c);
int launch_pack_bricks(vxpt_ctx* c);
// trace.cu
int launch_primary(vxpt_ctx* c, const VxCamera& cam, const VxPrimaryParams& p, const VxGBuffer& out_dev);
int launch_shadow(vxpt_ctx* c, const VxCamera& cam, const VxGBuffer& g_dev, const VxShadowParams& p, const VxShadowOut& out_dev);
int launch_diffuse(vxpt_ctx* c, const VxCamera& cam, const VxGBuffer& g_dev, const VxDiffuseParams& p, const VxDiffuseOut& out_dev);
// trace_gi.cu: bytes of wavefront scratch a slab needs (hit queue + per-pixel sample state when a pixel takes several samples)
size_t gi_scratch_bytes(size_t slab_px, size_t frame_px, bool spp1);
// trace_reflection.cu
int launch_reflection(vxpt_ctx* c, const VxCamera& cam, const VxGBuffer& g_dev, const VxReflectionIn& in_dev, const VxReflectionParams& p,
                      const VxReflectionOut& out_dev);
// df_consumers.cu
int launch_rays(vxpt_ctx* c, const float* origins, const float* directions, int n, int max_iterations, float* t, uint8_t* normal_id,
                uint8_t* block_id, int16_t* hit_voxel);
int launch_ambient(vxpt_ctx* c, const float player[3], int frame, unsigned* aggregate, unsigned* per_invocation);
// gbuffer.cu
int launch_gbuffer(vxpt_ctx* c, const VxCamera& cam, const VxGBuffer& g_dev, const VxMaterialParams& p, const VxMaterialOut& out_dev);
// denoise.cu
int launch_svgf_initial(vxpt_ctx* c, const VxCamera& cam, const VxSvgfInitialIn& in_dev, const VxSvgfInitialOut& out_dev);
int launch_svgf_temporal(vxpt_ctx* c, const VxCamera& cam, const VxSvgfTemporalIn& in_dev, const VxSvgfTemporalParams& p, const VxSvgfTemporalOut& out_dev);
int launch_svgf_variance(vxpt_ctx* c, const VxCamera& cam, const VxSvgfVarianceIn& in_dev, const VxSvgfVarianceParams& p, const VxSvgfVarianceOut& out_dev);
int launch_svgf_spatial(vxpt_ctx* c, const VxCamera& cam, const VxSvgfSpatialIn& in_dev, const VxSvgfSpatialParams& p, const VxSvgfSpatialOut& out_dev);
int launch_shadow_temporal(vxpt_ctx* c, const VxCamera& cam, const VxShadowTemporalIn& in_dev, const VxShadowTemporalParams& p, const VxShadowTemporalOut& out_dev);
int launch_shadow_filter(vxpt_ctx* c, const VxCamera& cam, const VxShadowFilterIn& in_dev, const VxShadowFilterParams& p, float* out_dev);
// l2_probe.cu
int run_l2_probe(vxpt_ctx* c, double* gbps);

SceneDev make_scene(const vxpt_ctx* c);

}  // namespace vxpt
