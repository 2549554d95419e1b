// kernels_on_host.cpp — TEST INFRASTRUCTURE: the per-pixel trace kernels' own source, compiled by g++ and executed on the CPU.
//
// The build container has no GPU, so a kernel written here is first seen by a GPU at the end of a round.  To shorten that
// loop this file includes voxelpathtracer_b200/csrc/trace.cu, trace_reflection.cu, df_consumers.cu, gbuffer.cu and denoise.cu UNCHANGED (kernels, device functions and
// their host launchers with the per-frame constants) and gives g++ what nvcc would: the CUDA vector types come from the toolkit's
// own headers (they are plain C++), the handful of device intrinsics the kernels use are defined below with their documented
// semantics, and VX_LAUNCH becomes a loop over blockIdx / threadIdx.  tests/test_kernels_on_host.py compares what comes out with
// the oracle and the reference-shader golden digests, so arithmetic or indexing mistakes in a kernel show up in the CPU suite.
//
// What this does NOT cover: kernels that use shared memory, warp shuffles or ballots (the DPX distance-field kernels, the
// wavefront GI pipeline) — those are only ever checked on the GPU (tests/test_gpu_parity.py).
// This library is never loaded by the product package; libvxpt.so is built by nvcc alone and has no host path.
#define VXPT_HOST_SHADOW 1
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

// ---- what nvcc provides implicitly ------------------------------------------------------------------------------------------
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__
static thread_local uint3 threadIdx, blockIdx;
using std::max;
using std::min;
// __fadd_rd(a, b): a + b rounded toward minus infinity.  The double sum of two floats is exact; round it down to float.
static inline float __fadd_rd(float a, float b) {
    const double s = (double)a + (double)b;
    float f = (float)s;  // round to nearest
    if ((double)f > s) f = std::nextafterf(f, -INFINITY);
    return f;
}
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
static inline unsigned __float2uint_rn(float f) { return (unsigned)std::nearbyintf(f); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

#define VX_CUDA(expr) do { } while (0)
#define VX_LAUNCH(kernel, grid, block, stream, ...)                                          \
    do {                                                                                     \
        const dim3 _g = (grid);                                                              \
        const int _nb = (int)(_g.x * _g.y);                                                  \
        _Pragma("omp parallel for schedule(dynamic, 4)") for (int _b = 0; _b < _nb; ++_b) {  \
            blockIdx = uint3{(unsigned)_b % _g.x, (unsigned)_b / _g.x, 0u};                  \
            for (unsigned _t = 0; _t < (unsigned)(block); ++_t) {                            \
                threadIdx = uint3{_t, 0u, 0u};                                               \
                kernel(__VA_ARGS__);                                                         \
            }                                                                                \
        }                                                                                    \
    } while (0)

#define VX_WARP_EMU_SET_DIMS(bs, g) do { } while (0)
#include "warp_emu.h"

#include "../../voxelpathtracer_b200/csrc/trace.cu"
#include "../../voxelpathtracer_b200/csrc/trace_reflection.cu"
#include "../../voxelpathtracer_b200/csrc/df_consumers.cu"
#include "../../voxelpathtracer_b200/csrc/gbuffer.cu"
#include "../../voxelpathtracer_b200/csrc/denoise.cu"

namespace vxpt {
// api.cu's make_scene, on a context whose "device" pointers are host pointers
SceneDev make_scene(const vxpt_ctx* c) {
    SceneDev S{};
    S.grid = c->d_grid; S.df = c->d_df; S.steps = c->d_steps; S.materials = c->d_materials;
    S.sobol = c->d_bluenoise; S.scramble = c->d_bluenoise ? c->d_bluenoise + 65536 : nullptr;
    S.rank = c->d_bluenoise ? c->d_bluenoise + 65536 + 131072 : nullptr;
    S.albedo_lod3 = c->d_albedo; S.pbr_lod2 = c->d_pbr; S.emissive = c->d_emissive;
    S.normal_lod3 = c->d_normal; S.emissive_lod2 = c->d_emissive2;
    S.sky = c->d_sky; S.shadow_noise = c->d_shadow_noise;
    S.n_layers = c->n_layers; S.n_emissive = c->n_emissive; S.sky_n = c->sky_n;
    S.counters = c->d_counters;
    S.alpha_mips = c->d_alpha_mips;
    S.n_alpha_layers = c->n_alpha_layers;
    S.albedo_mips = c->d_albedo_mips; S.normal_mips = c->d_normal_mips; S.pbr_mips = c->d_pbr_mips;
    S.srgb_lut = c->d_srgb_lut; S.n_mip_layers = c->n_mip_layers;
    S.lava_albedo = c->d_lava_albedo; S.lava_normal = c->d_lava_normal;
    S.lut = c->d_lut;
    return S;
}
// the wavefront GI pipeline needs a GPU (shared memory, ballots); the shadow runs the one-thread-per-pixel kernel
int launch_diffuse_wavefront(vxpt_ctx*, const VxCamera&, const DiffuseDev&, const VxGBuffer&, const VxDiffuseOut&) { return VXPT_E_UNSUPPORTED; }
void set_error(const std::string&) {}
}  // namespace vxpt

struct HostShadow {
    vxpt_ctx c;
    std::vector<uint8_t> steps, bluenoise;
    float srgb_lut[512];
    float trace_lut[vxpt::LUT_FLOATS];
    vxpt::DeviceCounters counters{};
};

extern "C" {
#define HS_API __attribute__((visibility("default")))

// mirrors oracle/vxo.py's VxoScene (same field order), so the tests hand both sides the same arrays
typedef struct HsScene {
    int32_t wx, wy, wz;
    const uint8_t* grid;
    const uint8_t* df;
    const int32_t* materials;
    const int32_t* sobol;
    const int32_t* scramble;
    const int32_t* rank;
    const float* albedo_lod3;
    const float* pbr_lod2;
    int32_t n_layers;
    const float* emissive_lod0;
    int32_t n_emissive_layers;
    const float* sky;
    int32_t sky_n;
    const uint8_t* shadow_noise;
    const float* normal_lod3;
    int32_t n_normal_layers;
    const float* emissive_lod2;
    const uint8_t* alpha_mips;
    int32_t n_alpha_layers;
    const uint8_t* albedo_mips;
    const uint8_t* normal_mips;
    const uint8_t* pbr_mips;
    int32_t n_mip_layers;
    const uint8_t* lava_albedo;
    const uint8_t* lava_normal;
} HsScene;

HS_API void* hs_create(const HsScene* s, int layout, int texel_format) {
    using namespace vxpt;
    HostShadow* h = new HostShadow();
    vxpt_ctx& c = h->c;
    c.d_grid = const_cast<uint8_t*>(s->grid);
    c.d_df = const_cast<uint8_t*>(s->df);
    c.opt_layout = layout; c.opt_wavefront = 0; c.opt_texel = texel_format;
    // pack_steps (df_build.cu): E(M) = (M == 1) ? 1 : floor(M * 0.57735026918f), linear or 8x4x4 bricks
    h->steps.assign(layout == 1 ? STEPS_TILED_BYTES : VOXELS, 0);
    if (s->df)
        for (int z = 0; z < WZ; ++z)
            for (int y = 0; y < WY; ++y)
                for (int x = 0; x < WX; ++x) {
                    const size_t lin = (size_t)x + (size_t)WX * ((size_t)y + (size_t)WY * z);
                    const float m = (float)s->df[lin];
                    const uint8_t e = (uint8_t)(int)std::floor(m == 1.0f ? 1.0f : m * 0.57735026918f);
                    h->steps[layout == 1 ? brick_offset(x, y, z) : lin] = e;
                }
    c.d_steps = h->steps.data();
    c.d_materials = const_cast<int32_t*>(s->materials);
    if (s->sobol) {
        h->bluenoise.resize(65536 + 131072 + 131072);
        for (int k = 0; k < 65536; ++k) h->bluenoise[k] = (uint8_t)s->sobol[k];
        for (int k = 0; k < 131072; ++k) h->bluenoise[65536 + k] = (uint8_t)s->scramble[k];
        for (int k = 0; k < 131072; ++k) h->bluenoise[65536 + 131072 + k] = (uint8_t)s->rank[k];
        c.d_bluenoise = h->bluenoise.data();
    }
    c.d_albedo = (float4*)s->albedo_lod3; c.d_pbr = (float4*)s->pbr_lod2; c.d_emissive = (float*)s->emissive_lod0;
    c.d_normal = (float4*)s->normal_lod3; c.d_emissive2 = (float*)s->emissive_lod2;
    c.n_layers = s->n_layers; c.n_emissive = s->n_emissive_layers; c.sky_n = s->sky_n;
    c.d_sky = (float*)s->sky;
    if (s->sky) c.h_sky.assign(s->sky, s->sky + (size_t)6 * s->sky_n * s->sky_n * 3);
    c.d_shadow_noise = (uchar4*)s->shadow_noise;
    c.d_alpha_mips = (uint8_t*)s->alpha_mips;
    c.n_alpha_layers = s->n_alpha_layers;
    c.d_counters = &h->counters;
    // vxpt_set_gbuffer_textures (api.cu): the mip chains and the sRGB decode table
    c.d_albedo_mips = (uchar4*)s->albedo_mips; c.d_normal_mips = (uchar4*)s->normal_mips; c.d_pbr_mips = (uchar4*)s->pbr_mips;
    c.n_mip_layers = s->n_mip_layers;
    c.d_lava_albedo = (uchar4*)s->lava_albedo; c.d_lava_normal = (uchar4*)s->lava_normal;
    for (int k = 0; k < 256; ++k) {
        const double cs = (double)k / 255.0;
        h->srgb_lut[k] = (float)(cs <= 0.04045 ? cs / 12.92 : std::pow((cs + 0.055) / 1.055, 2.4));
        h->srgb_lut[256 + k] = (float)k / 255.0f;
    }
    c.d_srgb_lut = h->srgb_lut;
    vxpt::fill_trace_lut(h->trace_lut);  // what vxpt_create uploads (api.cu)
    c.d_lut = h->trace_lut;
    return h;
}
HS_API void hs_destroy(void* p) { delete (HostShadow*)p; }
HS_API vxpt_ctx* hs_ctx(void* p) { return &((HostShadow*)p)->c; }
HS_API void hs_stats(void* p, uint64_t out[3], int reset) {
    HostShadow* h = (HostShadow*)p;
    out[0] = h->counters.rays; out[1] = h->counters.df_fetches; out[2] = h->counters.vox_fetches;
    if (reset) h->counters = vxpt::DeviceCounters{};
}
HS_API int hs_trace_primary(void* p, const VxCamera* cam, const VxPrimaryParams* prm, const VxGBuffer* out) {
    return vxpt::launch_primary(hs_ctx(p), *cam, *prm, *out);
}
HS_API int hs_trace_shadow(void* p, const VxCamera* cam, const VxGBuffer* g, const VxShadowParams* prm, const VxShadowOut* out) {
    return vxpt::launch_shadow(hs_ctx(p), *cam, *g, *prm, *out);
}
HS_API int hs_trace_diffuse(void* p, const VxCamera* cam, const VxGBuffer* g, const VxDiffuseParams* prm, const VxDiffuseOut* out) {
    return vxpt::launch_diffuse(hs_ctx(p), *cam, *g, *prm, *out);
}
HS_API int hs_trace_reflection(void* p, const VxCamera* cam, const VxGBuffer* g, const VxReflectionIn* in, const VxReflectionParams* prm,
                               const VxReflectionOut* out) {
    return vxpt::launch_reflection(hs_ctx(p), *cam, *g, *in, *prm, *out);
}
HS_API int hs_generate_gbuffer(void* p, const VxCamera* cam, const VxGBuffer* g, const VxMaterialParams* prm, const VxMaterialOut* out) {
    return vxpt::launch_gbuffer(hs_ctx(p), *cam, *g, *prm, *out);
}
// the SVGF denoiser passes need no scene: hs_create(NULL-scene) handles work too
HS_API int hs_svgf_initial(void* p, const VxCamera* cam, const VxSvgfInitialIn* in, const VxSvgfInitialOut* out) {
    return vxpt::launch_svgf_initial(hs_ctx(p), *cam, *in, *out);
}
HS_API int hs_svgf_temporal(void* p, const VxCamera* cam, const VxSvgfTemporalIn* in, const VxSvgfTemporalParams* prm, const VxSvgfTemporalOut* out) {
    return vxpt::launch_svgf_temporal(hs_ctx(p), *cam, *in, *prm, *out);
}
HS_API int hs_svgf_variance(void* p, const VxCamera* cam, const VxSvgfVarianceIn* in, const VxSvgfVarianceParams* prm, const VxSvgfVarianceOut* out) {
    return vxpt::launch_svgf_variance(hs_ctx(p), *cam, *in, *prm, *out);
}
HS_API int hs_svgf_spatial(void* p, const VxCamera* cam, const VxSvgfSpatialIn* in, const VxSvgfSpatialParams* prm, const VxSvgfSpatialOut* out) {
    return vxpt::launch_svgf_spatial(hs_ctx(p), *cam, *in, *prm, *out);
}
HS_API int hs_shadow_temporal(void* p, const VxCamera* cam, const VxShadowTemporalIn* in, const VxShadowTemporalParams* prm, const VxShadowTemporalOut* out) {
    return vxpt::launch_shadow_temporal(hs_ctx(p), *cam, *in, *prm, *out);
}
HS_API int hs_shadow_filter(void* p, const VxCamera* cam, const VxShadowFilterIn* in, const VxShadowFilterParams* prm, float* out) {
    return vxpt::launch_shadow_filter(hs_ctx(p), *cam, *in, *prm, out);
}
HS_API int hs_trace_rays(void* p, const float* origins, const float* directions, int n, int max_it, float* t, uint8_t* normal_id, uint8_t* block_id,
                         int16_t* hit_voxel) {
    return vxpt::launch_rays(hs_ctx(p), origins, directions, n, max_it, t, normal_id, block_id, hit_voxel);
}
HS_API int hs_ambient_sound(void* p, const float* player_pos, int frame, uint32_t* aggregate, uint32_t* per_invocation) {
    *aggregate = 0;
    return vxpt::launch_ambient(hs_ctx(p), player_pos, frame, aggregate, per_invocation);
}
// the denoiser's pinned transcendentals, element by element (tests/test_pinned_math.py compares them with (float)exp((double)x) / pow)
HS_API void hs_exp_cr(const float* x, float* y, long n) {
    for (long i = 0; i < n; ++i) y[i] = vxpt::exp_cr(x[i]);
}
HS_API void hs_pow01_cr(const float* x, const float* e, float* y, long n) {
    for (long i = 0; i < n; ++i) y[i] = vxpt::pow01_cr(x[i], e[i]);
}
// how many of the inputs the short evaluation hands to the library function (its rounding test)
HS_API long hs_exp_short_refused(const float* x, long n) {
    long c = 0;
    float y;
    for (long i = 0; i < n; ++i) c += !vxpt::exp_short((double)x[i], 16, &y);
    return c;
}
HS_API float hs_normal_weight(int a, int b, float at_floor, int power) {
    return vxpt::normal_weight(a, b, at_floor, power == 16 ? VXPT_POW3_16 : VXPT_POW3_32);
}
}  // extern "C"
