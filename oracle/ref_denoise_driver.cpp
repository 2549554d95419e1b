// ref_denoise_driver.cpp — the reference's SVGF denoiser shaders (Core/Shaders/SVGF/TemporalFilter.glsl, VarianceEstimate.glsl,
// SpatialFilter.glsl) and sun-shadow filters (Core/Shaders/ShadowTemporalFilter.glsl, ShadowFilter.glsl; Pipeline.cpp:2854-2944) compiled as C++; uniforms and binds follow Core/Pipeline.cpp:2335-2596, the vertex stage Core/Shaders/FBOVert.glsl.
// The attachments are bound with the filters Core/Pipeline.cpp:1094-1152 declares (hit distance LINEAR, ids NEAREST, the rest LINEAR) and
// GL_REPEAT (Core/GLClasses/Framebuffer.cpp:66-67).  See ref_shader_driver.cpp.  Test infrastructure only.
#include <cstdint>
#include <cstring>
#include <vector>

#include "glsl_compat.h"

namespace glsl {
// min / max / clamp of these shaders: pinned to IEEE minNum / maxNum (the non-NaN operand wins), which is what the GPUs the reference
// runs on do; glm's (a < b) ? b : a would propagate NaN instead.  The first frames of the accumulation divide by a zero frame count
// (VarianceEstimate.glsl:171), so infinities and 0 * inf do reach these calls.  Declared in the namespace that encloses the shader text,
// so unqualified calls resolve here (vector arguments still see glm's templates through ADL; the exact non-template overloads win).
namespace denoise {
inline float max(float a, float b) { return std::fmax(a, b); }
inline float max(int a, float b) { return std::fmax((float)a, b); }
inline float max(float a, int b) { return std::fmax(a, (float)b); }
inline float min(float a, float b) { return std::fmin(a, b); }
inline float clamp(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
inline vec2 clamp(const vec2& v, float lo, float hi) { return vec2(clamp(v.x, lo, hi), clamp(v.y, lo, hi)); }
inline vec3 clamp(const vec3& v, float lo, float hi) { return vec3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi)); }
inline vec4 clamp(const vec4& v, float lo, float hi) { return vec4(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi), clamp(v.w, lo, hi)); }
inline vec3 min(const vec3& a, const vec3& b) { return vec3(std::fmin(a.x, b.x), std::fmin(a.y, b.y), std::fmin(a.z, b.z)); }
inline vec3 max(const vec3& a, const vec3& b) { return vec3(std::fmax(a.x, b.x), std::fmax(a.y, b.y), std::fmax(a.z, b.z)); }
inline vec2 operator/(float a, const ivec2& b) { return vec2(a) / vec2(b); }
#include "_ref/ShadowTemporalFilter.inc"
#include "_ref/ShadowFilter.inc"
#include "_ref/Spatial3x3Initial.inc"
#include "_ref/TemporalFilter.inc"
#include "_ref/VarianceEstimate.inc"
#include "_ref/SpatialFilter.inc"
}  // namespace denoise
}  // namespace glsl

namespace {
struct IdPlanes {  // R8 attachments as the shaders read them: normal id / 10 (1.0 on a miss), block id / 255
    std::vector<float> normal, block;
    IdPlanes(const uint8_t* nid, const uint8_t* bid, size_t n) : normal(n), block(n) {
        for (size_t k = 0; k < n; ++k) {
            normal[k] = nid[k] > 5 ? 1.0f : (float)nid[k] / 10.0f;
            block[k] = bid ? (float)bid[k] / 255.0f : 0.0f;
        }
    }
};
inline glsl::sampler2D bind(const float* d, int w, int h, int comps, bool linear) {
    glsl::sampler2D s;
    s.data = d; s.w = w; s.h = h; s.comps = comps; s.linear = linear; s.repeat = true;
    return s;
}
}  // namespace

struct RefSvgfArgs {  // plain C layout, filled by oracle/ref_shaders.py; each entry point reads the members it needs
    const float* inv_view;
    const float* inv_proj;
    int32_t width, height, row_begin, row_end;
    const float* g_t;           const uint8_t* g_normal_id;      const uint8_t* g_block_id;
    const float* prev_t;        const uint8_t* prev_normal_id;   const uint8_t* prev_block_id;
    const float* sh;            const float* cocg;               const float* luma;            const float* ao_sky;
    const float* prev_sh;       const float* prev_cocg;          const float* prev_utility;    const float* prev_ao_sky;
    const float* utility;       const float* variance;           const float* temporal_utility;
    const float* prev_view;     const float* prev_projection;
    int32_t be_useful, do_spatial, aggressive, step, large_kernel;
    float color_phi_bias, time, resolution_scale;
    float* o_sh;  float* o_cocg;  float* o_utility;  float* o_variance;  float* o_ao_sky;
};

#define FOR_EACH_PIXEL(S, body)                                                                              \
    for (int j = a->row_begin; j < a->row_end; ++j)                                                         \
        for (int i = 0; i < W; ++i) {                                                                        \
            const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;                  \
            S::v_TexCoords = vec2(u, v);                                                                     \
            gl_FragCoord = vec4((float)i + 0.5f, (float)j + 0.5f, 0.0f, 1.0f);                               \
            S::shader_reset_globals();                                                                       \
            S::shader_main();                                                                                \
            const size_t px = (size_t)j * W + i;                                                             \
            body                                                                                             \
        }

extern "C" __attribute__((visibility("default"))) int ref_svgf_temporal(const RefSvgfArgs* a) {
    using namespace glsl;
    namespace S = glsl::denoise::ns_TemporalFilter;
    const int W = a->width, H = a->height;
    const size_t n = (size_t)W * H;
    IdPlanes cur(a->g_normal_id, a->g_block_id, n), prev(a->prev_normal_id, a->prev_block_id, n);
    std::memcpy(&S::u_InverseView[0][0], a->inv_view, 64);
    std::memcpy(&S::u_InverseProjection[0][0], a->inv_proj, 64);
    std::memcpy(&S::u_PrevView[0][0], a->prev_view, 64);
    std::memcpy(&S::u_PrevProjection[0][0], a->prev_projection, 64);
    S::u_BeUseful = a->be_useful != 0;
    S::u_Time = 0.0f;
    S::u_DeltaTime = 0.0f;
    S::v_RayOrigin = vec3(S::u_InverseView[3]);  // FBOVert.glsl:24
    S::u_CurrentPositionTexture = bind(a->g_t, W, H, 1, true);
    S::u_PreviousPositionTexture = bind(a->prev_t, W, H, 1, true);
    S::u_CurrentNormalTexture = bind(cur.normal.data(), W, H, 1, false);
    S::u_PreviousNormalTexture = bind(prev.normal.data(), W, H, 1, false);
    S::u_CurrentBlockIDTexture = bind(cur.block.data(), W, H, 1, false);
    S::u_PrevBlockIDTexture = bind(prev.block.data(), W, H, 1, false);
    S::u_CurrentSH = bind(a->sh, W, H, 4, true);
    S::u_CurrentCoCg = bind(a->cocg, W, H, 2, true);
    S::u_NoisyLuminosity = bind(a->luma, W, H, 1, true);
    S::u_CurrentAO = bind(a->ao_sky, W, H, 2, true);
    S::u_PreviousSH = bind(a->prev_sh, W, H, 4, true);
    S::u_PrevCoCg = bind(a->prev_cocg, W, H, 2, true);
    S::u_PreviousUtility = bind(a->prev_utility, W, H, 3, true);
    S::u_PreviousAO = bind(a->prev_ao_sky, W, H, 2, true);
    FOR_EACH_PIXEL(S, {
        std::memcpy(a->o_sh + 4 * px, &S::o_SH[0], 16);
        std::memcpy(a->o_cocg + 2 * px, &S::o_CoCg[0], 8);
        std::memcpy(a->o_utility + 3 * px, &S::o_Utility[0], 12);
        std::memcpy(a->o_ao_sky + 2 * px, &S::o_AOAndSkyLighting[0], 8);
    })
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ref_svgf_initial(const RefSvgfArgs* a) {
    using namespace glsl;
    namespace S = glsl::denoise::ns_Spatial3x3Initial;
    const int W = a->width, H = a->height;
    IdPlanes cur(a->g_normal_id, nullptr, (size_t)W * H);
    std::memcpy(&S::u_InverseView[0][0], a->inv_view, 64);
    std::memcpy(&S::u_InverseProjection[0][0], a->inv_proj, 64);
    S::u_Dimensions = vec2((float)W, (float)H);
    S::u_Time = a->time;
    S::u_DeltaTime = 0.0f;
    S::v_RayOrigin = vec3(S::u_InverseView[3]);
    S::u_SH = bind(a->sh, W, H, 4, true);
    S::u_CoCg = bind(a->cocg, W, H, 2, true);
    S::u_Utility = bind(a->luma, W, H, 1, true);
    S::u_AO = bind(a->ao_sky, W, H, 2, true);
    S::u_PositionTexture = bind(a->g_t, W, H, 1, true);
    S::u_NormalTexture = bind(cur.normal.data(), W, H, 1, false);
    FOR_EACH_PIXEL(S, {
        std::memcpy(a->o_sh + 4 * px, &S::o_SH[0], 16);
        std::memcpy(a->o_cocg + 2 * px, &S::o_CoCg[0], 8);
        a->o_utility[px] = S::o_Utility;
        std::memcpy(a->o_ao_sky + 2 * px, &S::o_AOSky[0], 8);
    })
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ref_svgf_variance(const RefSvgfArgs* a) {
    using namespace glsl;
    namespace S = glsl::denoise::ns_VarianceEstimate;
    const int W = a->width, H = a->height;
    IdPlanes cur(a->g_normal_id, nullptr, (size_t)W * H);
    std::memcpy(&S::u_InverseView[0][0], a->inv_view, 64);
    std::memcpy(&S::u_InverseProjection[0][0], a->inv_proj, 64);
    S::DO_SPATIAL = a->do_spatial != 0;
    S::AGGRESSIVE_DISOCCLUSION_HANDLING = a->aggressive != 0;
    S::v_RayOrigin = vec3(S::u_InverseView[3]);
    S::u_PositionTexture = bind(a->g_t, W, H, 1, true);
    S::u_NormalTexture = bind(cur.normal.data(), W, H, 1, false);
    S::u_SH = bind(a->sh, W, H, 4, true);
    S::u_CoCg = bind(a->cocg, W, H, 2, true);
    S::u_Utility = bind(a->utility, W, H, 3, true);
    FOR_EACH_PIXEL(S, {
        std::memcpy(a->o_sh + 4 * px, &S::o_SH[0], 16);
        std::memcpy(a->o_cocg + 2 * px, &S::o_CoCg[0], 8);
        a->o_variance[px] = S::o_Variance;
    })
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ref_svgf_spatial(const RefSvgfArgs* a) {
    using namespace glsl;
    namespace S = glsl::denoise::ns_SpatialFilter;
    const int W = a->width, H = a->height;
    IdPlanes cur(a->g_normal_id, a->g_block_id, (size_t)W * H);
    std::memcpy(&S::u_InverseView[0][0], a->inv_view, 64);
    std::memcpy(&S::u_InverseProjection[0][0], a->inv_proj, 64);
    S::u_Dimensions = vec2((float)W, (float)H);
    S::u_Step = a->step;
    S::u_ShouldDetailWeight = true;
    S::DO_SPATIAL = a->do_spatial != 0;
    S::u_LargeKernel = a->large_kernel != 0;
    S::AGGRESSIVE_DISOCCLUSION_HANDLING = a->aggressive != 0;
    S::u_ColorPhiBias = a->color_phi_bias;
    S::u_DeltaTime = 0.0f;
    S::u_Time = a->time;
    S::u_ResolutionScale = a->resolution_scale;
    S::v_RayOrigin = vec3(S::u_InverseView[3]);
    S::u_SH = bind(a->sh, W, H, 4, true);
    S::u_CoCg = bind(a->cocg, W, H, 2, true);
    S::u_Utility = bind(a->ao_sky, W, H, 2, true);          // unit 6 = DiffuseTemporalFBO attachment 3 (Pipeline.cpp:2581-2582); never used
    S::u_PositionTexture = bind(a->g_t, W, H, 1, true);
    S::u_NormalTexture = bind(cur.normal.data(), W, H, 1, false);
    S::u_BlockIDTexture = bind(cur.block.data(), W, H, 1, false);
    S::u_VarianceTexture = bind(a->variance, W, H, 1, true);
    S::u_TemporalMoment = bind(a->temporal_utility, W, H, 3, true);
    S::u_AO = bind(a->ao_sky, W, H, 2, true);
    FOR_EACH_PIXEL(S, {
        std::memcpy(a->o_sh + 4 * px, &S::o_SH[0], 16);
        std::memcpy(a->o_cocg + 2 * px, &S::o_CoCg[0], 8);
        a->o_variance[px] = S::o_Variance;
        std::memcpy(a->o_ao_sky + 2 * px, &S::o_AOAndSkylighting[0], 8);
    })
    return 0;
}

struct RefShadowFilterArgs {  // plain C layout, filled by oracle/ref_shaders.py
    const float* inv_view;
    const float* inv_proj;
    int32_t width, height, row_begin, row_end;
    const float* g_t;  const uint8_t* g_normal_id;  const float* prev_t;
    const uint8_t* shadow_u8;      // temporal pass: the raw 0 / 1 shadow plane
    const float* shadow;           // spatial pass: the temporal pass's output
    const float* transversal;
    const float* prev_shadow;  const float* frames;   // temporal: previous frame count; spatial: this frame's
    const float* prev_view;    const float* prev_projection;
    float filter_scale;
    float* o_shadow;  float* o_frames;
};

extern "C" __attribute__((visibility("default"))) int ref_shadow_temporal(const RefShadowFilterArgs* a) {
    using namespace glsl;
    namespace S = glsl::denoise::ns_ShadowTemporalFilter;
    const int W = a->width, H = a->height;
    const size_t n = (size_t)W * H;
    IdPlanes cur(a->g_normal_id, nullptr, n);
    std::vector<float> raw(n);
    for (size_t k = 0; k < n; ++k) raw[k] = (float)a->shadow_u8[k];   // o_Shadow 0 / 1 stored in R8 reads back as 0.0 / 1.0
    std::memcpy(&S::u_InverseView[0][0], a->inv_view, 64);
    std::memcpy(&S::u_InverseProjection[0][0], a->inv_proj, 64);
    std::memcpy(&S::u_PrevView[0][0], a->prev_view, 64);
    std::memcpy(&S::u_PrevProjection[0][0], a->prev_projection, 64);
    S::u_ShadowTemporal = true;                  // Pipeline.cpp:2879
    S::u_ShouldFilterShadows = true;
    S::v_RayOrigin = vec3(S::u_InverseView[3]);
    S::u_CurrentColorTexture = bind(raw.data(), W, H, 1, true);
    S::u_CurrentPositionTexture = bind(a->g_t, W, H, 1, true);
    S::u_PreviousColorTexture = bind(a->prev_shadow, W, H, 1, true);
    S::u_PreviousFramePositionTexture = bind(a->prev_t, W, H, 1, true);
    S::u_NormalTexture = bind(cur.normal.data(), W, H, 1, false);
    S::u_ShadowTransversals = bind(a->transversal, W, H, 1, true);
    S::u_FrameCount = bind(a->frames, W, H, 1, true);
    FOR_EACH_PIXEL(S, {
        a->o_shadow[px] = S::o_Color.x;
        a->o_frames[px] = S::o_Frames;
    })
    return 0;
}

extern "C" __attribute__((visibility("default"))) int ref_shadow_filter(const RefShadowFilterArgs* a) {
    using namespace glsl;
    namespace S = glsl::denoise::ns_ShadowFilter;
    const int W = a->width, H = a->height;
    IdPlanes cur(a->g_normal_id, nullptr, (size_t)W * H);
    std::memcpy(&S::u_InverseView[0][0], a->inv_view, 64);
    std::memcpy(&S::u_InverseProjection[0][0], a->inv_proj, 64);
    S::u_ShadowFilterScale = a->filter_scale;
    S::u_InputTexture = bind(a->shadow, W, H, 1, true);
    S::u_PositionTexture = bind(a->g_t, W, H, 1, true);
    S::u_NormalTexture = bind(cur.normal.data(), W, H, 1, false);
    S::u_IntersectionTransversals = bind(a->transversal, W, H, 1, true);
    S::u_FrameCount = bind(a->frames, W, H, 1, true);
    FOR_EACH_PIXEL(S, { a->o_shadow[px] = S::o_Color; })
    return 0;
}
