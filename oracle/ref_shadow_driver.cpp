// ref_shadow_driver.cpp — the reference's Core/Shaders/ShadowRayTraceFrag.glsl compiled as C++ (its own translation unit: the shader's
// #defines must not leak into the others).  See ref_shader_driver.cpp.  Test infrastructure only.
#include <cstdint>
#include <cstring>
#include <vector>

#include "glsl_compat.h"

namespace glsl {
#include "_ref/ShadowRayTraceFrag.inc"
}  // namespace glsl

static const int32_t* g_alpha_materials = nullptr;  // set by ref_trace_shadow_alpha around its call of ref_trace_shadow
static const uint8_t* g_alpha_mips = nullptr;
static int g_alpha_layers = 0;
static float g_alpha_fov = 60.0f;
extern "C" __attribute__((visibility("default"))) int ref_trace_shadow(
    const uint8_t* blocks, const uint8_t* df, const float* inv_view, const float* inv_proj, int width, int height, int row_begin, int row_end,
    const float* g_t, const uint8_t* g_normal_id, const float* light_dir, int frame, int soft, const float* halton, const uint8_t* noise_rgba8,
    float* o_shadow, float* o_transversal) {
    using namespace glsl;
    namespace S = ns_ShadowRayTraceFrag;
    S::u_VoxelData = sampler3D{blocks, 384, 128, 384};
    S::u_DistanceFieldTexture = sampler3D{df, 384, 128, 384};
    std::memcpy(&S::u_InverseView[0][0], inv_view, 16 * sizeof(float));
    std::memcpy(&S::u_InverseProjection[0][0], inv_proj, 16 * sizeof(float));
    S::u_Dimensions = vec2((float)width, (float)height);
    S::u_Halton = vec2(halton[0], halton[1]);
    S::u_LightDirection = vec3(light_dir[0], light_dir[1], light_dir[2]);
    S::u_CurrentFrame = frame;
    S::u_ContactHardeningShadows = soft != 0;
    S::u_ShouldAlphaTest = g_alpha_mips != nullptr;
    if (g_alpha_mips) {
        std::memcpy(S::SSBO_BlockData_storage, g_alpha_materials, 5 * 128 * sizeof(int32_t));
        S::u_AlbedoTextures = sampler2DArray{};
        S::u_AlbedoTextures.alpha_mips = g_alpha_mips;
        S::u_AlbedoTextures.alpha_layers = g_alpha_layers;
    }
    S::u_DoFullTrace = true;
    S::u_FOV = g_alpha_fov;
    S::u_Time = 0.0f;
    // o_Normal as the R8 attachment holds it: id / 10 (miss 1.0); the blue-noise texture as GL_RGBA8 unorm
    std::vector<float> normal((size_t)width * height), noise(256 * 256 * 4);
    for (size_t k = 0; k < normal.size(); ++k) normal[k] = g_normal_id[k] > 5 ? 1.0f : (float)g_normal_id[k] / 10.0f;
    for (size_t k = 0; k < noise.size(); ++k) noise[k] = (float)noise_rgba8[k] / 255.0f;
    S::u_PositionTexture = sampler2D{g_t, width, height, 1};
    S::u_NormalTexture = sampler2D{normal.data(), width, height, 1};
    S::u_BlueNoiseTexture = sampler2D{noise.data(), 256, 256, 4};
    S::v_RayOrigin = vec3(S::u_InverseView[3]);  // FBOVert.glsl:20
    for (int j = row_begin; j < row_end; ++j)
        for (int i = 0; i < width; ++i) {
            S::v_TexCoords = vec2(((float)i + 0.5f) / (float)width, ((float)j + 0.5f) / (float)height);
            gl_FragCoord = vec4((float)i + 0.5f, (float)j + 0.5f, 0.0f, 1.0f);
            S::shader_reset_globals();
            S::shader_main();
            o_shadow[(size_t)j * width + i] = S::o_Shadow;
            o_transversal[(size_t)j * width + i] = S::o_IntersectionTransversal;
        }
    return 0;
}

// the same with u_ShouldAlphaTest = true (ShadowRayTraceFrag.glsl:105-220, :502)
extern "C" __attribute__((visibility("default"))) int ref_trace_shadow_alpha(
    const uint8_t* blocks, const uint8_t* df, const float* inv_view, const float* inv_proj, int width, int height, int row_begin, int row_end,
    const float* g_t, const uint8_t* g_normal_id, const float* light_dir, int frame, int soft, const float* halton, const uint8_t* noise_rgba8,
    const int32_t* materials, const uint8_t* alpha_mips, int alpha_layers, float fov_degrees, float* o_shadow, float* o_transversal) {
    g_alpha_materials = materials; g_alpha_mips = alpha_mips; g_alpha_layers = alpha_layers; g_alpha_fov = fov_degrees;
    const int rc = ref_trace_shadow(blocks, df, inv_view, inv_proj, width, height, row_begin, row_end, g_t, g_normal_id, light_dir, frame, soft, halton,
                                    noise_rgba8, o_shadow, o_transversal);
    g_alpha_materials = nullptr; g_alpha_mips = nullptr; g_alpha_layers = 0; g_alpha_fov = 60.0f;
    return rc;
}
