// ref_reflection_driver.cpp — the reference's Core/Shaders/ReflectionTraceFrag.glsl compiled as C++, driven in the v1 parity profile of
// include/vxpt.h (no screen-space reprojection, LPV, cloud or player reflections; lava animation off); uniforms and binds follow
// Core/Pipeline.cpp:3003-3164.  See ref_shader_driver.cpp.  Test infrastructure only.
#include <cstdint>
#include <cstring>
#include <vector>

#include "glsl_compat.h"

namespace glsl {
#include "_ref/ReflectionTraceFrag.inc"
}  // namespace glsl

struct RefReflectionArgs {  // plain C layout, filled by oracle/ref_shaders.py
    const uint8_t* blocks;
    const uint8_t* df;
    const float* inv_view;
    const float* inv_proj;
    int32_t width, height, row_begin, row_end;
    const float* g_t;
    const uint8_t* g_normal_id;
    const uint8_t* g_block_id;
    const float* g_normal;     // 3 / pixel
    const float* g_pbr;        // 4 / pixel
    const float* sh;           // 4 / pixel
    const float* cocg;         // 2 / pixel
    const int32_t* materials;  // 6 x 128
    const int32_t* sobol;
    const int32_t* scramble;
    const int32_t* rank;
    const float* albedo_lod3;    // [layers][64][64][4]
    const float* normal_lod3;    // [layers][64][64][4]
    const float* pbr_lod2;       // [layers][128][128][4]
    const float* emissive_lod2;  // [layers][128][128]
    const float* sky;
    int32_t sky_n;
    int32_t spp, trace_length, frame, rough, roughness_bias, checkerboard;
    float sun_dir[3], moon_dir[3], stronger_dir[3], viewer_pos[3];
    float sun_strength, moon_strength, halton[2];
    int32_t grass_props[10];
    float* o_color;         // 4 / pixel
    float* o_hit_distance;  // 1 / pixel
    float* o_emissive_mask; // 1 / pixel
};

extern "C" __attribute__((visibility("default"))) int ref_trace_reflection(const RefReflectionArgs* a) {
    using namespace glsl;
    namespace S = ns_ReflectionTraceFrag;
    const int W = a->width, H = a->height;
    S::u_VoxelData = sampler3D{a->blocks, 384, 128, 384};
    S::u_DistanceFieldTexture = sampler3D{a->df, 384, 128, 384};
    std::memcpy(&S::u_InverseView[0][0], a->inv_view, 16 * sizeof(float));
    std::memcpy(&S::u_InverseProjection[0][0], a->inv_proj, 16 * sizeof(float));
    S::u_Dimensions = vec2((float)W, (float)H);
    S::u_Halton = vec2(a->halton[0], a->halton[1]);
    S::u_Time = 0.0f;
    S::u_SPP = a->spp;
    S::u_ReflectionTraceLength = a->trace_length;
    S::TEMPORAL_SPEC = a->frame >= 0;
    S::u_CurrentFrame = a->frame;
    S::u_CurrentFrameMod128 = a->frame >= 0 ? a->frame % 128 : 0;
    S::u_TemporalFilterReflections = true;
    S::u_RoughReflections = a->rough != 0;
    S::u_RoughnessBias = a->roughness_bias != 0;
    S::CHECKERBOARD_SPEC_SPP = a->checkerboard != 0;
    S::u_UseBlueNoise = true;
    S::u_ReprojectToScreenSpace = false;
    S::u_LPVGI = false;
    S::u_QualityLPVGI = false;
    S::u_CloudReflections = false;
    S::u_ReflectPlayer = false;
    S::u_DeriveFromDiffuseSH = false;
    S::u_UseDecoupledGI = false;
    S::u_ScreenSpaceSkylightingValid = false;
    S::u_LavaBlockID = -1;
    S::u_SunStrengthModifier = a->sun_strength;
    S::u_MoonStrengthModifier = a->moon_strength;
    S::u_SunDirection = vec3(a->sun_dir[0], a->sun_dir[1], a->sun_dir[2]);
    S::u_MoonDirection = vec3(a->moon_dir[0], a->moon_dir[1], a->moon_dir[2]);
    S::u_StrongerLightDirection = vec3(a->stronger_dir[0], a->stronger_dir[1], a->stronger_dir[2]);
    S::u_ViewerPosition = vec3(a->viewer_pos[0], a->viewer_pos[1], a->viewer_pos[2]);
    for (int k = 0; k < 10; ++k) S::u_GrassBlockProps[k] = a->grass_props[k];
    std::memcpy(S::BlockAlbedoData, a->materials + 0 * 128, 128 * sizeof(int));
    std::memcpy(S::BlockNormalData, a->materials + 1 * 128, 128 * sizeof(int));
    std::memcpy(S::BlockPBRData, a->materials + 2 * 128, 128 * sizeof(int));
    std::memcpy(S::BlockEmissiveData, a->materials + 3 * 128, 128 * sizeof(int));
    std::memcpy(S::BlockTransparentData, a->materials + 4 * 128, 128 * sizeof(int));
    std::memcpy(S::sobol_256spp_256d, a->sobol, 65536 * sizeof(int));
    std::memcpy(S::scramblingTile, a->scramble, 131072 * sizeof(int));
    std::memcpy(S::rankingTile, a->rank, 131072 * sizeof(int));
    for (int k = 0; k < 4096; ++k) S::rankingTile[131072 + k] = a->rank[131071];
    std::vector<float> normal((size_t)W * H), block((size_t)W * H);
    for (size_t k = 0; k < normal.size(); ++k) {
        normal[k] = a->g_normal_id[k] > 5 ? 1.0f : (float)a->g_normal_id[k] / 10.0f;
        block[k] = (float)a->g_block_id[k] / 255.0f;
    }
    // attachment 0 of the primary FBO (R16F distance) is GL_LINEAR, attachment 1 (normal id) GL_NEAREST, both GL_REPEAT (Core/Pipeline.cpp:1094,
    // Core/GLClasses/Framebuffer.cpp:64-67): the shader reads both at the Halton-jittered coordinate (:754-777)
    S::u_PositionTexture = sampler2D{a->g_t, W, H, 1, true, true};
    S::u_InitialTraceNormalTexture = sampler2D{normal.data(), W, H, 1, false, true};
    S::u_BlockIDTex = sampler2D{block.data(), W, H, 1};
    S::u_GBufferNormals = sampler2D{a->g_normal, W, H, 3};
    S::u_GBufferPBR = sampler2D{a->g_pbr, W, H, 4};
    S::u_DiffuseSH = sampler2D{a->sh, W, H, 4};
    S::u_DiffuseCoCg = sampler2D{a->cocg, W, H, 2};
    S::u_Skymap = samplerCube{a->sky, a->sky_n};
    S::u_BlockAlbedoTextures = sampler2DArray{a->albedo_lod3, 64, 4, true};
    S::u_BlockNormalTextures = sampler2DArray{a->normal_lod3, 64, 4, true};
    S::u_BlockPBRTextures = sampler2DArray{a->pbr_lod2, 128, 4, true};
    S::u_BlockEmissiveTextures = sampler2DArray{a->emissive_lod2, 128, 1, true};
    S::v_RayOrigin = vec3(S::u_InverseView[3]);
    for (int j = a->row_begin; j < a->row_end; ++j)
        for (int i = 0; i < W; ++i) {
            const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
            S::v_TexCoords = vec2(u, v);
            gl_FragCoord = vec4((float)i + 0.5f, (float)j + 0.5f, 0.0f, 1.0f);
            const vec4 clip = vec4(u * 2.0f - 1.0f, v * 2.0f - 1.0f, -1.0f, 1.0f);
            const vec4 eye = vec4(vec2(S::u_InverseProjection * clip), -1.0f, 0.0f);
            S::v_RayDirection = vec3(S::u_InverseView * eye);
            S::shader_reset_globals();
            S::shader_main();
            const size_t px = (size_t)j * W + i;
            std::memcpy(a->o_color + 4 * px, &S::o_Color[0], 4 * sizeof(float));
            a->o_hit_distance[px] = S::o_HitDistance;
            a->o_emissive_mask[px] = S::o_EmissivityHitMask;
        }
    return 0;
}
