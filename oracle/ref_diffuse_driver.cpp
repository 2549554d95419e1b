// ref_diffuse_driver.cpp — the reference's Core/Shaders/DiffuseRayTraceFrag.glsl (1-bounce diffuse GI, 1,300 lines) compiled as C++;
// uniforms and texture binds follow Core/Pipeline.cpp:2174-2281.  See ref_shader_driver.cpp.  Test infrastructure only.
#include <cstdint>
#include <cstring>
#include <vector>

#include "glsl_compat.h"

namespace glsl {
#include "_ref/DiffuseRayTraceFrag.inc"
}  // namespace glsl

struct RefDiffuseArgs {  // plain C layout, filled by oracle/ref_shaders.py
    const uint8_t* blocks;
    const uint8_t* df;
    const float* inv_view;
    const float* inv_proj;
    int32_t width, height, row_begin, row_end;
    const float* g_t;
    const uint8_t* g_normal_id;
    const int32_t* materials;  // 6 x 128 (BlockDataSSBO order)
    const int32_t* sobol;      // 65536
    const int32_t* scramble;   // 131072
    const int32_t* rank;       // 131072
    const float* albedo_lod3;  // [layers][64][64][4]
    const float* pbr_lod2;     // [layers][128][128][4]
    const float* emissive_lod0;  // [layers][512][512]
    const float* sky;          // [6][n][n][3]
    int32_t sky_n;
    int32_t spp, checker_spp, checkerboard, trace_length, frame, supersample;
    float halton[2], sun_dir[3], moon_dir[3];
    float sun_visibility, gi_sun_strength, gi_sky_strength, light_intensity;
    float* o_sh;      // 4 / pixel
    float* o_cocg;    // 2 / pixel
    float* o_utility; // 1 / pixel
    float* o_ao_sky;  // 2 / pixel
};

extern "C" __attribute__((visibility("default"))) int ref_trace_diffuse(const RefDiffuseArgs* a) {
    using namespace glsl;
    namespace S = ns_DiffuseRayTraceFrag;
    const int W = a->width, H = a->height;
    S::u_VoxelData = sampler3D{a->blocks, 384, 128, 384};
    S::u_DistanceFieldTexture = sampler3D{a->df, 384, 128, 384};
    std::memcpy(&S::u_InverseView[0][0], a->inv_view, 16 * sizeof(float));
    std::memcpy(&S::u_InverseProjection[0][0], a->inv_proj, 16 * sizeof(float));
    S::u_Dimensions = vec2((float)W, (float)H);
    S::u_Halton = vec2(a->halton[0], a->halton[1]);
    S::u_Time = 0.0f;
    S::u_Supersample = a->supersample != 0;
    S::u_APPLY_PLAYER_SHADOW = false;
    S::u_UseDirectSampling = false;
    S::u_SPP = a->spp;
    S::u_CheckerSPP = a->checker_spp;
    S::CHECKERBOARD_SPP = a->checkerboard != 0;
    S::u_DiffuseTraceLength = a->trace_length;
    S::u_CurrentFrame = a->frame;
    S::u_CurrentFrameMod512 = a->frame % 512;
    S::u_CurrentFrameMod128 = a->frame % 128;
    S::u_UseBlueNoise = true;
    S::u_GISunStrength = a->gi_sun_strength;
    S::u_GISkyStrength = a->gi_sky_strength;
    S::u_SunVisibility = a->sun_visibility;
    S::u_DiffuseLightIntensity = a->light_intensity;
    S::u_ViewerPosition = vec3(S::u_InverseView[3]);
    S::u_SunDirection = vec3(a->sun_dir[0], a->sun_dir[1], a->sun_dir[2]);
    S::u_MoonDirection = vec3(a->moon_dir[0], a->moon_dir[1], a->moon_dir[2]);
    // SSBOs: BlockDataSSBO (Core/BlockDataSSBO.cpp:28-35) and the three blue-noise tables; the slack behind the last table holds its
    // last element, the value a clamped out-of-range read returns
    std::memcpy(S::BlockAlbedoData, a->materials + 0 * 128, 128 * sizeof(int));
    std::memcpy(S::BlockNormalData, a->materials + 1 * 128, 128 * sizeof(int));
    std::memcpy(S::BlockPBRData, a->materials + 2 * 128, 128 * sizeof(int));
    std::memcpy(S::BlockEmissiveData, a->materials + 3 * 128, 128 * sizeof(int));
    std::memcpy(S::BlockTransparentData, a->materials + 4 * 128, 128 * sizeof(int));
    std::memcpy(S::sobol_256spp_256d, a->sobol, 65536 * sizeof(int));
    std::memcpy(S::scramblingTile, a->scramble, 131072 * sizeof(int));
    std::memcpy(S::rankingTile, a->rank, 131072 * sizeof(int));
    for (int k = 0; k < 4096; ++k) S::rankingTile[131072 + k] = a->rank[131071];
    std::vector<float> normal((size_t)W * H);
    for (size_t k = 0; k < normal.size(); ++k) normal[k] = a->g_normal_id[k] > 5 ? 1.0f : (float)a->g_normal_id[k] / 10.0f;
    S::u_PositionTexture = sampler2D{a->g_t, W, H, 1};
    S::u_NormalTexture = sampler2D{normal.data(), W, H, 1};
    S::u_Skymap = samplerCube{a->sky, a->sky_n};
    S::u_BlockAlbedoTextures = sampler2DArray{a->albedo_lod3, 64, 4};    // textureLod(.., 3.0f)
    S::u_BlockPBRTextures = sampler2DArray{a->pbr_lod2, 128, 4};         // textureLod(.., 2.0f)
    S::u_BlockEmissiveTextures = sampler2DArray{a->emissive_lod0, 512, 1};  // texture(): level 0
    S::v_RayOrigin = vec3(S::u_InverseView[3]);  // FBOVert.glsl:20
    for (int j = a->row_begin; j < a->row_end; ++j)
        for (int i = 0; i < W; ++i) {
            const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
            S::v_TexCoords = vec2(u, v);
            gl_FragCoord = vec4((float)i + 0.5f, (float)j + 0.5f, 0.0f, 1.0f);
            // FBOVert.glsl:17-19 evaluated at the pixel (the varying is linear in the quad position)
            const vec4 clip = vec4(u * 2.0f - 1.0f, v * 2.0f - 1.0f, -1.0f, 1.0f);
            const vec4 eye = vec4(vec2(S::u_InverseProjection * clip), -1.0f, 0.0f);
            S::v_RayDirection = vec3(S::u_InverseView * eye);
            S::shader_reset_globals();
            S::shader_main();
            const size_t px = (size_t)j * W + i;
            std::memcpy(a->o_sh + 4 * px, &S::o_SH[0], 4 * sizeof(float));
            std::memcpy(a->o_cocg + 2 * px, &S::o_CoCg[0], 2 * sizeof(float));
            a->o_utility[px] = S::o_Utility;
            std::memcpy(a->o_ao_sky + 2 * px, &S::o_AOAndSkyLighting[0], 2 * sizeof(float));
        }
    return 0;
}
