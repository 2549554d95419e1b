// ref_gbuffer_driver.cpp — the reference's Core/Shaders/GenerateGBuffer.glsl compiled as C++, driven in the v1 parity profile of
// include/vxpt.h (no lava block); uniforms and binds follow Core/Pipeline.cpp:2066-2136.  The shader takes screen-space
// derivatives, so it runs in 2x2 quads: a record run and a replay run per quad (glsl_compat.h, QuadDerivatives).  See
// ref_shader_driver.cpp.  Test infrastructure only.
#include <cstdint>
#include <cstring>
#include <vector>

#include "glsl_compat.h"

namespace glsl {
#define discard return
#include "_ref/GenerateGBuffer.inc"
#undef discard
}  // namespace glsl

struct RefGBufferArgs {  // plain C layout, filled by oracle/ref_shaders.py
    const float* inv_view;
    const float* inv_proj;
    int32_t width, height, row_begin, row_end;
    const float* g_inv_t;
    const uint8_t* g_normal_id;
    const uint8_t* g_block_id;
    const int32_t* materials;   // 6 x 128
    const uint8_t* albedo_mips;
    const uint8_t* normal_mips;
    const uint8_t* pbr_mips;
    int32_t n_mip_layers;
    const float* emissive_lod0;  // [layers][512][512]
    int32_t update_this_frame;
    int32_t grass_props[10];
    int32_t pom, high_quality_pom, dither_pom, frame;
    float pom_height, pom_exp;
    int32_t lava_block_id;
    float time;
    const uint8_t* lava_albedo;  // [8][256][256][4]
    const uint8_t* lava_normal;
    float* o_albedo;      // 3 / pixel
    float* o_normal;      // 3 / pixel
    float* o_pbr;         // 4 / pixel
    float* o_texture_ao;  // 1 / pixel
};

extern "C" __attribute__((visibility("default"))) int ref_generate_gbuffer(const RefGBufferArgs* a) {
    using namespace glsl;
    namespace S = ns_GenerateGBuffer;
    const int W = a->width, H = a->height;
    std::memcpy(&S::u_InverseView[0][0], a->inv_view, 16 * sizeof(float));
    std::memcpy(&S::u_InverseProjection[0][0], a->inv_proj, 16 * sizeof(float));
    S::u_LavaBlockID = a->lava_block_id;
    S::u_Time = a->time;
    S::uTime = a->time;
    S::u_LavaTextures[0] = sampler3D();
    S::u_LavaTextures[1] = sampler3D();
    S::u_LavaTextures[0].rgba = a->lava_albedo; S::u_LavaTextures[0].rn = 256; S::u_LavaTextures[0].rframes = 8;
    S::u_LavaTextures[1].rgba = a->lava_normal; S::u_LavaTextures[1].rn = 256; S::u_LavaTextures[1].rframes = 8;
    S::u_Frame = a->frame;
    S::u_UpdateGBufferThisFrame = a->update_this_frame != 0;
    S::u_POM = a->pom != 0;
    S::u_HighQualityPOM = a->high_quality_pom != 0;
    S::u_DitherPOM = a->dither_pom != 0;
    S::u_POMHeight = a->pom_height;
    S::u_POMExp = a->pom_exp;
    for (int k = 0; k < 10; ++k) S::u_GrassBlockProps[k] = a->grass_props[k];
    std::memcpy(S::BlockAlbedoData, a->materials + 0 * 128, 128 * sizeof(int));
    std::memcpy(S::BlockNormalData, a->materials + 1 * 128, 128 * sizeof(int));
    std::memcpy(S::BlockPBRData, a->materials + 2 * 128, 128 * sizeof(int));
    std::memcpy(S::BlockEmissiveData, a->materials + 3 * 128, 128 * sizeof(int));
    std::memcpy(S::BlockTransparentData, a->materials + 4 * 128, 128 * sizeof(int));
    std::vector<float> normal((size_t)W * H), block((size_t)W * H);
    for (size_t k = 0; k < normal.size(); ++k) {
        normal[k] = a->g_normal_id[k] > 5 ? 1.0f : (float)a->g_normal_id[k] / 10.0f;
        block[k] = (float)a->g_block_id[k] / 255.0f;
    }
    S::u_NonLinearDepth = sampler2D{a->g_inv_t, W, H, 1};
    S::u_Normals = sampler2D{normal.data(), W, H, 1};
    S::u_BlockIDs = sampler2D{block.data(), W, H, 1};
    sampler2DArray arr;
    arr.mips = a->albedo_mips; arr.mip_layers = a->n_mip_layers; arr.srgb = true; arr.mag_linear = false;   // TextureArray.cpp:74 defaults
    S::u_BlockAlbedos = arr;
    arr.mips = a->normal_mips; arr.srgb = false; arr.mag_linear = true;                                      // BlockDatabase.cpp:78
    S::u_BlockNormals = arr;
    arr.mips = a->pbr_mips;                                                                                   // BlockDatabase.cpp:82
    S::u_BlockPBR = arr;
    S::u_BlockEmissive = sampler2DArray{a->emissive_lod0, 512, 1, false};  // texture(): bilinear on level 0, as in the GI pass
    for (int j0 = a->row_begin & ~1; j0 < a->row_end; j0 += 2)
        for (int i0 = 0; i0 < W; i0 += 2) {
            gl_Quad = QuadDerivatives();
            for (int run = 0; run < 2; ++run) {
                gl_Quad.mode = run;
                for (int lane = 0; lane < 4; ++lane) {
                    const int i = i0 + (lane & 1), j = j0 + (lane >> 1);
                    if (i >= W || j >= H) continue;  // helper invocation outside the frame: records nothing
                    gl_Quad.lane = lane;
                    gl_Quad.call = 0;
                    S::v_TexCoords = vec2(((float)i + 0.5f) / (float)W, ((float)j + 0.5f) / (float)H);
                    gl_FragCoord = vec4((float)i + 0.5f, (float)j + 0.5f, 0.0f, 1.0f);
                    S::shader_reset_globals();
                    const size_t px = (size_t)j * W + i;
                    // a discarded fragment leaves the attachment untouched: seed the outputs with what the planes hold
                    S::o_Albedo = vec3(a->o_albedo[3 * px], a->o_albedo[3 * px + 1], a->o_albedo[3 * px + 2]);
                    S::o_Normal = vec3(a->o_normal[3 * px], a->o_normal[3 * px + 1], a->o_normal[3 * px + 2]);
                    S::o_PBR = vec4(a->o_pbr[4 * px], a->o_pbr[4 * px + 1], a->o_pbr[4 * px + 2], a->o_pbr[4 * px + 3]);
                    S::o_TextureAO = a->o_texture_ao[px];
                    S::shader_main();
                    if (run == 1 && j >= a->row_begin && j < a->row_end) {
                        std::memcpy(a->o_albedo + 3 * px, &S::o_Albedo[0], 3 * sizeof(float));
                        std::memcpy(a->o_normal + 3 * px, &S::o_Normal[0], 3 * sizeof(float));
                        std::memcpy(a->o_pbr + 4 * px, &S::o_PBR[0], 4 * sizeof(float));
                        a->o_texture_ao[px] = S::o_TextureAO;
                    }
                }
            }
        }
    return 0;
}
