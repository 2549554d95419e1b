// ref_shader_driver.cpp — runs the REFERENCE'S OWN shader source on the CPU (test infrastructure only).
//
// oracle/Makefile translates Core/Shaders/ManhattanDistance{X,Y,Z}.comp and Core/Shaders/InitialRayTraceFrag.glsl, read where they lie
// under /root/reference, into oracle/_ref/*.inc (oracle/glsl2cpp.py rewrites declarations only) and compiles them here against the
// reference's vendored glm with -O2 -ffp-contract=off.  The result, oracle/_ref/libref_shaders.so, executes every expression of those
// shaders as written: it is the "reference run here" that pins oracle/vxo_oracle.cpp (tests/test_oracle_vs_reference_shaders.py)
// and that made tests/golden/ref_shader_digests.json (tools/make_ref_shader_golden.py).
//
// Dispatch shapes follow Core/World.cpp:69-113 (one invocation per grid line, X then Y then Z with a barrier between) and the
// full-screen quad of Core/Shaders/FBOVert.glsl:13-21 (v_TexCoords = pixel centre / dimensions).
#include <cstdint>
#include <cstring>

#include "glsl_compat.h"

namespace glsl {
#include "_ref/ManhattanDistanceX.inc"
#include "_ref/ManhattanDistanceY.inc"
#include "_ref/ManhattanDistanceZ.inc"
#include "_ref/InitialRayTraceFrag.inc"
}  // namespace glsl

extern "C" {

// World::GenerateDistanceField: blocks[384*128*384] -> df (same layout)
__attribute__((visibility("default"))) int ref_df_build(const uint8_t* blocks, uint8_t* df) {
    using namespace glsl;
    const int SX = 384, SY = 128, SZ = 384;
    std::memset(df, 0, (size_t)SX * SY * SZ);
    ns_ManhattanDistanceX::u_BlockData = sampler3D{blocks, SX, SY, SZ};
    ns_ManhattanDistanceX::o_DistanceBuffer = image3D{df, SX, SY, SZ};
    ns_ManhattanDistanceY::o_DistanceBuffer = image3D{df, SX, SY, SZ};
    ns_ManhattanDistanceZ::o_DistanceBuffer = image3D{df, SX, SY, SZ};
    // invocations of one dispatch touch disjoint grid lines, so they may run on any number of threads; the loops end = the barriers
#pragma omp parallel for schedule(static)
    for (int z = 0; z < SZ; ++z)  // glDispatchCompute(1, WORLD_SIZE_Y / 32, WORLD_SIZE_Z / 32), local size (1, 32, 32)
        for (int y = 0; y < SY; ++y) {
            gl_GlobalInvocationID = uvec3(0, y, z);
            ns_ManhattanDistanceX::shader_main();
        }
#pragma omp parallel for schedule(static)
    for (int z = 0; z < SZ; ++z)
        for (int x = 0; x < SX; ++x) {
            gl_GlobalInvocationID = uvec3(x, 0, z);
            ns_ManhattanDistanceY::shader_main();
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < SY; ++y)
        for (int x = 0; x < SX; ++x) {
            gl_GlobalInvocationID = uvec3(x, y, 0);
            ns_ManhattanDistanceZ::shader_main();
        }
    return 0;
}

// InitialRayTraceFrag.glsl main() over rows [row_begin, row_end) of a width x height frame.
// Outputs are the four colour attachments as the shader writes them (fp32, before any render-target conversion).
static const int32_t* g_alpha_materials = nullptr;  // set by ref_trace_primary_alpha around its call of ref_trace_primary
static const uint8_t* g_alpha_mips = nullptr;
static int g_alpha_layers = 0;
static float g_alpha_fov = 60.0f;
__attribute__((visibility("default"))) int ref_trace_primary(const uint8_t* blocks, const uint8_t* df, const float* inv_view, const float* inv_proj, int width,
                                                             int height, int row_begin, int row_end, int render_distance, int jitter_enable,
                                                             const float* jitter, float* o_hit_distance, float* o_normal, float* o_block_id,
                                                             float* o_depth_nonlinear) {
    using namespace glsl;
    namespace S = ns_InitialRayTraceFrag;
    S::u_VoxelDataTexture = sampler3D{blocks, 384, 128, 384};
    S::u_DistanceFieldTexture = sampler3D{df, 384, 128, 384};
    std::memcpy(&S::u_InverseView[0][0], inv_view, 16 * sizeof(float));       // column-major, as glUniformMatrix4fv(GL_FALSE)
    std::memcpy(&S::u_InverseProjection[0][0], inv_proj, 16 * sizeof(float));
    S::u_Dimensions = vec2((float)width, (float)height);
    S::u_ShouldAlphaTest = g_alpha_mips != nullptr;
    if (g_alpha_mips) {  // BlockDataSSBO (Core/BlockDataSSBO.cpp:28-35 order) and the albedo array's alpha mip chain
        std::memcpy(S::SSBO_BlockData_storage, g_alpha_materials, 5 * 128 * sizeof(int32_t));
        S::u_AlbedoTextures = sampler2DArray{};
        S::u_AlbedoTextures.alpha_mips = g_alpha_mips;
        S::u_AlbedoTextures.alpha_layers = g_alpha_layers;
    }
    S::u_RenderDistance = render_distance;
    S::u_JitterSceneForTAA = jitter_enable != 0;
    S::u_CurrentTAAJitter = vec2(jitter[0], jitter[1]);
    S::u_FOV = g_alpha_fov;
    S::u_Time = 0.0f;
    for (int j = row_begin; j < row_end; ++j)
        for (int i = 0; i < width; ++i) {
            S::v_TexCoords = vec2(((float)i + 0.5f) / (float)width, ((float)j + 0.5f) / (float)height);
            S::shader_reset_globals();
            S::shader_main();
            const size_t px = (size_t)j * width + i;
            o_hit_distance[px] = S::o_HitDistance;
            o_normal[px] = S::o_Normal;
            o_block_id[px] = S::o_BlockID;
            o_depth_nonlinear[px] = S::o_DepthNonLinear;
        }
    return 0;
}

// the same with u_ShouldAlphaTest = true (VoxelTraversalDF_AlphaTest + StopRay, InitialRayTraceFrag.glsl:189-305)
__attribute__((visibility("default"))) int ref_trace_primary_alpha(const uint8_t* blocks, const uint8_t* df, const float* inv_view, const float* inv_proj,
                                                                   int width, int height, int row_begin, int row_end, int render_distance,
                                                                   int jitter_enable, const float* jitter, const int32_t* materials,
                                                                   const uint8_t* alpha_mips, int alpha_layers, float fov_degrees,
                                                                   float* o_hit_distance, float* o_normal, float* o_block_id, float* o_depth_nonlinear) {
    g_alpha_materials = materials; g_alpha_mips = alpha_mips; g_alpha_layers = alpha_layers; g_alpha_fov = fov_degrees;
    const int rc = ref_trace_primary(blocks, df, inv_view, inv_proj, width, height, row_begin, row_end, render_distance, jitter_enable, jitter,
                                     o_hit_distance, o_normal, o_block_id, o_depth_nonlinear);
    g_alpha_materials = nullptr; g_alpha_mips = nullptr; g_alpha_layers = 0; g_alpha_fov = 60.0f;
    return rc;
}

}  // extern "C"
