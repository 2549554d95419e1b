// ref_ambient_driver.cpp — the reference's Core/Shaders/EstimateAmbientSoundLevel.comp compiled as C++ (own translation unit; see
// ref_shader_driver.cpp).  Test infrastructure only.  Dispatch shape: Core/Pipeline.cpp:1921 glDispatchCompute(2, 1, 1) with the shader's
// local size 4 x 4 -> gl_GlobalInvocationID.xy in [0, 8) x [0, 4); the SSBO is cleared before every dispatch (:1931-1934).
#include <cstdint>

#include "glsl_compat.h"

namespace glsl {
#include "_ref/EstimateAmbientSoundLevel.inc"
}  // namespace glsl

extern "C" __attribute__((visibility("default"))) int ref_ambient_sound(const uint8_t* df, const float* player_pos, int frame, uint32_t* aggregate,
                                                                       uint32_t* per_invocation /* 32 */) {
    using namespace glsl;
    namespace S = ns_EstimateAmbientSoundLevel;
    S::u_DistanceField = sampler3D{df, 384, 128, 384};
    S::u_PlayerPosition = vec3(player_pos[0], player_pos[1], player_pos[2]);
    S::u_Frame = frame;
    S::SkyLevelAggregate = 0u;
    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 8; ++x) {
            gl_GlobalInvocationID = uvec3(x, y, 0);
            S::shader_reset_globals();
            const uint32_t before = S::SkyLevelAggregate;
            S::shader_main();
            if (per_invocation) per_invocation[y * 8 + x] = S::SkyLevelAggregate - before;
        }
    *aggregate = S::SkyLevelAggregate;
    return 0;
}
