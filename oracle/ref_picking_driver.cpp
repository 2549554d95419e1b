// ref_picking_driver.cpp — the reference's OWN CPU picking, World::Raycast and World::RaycastDetect (Core/World.cpp:215-546), compiled from
// the source where it lies: oracle/Makefile cuts those lines into _ref/World_picking_body.inc (never copied into the repository) and this
// file supplies just enough of the surrounding classes for them to compile — the block grid, and do-nothing stand-ins for what picking
// touches besides the grid (light-propagation queues, particles, sound, the GL texture update).  Test infrastructure only: it pins
// voxelpathtracer_b200/world.py::World.raycast / raycast_detect and host/VoxelRT.h (tests/test_picking_vs_reference.py).
// Built at -O0: RaycastDetect falls off its end without a return when nothing is hit (the port pins that case as "no hit"); the test
// calls it only for rays that do hit.
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <queue>

#include <glm/glm.hpp>

#define WORLD_SIZE_X 384
#define WORLD_SIZE_Y 128
#define WORLD_SIZE_Z 384
#define GL_TEXTURE_3D 0
#define GL_RED 0
#define GL_UNSIGNED_BYTE 0
static inline void glBindTexture(int, unsigned) {}
static inline void glTexSubImage3D(int, int, int, int, int, int, int, int, int, int, const void*) {}

namespace VoxelRT {
struct Block {
    uint8_t block;
};
struct LightRemovalNode {
    LightRemovalNode(const glm::vec3&, int) {}
};
struct LightNode {
    LightNode(const glm::vec3&) {}
};
namespace Volumetrics {
static std::queue<LightRemovalNode> g_removal;
static std::queue<LightNode> g_light;
static inline std::queue<LightRemovalNode>& GetLightRemovalBFSQueue() { return g_removal; }
static inline std::queue<LightNode>& GetLightBFSQueue() { return g_light; }
static inline int GetLightValue(const glm::ivec3&) { return 0; }
static inline void SetLightValue(const glm::ivec3&, int, int) {}
static inline void UploadLight(const glm::ivec3&, int, int, bool) {}
static inline void AddLightToVolume(const glm::ivec3&, uint8_t) {}
static inline void DepropogateVolume() {}
static inline void PropogateVolume() {}
}  // namespace Volumetrics
namespace BlockDatabase {
static const int32_t* g_emissive = nullptr;  // BlockEmissiveData (materials table rows 384..511); -1 = none
static inline int GetBlockEmissiveTexture(uint8_t b) { return (g_emissive && b < 128) ? g_emissive[b] : -1; }
static inline bool HasEmissiveTexture(uint8_t b) { return GetBlockEmissiveTexture(b) >= 0; }
}  // namespace BlockDatabase
namespace SoundManager {
static inline void PlayBlockSound(uint8_t, const glm::vec3&, bool) {}
}
struct ParticleEmitterStub {
    void EmitParticlesAt(const glm::vec3&, float, int, const glm::vec3&, const glm::vec3&, const glm::vec3&, uint8_t) {}
};
struct TextureStub {
    unsigned GetTextureID() const { return 0; }
};
class World {
public:
    uint8_t* m_WorldData = nullptr;  // x + 384 * (y + 128 * z), Core/World.h:46-70
    uint8_t m_CurrentlyHeldBlock = 0;
    bool m_Buffered = false;         // no GL texture behind this world: the edit stays in the host grid
    TextureStub m_DataTexture;
    ParticleEmitterStub m_ParticleEmitter;
    Block GetBlock(int x, int y, int z) const { return Block{m_WorldData[(size_t)x + 384 * ((size_t)y + 128 * (size_t)z)]}; }
    void SetBlock(int x, int y, int z, Block b) { m_WorldData[(size_t)x + 384 * ((size_t)y + 128 * (size_t)z)] = b.block; }
    void GenerateDistanceField() {}
    void InsertToLightList(const glm::vec3&) {}
    void RemoveFromLightList(const glm::vec3&) {}
    bool Raycast(uint8_t op, glm::vec3 pos, const glm::vec3& dir, const glm::vec3& acceleration, bool is_falling, float dt);
    glm::ivec4 RaycastDetect(const glm::vec3& pos, const glm::vec3& dir);
};
}  // namespace VoxelRT

using namespace VoxelRT;
using namespace glm;
#include "_ref/World_picking_body.inc"

extern "C" {
// op 0 break / 1 place / 2 pick on `blocks` (edited in place).  Returns the reference's return value; *held receives m_CurrentlyHeldBlock
__attribute__((visibility("default"))) int ref_world_raycast(uint8_t* blocks, const int32_t* emissive_table, int op, const float* pos, const float* dir,
                                                             int held_in, int* held_out) {
    std::cout.setstate(std::ios_base::failbit);  // "LAMP PLACED" chatter
    World w;
    w.m_WorldData = blocks;
    w.m_CurrentlyHeldBlock = (uint8_t)held_in;
    BlockDatabase::g_emissive = emissive_table;
    const bool r = w.Raycast((uint8_t)op, glm::vec3(pos[0], pos[1], pos[2]), glm::vec3(dir[0], dir[1], dir[2]), glm::vec3(0.0f), false, 0.0f);
    *held_out = w.m_CurrentlyHeldBlock;
    return r ? 1 : 0;
}
// only defined for rays that hit something within the 48 steps (see the header)
__attribute__((visibility("default"))) void ref_world_raycast_detect(uint8_t* blocks, const float* pos, const float* dir, int out[4]) {
    World w;
    w.m_WorldData = blocks;
    const glm::ivec4 r = w.RaycastDetect(glm::vec3(pos[0], pos[1], pos[2]), glm::vec3(dir[0], dir[1], dir[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
}
