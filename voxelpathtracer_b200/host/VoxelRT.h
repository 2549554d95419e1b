// VoxelRT.h — headless C++17 host mirror of the reference's interfaces for the hot path, over the C ABI (include/vxpt.h).
//
// Same names, argument meaning and error behaviour as the reference classes it stands in for, minus everything that
// needs a window, GL context or audio device:
//   VoxelRT::Block, VoxelRT::World           Core/Block.h:7-10, Core/World.h:35-70,167-171,192, Core/World.cpp:48-113
//   VoxelRT::World::Raycast / RaycastDetect  Core/World.cpp:215-546 (CPU picking: place / break / pick the block looked at)
//   VoxelRT::GenerateWorld                   Core/WorldGenerator.cpp:69-121 (superflat; plains from a column table)
//   VoxelRT::SaveWorld / LoadWorld           Core/WorldFileHandler.cpp:10-83
//   VoxelRT::BlockDataSSBO                   Core/BlockDataSSBO.cpp:5-46
//   VoxelRT::BlueNoiseDataSSBO               Core/BlueNoiseDataSSBO.cpp:17-37
//   VoxelRT::FPSCamera                       Core/FpsCamera.cpp:23-24,150-163 (lookAt / perspective, right-handed)
//   VoxelRT::GetTAAJitter(…)                 Core/TAAJitter.cpp:6-47
// The reference logs failures and carries on (Core/GLClasses/Shader.cpp:39-155); so does this layer: every failing ABI
// call is reported through VoxelRT::Logger::Log and the method returns false.
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/vxpt.h"

#define WORLD_SIZE_X VXPT_WORLD_SIZE_X
#define WORLD_SIZE_Y VXPT_WORLD_SIZE_Y
#define WORLD_SIZE_Z VXPT_WORLD_SIZE_Z

namespace VoxelRT {

struct Logger {
    static void Log(const std::string& txt) { std::fprintf(stderr, "[VoxelRT] %s\n", txt.c_str()); }
};

inline bool vx_ok(int rc, const char* what) {
    if (rc == VXPT_OK) return true;
    Logger::Log(std::string(what) + " failed: " + vxpt_last_error());
    return false;
}

struct Block {
    uint8_t block;
};

// block ids = position in blockdb.txt (Core/BlockDatabaseParser.cpp:31-42)
namespace BlockID {
constexpr uint8_t Air = 0, Grass = 1, Dirt = 2, Stone = 3, Cobblestone = 4, Sand = 5, Lamp = 12;
}

class World {
public:
    World() : m_WorldData(new std::array<Block, WORLD_SIZE_X * WORLD_SIZE_Y * WORLD_SIZE_Z>()) {
        std::memset(m_WorldData->data(), 0, m_WorldData->size());
        m_Buffered = false;
    }
    ~World() {
        if (m_Vx) vxpt_destroy(m_Vx);
    }
    World(const World&) = delete;
    World& operator=(const World&) = delete;

    const Block& GetBlock(uint16_t x, uint16_t y, uint16_t z) const { return (*m_WorldData)[x + y * WORLD_SIZE_X + z * WORLD_SIZE_X * WORLD_SIZE_Y]; }
    void SetBlock(uint16_t x, uint16_t y, uint16_t z, Block block) { (*m_WorldData)[x + y * WORLD_SIZE_X + z * WORLD_SIZE_X * WORLD_SIZE_Y] = block; }
    const Block* Data() const { return m_WorldData->data(); }  // the grid as World::Buffer uploads it (one byte per block)

    // World::Buffer (Core/World.h:167-171): upload the grid; creates the device context on first use
    bool Buffer(int device = 0) {
        if (!m_Vx && !vx_ok(vxpt_create(device, &m_Vx), "vxpt_create")) return false;
        m_Buffered = vx_ok(vxpt_upload_world(m_Vx, reinterpret_cast<const uint8_t*>(m_WorldData->data())), "vxpt_upload_world");
        return m_Buffered;
    }
    // World::InitializeDistanceGenerator (Core/World.cpp:48-67): nothing to compile here; kept for call-site parity
    bool InitializeDistanceGenerator() { return m_Vx != nullptr; }
    // World::GenerateDistanceField (Core/World.cpp:69-113)
    bool GenerateDistanceField() { return vx_ok(vxpt_build_distance_field(m_Vx), "vxpt_build_distance_field"); }
    // block edit of World::Raycast (Core/World.cpp:367-374, 456-460): host grid + 1-voxel device update + full rebuild
    bool EditBlock(uint16_t x, uint16_t y, uint16_t z, Block block) {
        SetBlock(x, y, z, block);
        if (!vx_ok(vxpt_set_block(m_Vx, x, y, z, block.block), "vxpt_set_block")) return false;
        return GenerateDistanceField();
    }
    // ---- CPU picking: World::RaycastDetect / World::Raycast (Core/World.cpp:215-546) ----
    // The block the player looks at, found on the host grid by stepping from cell face to cell face (48 steps of reach); it does not
    // read the distance field.  Without the reference's side effects that are out of scope (particles, sound, light-propagation queues).
    struct PickResult {
        bool changed;   // the reference's return value
        int x, y, z;    // edited / picked voxel, -1 when none
        uint8_t block;  // placed, removed or picked id
    };
    static bool PickOutside(const float p[3]) {
        return (int)std::floor(p[0]) >= WORLD_SIZE_X || (int)std::floor(p[1]) >= WORLD_SIZE_Y || (int)std::floor(p[2]) >= WORLD_SIZE_Z ||
               (int)std::floor(p[0]) <= 0 || (int)std::floor(p[1]) <= 0 || (int)std::floor(p[2]) <= 0;
    }
    // RaycastDetect (:497-546): (x, y, z, block) of the first solid cell; x = -1 when the hit cell is outside; false when nothing is hit
    bool RaycastDetect(const float pos[3], const float dir[3], int out[4]) const {
        float p[3] = {pos[0], pos[1], pos[2]}, sign[3];
        for (int i = 0; i < 3; ++i) sign[i] = dir[i] > 0;
        for (int i = 0; i < 48; ++i) {
            float tvec[3];
            for (int k = 0; k < 3; ++k) tvec[k] = (std::floor(p[k] + sign[k]) - p[k]) / dir[k];
            const float t = std::min(tvec[0], std::min(tvec[1], tvec[2]));
            for (int k = 0; k < 3; ++k) p[k] += dir[k] * (t + 0.001f);
            if (!PickOutside(p) && GetBlock((uint16_t)(int)p[0], (uint16_t)(int)p[1], (uint16_t)(int)p[2]).block != 0) {
                out[0] = (int)p[0]; out[1] = (int)p[1]; out[2] = (int)p[2];
                out[3] = GetBlock((uint16_t)out[0], (uint16_t)out[1], (uint16_t)out[2]).block;
                return true;
            }
        }
        return false;
    }
    // Raycast (:215-495): op 0 = break, 1 = place `held`, 2 = pick.  Edits the host grid and, when the world is buffered, the device grid
    // + distance field (glTexSubImage3D + GenerateDistanceField in the reference).
    PickResult Raycast(uint8_t op, const float pos[3], const float dir[3], uint8_t held = BlockID::Stone) {
        const PickResult none{false, -1, -1, -1, 0};
        float p[3] = {pos[0], pos[1], pos[2]}, sign[3];
        for (int i = 0; i < 3; ++i) sign[i] = dir[i] > 0;
        for (int i = 0; i < 48; ++i) {
            float tvec[3];
            for (int k = 0; k < 3; ++k) tvec[k] = (std::floor(p[k] + sign[k]) - p[k]) / dir[k];
            const float t = std::min(tvec[0], std::min(tvec[1], tvec[2]));
            for (int k = 0; k < 3; ++k) p[k] += dir[k] * (t + 0.001f);
            if (PickOutside(p) || GetBlock((uint16_t)(int)p[0], (uint16_t)(int)p[1], (uint16_t)(int)p[2]).block == 0) continue;
            float normal[3];
            for (int k = 0; k < 3; ++k) {
                normal[k] = (t == tvec[k]);
                if (sign[k]) normal[k] = -normal[k];
            }
            if (op == 1)
                for (int k = 0; k < 3; ++k) p[k] = p[k] + normal[k];
            for (int k = 0; k < 3; ++k) p[k] = std::floor(p[k]);
            if (PickOutside(p)) return none;
            const int x = (int)p[0], y = (int)p[1], z = (int)p[2];
            auto dist = [](const float a[3], float bx, float by, float bz) {
                const float dx = bx - a[0], dy = by - a[1], dz = bz - a[2];
                return std::sqrt((dx * dx + dy * dy) + dz * dz);
            };
            if (op == 1) {
                const uint8_t prev = GetBlock((uint16_t)x, (uint16_t)y, (uint16_t)z).block;
                const uint16_t ox = (uint16_t)(int)pos[0], oz = (uint16_t)(int)pos[2];
                const uint8_t u1 = GetBlock(ox, (uint16_t)(((int)pos[1]) - 1.0f), oz).block, u2 = GetBlock(ox, (uint16_t)(int)(pos[1] - 1.0f), oz).block,
                              u3 = GetBlock(ox, (uint16_t)(int)(pos[1] - 1.1f), oz).block;
                if (prev != 0 || dist(p, pos[0], pos[1], pos[2]) < 1.25f) return none;
                if ((u1 == 0 || u2 == 0 || u3 == 0) && dist(p, pos[0], pos[1] - 1.0f, pos[2]) < 1.35f) return none;
                const bool placed = prev != held;
                SetBlock((uint16_t)x, (uint16_t)y, (uint16_t)z, {held});
                if (m_Buffered && placed) EditBlock((uint16_t)x, (uint16_t)y, (uint16_t)z, {held});
                return PickResult{placed, x, y, z, held};
            }
            if (op == 0) {
                const uint8_t old = GetBlock((uint16_t)x, (uint16_t)y, (uint16_t)z).block;
                SetBlock((uint16_t)x, (uint16_t)y, (uint16_t)z, {0});
                if (m_Buffered) EditBlock((uint16_t)x, (uint16_t)y, (uint16_t)z, {0});
                return PickResult{true, x, y, z, old};
            }
            if (op == 2) return PickResult{false, x, y, z, GetBlock((uint16_t)x, (uint16_t)y, (uint16_t)z).block};
            return none;
        }
        return none;
    }
    bool DownloadDistanceField(std::vector<uint8_t>& out) const {
        out.resize(VXPT_WORLD_VOXELS);
        return vx_ok(vxpt_download_distance_field(m_Vx, out.data()), "vxpt_download_distance_field");
    }
    vxpt_handle Handle() const { return m_Vx; }

    std::unique_ptr<std::array<Block, WORLD_SIZE_X * WORLD_SIZE_Y * WORLD_SIZE_Z>> m_WorldData;
    std::string m_Name;

private:
    bool m_Buffered;
    vxpt_handle m_Vx = nullptr;
};

// SetVerticalBlocks / GenerateWorld (Core/WorldGenerator.cpp:28-67, 69-121).  gen_type false = superflat.  For the plains
// the per-column (height, biome) table comes from the reference's FastNoise (voxelpathtracer_b200/data/plains_columns.u8).
inline void SetVerticalBlocks(World* world, int x, int z, int y_level, int biome) {
    for (int y = 0; y < y_level && y < WORLD_SIZE_Y; y++) {
        uint8_t id;
        if (biome == 1) id = (y >= y_level - 1) ? BlockID::Grass : (y >= y_level - 5 ? BlockID::Dirt : BlockID::Stone);
        else id = (y >= y_level - 8) ? BlockID::Sand : BlockID::Stone;
        world->SetBlock((uint16_t)x, (uint16_t)y, (uint16_t)z, {id});
    }
}
inline void GenerateWorld(World* world, bool gen_type, const uint8_t* plains_columns = nullptr) {
    for (int x = 0; x < WORLD_SIZE_X; x++)
        for (int z = 0; z < WORLD_SIZE_Z; z++) {
            if (gen_type && plains_columns) SetVerticalBlocks(world, x, z, plains_columns[(x * WORLD_SIZE_Z + z) * 2], plains_columns[(x * WORLD_SIZE_Z + z) * 2 + 1]);
            else SetVerticalBlocks(world, x, z, 50, 1);
        }
}

// SaveWorld / LoadWorld (Core/WorldFileHandler.cpp:10-83): the raw 18,874,368-byte grid
inline bool SaveWorld(World* world, const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "wb+");
    if (!f) { Logger::Log("COULD NOT SAVE WORLD"); return false; }
    size_t n = std::fwrite(world->m_WorldData->data(), sizeof(Block), world->m_WorldData->size(), f);
    std::fclose(f);
    return n == world->m_WorldData->size();
}
inline bool LoadWorld(World* world, const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (!f) { Logger::Log("COULD NOT OPEN WORLD FILE"); return false; }
    size_t n = std::fread(world->m_WorldData->data(), sizeof(Block), world->m_WorldData->size(), f);
    std::fclose(f);
    return n == world->m_WorldData->size();
}

// BlockDataSSBO (Core/BlockDataSSBO.cpp:15-46): 6 x 128 ints, Front-face layers
class BlockDataSSBO {
public:
    bool CreateBuffers(World* world, const std::array<int32_t, 768>& total_data) { return vx_ok(vxpt_set_materials(world->Handle(), total_data.data()), "vxpt_set_materials"); }
};
// BlueNoiseDataSSBO (Core/BlueNoiseDataSSBO.cpp:17-37): sobol | scramble | rank
class BlueNoiseDataSSBO {
public:
    bool CreateBuffers(World* world, const int32_t* sobol, const int32_t* scramble, const int32_t* rank) {
        return vx_ok(vxpt_set_blue_noise(world->Handle(), sobol, scramble, rank), "vxpt_set_blue_noise");
    }
};

// ---- minimal column-major mat4 (glm::value_ptr order) -------------------------------------------------------------------
struct Mat4 {
    float m[16];
};
inline Mat4 InverseRigid(const float s[3], const float u[3], const float f[3], const float eye[3]) {
    // inverse of glm::lookAt(eye, eye + f, up): columns = right, up, -front, position
    Mat4 r{};
    for (int k = 0; k < 3; ++k) { r.m[k] = s[k]; r.m[4 + k] = u[k]; r.m[8 + k] = -f[k]; r.m[12 + k] = eye[k]; }
    r.m[15] = 1.0f;
    return r;
}

class FPSCamera {
public:
    // fov / aspect / near / far are kept in double like the Python mirror's (camera.FpsCamera), so both derive bit-identical matrices
    FPSCamera(double fov, double aspect, double zNear = 0.1, double zFar = 1000.0) : m_Fov(fov), m_Aspect(aspect), m_zNear(zNear), m_zFar(zFar) {}
    void SetPosition(float x, float y, float z) { m_Position[0] = x; m_Position[1] = y; m_Position[2] = z; }
    // FPSCamera::UpdateOnMouseMovement (Core/FpsCamera.cpp:66-70)
    void SetYawPitch(float yaw_deg, float pitch_deg) {
        const double ry = yaw_deg * M_PI / 180.0, rp = pitch_deg * M_PI / 180.0;
        m_Front[0] = (float)(std::cos(rp) * std::cos(ry));
        m_Front[1] = (float)std::sin(rp);
        m_Front[2] = (float)(std::cos(rp) * std::sin(ry));
    }
    // u_InverseView / u_InverseProjection as Pipeline.cpp:1823-1824 hands them to the shaders
    VxCamera GetVxCamera(int width, int height) const {
        double f[3] = {m_Front[0], m_Front[1], m_Front[2]};
        double fl = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        for (double& c : f) c /= fl;
        double s[3] = {f[1] * 0.0 - f[2] * 1.0, f[2] * 0.0 - f[0] * 0.0, f[0] * 1.0 - f[1] * 0.0};  // cross(f, up = (0,1,0))
        double sl = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
        for (double& c : s) c /= sl;
        double u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};  // cross(s, f)
        float sf[3] = {(float)s[0], (float)s[1], (float)s[2]}, uf[3] = {(float)u[0], (float)u[1], (float)u[2]}, ff[3] = {(float)f[0], (float)f[1], (float)f[2]};
        VxCamera cam{};
        Mat4 iv = InverseRigid(sf, uf, ff, m_Position);
        std::memcpy(cam.inv_view, iv.m, sizeof iv.m);
        // inverse of glm::perspective(radians(fov), aspect, near, far)
        const double t = std::tan(m_Fov * M_PI / 180.0 / 2.0);
        const double A = -(m_zFar + m_zNear) / (m_zFar - m_zNear), B = -(2.0 * m_zFar * m_zNear) / (m_zFar - m_zNear);
        cam.inv_proj[0] = (float)(m_Aspect * t);
        cam.inv_proj[5] = (float)t;
        cam.inv_proj[11] = (float)(1.0 / B);
        cam.inv_proj[14] = -1.0f;
        cam.inv_proj[15] = (float)(A / B);
        cam.width = width; cam.height = height; cam.row_begin = 0; cam.row_end = height;
        return cam;
    }
    // u_View / u_Projection (glm::lookAt, glm::perspective; Core/FpsCamera.cpp:23-24,150-163), column-major: what a frame hands to the NEXT
    // frame's temporal filters as u_PrevView / u_PrevProjection.  Same formulas, in the same order, as camera.FpsCamera.view_projection_f32.
    void GetViewProjection(float view[16], float proj[16]) const {
        double f[3] = {m_Front[0], m_Front[1], m_Front[2]};
        double fl = std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
        for (double& c : f) c /= fl;
        double s[3] = {f[1] * 0.0 - f[2] * 1.0, f[2] * 0.0 - f[0] * 0.0, f[0] * 1.0 - f[1] * 0.0};
        double sl = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
        for (double& c : s) c /= sl;
        double u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
        const double e[3] = {m_Position[0], m_Position[1], m_Position[2]};
        for (int k = 0; k < 16; ++k) view[k] = proj[k] = 0.0f;
        for (int c = 0; c < 3; ++c) {
            view[4 * c + 0] = (float)s[c];
            view[4 * c + 1] = (float)u[c];
            view[4 * c + 2] = (float)(-f[c]);
        }
        view[12] = (float)(-((s[0] * e[0] + s[1] * e[1]) + s[2] * e[2]));
        view[13] = (float)(-((u[0] * e[0] + u[1] * e[1]) + u[2] * e[2]));
        view[14] = (float)((f[0] * e[0] + f[1] * e[1]) + f[2] * e[2]);
        view[15] = 1.0f;
        const double fov = m_Fov, aspect = m_Aspect, zn = m_zNear, zf = m_zFar;
        const double t = std::tan(fov * M_PI / 180.0 / 2.0);
        proj[0] = (float)(1.0 / (aspect * t));
        proj[5] = (float)(1.0 / t);
        proj[10] = (float)(-(zf + zn) / (zf - zn));
        proj[11] = -1.0f;
        proj[14] = (float)(-(2.0 * zf * zn) / (zf - zn));
    }
    float m_Position[3] = {192.0f, 75.0f, 192.0f};  // Pipeline.cpp:1500
    float m_Front[3] = {0.0f, 0.0f, 1.0f};

private:
    double m_Fov, m_Aspect, m_zNear, m_zFar;
};

// TAAJitter.cpp:6-47
inline float HaltonSequence(int Prime, int index) {
    float r = 0.0f, f = 1.0f;
    int i = index;
    while (i > 0) {
        f /= Prime;
        r += f * (i % Prime);
        i = (int)std::floor(i / float(Prime));
    }
    return r;
}
inline void GetTAAJitter(int CurrentFrame, float out[2]) {
    out[0] = HaltonSequence(2, CurrentFrame % 64 + 1);
    out[1] = HaltonSequence(3, CurrentFrame % 64 + 1);
}

}  // namespace VoxelRT
