// ref_mc_import_driver.cpp — runs the REFERENCE'S OWN Minecraft importer (Core/NBT/Importer.cpp over its vendored enkiMI + miniz) on a
// directory of region files (test infrastructure only).  oracle/Makefile strips Importer.cpp's `#include "Importer.h"` (which drags in the
// GL block database) into oracle/_ref/Importer_body.inc; this driver supplies what that header would have: glm, the world size macros and
// BlockDatabase::GetIDFromMCID backed by a 256-byte table read from a file (the table is an INPUT on both sides of the comparison).
//   usage: ref_mc_import <region directory> <ox> <oy> <oz> <mcid_lut.u8> <out.u8>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <filesystem>
#include <string>

#include <glm/glm.hpp>

extern "C" {
#include "enkimi.h"
}

#define WORLD_SIZE_X 384
#define WORLD_SIZE_Y 128
#define WORLD_SIZE_Z 384

namespace VoxelRT {
namespace BlockDatabase {
static uint8_t g_lut[256];
inline uint8_t GetIDFromMCID(uint8_t mcid) { return g_lut[mcid]; }
}  // namespace BlockDatabase
}  // namespace VoxelRT

#include "_ref/Importer_body.inc"

int main(int argc, char** argv) {
    if (argc < 7) { std::fprintf(stderr, "usage: %s dir ox oy oz lut out\n", argv[0]); return 2; }
    FILE* f = std::fopen(argv[5], "rb");
    if (!f || std::fread(VoxelRT::BlockDatabase::g_lut, 1, 256, f) != 256) { std::fprintf(stderr, "cannot read the id table\n"); return 2; }
    std::fclose(f);
    static uint8_t out[384 * 128 * 384];
    try {
        VoxelRT::MCWorldImporter::ImportWorld(argv[1], out, glm::vec3(std::atof(argv[2]), std::atof(argv[3]), std::atof(argv[4])));
    } catch (const char* e) { std::fprintf(stderr, "importer threw: %s\n", e); return 1; }
    f = std::fopen(argv[6], "wb");
    if (!f || std::fwrite(out, 1, sizeof out, f) != sizeof out) return 1;
    std::fclose(f);
    return 0;
}
