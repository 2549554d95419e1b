// ref_worldgen_driver.cpp — runs the reference's terrain noise (its vendored FastNoise, compiled from
// /root/reference/Dependencies/fast_noise, not copied) with the call sequence of
// Core/WorldGenerator.cpp:71-107 and writes the per-column result: for every (x, z) the column height
// `int(height + 40)` handed to SetVerticalBlocks and the biome (0 = sand, 1 = grass).
// Output: 384*384 pairs of bytes, index (x * 384 + z) * 2.  Used once to create voxelpathtracer_b200/data/plains_columns.u8
// (tools/make_fixtures.py); test infrastructure only.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include "FastNoise.h"

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: %s out.u8\n", argv[0]); return 2; }
    // WorldGenerator.cpp:73-74 — rand() without srand(): glibc yields 1804289383, 846930886 -> 9383, 6886
    int biome_seed = rand() % 10000;
    int noise_seed = rand() % 6942;
    FastNoise BiomeGenerator(biome_seed);
    FastNoise NoiseGenerator(noise_seed);
    BiomeGenerator.SetNoiseType(FastNoise::Simplex);
    NoiseGenerator.SetNoiseType(FastNoise::SimplexFractal);
    NoiseGenerator.SetFrequency(0.0035);
    NoiseGenerator.SetFractalOctaves(5);
    std::vector<uint8_t> out(384 * 384 * 2);
    int hmin = 1000, hmax = -1;
    for (int x = 0; x < 384; x++)
        for (int z = 0; z < 384; z++) {
            float real_x = x, real_z = z;
            float h = NoiseGenerator.GetNoise(real_x, real_z);
            float height = ((h + 1.0f) / 2.0f) * 24.0f;
            float column_noise = BiomeGenerator.GetNoise(real_x / 2.0f, real_z / 2.0f);
            column_noise = ((column_noise + 1.0f) / 2) * 240;
            int y_level = (int)(height + 40);       // float -> int parameter conversion of SetVerticalBlocks
            int biome = column_noise < 130 ? 0 : 1;  // GetBiome, WorldGenerator.cpp:13-26
            out[(x * 384 + z) * 2 + 0] = (uint8_t)y_level;
            out[(x * 384 + z) * 2 + 1] = (uint8_t)biome;
            if (y_level < hmin) hmin = y_level;
            if (y_level > hmax) hmax = y_level;
        }
    FILE* f = std::fopen(argv[1], "wb");
    if (!f) return 1;
    std::fwrite(out.data(), 1, out.size(), f);
    std::fclose(f);
    std::printf("seeds %d %d heights %d..%d\n", biome_seed, noise_seed, hmin, hmax);
    return 0;
}
