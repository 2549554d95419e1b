// picking_main.cpp — TEST DRIVER for the C++ host mirror's CPU picking (VoxelRT::World::Raycast / RaycastDetect): reads
// "op px py pz dx dy dz held" lines from stdin on a superflat world (op 3 = RaycastDetect) and prints one result line per query, so that
// tests/test_host.py can compare the C++ mirror with the Python mirror.  Never touches the GPU (the world is not buffered).
#include <cstdio>

#include "../../voxelpathtracer_b200/host/VoxelRT.h"

int main() {
    VoxelRT::World world;
    VoxelRT::GenerateWorld(&world, false);
    int op, held;
    float p[3], d[3];
    while (std::scanf("%d %f %f %f %f %f %f %d", &op, &p[0], &p[1], &p[2], &d[0], &d[1], &d[2], &held) == 8) {
        if (op == 3) {
            int out[4] = {-1, -1, -1, -1};
            const bool hit = world.RaycastDetect(p, d, out);
            std::printf("%d %d %d %d %d\n", hit ? 1 : 0, out[0], out[1], out[2], out[3]);
        } else {
            const VoxelRT::World::PickResult r = world.Raycast((uint8_t)op, p, d, (uint8_t)held);
            std::printf("%d %d %d %d %d\n", r.changed ? 1 : 0, r.x, r.y, r.z, (int)r.block);
        }
    }
    return 0;
}
