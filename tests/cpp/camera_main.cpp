// camera_main.cpp — TEST DRIVER: prints the C++ host mirror's camera matrices (bit patterns) for "x y z yaw pitch aspect" lines on stdin, so
// that tests/test_host_cpp.py can check them against the Python mirror bit for bit.  Never touches the GPU.
#include <cstdio>
#include <cstring>

#include "../../voxelpathtracer_b200/host/VoxelRT.h"

static void dump(const float* m) {
    for (int k = 0; k < 16; ++k) {
        unsigned u;
        std::memcpy(&u, &m[k], 4);
        std::printf("%08x ", u);
    }
}

int main() {
    float x, y, z, yaw, pitch;
    double aspect;
    while (std::scanf("%f %f %f %f %f %lf", &x, &y, &z, &yaw, &pitch, &aspect) == 6) {
        VoxelRT::FPSCamera cam(60.0, aspect);
        cam.SetPosition(x, y, z);
        cam.SetYawPitch(yaw, pitch);
        float view[16], proj[16];
        cam.GetViewProjection(view, proj);
        const VxCamera vc = cam.GetVxCamera(64, 36);
        dump(view); dump(proj); dump(vc.inv_view); dump(vc.inv_proj);
        std::printf("\n");
    }
    return 0;
}
