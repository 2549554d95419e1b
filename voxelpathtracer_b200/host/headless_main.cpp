// headless_main.cpp — the reference's start-up + first frames without a window (main.cpp:66-70 -> Pipeline.cpp:1164-1294,
// 1959-2062, 2790-2852), through the C++ host mirror (VoxelRT.h) and the C ABI.  Builds a world, buffers it, generates
// the distance field, traces primary + hard sun shadow rays for a few frames, edits a block (full rebuild) and prints
// FNV-1a digests of the planes (tests/test_gpu_host_cpp.py compares them with the Python driver's).
//   usage: [VXPT_HEADLESS_FILTERS=1] [VXPT_HEADLESS_DEVICES=0,0] vxpt_headless [width height [plains_columns.u8]]
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "VoxelRT.h"

static uint64_t fnv1a(const void* p, size_t n) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char** argv) {
    using namespace VoxelRT;
    const int W = argc > 2 ? std::atoi(argv[1]) : 640, H = argc > 2 ? std::atoi(argv[2]) : 360;
    std::vector<uint8_t> columns;
    if (argc > 3) {
        FILE* f = std::fopen(argv[3], "rb");
        if (!f) { Logger::Log("cannot open the plains column table"); return 2; }
        columns.resize(WORLD_SIZE_X * WORLD_SIZE_Z * 2);
        if (std::fread(columns.data(), 1, columns.size(), f) != columns.size()) { std::fclose(f); return 2; }
        std::fclose(f);
    }
    World world;
    GenerateWorld(&world, !columns.empty(), columns.empty() ? nullptr : columns.data());
    if (!world.Buffer(0)) return 1;  // no CUDA device -> loud failure, there is no CPU path
    world.InitializeDistanceGenerator();
    if (!world.GenerateDistanceField()) return 1;
    std::vector<uint8_t> df;
    world.DownloadDistanceField(df);
    std::printf("df %016" PRIx64 "\n", fnv1a(df.data(), df.size()));

    FPSCamera camera(60.0, (double)W / (double)H);
    camera.SetYawPitch(90.0f, -20.0f);
    VxCamera cam = camera.GetVxCamera(W, H);
    std::vector<float> t((size_t)W * H), inv_t((size_t)W * H), transversal((size_t)W * H);
    std::vector<uint8_t> normal((size_t)W * H), block((size_t)W * H), shadow((size_t)W * H);
    VxGBuffer gbuf{t.data(), normal.data(), block.data(), inv_t.data(), nullptr};
    VxShadowOut sout{shadow.data(), transversal.data()};
    const float sun[3] = {-0.66896474f, 0.46841538f, 0.57735026f};  // SunTick = 50 (Pipeline.cpp:64,1653-1671)
    for (int frame = 0; frame < 3; ++frame) {
        VxPrimaryParams pp{frame == 0 ? 475 : 350, 1, {0.f, 0.f}, 0, 0};  // u_RenderDistance: 475 then 350 (Pipeline.cpp:55,4824)
        GetTAAJitter(frame, pp.jitter);
        if (!vx_ok(vxpt_trace_primary(world.Handle(), &cam, &pp, &gbuf), "vxpt_trace_primary")) return 1;
        VxShadowParams sp{{sun[0], sun[1], sun[2]}, frame, /*soft*/ 0, {0.f, 0.f}, 0};
        if (!vx_ok(vxpt_trace_shadow(world.Handle(), &cam, &gbuf, &sp, &sout), "vxpt_trace_shadow")) return 1;
        std::printf("frame %d t %016" PRIx64 " normal %016" PRIx64 " block %016" PRIx64 " shadow %016" PRIx64 "\n", frame, fnv1a(t.data(), t.size() * 4),
                    fnv1a(normal.data(), normal.size()), fnv1a(block.data(), block.size()), fnv1a(shadow.data(), shadow.size()));
    }
    // the same frame through the whole-frame entry point (Pipeline.cpp's pass sequence in one call, G-buffer resident on the device)
    {
        std::vector<float> t2((size_t)W * H), inv2((size_t)W * H), tr2((size_t)W * H);
        std::vector<uint8_t> n2((size_t)W * H), b2((size_t)W * H), s2((size_t)W * H);
        VxPrimaryParams pp{350, 1, {0.f, 0.f}, 0, 0};
        GetTAAJitter(2, pp.jitter);
        VxShadowParams sp{{sun[0], sun[1], sun[2]}, 2, 0, {0.f, 0.f}, 0};
        VxFrameParams fp{&pp, &sp, nullptr, nullptr, nullptr, nullptr};
        VxFrameOut fo{};
        fo.gbuffer = VxGBuffer{t2.data(), n2.data(), b2.data(), inv2.data(), nullptr};
        fo.shadow = VxShadowOut{s2.data(), tr2.data()};
        if (!vx_ok(vxpt_render_frame(world.Handle(), &cam, &fp, &fo), "vxpt_render_frame")) return 1;
        std::printf("render_frame t %016" PRIx64 " normal %016" PRIx64 " block %016" PRIx64 " shadow %016" PRIx64 "\n", fnv1a(t2.data(), t2.size() * 4),
                    fnv1a(n2.data(), n2.size()), fnv1a(b2.data(), b2.size()), fnv1a(s2.data(), s2.size()));
    }
    // the same frame once more, sharded over several handles from this one thread (vxpt_mg_*: scene state replicated by the same calls,
    // rows cut into slabs, host planes filled by every device's own copy stream).  VXPT_HEADLESS_DEVICES = "0,1,..." (ids may repeat:
    // "0,0,0" = three slabs on one GPU); the line must equal the render_frame line above.
    if (const char* devs = std::getenv("VXPT_HEADLESS_DEVICES")) {
        std::vector<int> ids;
        for (const char* q = devs; *q;) {
            ids.push_back(std::atoi(q));
            while (*q && *q != ',') ++q;
            if (*q == ',') ++q;
        }
        vxpt_mg_handle mg = nullptr;
        if (!vx_ok(vxpt_mg_create((int)ids.size(), ids.data(), &mg), "vxpt_mg_create")) return 1;
        bool ok = vx_ok(vxpt_mg_upload_world(mg, reinterpret_cast<const uint8_t*>(world.Data())), "vxpt_mg_upload_world") &&
                  vx_ok(vxpt_mg_build_distance_field(mg), "vxpt_mg_build_distance_field");
        std::vector<float> t3((size_t)W * H), inv3((size_t)W * H), tr3((size_t)W * H);
        std::vector<uint8_t> n3((size_t)W * H), b3((size_t)W * H), s3((size_t)W * H);
        VxPrimaryParams pp{350, 1, {0.f, 0.f}, 0, 0};
        GetTAAJitter(2, pp.jitter);
        VxShadowParams sp{{sun[0], sun[1], sun[2]}, 2, 0, {0.f, 0.f}, 0};
        VxFrameParams fp{&pp, &sp, nullptr, nullptr, nullptr, nullptr};
        VxFrameOut fo{};
        fo.gbuffer = VxGBuffer{t3.data(), n3.data(), b3.data(), inv3.data(), nullptr};
        fo.shadow = VxShadowOut{s3.data(), tr3.data()};
        ok = ok && vx_ok(vxpt_mg_render_frame(mg, &cam, &fp, &fo), "vxpt_mg_render_frame");
        if (ok)
            std::printf("mg_render_frame devices %d t %016" PRIx64 " normal %016" PRIx64 " block %016" PRIx64 " shadow %016" PRIx64 "\n", vxpt_mg_size(mg),
                        fnv1a(t3.data(), t3.size() * 4), fnv1a(n3.data(), n3.size()), fnv1a(b3.data(), b3.size()), fnv1a(s3.data(), s3.size()));
        vxpt_mg_destroy(mg);
        if (!ok) return 1;
    }
    // the sun-shadow filters on the last traced frame (Pipeline.cpp:2854-2944): first frame of a history, so the previous planes are zero and
    // the previous camera is the current one.  Reported on its own line; a failure here does not stop the run.  Opt-in (VXPT_HEADLESS_FILTERS=1).
    if (std::getenv("VXPT_HEADLESS_FILTERS")) {
        const size_t n = (size_t)W * H;
        std::vector<float> zero(n, 0.0f), st_shadow(n), st_frames(n), filtered(n);
        VxShadowTemporalIn ti{};
        ti.current = gbuf; ti.previous = gbuf;
        ti.shadow = shadow.data(); ti.transversal = transversal.data(); ti.prev_shadow = zero.data(); ti.prev_frames = zero.data();
        VxShadowTemporalParams tp{};
        camera.GetViewProjection(tp.prev_view, tp.prev_projection);
        VxShadowTemporalOut to{st_shadow.data(), st_frames.data()};
        VxShadowFilterIn fi{};
        fi.current = gbuf; fi.shadow = st_shadow.data(); fi.transversal = transversal.data(); fi.frames = st_frames.data();
        VxShadowFilterParams fp{1.0f};
        if (vx_ok(vxpt_shadow_temporal(world.Handle(), &cam, &ti, &tp, &to), "vxpt_shadow_temporal") &&
            vx_ok(vxpt_shadow_filter(world.Handle(), &cam, &fi, &fp, filtered.data()), "vxpt_shadow_filter"))
            std::printf("shadow_filters temporal %016" PRIx64 " frames %016" PRIx64 " filtered %016" PRIx64 "\n", fnv1a(st_shadow.data(), n * 4),
                        fnv1a(st_frames.data(), n * 4), fnv1a(filtered.data(), n * 4));
        else
            std::printf("shadow_filters failed\n");
    }
    // place a block in front of the camera: the ABI refuses to trace over a stale field until the rebuild
    world.EditBlock(192, 70, 200, {BlockID::Stone});
    world.DownloadDistanceField(df);
    std::printf("df_after_edit %016" PRIx64 "\n", fnv1a(df.data(), df.size()));
    VxStats st{};
    vxpt_get_stats(world.Handle(), &st);
    std::printf("rays %" PRIu64 " df_fetches %" PRIu64 " df_build_ms %.4f\n", (uint64_t)st.rays, (uint64_t)st.df_fetches, st.df_build_ms);
    return 0;
}
