// warp_emu.h — test infrastructure: warp shuffles for the host builds of the kernels (kernels_on_host.cpp, vxpt_hostemu.cpp).
// The host builds run the "threads" of a block one after another, so a shuffle cannot see its partner's value at the time of the
// call.  VX_LAUNCH_WARPSYNC therefore runs every warp TWICE: a recording pass in which __shfl_xor_sync stores the calling lane's
// operand (call by call) and returns it unchanged, then a replay pass in which it returns what lane ^ mask recorded for the same
// call.  That is exact for kernels whose control flow up to each shuffle does not depend on shuffled values and in which all 32 lanes
// execute the same sequence of shuffles (the G-buffer pass's quad-shuffle instantiation: its shuffles come first); whatever the
// recording pass stored to memory is overwritten by the replay pass.
#pragma once
#include <cstdint>
#include <cstring>

static thread_local int vx_warp_mode = 0;  // 0 = recording, 1 = replay
static thread_local int vx_warp_call = 0;  // index of the next shuffle of the running thread
static thread_local unsigned vx_warp_lane = 0;
static thread_local uint32_t vx_warp_table[64][32];

template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
    static_assert(sizeof(T) == 4, "32-bit operands only");
    const int call = vx_warp_call++;
    if (call >= 64) std::abort();
    if (vx_warp_mode == 0) {
        std::memcpy(&vx_warp_table[call][vx_warp_lane], &v, 4);
        return v;
    }
    T r;
    std::memcpy(&r, &vx_warp_table[call][(vx_warp_lane ^ (unsigned)lane_mask) & 31u], 4);
    return r;
}

#define VX_LAUNCH_WARPSYNC(kernel, grid, block, stream, ...)                                     \
    do {                                                                                         \
        const dim3 _g = (grid);                                                                  \
        const int _nb = (int)(_g.x * _g.y);                                                      \
        const unsigned _bs = (unsigned)(block);                                                  \
        _Pragma("omp parallel for schedule(dynamic, 4)") for (int _b = 0; _b < _nb; ++_b) {      \
            blockIdx = uint3{(unsigned)_b % _g.x, (unsigned)_b / _g.x, 0u};                      \
            VX_WARP_EMU_SET_DIMS(_bs, _g);                                                       \
            for (unsigned _w = 0; _w < _bs; _w += 32) {                                          \
                for (int _pass = 0; _pass < 2; ++_pass) {                                        \
                    vx_warp_mode = _pass;                                                        \
                    for (unsigned _t = _w; _t < _w + 32 && _t < _bs; ++_t) {                     \
                        threadIdx = uint3{_t, 0u, 0u};                                           \
                        vx_warp_lane = _t & 31u;                                                 \
                        vx_warp_call = 0;                                                        \
                        kernel(__VA_ARGS__);                                                     \
                    }                                                                            \
                }                                                                                \
            }                                                                                    \
        }                                                                                        \
    } while (0)
