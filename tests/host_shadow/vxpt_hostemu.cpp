// vxpt_hostemu.cpp — TEST INFRASTRUCTURE: the whole C ABI (voxelpathtracer_b200/csrc/api.cu) compiled by g++ against a miniature
// stand-in for the CUDA runtime, with the per-pixel kernels' own source run thread after thread (see kernels_on_host.cpp).
//
// Why: the build container has no GPU, so the argument checking, plane staging, slab arithmetic and call order of a new export — and
// the `-m gpu` tests written for it — would first execute at the end of a round.  `pytest --host-emulation` points the ctypes binding
// at this library instead of libvxpt.so and runs the GPU tests that use host planes on the CPU.  It proves nothing about the GPU
// (no shared memory, warps, streams or real device pointers here) and it is never built or loaded by the product: libvxpt.so is
// nvcc's build of the same sources and has no host path.
//
// What stands in for what:
//   cudaMalloc / cudaFree / cudaMemcpy* / cudaMemset*   malloc / free / memcpy / memset; allocations are remembered, so
//   cudaPointerGetAttributes                             can tell "device" memory (ours) from the caller's host planes
//   streams, events                                      everything runs synchronously; events carry a wall-clock stamp
//   the distance-field kernels (DPX, shared memory)      a plain three-sweep build + step-field packing (launch_df_build, launch_pack_bricks)
//   the wavefront GI pipeline (ballots, queues)          the one-thread-per-pixel diffuse kernel (bit-identical planes by construction)
//   release / acquire flag words, the global timer       plain stores / loads, a clock; a wait on a flag nobody set times out at once
#define VXPT_HOST_SHADOW 1
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

// ---- the CUDA runtime, miniature ------------------------------------------------------------------------------------------------
namespace {
std::mutex g_mu;
std::map<uintptr_t, size_t> g_allocs;  // base -> bytes
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct FakeEvent {
    double ms;
};
}  // namespace

extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaDeviceCanAccessPeer(int* can, int, int) { *can = 1; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "host emulation"; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (cudaStream_t) new int(0); return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest) { *least = 0; *greatest = 0; return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned f, int) { return cudaStreamCreateWithFlags(s, f); }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete (int*)s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamIsCapturing(cudaStream_t, cudaStreamCaptureStatus* st) { *st = cudaStreamCaptureStatusNone; return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = (cudaEvent_t) new FakeEvent{0.0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete (FakeEvent*)e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { ((FakeEvent*)e)->ms = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(((FakeEvent*)b)->ms - ((FakeEvent*)a)->ms); return cudaSuccess; }
cudaError_t cudaMalloc(void** p, size_t bytes) {
    *p = std::malloc(bytes ? bytes : 1);
    if (!*p) return cudaErrorMemoryAllocation;
    std::lock_guard<std::mutex> lk(g_mu);
    g_allocs[(uintptr_t)*p] = bytes;
    return cudaSuccess;
}
cudaError_t cudaFree(void* p) {
    if (!p) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        g_allocs.erase((uintptr_t)p);
    }
    std::free(p);
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { std::memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    std::memset(a, 0, sizeof *a);
    a->type = cudaMemoryTypeUnregistered;
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_allocs.upper_bound((uintptr_t)p);
    if (it != g_allocs.begin()) {
        --it;
        if ((uintptr_t)p < it->first + std::max<size_t>(it->second, 1)) a->type = cudaMemoryTypeDevice;
    }
    return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p) { std::memset(h, 0, sizeof *h); std::memcpy(h, &p, sizeof p); return cudaSuccess; }
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned) { std::memcpy(p, &h, sizeof *p); return cudaSuccess; }  // same process
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaSuccess; }
}  // extern "C"

// ---- what nvcc provides implicitly (as in kernels_on_host.cpp) ----------------------------------------------------------------------
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __grid_constant__
#define __grid_constant__
#undef __shared__
#define __shared__ static
static thread_local uint3 threadIdx, blockIdx;
static thread_local dim3 blockDim, gridDim;
using std::max;
using std::min;
static inline float __fadd_rd(float a, float b) {
    const double s = (double)a + (double)b;
    float f = (float)s;
    if ((double)f > s) f = std::nextafterf(f, -INFINITY);
    return f;
}
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline long long __double_as_longlong(double d) { long long i; std::memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; std::memcpy(&d, &i, 8); return d; }
static inline unsigned __float2uint_rn(float f) { return (unsigned)std::nearbyintf(f); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_RELAXED); }
static inline void __syncthreads() {}
static inline void __nanosleep(unsigned) {}
namespace vxpt {
static inline void store_release_sys(uint32_t* flag, uint32_t value) { __atomic_store_n(flag, value, __ATOMIC_RELEASE); }
static inline uint32_t load_acquire_sys(const uint32_t* f) { return __atomic_load_n(f, __ATOMIC_ACQUIRE); }
// nothing runs concurrently here: a flag that is not there yet never will be, so every wait has already timed out
static inline unsigned long long global_timer_ns() { static unsigned long long t = 0; return t += (1ull << 40); }
}  // namespace vxpt

#define VX_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) return ::vxpt::cuda_fail(_e, #expr, __FILE__, __LINE__); \
    } while (0)
// kernels whose blocks are independent run their blocks in parallel; the single-block kernels of api.cu (flags, scatter) run in order
#define VX_LAUNCH(kernel, grid, block, stream, ...)                                          \
    do {                                                                                     \
        const dim3 _g = (grid);                                                              \
        const int _nb = (int)(_g.x * _g.y);                                                  \
        const unsigned _bs = (unsigned)(block);                                              \
        _Pragma("omp parallel for schedule(dynamic, 4)") for (int _b = 0; _b < _nb; ++_b) {  \
            blockIdx = uint3{(unsigned)_b % _g.x, (unsigned)_b / _g.x, 0u};                  \
            blockDim = dim3(_bs, 1, 1);                                                      \
            gridDim = _g;                                                                    \
            for (unsigned _t = 0; _t < _bs; ++_t) {                                          \
                threadIdx = uint3{_t, 0u, 0u};                                               \
                kernel(__VA_ARGS__);                                                         \
            }                                                                                \
        }                                                                                    \
    } while (0)

#define VX_WARP_EMU_SET_DIMS(bs, g) do { blockDim = dim3((bs), 1, 1); gridDim = (g); } while (0)
#include "warp_emu.h"

#include "../../voxelpathtracer_b200/csrc/trace.cu"
#include "../../voxelpathtracer_b200/csrc/trace_reflection.cu"
#include "../../voxelpathtracer_b200/csrc/df_consumers.cu"
#include "../../voxelpathtracer_b200/csrc/gbuffer.cu"
#include "../../voxelpathtracer_b200/csrc/denoise.cu"
#include "../../voxelpathtracer_b200/csrc/api.cu"
#include "../../voxelpathtracer_b200/csrc/mg.cu"

// ---- the launchers of the translation units that need a GPU (df_build.cu, trace_gi.cu, l2_probe.cu) -----------------------------------
namespace vxpt {
int init_df_kernels(vxpt_ctx*) { return VXPT_OK; }
// ManhattanDistance{X,Y,Z}.comp as three plain sweeps (the oracle's definition), on the "device" buffers
int launch_df_build(vxpt_ctx* c) {
    const uint8_t* g = c->d_grid;
    uint8_t* d = c->d_df;
#pragma omp parallel for
    for (int line = 0; line < WY * WZ; ++line) {
        const size_t base = (size_t)line * WX;
        int v = 254;
        for (int x = 0; x < WX; ++x) { v = g[base + x] ? 0 : std::min(254, v + 1); d[base + x] = (uint8_t)v; }
        for (int x = WX - 2; x >= 0; --x) if (d[base + x + 1] < d[base + x]) d[base + x] = (uint8_t)(1 + d[base + x + 1]);
    }
#pragma omp parallel for
    for (int line = 0; line < WX * WZ; ++line) {
        const int x = line % WX, z = line / WX;
        auto at = [&](int y) -> uint8_t& { return d[(size_t)x + (size_t)WX * ((size_t)y + (size_t)WY * z)]; };
        for (int y = 1; y < WY; ++y) at(y) = (uint8_t)std::min<int>(at(y), at(y - 1) + 1);
        for (int y = WY - 2; y >= 0; --y) at(y) = (uint8_t)std::min<int>(at(y), at(y + 1) + 1);
    }
#pragma omp parallel for
    for (int line = 0; line < WX * WY; ++line) {
        const int x = line % WX, y = line / WX;
        auto at = [&](int z) -> uint8_t& { return d[(size_t)x + (size_t)WX * ((size_t)y + (size_t)WY * z)]; };
        for (int z = 1; z < WZ; ++z) at(z) = (uint8_t)std::min<int>(at(z), at(z - 1) + 1);
        for (int z = WZ - 2; z >= 0; --z) at(z) = (uint8_t)std::min<int>(at(z), at(z + 1) + 1);
    }
    c->steps_layout = -1;  // the step field follows in launch_pack_bricks (vxpt_build_distance_field)
    return VXPT_OK;
}
// pack_steps: E(M) = (M == 1) ? 1 : floor(M * 0.57735026918f), linear or 8x4x4 bricks
int launch_pack_bricks(vxpt_ctx* c) {
#pragma omp parallel for
    for (int z = 0; z < WZ; ++z)
        for (int y = 0; y < WY; ++y)
            for (int x = 0; x < WX; ++x) {
                const size_t lin = (size_t)x + (size_t)WX * ((size_t)y + (size_t)WY * z);
                const float m = (float)c->d_df[lin];
                const uint8_t e = (uint8_t)(int)std::floor(m == 1.0f ? 1.0f : m * 0.57735026918f);
                c->d_steps[c->opt_layout == 1 ? brick_offset(x, y, z) : lin] = e;
            }
    c->steps_layout = c->opt_layout;
    return VXPT_OK;
}
// the wavefront pipeline produces the planes of the one-thread-per-pixel kernel bit for bit (which thread traces a ray does not change
// what is computed for its pixel): the tail of trace.cu's launch_diffuse
int launch_diffuse_wavefront(vxpt_ctx* c, const VxCamera& cam, const DiffuseDev& d, const VxGBuffer& g, const VxDiffuseOut& out) {
    const SceneDev S = make_scene(c);
    const DiffuseOutDev od{reinterpret_cast<float4*>(out.sh), reinterpret_cast<float2*>(out.cocg), out.luma, reinterpret_cast<float2*>(out.ao_sky), c->opt_texel};
    const dim3 grid = pixel_grid(cam);
    if (c->opt_layout == 1) VX_LAUNCH((diffuse_kernel<1>), grid, 256, c->stream, S, to_dev(cam), d, to_dev(c, g), od);
    else VX_LAUNCH((diffuse_kernel<0>), grid, 256, c->stream, S, to_dev(cam), d, to_dev(c, g), od);
    c->launches += 1;
    return VXPT_OK;
}
size_t gi_scratch_bytes(size_t, size_t, bool) { return 256; }
int run_l2_probe(vxpt_ctx*, double* gbps) { *gbps = 1000.0; return VXPT_OK; }
}  // namespace vxpt
