// df_build.cu — Manhattan (L1) distance field over the 384x128x384 block grid, sm_100a.
//
// Replaces World::GenerateDistanceField (Core/World.cpp:69-113) and the three serial-scan compute shaders
// Core/Shaders/ManhattanDistance{X,Y,Z}.comp.  Result:
//     DF[p] = min(254, min over solid q of |p - q|_1),   solid = block byte > 0.
// The separable min-plus sweeps commute, so any order of the three axes yields the same bytes.
//
// Design (algo 1, default) — two kernels, both built on the DPX instruction VIADDMNMX.U16x2
// (__viaddmin_u16x2: per 16-bit lane min(a + b, c)), i.e. one instruction per sweep step per TWO voxels:
//   df_xy_dpx : persistent CTAs (two per SM) stream z-slices (49,152 contiguous bytes) through two shared-memory buffers each: one TMA
//               bulk copy (cp.async.bulk + mbarrier) brings a slice in while the previous one is swept; x sweep with rows paired in the
//               two 16-bit lanes (24 voxels x 2 rows per lane, cross-lane carries by shuffle min-plus scans over half warps), y sweep
//               with the column streamed through shared memory (4 voxels per thread); one TMA bulk store writes the slice back.
//   df_z_dpx  : the z sweep, launched as a programmatic dependent of df_xy_dpx; each thread owns a 24-voxel z-segment of one 4-byte
//               x-word in registers, 16 segments per CTA exchange their edge values through shared memory (min-plus carries).  The
//               finished words also leave as the traversal's step field E(M) (trace_device.cuh), converted two voxels per table read
//               and stored in the 8x4x4-brick layout (16-byte stores) — no second pass over the distance field.
//   pack_steps: distance field -> step field on its own (after the reference-shaped build, or when the layout option changes).
// Algo 0 keeps the reference's shape (one thread per grid line, three launches) as an on-device cross-check.
#include <algorithm>
#include <cstdlib>

#include "vxpt_internal.h"

namespace vxpt {

// ------------------------------------------------------------------------------------------------------------
// algo 0: reference-shaped, one thread per line (ManhattanDistanceX.comp:45-69, Y.comp:26-51, Z.comp:24-47)
// ------------------------------------------------------------------------------------------------------------
__global__ void df_x_lines(const uint8_t* __restrict__ grid, uint8_t* __restrict__ df) {
    int line = blockIdx.x * blockDim.x + threadIdx.x;  // y + WY * z
    if (line >= WY * WZ) return;
    size_t base = (size_t)line * WX;
    int prev = grid[base] > 0 ? 0 : 254;
    df[base] = (uint8_t)prev;
    for (int x = 1; x < WX; ++x) {
        prev = grid[base + x] > 0 ? 0 : min(254, prev + 1);
        df[base + x] = (uint8_t)prev;
    }
    for (int x = WX - 2; x >= 0; --x) {
        int cur = df[base + x];
        if (prev < cur) { cur = prev + 1; df[base + x] = (uint8_t)cur; }
        prev = cur;
    }
}
__global__ void df_y_lines(uint8_t* __restrict__ df) {
    int line = blockIdx.x * blockDim.x + threadIdx.x;  // x + WX * z
    if (line >= WX * WZ) return;
    int x = line % WX, z = line / WX;
    size_t base = (size_t)x + (size_t)z * WX * WY;
    int prev = df[base];
    for (int y = 1; y < WY; ++y) {
        int cur = df[base + (size_t)y * WX];
        if (prev < cur) { cur = prev + 1; df[base + (size_t)y * WX] = (uint8_t)cur; }
        prev = cur;
    }
    for (int y = WY - 2; y >= 0; --y) {
        int cur = df[base + (size_t)y * WX];
        if (prev < cur) { cur = prev + 1; df[base + (size_t)y * WX] = (uint8_t)cur; }
        prev = cur;
    }
}
__global__ void df_z_lines(uint8_t* __restrict__ df) {
    int line = blockIdx.x * blockDim.x + threadIdx.x;  // x + WX * y
    if (line >= WX * WY) return;
    size_t base = (size_t)line;
    const size_t sz = (size_t)WX * WY;
    int prev = df[base];
    for (int z = 1; z < WZ; ++z) {
        int cur = df[base + z * sz];
        if (prev < cur) { cur = prev + 1; df[base + z * sz] = (uint8_t)cur; }
        prev = cur;
    }
    for (int z = WZ - 2; z >= 0; --z) {
        int cur = df[base + z * sz];
        if (prev < cur) { cur = prev + 1; df[base + z * sz] = (uint8_t)cur; }
        prev = cur;
    }
}

// ------------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copies (SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Development aid (build.py -DVXPT_DF_TRACE --out=...; tools/df_timeline.py): thread 0 of every CTA of the two sweep kernels logs
// %globaltimer at its phase boundaries.  Not compiled into the product library.
#ifdef VXPT_DF_TRACE
constexpr int DF_TRACE_SLOTS = 16;
__device__ unsigned long long g_df_trace[2][2048][DF_TRACE_SLOTS];
__device__ __forceinline__ void df_trace(int kernel, int cta, int slot) {
    if (threadIdx.x == 0 && threadIdx.y == 0 && cta < 2048 && slot < DF_TRACE_SLOTS) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_df_trace[kernel][cta][slot] = t;
    }
}
#define DF_TRACE(kernel, cta, slot) df_trace(kernel, cta, slot)
#else
#define DF_TRACE(kernel, cta, slot)
#endif

constexpr uint32_t ONE2 = 0x00010001u;  // +1 in both 16-bit lanes
constexpr uint32_t INF2 = 0x00FE00FEu;  // 254 in both lanes ("no solid voxel seen")

// ------------------------------------------------------------------------------------------------------------
// df_xy_dpx: x and y sweeps of whole z-slices in shared memory.  Persistent CTAs (two per SM, 2 x 48 KB of shared memory each):
// CTA b takes slices b, b + gridDim.x, ...; while it sweeps one slice the TMA bulk load of its next slice is already in flight
// into the other buffer, and the bulk store of the previous slice drains behind it (r01h: one CTA per slice, load -> sweep -> store
// in lock-step on every SM, the SMs idle 30 % of the kernel).
// ------------------------------------------------------------------------------------------------------------
constexpr int XY_THREADS = 256;  // 8 warps
constexpr int XSEG = 24;         // voxels of a row one lane holds in the x sweep: 16 lanes cover a row, a warp sweeps two row pairs at once

// block bytes -> initial distances (solid 0, air 254; ManhattanDistanceX.comp:51-52) for rows A and B, two voxels of the same x in the
// halves of a register (row A low, row B high); raw = the 24 block bytes of row A (6 words) followed by those of row B
__device__ __forceinline__ void load_raw_pair24(const uint32_t* rowA, uint32_t (&raw)[12]) {
    const uint32_t* rowB = rowA + WX / 4;
#pragma unroll
    for (int k = 0; k < XSEG / 8; ++k) {
        const uint2 A = *reinterpret_cast<const uint2*>(rowA + 2 * k), B = *reinterpret_cast<const uint2*>(rowB + 2 * k);
        raw[2 * k] = A.x; raw[2 * k + 1] = A.y; raw[6 + 2 * k] = B.x; raw[6 + 2 * k + 1] = B.y;
    }
}
__device__ __forceinline__ void convert_row_pair24(const uint32_t (&raw)[12], uint32_t (&r)[XSEG]) {
#pragma unroll
    for (int k = 0; k < XSEG / 8; ++k) {
        const uint32_t a[2] = {raw[2 * k], raw[2 * k + 1]}, b[2] = {raw[6 + 2 * k], raw[6 + 2 * k + 1]};
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t t0 = __byte_perm(a[h], b[h], 0x6420);  // a0 a2 b0 b2
            const uint32_t t1 = __byte_perm(a[h], b[h], 0x7531);  // a1 a3 b1 b3
            const uint32_t v[4] = {t0 & 0x00FF00FFu, t1 & 0x00FF00FFu, (t0 >> 8) & 0x00FF00FFu, (t1 >> 8) & 0x00FF00FFu};
#pragma unroll
            for (int q = 0; q < 4; ++q) r[8 * k + 4 * h + q] = INF2 - __vminu2(v[q], ONE2) * 0xFEu;
        }
    }
}
// one word of the distance field for every word of two rows (what a uniform row pair turns into: 254 in the air, 0 in the ground)
__device__ __forceinline__ void fill_row_pair24(uint32_t* rowA, uint32_t word) {
    uint32_t* rowB = rowA + WX / 4;
#pragma unroll
    for (int k = 0; k < XSEG / 8; ++k) {
        *reinterpret_cast<uint2*>(rowA + 2 * k) = make_uint2(word, word);
        *reinterpret_cast<uint2*>(rowB + 2 * k) = make_uint2(word, word);
    }
}
__device__ __forceinline__ void store_row_pair24(uint32_t* rowA, const uint32_t (&r)[XSEG]) {
    uint32_t* rowB = rowA + WX / 4;
#pragma unroll
    for (int k = 0; k < XSEG / 8; ++k) {
        uint32_t a[2], b[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const uint32_t* q = r + 8 * k + 4 * h;
            const uint32_t t0 = q[0] | (q[2] << 8);  // a0 a2 b0 b2
            const uint32_t t1 = q[1] | (q[3] << 8);  // a1 a3 b1 b3
            a[h] = __byte_perm(t0, t1, 0x5140);
            b[h] = __byte_perm(t0, t1, 0x7362);
        }
        *reinterpret_cast<uint2*>(rowA + 2 * k) = make_uint2(a[0], a[1]);
        *reinterpret_cast<uint2*>(rowB + 2 * k) = make_uint2(b[0], b[1]);
    }
}
// x sweep of a row pair spread over 16 lanes (l16 = lane within the half warp): local forward and backward sweeps, min-plus scans of the
// segments' edge values over the half warp, then both carries applied in one pass
__device__ __forceinline__ void x_sweep_rows24(uint32_t (&r)[XSEG], int l16) {
#pragma unroll
    for (int k = 1; k < XSEG; ++k) r[k] = __viaddmin_u16x2(r[k - 1], ONE2, r[k]);
#pragma unroll
    for (int k = XSEG - 2; k >= 0; --k) r[k] = __viaddmin_u16x2(r[k + 1], ONE2, r[k]);
    uint32_t cf = r[XSEG - 1], cb = r[0];  // best value at the last / first voxel of the segment from sources inside it
#pragma unroll
    for (int d = 1; d < 16; d <<= 1) {
        const uint32_t tf = __shfl_up_sync(0xffffffffu, cf, d, 16), tb = __shfl_down_sync(0xffffffffu, cb, d, 16);
        if (l16 >= d) cf = __viaddmin_u16x2(tf, (uint32_t)(XSEG * d) * ONE2, cf);
        if (l16 + d < 16) cb = __viaddmin_u16x2(tb, (uint32_t)(XSEG * d) * ONE2, cb);
    }
    uint32_t cin_f = __shfl_up_sync(0xffffffffu, cf, 1, 16), cin_b = __shfl_down_sync(0xffffffffu, cb, 1, 16);
    if (l16 == 0) cin_f = INF2;
    if (l16 == 15) cin_b = INF2;
    // After the local sweeps a segment is 1-Lipschitz, so a carry from the left improves something in it only if it improves its FIRST
    // voxel (cin + 1 + k < r[k] <= r[0] + k), one from the right only if it improves its LAST: two DPX steps decide for the segment,
    // one vote for the warp, and rows whose segments do not see each other (the open air, the ground) skip 2 x 24 steps per lane.
    const bool need_f = __viaddmin_u16x2(cin_f, ONE2, r[0]) != r[0];
    const bool need_b = __viaddmin_u16x2(cin_b, ONE2, r[XSEG - 1]) != r[XSEG - 1];
    if (__any_sync(0xffffffffu, need_f)) {
#pragma unroll
        for (int k = 0; k < XSEG; ++k) r[k] = __viaddmin_u16x2(cin_f, (uint32_t)(k + 1) * ONE2, r[k]);
    }
    if (__any_sync(0xffffffffu, need_b)) {
#pragma unroll
        for (int k = 0; k < XSEG; ++k) r[k] = __viaddmin_u16x2(cin_b, (uint32_t)(XSEG - k) * ONE2, r[k]);
    }
}

// NBUF = 2 (default): persistent CTAs (two per SM), the next slice's load in flight while the current one is swept.  NBUF = 1: one CTA per
// slice, 48 KB each, four per SM, every load issued at the start.  Timeline r02s: with all 18.9 MB requested at once the slices arrive after
// 2.5 us (median) instead of 1.1, and three CTAs per SM run their y sweeps on the same three schedulers (warps 0-2 of each): 33.9 us per
// rebuild against 31.0 us persistent.
// One row of the y sweep on an x-word (4 voxels), two DPX chains.
// A chain only needs the HIGH byte of each 16-bit lane to be right: min((v << 8 | junk) + 0x0100, (b << 8 | junk')) has min(v + 1, b) in its high
// byte whatever the low bytes hold (a lexicographic minimum's first component is the minimum of the first components; v <= 254, so the add
// never leaves the lane).  The odd bytes of the word ARE the high bytes of its lanes, the even bytes get there by a shift (IMAD.SHL, FMA
// pipe), one PRMT collects the four high bytes: three ALU-pipe instructions per row where unpacking byte pairs into clean lanes took five
// (PRMT, PRMT, DPX, DPX, PRMT: -DVXPT_DF_Y_CLEAN_LANES).  The sweep is bound by that pipe (one instruction per two cycles per scheduler).
#ifdef VXPT_DF_Y_CLEAN_LANES
constexpr uint32_t Y_CHAIN_INIT = INF2;
#define DF_Y_STEP(lo, hi, word)                                                   \
    do {                                                                          \
        lo = __viaddmin_u16x2(lo, ONE2, __byte_perm((word), 0u, 0x4140));         \
        hi = __viaddmin_u16x2(hi, ONE2, __byte_perm((word), 0u, 0x4342));         \
        word = __byte_perm(lo, hi, 0x6420);                                       \
    } while (0)
#else
constexpr uint32_t Y_CHAIN_INIT = 0xFE00FE00u;
#define DF_Y_STEP(ce, co, word)                                                   \
    do {                                                                          \
        co = __viaddmin_u16x2(co, 0x01000100u, (word));                           \
        ce = __viaddmin_u16x2(ce, 0x01000100u, (word) << 8);                      \
        word = __byte_perm(ce, co, 0x7351);                                       \
    } while (0)
#endif
template <int NBUF>
__global__ void __launch_bounds__(XY_THREADS, NBUF == 1 ? 4 : 2) df_xy_dpx(const uint8_t* __restrict__ grid, uint8_t* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + NBUF * SLICE_BYTES);  // one mbarrier per buffer
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    DF_TRACE(0, blockIdx.x, 0);
    // the dependent z sweep may be scheduled as soon as SMs free up; it waits for this grid's memory with cudaGridDependencySynchronize()
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (tid == 0) {
        mbar_init(bar, 1);
        if (NBUF == 2) mbar_init(bar + 1, 1);
        fence_mbar_init();
    }
    __syncthreads();
    int z = blockIdx.x;
    if (tid == 0 && z < WZ) {
        mbar_arrive_expect_tx(bar, SLICE_BYTES);
        bulk_g2s(smem, grid + (size_t)z * SLICE_BYTES, SLICE_BYTES, bar);
    }
    for (int it = 0; z < WZ; ++it, z += gridDim.x) {
        const int b = NBUF == 2 ? (it & 1) : 0;
        uint8_t* tile = smem + b * SLICE_BYTES;  // [128][384] bytes
        if (NBUF == 1 && it > 0 && tid == 0) {  // the one buffer again: its store must have read it before the next slice lands
            bulk_wait_read0();
            mbar_arrive_expect_tx(bar, SLICE_BYTES);
            bulk_g2s(smem, grid + (size_t)z * SLICE_BYTES, SLICE_BYTES, bar);
        }
        mbar_wait(bar + b, NBUF == 2 ? ((it >> 1) & 1) : (it & 1));
        DF_TRACE(0, blockIdx.x, 1 + 4 * it);

        // ---- x sweep: half warp h of warp w takes row pairs (2p, 2p+1), p = 2w + h + 16j; lane l16 holds x = 24 l16 .. 24 l16 + 23
        uint32_t* t32 = reinterpret_cast<uint32_t*>(tile);
        const int l16 = lane & 15;
#pragma unroll 1
        for (int p = 2 * warp + (lane >> 4); p < WY / 2; p += 2 * (XY_THREADS / 32)) {
            uint32_t* rowA = t32 + (2 * p) * (WX / 4) + l16 * (XSEG / 4);
            uint32_t raw[12];
            load_raw_pair24(rowA, raw);
            // the four rows of this warp step are often all air or all ground: then the x sweep changes nothing but the encoding
            uint32_t any = 0u, zero_byte = 0u;
#pragma unroll
            for (int k = 0; k < 12; ++k) {
                any |= raw[k];
                zero_byte |= (raw[k] - 0x01010101u) & ~raw[k] & 0x80808080u;  // a byte of raw[k] is 0
            }
            if (__all_sync(0xffffffffu, any == 0u)) {
                fill_row_pair24(rowA, 0xFEFEFEFEu);
                continue;
            }
            if (__all_sync(0xffffffffu, zero_byte == 0u)) {
                fill_row_pair24(rowA, 0u);
                continue;
            }
            uint32_t r[XSEG];
            convert_row_pair24(raw, r);
            x_sweep_rows24(r, l16);
            store_row_pair24(rowA, r);
        }
        __syncthreads();
        DF_TRACE(0, blockIdx.x, 2 + 4 * it);
        // the other buffer is free once the bulk store of the previous slice has read it: start the load of this CTA's next slice
        if (NBUF == 2 && tid == 0 && z + (int)gridDim.x < WZ) {
            bulk_wait_read0();
            mbar_arrive_expect_tx(bar + (b ^ 1), SLICE_BYTES);
            bulk_g2s(smem + (b ^ 1) * SLICE_BYTES, grid + (size_t)(z + gridDim.x) * SLICE_BYTES, SLICE_BYTES, bar + (b ^ 1));
        }

        // ---- y sweep: thread t < 96 owns the x-word t (4 voxels = two DPX chains) and streams the column through shared memory, 16 rows
        // at a time: the loads of a batch are issued together and the next batch's before this one's chain (a compiler barrier keeps them
        // there).  Written row by row, every load waits behind the previous row's store to the same array and each step costs a
        // shared-memory round trip: 27 cycles per row, 3.5 us per slice (timeline r02r) against 4.5 cycles of dependent DPX latency.
        if (tid < WX / 4) {
            constexpr int YB = 16, STR = WX / 4;
            uint32_t* col = t32 + tid;
            uint32_t lo = Y_CHAIN_INIT, hi = Y_CHAIN_INIT, w[YB], nxt[YB];
#pragma unroll
            for (int k = 0; k < YB; ++k) w[k] = col[k * STR];
#pragma unroll 1
            for (int yb = 0; yb < WY; yb += YB) {   // forward
                if (yb + YB < WY) {
#pragma unroll
                    for (int k = 0; k < YB; ++k) nxt[k] = col[(yb + YB + k) * STR];
                }
                asm volatile("" ::: "memory");
#pragma unroll
                for (int k = 0; k < YB; ++k) {
                    DF_Y_STEP(lo, hi, w[k]);
                }
#pragma unroll
                for (int k = 0; k < YB; ++k) col[(yb + k) * STR] = w[k];
#pragma unroll
                for (int k = 0; k < YB; ++k) w[k] = nxt[k];
            }
            // backward: the last batch is still in w[] as stored (the first step re-reads the running value: min(v, v + 1) = v)
#pragma unroll
            for (int k = 0; k < YB; ++k) w[k] = col[(WY - 1 - k) * STR];
#pragma unroll 1
            for (int yb = WY - 1; yb >= 0; yb -= YB) {
                if (yb - YB >= 0) {
#pragma unroll
                    for (int k = 0; k < YB; ++k) nxt[k] = col[(yb - YB - k) * STR];
                }
                asm volatile("" ::: "memory");
#pragma unroll
                for (int k = 0; k < YB; ++k) {
                    DF_Y_STEP(lo, hi, w[k]);
                }
#pragma unroll
                for (int k = 0; k < YB; ++k) col[(yb - k) * STR] = w[k];
#pragma unroll
                for (int k = 0; k < YB; ++k) w[k] = nxt[k];
            }
        }
        fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the bulk-copy (async) proxy
        __syncthreads();
        DF_TRACE(0, blockIdx.x, 3 + 4 * it);
        if (tid == 0) {
            bulk_s2g(out + (size_t)z * SLICE_BYTES, tile, SLICE_BYTES);
            bulk_commit();
        }
        DF_TRACE(0, blockIdx.x, 4 + 4 * it);
    }
    if (tid == 0) bulk_wait_read0();  // shared memory must outlive the last store's reads
    DF_TRACE(0, blockIdx.x, 15);
}

// ------------------------------------------------------------------------------------------------------------
// df_z_dpx: z sweep + the traversal's step field
// ------------------------------------------------------------------------------------------------------------
constexpr int ZSEG = 24;          // voxels per thread along z
constexpr int ZSEGS = WZ / ZSEG;  // 16 segments per CTA
// x-words (4 bytes) per segment row of a CTA.  8 (= one 32-byte sector per row, 128-thread CTAs, 1536 of them) instead of a full warp
// (128-byte lines, 512-thread CTAs, 384 of them): with 63 registers only two 512-thread CTAs fit an SM, so 384 CTAs ran as 1.3 waves
// on 296 slots and the SMs idled 41 % of the kernel (ncu r01g); small CTAs keep eight resident per SM and refill as they retire.
constexpr int ZXW = 8;

__constant__ uint8_t c_step_lut[256];
// Step values of two x-neighbours at once.  The finished field is 1-Lipschitz in the L1 metric, so x-neighbours (M0, M1) differ by at most
// one and 2 * M0 + M1 + 1 = 3 * M0 + (M1 - M0 + 1) indexes the pair uniquely: one IDP.2A + one 16-bit shared-memory read per two voxels
// instead of an 8-bit table read (and its index extraction) per voxel.  Entry = E(M0) | E(M1) << 8.
constexpr int PAIR_LUT = 768;
__constant__ uint16_t c_pair_lut[PAIR_LUT];

__device__ __forceinline__ uint32_t lut4(const uint8_t* lut, uint32_t w) {
    return (uint32_t)lut[w & 0xFF] | ((uint32_t)lut[(w >> 8) & 0xFF] << 8) | ((uint32_t)lut[(w >> 16) & 0xFF] << 16) |
           ((uint32_t)lut[w >> 24] << 24);
}
// lo = M0 | M1 << 16, hi = M2 | M3 << 16 -> E0 | E1 << 8 | E2 << 16 | E3 << 24
__device__ __forceinline__ uint32_t steps4(const uint16_t* plut, uint32_t lo, uint32_t hi) {
    const uint32_t e01 = plut[__dp2a_lo(lo, 0x0102u, 1u)], e23 = plut[__dp2a_lo(hi, 0x0102u, 1u)];
    return __byte_perm(e01, e23, 0x5410);
}

// PACK: 0 / 1 = the step field E(M) in the linear / brick layout, written from the registers that hold the finished words (no second
// pass over the distance field).  In the brick layout the four z-neighbours of an x-word are 16 contiguous bytes (brick_offset: z & 3
// has stride 4), so a thread issues one 16-byte store per four z values; ZSEG and the segment origins are multiples of 4.
template <int PACK>
__global__ void __launch_bounds__(ZXW * ZSEGS, 8) df_z_dpx(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint8_t* __restrict__ steps) {
    __shared__ uint16_t plut[PAIR_LUT];
    {
        const int t = threadIdx.y * ZXW + threadIdx.x;
#pragma unroll
        for (int k = 0; k < PAIR_LUT / (ZXW * ZSEGS); ++k) plut[t + k * ZXW * ZSEGS] = c_pair_lut[t + k * ZXW * ZSEGS];
    }
    __shared__ uint2 edge_first[ZSEGS][ZXW];  // local value at the first voxel of a segment (lo pair, hi pair)
    __shared__ uint2 edge_last[ZSEGS][ZXW];
#ifdef VXPT_DF_SERIAL_CARRIES
    __shared__ uint2 carry_f[ZSEGS][ZXW];
    __shared__ uint2 carry_b[ZSEGS][ZXW];
#endif
    const int lane = threadIdx.x, seg = threadIdx.y;
    const int y = blockIdx.y;
    const size_t base = (size_t)y * WX + (size_t)(blockIdx.x * ZXW + lane) * 4 + (size_t)seg * ZSEG * SLICE_BYTES;
    DF_TRACE(1, blockIdx.y * gridDim.x + blockIdx.x, 0);
    // programmatic dependent launch: everything above overlapped the tail of df_xy_dpx; its stores are visible from here on
    asm volatile("griddepcontrol.wait;" ::: "memory");
    DF_TRACE(1, blockIdx.y * gridDim.x + blockIdx.x, 1);

    uint32_t lo[ZSEG], hi[ZSEG];
#pragma unroll
    for (int i = 0; i < ZSEG; ++i) {
        uint32_t w = __ldcg(reinterpret_cast<const uint32_t*>(in + base + (size_t)i * SLICE_BYTES));
        lo[i] = __byte_perm(w, 0u, 0x4140);
        hi[i] = __byte_perm(w, 0u, 0x4342);
    }
#pragma unroll
    for (int i = 1; i < ZSEG; ++i) {
        lo[i] = __viaddmin_u16x2(lo[i - 1], ONE2, lo[i]);
        hi[i] = __viaddmin_u16x2(hi[i - 1], ONE2, hi[i]);
    }
#pragma unroll
    for (int i = ZSEG - 2; i >= 0; --i) {
        lo[i] = __viaddmin_u16x2(lo[i + 1], ONE2, lo[i]);
        hi[i] = __viaddmin_u16x2(hi[i + 1], ONE2, hi[i]);
    }
#ifdef VXPT_DF_SERIAL_CARRIES   // experiment (build.py -D...): the round-1 form, 16 threads walking the segments between two barriers
    edge_first[seg][lane] = make_uint2(lo[0], hi[0]);
    edge_last[seg][lane] = make_uint2(lo[ZSEG - 1], hi[ZSEG - 1]);
    __syncthreads();
    DF_TRACE(1, blockIdx.y * gridDim.x + blockIdx.x, 2);
    if (seg == 0) {
        uint2 c = make_uint2(INF2, INF2);
#pragma unroll
        for (int s = 0; s < ZSEGS; ++s) {
            carry_f[s][lane] = c;
            uint2 e = edge_last[s][lane];
            c.x = __viaddmin_u16x2(c.x, (uint32_t)ZSEG * ONE2, e.x);
            c.y = __viaddmin_u16x2(c.y, (uint32_t)ZSEG * ONE2, e.y);
        }
    } else if (seg == 1) {
        uint2 c = make_uint2(INF2, INF2);
#pragma unroll
        for (int s = ZSEGS - 1; s >= 0; --s) {
            carry_b[s][lane] = c;
            uint2 e = edge_first[s][lane];
            c.x = __viaddmin_u16x2(c.x, (uint32_t)ZSEG * ONE2, e.x);
            c.y = __viaddmin_u16x2(c.y, (uint32_t)ZSEG * ONE2, e.y);
        }
    }
    __syncthreads();
    uint2 cf = carry_f[seg][lane], cb = carry_b[seg][lane];
#else
    // Carries in two levels.  A warp holds four consecutive segments of eight x-words (lane = 8 j + x): min-plus scans over j by shuffles give
    // every segment the best value at its last / first voxel from sources inside the warp; the four warps exchange one value per x-word
    // and direction through shared memory (the only barrier of the kernel) and add the contribution of the warps before / after them.
    // (r02r: the round-1 form, 16 threads walking the 16 segments between two barriers, held every CTA for 2.9 us; every thread reading
    // all 16 segment edges, r02s, cost 0.8 M more warp instructions and as long.)
    const int j = seg & 3, w4 = seg >> 2;
    uint2 pf = make_uint2(lo[ZSEG - 1], hi[ZSEG - 1]), pb = make_uint2(lo[0], hi[0]);
#pragma unroll
    for (int d = 1; d < 4; d <<= 1) {
        const uint32_t fx = __shfl_up_sync(0xffffffffu, pf.x, 8 * d), fy = __shfl_up_sync(0xffffffffu, pf.y, 8 * d);
        const uint32_t bx = __shfl_down_sync(0xffffffffu, pb.x, 8 * d), by = __shfl_down_sync(0xffffffffu, pb.y, 8 * d);
        if (j >= d) {
            pf.x = __viaddmin_u16x2(fx, (uint32_t)(ZSEG * d) * ONE2, pf.x);
            pf.y = __viaddmin_u16x2(fy, (uint32_t)(ZSEG * d) * ONE2, pf.y);
        }
        if (j + d < 4) {
            pb.x = __viaddmin_u16x2(bx, (uint32_t)(ZSEG * d) * ONE2, pb.x);
            pb.y = __viaddmin_u16x2(by, (uint32_t)(ZSEG * d) * ONE2, pb.y);
        }
    }
    if (j == 3) edge_last[w4][lane] = pf;    // best value at the last voxel of the warp's four segments, from sources inside them
    if (j == 0) edge_first[w4][lane] = pb;   // ... at their first voxel
    __syncthreads();
    DF_TRACE(1, blockIdx.y * gridDim.x + blockIdx.x, 2);
    uint2 Cf = make_uint2(INF2, INF2), Cb = make_uint2(INF2, INF2);  // best value one voxel before / after the warp's segments
#pragma unroll
    for (int w2 = 0; w2 < 4; ++w2) {
        const uint2 el = edge_last[w2][lane], ef = edge_first[w2][lane];
        if (w2 < w4) {
            const uint32_t dd = (uint32_t)(4 * ZSEG * (w4 - 1 - w2)) * ONE2;
            Cf.x = __viaddmin_u16x2(el.x, dd, Cf.x);
            Cf.y = __viaddmin_u16x2(el.y, dd, Cf.y);
        } else if (w2 > w4) {
            const uint32_t dd = (uint32_t)(4 * ZSEG * (w2 - w4 - 1)) * ONE2;
            Cb.x = __viaddmin_u16x2(ef.x, dd, Cb.x);
            Cb.y = __viaddmin_u16x2(ef.y, dd, Cb.y);
        }
    }
    // the voxel in front of segment j: the last voxel of segment j - 1 of this warp (sources inside the warp), or ZSEG * j voxels past
    // the voxel in front of the warp; likewise behind
    uint2 cf, cb;
    {
        uint32_t px_ = __shfl_up_sync(0xffffffffu, pf.x, 8), py_ = __shfl_up_sync(0xffffffffu, pf.y, 8);
        uint32_t nx_ = __shfl_down_sync(0xffffffffu, pb.x, 8), ny_ = __shfl_down_sync(0xffffffffu, pb.y, 8);
        if (j == 0) { px_ = INF2; py_ = INF2; }
        if (j == 3) { nx_ = INF2; ny_ = INF2; }
        cf.x = __viaddmin_u16x2(Cf.x, (uint32_t)(ZSEG * j) * ONE2, px_);
        cf.y = __viaddmin_u16x2(Cf.y, (uint32_t)(ZSEG * j) * ONE2, py_);
        cb.x = __viaddmin_u16x2(Cb.x, (uint32_t)(ZSEG * (3 - j)) * ONE2, nx_);
        cb.y = __viaddmin_u16x2(Cb.y, (uint32_t)(ZSEG * (3 - j)) * ONE2, ny_);
    }
#endif
    DF_TRACE(1, blockIdx.y * gridDim.x + blockIdx.x, 3);
    const int x0 = (blockIdx.x * ZXW + lane) * 4, z0 = seg * ZSEG;
    // a carry improves a (1-Lipschitz) segment only if it improves the voxel it enters through: decide per warp, skip 2 x 24 steps each
    const bool need_f = __viaddmin_u16x2(cf.x, ONE2, lo[0]) != lo[0] || __viaddmin_u16x2(cf.y, ONE2, hi[0]) != hi[0];
    const bool need_b = __viaddmin_u16x2(cb.x, ONE2, lo[ZSEG - 1]) != lo[ZSEG - 1] || __viaddmin_u16x2(cb.y, ONE2, hi[ZSEG - 1]) != hi[ZSEG - 1];
    if (__any_sync(0xffffffffu, need_f)) {
#pragma unroll
        for (int i = 0; i < ZSEG; ++i) {
            lo[i] = __viaddmin_u16x2(cf.x, (uint32_t)(i + 1) * ONE2, lo[i]);
            hi[i] = __viaddmin_u16x2(cf.y, (uint32_t)(i + 1) * ONE2, hi[i]);
        }
    }
    if (__any_sync(0xffffffffu, need_b)) {
#pragma unroll
        for (int i = 0; i < ZSEG; ++i) {
            lo[i] = __viaddmin_u16x2(cb.x, (uint32_t)(ZSEG - i) * ONE2, lo[i]);
            hi[i] = __viaddmin_u16x2(cb.y, (uint32_t)(ZSEG - i) * ONE2, hi[i]);
        }
    }
    uint32_t e[4];
#pragma unroll
    for (int i = 0; i < ZSEG; ++i) {
        const uint32_t a = lo[i], b = hi[i];
        *reinterpret_cast<uint32_t*>(out + base + (size_t)i * SLICE_BYTES) = __byte_perm(a, b, 0x6420);
        if (PACK == 0) *reinterpret_cast<uint32_t*>(steps + base + (size_t)i * SLICE_BYTES) = steps4(plut, a, b);
        if (PACK == 1) {
            e[i & 3] = steps4(plut, a, b);
            if ((i & 3) == 3) *reinterpret_cast<uint4*>(steps + brick_offset(x0, y, z0 + i - 3)) = make_uint4(e[0], e[1], e[2], e[3]);
        }
    }
    DF_TRACE(1, blockIdx.y * gridDim.x + blockIdx.x, 4);
}

// ------------------------------------------------------------------------------------------------------------
// pack_steps: distance field M (linear) -> step field E(M) = (M == 1) ? 1 : floor(M * 0.57735026918f)
// (ToConservativeEuclidean + floor, InitialRayTraceFrag.glsl:90-93,329), through a 256-entry table computed on the
// host with the same single IEEE multiply.  LAYOUT 1: 8x4x4 bricks (brick_offset) — a warp moves 4 x-adjacent bricks
// (32 x-bytes x 4 y x 4 z): every lane reads one 16-byte run, every store instruction fills whole 64-byte half lines.
// LAYOUT 0: linear, 16 bytes in / 16 bytes out per lane.
// ------------------------------------------------------------------------------------------------------------
template <int LAYOUT>
__global__ void __launch_bounds__(256) pack_steps(const uint8_t* __restrict__ df, uint8_t* __restrict__ steps) {
    __shared__ uint8_t lut[256];
    lut[threadIdx.x] = c_step_lut[threadIdx.x];
    __syncthreads();
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (LAYOUT == 1) {
        constexpr int BRICKS_X = WX / 8, BRICKS_Y = WY / 4, BRICKS_Z = WZ / 4;  // 48 x 32 x 96 bricks hold voxels
        constexpr int GROUPS_X = BRICKS_X / 4;  // 12 groups of 4 bricks along x
        if (warp_global >= GROUPS_X * BRICKS_Y * BRICKS_Z) return;
        const int gx = warp_global % GROUPS_X;
        const int by = (warp_global / GROUPS_X) % BRICKS_Y;
        const int bz = warp_global / (GROUPS_X * BRICKS_Y);
        const int row = lane >> 1, half = lane & 1;  // row = (z&3)*4 + (y&3)
        const int y = by * 4 + (row & 3), z = bz * 4 + (row >> 2);
        const int x0 = gx * 32 + half * 16;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(df + (size_t)x0 + (size_t)WX * ((size_t)y + (size_t)WY * z)));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) *reinterpret_cast<uint32_t*>(steps + brick_offset(x0 + 4 * j, y, z)) = lut4(lut, w[j]);
    } else {
        const size_t i = ((size_t)warp_global * 32 + lane) * 16;
        if (i >= VOXELS) return;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(df + i));
        *reinterpret_cast<uint4*>(steps + i) = make_uint4(lut4(lut, v.x), lut4(lut, v.y), lut4(lut, v.z), lut4(lut, v.w));
    }
}

// ------------------------------------------------------------------------------------------------------------
// per-handle, per-device one-time set-up (called by vxpt_create on the handle's device)
constexpr int XY_SMEM2 = 2 * SLICE_BYTES + 16, XY_SMEM1 = SLICE_BYTES + 16;
int init_df_kernels(vxpt_ctx* c) {
    VX_CUDA(cudaFuncSetAttribute(df_xy_dpx<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, XY_SMEM1));
    VX_CUDA(cudaFuncSetAttribute(df_xy_dpx<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, XY_SMEM2));
    uint8_t lut[256];
    for (int m = 0; m < 256; ++m) lut[m] = (uint8_t)((m == 1) ? 1 : (int)floorf((float)m * 0.57735026918f));
    uint16_t pair[PAIR_LUT];
    for (int i = 0; i < PAIR_LUT; ++i) {  // index 3 * M0 + (M1 - M0 + 1)
        const int m0 = i / 3, m1 = std::min(std::max(m0 + i % 3 - 1, 0), 255);
        pair[i] = (uint16_t)(lut[m0] | (lut[m1] << 8));
    }
    VX_CUDA(cudaMemcpyToSymbolAsync(c_step_lut, lut, sizeof lut, 0, cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaMemcpyToSymbolAsync(c_pair_lut, pair, sizeof pair, 0, cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXPT_OK;
}

// Distance field AND the traversal's step field in the handle's layout (c->steps_layout says which layout the step field holds; -1 =
// stale: the reference-shaped algo 0 leaves it to launch_pack_bricks).
int launch_df_build(vxpt_ctx* c) {
    cudaStream_t s = c->stream;
    if (c->opt_df_algo == 0) {
        df_x_lines<<<(WY * WZ + 127) / 128, 128, 0, s>>>(c->d_grid, c->d_df);
        df_y_lines<<<(WX * WZ + 127) / 128, 128, 0, s>>>(c->d_df);
        df_z_lines<<<(WX * WY + 127) / 128, 128, 0, s>>>(c->d_df);
        c->steps_layout = -1;
        c->launches += 3;
    } else {
        static const int xy_ctas = std::getenv("VXPT_DF_XY_CTAS") ? std::atoi(std::getenv("VXPT_DF_XY_CTAS")) : 2 * 148;  // experiment knob: >= 384 = one CTA per slice
        if (xy_ctas >= WZ) df_xy_dpx<1><<<WZ, XY_THREADS, XY_SMEM1, s>>>(c->d_grid, c->d_tmp);   // one CTA per slice
        else df_xy_dpx<2><<<std::max(xy_ctas, 1), XY_THREADS, XY_SMEM2, s>>>(c->d_grid, c->d_tmp);      // persistent, double-buffered (default: two per SM)
        // the z sweep is a programmatic dependent of the xy sweep: its CTAs are scheduled (and load their tables) while the xy grid drains
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(WX / (4 * ZXW), WY);
        cfg.blockDim = dim3(ZXW, ZSEGS);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        const uint8_t* in = c->d_tmp;
        if (c->opt_layout == 1) VX_CUDA(cudaLaunchKernelEx(&cfg, df_z_dpx<1>, in, c->d_df, c->d_steps));
        else VX_CUDA(cudaLaunchKernelEx(&cfg, df_z_dpx<0>, in, c->d_df, c->d_steps));
        c->steps_layout = c->opt_layout;
        c->launches += 2;
    }
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

// (re)write the step field from the distance field in the handle's layout: after the reference-shaped build, or when the layout option changes
int launch_pack_bricks(vxpt_ctx* c) {
    if (c->opt_layout == 1) {
        const int warps = (WX / 32) * (WY / 4) * (WZ / 4);
        pack_steps<1><<<(warps * 32 + 255) / 256, 256, 0, c->stream>>>(c->d_df, c->d_steps);
    } else {
        pack_steps<0><<<(int)((VOXELS / 16 + 255) / 256), 256, 0, c->stream>>>(c->d_df, c->d_steps);
    }
    c->steps_layout = c->opt_layout;
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

}  // namespace vxpt

#ifdef VXPT_DF_TRACE
extern "C" __attribute__((visibility("default"))) int vxpt_debug_df_trace(unsigned long long* out /* [2][2048][16] */) {
    return (int)cudaMemcpyFromSymbol(out, vxpt::g_df_trace, sizeof(vxpt::g_df_trace));
}
extern "C" __attribute__((visibility("default"))) int vxpt_debug_df_trace_clear() {
    static unsigned long long zero[2 * 2048 * 16];
    return (int)cudaMemcpyToSymbol(vxpt::g_df_trace, zero, sizeof(zero));
}
#endif
