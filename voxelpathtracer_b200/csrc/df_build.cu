// df_build.cu — Manhattan (L1) distance field over the 384x128x384 block grid, sm_100a.
//
// Replaces World::GenerateDistanceField (Core/World.cpp:69-113) and the three serial-scan compute shaders
// Core/Shaders/ManhattanDistance{X,Y,Z}.comp.  Result:
//     DF[p] = min(254, min over solid q of |p - q|_1),   solid = block byte > 0.
// The separable min-plus sweeps commute, so any order of the three axes yields the same bytes.
//
// Design (algo 1, default) — two kernels, both built on the DPX instruction VIADDMNMX.U16x2
// (__viaddmin_u16x2: per 16-bit lane min(a + b, c)), i.e. one instruction per sweep step per TWO voxels:
//   df_xy_dpx : one CTA per z-slice (49,152 contiguous bytes).  The slice is brought into shared memory with
//               one TMA bulk copy (cp.async.bulk + mbarrier), swept along x with rows paired in the two 16-bit
//               lanes (12 voxels x 2 rows per lane, cross-lane carries by a shuffle min-plus scan), swept
//               along y with x-neighbours paired (a whole 128-voxel column pair lives in registers), and
//               written back with one TMA bulk store.
//   df_z_dpx  : the z sweep; each thread owns a 24-voxel z-segment of one 4-byte x-word in registers, 16
//               segments per CTA exchange their edge values through shared memory (min-plus carries), every
//               global access is a fully used 128-byte line.
//   pack_bricks: permutes the linear field into 8x4x4 bricks (one 128-B line each, 4x2x4 per 32-B sector) for
//               the traversal kernels.
// Algo 0 keeps the reference's shape (one thread per grid line, three launches) as an on-device cross-check.
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

constexpr uint32_t ONE2 = 0x00010001u;  // +1 in both 16-bit lanes
constexpr uint32_t INF2 = 0x00FE00FEu;  // 254 in both lanes ("no solid voxel seen")

// ------------------------------------------------------------------------------------------------------------
// df_xy_dpx: x and y sweeps of one z-slice in shared memory
// 256 threads, <= 64 registers, 48 KB of shared memory: 4 CTAs per SM, so all 384 slices are resident at once (592 slots)
// and the TMA loads, the sweeps and the TMA stores of different slices overlap on every SM.
// ------------------------------------------------------------------------------------------------------------
constexpr int XY_THREADS = 256;  // 8 warps; threads 0..191 each own one x-pair column in the y sweep

// one x-sweep of a pair of rows held as r[0..11] (row A in the low 16-bit lane, row B in the high lane)
__device__ __forceinline__ void x_sweep_rows(uint32_t (&r)[12], int lane) {
    // block byte -> initial distance: solid 0, air 254 (ManhattanDistanceX.comp:51-52)
#pragma unroll
    for (int k = 0; k < 12; ++k) r[k] = INF2 - __vminu2(r[k], ONE2) * 0xFEu;
    // forward (x ascending): local sweep, min-plus scan of the lanes' last values, carry-in from the lanes to the left
#pragma unroll
    for (int k = 1; k < 12; ++k) r[k] = __viaddmin_u16x2(r[k - 1], ONE2, r[k]);
    uint32_t c = r[11];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, c, d);
        if (lane >= d) c = __viaddmin_u16x2(t, (uint32_t)(12 * d) * ONE2, c);
    }
    uint32_t cin = __shfl_up_sync(0xffffffffu, c, 1);
    if (lane == 0) cin = INF2;
#pragma unroll
    for (int k = 0; k < 12; ++k) r[k] = __viaddmin_u16x2(cin, (uint32_t)(k + 1) * ONE2, r[k]);
    // backward (x descending)
#pragma unroll
    for (int k = 10; k >= 0; --k) r[k] = __viaddmin_u16x2(r[k + 1], ONE2, r[k]);
    c = r[0];
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_down_sync(0xffffffffu, c, d);
        if (lane + d < 32) c = __viaddmin_u16x2(t, (uint32_t)(12 * d) * ONE2, c);
    }
    cin = __shfl_down_sync(0xffffffffu, c, 1);
    if (lane == 31) cin = INF2;
#pragma unroll
    for (int k = 0; k < 12; ++k) r[k] = __viaddmin_u16x2(cin, (uint32_t)(12 - k) * ONE2, r[k]);
}

__device__ __forceinline__ void load_row_pair(const uint32_t* rowA, uint32_t (&r)[12]) {
    const uint32_t* rowB = rowA + WX / 4;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint32_t A = rowA[k], B = rowB[k];
        uint32_t t0 = __byte_perm(A, B, 0x6420);  // a0 a2 b0 b2
        uint32_t t1 = __byte_perm(A, B, 0x7531);  // a1 a3 b1 b3
        r[4 * k + 0] = t0 & 0x00FF00FFu;
        r[4 * k + 2] = (t0 >> 8) & 0x00FF00FFu;
        r[4 * k + 1] = t1 & 0x00FF00FFu;
        r[4 * k + 3] = (t1 >> 8) & 0x00FF00FFu;
    }
}
__device__ __forceinline__ void store_row_pair(uint32_t* rowA, const uint32_t (&r)[12]) {
    uint32_t* rowB = rowA + WX / 4;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        uint32_t t0 = r[4 * k + 0] | (r[4 * k + 2] << 8);  // a0 a2 b0 b2
        uint32_t t1 = r[4 * k + 1] | (r[4 * k + 3] << 8);  // a1 a3 b1 b3
        rowA[k] = __byte_perm(t0, t1, 0x5140);
        rowB[k] = __byte_perm(t0, t1, 0x7362);
    }
}

__global__ void __launch_bounds__(XY_THREADS, 4) df_xy_dpx(const uint8_t* __restrict__ grid, uint8_t* __restrict__ out) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* tile = smem;                                            // [128][384] bytes
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + SLICE_BYTES);  // mbarrier
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const size_t slice = (size_t)blockIdx.x * SLICE_BYTES;

    if (tid == 0) {
        mbar_init(bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(bar, SLICE_BYTES);
        bulk_g2s(tile, grid + slice, SLICE_BYTES, bar);
    }
    mbar_wait(bar, 0);

    // ---- x sweep: a warp takes row pairs (2p, 2p+1); lane l holds x = 12l .. 12l+11 of both rows.  Two row pairs are in
    //      flight per warp (independent dependency chains) to cover the latency of the DPX / shuffle chains.
    uint32_t* t32 = reinterpret_cast<uint32_t*>(tile);
    constexpr int WARPS = XY_THREADS / 32;  // 8 warps x 8 row pairs = 64 row pairs
#pragma unroll 1
    for (int p = warp; p < WY / 2; p += 2 * WARPS) {
        uint32_t* rowA0 = t32 + (2 * p) * (WX / 4) + lane * 3;
        uint32_t* rowA1 = rowA0 + 2 * WARPS * (WX / 4);  // row pair p + 8 = 16 rows further
        uint32_t r0[12], r1[12];
        load_row_pair(rowA0, r0);
        load_row_pair(rowA1, r1);
        x_sweep_rows(r0, lane);
        x_sweep_rows(r1, lane);
        store_row_pair(rowA0, r0);
        store_row_pair(rowA1, r1);
    }
    __syncthreads();

    // ---- y sweep: thread t owns voxels x = 2t, 2t+1 (x-neighbours in the two 16-bit lanes); the column is streamed
    //      through shared memory (loads run ahead of the one-instruction dependency chain), not held in registers.
    if (tid < WX / 2) {
        uint16_t* col = reinterpret_cast<uint16_t*>(tile) + tid;
        uint32_t prev = __byte_perm((uint32_t)col[0], 0u, 0x4140);
#pragma unroll 16
        for (int y = 1; y < WY; ++y) {
            uint32_t cur = __byte_perm((uint32_t)col[y * (WX / 2)], 0u, 0x4140);
            prev = __viaddmin_u16x2(prev, ONE2, cur);
            col[y * (WX / 2)] = (uint16_t)__byte_perm(prev, 0u, 0x4420);
        }
#pragma unroll 16
        for (int y = WY - 2; y >= 0; --y) {
            uint32_t cur = __byte_perm((uint32_t)col[y * (WX / 2)], 0u, 0x4140);
            prev = __viaddmin_u16x2(prev, ONE2, cur);
            col[y * (WX / 2)] = (uint16_t)__byte_perm(prev, 0u, 0x4420);
        }
    }
    fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the bulk-copy (async) proxy
    __syncthreads();
    if (tid == 0) {
        bulk_s2g(out + slice, tile, SLICE_BYTES);
        bulk_commit();
        bulk_wait_read0();  // shared memory must outlive the copy's reads
    }
}

// ------------------------------------------------------------------------------------------------------------
// df_z_dpx: z sweep, in place capable (in may equal out)
// ------------------------------------------------------------------------------------------------------------
constexpr int ZSEG = 24;          // voxels per thread along z
constexpr int ZSEGS = WZ / ZSEG;  // 16 segments per CTA
// x-words (4 bytes) per segment row of a CTA.  8 (= one 32-byte sector per row, 128-thread CTAs, 1536 of them) instead of a full warp
// (128-byte lines, 512-thread CTAs, 384 of them): with 63 registers only two 512-thread CTAs fit an SM, so 384 CTAs ran as 1.3 waves
// on 296 slots and the SMs idled 41 % of the kernel (ncu r01g); small CTAs keep eight resident per SM and refill as they retire.
constexpr int ZXW = 8;

__constant__ uint8_t c_step_lut[256];

__device__ __forceinline__ uint32_t lut4(const uint8_t* lut, uint32_t w) {
    return (uint32_t)lut[w & 0xFF] | ((uint32_t)lut[(w >> 8) & 0xFF] << 8) | ((uint32_t)lut[(w >> 16) & 0xFF] << 16) |
           ((uint32_t)lut[w >> 24] << 24);
}

// PACK (VXPT_OPT_DF_ALGO = 2): -1 = distance field only; 0 / 1 = also write the step field E(M) of pack_steps<0 / 1> from the registers that
// hold the finished words, which saves pack_steps' launch and its 18.9 MB re-read of the distance field.  In the brick layout the four
// z-neighbours of an x-word are 16 contiguous bytes (brick_offset: z & 3 has stride 4), so a thread issues one 16-byte store per four
// z values; ZSEG and the segment origins are multiples of 4.
template <int PACK>
__global__ void __launch_bounds__(ZXW * ZSEGS, 8) df_z_dpx(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, uint8_t* __restrict__ steps) {
    __shared__ uint8_t lut[PACK >= 0 ? 256 : 1];
    if (PACK >= 0) {
        const int t = threadIdx.y * ZXW + threadIdx.x;
        lut[t] = c_step_lut[t];
        lut[t + ZXW * ZSEGS] = c_step_lut[t + ZXW * ZSEGS];
    }
    __shared__ uint2 edge_first[ZSEGS][ZXW];  // local value at the first voxel of a segment (lo pair, hi pair)
    __shared__ uint2 edge_last[ZSEGS][ZXW];
    __shared__ uint2 carry_f[ZSEGS][ZXW];
    __shared__ uint2 carry_b[ZSEGS][ZXW];
    const int lane = threadIdx.x, seg = threadIdx.y;
    const int y = blockIdx.y;
    const size_t base = (size_t)y * WX + (size_t)(blockIdx.x * ZXW + lane) * 4 + (size_t)seg * ZSEG * SLICE_BYTES;

    uint32_t lo[ZSEG], hi[ZSEG];
#pragma unroll
    for (int i = 0; i < ZSEG; ++i) {
        uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(in + base + (size_t)i * SLICE_BYTES));
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
    edge_first[seg][lane] = make_uint2(lo[0], hi[0]);
    edge_last[seg][lane] = make_uint2(lo[ZSEG - 1], hi[ZSEG - 1]);
    __syncthreads();
    if (seg == 0) {  // ZXW threads run the 16-step min-plus scans for their x-words
        uint2 c = make_uint2(INF2, INF2);
#pragma unroll
        for (int s = 0; s < ZSEGS; ++s) {
            carry_f[s][lane] = c;  // best value one voxel before segment s
            uint2 e = edge_last[s][lane];
            c.x = __viaddmin_u16x2(c.x, (uint32_t)ZSEG * ONE2, e.x);
            c.y = __viaddmin_u16x2(c.y, (uint32_t)ZSEG * ONE2, e.y);
        }
    } else if (seg == 1) {
        uint2 c = make_uint2(INF2, INF2);
#pragma unroll
        for (int s = ZSEGS - 1; s >= 0; --s) {
            carry_b[s][lane] = c;  // best value one voxel after segment s
            uint2 e = edge_first[s][lane];
            c.x = __viaddmin_u16x2(c.x, (uint32_t)ZSEG * ONE2, e.x);
            c.y = __viaddmin_u16x2(c.y, (uint32_t)ZSEG * ONE2, e.y);
        }
    }
    __syncthreads();
    const uint2 cf = carry_f[seg][lane], cb = carry_b[seg][lane];
    const int x0 = (blockIdx.x * ZXW + lane) * 4, z0 = seg * ZSEG;
    uint32_t e[4];
#pragma unroll
    for (int i = 0; i < ZSEG; ++i) {
        uint32_t a = __viaddmin_u16x2(cf.x, (uint32_t)(i + 1) * ONE2, lo[i]);
        a = __viaddmin_u16x2(cb.x, (uint32_t)(ZSEG - i) * ONE2, a);
        uint32_t b = __viaddmin_u16x2(cf.y, (uint32_t)(i + 1) * ONE2, hi[i]);
        b = __viaddmin_u16x2(cb.y, (uint32_t)(ZSEG - i) * ONE2, b);
        const uint32_t w = __byte_perm(a, b, 0x6420);
        *reinterpret_cast<uint32_t*>(out + base + (size_t)i * SLICE_BYTES) = w;
        if (PACK == 0) *reinterpret_cast<uint32_t*>(steps + base + (size_t)i * SLICE_BYTES) = lut4(lut, w);
        if (PACK == 1) {
            e[i & 3] = lut4(lut, w);
            if ((i & 3) == 3) *reinterpret_cast<uint4*>(steps + brick_offset(x0, y, z0 + i - 3)) = make_uint4(e[0], e[1], e[2], e[3]);
        }
    }
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
int init_df_kernels(vxpt_ctx* c) {
    VX_CUDA(cudaFuncSetAttribute(df_xy_dpx, cudaFuncAttributeMaxDynamicSharedMemorySize, SLICE_BYTES + 16));
    uint8_t lut[256];
    for (int m = 0; m < 256; ++m) lut[m] = (uint8_t)((m == 1) ? 1 : (int)floorf((float)m * 0.57735026918f));
    VX_CUDA(cudaMemcpyToSymbolAsync(c_step_lut, lut, sizeof lut, 0, cudaMemcpyHostToDevice, c->stream));
    VX_CUDA(cudaStreamSynchronize(c->stream));
    return VXPT_OK;
}

int launch_df_build(vxpt_ctx* c) {
    cudaStream_t s = c->stream;
    if (c->opt_df_algo == 0) {
        c->steps_fused = false;
        df_x_lines<<<(WY * WZ + 127) / 128, 128, 0, s>>>(c->d_grid, c->d_df);
        df_y_lines<<<(WX * WZ + 127) / 128, 128, 0, s>>>(c->d_df);
        df_z_lines<<<(WX * WY + 127) / 128, 128, 0, s>>>(c->d_df);
        c->launches += 3;
    } else {
        const int smem = SLICE_BYTES + 16;
        df_xy_dpx<<<WZ, XY_THREADS, smem, s>>>(c->d_grid, c->d_tmp);
        const dim3 zg(WX / (4 * ZXW), WY), zb(ZXW, ZSEGS);
        c->steps_fused = false;
        if (c->opt_df_algo == 2) {  // the z sweep writes the step field too; launch_pack_bricks has nothing left to do for this build
            if (c->opt_layout == 1) df_z_dpx<1><<<zg, zb, 0, s>>>(c->d_tmp, c->d_df, c->d_steps);
            else df_z_dpx<0><<<zg, zb, 0, s>>>(c->d_tmp, c->d_df, c->d_steps);
            c->steps_layout = c->opt_layout;
            c->steps_fused = true;
        } else {
            df_z_dpx<-1><<<zg, zb, 0, s>>>(c->d_tmp, c->d_df, nullptr);
        }
        c->launches += 2;
    }
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

int launch_pack_bricks(vxpt_ctx* c) {
    if (c->steps_fused && c->steps_layout == c->opt_layout) {  // df_z_dpx<PACK> of this build already wrote the step field in this layout
        c->steps_fused = false;
        return VXPT_OK;
    }
    c->steps_fused = false;
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
