// trace_device.cuh — device functions shared by the trace kernels (sm_100a).
//
// VoxelTraversalDF and the ray set-up of Core/Shaders/InitialRayTraceFrag.glsl:307-374,398-414 (identical copies
// in ShadowRayTraceFrag.glsl:222-289, DiffuseRayTraceFrag.glsl:1043-1110, ReflectionTraceFrag.glsl:1088-1155).
// Bit-exactness contract: this translation unit is compiled with -fmad=false, IEEE division and
// square root; every expression keeps the association of the GLSL source.  dot = (x*x' + y*y') + z*z',
// normalize(v) = v * (1/sqrt(dot(v,v))), mat4*vec4 = (m0*x + m1*y) + (m2*z + m3*w).
#pragma once
#include <cuda_fp16.h>

#include "vxpt_internal.h"

namespace vxpt {

struct V3 {
    float x, y, z;
};
__device__ __forceinline__ V3 mk3(float x, float y, float z) { return V3{x, y, z}; }
__device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ V3 operator*(float s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ __forceinline__ float length3(V3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ V3 normalize3(V3 a) { return a * (1.0f / sqrtf(dot3(a, a))); }
__device__ __forceinline__ V3 cross3(V3 a, V3 b) { return mk3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float fractf(float x) { return x - floorf(x); }
__device__ __forceinline__ int wrap_repeat(int i, int n) {  // GL_REPEAT on a texel index
    if ((unsigned)i < (unsigned)n) return i;           // inside the texture: nearly every tap (an integer division costs ~20 instructions)
    if ((unsigned)(i + n) < (unsigned)n) return i + n;  // one period below
    if ((unsigned)(i - n) < (unsigned)n) return i - n;  // one period above
    const int m = i % n;
    return m < 0 ? m + n : m;
}
// GL_REPEAT for an index known to lie in [-n, 2n): two predicated adds, no branch
__device__ __forceinline__ int wrap_near(int i, int n) {
    i += i < 0 ? n : 0;
    return i - (i >= n ? n : 0);
}
// GLSL leaves sin/cos/pow precision open; the parity contract pins the correctly rounded fp32 value
// (double evaluation, one rounding) — these run a handful of times per GI sample, never in the traversal loop.
__device__ __forceinline__ float sin_cr(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cos_cr(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float pow_cr(float x, float y) { return (float)pow((double)x, (double)y); }

struct CameraDev {
    float inv_view[16];
    float inv_proj[16];
    int width, height, row_begin, row_end;
    int il_n, il_rank, il_band;  // interleaved row bands (VxCamera::interleave_*); il_n <= 1 = off
};

// image row of virtual row v (identity unless bands are interleaved across handles)
__device__ __forceinline__ int image_row(const CameraDev& cam, int v) {
    return cam.il_n > 1 ? ((v / cam.il_band) * cam.il_n + cam.il_rank) * cam.il_band + v % cam.il_band : v;
}

// GetRayDirectionAt (ShadowRayTraceFrag.glsl:303-308) == GetRayStuff (InitialRayTraceFrag.glsl:410-413)
__device__ __forceinline__ V3 ray_direction_at(const CameraDev& cam, float u, float v) {
    const float cx = u * 2.0f - 1.0f, cy = v * 2.0f - 1.0f, cz = -1.0f, cw = 1.0f;
    const float* P = cam.inv_proj;
    const float ex = (P[0] * cx + P[4] * cy) + (P[8] * cz + P[12] * cw);
    const float ey = (P[1] * cx + P[5] * cy) + (P[9] * cz + P[13] * cw);
    const float* M = cam.inv_view;
    const float ez = -1.0f, ew = 0.0f;
    return mk3((M[0] * ex + M[4] * ey) + (M[8] * ez + M[12] * ew), (M[1] * ex + M[5] * ey) + (M[9] * ez + M[13] * ew),
               (M[2] * ex + M[6] * ey) + (M[10] * ez + M[14] * ew));
}
__device__ __forceinline__ V3 ray_origin(const CameraDev& cam) { return mk3(cam.inv_view[12], cam.inv_view[13], cam.inv_view[14]); }

// ---- plane texel formats (VXPT_OPT_TEXEL_FORMAT) ----------------------------------------------------------------
// fmt 0: fp32 planes (parity default).  fmt 1: the reference's FBO attachment formats (Core/Pipeline.cpp:1094-1095, 1102, 1141, 1152):
// R16F / RG16F / RGBA16F texels are IEEE halves rounded to nearest even from the fp32 value, RG8 is unorm8 = round(v * 255).
// The pointer types of the plane structs are nominal; the format decides the element size.
__device__ __forceinline__ float load_f1(const float* plane, size_t px, int fmt) {
    return fmt ? __half2float(reinterpret_cast<const __half*>(plane)[px]) : plane[px];
}
__device__ __forceinline__ void store_f1(float* plane, size_t px, float v, int fmt) {
    if (fmt) reinterpret_cast<__half*>(plane)[px] = __float2half_rn(v);
    else plane[px] = v;
}
__device__ __forceinline__ float2 load_f2(const float2* plane, size_t px, int fmt) {
    if (fmt) return __half22float2(reinterpret_cast<const __half2*>(plane)[px]);
    return plane[px];
}
__device__ __forceinline__ void store_f2(float2* plane, size_t px, float a, float b, int fmt) {
    if (fmt) reinterpret_cast<__half2*>(plane)[px] = __halves2half2(__float2half_rn(a), __float2half_rn(b));
    else plane[px] = make_float2(a, b);
}
__device__ __forceinline__ float4 load_f4(const float4* plane, size_t px, int fmt) {
    if (fmt) {
        const uint2 r = reinterpret_cast<const uint2*>(plane)[px];
        const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&r.x)), hi = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
        return make_float4(lo.x, lo.y, hi.x, hi.y);
    }
    return plane[px];
}
__device__ __forceinline__ void store_f4(float4* plane, size_t px, float a, float b, float c, float d, int fmt) {
    if (fmt) {
        const __half2 lo = __halves2half2(__float2half_rn(a), __float2half_rn(b)), hi = __halves2half2(__float2half_rn(c), __float2half_rn(d));
        uint2 r;
        r.x = *reinterpret_cast<const unsigned*>(&lo);
        r.y = *reinterpret_cast<const unsigned*>(&hi);
        reinterpret_cast<uint2*>(plane)[px] = r;
    } else {
        plane[px] = make_float4(a, b, c, d);
    }
}
// RG8 unorm (values already clamped to [0, 1])
__device__ __forceinline__ void store_unorm2(float2* plane, size_t px, float a, float b, int fmt) {
    if (fmt) reinterpret_cast<uchar2*>(plane)[px] = make_uchar2((unsigned char)__float2uint_rn(a * 255.0f), (unsigned char)__float2uint_rn(b * 255.0f));
    else plane[px] = make_float2(a, b);
}

struct Counters {
    unsigned int rays, df, vox;
};

struct TraceHit {
    float t;
    int min_idx;  // axis of the last DDA step
    int sgn;      // ray sign on that axis; face normal = -sgn on min_idx
    int block;    // raw block byte at the hit
    int vx, vy, vz;
};

// ---- exact floor without the quarter-rate conversion pipe ---------------------------------------------------
// For |x| < 2^22, x + 1.5*2^23 rounded toward -inf lands on an integer of [2^23, 2^24), whose mantissa holds
// floor(x) in two's complement offset form: floor(x) = bits(sum) - bits(1.5*2^23).  FADD.RM and IADD are full-rate,
// FRND/F2I/I2F are not (ncu r01a: the loop was bound by the XU pipe).  Anything outside that range, +-inf and NaN map
// to an index outside [0, 383], i.e. "outside the volume" — the same verdict IsInVolume gives (NaN compares false;
// SURVEY.md A.3 note 5).
constexpr float FLOOR_MAGIC = 12582912.0f;      // 1.5 * 2^23
constexpr int FLOOR_MAGIC_BITS = 0x4B400000;    // its bit pattern
__device__ __forceinline__ float floor_biased(float x) { return __fadd_rd(x, FLOOR_MAGIC); }
__device__ __forceinline__ int biased_to_int(float b) { return __float_as_int(b) - FLOOR_MAGIC_BITS; }
__device__ __forceinline__ float biased_to_float(float b) { return b - FLOOR_MAGIC; }  // exact
// IsInVolume (InitialRayTraceFrag.glsl:68-78) on floor()ed coordinates
__device__ __forceinline__ bool in_volume_i(int x, int y, int z) {
    return (unsigned)x < (unsigned)WX && (unsigned)y < (unsigned)WY && (unsigned)z < (unsigned)WZ;
}

// Step field: E(M) = (M == 1) ? 1 : floor(M * 0.57735026918f) per voxel (ToConservativeEuclidean + floor,
// InitialRayTraceFrag.glsl:90-93,329), converted once per distance-field build (pack_steps in df_build.cu).
template <int LAYOUT>
__device__ __forceinline__ int fetch_step(const SceneDev& S, int x, int y, int z) {
    if (LAYOUT == 1) return S.steps[brick_offset(x, y, z)];
    return S.steps[(uint32_t)x + (uint32_t)WX * ((uint32_t)y + (uint32_t)WY * (uint32_t)z)];
}

// GetVoxel(ivec3(floor(p))) — InitialRayTraceFrag.glsl:80-88
__device__ __forceinline__ int get_voxel_at(const SceneDev& S, V3 p, Counters& cnt, int* vx = nullptr, int* vy = nullptr, int* vz = nullptr) {
    const int x = biased_to_int(floor_biased(p.x)), y = biased_to_int(floor_biased(p.y)), z = biased_to_int(floor_biased(p.z));
    if (!in_volume_i(x, y, z)) return 0;
    if (vx) { *vx = x; *vy = y; *vz = z; }
    cnt.vox++;
    return S.grid[(uint32_t)x + (uint32_t)WX * ((uint32_t)y + (uint32_t)WY * (uint32_t)z)];
}

// VoxelTraversalDF — InitialRayTraceFrag.glsl:307-374.  SURVEY.md A.3.
// Every arithmetic result equals the GLSL expression it stands for:
//  * int<->float conversions of small integers are exact whichever way they are computed;
//  * on the stepped axis the shader computes float(Grid + RaySign) + float(1 - half), an exact small integer, and then
//    adds RaySign * 0.0001f (one rounding): written here as (g + k) + n with k = RaySign + (1 - half);
//  * "Intersection = false" on leaving the volume is dropped: the final position is then outside the volume, GetVoxel
//    returns 0 and the function returns -1 either way (:362-367), so a sticky "a DDA step happened" flag suffices;
//  * the loop counter doubles as the distance-field fetch count (an iteration that fails the bounds test ends the loop).
// The loop is written as a resumable state machine (init / step / finish) so that the persistent GI kernel can interleave
// the iterations of different rays in one lane; traverse_df() is the plain composition of the three.
// (Tried and rejected, r01g: a "while-while" shape — consecutive sphere-skips in an inner loop, the lanes that need the longer
// DDA step waiting at its exit to take it together.  It executes fewer instructions but lengthens every warp's chain of dependent
// distance-field loads (sum over rounds of the longest skip run instead of the longest ray), and the loop is latency-bound:
// primary 0.148 -> 0.177 ms, shadow 0.152 -> 0.203 ms, diffuse 0.39 -> 0.85 ms.)
struct TravState {
    float dx, dy, dz;        // direction
    float ivx, ivy, ivz;     // (1.0f / direction)
    float hx, hy, hz;        // (1 + RaySign) >> 1 as floats
    float kx, ky, kz;        // RaySign + (1 - half)
    float nx, ny, nz;        // RaySign * 0.0001f
    float ox, oy, oz;        // current position
    int sx, sy, sz;          // RaySign
    int n, max_it, stepped, min_idx;
};

__device__ __forceinline__ void trav_init(TravState& t, const V3 origin, const V3 dir, int max_it) {
    t.dx = dir.x; t.dy = dir.y; t.dz = dir.z;
    t.sx = (dir.x > 0.0f) - (dir.x < 0.0f); t.sy = (dir.y > 0.0f) - (dir.y < 0.0f); t.sz = (dir.z > 0.0f) - (dir.z < 0.0f);
    const float fsx = (float)t.sx, fsy = (float)t.sy, fsz = (float)t.sz;
    t.hx = (float)((1 + t.sx) >> 1); t.hy = (float)((1 + t.sy) >> 1); t.hz = (float)((1 + t.sz) >> 1);
    t.kx = fsx + (1.0f - t.hx); t.ky = fsy + (1.0f - t.hy); t.kz = fsz + (1.0f - t.hz);
    t.nx = fsx * 0.0001f; t.ny = fsy * 0.0001f; t.nz = fsz * 0.0001f;
    t.ivx = 1.0f / dir.x; t.ivy = 1.0f / dir.y; t.ivz = 1.0f / dir.z;
    t.ox = origin.x; t.oy = origin.y; t.oz = origin.z;
    t.n = 0; t.max_it = max_it; t.stepped = 0; t.min_idx = 0;
}

// the one-voxel DDA step of VoxelTraversalDF (:340-356); gx, gy, gz = float(ivec3(origin))
__device__ __forceinline__ void dda_advance(TravState& t, const float gx, const float gy, const float gz) {
    const float wx = t.ox - gx, wy = t.oy - gy, wz = t.oz - gz;
    const float dfx = (t.hx - wx) * t.ivx, dfy = (t.hy - wy) * t.ivy, dfz = (t.hz - wz) * t.ivz;
    const bool p = dfx < dfy && t.sx != 0;
    const float c = p ? dfx : dfy;
    const bool q = c < dfz || t.sz == 0;
    const float fm = q ? c : dfz;
    t.min_idx = q ? (p ? 0 : 1) : 2;
    const float ax = gx + (wx + t.dx * fm), ay = gy + (wy + t.dy * fm), az = gz + (wz + t.dz * fm);
    const float sxp = (gx + t.kx) + t.nx, syp = (gy + t.ky) + t.ny, szp = (gz + t.kz) + t.nz;
    t.ox = (q && p) ? sxp : ax;
    t.oy = (q && !p) ? syp : ay;
    t.oz = q ? az : szp;
}

// one loop iteration; returns true when the loop has ended (cap reached, left the volume, or E == 0)
template <int LAYOUT>
__device__ __forceinline__ bool trav_step(const SceneDev& S, TravState& t) {
    if (t.n >= t.max_it) return true;
    const float bx = floor_biased(t.ox), by = floor_biased(t.oy), bz = floor_biased(t.oz);
    const int lx = biased_to_int(bx), ly = biased_to_int(by), lz = biased_to_int(bz);
    if (!in_volume_i(lx, ly, lz)) return true;
    ++t.n;
    const int euclid = fetch_step<LAYOUT>(S, lx, ly, lz);
    if (euclid >= 2) {
        const float k = (float)(euclid - 1);
        t.ox = t.ox + k * t.dx;
        t.oy = t.oy + k * t.dy;
        t.oz = t.oz + k * t.dz;
        return false;
    }
    if (euclid == 0) return true;
    // euclid == 1: one DDA step.  in-volume => origin >= 0, so ivec3(origin) (truncation) == Loc
    dda_advance(t, biased_to_float(bx), biased_to_float(by), biased_to_float(bz));
    t.stepped = 1;
    return false;
}

__device__ __forceinline__ float trav_finish(const SceneDev& S, const TravState& t, const V3 origin, TraceHit& h, Counters& cnt) {
    cnt.df += t.n;
    h.min_idx = t.min_idx;
    h.sgn = (t.min_idx == 0) ? t.sx : ((t.min_idx == 1) ? t.sy : t.sz);
    h.block = 0;
    h.vx = h.vy = h.vz = -1;
    h.t = -1.0f;
    if (t.stepped) {
        const V3 pos = mk3(t.ox, t.oy, t.oz);
        h.block = get_voxel_at(S, pos, cnt, &h.vx, &h.vy, &h.vz);
        if (h.block > 0) h.t = length3(pos - origin);
        else h.vx = h.vy = h.vz = -1;
    }
    return h.t;
}

template <int LAYOUT>
__device__ __forceinline__ float traverse_df(const SceneDev& S, const V3 origin, const V3 dir, const int max_it, TraceHit& h, Counters& cnt) {
    TravState t;
    trav_init(t, origin, dir, max_it);
    cnt.rays++;
    while (!trav_step<LAYOUT>(S, t)) {}
    return trav_finish(S, t, origin, h, cnt);
}


// ---- alpha-tested traversal (u_ShouldAlphaTest; off by default, Core/Pipeline.cpp:141-142) --------------------------------------
struct AlphaDev {
    V3 cam;          // u_InverseView[3].xyz
    float g_K;       // 1 / (tan(radians(u_FOV) / (2 * u_Dimensions.x)) * 2), main() :421 (primary) / :419 (shadow); host-computed
    float lod_bias;  // 0 in the primary shader, 2 in the shadow shader (clamp(LOD - 2.0f, 0, 8), ShadowRayTraceFrag.glsl:115)
    int flip_x;      // the primary shader flips both texture coordinates (:198-199), the shadow shader only y (:112)
};
// StopRay — InitialRayTraceFrag.glsl:189-203, ShadowRayTraceFrag.glsl:105-117.  textureLod with the integer LOD reads the nearest texel
// of exactly that level of the albedo array's alpha (vxpt_set_albedo_alpha_mips); level k starts at (4^10 - 4^(10-k)) / 3 in a layer.
__device__ __forceinline__ bool stop_ray(const SceneDev& S, const AlphaDev& A, const V3 P, const int axis, const int axis_sign, const int block) {
    const int id = min(max(block, 0), 127);         // GetBlockID
    if (S.materials[512 + id] == 0) return true;    // BlockTransparentData
    float u = 0.0f, v = 0.0f;                       // CalculateUV; a zero normal matches no branch (pinned: uv = 0)
    if (axis_sign != 0) {
        if (axis == 1) { u = fractf(P.x); v = fractf(P.z); }
        else if (axis == 0) { u = fractf(P.z); v = fractf(P.y); }
        else { u = fractf(P.x); v = fractf(P.y); }
    }
    v = 1.0f - v;
    if (A.flip_x) u = 1.0f - u;
    const float D = length3(A.cam - P);
    const float l2 = (float)log2((double)(512.0f / (1.0f / D * A.g_K)));  // pinned: correctly rounded fp32
    const int lod = (l2 > -2147483000.0f && l2 < 2147483000.0f) ? (int)l2 : 0;
    const int level = (int)clampf((float)lod - A.lod_bias, 0.0f, 8.0f);
    const int n = 512 >> level;
    const unsigned off = (1048576u - (1048576u >> (2 * level))) / 3u;
    const int i = ((int)floorf(u * (float)n)) & (n - 1), j = ((int)floorf(v * (float)n)) & (n - 1);
    const int layer = min(max(S.materials[id], 0), S.n_alpha_layers - 1);  // BlockAlbedoData
    const float alpha = (float)S.alpha_mips[(size_t)layer * VXPT_ALPHA_MIP_TEXELS + off + (unsigned)(j * n + i)] / 255.0f;
    return alpha > 0.975f;
}

// VoxelTraversalDF_AlphaTest — InitialRayTraceFrag.glsl:205-305 (cap u_RenderDistance), ShadowRayTraceFrag.glsl:119-220 (cap 350).
// As written, "known artifacts" (Pipeline.cpp:836) included: when a cut-out texel lets the ray through and the four inner DDA steps
// find nothing that stops it, control falls into the else of `if (Euclidean == 1)` with Euclidean == 0 and the ray moves back by one
// direction vector (:294-297).  The inner steps truncate (ivec3(origin)) and run without a bounds test, exactly like the shader.
// Block fetches are counted per GetVoxel call of the shader (the value is of course fetched once).
template <int LAYOUT>
__device__ __forceinline__ float traverse_df_alpha(const SceneDev& S, const AlphaDev& A, const V3 origin, const V3 dir, const int max_it,
                                                   TraceHit& h, Counters& cnt) {
    TravState t;
    trav_init(t, origin, dir, max_it);
    cnt.rays++;
    bool early = false;
    while (t.n < t.max_it) {
        const float bx = floor_biased(t.ox), by = floor_biased(t.oy), bz = floor_biased(t.oz);
        const int lx = biased_to_int(bx), ly = biased_to_int(by), lz = biased_to_int(bz);
        if (!in_volume_i(lx, ly, lz)) break;
        ++t.n;
        const int euclid = fetch_step<LAYOUT>(S, lx, ly, lz);
        if (euclid == 0) {
            int bt = get_voxel_at(S, mk3(t.ox, t.oy, t.oz), cnt);
            int sg = (t.min_idx == 0) ? t.sx : ((t.min_idx == 1) ? t.sy : t.sz);
            if (stop_ray(S, A, mk3(t.ox, t.oy, t.oz), t.min_idx, sg, bt)) break;
#pragma unroll 1
            for (int i = 0; i < 4; ++i) {
                dda_advance(t, truncf(t.ox), truncf(t.oy), truncf(t.oz));
                bt = get_voxel_at(S, mk3(t.ox, t.oy, t.oz), cnt);
                if (bt > 0) {
                    cnt.vox++;  // the shader fetches the voxel a second time (:247)
                    sg = (t.min_idx == 0) ? t.sx : ((t.min_idx == 1) ? t.sy : t.sz);
                    if (stop_ray(S, A, mk3(t.ox, t.oy, t.oz), t.min_idx, sg, bt)) { early = true; break; }
                }
            }
            if (early) break;
            t.ox = t.ox + -1.0f * t.dx;  // origin += int(Euclidean - 1) * direction with Euclidean == 0
            t.oy = t.oy + -1.0f * t.dy;
            t.oz = t.oz + -1.0f * t.dz;
        } else if (euclid == 1) {
            dda_advance(t, biased_to_float(bx), biased_to_float(by), biased_to_float(bz));
            t.stepped = 1;
        } else {
            const float k = (float)(euclid - 1);
            t.ox = t.ox + k * t.dx;
            t.oy = t.oy + k * t.dy;
            t.oz = t.oz + k * t.dz;
        }
    }
    if (early) t.stepped = 1;  // :249-253 returns through the same three statements as :300-305
    return trav_finish(S, t, origin, h, cnt);
}

// GetNormalID — InitialRayTraceFrag.glsl:143-185
__device__ __forceinline__ int normal_id_of(const TraceHit& h) {
    const int s = -h.sgn;
    if (h.min_idx == 2) return s > 0 ? 0 : 1;
    if (h.min_idx == 1) return s > 0 ? 2 : 3;
    return s < 0 ? 4 : 5;
}
__device__ __forceinline__ V3 hit_normal(const TraceHit& h) {
    const float s = (float)(-h.sgn);
    return mk3(h.min_idx == 0 ? s : 0.0f, h.min_idx == 1 ? s : 0.0f, h.min_idx == 2 ? s : 0.0f);
}
// GetNormalFromID — ShadowRayTraceFrag.glsl:317-328 (miss (1,1,1)), DiffuseRayTraceFrag.glsl:790-801 (miss 0.5)
__device__ __forceinline__ V3 normal_from_id(int id, float miss) {
    switch (id) {
        case 0: return mk3(0.f, 0.f, 1.f);
        case 1: return mk3(0.f, 0.f, -1.f);
        case 2: return mk3(0.f, 1.f, 0.f);
        case 3: return mk3(0.f, -1.f, 0.f);
        case 4: return mk3(-1.f, 0.f, 0.f);
        case 5: return mk3(1.f, 0.f, 0.f);
        default: return mk3(miss, miss, miss);
    }
}

// one atomic per counter per warp
__device__ __forceinline__ void flush_counters(const SceneDev& S, const Counters& c) {
#ifdef VXPT_HOST_SHADOW  // tests/host_shadow: the "threads" of the g++ build run one after another, there is no warp to reduce over
    if (S.counters) {
        __atomic_fetch_add(&S.counters->rays, (unsigned long long)c.rays, __ATOMIC_RELAXED);
        __atomic_fetch_add(&S.counters->df_fetches, (unsigned long long)c.df, __ATOMIC_RELAXED);
        __atomic_fetch_add(&S.counters->vox_fetches, (unsigned long long)c.vox, __ATOMIC_RELAXED);
    }
    return;
#else
    unsigned int r = __reduce_add_sync(0xffffffffu, c.rays);
    unsigned int d = __reduce_add_sync(0xffffffffu, c.df);
    unsigned int v = __reduce_add_sync(0xffffffffu, c.vox);
    if ((threadIdx.x & 31) == 0 && S.counters) {
        atomicAdd(&S.counters->rays, (unsigned long long)r);
        atomicAdd(&S.counters->df_fetches, (unsigned long long)d);
        atomicAdd(&S.counters->vox_fetches, (unsigned long long)v);
    }
#endif
}

// pixel of this thread: a warp covers an 8x4 pixel tile, a 256-thread CTA covers 32x8 pixels.
// i, j = image pixel (all arithmetic); prow = row of the planes handed to the call (== j unless bands are interleaved)
template <int CTA_ROWS = 8>  // 8: 256 threads cover 32x8 pixels; 4: 128 threads cover 32x4 (experiment knob of the primary / shadow kernels)
__device__ __forceinline__ bool thread_pixel(const CameraDev& cam, int& i, int& j, int& prow) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    i = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    prow = cam.row_begin + blockIdx.y * CTA_ROWS + (warp >> 2) * 4 + (lane >> 3);
    j = image_row(cam, prow);
    return i < cam.width && prow < cam.row_end;
}

}  // namespace vxpt
