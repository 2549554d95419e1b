// trace_device.cuh — device functions shared by the trace kernels (sm_100a).
//
// VoxelTraversalDF and the ray set-up of Core/Shaders/InitialRayTraceFrag.glsl:307-374,398-414 (identical copies
// in ShadowRayTraceFrag.glsl:222-289, DiffuseRayTraceFrag.glsl:1043-1110, ReflectionTraceFrag.glsl:1088-1155).
// Bit-exactness contract: this translation unit is compiled with -fmad=false, IEEE division and
// square root; every expression keeps the association of the GLSL source.  dot = (x*x' + y*y') + z*z',
// normalize(v) = v * (1/sqrt(dot(v,v))), mat4*vec4 = (m0*x + m1*y) + (m2*z + m3*w).
#pragma once
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
// GLSL leaves sin/cos/pow precision open; the parity contract pins the correctly rounded fp32 value
// (double evaluation, one rounding) — these run a handful of times per GI sample, never in the traversal loop.
__device__ __forceinline__ float sin_cr(float x) { return (float)sin((double)x); }
__device__ __forceinline__ float cos_cr(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float pow_cr(float x, float y) { return (float)pow((double)x, (double)y); }

struct CameraDev {
    float inv_view[16];
    float inv_proj[16];
    int width, height, row_begin, row_end;
};

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

// IsInVolume (InitialRayTraceFrag.glsl:68-78) applied to floor()ed coordinates.  NaN compares false -> outside.
__device__ __forceinline__ bool in_volume_f(float fx, float fy, float fz) {
    return fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx <= (float)(WX - 1) && fy <= (float)(WY - 1) && fz <= (float)(WZ - 1);
}

template <int LAYOUT>
__device__ __forceinline__ int fetch_manhattan(const SceneDev& S, int x, int y, int z) {
    if (LAYOUT == 1) return S.steps[brick_offset(x, y, z)];
    return S.df[(size_t)x + (size_t)WX * ((size_t)y + (size_t)WY * (size_t)z)];
}

// GetVoxel(ivec3(floor(p))) — InitialRayTraceFrag.glsl:80-88
__device__ __forceinline__ int get_voxel_at(const SceneDev& S, V3 p, Counters& cnt, int* vx = nullptr, int* vy = nullptr, int* vz = nullptr) {
    const float fx = floorf(p.x), fy = floorf(p.y), fz = floorf(p.z);
    if (!in_volume_f(fx, fy, fz)) return 0;
    const int x = (int)fx, y = (int)fy, z = (int)fz;
    if (vx) { *vx = x; *vy = y; *vz = z; }
    cnt.vox++;
    return S.grid[(size_t)x + (size_t)WX * ((size_t)y + (size_t)WY * (size_t)z)];
}

// VoxelTraversalDF — InitialRayTraceFrag.glsl:307-374.  SURVEY.md A.3.
template <int LAYOUT>
__device__ __forceinline__ float traverse_df(const SceneDev& S, V3 origin, const V3 dir, const int max_it, TraceHit& h, Counters& cnt) {
    const V3 initial_origin = origin;
    bool intersection = false;
    int min_idx = 0;
    const int sx = (dir.x > 0.0f) - (dir.x < 0.0f), sy = (dir.y > 0.0f) - (dir.y < 0.0f), sz = (dir.z > 0.0f) - (dir.z < 0.0f);
    const int hx = (1 + sx) >> 1, hy = (1 + sy) >> 1, hz = (1 + sz) >> 1;
    const float ivx = 1.0f / dir.x, ivy = 1.0f / dir.y, ivz = 1.0f / dir.z;  // (1.0f / direction), loop invariant
    cnt.rays++;
    for (int itr = 0; itr < max_it; ++itr) {
        const float fx = floorf(origin.x), fy = floorf(origin.y), fz = floorf(origin.z);
        if (!in_volume_f(fx, fy, fz)) { intersection = false; break; }
        const int lx = (int)fx, ly = (int)fy, lz = (int)fz;
        cnt.df++;
        const int m = fetch_manhattan<LAYOUT>(S, lx, ly, lz);
        // ToConservativeEuclidean + floor (:90-93, :329)
        const int euclid = (m == 1) ? 1 : (int)floorf((float)m * 0.57735026918f);
        if (euclid == 0) break;
        if (euclid == 1) {
            // in-volume => origin >= 0, so ivec3(origin) (truncation) == Loc
            int gx = lx, gy = ly, gz = lz;
            float wx = origin.x - (float)gx, wy = origin.y - (float)gy, wz = origin.z - (float)gz;
            const float dfx = ((float)hx - wx) * ivx, dfy = ((float)hy - wy) * ivy, dfz = ((float)hz - wz) * ivz;
            min_idx = (dfx < dfy && sx != 0) ? ((dfx < dfz || sz == 0) ? 0 : 2) : ((dfy < dfz || sz == 0) ? 1 : 2);
            const float fm = (min_idx == 0) ? dfx : ((min_idx == 1) ? dfy : dfz);
            wx = wx + dir.x * fm;
            wy = wy + dir.y * fm;
            wz = wz + dir.z * fm;
            if (min_idx == 0) { gx += sx; wx = (float)(1 - hx); }
            else if (min_idx == 1) { gy += sy; wy = (float)(1 - hy); }
            else { gz += sz; wz = (float)(1 - hz); }
            origin.x = (float)gx + wx;
            origin.y = (float)gy + wy;
            origin.z = (float)gz + wz;
            if (min_idx == 0) origin.x += (float)sx * 0.0001f;
            else if (min_idx == 1) origin.y += (float)sy * 0.0001f;
            else origin.z += (float)sz * 0.0001f;
            intersection = true;
        } else {
            const float k = (float)(euclid - 1);
            origin.x = origin.x + k * dir.x;
            origin.y = origin.y + k * dir.y;
            origin.z = origin.z + k * dir.z;
        }
    }
    h.min_idx = min_idx;
    h.sgn = (min_idx == 0) ? sx : ((min_idx == 1) ? sy : sz);
    h.block = 0;
    h.vx = h.vy = h.vz = -1;
    h.t = -1.0f;
    if (intersection) {
        h.block = get_voxel_at(S, origin, cnt, &h.vx, &h.vy, &h.vz);
        if (h.block > 0) h.t = length3(origin - initial_origin);
        else h.vx = h.vy = h.vz = -1;
    }
    return h.t;
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
    unsigned int r = __reduce_add_sync(0xffffffffu, c.rays);
    unsigned int d = __reduce_add_sync(0xffffffffu, c.df);
    unsigned int v = __reduce_add_sync(0xffffffffu, c.vox);
    if ((threadIdx.x & 31) == 0 && S.counters) {
        atomicAdd(&S.counters->rays, (unsigned long long)r);
        atomicAdd(&S.counters->df_fetches, (unsigned long long)d);
        atomicAdd(&S.counters->vox_fetches, (unsigned long long)v);
    }
}

// pixel of this thread: a warp covers an 8x4 pixel tile, a 256-thread CTA covers 32x8 pixels
__device__ __forceinline__ bool thread_pixel(const CameraDev& cam, int& i, int& j) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    i = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
    j = cam.row_begin + blockIdx.y * 8 + (warp >> 2) * 4 + (lane >> 3);
    return i < cam.width && j < cam.row_end;
}

}  // namespace vxpt
