// gi_device.cuh — device helpers of the diffuse-GI pass shared by the one-thread-per-pixel kernel (trace.cu) and the
// wavefront pipeline (trace_gi.cu).  Restates Core/Shaders/DiffuseRayTraceFrag.glsl (line numbers in comments).
#pragma once
#include "trace_device.cuh"

namespace vxpt {

struct GBufferDev {
    float* t;
    uint8_t* normal_id;
    uint8_t* block_id;
    float* inv_t;
    int16_t* hit_voxel;
    int fmt;  // texel format of t (VXPT_OPT_TEXEL_FORMAT); inv_t is R32F in both
};

struct DiffuseDev {
    V3 light_color, stronger_dir;  // LIGHT_COLOR, StrongerLightDirection (:832-838), computed on the host in fp32
    int moon_stronger;
    float emissivity_mult;
    int spp, checker_spp, checkerboard, trace_length, frame, supersample;
    float hx, hy;
    float sun_visibility, gi_sky_strength, light_intensity;
};
struct DiffuseOutDev {
    float4* sh;
    float2* cocg;
    float* luma;
    float2* ao_sky;
    int fmt;
};

// trace_gi.cu
int launch_diffuse_wavefront(vxpt_ctx* c, const VxCamera& cam, const DiffuseDev& d, const VxGBuffer& g, const VxDiffuseOut& out);

constexpr float PI_F = 3.14159265359f;

// samplerBlueNoiseErrorDistribution_128x128_OptimizedFor_2d2d2d2d_32spp — DiffuseRayTraceFrag.glsl:126-149
// the table value (0..255); the sample is (0.5 + value) / 256
__device__ __forceinline__ int blue_noise_byte(const SceneDev& S, int px, int py, int sample_index, int sample_dim) {
    const int pi = px & 127, pj = py & 127;
    sample_index &= 255;
    sample_dim &= 255;
    int ridx = sample_dim + (pi + pj * 128) * 8;
    if (ridx > 131071) ridx = 131071;  // the shader runs past rankingTile here; pinned by clamping
    const int ranked = (sample_index ^ (int)S.rank[ridx]) & 255;
    int value = S.sobol[sample_dim + ranked * 256];
    value = value ^ (int)S.scramble[(sample_dim % 8) + (pi + pj * 128) * 8];
    return value;
}
__device__ __forceinline__ float blue_noise_1d(const SceneDev& S, int px, int py, int sample_index, int sample_dim) {
    return (0.5f + (float)blue_noise_byte(S, px, py, sample_index, sample_dim)) / 256.0f;
}

// texture(u_Skymap, d): bilinear inside the major-axis face, clamped at the face edge (pinned)
__device__ __forceinline__ V3 sky_sample(const SceneDev& S, V3 d) {
    const int N = S.sky_n;
    const float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int face;
    float sc, tc, ma;
    if (ax >= ay && ax >= az) { face = d.x > 0.f ? 0 : 1; sc = d.x > 0.f ? -d.z : d.z; tc = -d.y; ma = ax; }
    else if (ay >= az)        { face = d.y > 0.f ? 2 : 3; sc = d.x; tc = d.y > 0.f ? d.z : -d.z; ma = ay; }
    else                      { face = d.z > 0.f ? 4 : 5; sc = d.z > 0.f ? d.x : -d.x; tc = -d.y; ma = az; }
    const float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
    const float u = s * (float)N - 0.5f, v = t * (float)N - 0.5f;
    const float fu0 = floorf(u), fv0 = floorf(v);
    const float fu = u - fu0, fv = v - fv0;
    int i0 = (int)fu0, j0 = (int)fv0, i1 = i0 + 1, j1 = j0 + 1;
    i0 = min(max(i0, 0), N - 1); i1 = min(max(i1, 0), N - 1);
    j0 = min(max(j0, 0), N - 1); j1 = min(max(j1, 0), N - 1);
    const float* F = S.sky + (size_t)face * N * N * 3;
    const float* p00 = F + (j0 * N + i0) * 3;
    const float* p10 = F + (j0 * N + i1) * 3;
    const float* p01 = F + (j1 * N + i0) * 3;
    const float* p11 = F + (j1 * N + i1) * 3;
    const V3 a = mk3(p00[0], p00[1], p00[2]) * (1.0f - fu) + mk3(p10[0], p10[1], p10[2]) * fu;
    const V3 b = mk3(p01[0], p01[1], p01[2]) * (1.0f - fu) + mk3(p11[0], p11[1], p11[2]) * fu;
    return a * (1.0f - fv) + b * fv;
}

// GL clamps the array layer to [0, d-1].  The material table holds -1 for every id the block database does not name
// (BlockDataSSBO.cpp:15-26), and worlds may contain such ids, so a negative layer reads layer 0; layers >= d are rejected when the
// pass is called (check_diffuse / check_reflection in api.cu).
__device__ __forceinline__ V3 tex_nearest(const float4* base, int layer, int n, float u, float v) {
    const int i = ((int)floorf(u * (float)n)) & (n - 1);
    const int j = ((int)floorf(v * (float)n)) & (n - 1);
    const float4 c = __ldg(base + ((size_t)max(layer, 0) * n + j) * n + i);
    return mk3(c.x, c.y, c.z);
}
__device__ __forceinline__ float tex_bilinear1(const float* base, int layer, int n, float u, float v) {
    const float x = u * (float)n - 0.5f, y = v * (float)n - 0.5f;
    const float fx0 = floorf(x), fy0 = floorf(y);
    const float fx = x - fx0, fy = y - fy0;
    const int i0 = ((int)fx0) & (n - 1), i1 = ((int)fx0 + 1) & (n - 1);
    const int j0 = ((int)fy0) & (n - 1), j1 = ((int)fy0 + 1) & (n - 1);
    const float* L = base + (size_t)layer * n * n;
    const float a = L[j0 * n + i0] * (1.0f - fx) + L[j0 * n + i1] * fx;
    const float b = L[j1 * n + i0] * (1.0f - fx) + L[j1 * n + i1] * fx;
    return a * (1.0f - fy) + b * fy;
}

// InverseSchlick :1305-1308, DiffuseHammon :1311-1332 (rcp(x) == 1.0f / x)
// pow(x, 5.0f), pinned as the correctly rounded fp32 value: x * x is exact in double (48 bits), the two further products round at 2^-53
// each, so the double result is within 2^-52 of the true power — as close as a double-precision pow() gets — and one rounding to fp32
// follows.  x < 0 cannot occur (1 - clamped cosine); pow(0, 5) = 0 and pow(1, 5) = 1 come out exactly.
__device__ __forceinline__ float pow5_cr(float x) {
    const double d = (double)x, d2 = d * d;
    return (float)((d2 * d2) * d);
}
__device__ __forceinline__ float inverse_schlick(float f0, float voh) {
    return 1.0f - clampf(f0 + (1.0f - f0) * pow5_cr(1.0f - voh), 0.0f, 1.0f);
}
__device__ __forceinline__ float diffuse_hammon(V3 n, V3 view, V3 light, float rough) {
    const float ndl = fmaxf(dot3(n, light), 0.0f);
    if (ndl <= 0.0f) return 0.0f;
    const float ndv = fmaxf(dot3(n, view), 0.0f);
    const float ldv = fmaxf(dot3(light, view), 0.0f);
    const V3 hw = normalize3(view + light);
    const float ndh = fmaxf(dot3(n, hw), 0.0f);
    const float facing = ldv * 0.5f + 0.5f;
    const float single_rough = facing * (0.9f - 0.4f * facing) * ((0.5f + ndh) * (1.0f / fmaxf(ndh, 0.02f)));
    const float single_smooth = 1.05f * inverse_schlick(0.0f, ndl) * inverse_schlick(0.0f, fmaxf(ndv, 0.0f));
    const float single = clampf(mixf(single_smooth, single_rough, rough) * (1.0f / PI_F), 0.0f, 1.0f);
    const float multi = 0.1159f * rough;
    return clampf((multi + single) * ndl, 0.0f, 1.0f);
}

// SampleBlueNoise2D :811-818 + cosWeightedRandomHemisphereDirection :945-967.
// nid = the face (GetNormalFromID order 0..5) when n is one of the six axis normals, -1 otherwise: the tangent frame of an axis normal
// and sin / cos of the byte-valued angle come from S.lut (vxpt_internal.h: the same fp32 values, tabulated on the host).
__device__ __forceinline__ V3 cos_hemisphere(const SceneDev& S, int px, int py, int frame_mod128, int& bl_sample, V3 n, int nid = -1) {
    const int v1 = blue_noise_byte(S, px, py, frame_mod128, 1 + bl_sample);
    const float r2 = blue_noise_1d(S, px, py, frame_mod128, 2 + bl_sample);
    bl_sample += 2;
    V3 uu, vv;
    if ((unsigned)nid < 6u) {
        const float* b = S.lut + LUT_BASIS + 6 * nid;
        uu = mk3(b[0], b[1], b[2]);
        vv = mk3(b[3], b[4], b[5]);
    } else {
        uu = normalize3(cross3(n, mk3(0.0f, 1.0f, 1.0f)));
        vv = cross3(uu, n);
    }
    const float2 cs = *reinterpret_cast<const float2*>(S.lut + LUT_TRIG_GI + 2 * v1);  // cos, sin of 2 PI r1
    const float ra = sqrtf(r2);
    const float rx = ra * cs.x;
    const float ry = ra * cs.y;
    const float rz = sqrtf(1.0f - r2);
    const V3 rr = (rx * uu + ry * vv) + rz * n;
    return normalize3(rr);
}
// face of an axis normal given as (axis, sign of the normal on it); -1 when the sign is 0 (no face)
__device__ __forceinline__ int face_of(int axis, int s) {
    if (s == 0) return -1;
    if (axis == 2) return s > 0 ? 0 : 1;
    if (axis == 1) return s > 0 ? 2 : 3;
    return s < 0 ? 4 : 5;
}

// CalculateUV :1235-1273 on an exact axis normal
__device__ __forceinline__ void calc_uv(V3 p, int axis, float& u, float& v) {
    if (axis == 1) { u = fractf(p.x); v = fractf(p.z); }
    else if (axis == 0) { u = fractf(p.z); v = fractf(p.y); }
    else { u = fractf(p.x); v = fractf(p.y); }
}

#ifndef VXPT_HOST_SHADOW
// one warp ballot + one atomic per warp; returns this lane's slot (valid where pred)
__device__ __forceinline__ unsigned warp_push(unsigned* counter, bool pred) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    const unsigned lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0 && mask) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + __popc(mask & ((1u << lane) - 1u));
}
// exclusive prefix over the 256 bins of a CTA's key histogram by warp 0 (8 bins per lane, then a warp scan of the lane sums);
// hist[256] receives the total
__device__ __forceinline__ void prefix_256(unsigned* hist, unsigned tid) {
    if (tid < 32) {
        unsigned c[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { c[k] = hist[tid * 8 + k]; sum += c[k]; }
        unsigned incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
            if (tid >= (unsigned)d) incl += t;
        }
        unsigned base = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) { hist[tid * 8 + k] = base; base += c[k]; }
        if (tid == 31) hist[256] = incl;
    }
}
// sort key of a hemisphere ray: small = grazing = long-lived
__device__ __forceinline__ unsigned life_key(V3 d) { return (unsigned)min((int)(fabsf(d.y) * 255.0f), 255); }
#endif

// IrridianceToSH :766-784
__device__ __forceinline__ void irradiance_to_sh(V3 rad, V3 dir, float out[6]) {
    const float Co = rad.x - rad.z;
    const float T = rad.z + Co * 0.5f;
    const float Cg = rad.y - T;
    const float Y = fmaxf(T + Cg * 0.5f, 0.0f);
    const float L00 = 0.282095f;
    const float L1_1 = 0.488603f * dir.y, L10 = 0.488603f * dir.z, L11 = 0.488603f * dir.x;
    out[0] = fmaxf(L11 * Y, -100.0f);
    out[1] = fmaxf(L1_1 * Y, -100.0f);
    out[2] = fmaxf(L10 * Y, -100.0f);
    out[3] = fmaxf(L00 * Y, -100.0f);
    out[4] = Co;
    out[5] = Cg;
}

}  // namespace vxpt
