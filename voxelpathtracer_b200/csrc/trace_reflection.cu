// trace_reflection.cu — rough-reflection pass (Core/Shaders/ReflectionTraceFrag.glsl main() :717-1038) in the v1 parity
// profile of SURVEY.md A.6: u_ReprojectToScreenSpace, u_LPVGI, u_CloudReflections, u_ReflectPlayer, u_DeriveFromDiffuseSH
// off; lava UV distortion / flicker (functions of the wall clock) never apply.  One thread per pixel; per sample one
// GGX-sampled reflection ray (cap u_ReflectionTraceLength) and, for the first max(SPP/4,1) hits, a sun shadow ray (cap 150).
// Compiled with -fmad=false; sin/cos/pow/log are the pinned correctly rounded fp32 values (double evaluation).
#include <algorithm>
#include <cmath>

#include "gi_device.cuh"

namespace vxpt {

struct ReflDev {
    int spp, trace_length, bn_index, rough, checkerboard, frame;
    float rough_bias;                // mix(1, 0.85, u_RoughnessBias)
    float hx, hy;                    // clamp(u_Halton, -2, 2)
    V3 stronger, viewer, color_mixed;  // u_StrongerLightDirection, u_ViewerPosition, SAMPLED_COLOR_MIXED (:731)
    int grass[10];
};
struct ReflInDev {
    const float* g_normal;
    const float4* g_pbr;
    const float4* sh;
    const float2* cocg;
    int fmt;  // texel format of sh / cocg (g_normal and g_pbr stay fp32)
};
struct ReflOutDev {
    float4* color;
    float* hit_distance;
    uint8_t* emissive_mask;
    int fmt;
};

// CalculateVectors :1376-1449 on an exact axis normal (normal id 0..5)
__device__ __forceinline__ void calc_vectors(V3 p, int nid, V3& tangent, V3& bitangent, float& u, float& v) {
    if (nid <= 1) { u = fractf(p.x); v = fractf(p.y); tangent = mk3(1.f, 0.f, 0.f); bitangent = mk3(0.f, 1.f, 0.f); }
    else if (nid <= 3) { u = fractf(p.x); v = fractf(p.z); tangent = mk3(1.f, 0.f, 0.f); bitangent = mk3(0.f, 0.f, 1.f); }
    else { u = fractf(p.z); v = fractf(p.y); tangent = mk3(0.f, 0.f, -1.f); bitangent = mk3(0.f, -1.f, 0.f); }
}

// capIntersect :1264-1292, GetPlayerIntersect :1301-1307
__device__ __forceinline__ float cap_intersect(V3 ro, V3 rd, V3 pa, V3 pb, float r) {
    const V3 ba = pb - pa, oa = ro - pa;
    const float baba = dot3(ba, ba), bard = dot3(ba, rd), baoa = dot3(ba, oa), rdoa = dot3(rd, oa), oaoa = dot3(oa, oa);
    const float a = baba - bard * bard;
    float b = baba * rdoa - baoa * bard;
    float c = baba * oaoa - baoa * baoa - r * r * baba;
    float h = b * b - a * c;
    if (h >= 0.0f) {
        const float t = (-b - sqrtf(h)) / a;
        const float y = baoa + t * bard;
        if (y > 0.0f && y < baba) return t;
        const V3 oc = (y <= 0.0f) ? oa : ro - pb;
        b = dot3(rd, oc);
        c = dot3(oc, oc) - r * r;
        h = b * b - c;
        if (h > 0.0f) return -b - sqrtf(h);
    }
    return -1.0f;
}
__device__ __forceinline__ bool player_intersect(V3 viewer, V3 pos, V3 d) {
    const float x = 0.4f;
    const V3 vp = viewer + mk3(-x, -x, +x);
    return cap_intersect(pos, d, vp, vp + mk3(0.0f, 1.0f, 0.0f), 0.5f) > 0.0f;
}

// ImportanceSampleGGX :345-365
// Xi.x is 0.9 * a blue-noise sample, i.e. a function of one byte: cos / sin of phi = 2 PI Xi.x come from the handle's table
// (LUT_TRIG_GGX, the same fp32 values a double evaluation rounds to) instead of two double-precision evaluations per candidate direction
__device__ __forceinline__ V3 importance_sample_ggx(V3 N, float roughness, float2 cs_phi, float xi_y) {
    const float alpha = roughness * roughness;
    const float alpha2 = alpha * alpha;
    const float cos_theta = sqrtf((1.0f - xi_y) / (1.0f + (alpha2 - 1.0f) * xi_y));
    const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
    const V3 H = mk3(cs_phi.x * sin_theta, cs_phi.y * sin_theta, cos_theta);
    const V3 up = fabsf(N.z) < 0.999f ? mk3(0.0f, 0.0f, 1.0f) : mk3(1.0f, 0.0f, 0.0f);
    const V3 tangent = normalize3(cross3(up, N));
    const V3 bitangent = cross3(N, tangent);
    const V3 sv = (tangent * H.x + bitangent * H.y) + N * H.z;
    return normalize3(sv);
}

__device__ __forceinline__ V3 mix3(V3 a, V3 b, float t) { return a * (1.0f - t) + b * t; }

// CalculateDirectionalLight :313-336 (specular part multiplied by radiance * 0.05 * 0, as in the shader)
__device__ __forceinline__ V3 directional_light(V3 viewer, V3 world_pos, V3 light_dir, V3 radiance, V3 albedo, V3 normal, V3 pbr, float shadow) {
    const float Epsilon = 0.00001f;
    const float Shadow = fminf(shadow, 1.0f);
    const V3 Lo = normalize3(viewer - world_pos);
    const V3 N = normal;
    const float cosLo = fmaxf(0.0f, dot3(N, Lo));
    const V3 F0 = mix3(mk3(0.04f, 0.04f, 0.04f), albedo, pbr.y);
    const V3 Li = light_dir;
    const V3 Lh = normalize3(Li + Lo);
    const float cosLi = fmaxf(0.0f, dot3(N, Li));
    const float cosLh = fmaxf(0.0f, dot3(N, Lh));
    const float ct = fmaxf(0.0f, dot3(Lh, Lo));
    const V3 F = F0 + (mk3(1.f, 1.f, 1.f) - F0) * pow5_cr(1.0f - ct);  // pow(x, 5.0): gi_device.cuh
    const float alpha = pbr.x * pbr.x, alphaSq = alpha * alpha;
    const float denom = (cosLh * cosLh) * (alphaSq - 1.0f) + 1.0f;
    const float D = alphaSq / (PI_F * denom * denom);
    const float rr = pbr.x + 1.0f;
    const float k = (rr * rr) / 8.0f;
    const float G = (cosLi / (cosLi * (1.0f - k) + k)) * (cosLo / (cosLo * (1.0f - k) + k));
    const V3 kd = mix3(mk3(1.f, 1.f, 1.f) - F, mk3(0.f, 0.f, 0.f), pbr.y);
    const V3 diffuseBRDF = kd * albedo;
    const V3 specularBRDF = ((F * D) * G) / fmaxf(Epsilon, 4.0f * cosLi * cosLo);
    const V3 radiance_s = (radiance * 0.05f) * 0.0f;
    V3 res = ((diffuseBRDF * radiance) * cosLi) + ((specularBRDF * radiance_s) * cosLi);
    res = mk3(fmaxf(res.x, 0.0f), fmaxf(res.y, 0.0f), fmaxf(res.z, 0.0f));
    return res * clampf(1.0f - Shadow, 0.0f, 1.0f);
}

__device__ __forceinline__ float4 tex_nearest_rgba(const float4* base, int layer, int n, float u, float v) {
    const int i = ((int)floorf(u * (float)n)) & (n - 1);
    const int j = ((int)floorf(v * (float)n)) & (n - 1);
    return __ldg(base + ((size_t)max(layer, 0) * n + j) * n + i);  // negative layer -> 0, as tex_nearest (GL clamps the array layer)
}

// ---- the pass, cut into the pieces both kernel shapes are made of (same expressions, same order) ---------------------------------------
// what the samples of a pixel share: main() :754-818
struct ReflPixel {
    V3 pos;            // biased world position the rays start from
    V3 I;              // incident direction
    V3 nmapped;        // u_GBufferNormals (or the face normal)
    V3 base_indirect;  // SHToIrradianceA of the pixel's GI planes
    float roughness_at, metalness_at;
    int spp;
};
// false: the (jittered) G-buffer texel is the sky
__device__ __forceinline__ bool refl_pixel_setup(const SceneDev& S, const CameraDev& cam, const ReflDev& P, const GBufferDev& g, const ReflInDev& in,
                                                 int i, int j, size_t px, ReflPixel& q) {
    const float u = ((float)i + 0.5f) / (float)cam.width, v = ((float)j + 0.5f) / (float)cam.height;
    const float ju = u + (P.hx / (float)cam.width) * 1.0f;  // u_TemporalFilterReflections = true
    const float jv = v + (P.hy / (float)cam.height) * 1.0f;
    // texture(u_PositionTexture, JitteredUV) / SampleNormalFromTex(u_InitialTraceNormalTexture, JitteredUV) :766,777 — attachment 0 of
    // the primary FBO (distance) is GL_LINEAR, attachment 1 (normal id) GL_NEAREST, both GL_REPEAT (Core/Pipeline.cpp:1094,
    // Core/GLClasses/Framebuffer.cpp:64-67).  OpenGL 4.3 8.14.2: x = u * w - 0.5, i0 = floor(x) mod w, weights a * (1 - f) + b * f, x first.
    // The rows read are those of the jittered coordinate: up to ceil(|halton.y|) + 1 rows beyond the slab (vxpt.h, VxReflectionParams).
    float dist;
    int nid;
    {
        const int W = cam.width, H = cam.height;
        const float x = ju * (float)W - 0.5f, y = jv * (float)H - 0.5f;
        const float fx0 = floorf(x), fy0 = floorf(y);
        const float fx = x - fx0, fy = y - fy0;
        const int i0 = wrap_repeat((int)fx0, W), i1 = wrap_repeat((int)fx0 + 1, W), j0 = wrap_repeat((int)fy0, H), j1 = wrap_repeat((int)fy0 + 1, H);
        const float a = load_f1(g.t, (size_t)j0 * W + i0, g.fmt) * (1.0f - fx) + load_f1(g.t, (size_t)j0 * W + i1, g.fmt) * fx;
        const float b = load_f1(g.t, (size_t)j1 * W + i0, g.fmt) * (1.0f - fx) + load_f1(g.t, (size_t)j1 * W + i1, g.fmt) * fx;
        dist = a * (1.0f - fy) + b * fy;
        const int ni = wrap_repeat((int)floorf(ju * (float)W), W), nj = wrap_repeat((int)floorf(jv * (float)H), H);
        nid = g.normal_id[(size_t)nj * W + ni];
    }
    if (dist < 0.0f) return false;
    int spp = min(max(P.spp, 1), 16);
    if (P.checkerboard) {
        const bool checker = ((int)(((float)i + 0.5f) + ((float)j + 0.5f))) % 2 == (P.frame % 2);
        spp = (int)mixf((float)P.spp, (float)((P.spp + P.spp % 2) / 2), checker ? 1.0f : 0.0f);
    }
    q.spp = min(max(spp, 1), 16);
    V3 pos = ray_origin(cam) + normalize3(ray_direction_at(cam, ju, jv)) * dist;
    const V3 face_n = normal_from_id(nid, 1.0f);
    if (in.g_pbr) {
        const float4 m = in.g_pbr[px];
        q.roughness_at = m.x; q.metalness_at = m.y;
    } else {  // stand-in for the G-buffer material pass: level-2 PBR texel of the block at the primary hit
        V3 tg, bt;
        float tu, tv;
        calc_vectors(pos, nid, tg, bt, tu, tv);
        tv = 1.0f - tv;
        const float4 m = tex_nearest_rgba(S.pbr_lod2, S.materials[256 + min((int)g.block_id[px], 127)], 128, tu, tv);
        q.roughness_at = m.x; q.metalness_at = m.y;
    }
    q.I = normalize3(pos - P.viewer);
    q.pos = pos + face_n * 0.035f;
    q.nmapped = in.g_normal ? mk3(in.g_normal[3 * px], in.g_normal[3 * px + 1], in.g_normal[3 * px + 2]) : face_n;
    {  // SHToIrradianceA :469-478
        const float4 sh = load_f4(in.sh, px, in.fmt);
        const float2 cg = load_f2(in.cocg, px, in.fmt);
        const float Y = fmaxf(0.0f, 3.544905f * sh.w);
        const float sc = (Y * 0.282095f) / (sh.w + 1e-6f);
        const float c0 = cg.x * sc, c1 = cg.y * sc;
        const float T = Y - c1 * 0.5f, G = c1 + T, B = T - c0 * 0.5f, R = B + c0;
        q.base_indirect = mk3(fmaxf(R, 0.0f), fmaxf(G, 0.0f), fmaxf(B, 0.0f));
    }
    return true;
}
// direction of one sample (:841-848): best of three GGX-sampled normals (GetReflectionDirection :621-644), reflected; bl = the blue-noise
// dimension counter, which every sample advances by 6 (mod 128) when reflections are rough
__device__ __forceinline__ V3 refl_direction(const SceneDev& S, const ReflDev& P, const ReflPixel& q, int i, int j, int& bl) {
    V3 rn = q.nmapped;
    if (P.rough) {
        const float R = fmaxf(clampf(q.roughness_at * P.rough_bias, 0.01f, 1.0f), 0.05f);
        float nearest = -100.0f;
        V3 best = mk3(0.f, 0.f, 0.f);
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            const int bx = blue_noise_byte(S, i, j, P.bn_index, 1 + bl);
            const float xy = blue_noise_1d(S, i, j, P.bn_index, 2 + bl);
            bl += 2;
            bl = bl % 128;
            const float2 cs = *reinterpret_cast<const float2*>(S.lut + LUT_TRIG_GGX + 2 * bx);
            const V3 smp = importance_sample_ggx(q.nmapped, R, cs, xy * 0.65f);
            const float d = dot3(smp, q.nmapped);
            if (d > nearest) { best = smp; nearest = d; }
        }
        rn = best;
    }
    return q.I - rn * (2.0f * dot3(rn, q.I));  // reflect(I, N)
}
// per-pixel running sums of main()'s sample loop (:836-1012)
struct ReflAcc {
    float t0, t1, t2, t3, avg_hit, meaningful, mask, computed_shadow;
    int shadow_itr, total_hits;
};
__device__ __forceinline__ ReflAcc refl_acc_init() { return ReflAcc{0.f, 0.f, 0.f, 0.f, 0.001f, 0.0f, 0.0f, 0.0f, 0, 0}; }
// a sample whose ray left the scene (:1003-1010)
__device__ __forceinline__ void refl_add_miss(ReflAcc& A, const SceneDev& S, float metalness_at, V3 R) {
    const V3 atmo = sky_sample(S, normalize3(R));
    const float m = mixf(1.0f, 1.175f, metalness_at > 0.05f ? 1.0f : 0.0f);
    A.t0 += atmo.x * m; A.t1 += atmo.y * m; A.t2 += atmo.z * m; A.t3 += 1.0f;
    A.total_hits++;
}
// a sample whose ray hit (:861-1002): shading of the hit, sun shadow ray for the first max(spp / 4, 1) hits of the pixel
template <int LAYOUT>
__device__ __forceinline__ void refl_add_hit(ReflAcc& A, const SceneDev& S, const ReflDev& P, V3 pos, V3 base_indirect, int spp, V3 R, float T,
                                             const TraceHit& h, Counters& cnt) {
    const V3 hit_pos = pos + (R * T);
    const int hnid = normal_id_of(h);
    const V3 hn = hit_normal(h);
    V3 tg, bt;
    float tu, tv;
    calc_vectors(hit_pos, hnid, tg, bt, tu, tv);
    tv = 1.0f - tv;
    const int ref = min(max(h.block, 0), 127);
    int t_albedo = S.materials[ref], t_normal = S.materials[128 + ref], t_pbr = S.materials[256 + ref], t_emis = S.materials[384 + ref];
    if (ref == P.grass[0]) {  // :896-918
        if (hnid == 4 || hnid == 5 || hnid == 0 || hnid == 1) { t_albedo = P.grass[4]; t_normal = P.grass[5]; t_pbr = P.grass[6]; }
        else if (hnid == 2) { t_albedo = P.grass[1]; t_normal = P.grass[2]; t_pbr = P.grass[3]; }
        else { t_albedo = P.grass[7]; t_normal = P.grass[8]; t_pbr = P.grass[9]; }
    }
    V3 ambient = base_indirect;
    const V3 albedo = tex_nearest(S.albedo_lod3, t_albedo, 64, tu, tv);
    const V3 radiance = P.color_mixed * 0.6f;
    const float4 pbr4 = tex_nearest_rgba(S.pbr_lod2, t_pbr, 128, tu, tv);
    const float AO = pbr4.w * pbr4.w;  // pow(x, 2.0f) pinned as the correctly rounded square = the fp32 product (x * x is exact in double)
    const bool player_shadow = player_intersect(P.viewer, hit_pos + hn * 0.035f, P.stronger);
    if (A.shadow_itr < max(spp / 4, 1)) {
        if (!player_shadow) {  // GetShadowAt :1327-1346
            const V3 so = hit_pos + hn * 0.055f;
            if (player_intersect(P.viewer, so, P.stronger)) A.computed_shadow = 1.0f;
            else {
                TraceHit hs;
                const float Ts = traverse_df<LAYOUT>(S, so, P.stronger, 150, hs, cnt);
                A.computed_shadow = Ts > 0.0f ? 1.0f : 0.0f;
            }
        } else A.computed_shadow = 1.0f;
        A.shadow_itr = A.shadow_itr + 1;
    }
    ambient = ((ambient * 1.0f) * clampf(AO, 0.1f, 1.0f)) * albedo;
    const V3 nm = tex_nearest(S.normal_lod3, t_normal, 64, tu, tv) * 2.0f - mk3(1.f, 1.f, 1.f);
    const V3 nmap = (tg * nm.x + bt * nm.y) + hn * nm.z;  // TBN * n
    V3 direct = ambient + directional_light(P.viewer, hit_pos, P.stronger, radiance, albedo, nmap, mk3(pbr4.x, pbr4.y, pbr4.z), A.computed_shadow);
    if ((float)t_emis > -0.5f) {
        const int ei = ((int)floorf(tu * 128.0f)) & 127, ej = ((int)floorf(tv * 128.0f)) & 127;
        float e = S.emissive_lod2[((size_t)t_emis * 128 + ej) * 128 + ei];
        if (e > 0.1f) {
            const float lbx = 0.02501f, lby = 0.03001f;
            e *= (tu > lbx && tu < 1.0f - lbx && tv > lby && tv < 1.0f - lby) ? 1.0f : 0.0f;
            direct = albedo * fmaxf((e * 19.0f) * 1.0f, 2.0f);
            A.mask = 1.0f;
        }
    }
    A.t0 += direct.x; A.t1 += direct.y; A.t2 += direct.z; A.t3 += 1.0f;
    A.avg_hit += T;
    A.meaningful += 1.0f;
    A.total_hits++;
}
// :1014-1037
__device__ __forceinline__ void refl_store(const ReflOutDev& out, size_t px, ReflAcc A) {
    A.avg_hit /= fmaxf(A.meaningful, 0.01f);
    const float th = (float)A.total_hits;
    A.t0 /= th; A.t1 /= th; A.t2 /= th; A.t3 /= th;
    if (out.color) store_f4(out.color, px, clampf(A.t0, 0.0000001f, 100.0f), clampf(A.t1, 0.0000001f, 100.0f), clampf(A.t2, 0.0000001f, 100.0f), clampf(A.t3, 0.0000001f, 100.0f), out.fmt);
    if (out.hit_distance) store_f1(out.hit_distance, px, clampf(A.meaningful > 0.01f ? A.avg_hit : -1.0f, -10.0f, 200.0f), out.fmt);
    if (out.emissive_mask) out.emissive_mask[px] = clampf(A.mask, 0.0f, 1.0f) > 0.5f ? 1 : 0;
}
__device__ __forceinline__ void refl_store_sky(const ReflOutDev& out, size_t px) {  // :768-774
    if (out.color) store_f4(out.color, px, 0.f, 0.f, 0.f, 0.f, out.fmt);
    if (out.hit_distance) store_f1(out.hit_distance, px, -1.0f, out.fmt);
    if (out.emissive_mask) out.emissive_mask[px] = 0;
}

// ---- shape 0: one thread per pixel, the shader's loop as it stands (cross-check; also what tests/host_shadow compiles for the CPU) -----
template <int LAYOUT>
__global__ void __launch_bounds__(256) reflection_kernel(const SceneDev S, const __grid_constant__ CameraDev cam, const __grid_constant__ ReflDev P,
                                                         const GBufferDev g, const ReflInDev in, const ReflOutDev out) {
    int i, j, prow;
    const bool active = thread_pixel(cam, i, j, prow);
    Counters cnt = {0u, 0u, 0u};
    if (active) {
        const size_t px = (size_t)prow * cam.width + i;
        ReflPixel q;
        if (!refl_pixel_setup(S, cam, P, g, in, i, j, px, q)) {
            refl_store_sky(out, px);
        } else {
            ReflAcc A = refl_acc_init();
            int bl = 0;
#pragma unroll 1
            for (int s = 0; s < q.spp; ++s) {
                const V3 R = refl_direction(S, P, q, i, j, bl);
                TraceHit h;
                const float T = traverse_df<LAYOUT>(S, q.pos, R, P.trace_length, h, cnt);
                if (T > 0.0f) refl_add_hit<LAYOUT>(A, S, P, q.pos, q.base_indirect, q.spp, R, T, h, cnt);
                else refl_add_miss(A, S, q.metalness_at, R);
            }
            refl_store(out, px, A);
        }
    }
    flush_counters(S, cnt);
}

#ifndef VXPT_HOST_SHADOW
// ---- shape 1 (default): the samples re-queued, as in the GI pass (trace_gi.cu) ---------------------------------------------------------
// r02g, one thread per pixel at 1080p / 1 spp on plains: 0.75 ms, 14.6 of 32 lanes per instruction, 122 registers — three quarters of the
// reflection rays of open terrain leave the scene after a few iterations while the lanes that hit drag their warp through four texture
// fetches, two capsule tests, a Cook-Torrance term and a 150-iteration shadow ray.  Here, per sample s (launch after launch, so a pixel's
// sums grow in the shader's order and the planes are bit-identical to shape 0's):
//   refl_gen_trace  one thread per pixel: set-up, the sample's direction, the reflection ray; a miss is added on the spot (sky), a hit is
//                   compacted into a queue (one warp ballot + one atomicAdd per warp) — traversal registers only
//   refl_shade      one thread per queued hit, dense warps: the hit's shading and its sun shadow ray
//   refl_finalize   only when some pixel takes several samples: the pixel's sums -> planes
// With one sample per pixel (SPP1) the sums never leave registers: whichever kernel ends the pixel's only sample writes its planes.
struct ReflState {  // 48 B per pixel between the launches of a multi-sample pass
    float4 t;       // colour sums, alpha sum
    float4 m;       // avg_hit, meaningful, mask, computed_shadow
    int4 k;         // shadow_itr, total_hits, spp (0 = sky), -
};
__device__ __forceinline__ ReflAcc refl_state_load(const ReflState& st) {
    return ReflAcc{st.t.x, st.t.y, st.t.z, st.t.w, st.m.x, st.m.y, st.m.z, st.m.w, st.k.x, st.k.y};
}
__device__ __forceinline__ void refl_state_store(ReflState& st, const ReflAcc& A, int spp) {
    st.t = make_float4(A.t0, A.t1, A.t2, A.t3);
    st.m = make_float4(A.avg_hit, A.meaningful, A.mask, A.computed_shadow);
    st.k = make_int4(A.shadow_itr, A.total_hits, spp, 0);
}

// The rays of a CTA are SORTED before they are traced, as the first-bounce GI rays are (gi_gen_trace0): a 256-thread CTA sets up the pixels
// of RPT 32x8 tiles, files their rays under |R.y| (steep reflections off the ground leave the scene at once, grazing ones creep along it) by
// a counting sort in shared memory, and its warps claim groups of 32 rays, longest-lived first.  Sky pixels produce no ray, so they cost
// their set-up only (r02n, unsorted, one thread per pixel: 15.5 of 32 lanes per instruction, a third of the pixels of the bench frame sky).
// resident CTAs per SM the register allocation aims for (build.py -D... to experiment).  r03w, 1080p / 1 spp: uncapped (58-63 and 68-72
// registers, 4 and 3 CTAs) 0.331 ms; refl_gen_trace at 5 (48 registers) 0.318 ms; and refl_shade at 4 (64 registers) 0.311 ms
#ifndef VXPT_REFL_SHADE_MINB
#define VXPT_REFL_SHADE_MINB 4
#endif
#ifndef VXPT_REFL_GEN_MINB
#define VXPT_REFL_GEN_MINB 5
#endif
template <int LAYOUT, bool SPP1, int RPT>
__global__ void __launch_bounds__(256, VXPT_REFL_GEN_MINB) refl_gen_trace(const SceneDev S, const __grid_constant__ CameraDev cam, const __grid_constant__ ReflDev P,
                                                      const GBufferDev g, const ReflInDev in, const ReflOutDev out, ReflState* __restrict__ state,
                                                      float4* __restrict__ queue, unsigned* __restrict__ queue_count, const int sample) {
    __shared__ float4 s_a[256 * RPT];            // ray origin, pixel index (bits)
    __shared__ float4 s_b[256 * RPT];            // ray direction, (metalness > 0.05) | spp << 1 (bits)
    __shared__ unsigned s_hist[258];             // 256 bins, [256] = ray count, [257] = next group
    __shared__ unsigned short s_order[256 * RPT];
    const unsigned tid = threadIdx.x, lane = tid & 31;
    Counters cnt = {0u, 0u, 0u};
    s_hist[tid] = 0u;
    if (tid < 2) s_hist[256 + tid] = 0u;
    __syncthreads();
    unsigned kr[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
        kr[r] = ~0u;
        const int warp = tid >> 5;
        const int i = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
        const int prow = cam.row_begin + (blockIdx.y * RPT + r) * 8 + (warp >> 2) * 4 + (lane >> 3);
        if (i >= cam.width || prow >= cam.row_end) continue;
        const int j = image_row(cam, prow);
        const size_t px = (size_t)prow * cam.width + i;
        ReflPixel q;
        if (!refl_pixel_setup(S, cam, P, g, in, i, j, px, q)) {
            if (sample == 0) {
                refl_store_sky(out, px);
                if (!SPP1) state[px].k = make_int4(0, 0, 0, 0);
            }
        } else if (sample < q.spp) {
            int bl = P.rough ? (6 * sample) % 128 : 0;
            const V3 R = refl_direction(S, P, q, i, j, bl);
            const unsigned key = life_key(R);
            kr[r] = (key << 16) | atomicAdd(&s_hist[key], 1u);
            s_a[r * 256 + tid] = make_float4(q.pos.x, q.pos.y, q.pos.z, __int_as_float((int)px));
            s_b[r * 256 + tid] = make_float4(R.x, R.y, R.z, __int_as_float((q.metalness_at > 0.05f ? 1 : 0) | (q.spp << 1)));
        }
    }
    __syncthreads();
    prefix_256(s_hist, tid);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPT; ++r)
        if (kr[r] != ~0u) s_order[s_hist[kr[r] >> 16] + (kr[r] & 0xFFFFu)] = (unsigned short)(r * 256 + tid);
    __syncthreads();
    const unsigned n_rays = s_hist[256], n_groups = (n_rays + 31u) / 32u;
    while (true) {
        unsigned grp = 0;
        if (lane == 0) grp = atomicAdd(&s_hist[257], 1u);
        grp = __shfl_sync(0xffffffffu, grp, 0);
        if (grp >= n_groups) break;
        const unsigned idx = grp * 32u + lane;
        bool push = false;
        float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra, rc = ra;
        if (idx < n_rays) {
            const unsigned slot_ab = s_order[idx];
            const float4 a = s_a[slot_ab], b = s_b[slot_ab];
            const V3 pos = mk3(a.x, a.y, a.z), R = mk3(b.x, b.y, b.z);
            const size_t px = (size_t)(unsigned)__float_as_int(a.w);
            const int meta = __float_as_int(b.w), spp = meta >> 1;
            TraceHit h;
            const float T = traverse_df<LAYOUT>(S, pos, R, P.trace_length, h, cnt);
            ReflAcc A = refl_acc_init();
            if (!SPP1 && sample > 0) A = refl_state_load(state[px]);
            if (T > 0.0f) {
                push = true;
                ra = make_float4(pos.x, pos.y, pos.z, T);
                rb = make_float4(R.x, R.y, R.z, a.w);
                rc = make_float4(__int_as_float(h.min_idx | ((h.sgn + 1) << 2) | (h.block << 8)), __int_as_float(spp), 0.f, 0.f);
                if (!SPP1 && sample == 0) refl_state_store(state[px], A, spp);  // the sums start here; refl_shade adds this sample
            } else {
                refl_add_miss(A, S, (meta & 1) ? 1.0f : 0.0f, R);  // only metalness > 0.05 matters (:1005)
                if (SPP1) refl_store(out, px, A);
                else refl_state_store(state[px], A, spp);
            }
        }
        const unsigned slot = warp_push(queue_count, push);
        if (push) { queue[3 * (size_t)slot] = ra; queue[3 * (size_t)slot + 1] = rb; queue[3 * (size_t)slot + 2] = rc; }
    }
    flush_counters(S, cnt);
}

template <int LAYOUT, bool SPP1>
__global__ void __launch_bounds__(256, VXPT_REFL_SHADE_MINB) refl_shade(const SceneDev S, const __grid_constant__ ReflDev P, const ReflInDev in, const ReflOutDev out,
                                                  ReflState* __restrict__ state, const float4* __restrict__ queue, const unsigned* __restrict__ queue_count) {
    const unsigned count = queue_count[0];
    Counters cnt = {0u, 0u, 0u};
    for (unsigned rec = blockIdx.x * 256u + threadIdx.x; rec < ((count + 31u) & ~31u); rec += gridDim.x * 256u) {
        if (rec >= count) continue;
        const float4 ra = queue[3 * (size_t)rec], rb = queue[3 * (size_t)rec + 1], rc = queue[3 * (size_t)rec + 2];
        const size_t px = (size_t)(unsigned)__float_as_int(rb.w);
        const int code = __float_as_int(rc.x), spp = __float_as_int(rc.y);
        TraceHit h;
        h.min_idx = code & 3; h.sgn = ((code >> 2) & 3) - 1; h.block = (code >> 8) & 255;
        V3 base_indirect;
        {  // SHToIrradianceA :469-478 of the pixel's GI planes (as in refl_pixel_setup)
            const float4 sh = load_f4(in.sh, px, in.fmt);
            const float2 cg = load_f2(in.cocg, px, in.fmt);
            const float Y = fmaxf(0.0f, 3.544905f * sh.w);
            const float sc = (Y * 0.282095f) / (sh.w + 1e-6f);
            const float c0 = cg.x * sc, c1 = cg.y * sc;
            const float T = Y - c1 * 0.5f, G = c1 + T, B = T - c0 * 0.5f, R = B + c0;
            base_indirect = mk3(fmaxf(R, 0.0f), fmaxf(G, 0.0f), fmaxf(B, 0.0f));
        }
        ReflAcc A = SPP1 ? refl_acc_init() : refl_state_load(state[px]);
        refl_add_hit<LAYOUT>(A, S, P, mk3(ra.x, ra.y, ra.z), base_indirect, spp, mk3(rb.x, rb.y, rb.z), ra.w, h, cnt);
        if (SPP1) refl_store(out, px, A);
        else refl_state_store(state[px], A, spp);
    }
    flush_counters(S, cnt);
}

__global__ void __launch_bounds__(256) refl_finalize(const __grid_constant__ CameraDev cam, const ReflOutDev out, const ReflState* __restrict__ state) {
    int i, j, prow;
    if (!thread_pixel(cam, i, j, prow)) return;
    const size_t px = (size_t)prow * cam.width + i;
    const ReflState st = state[px];
    if (st.k.z > 0) refl_store(out, px, refl_state_load(st));
}
#endif  // VXPT_HOST_SHADOW

// ---- host side: per-frame constants (:648-668, :727-731) in fp32 with the pinned transcendental definitions -------------
static inline float h_clamp(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
static inline float h_pow(float x, float y) { return (float)pow((double)x, (double)y); }
static inline float h_log(float x) { return (float)log((double)x); }
static inline float h_srgb_to_linear(float x) { return x > 0.04045f ? h_pow(x * (1.0f / 1.055f) + 0.0521327f, 2.4f) : x / 12.92f; }
static void h_temperature_to_rgb(float kelvin, float c[3]) {
    const float t = h_clamp(kelvin, 1000.0f, 50000.0f) / 100.0f;
    if (t <= 66.0f) {
        c[0] = 1.0f;
        c[1] = h_clamp(0.39008157876901960784f * h_log(t) - 0.63184144378862745098f, 0.0f, 1.0f);
    } else {
        const float u = t - 60.0f;
        c[0] = h_clamp(1.29293618606274509804f * h_pow(u, -0.1332047592f), 0.0f, 1.0f);
        c[1] = h_clamp(1.12989086089529411765f * h_pow(u, -0.0755148492f), 0.0f, 1.0f);
    }
    if (t >= 66.0f) c[2] = 1.0f;
    else if (t <= 19.0f) c[2] = 0.0f;
    else c[2] = h_clamp(0.54320678911019607843f * h_log(t - 10.0f) - 1.19625408914f, 0.0f, 1.0f);
    for (int k = 0; k < 3; ++k) c[k] = h_srgb_to_linear(c[k]);
}
// host copy of sky_sample() (same arithmetic, same order) on the handle's host copy of the cubemap
static void h_sky_sample(const float* sky, int N, const float d[3], float out[3]) {
    const float ax = fabsf(d[0]), ay = fabsf(d[1]), az = fabsf(d[2]);
    int face;
    float sc, tc, ma;
    if (ax >= ay && ax >= az) { face = d[0] > 0.f ? 0 : 1; sc = d[0] > 0.f ? -d[2] : d[2]; tc = -d[1]; ma = ax; }
    else if (ay >= az)        { face = d[1] > 0.f ? 2 : 3; sc = d[0]; tc = d[1] > 0.f ? d[2] : -d[2]; ma = ay; }
    else                      { face = d[2] > 0.f ? 4 : 5; sc = d[2] > 0.f ? d[0] : -d[0]; tc = -d[1]; ma = az; }
    const float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
    const float u = s * (float)N - 0.5f, v = t * (float)N - 0.5f;
    const float fu0 = floorf(u), fv0 = floorf(v);
    const float fu = u - fu0, fv = v - fv0;
    int i0 = (int)fu0, j0 = (int)fv0, i1 = i0 + 1, j1 = j0 + 1;
    i0 = i0 < 0 ? 0 : (i0 > N - 1 ? N - 1 : i0); i1 = i1 < 0 ? 0 : (i1 > N - 1 ? N - 1 : i1);
    j0 = j0 < 0 ? 0 : (j0 > N - 1 ? N - 1 : j0); j1 = j1 < 0 ? 0 : (j1 > N - 1 ? N - 1 : j1);
    const float* F = sky + (size_t)face * N * N * 3;
    for (int k = 0; k < 3; ++k) {
        const float a = F[(j0 * N + i0) * 3 + k] * (1.0f - fu) + F[(j0 * N + i1) * 3 + k] * fu;
        const float b = F[(j1 * N + i0) * 3 + k] * (1.0f - fu) + F[(j1 * N + i1) * 3 + k] * fu;
        out[k] = a * (1.0f - fv) + b * fv;
    }
}

int launch_reflection(vxpt_ctx* c, const VxCamera& cam, const VxGBuffer& g, const VxReflectionIn& in, const VxReflectionParams& p,
                      const VxReflectionOut& out) {
    const SceneDev S = make_scene(c);
    ReflDev d;
    d.spp = p.spp; d.trace_length = p.trace_length; d.rough = p.rough; d.checkerboard = p.checkerboard; d.frame = p.frame;
    d.bn_index = p.frame >= 0 ? p.frame % 128 : 100;
    d.rough_bias = 1.0f * (1.0f - (p.roughness_bias ? 1.0f : 0.0f)) + 0.85f * (p.roughness_bias ? 1.0f : 0.0f);
    d.hx = h_clamp(p.halton[0], -2.0f, 2.0f);
    d.hy = h_clamp(p.halton[1], -2.0f, 2.0f);
    d.stronger = V3{p.stronger_dir[0], p.stronger_dir[1], p.stronger_dir[2]};
    d.viewer = V3{p.viewer_pos[0], p.viewer_pos[1], p.viewer_pos[2]};
    for (int k = 0; k < 10; ++k) d.grass[k] = p.grass_props[k];
    float temp[3], sky_sun[3], sky_moon[3], sun_c[3], moon_c[3];
    h_temperature_to_rgb(5778.0f, temp);
    h_sky_sample(c->h_sky.data(), c->sky_n, p.sun_dir, sky_sun);
    h_sky_sample(c->h_sky.data(), c->sky_n, p.moon_dir, sky_moon);
    for (int k = 0; k < 3; ++k) {
        sun_c[k] = (((sky_sun[k] * temp[k]) * 3.14159265359f) * 2.2f) * p.sun_strength;
        moon_c[k] = (sky_moon[k] * 3.14159265359f) * p.moon_strength;
    }
    const float lum = (moon_c[0] * 0.2125f + moon_c[1] * 0.7154f) + moon_c[2] * 0.0721f;
    for (int k = 0; k < 3; ++k) moon_c[k] = ((lum * (1.0f - 1.3f) + moon_c[k] * 1.3f) * 0.42525f) * p.moon_strength;
    float sun_vis = h_clamp(((p.sun_dir[0] * 0.0f + p.sun_dir[1] * 1.0f) + p.sun_dir[2] * 0.0f) + 0.05f, 0.0f, 0.1f) * 12.0f;
    sun_vis = 1.0f - sun_vis;
    float mixed[3];
    for (int k = 0; k < 3; ++k) mixed[k] = sun_c[k] * (1.0f - sun_vis) + moon_c[k] * sun_vis;
    d.color_mixed = V3{mixed[0], mixed[1], mixed[2]};
    CameraDev cd;
    for (int k = 0; k < 16; ++k) { cd.inv_view[k] = cam.inv_view[k]; cd.inv_proj[k] = cam.inv_proj[k]; }
    cd.width = cam.width; cd.height = cam.height; cd.row_begin = cam.row_begin; cd.row_end = cam.row_end;
    cd.il_n = cam.interleave_n; cd.il_rank = cam.interleave_rank; cd.il_band = cam.band_rows > 0 ? cam.band_rows : 1;
    const GBufferDev gd{g.t, g.normal_id, g.block_id, g.inv_t, g.hit_voxel, c->opt_texel};
    const ReflInDev id{in.g_normal, reinterpret_cast<const float4*>(in.g_pbr), reinterpret_cast<const float4*>(in.sh),
                       reinterpret_cast<const float2*>(in.cocg), c->opt_texel};
    const ReflOutDev od{reinterpret_cast<float4*>(out.color), out.hit_distance, out.emissive_mask, c->opt_texel};
    const dim3 grid((cam.width + 31) / 32, (cam.row_end - cam.row_begin + 7) / 8);
#ifndef VXPT_HOST_SHADOW
    if (c->opt_refl_wavefront) {
        // the largest per-pixel sample count (:736-742 on the host)
        int max_spp = std::min(std::max(p.spp, 1), 16);
        if (p.checkerboard) max_spp = std::max(max_spp, std::min(std::max((p.spp + p.spp % 2) / 2, 1), 16));
        const bool spp1 = max_spp == 1;
        const size_t slab_px = (size_t)(cam.row_end - cam.row_begin) * cam.width, frame_px = (size_t)cam.width * cam.height;
        // same scratch as the GI wavefront (256 B of counters | one 48-byte record per slab pixel | 48 B of sums per frame pixel when a pixel
        // takes several samples); the GI pass of the frame has finished with it (stream order).  Never grown under stream capture.
        if (int rc = grow_scratch(c, &c->d_queue, &c->queue_bytes, gi_scratch_bytes(slab_px, frame_px, spp1), "the reflection wavefront queue")) return rc;
        unsigned* count = static_cast<unsigned*>(c->d_queue);
        float4* queue = reinterpret_cast<float4*>(static_cast<char*>(c->d_queue) + 256);
        ReflState* state = spp1 ? nullptr : reinterpret_cast<ReflState*>(static_cast<char*>(c->d_queue) + 256 + slab_px * 48);
        const unsigned shade_ctas = (unsigned)std::min<size_t>(148 * 8, std::max<size_t>((slab_px + 255) / 256, 1));
#ifndef VXPT_REFL_RPT
#define VXPT_REFL_RPT 2   // experiment knob (build.py -D...): 32x8 tiles per CTA = 256 x this many rays sorted together
#endif
        constexpr int RPT = VXPT_REFL_RPT;
        const dim3 grid_s((cam.width + 31) / 32, (cam.row_end - cam.row_begin + 8 * RPT - 1) / (8 * RPT));
        for (int s = 0; s < max_spp; ++s) {
            VX_CUDA(cudaMemsetAsync(count, 0, 16, c->stream));
            if (c->opt_layout == 1) {
                if (spp1) { refl_gen_trace<1, true, RPT><<<grid_s, 256, 0, c->stream>>>(S, cd, d, gd, id, od, state, queue, count, s); refl_shade<1, true><<<shade_ctas, 256, 0, c->stream>>>(S, d, id, od, state, queue, count); }
                else { refl_gen_trace<1, false, RPT><<<grid_s, 256, 0, c->stream>>>(S, cd, d, gd, id, od, state, queue, count, s); refl_shade<1, false><<<shade_ctas, 256, 0, c->stream>>>(S, d, id, od, state, queue, count); }
            } else {
                if (spp1) { refl_gen_trace<0, true, RPT><<<grid_s, 256, 0, c->stream>>>(S, cd, d, gd, id, od, state, queue, count, s); refl_shade<0, true><<<shade_ctas, 256, 0, c->stream>>>(S, d, id, od, state, queue, count); }
                else { refl_gen_trace<0, false, RPT><<<grid_s, 256, 0, c->stream>>>(S, cd, d, gd, id, od, state, queue, count, s); refl_shade<0, false><<<shade_ctas, 256, 0, c->stream>>>(S, d, id, od, state, queue, count); }
            }
            c->launches += 2;
        }
        if (!spp1) {
            refl_finalize<<<grid, 256, 0, c->stream>>>(cd, od, state);
            c->launches += 1;
        }
        VX_CUDA(cudaGetLastError());
        return VXPT_OK;
    }
#endif
    if (c->opt_layout == 1) VX_LAUNCH((reflection_kernel<1>), grid, 256, c->stream, S, cd, d, gd, id, od);
    else VX_LAUNCH((reflection_kernel<0>), grid, 256, c->stream, S, cd, d, gd, id, od);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

}  // namespace vxpt
