// trace.cu — primary, sun-shadow and diffuse-GI passes over the resident grid + distance field (sm_100a).
//
//   primary_kernel  : Core/Shaders/InitialRayTraceFrag.glsl main() :417-468
//   shadow_kernel   : Core/Shaders/ShadowRayTraceFrag.glsl main() :414-513
//   diffuse_kernel  : Core/Shaders/DiffuseRayTraceFrag.glsl main() :822-935, CalculateDiffuse :535-664
// Compiled with -fmad=false (see trace_device.cuh).  No CPU fallback exists for any of these.
#include "gi_device.cuh"

namespace vxpt {

// ============================================================================================= primary rays
struct PrimaryDev {
    int max_iterations, jitter_enable;
    float jx, jy;
    AlphaDev alpha;  // read by the ALPHA instantiation only (u_ShouldAlphaTest)
};
#ifndef VXPT_TRACE_CTA
#define VXPT_TRACE_CTA 256  // experiment knob: threads per CTA of the primary / shadow kernels, 256 (32x8 pixels) or 128 (32x4)
#endif
static_assert(VXPT_TRACE_CTA == 256 || VXPT_TRACE_CTA == 128, "primary / shadow CTAs are 32x8 or 32x4 pixels");
template <int LAYOUT, bool ALPHA>
#ifndef VXPT_TRACE_MINB
#define VXPT_TRACE_MINB 1   // experiment knob (build.py -D...): resident CTAs per SM the primary / shadow kernels' register allocation aims for
                            // (r02z: 8 = 32 registers, 100 % occupancy: 0.1505 / 0.1482 ms against 0.1470 / 0.1449 ms uncapped at 39 / 43)
                            // (r03x: 6 = 40 registers: 0.1497 / 0.1446 ms)
#endif
__global__ void __launch_bounds__(VXPT_TRACE_CTA, VXPT_TRACE_MINB) primary_kernel(const SceneDev S, const __grid_constant__ CameraDev cam, const PrimaryDev p,
                                                      const GBufferDev out) {
    int i, j, prow;
    const bool active = thread_pixel<VXPT_TRACE_CTA / 32>(cam, i, j, prow);
    Counters cnt = {0u, 0u, 0u};
    if (active) {
        float u = ((float)i + 0.5f) / (float)cam.width;
        float v = ((float)j + 0.5f) / (float)cam.height;
        if (p.jitter_enable) {  // GetRayStuff :404-408, TexelSize = 1.0f / u_Dimensions
            u -= p.jx * (1.0f / (float)cam.width);
            v -= p.jy * (1.0f / (float)cam.height);
        }
        const V3 dir = normalize3(ray_direction_at(cam, u, v));
        TraceHit h;
        const float t = ALPHA ? traverse_df_alpha<LAYOUT>(S, p.alpha, ray_origin(cam), dir, p.max_iterations, h, cnt)
                              : traverse_df<LAYOUT>(S, ray_origin(cam), dir, p.max_iterations, h, cnt);
        const bool intersect = t > 0.0f && h.block > 0;
        const size_t px = (size_t)prow * cam.width + i;
        if (out.t) store_f1(out.t, px, t, out.fmt);
        if (out.inv_t) out.inv_t[px] = 1.0f / t;
        if (out.normal_id) out.normal_id[px] = intersect ? (uint8_t)normal_id_of(h) : (uint8_t)VXPT_NORMAL_MISS;
        if (out.block_id) out.block_id[px] = intersect ? (uint8_t)h.block : (uint8_t)0;
        if (out.hit_voxel) {
            out.hit_voxel[3 * px + 0] = intersect ? (int16_t)h.vx : (int16_t)-1;
            out.hit_voxel[3 * px + 1] = intersect ? (int16_t)h.vy : (int16_t)-1;
            out.hit_voxel[3 * px + 2] = intersect ? (int16_t)h.vz : (int16_t)-1;
        }
    }
    flush_counters(S, cnt);
}

// ============================================================================================= sun shadow
struct ShadowDev {
    V3 light;
    int soft;
    int ioffx, ioffy;  // floor(off) of the per-frame blue-noise texel offset (:456-459), computed on the host in fp32
    float hx, hy;      // u_Halton
    AlphaDev alpha;    // read by the ALPHA instantiation only (u_ShouldAlphaTest = ShouldAlphaTestShadows)
};
struct ShadowOutDev {
    uint8_t* shadow;
    float* transversal;
    int fmt;
};

template <int LAYOUT, bool ALPHA>
__global__ void __launch_bounds__(VXPT_TRACE_CTA, VXPT_TRACE_MINB) shadow_kernel(const SceneDev S, const __grid_constant__ CameraDev cam, const ShadowDev p,
                                                     const GBufferDev g, const ShadowOutDev out) {
    int i, j, prow;
    const bool active = thread_pixel<VXPT_TRACE_CTA / 32>(cam, i, j, prow);
    Counters cnt = {0u, 0u, 0u};
    if (active) {
        const size_t px = (size_t)prow * cam.width + i;
        float u = ((float)i + 0.5f) / (float)cam.width;
        float v = ((float)j + 0.5f) / (float)cam.height;
        u += p.hx * (1.0f / (float)cam.width);
        v += p.hy * (1.0f / (float)cam.height);
        const float dist = load_f1(g.t, px, g.fmt);
        uint8_t o_shadow;
        float o_trans;
        if (dist < 0.0f) {
            o_shadow = 0;
            o_trans = 64.0f;
        } else {
            const V3 pos = ray_origin(cam) + normalize3(ray_direction_at(cam, u, v)) * dist;  // GetPositionAt :310-314
            const V3 L = p.light;
            V3 dir = L;
            if (p.soft) {
                const int tx = ((int)(((float)i + 0.5f) + (float)p.ioffx)) % 256;
                const int ty = ((int)(((float)j + 0.5f) + (float)p.ioffy)) % 256;
                const uchar4 tex = S.shadow_noise[ty * 256 + tx];
                const float xi_x = (float)tex.x / 255.0f, xi_y = (float)tex.y / 255.0f;
                const V3 T = normalize3(cross3(L, mk3(0.0f, 1.0f, 1.0f)));
                const V3 B = cross3(T, L);
                const float cos_theta_max = 0.9999505604617f;  // SampleCone :388-398, :474
                const float cos_theta = (1.0f - xi_x) + xi_x * cos_theta_max;
                const float sin_theta = sqrtf(1.0f - cos_theta * cos_theta);
                // phi = xi_y * 3.14159265359f * 2.0f is a function of the texel byte: (cos, sin) come from the host-built table (S.lut)
                const float2 cs = *reinterpret_cast<const float2*>(S.lut + LUT_TRIG_CONE + 2 * (int)tex.y);
                const V3 c = mk3(sin_theta * cs.x, sin_theta * cs.y, cos_theta);
                dir = (T * c.x + B * c.y) + L * c.z;  // mat3(T,B,L) * c
            }
            const V3 N = normal_from_id(g.normal_id[px], 1.0f);
            const float ndotl = dot3(N, dir);
            if (ndotl <= 0.01f) {
                o_shadow = 1;
                o_trans = 1.0f / 100.0f;
            } else {
                const V3 o = pos + N * mk3(0.06f, 0.06f, 0.06f);
                const int block_at = get_voxel_at(S, o, cnt);
                float T = -1.0f;
                if (dist > 0.0f) {
                    TraceHit h;
                    T = ALPHA ? traverse_df_alpha<LAYOUT>(S, p.alpha, o, dir, 350, h, cnt) : traverse_df<LAYOUT>(S, o, dir, 350, h, cnt);
                }
                o_shadow = (T > 0.0f || block_at > 0) ? 1 : 0;
                o_trans = clampf(T / 100.0f, 0.00001f, 196.0f);
                if (T < 0.0f) o_trans = 4.25f / 100.0f;
            }
        }
        if (out.shadow) out.shadow[px] = o_shadow;
        if (out.transversal) store_f1(out.transversal, px, o_trans, out.fmt);
    }
    flush_counters(S, cnt);
}

// ============================================================================================= diffuse GI
// CalculateDiffuse :535-664, one sample
template <int LAYOUT>
__device__ __forceinline__ void calculate_diffuse(const SceneDev& S, const DiffuseDev& P, int px, int py, int& bl_sample, V3 initial_origin,
                                                  V3 input_normal, int input_nid, V3& out_rad, float& out_ao, V3& odir, bool& skyhit, Counters& cnt) {
    skyhit = false;
    const float bias = 0.06f;
    const int fm = P.frame % 128;
    V3 ro = initial_origin + input_normal * bias;
    V3 rd = cos_hemisphere(S, px, py, fm, bl_sample, input_normal, input_nid);
    float ao = 1.0f;
    V3 contrib = mk3(0.f, 0.f, 0.f), thr = mk3(1.f, 1.f, 1.f);
    odir = rd;
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {  // MAX_BOUNCE_LIMIT :17
        TraceHit h;
        const float T = traverse_df<LAYOUT>(S, ro, rd, P.trace_length, h, cnt);
        const int tex_ref = min(max(h.block, 0), 127);
        const V3 ipos = ro + (rd * T);
        if (T > 0.0f && h.block > 0) {
            const V3 hn = hit_normal(h);
            float tu, tv;
            calc_uv(ipos, h.min_idx, tu, tv);
            const int albedo_layer = S.materials[tex_ref], emissive_layer = S.materials[384 + tex_ref];
            const V3 albedo = tex_nearest(S.albedo_lod3, albedo_layer, 64, tu, tv);
            const V3 pbr = tex_nearest(S.pbr_lod2, albedo_layer, 128, tu, tv);  // sic: albedo layer (:578)
            float emis = 0.0f;
            if ((float)emissive_layer >= 0.0f) {
                const float se = tex_bilinear1(S.emissive, emissive_layer, 512, tu, tv);
                emis = se * P.emissivity_mult * P.light_intensity;
            }
            const float ndl = fmaxf(dot3(hn, P.stronger_dir), 0.0f);
            float shadow_at;
            if (P.moon_stronger) shadow_at = 1.0f;
            else if (ndl < 0.001f) shadow_at = 0.0f;
            else {  // GetShadowAt :1202-1222 (u_APPLY_PLAYER_SHADOW = false)
                TraceHit hs;
                const float Ts = traverse_df<LAYOUT>(S, ipos + hn * 0.045f, P.stronger_dir, 128, hs, cnt);
                shadow_at = Ts > 0.0f ? 1.0f : 0.0f;
            }
            const V3 emis_color = (emis * mixf(1.0f, 1.0f, P.sun_visibility)) * albedo;
            const V3 neg_rd = -rd;
            const V3 sunbrdf =
                (((albedo * diffuse_hammon(hn, neg_rd, P.stronger_dir, pbr.x)) * (P.light_color * 3.5f)) * (1.0f - shadow_at)) * PI_F;
            const V3 new_dir = cos_hemisphere(S, px, py, fm, bl_sample, hn, face_of(h.min_idx, -h.sgn));
            const float cos_theta = clampf(dot3(hn, new_dir), 0.0f, 1.0f);
            const float pdf = fmaxf(cos_theta / PI_F, 0.00001f);
            const V3 atten = mk3(1.f, 1.f, 1.f) * diffuse_hammon(hn, neg_rd, new_dir, pbr.x);
            contrib = contrib + thr * sunbrdf;
            contrib = contrib + emis_color * thr;
            thr = thr * ((albedo * atten) / pdf);
            rd = new_dir;
            ro = ipos + hn * bias;
        } else {
            float x = mixf(1.0f, 1.05f, P.sun_visibility);
            x = clampf(x * 1.0f * P.gi_sky_strength, 0.0f, 5.0f);
            V3 sd = rd;
            sd.y = clampf(sd.y, 0.125f, 1.5f);  // GetSkyColorAt :984-988
            const V3 sky = sky_sample(S, sd) * x;
            contrib = contrib + sky * thr;
            skyhit = true;
            break;
        }
        if (i == 0) {
            const float dao = 2.0f;
            if (T < dao && T > 0.0f) ao = fmaxf(T / dao, 0.0f);
        }
    }
    out_rad = contrib;
    out_ao = ao;
}

template <int LAYOUT>
__global__ void __launch_bounds__(256) diffuse_kernel(const SceneDev S, const __grid_constant__ CameraDev cam, const DiffuseDev P,
                                                      const GBufferDev g, const DiffuseOutDev out) {
    int i, j, prow;
    const bool active = thread_pixel(cam, i, j, prow);
    Counters cnt = {0u, 0u, 0u};
    if (active) {
        const size_t px = (size_t)prow * cam.width + i;
        float u = ((float)i + 0.5f) / (float)cam.width;
        float v = ((float)j + 0.5f) / (float)cam.height;
        const float u0 = u, v0 = v;
        if (P.supersample) {
            u += (P.hx * 0.75f) / (float)cam.width;
            v += (P.hy * 0.75f) / (float)cam.height;
        }
        float o_sh[4], o_cocg[2], o_util = 0.0f, o_ao0 = 1.0f, o_ao1 = 0.0f;
        const float dist = load_f1(g.t, px, g.fmt);
        const int nid = g.normal_id[px];
        const V3 normal = normal_from_id(nid, 0.5f);
        if (dist < 0.0f) {
            float sh[6];
            const V3 vdir = normalize3(ray_direction_at(cam, u0, v0));
            irradiance_to_sh(sky_sample(S, vdir) * 2.66f, normal, sh);
            o_sh[0] = sh[0]; o_sh[1] = sh[1]; o_sh[2] = sh[2]; o_sh[3] = sh[3];
            o_cocg[0] = sh[4]; o_cocg[1] = sh[5];
        } else {
            const V3 pos = ray_origin(cam) + normalize3(ray_direction_at(cam, u, v)) * dist;
            int spp = min(max(P.spp, 1), 32);
            if (P.checkerboard) {
                const bool checker = ((int)(((float)i + 0.5f) + ((float)j + 0.5f))) % 2 == P.frame % 2;
                spp = (int)mixf((float)P.spp, (float)P.checker_spp, checker ? 1.0f : 0.0f);
            }
            spp = min(max(spp, 1), 32);
            if (P.moon_stronger) spp *= 2;
            int bl_sample = 0;
            float tot0 = 0.f, tot1 = 0.f, tot2 = 0.f, tot3 = 0.f, cocg0 = 0.f, cocg1 = 0.f, acc_ao = 0.f, skyhits = 0.f;
            V3 radiance = mk3(0.f, 0.f, 0.f);
#pragma unroll 1
            for (int s = 0; s < spp; ++s) {
                V3 rad, d;
                float ao;
                bool ss;
                calculate_diffuse<LAYOUT>(S, P, i, j, bl_sample, pos, normal, nid, rad, ao, d, ss, cnt);
                rad = mk3(clampf(rad.x, 0.0f, 8.0f), clampf(rad.y, 0.0f, 8.0f), clampf(rad.z, 0.0f, 8.0f));
                radiance = radiance + rad;
                acc_ao += ao;
                float sh[6];
                irradiance_to_sh(rad, d, sh);
                tot0 += sh[0]; tot1 += sh[1]; tot2 += sh[2]; tot3 += sh[3];
                cocg0 += sh[4]; cocg1 += sh[5];
                skyhits += ss ? 1.0f : 0.0f;
            }
            const float fs = (float)spp;
            acc_ao /= fs;
            tot0 /= fs; tot1 /= fs; tot2 /= fs; tot3 /= fs;
            cocg0 /= fs; cocg1 /= fs;
            radiance = radiance / fs;
            skyhits /= fs;
            const float lum = dot3(radiance, mk3(0.299f, 0.587f, 0.114f));
            o_util = fmaxf(lum, 0.01f);
            o_ao0 = clampf(acc_ao, 0.0f, 1.0f);
            o_ao1 = clampf(skyhits, 0.0f, 1.0f);
            o_sh[0] = clampf(tot0, -100.0f, 100.0f); o_sh[1] = clampf(tot1, -100.0f, 100.0f);
            o_sh[2] = clampf(tot2, -100.0f, 100.0f); o_sh[3] = clampf(tot3, -100.0f, 100.0f);
            o_cocg[0] = clampf(cocg0, -100.0f, 100.0f);
            o_cocg[1] = clampf(cocg1, -100.0f, 100.0f);
            o_util = clampf(o_util, 0.001f, 64.0f);
        }
        if (out.sh) store_f4(out.sh, px, o_sh[0], o_sh[1], o_sh[2], o_sh[3], out.fmt);
        if (out.cocg) store_f2(out.cocg, px, o_cocg[0], o_cocg[1], out.fmt);
        if (out.luma) store_f1(out.luma, px, o_util, out.fmt);
        if (out.ao_sky) store_unorm2(out.ao_sky, px, o_ao0, o_ao1, out.fmt);
    }
    flush_counters(S, cnt);
}

// ============================================================================================= host launchers
static CameraDev to_dev(const VxCamera& cam) {
    CameraDev c;
    for (int k = 0; k < 16; ++k) { c.inv_view[k] = cam.inv_view[k]; c.inv_proj[k] = cam.inv_proj[k]; }
    c.width = cam.width; c.height = cam.height; c.row_begin = cam.row_begin; c.row_end = cam.row_end;
    c.il_n = cam.interleave_n; c.il_rank = cam.interleave_rank; c.il_band = cam.band_rows > 0 ? cam.band_rows : 1;
    return c;
}
static dim3 pixel_grid(const VxCamera& cam, int cta_rows = 8) { return dim3((cam.width + 31) / 32, (cam.row_end - cam.row_begin + cta_rows - 1) / cta_rows); }
static GBufferDev to_dev(const vxpt_ctx* c, const VxGBuffer& g) { return GBufferDev{g.t, g.normal_id, g.block_id, g.inv_t, g.hit_voxel, c->opt_texel}; }

// g_K of the alpha test's LOD (InitialRayTraceFrag.glsl:421, ShadowRayTraceFrag.glsl:419), fp32 with the pinned tan
static AlphaDev alpha_dev(const VxCamera& cam, float fov_degrees, float lod_bias, int flip_x) {
    AlphaDev a;
    a.cam = V3{cam.inv_view[12], cam.inv_view[13], cam.inv_view[14]};
    const float radians = fov_degrees * 0.01745329251994329576923690768489f;  // glm::radians
    a.g_K = 1.0f / ((float)tan((double)(radians / (2.0f * (float)cam.width))) * 2.0f);
    a.lod_bias = lod_bias;
    a.flip_x = flip_x;
    return a;
}

int launch_primary(vxpt_ctx* c, const VxCamera& cam, const VxPrimaryParams& p, const VxGBuffer& out) {
    const SceneDev S = make_scene(c);
    const PrimaryDev pd{p.max_iterations, p.jitter_enable, p.jitter[0], p.jitter[1], alpha_dev(cam, p.fov_degrees, 0.0f, 1)};
    const dim3 grid = pixel_grid(cam, VXPT_TRACE_CTA / 32);
    if (p.alpha_test) {
        if (c->opt_layout == 1) VX_LAUNCH((primary_kernel<1, true>), grid, VXPT_TRACE_CTA, c->stream, S, to_dev(cam), pd, to_dev(c, out));
        else VX_LAUNCH((primary_kernel<0, true>), grid, VXPT_TRACE_CTA, c->stream, S, to_dev(cam), pd, to_dev(c, out));
    } else if (c->opt_layout == 1) VX_LAUNCH((primary_kernel<1, false>), grid, VXPT_TRACE_CTA, c->stream, S, to_dev(cam), pd, to_dev(c, out));
    else VX_LAUNCH((primary_kernel<0, false>), grid, VXPT_TRACE_CTA, c->stream, S, to_dev(cam), pd, to_dev(c, out));
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

static inline float host_fract(float x) { return x - floorf(x); }

int launch_shadow(vxpt_ctx* c, const VxCamera& cam, const VxGBuffer& g, const VxShadowParams& p, const VxShadowOut& out) {
    const SceneDev S = make_scene(c);
    ShadowDev sd;
    sd.light = V3{p.light_dir[0], p.light_dir[1], p.light_dir[2]};
    sd.soft = p.soft;
    // ShadowRayTraceFrag.glsl:456-459 — int products wrap like GLSL ints, the rest is fp32
    const int n = p.frame % 1024;
    const int32_t ax = (int32_t)((uint32_t)n * 12664745u), ay = (int32_t)((uint32_t)n * 9560333u);
    const float offx = host_fract((float)ax / 16777216.0f) * 1024.0f, offy = host_fract((float)ay / 16777216.0f) * 1024.0f;
    sd.ioffx = (int)floorf(offx);
    sd.ioffy = (int)floorf(offy);
    sd.hx = p.halton[0];
    sd.hy = p.halton[1];
    sd.alpha = alpha_dev(cam, p.fov_degrees, 2.0f, 0);
    const ShadowOutDev od{out.shadow, out.transversal, c->opt_texel};
    const dim3 grid = pixel_grid(cam, VXPT_TRACE_CTA / 32);
    if (p.alpha_test) {
        if (c->opt_layout == 1) VX_LAUNCH((shadow_kernel<1, true>), grid, VXPT_TRACE_CTA, c->stream, S, to_dev(cam), sd, to_dev(c, g), od);
        else VX_LAUNCH((shadow_kernel<0, true>), grid, VXPT_TRACE_CTA, c->stream, S, to_dev(cam), sd, to_dev(c, g), od);
    } else if (c->opt_layout == 1) VX_LAUNCH((shadow_kernel<1, false>), grid, VXPT_TRACE_CTA, c->stream, S, to_dev(cam), sd, to_dev(c, g), od);
    else VX_LAUNCH((shadow_kernel<0, false>), grid, VXPT_TRACE_CTA, c->stream, S, to_dev(cam), sd, to_dev(c, g), od);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

int launch_diffuse(vxpt_ctx* c, const VxCamera& cam, const VxGBuffer& g, const VxDiffuseParams& p, const VxDiffuseOut& out) {
    const SceneDev S = make_scene(c);
    DiffuseDev d;
    // per-frame constants of DiffuseRayTraceFrag.glsl main() :832-838 (uniform over the frame), fp32 on the host;
    // built with -fmad=false like the device code
    const float sun[3] = {p.sun_dir[0], p.sun_dir[1], p.sun_dir[2]};
    const bool sun_stronger = -sun[1] < 0.01f;
    const float SUN_COLOR[3] = {(192.0f / 255.0f) * 16.0f, (216.0f / 255.0f) * 16.0f, (255.0f / 255.0f) * 16.0f};
    const float NIGHT_COLOR[3] = {(96.0f / 255.0f) * 1.5f, (192.0f / 255.0f) * 1.5f, (255.0f / 255.0f) * 1.5f};
    const float DUSK_COLOR[3] = {(96.0f / 255.0f) * 0.9f, (192.0f / 255.0f) * 0.9f, (255.0f / 255.0f) * 0.9f};
    float dusk = (float)pow((double)fabsf(sun[1] - 1.0f), (double)2.9f);
    dusk = fminf(fmaxf(dusk, 0.0f), 1.0f);
    float lc[3];
    for (int k = 0; k < 3; ++k) {
        const float sc = SUN_COLOR[k] * (1.0f - dusk) + DUSK_COLOR[k] * dusk;
        lc[k] = (sun_stronger ? sc : NIGHT_COLOR[k]) * (0.4f * p.gi_sun_strength);
    }
    d.light_color = V3{lc[0], lc[1], lc[2]};
    d.stronger_dir = sun_stronger ? V3{sun[0], sun[1], sun[2]} : V3{p.moon_dir[0], p.moon_dir[1], p.moon_dir[2]};
    d.moon_stronger = sun_stronger ? 0 : 1;
    d.emissivity_mult = sun_stronger ? 12.0f : 13.0f;
    d.spp = p.spp; d.checker_spp = p.checker_spp; d.checkerboard = p.checkerboard; d.trace_length = p.trace_length;
    d.frame = p.frame; d.supersample = p.supersample;
    d.hx = p.halton[0]; d.hy = p.halton[1];
    d.sun_visibility = p.sun_visibility; d.gi_sky_strength = p.gi_sky_strength; d.light_intensity = p.light_intensity;
    if (c->opt_wavefront) return launch_diffuse_wavefront(c, cam, d, g, out);
    const DiffuseOutDev od{reinterpret_cast<float4*>(out.sh), reinterpret_cast<float2*>(out.cocg), out.luma,
                           reinterpret_cast<float2*>(out.ao_sky), c->opt_texel};
    const dim3 grid = pixel_grid(cam);
    if (c->opt_layout == 1) VX_LAUNCH((diffuse_kernel<1>), grid, 256, c->stream, S, to_dev(cam), d, to_dev(c, g), od);
    else VX_LAUNCH((diffuse_kernel<0>), grid, 256, c->stream, S, to_dev(cam), d, to_dev(c, g), od);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

}  // namespace vxpt
