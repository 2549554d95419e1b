// vxo_denoise.cpp — CPU restatement of the reference's SVGF diffuse denoiser and shadow filters, the passes that
// consume the trace passes' planes: Core/Shaders/SVGF/TemporalFilter.glsl, VarianceEstimate.glsl, SpatialFilter.glsl,
// Core/Shaders/ShadowTemporalFilter.glsl, ShadowFilter.glsl as dispatched by Core/Pipeline.cpp:2284-2640, 2854-2994.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE (see vxo_oracle.cpp's header; the same rules apply).  Pinned by the reference itself:
// oracle/_ref/libref_shaders.so holds these shaders compiled as C++ (oracle/ref_denoise_driver.cpp) and every function here equals
// them bit for bit (tests/test_svgf_denoise.py, tests/golden/ref_denoise_digests.json).
//
// Pinned definitions GL leaves to the driver, beyond vxo_oracle.cpp's list:
//   texture(sampler2D, uv) on an FBO attachment: GL_REPEAT; GL_NEAREST texel floor(u*w) mod w; GL_LINEAR (OpenGL 4.3 section 8.14.2)
//   x = u*w - 0.5, i0 = floor(x) mod w, i1 = (i0 + 1) mod w, f = x - floor(x), value = (t00*(1-fx) + t10*fx)*(1-fy) + (t01*(1-fx) + t11*fx)*fy.
//   Filters per attachment as Core/Pipeline.cpp:1094-1152 declares them: hit distance LINEAR, normal id / block id NEAREST, every
//   SH / CoCg / utility / AO / variance / shadow plane LINEAR.  Planes are fp32 (the storage rounding of the FBO formats is the caller's).
//   exp and pow are the correctly rounded fp32 values (double evaluation, one rounding).
//   mat4 * mat4: column j = ((A0*b0j + A1*b1j) + A2*b2j) + A3*b3j   (glm type_mat4x4.inl)
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../include/vxpt.h"

namespace {

struct v2 { float x, y; };
struct v3 { float x, y, z; };
struct v4 { float x, y, z, w; };
inline v4 operator+(v4 a, v4 b) { return v4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
inline v4 operator*(v4 a, float s) { return v4{a.x * s, a.y * s, a.z * s, a.w * s}; }
inline v4 operator/(v4 a, float s) { return v4{a.x / s, a.y / s, a.z / s, a.w / s}; }
inline v2 operator+(v2 a, v2 b) { return v2{a.x + b.x, a.y + b.y}; }
inline v2 operator*(v2 a, float s) { return v2{a.x * s, a.y * s}; }
inline v2 operator/(v2 a, float s) { return v2{a.x / s, a.y / s}; }
inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float fractf(float x) { return x - std::floor(x); }
inline float exp_cr(float x) { return (float)std::exp((double)x); }
inline float pow_cr(float x, float y) { return (float)std::pow((double)x, (double)y); }
inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline int wrap(int i, int n) { const int m = i % n; return m < 0 ? m + n : m; }

inline void mat4_mul_vec(const float* m, const float v[4], float out[4]) {
    for (int r = 0; r < 4; ++r) out[r] = (m[0 + r] * v[0] + m[4 + r] * v[1]) + (m[8 + r] * v[2] + m[12 + r] * v[3]);
}
inline void mat4_mul_mat(const float* a, const float* b, float* out) {
    for (int j = 0; j < 4; ++j)
        for (int r = 0; r < 4; ++r)
            out[4 * j + r] = ((a[0 + r] * b[4 * j + 0] + a[4 + r] * b[4 * j + 1]) + a[8 + r] * b[4 * j + 2]) + a[12 + r] * b[4 * j + 3];
}

// one FBO attachment
struct Tex {
    const float* d;
    int w, h, c;
    void texel(int i, int j, float out[4]) const {
        const float* p = d + ((size_t)j * w + i) * c;
        for (int k = 0; k < 4; ++k) out[k] = k < c ? p[k] : 0.0f;
    }
    void linear(float u, float v, float out[4]) const {
        const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
        const float x0 = std::floor(x), y0 = std::floor(y);
        const float fx = x - x0, fy = y - y0;
        const int i0 = wrap((int)x0, w), i1 = wrap((int)x0 + 1, w), j0 = wrap((int)y0, h), j1 = wrap((int)y0 + 1, h);
        float a[4], b[4], c2[4], e[4];
        texel(i0, j0, a); texel(i1, j0, b); texel(i0, j1, c2); texel(i1, j1, e);
        for (int k = 0; k < c; ++k) out[k] = (a[k] * (1.0f - fx) + b[k] * fx) * (1.0f - fy) + (c2[k] * (1.0f - fx) + e[k] * fx) * fy;
        for (int k = c; k < 4; ++k) out[k] = 0.0f;
    }
    float linear1(float u, float v) const { float o[4]; linear(u, v, o); return o[0]; }
    v2 linear2(float u, float v) const { float o[4]; linear(u, v, o); return v2{o[0], o[1]}; }
    v4 linear4(float u, float v) const { float o[4]; linear(u, v, o); return v4{o[0], o[1], o[2], o[3]}; }
};
// R8 id attachments (GL_NEAREST)
struct TexU8 {
    const uint8_t* d;
    int w, h;
    int nearest(float u, float v) const { return d[(size_t)wrap((int)std::floor(v * (float)h), h) * w + wrap((int)std::floor(u * (float)w), w)]; }
};
// GetNormalFromID on a normal-id texel (id / 10 in the R8 attachment, 1.0 on a miss; int(round(n * 10)) recovers the id)
inline v3 normal_of(int id) {
    switch (id) {
        case 0: return v3{0.f, 0.f, 1.f};
        case 1: return v3{0.f, 0.f, -1.f};
        case 2: return v3{0.f, 1.f, 0.f};
        case 3: return v3{0.f, -1.f, 0.f};
        case 4: return v3{-1.f, 0.f, 0.f};
        case 5: return v3{1.f, 0.f, 0.f};
        default: return v3{1.f, 1.f, 1.f};
    }
}
inline v3 ray_direction_at(const VxCamera& cam, float u, float v) {  // GetRayDirectionAt (every one of these shaders carries a copy)
    const float clip[4] = {u * 2.0f - 1.0f, v * 2.0f - 1.0f, -1.0f, 1.0f};
    float e4[4], r4[4];
    mat4_mul_vec(cam.inv_proj, clip, e4);
    const float eye[4] = {e4[0], e4[1], -1.0f, 0.0f};
    mat4_mul_vec(cam.inv_view, eye, r4);
    return v3{r4[0], r4[1], r4[2]};
}
inline v3 normalize3(v3 a) { const float s = 1.0f / std::sqrt(dot3(a, a)); return v3{a.x * s, a.y * s, a.z * s}; }
// GetPositionAt: origin + normalize(dir(txc)) * texture(pos_tex, txc).r ; .w = the distance
inline void position_at(const VxCamera& cam, const Tex& pos, float u, float v, v3& p, float& dist) {
    dist = pos.linear1(u, v);
    const v3 d = normalize3(ray_direction_at(cam, u, v));
    p = v3{cam.inv_view[12] + d.x * dist, cam.inv_view[13] + d.y * dist, cam.inv_view[14] + d.z * dist};
}
inline float sh_to_y(v4 sh) { return std::fmax(0.0f, 3.544905f * sh.w); }  // SHToY
// GradientNoise — SpatialFilter.glsl:177-182
inline float gradient_noise(int i, int j, float time) {
    const float m = time * 100.493850275f;
    const float md = m - 500.0f * std::floor(m / 500.0f);  // mod(x, y) = x - y * floor(x / y)
    const float cx = ((float)i + 0.5f) + md, cy = ((float)j + 0.5f) + md;
    return fractf(52.9829189f * fractf(0.06711056f * cx + 0.00583715f * cy));
}

}  // namespace

extern "C" {

// SVGF/TemporalFilter.glsl main() :137-362 (Core/Pipeline.cpp:2335-2432)
int vxo_svgf_temporal(const VxCamera* cam, const VxSvgfTemporalIn* in, const VxSvgfTemporalParams* prm, const VxSvgfTemporalOut* out) {
    const int W = cam->width, H = cam->height;
    const Tex cur_pos{in->current.t, W, H, 1}, prev_pos{in->previous.t, W, H, 1};
    const TexU8 cur_nrm{in->current.normal_id, W, H}, prev_nrm{in->previous.normal_id, W, H};
    const TexU8 cur_blk{in->current.block_id, W, H}, prev_blk{in->previous.block_id, W, H};
    const Tex cur_sh{in->sh, W, H, 4}, cur_cocg{in->cocg, W, H, 2}, cur_luma{in->luma, W, H, 1}, cur_ao{in->ao_sky, W, H, 2};
    const Tex prev_sh{in->prev_sh, W, H, 4}, prev_cocg{in->prev_cocg, W, H, 2}, prev_util{in->prev_utility, W, H, 3}, prev_ao{in->prev_ao_sky, W, H, 2};
    float prev_vp[16];
    mat4_mul_mat(prm->prev_projection, prm->prev_view, prev_vp);  // u_PrevProjection * u_PrevView
    const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;
    static const float kWeights[5] = {3.0f / 32.0f, 3.0f / 32.0f, 9.0f / 64.0f, 3.0f / 32.0f, 3.0f / 32.0f};
    static const float kOff[5][2] = {{1, 0}, {0, 1}, {0, 0}, {-1, 0}, {0, -1}};
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = cam->row_begin; j < cam->row_end; ++j)
        for (int i = 0; i < W; ++i) {
            const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
            v3 base_p;
            float base_w;
            position_at(*cam, cur_pos, u, v, base_p, base_w);
            const v3 base_n = normal_of(cur_nrm.nearest(u, v));
            const v4 base_sh = cur_sh.linear4(u, v);
            const v2 base_cocg = cur_cocg.linear2(u, v);
            const v2 base_ao = cur_ao.linear2(u, v);
            // Reprojection :66-79
            const float wp[4] = {base_p.x, base_p.y, base_p.z, 1.0f};
            float pr[4];
            mat4_mul_vec(prev_vp, wp, pr);
            const float ru = (pr[0] / pr[3]) * 0.5f + 0.5f, rv = (pr[1] / pr[3]) * 0.5f + 0.5f;
            const float base_lum = cur_luma.linear1(u, v);
            const int base_block = std::min(std::max((int)std::floor(((float)cur_blk.nearest(u, v) / 255.0f) * 255.0f), 0), 127);
            // Jitter = ivec2((GradientNoise() - 0.5f) * 1.0f) is (0, 0) for every pixel: |noise - 0.5| < 1 truncates to zero
            const float dx = base_p.x - cam->inv_view[12], dy = base_p.y - cam->inv_view[13], dz = base_p.z - cam->inv_view[14];
            const float dist_player = std::sqrt(dot3(v3{-dx, -dy, -dz}, v3{-dx, -dy, -dz}));  // distance(a, b) = length(b - a)
            bool block_w = true, normal_w = true;
            float tol = 0.75f;
            if (dist_player < 4.0f) tol = 0.3f;
            else if (dist_player < 6.0f) tol = 0.65f;
            else if (dist_player < 8.0f) tol = 0.85f;
            else if (dist_player < 16.0f) tol = 1.414f;
            else if (dist_player < 32.0f) tol = 2.4f;
            else if (dist_player < 48.0f) { tol = 3.5f; block_w = false; }
            else if (dist_player < 64.0f) { tol = 4.2f; block_w = false; }
            else if (dist_player < 96.0f) { tol = 6.25f; block_w = false; normal_w = false; }
            else if (dist_player < 128.0f) { tol = 9.0f; block_w = false; normal_w = false; }
            else if (dist_player < 200.0f) { tol = 14.0f; block_w = false; normal_w = false; }
            float total_w = 0.0f, sum_spp = 0.0f, sum_moment = 0.0f, sum_lum = 0.0f;
            v4 sum_sh{0, 0, 0, 0};
            v2 sum_cocg{0, 0}, sum_ao{0, 0};
            int ok = 0;
            for (int k = 0; k < 5; ++k) {
                const float su = ru + (kOff[k][0] + 0.0f) * tsx, sv = rv + (kOff[k][1] + 0.0f) * tsy;
                const float b = 0.0035f;  // InThresholdedScreenSpace :101-105
                if (!(su < 1.0f - b && su > b && sv < 1.0f - b && sv > b)) continue;
                v3 pp;
                float pw;
                position_at(*cam, prev_pos, su, sv, pp, pw);
                const v3 pn = normal_of(prev_nrm.nearest(su, sv));
                const v3 diff{std::fabs(base_p.x - pp.x), std::fabs(base_p.y - pp.y), std::fabs(base_p.z - pp.z)};
                const float err = dot3(diff, diff);
                const int sblock = std::min(std::max((int)std::floor(((float)prev_blk.nearest(su, sv) / 255.0f) * 255.0f), 0), 127);
                bool valid = false;
                if (err < tol && ((pw < 0.0f) == (base_w < 0.0f))) {
                    valid = true;
                    if (normal_w && (pn.x != base_n.x || pn.y != base_n.y || pn.z != base_n.z)) valid = false;
                    if (block_w && base_block != sblock) valid = false;
                }
                if (valid) {
                    float ut[4];
                    prev_util.linear(su, sv, ut);
                    const float cw = kWeights[k];
                    sum_sh = sum_sh + prev_sh.linear4(su, sv) * cw;
                    sum_cocg = sum_cocg + prev_cocg.linear2(su, sv) * cw;
                    sum_spp += ut[0] * cw;
                    sum_moment += ut[1] * cw;
                    sum_lum += ut[2] * cw;
                    sum_ao = sum_ao + prev_ao.linear2(su, sv) * cw;
                    total_w += cw;
                    ok++;
                }
            }
            if (total_w > 0.001f) {
                sum_sh = sum_sh / total_w; sum_cocg = sum_cocg / total_w; sum_moment /= total_w; sum_spp /= total_w; sum_lum /= total_w;
                sum_ao = sum_ao / total_w;
            } else {
                ok = 0;
            }
            const float increment = prm->be_useful ? 1.0f : 0.0f;
            float spp_inc = sum_spp + increment;
            if (ok <= 0) spp_inc = 0.01f;
            float blend = std::fmax(1.0f / spp_inc, 0.05f);
            const float moment_factor = std::fmax(1.0f / spp_inc, 0.05f);
            if (!prm->be_useful) blend = 0.99f;
            float util_spp = spp_inc;
            if (ok <= 0) util_spp = 0.0f;
            const float util_moment = (1.0f - moment_factor) * sum_moment + moment_factor * (base_lum * base_lum);
            const float store_luma = mixf(sum_lum, base_lum, blend);
            v4 o_sh{mixf(sum_sh.x, base_sh.x, blend), mixf(sum_sh.y, base_sh.y, blend), mixf(sum_sh.z, base_sh.z, blend), mixf(sum_sh.w, base_sh.w, blend)};
            v2 o_cocg{mixf(sum_cocg.x, base_cocg.x, blend), mixf(sum_cocg.y, base_cocg.y, blend)};
            v2 o_ao{mixf(sum_ao.x, base_ao.x, blend), mixf(sum_ao.y, base_ao.y, blend)};
            if (ok <= 0) { o_sh = base_sh; o_cocg = base_cocg; o_ao = base_ao; }
            const size_t p = (size_t)j * W + i;
            if (out->sh) { out->sh[4 * p] = clampf(o_sh.x, -100.f, 100.f); out->sh[4 * p + 1] = clampf(o_sh.y, -100.f, 100.f);
                           out->sh[4 * p + 2] = clampf(o_sh.z, -100.f, 100.f); out->sh[4 * p + 3] = clampf(o_sh.w, -100.f, 100.f); }
            if (out->cocg) { out->cocg[2 * p] = clampf(o_cocg.x, -10.f, 100.f); out->cocg[2 * p + 1] = clampf(o_cocg.y, -10.f, 100.f); }
            if (out->ao_sky) { out->ao_sky[2 * p] = clampf(o_ao.x, 0.f, 1.f); out->ao_sky[2 * p + 1] = clampf(o_ao.y, 0.f, 1.f); }
            if (out->utility) { out->utility[3 * p] = clampf(util_spp, -150.f, 150.f); out->utility[3 * p + 1] = clampf(util_moment, -150.f, 150.f);
                                out->utility[3 * p + 2] = clampf(store_luma, -150.f, 150.f); }
        }
    return VXPT_OK;
}

// SVGF/VarianceEstimate.glsl main() :76-192 (Core/Pipeline.cpp:2434-2466)
int vxo_svgf_variance(const VxCamera* cam, const VxSvgfVarianceIn* in, const VxSvgfVarianceParams* prm, const VxSvgfVarianceOut* out) {
    const int W = cam->width, H = cam->height;
    const Tex pos{in->current.t, W, H, 1}, sh{in->sh, W, H, 4}, cocg{in->cocg, W, H, 2}, util{in->utility, W, H, 3};
    const TexU8 nrm{in->current.normal_id, W, H};
    const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;
    const bool aggr = prm->aggressive_disocclusion != 0;
    const float thresh = aggr ? 4.0f + 4.0f + 4.0f : 4.0f + 4.0f;
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = cam->row_begin; j < cam->row_end; ++j)
        for (int i = 0; i < W; ++i) {
            const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
            const float base_w = pos.linear1(u, v);  // GetPositionAt(...).w: the xyz of the position is never used
            const v3 base_n = normal_of(nrm.nearest(u, v));
            float bu[4];
            util.linear(u, v, bu);
            const v4 base_sh = sh.linear4(u, v);
            const v2 base_cocg = cocg.linear2(u, v);
            const float base_lum = sh_to_y(base_sh), frames = bu[0], base_moment = bu[1];
            v4 o_sh;
            v2 o_cocg;
            float variance;
            if (!prm->do_spatial) {
                o_sh = base_sh; o_cocg = base_cocg;
                variance = base_moment - base_lum * base_lum;  // :94-99 returns before the clamps
                const size_t p = (size_t)j * W + i;
                if (out->sh) { out->sh[4 * p] = o_sh.x; out->sh[4 * p + 1] = o_sh.y; out->sh[4 * p + 2] = o_sh.z; out->sh[4 * p + 3] = o_sh.w; }
                if (out->cocg) { out->cocg[2 * p] = o_cocg.x; out->cocg[2 * p + 1] = o_cocg.y; }
                if (out->variance) out->variance[p] = variance;
                continue;
            }
            if (frames < thresh) {
                const float color_phi = aggr ? 5.0f : 5.0f * 2.0f;
                const int K = aggr ? 4 : 1;
                float tw = 0.0f, tm = 0.0f, tl = 0.0f, tw2 = 0.0f;
                v4 tsh{0, 0, 0, 0};
                v2 tcc{0, 0};
                for (int x = -K; x <= K; ++x)
                    for (int y = -K; y <= K; ++y) {
                        const float su = u + (float)x * tsx, sv = v + (float)y * tsy;
                        if (!(su < 1.0f && su > 0.0f && sv < 1.0f && sv > 0.0f)) continue;
                        const float sw = pos.linear1(su, sv);
                        const v3 sn = normal_of(nrm.nearest(su, sv));
                        float sut[4];
                        util.linear(su, sv, sut);
                        const v4 ssh = sh.linear4(su, sv);
                        const v2 scc = cocg.linear2(su, sv);
                        const float slum = sh_to_y(ssh);
                        const float nw = pow_cr(std::fmax(dot3(base_n, sn), 0.0f), 16.0f);
                        const float dw = pow_cr(exp_cr(-std::fabs(sw - base_w)), 2.0f);
                        const float lw = std::fabs(slum - base_lum) / color_phi;
                        float w1 = exp_cr(-lw) * nw * dw;
                        float w2 = w1;
                        w1 = std::fmax(w1, 0.000000015f);
                        w2 = std::fmax(w2, 0.0000000015f);
                        tw += w1;
                        tm += sut[1] * w2;
                        tsh = tsh + ssh * w1;
                        tcc = tcc + scc * w1;
                        tl += slum * w2;
                        tw2 += w2;
                    }
                if (tw > 0.0f) { tm /= tw2; tl /= tw2; tcc = tcc / tw; tsh = tsh / tw; }
                o_sh = tsh; o_cocg = tcc;
                variance = (tm - tl * tl) * 3.0f;
            } else {
                o_sh = base_sh; o_cocg = base_cocg;
                variance = base_moment - base_lum * base_lum;
            }
            variance *= thresh / frames;
            const size_t p = (size_t)j * W + i;
            if (out->sh) { out->sh[4 * p] = clampf(o_sh.x, -100.f, 100.f); out->sh[4 * p + 1] = clampf(o_sh.y, -100.f, 100.f);
                           out->sh[4 * p + 2] = clampf(o_sh.z, -100.f, 100.f); out->sh[4 * p + 3] = clampf(o_sh.w, -100.f, 100.f); }
            if (out->cocg) { out->cocg[2 * p] = clampf(o_cocg.x, -10.f, 100.f); out->cocg[2 * p + 1] = clampf(o_cocg.y, -10.f, 100.f); }
            if (out->variance) out->variance[p] = clampf(variance, -1.0f, 50.0f);
        }
    return VXPT_OK;
}

// SVGF/SpatialFilter.glsl main() :193-354 (Core/Pipeline.cpp:2468-2596), one a-trous pass
int vxo_svgf_spatial(const VxCamera* cam, const VxSvgfSpatialIn* in, const VxSvgfSpatialParams* prm, const VxSvgfSpatialOut* out) {
    const int W = cam->width, H = cam->height;
    const Tex pos{in->current.t, W, H, 1}, sh{in->sh, W, H, 4}, cocg{in->cocg, W, H, 2}, var{in->variance, W, H, 1}, ao{in->ao_sky, W, H, 2},
        moment{in->temporal_utility, W, H, 3};
    const TexU8 nrm{in->current.normal_id, W, H};
    const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;   // 1 / u_Dimensions == 1 / textureSize(u_SH, 0)
    static const float kAtrous[3] = {1.0f, 2.0f / 3.0f, 1.0f / 6.0f};
    static const float kGauss[2] = {0.60283f, 0.198585f};
    const int step = prm->step;
    const bool filter_ao = step <= 4;
    const bool filter_sky = step <= 6 || filter_ao;
    const int K = prm->large_kernel ? 2 : 1;
    const float add_scale = mixf(1.0f, 2.4f, prm->resolution_scale);
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = cam->row_begin; j < cam->row_end; ++j)
        for (int i = 0; i < W; ++i) {
            const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
            const float jit_f = (gradient_noise(i, j, prm->time) - 0.5f) * ((float)step * 0.8f);
            const int jit = (int)jit_f;  // ivec2(float): both components
            const float base_depth = pos.linear1(u, v);
            const v3 base_n = normal_of(nrm.nearest(u, v));
            const v4 base_sh = sh.linear4(u, v);
            const v2 base_cocg = cocg.linear2(u, v);
            const float base_lum = sh_to_y(base_sh);
            // GaussianVariance :98-132
            float base_var = 0.0f, vsum = 0.0f, ksum = 0.0f;
            for (int x = -1; x <= 1; ++x)
                for (int y = -1; y <= 1; ++y) {
                    const float su = u + (float)x * tsx, sv = v + (float)y * tsy;
                    if (!(su > 0.0f && su < 1.0f && sv > 0.0f && sv < 1.0f)) continue;
                    const float kv = kGauss[std::abs(x)] * kGauss[std::abs(y)];
                    const float V = var.linear1(su, sv);
                    if (x == 0 && y == 0) base_var = V;
                    vsum += V * kv;
                    ksum += kv;
                }
            const float var_est = vsum / std::fmax(ksum, 0.01f);
            const v2 base_ao = ao.linear2(u, v);
            const size_t p = (size_t)j * W + i;
            if (!prm->do_spatial) {  // :217-223 returns before the clamps
                if (out->sh) { out->sh[4 * p] = base_sh.x; out->sh[4 * p + 1] = base_sh.y; out->sh[4 * p + 2] = base_sh.z; out->sh[4 * p + 3] = base_sh.w; }
                if (out->cocg) { out->cocg[2 * p] = base_cocg.x; out->cocg[2 * p + 1] = base_cocg.y; }
                if (out->variance) out->variance[p] = base_var;
                if (out->ao_sky) { out->ao_sky[2 * p] = base_ao.x; out->ao_sky[2 * p + 1] = base_ao.y; }
                continue;
            }
            v4 tsh = base_sh;
            v2 tcc = base_cocg, tao = base_ao;
            float tw = 1.0f, tvar = base_var, taow = 1.0f;
            float mt[4];
            moment.linear(u, v, mt);
            const bool strong = mt[0] <= 8.0f && prm->aggressive_disocclusion && step <= 8;
            float curve = 0.0f;
            if (var_est < 0.01f) curve = 128.0f;
            else if (var_est < 0.025f) curve = 112.0f;
            else if (var_est < 0.05f) curve = 96.0f;
            else if (var_est < 0.075f) curve = 84.0f;
            else if (var_est < 0.1f) curve = 70.0f;
            float tweaked = var_est;
            if (var_est < 0.1f) {  // TweakVariance :184-190
                const float f = clampf(var_est, 0.0f, 1.0f);
                tweaked = f * pow_cr(1.0f - f, curve + 6.0f);
            }
            float phi = std::sqrt(std::fmax(0.0f, 0.000001f + tweaked));
            phi /= std::fmax(prm->color_phi_bias, 0.1f);
            for (int x = -K; x <= K; ++x)
                for (int y = -K; y <= K; ++y) {
                    if (x == 0 && y == 0) continue;
                    const float su = u + ((((float)x * (float)step) * add_scale) + ((float)jit * 0.5f)) * tsx;
                    const float sv = v + ((((float)y * (float)step) * add_scale) + ((float)jit * 0.5f)) * tsy;
                    if (!(su > 0.0f && su < 1.0f && sv > 0.0f && sv < 1.0f)) continue;
                    const float sdepth = pos.linear1(su, sv);
                    const float ddiff = std::fabs(sdepth - base_depth);
                    const v3 sn = normal_of(nrm.nearest(su, sv));
                    if ((base_depth < 0.0f) == (ddiff < 0.0f)) {
                        const v4 ssh = sh.linear4(su, sv);
                        const v2 scc = cocg.linear2(su, sv);
                        const float slum = sh_to_y(ssh);
                        const float svar = var.linear1(su, sv);
                        float nw = pow_cr(std::fmax(dot3(base_n, sn), 0.0f), 32.0f);
                        nw = clampf(nw, 0.001f, 1.0f);
                        const float lw = std::fabs(slum - base_lum) / phi;
                        const float dw = clampf(pow_cr(exp_cr(-std::fmax(ddiff, 0.00001f)), 2.0f), 0.0001f, 1.0f);
                        float w = strong ? (nw * dw) : (exp_cr(-lw) * nw * dw);
                        w = clampf(w, 0.001f, 1.0f);
                        const float xw = kAtrous[std::abs(x)], yw = kAtrous[std::abs(y)];
                        w = (xw * yw) * w;
                        w = std::fmax(w, 0.00000001f);
                        tsh = tsh + ssh * w;
                        tcc = tcc + scc * w;
                        tvar += (w * w) * svar;
                        tw += w;
                        if (filter_sky || filter_ao) {
                            const float aw = clampf((xw * yw) * nw * dw, 0.000001f, 1.0f);
                            const v2 s = ao.linear2(su, sv);
                            tao.x += s.x * aw;
                            tao.y += s.y * aw;
                            taow += aw;
                        }
                    }
                }
            tsh = tsh / tw;
            tcc = tcc / tw;
            tvar /= (tw * tw);
            tao = tao / taow;
            if (!filter_ao) tao.x = base_ao.x;
            if (out->sh) { out->sh[4 * p] = clampf(tsh.x, -100.f, 100.f); out->sh[4 * p + 1] = clampf(tsh.y, -100.f, 100.f);
                           out->sh[4 * p + 2] = clampf(tsh.z, -100.f, 100.f); out->sh[4 * p + 3] = clampf(tsh.w, -100.f, 100.f); }
            if (out->cocg) { out->cocg[2 * p] = clampf(tcc.x, -10.f, 100.f); out->cocg[2 * p + 1] = clampf(tcc.y, -10.f, 100.f); }
            if (out->variance) out->variance[p] = clampf(tvar, -1.0f, 50.0f);
            if (out->ao_sky) { out->ao_sky[2 * p] = clampf(tao.x, 0.f, 1.f); out->ao_sky[2 * p + 1] = clampf(tao.y, 0.f, 1.f); }
        }
    return VXPT_OK;
}

// Spatial3x3Initial.glsl main() :102-176 (Core/Pipeline.cpp:2288-2330), the pre-temporal 3x3 pass
int vxo_svgf_initial(const VxCamera* cam, const VxSvgfInitialIn* in, const VxSvgfInitialOut* out) {
    const int W = cam->width, H = cam->height;
    const Tex pos{in->current.t, W, H, 1}, sh{in->sh, W, H, 4}, cocg{in->cocg, W, H, 2}, luma{in->luma, W, H, 1}, ao{in->ao_sky, W, H, 2};
    const TexU8 nrm{in->current.normal_id, W, H};
    const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;
    static const float kAtrous[3] = {1.0f, 2.0f / 3.0f, 1.0f / 6.0f};
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = cam->row_begin; j < cam->row_end; ++j)
        for (int i = 0; i < W; ++i) {
            const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
            v3 bp;
            float bw;
            position_at(*cam, pos, u, v, bp, bw);
            const v3 bn = normal_of(nrm.nearest(u, v));
            const v4 bsh = sh.linear4(u, v);
            const v2 bcc = cocg.linear2(u, v), bao = ao.linear2(u, v);
            const float blum = sh_to_y(bsh);
            v4 tsh = bsh;
            v2 tcc = bcc, tao = bao;
            float tw = 1.0f, taw = 1.0f;
            for (int x = -1; x <= 1; ++x)
                for (int y = -1; y <= 1; ++y) {
                    if (x == 0 && y == 0) continue;
                    const float su = u + ((float)x * 1.0f) * tsx, sv = v + ((float)y * 1.0f) * tsy;
                    if (!(su > 0.0f && su < 1.0f && sv > 0.0f && sv < 1.0f)) continue;
                    v3 sp;
                    float sw;
                    position_at(*cam, pos, su, sv, sp, sw);
                    const v3 df{std::fabs(sp.x - bp.x), std::fabs(sp.y - bp.y), std::fabs(sp.z - bp.z)};
                    if (dot3(df, df) < 1.0f) {
                        const v4 ssh = sh.linear4(su, sv);
                        const v2 scc = cocg.linear2(su, sv);
                        const v3 sn = normal_of(nrm.nearest(su, sv));
                        const float nw = pow_cr(std::fmax(dot3(bn, sn), 0.0f), 16.0f);
                        const float lw = std::fabs(sh_to_y(ssh) - blum) / 4.0f;
                        float w = exp_cr(-lw - nw);       // sic: the normal term is subtracted in the exponent
                        w = std::fmax(w, 0.01f);
                        w = (kAtrous[std::abs(x)] * kAtrous[std::abs(y)]) * w;
                        w = clampf(std::fmax(w, 0.01f), 0.0f, 1.0f);
                        tsh = tsh + ssh * w;
                        tcc = tcc + scc * w;
                        tw += w;
                        tao = tao + ao.linear2(su, sv) * w;
                        taw += w;
                    }
                }
            tw = std::fmax(tw, 0.01f);
            tsh = tsh / tw;
            tcc = tcc / tw;
            tao = tao / std::fmax(taw, 0.01f);
            const size_t p = (size_t)j * W + i;
            if (out->sh) { out->sh[4 * p] = tsh.x; out->sh[4 * p + 1] = tsh.y; out->sh[4 * p + 2] = tsh.z; out->sh[4 * p + 3] = tsh.w; }
            if (out->cocg) { out->cocg[2 * p] = tcc.x; out->cocg[2 * p + 1] = tcc.y; }
            if (out->ao_sky) { out->ao_sky[2 * p] = tao.x; out->ao_sky[2 * p + 1] = tao.y; }
            if (out->luma) out->luma[p] = luma.linear1(u, v);
        }
    return VXPT_OK;
}

// ShadowTemporalFilter.glsl main() :201-298 with u_ShadowTemporal = true (Core/Pipeline.cpp:2854-2903).  The colour attachments hold one
// channel (R8 shadow, R16F frame count): only the .x of the shader's vec3 / vec4 arithmetic reaches an output, and only .x is restated.
int vxo_shadow_temporal(const VxCamera* cam, const VxShadowTemporalIn* in, const VxShadowTemporalParams* prm, const VxShadowTemporalOut* out) {
    const int W = cam->width, H = cam->height;
    std::vector<float> cur_f((size_t)W * H);
    for (size_t k = 0; k < cur_f.size(); ++k) cur_f[k] = (float)in->shadow[k];  // o_Shadow 0 / 1 in an R8 attachment reads back as 0.0 / 1.0
    const Tex cur{cur_f.data(), W, H, 1}, pos{in->current.t, W, H, 1}, ppos{in->previous.t, W, H, 1}, trans{in->transversal, W, H, 1},
        prev{in->prev_shadow, W, H, 1}, frames{in->prev_frames, W, H, 1};
    const TexU8 nrm{in->current.normal_id, W, H};
    float prev_vp[16];
    mat4_mul_mat(prm->prev_projection, prm->prev_view, prev_vp);
    const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;
    const float unit_diagonal = std::sqrt(2.0f);
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = cam->row_begin; j < cam->row_end; ++j)
        for (int i = 0; i < W; ++i) {
            const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
            const size_t p = (size_t)j * W + i;
            v3 cp;
            float cw;
            position_at(*cam, pos, u, v, cp, cw);
            float o_color, o_frames;
            if (cw > 0.0f) {
                const float wp[4] = {cp.x, cp.y, cp.z, 1.0f};
                float pr[4];
                mat4_mul_vec(prev_vp, wp, pr);
                const float ru = (pr[0] / pr[3]) * 0.5f + 0.5f, rv = (pr[1] / pr[3]) * 0.5f + 0.5f;
                const float tr = trans.linear1(u, v) * 100.0f;
                // GetShadowSpatial :109-155
                float cur_color;
                if (tr <= unit_diagonal * 2.0f) {
                    cur_color = 1.0f;
                } else {
                    float total = cur.linear1(u, v);
                    const float base = total;
                    float weight = 1.0f;
                    const int bn = nrm.nearest(u, v);
                    for (int x = -1; x <= 1; ++x)
                        for (int y = -1; y <= 1; ++y) {
                            if (x == 0 && y == 0) continue;
                            const float su = u + (float)x * tsx, sv = v + (float)y * tsy;
                            const float b = 0.03f;
                            if (!(su > b && su < 1.0f - b && sv > b && sv < 1.0f - b)) continue;
                            const float sd = pos.linear1(su, sv);
                            const int sn = nrm.nearest(su, sv);
                            if (std::min(sn, 6) == std::min(bn, 6) && std::fabs(sd - cw) < 1.0f) {
                                const float smp = cur.linear1(su, sv);
                                float wa = clampf(1.0f - clampf(std::fabs(smp - base) / 3.0f, 0.0f, 1.0f), 0.0f, 1.0f);
                                wa = clampf(pow_cr(wa, 7.0f), 0.000001f, 1.0f);
                                total += smp * wa;
                                weight += wa;
                            }
                        }
                    cur_color = total / weight;
                }
                const float prev_orig = prev.linear1(ru, rv);
                float prev_color = prev_orig;
                if (tr < 1.414f * 3.0f) {  // ClipShadow :168-184, clipAABB :157-166
                    float mn = 100.0f, mx = -100.0f;
                    static const float off[5][2] = {{-1, 0}, {1, 0}, {0, 0}, {0, -1}, {0, 1}};
                    for (int s2 = 0; s2 < 5; ++s2) {
                        const float smp = cur.linear1(u + off[s2][0] * tsx, v + off[s2][1] * tsy);
                        mn = std::fmin(smp, mn);
                        mx = std::fmax(smp, mx);
                    }
                    const float lo = mn - 0.025f, hi = mx + 0.025f;
                    const float pc = 0.5f * (hi + lo), ec = 0.5f * (hi - lo);
                    const float vc = prev_orig - pc;
                    const float au = std::fabs(vc / ec);
                    const float denom = std::fmax(au, std::fmax(au, au));
                    prev_color = denom > 1.0f ? pc + vc / denom : prev_orig;
                }
                v3 pp;
                float pw;
                position_at(*cam, ppos, ru, rv, pp, pw);
                const float bias = 0.005f;
                const bool reject = !(ru > 0.0f + bias && ru < 1.0f - bias && rv > 0.0f + bias && rv < 1.0f - bias);
                if (!reject) {
                    const v3 dd{cp.x - pp.x, cp.y - pp.y, cp.z - pp.z};
                    const float d = std::sqrt(dot3(dd, dd));
                    const float cc = clampf(cur_color, 0.0f, 1.0f), pc2 = clampf(prev_color, 0.0f, 1.0f);
                    const float vx = (u - ru) * (float)W, vy = (v - rv) * (float)H;
                    const float clip_error = std::fabs(prev_orig - pc2);
                    const float inc = clip_error < 0.2f ? 1.0f : 0.6f;
                    const float fetch = frames.linear1(ru, rv);
                    const float incd = fetch + inc;
                    float blend = clampf((1.0f - (1.0f / incd)) * 1.2f, 0.01f, 0.97f);
                    const float vrf = clampf(exp_cr(-std::sqrt(vx * vx + vy * vy)) * 0.8f + 0.6f, 0.00000001f, 1.0f);
                    blend *= vrf;
                    float depth_rej = 1.0f;
                    if (d > 0.4f) {
                        depth_rej = pow_cr(exp_cr(-d), 48.0f);
                        blend *= clampf(depth_rej, 0.0f, 1.0f);
                    }
                    o_color = mixf(cc, pc2, clampf(blend, 0.0f, 0.97f));
                    const float mult = depth_rej * vrf;
                    o_frames = fetch + clampf(mult * 1.1f, 0.0f, 1.0f);
                    if (mult < 0.1f) o_frames = 0.0f;
                    else if (mult <= 0.2f + 0.001f) o_frames = 2.0f;
                    else if (mult <= 0.3f + 0.001f) o_frames = 3.25f;
                } else {
                    o_color = cur_color;
                    o_frames = 0.0f;
                }
            } else {
                o_color = cur.linear1(u, v);
                o_frames = 0.0f;
            }
            if (out->shadow) out->shadow[p] = o_color;
            if (out->frames) out->frames[p] = clampf(o_frames, 0.0f, 256.0f);
        }
    return VXPT_OK;
}

// ShadowFilter.glsl ShadowSpatial :68-160 (Core/Pipeline.cpp:2905-2944)
int vxo_shadow_filter(const VxCamera* cam, const VxShadowFilterIn* in, const VxShadowFilterParams* prm, float* out) {
    const int W = cam->width, H = cam->height;
    const Tex inp{in->shadow, W, H, 1}, pos{in->current.t, W, H, 1}, trans{in->transversal, W, H, 1}, frames{in->frames, W, H, 1};
    const TexU8 nrm{in->current.normal_id, W, H};
    const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;
    const float cutoff = std::sqrt(2.0f);
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = cam->row_begin; j < cam->row_end; ++j)
        for (int i = 0; i < W; ++i) {
            const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
            const size_t p = (size_t)j * W + i;
            const float fr = frames.linear1(u, v);
            const bool luma_weight = fr > 7.5f;
            const float center_w = pos.linear1(u, v);
            const v3 cn = normal_of(nrm.nearest(u, v));
            const float center = inp.linear1(u, v);
            const float tr = trans.linear1(u, v) * 100.0f;
            if ((tr > 0.0f && tr < cutoff) || center_w < 0.0f) { out[p] = center; continue; }
            const bool reduced = tr < cutoff * 1.414f;
            const int K = reduced ? 1 : 3;
            float scale = 1.0f;
            if (tr > 6.0f) scale = 2.0f;
            if (tr > 16.0f) scale = 2.4f;
            if (tr > 32.0f) scale = 2.6f;
            const float ct = clampf(tr, 0.0f, 10.0f);
            float var_est = mixf(20.0f, 6.0f, ct / 10.0f) + (tr < 6.0f ? 5.0f : 2.0f);
            var_est = clampf(var_est - 1.75f, 0.0000001f, 64.0f);
            float luma_mixer = 1.0f;
            if (!luma_weight) luma_mixer = mixf(0.1f, 0.5f, fr / 7.5f);
            float tw = 0.0f, ts = 0.0f;
            for (int x = -K; x <= K; ++x)
                for (int y = -K; y <= K; ++y) {
                    const float su = u + ((((float)x * tsx) * 1.2f) * scale) * prm->filter_scale;
                    const float sv = v + ((((float)y * tsy) * 1.2f) * scale) * prm->filter_scale;
                    const float sd = pos.linear1(su, sv);
                    const v3 sn = normal_of(nrm.nearest(su, sv));
                    const float dw = pow_cr(exp_cr(-(std::fabs(center_w - sd))), 3.0f);
                    const float nw = pow_cr(std::fmax(dot3(cn, sn), 0.000000001f), 32.0f);
                    const float sa = inp.linear1(su, sv);
                    const float le = clampf(1.0f - clampf(std::fabs(sa - center) / 3.0f, 0.0f, 1.0f), 0.0f, 1.0f);
                    float w = 1.0f;
                    w *= clampf(pow_cr(le, var_est * luma_mixer * 0.9f), 0.0f, 1.0f);
                    w *= dw;
                    w *= nw;
                    w = clampf(w, 0.000000001f, 1.0f);
                    ts += sa * w;
                    tw += w;
                }
            out[p] = ts / std::fmax(tw, 0.01f);
        }
    return VXPT_OK;
}

}  // extern "C"
