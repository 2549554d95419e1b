// denoise.cu — SVGF diffuse denoiser and sun-shadow filters: Core/Shaders/SVGF/TemporalFilter.glsl,
// VarianceEstimate.glsl, SpatialFilter.glsl (Core/Pipeline.cpp:2335-2596) and ShadowTemporalFilter.glsl, ShadowFilter.glsl (:2854-2944),
// the passes that consume the diffuse-GI and shadow planes.
//
// One thread per output pixel on the trace passes' 8x4-pixel warp tiles, so the taps of a warp (neighbouring texels, or the same
// a-trous offset for every lane) fall into a few 128-byte lines.  Every read goes through tex_linear / tex_nearest below: the pinned
// texture() of include/vxpt.h (GL_REPEAT, OpenGL 4.3 section 8.14.2 bilinear in fp32) — also for taps that land on a texel centre, because
// u * w - 0.5 is not always an integer in fp32 and the parity contract is bit-exactness with the reference's shaders.
// HBM per 1080p pass: temporal 189 MB, variance 141 MB, one a-trous pass 151 MB of planes; the 4-tap gathers around them hit L1 / L2.
// Compiled with -fmad=false; exp / pow are the pinned correctly rounded fp32 values; fminf / fmaxf return the non-NaN operand.
#include <cmath>

#include "gi_device.cuh"

namespace vxpt {

// The pinned transcendentals are double-precision library calls (about a hundred FP64 instructions each), and the filters evaluate several
// per tap.  Most of them have exact short cuts that return the SAME fp32 value (checked bit for bit against the oracle on the host):
//   pow(x, y) with x == 1 is 1 for every y, with x == 0 and y > 0 is 0: face-normal dot products are 0 or 1, luminance-error terms are
//   exactly 1 wherever a tap equals the centre;
//   pow(y, 2): y * y is exact in double (48 significant bits), so the correctly rounded square is the fp32 product;
//   pow(y, 3): (y * y) * y in double carries one rounding at 53 bits before the one to fp32;
//   exp(-0) = 1.
//
// What is left goes through a short double-precision evaluation with a rounding test.  exp(x) = 2^(k / 64) * e^r with k = rint(64 x / ln 2)
// and |r| <= ln 2 / 128: a 64-entry table and a degree-5 polynomial, relative error below 2^-51 (table entry, polynomial, product: each 2^-53;
// truncation r^6 / 720 < 2^-54), i.e. within 4 units of the double's last place.  Rounding that double to fp32 gives the correctly rounded
// value — and the value (float)exp((double)x) gives, whose own error is below one unit — unless the double lies within a few units of the
// midpoint of two floats (mantissa bits 28..0 = 0x10000000): those calls, about one in 10^7, take the library function.  pow(x, y) for
// 0 < x < 1 is the same evaluation at y * log(x): the library logarithm is good to a unit in the last place and the product is below 104 in
// magnitude, so the argument carries an absolute error under 2^-45 and the rounding test is as many units wider.
__device__ const double kExp2Tab[64] = {
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0, 0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0,
    0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0, 0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0,
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0, 0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0,
    0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0, 0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0,
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0, 0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0,
    0x1.6247eb03a5585p+0, 0x1.6623882552225p+0, 0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0,
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0, 0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0,
    0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0, 0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0,
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0, 0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0,
    0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0, 0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0,
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0};
// e^a for -87 < a < 0 (a normal float comes out), rounded to fp32; false when the rounding test asks for the library function.
// (r03z: the constants as literals; read from a __constant__ array instead, DFMA takes them as operands without the two moves a literal
// costs, but the compiler hoists the loads out of the filter loops and the a-trous kernel goes from 64 to 78 registers: 0.288 -> 0.313 ms.)
__device__ __forceinline__ bool exp_short(double a, int margin, float* out) {
    const double magic = 0x1.8p52;                           // 1.5 * 2^52: rounds the sum to an integer
    const double t = fma(a, 0x1.71547652b82fep+6, magic);    // 64 / ln 2
    const int k = (int)__double_as_longlong(t);              // rint(64 a / ln 2): the low word of the magic sum
    const double kd = t - magic;
    double r = fma(kd, -0x1.62e42fee00000p-7, a);            // ln 2 / 64 in two pieces, the first short enough for kd * hi to be exact
    r = fma(kd, -0x1.a39ef35793c76p-39, r);
    double p = fma(r, 0x1.1111111111111p-7, 0x1.5555555555555p-5);  // 1 / 120, 1 / 24
    p = fma(r, p, 0x1.5555555555555p-3);                     // 1 / 6
    p = fma(r, p, 0.5);
    p = fma(r, p, 1.0);
    p = fma(r, p, 1.0);
    const long long sb = __double_as_longlong(kExp2Tab[k & 63]) + ((long long)(k >> 6) << 52);  // 2^(k / 64): exponent field >= 1023 - 126
    const double y = __longlong_as_double(sb) * p;
    const int low = (int)__double_as_longlong(y) & 0x1fffffff;  // the 29 mantissa bits fp32 drops
    *out = (float)y;
    return (unsigned)(low - 0x10000000 + margin) > (unsigned)(2 * margin);
}
__device__ __forceinline__ float exp_cr(float x) {
    if (x == 0.0f) return 1.0f;
    if (x <= -104.0f) return 0.0f;  // e^-104 < 2^-150: rounds to zero
    float y;
    if (x < 0.0f && x > -87.0f && exp_short((double)x, 16, &y)) return y;
    return (float)exp((double)x);
}
__device__ __forceinline__ float pow_lt1_cr(float x, float y) {  // the library's pow for every input; short evaluation for 0 < x < 1, 0 < y
    if (x > 0.0f && x < 1.0f && y > 0.0f && y < 1.0e6f) {
        const double a = (double)y * log((double)x);
        float r;
        if (a > -87.0 && a < 0.0 && exp_short(a, 2048, &r)) return r;
    }
    return pow_cr(x, y);
}
__device__ __forceinline__ float pow01_cr(float x, float y) {  // y > 0
    if (x == 1.0f) return 1.0f;
    if (x == 0.0f) return 0.0f;
    return pow_lt1_cr(x, y);
}
// pow(max(dot(n_a, n_b), floor), y) for two face-normal ids (normal_from_id(id, 1): six axis normals, every other id (1, 1, 1)): the dot
// product is exactly -1, 0, 1 or 3, so the power is `at_floor` (the value for the clamped -1 / 0: pow(0, y) = 0, pow(1e-9, 32) underflows
// to 0), 1, or 3^y — 3^16 and 3^32 are integers below 2^53, so the pinned (float)pow((double)3, y) is the literal rounded once.
__device__ __forceinline__ float normal_weight(int a, int b, float at_floor, float pow3) {
    const bool a6 = a > 5, b6 = b > 5;
    if (a6 && b6) return pow3;
    const int axis = a6 ? b : a;                                     // the id that is an axis normal when exactly one is not
    const bool one = (a6 || b6) ? ((0x25 >> axis) & 1) : (a == b);  // ids 0, 2, 5 point along +z, +y, +x
    return one ? 1.0f : at_floor;
}
#define VXPT_POW3_16 ((float)43046721.0)
#define VXPT_POW3_32 ((float)1853020188851841.0)
__device__ __forceinline__ float sq_cr(float y) { return y * y; }
__device__ __forceinline__ float cube_cr(float y) { return (float)(((double)y * (double)y) * (double)y); }
struct Bilinear {  // the four texels and two weights of one GL_LINEAR tap
    int o00, o10, o01, o11;
    float fx, fy;
};
__device__ __forceinline__ Bilinear bilinear_at(int w, int h, float u, float v) {
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float x0 = floorf(x), y0 = floorf(y);
    const int ix = (int)x0, iy = (int)y0;
    int i0, i1, j0, j1;
    if ((unsigned)ix + (unsigned)w < (unsigned)(3 * w - 1) && (unsigned)iy + (unsigned)h < (unsigned)(3 * h - 1)) {  // within a period of the texture: one branch per tap
        i0 = wrap_near(ix, w), i1 = wrap_near(ix + 1, w), j0 = wrap_near(iy, h), j1 = wrap_near(iy + 1, h);
    } else {
        i0 = wrap_repeat(ix, w), i1 = wrap_repeat(ix + 1, w), j0 = wrap_repeat(iy, h), j1 = wrap_repeat(iy + 1, h);
    }
    return Bilinear{j0 * w + i0, j0 * w + i1, j1 * w + i0, j1 * w + i1, x - x0, y - y0};
}
// the same tap for 0 < u < 1 and 0 < v < 1 (every filter loop tests that before it samples): u * w <= w in fp32, so floor(u * w - 0.5) lies in
// [-1, w - 1] and only the texel left of column 0 / right of column w - 1 wraps
__device__ __forceinline__ Bilinear bilinear_in01(int w, int h, float u, float v) {
    const float x = u * (float)w - 0.5f, y = v * (float)h - 0.5f;
    const float x0 = floorf(x), y0 = floorf(y);
    const int ix = (int)x0, iy = (int)y0;
    const int i0 = ix < 0 ? ix + w : ix, i1 = ix + 1 >= w ? ix + 1 - w : ix + 1;
    const int j0 = (iy < 0 ? iy + h : iy) * w, j1 = (iy + 1 >= h ? iy + 1 - h : iy + 1) * w;
    return Bilinear{j0 + i0, j0 + i1, j1 + i0, j1 + i1, x - x0, y - y0};
}
__device__ __forceinline__ float blend(const Bilinear& b, float t00, float t10, float t01, float t11) {
    return (t00 * (1.0f - b.fx) + t10 * b.fx) * (1.0f - b.fy) + (t01 * (1.0f - b.fx) + t11 * b.fx) * b.fy;
}
__device__ __forceinline__ float tex1(const float* d, const Bilinear& b) { return blend(b, d[b.o00], d[b.o10], d[b.o01], d[b.o11]); }
__device__ __forceinline__ float2 tex2(const float* d, const Bilinear& b) {
    const float2* p = reinterpret_cast<const float2*>(d);
    const float2 a = p[b.o00], c = p[b.o10], e = p[b.o01], f = p[b.o11];
    return make_float2(blend(b, a.x, c.x, e.x, f.x), blend(b, a.y, c.y, e.y, f.y));
}
__device__ __forceinline__ V3 tex3(const float* d, const Bilinear& b) {
    const float *a = d + 3 * b.o00, *c = d + 3 * b.o10, *e = d + 3 * b.o01, *f = d + 3 * b.o11;
    return mk3(blend(b, a[0], c[0], e[0], f[0]), blend(b, a[1], c[1], e[1], f[1]), blend(b, a[2], c[2], e[2], f[2]));
}
__device__ __forceinline__ float4 tex4(const float* d, const Bilinear& b) {
    const float4* p = reinterpret_cast<const float4*>(d);
    const float4 a = p[b.o00], c = p[b.o10], e = p[b.o01], f = p[b.o11];
    return make_float4(blend(b, a.x, c.x, e.x, f.x), blend(b, a.y, c.y, e.y, f.y), blend(b, a.z, c.z, e.z, f.z), blend(b, a.w, c.w, e.w, f.w));
}
__device__ __forceinline__ int tex_nearest_u8(const uint8_t* d, int w, int h, float u, float v) {
    return d[wrap_repeat((int)floorf(v * (float)h), h) * w + wrap_repeat((int)floorf(u * (float)w), w)];
}
// 0 < u < 1, 0 < v < 1: floor(u * w) lies in [0, w] (u * w can round up to w), so only index w wraps
__device__ __forceinline__ int tex_nearest_u8_in01(const uint8_t* d, int w, int h, float u, float v) {
    const int ix = (int)floorf(u * (float)w), iy = (int)floorf(v * (float)h);
    return d[(iy >= h ? iy - h : iy) * w + (ix >= w ? ix - w : ix)];
}
__device__ __forceinline__ float sh_to_y(float4 sh) { return fmaxf(0.0f, 3.544905f * sh.w); }
__device__ __forceinline__ float4 operator+(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 operator*(float4 a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
__device__ __forceinline__ float4 operator/(float4 a, float s) { return make_float4(a.x / s, a.y / s, a.z / s, a.w / s); }
__device__ __forceinline__ float2 operator+(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 operator*(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 operator/(float2 a, float s) { return make_float2(a.x / s, a.y / s); }
__device__ __forceinline__ float4 clamp4(float4 a, float lo, float hi) {
    return make_float4(clampf(a.x, lo, hi), clampf(a.y, lo, hi), clampf(a.z, lo, hi), clampf(a.w, lo, hi));
}
__device__ __forceinline__ float2 clamp2(float2 a, float lo, float hi) { return make_float2(clampf(a.x, lo, hi), clampf(a.y, lo, hi)); }

// __launch_bounds__ per kernel: threads, and optionally the resident CTAs per SM the register allocation aims for (build.py
// -DVXPT_DN_..._BOUNDS=256,4 to experiment; r03z / r04a)
#ifndef VXPT_DN_INITIAL_BOUNDS
#define VXPT_DN_INITIAL_BOUNDS 256
#endif
#ifndef VXPT_DN_TEMPORAL_BOUNDS
#define VXPT_DN_TEMPORAL_BOUNDS 256, 4
#endif
#ifndef VXPT_DN_VARIANCE_BOUNDS
#define VXPT_DN_VARIANCE_BOUNDS 256, 4
#endif
#ifndef VXPT_DN_SPATIAL_BOUNDS
#define VXPT_DN_SPATIAL_BOUNDS 256
#endif
#ifndef VXPT_DN_SHADOW_BOUNDS
#define VXPT_DN_SHADOW_BOUNDS 256, 5
#endif
struct SvgfPlanes {  // device pointers of one call; unused members are null
    const float *t, *prev_t;
    const uint8_t *nid, *prev_nid, *bid, *prev_bid;
    const float *sh, *cocg, *luma, *ao, *utility, *variance;
    const float *prev_sh, *prev_cocg, *prev_utility, *prev_ao;
    float *o_sh, *o_cocg, *o_utility, *o_variance, *o_ao;
};

// ============================================================================================= pre-temporal 3x3 pass
// Spatial3x3Initial.glsl main() :102-176
__global__ void __launch_bounds__(VXPT_DN_INITIAL_BOUNDS) svgf_initial_kernel(const __grid_constant__ CameraDev cam, const SvgfPlanes pl) {
    int i, j, prow;
    if (!thread_pixel(cam, i, j, prow)) return;
    const int W = cam.width, H = cam.height;
    const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
    const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;
    const Bilinear bc = bilinear_in01(W, H, u, v);
    const V3 origin = ray_origin(cam);
    const V3 bp = origin + normalize3(ray_direction_at(cam, u, v)) * tex1(pl.t, bc);
    const int bn = tex_nearest_u8_in01(pl.nid, W, H, u, v);
    const float4 bsh = tex4(pl.sh, bc);
    const float2 bcc = tex2(pl.cocg, bc), bao = tex2(pl.ao, bc);
    const float blum = sh_to_y(bsh);
    float4 tsh = bsh;
    float2 tcc = bcc, tao = bao;
    float tw = 1.0f, taw = 1.0f;
#pragma unroll 1
    for (int x = -1; x <= 1; ++x)
#pragma unroll 1
        for (int y = -1; y <= 1; ++y) {
            if (x == 0 && y == 0) continue;
            const float su = u + ((float)x * 1.0f) * tsx, sv = v + ((float)y * 1.0f) * tsy;
            if (!(su > 0.0f && su < 1.0f && sv > 0.0f && sv < 1.0f)) continue;
            const Bilinear bs = bilinear_in01(W, H, su, sv);
            const V3 sp = origin + normalize3(ray_direction_at(cam, su, sv)) * tex1(pl.t, bs);
            const V3 df = mk3(fabsf(sp.x - bp.x), fabsf(sp.y - bp.y), fabsf(sp.z - bp.z));
            if (!(dot3(df, df) < 1.0f)) continue;
            const float4 ssh = tex4(pl.sh, bs);
            const float nw = normal_weight(bn, tex_nearest_u8_in01(pl.nid, W, H, su, sv), 0.0f, VXPT_POW3_16);  // pow(max(dot(bn, sn), 0), 16)
            const float lw = fabsf(sh_to_y(ssh) - blum) / 4.0f;
            float w = fmaxf(exp_cr(-lw - nw), 0.01f);  // sic: the normal term is subtracted in the exponent
            const float xw = x == 0 ? 1.0f : 2.0f / 3.0f, yw = y == 0 ? 1.0f : 2.0f / 3.0f;
            w = clampf(fmaxf((xw * yw) * w, 0.01f), 0.0f, 1.0f);
            tsh = tsh + ssh * w;
            tcc = tcc + tex2(pl.cocg, bs) * w;
            tw += w;
            tao = tao + tex2(pl.ao, bs) * w;
            taw += w;
        }
    tw = fmaxf(tw, 0.01f);
    const size_t px = (size_t)prow * W + i;
    if (pl.o_sh) reinterpret_cast<float4*>(pl.o_sh)[px] = tsh / tw;
    if (pl.o_cocg) reinterpret_cast<float2*>(pl.o_cocg)[px] = tcc / tw;
    if (pl.o_ao) reinterpret_cast<float2*>(pl.o_ao)[px] = tao / fmaxf(taw, 0.01f);
    if (pl.o_utility) pl.o_utility[px] = tex1(pl.luma, bc);
}

// ============================================================================================= temporal accumulation
struct TemporalDev {
    float prev_vp[16];  // u_PrevProjection * u_PrevView, multiplied on the host in glm's order
    int be_useful;
};
// SVGF/TemporalFilter.glsl main() :137-362
__global__ void __launch_bounds__(VXPT_DN_TEMPORAL_BOUNDS) svgf_temporal_kernel(const __grid_constant__ CameraDev cam, const __grid_constant__ TemporalDev P,
                                                            const SvgfPlanes pl) {
    int i, j, prow;
    if (!thread_pixel(cam, i, j, prow)) return;
    const int W = cam.width, H = cam.height;
    const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
    const Bilinear bc = bilinear_in01(W, H, u, v);
    const V3 origin = ray_origin(cam);
    const float base_w = tex1(pl.t, bc);
    const V3 base_p = origin + normalize3(ray_direction_at(cam, u, v)) * base_w;  // GetPositionAt :81-85
    const int base_nid = tex_nearest_u8_in01(pl.nid, W, H, u, v);
    const float4 base_sh = tex4(pl.sh, bc);
    const float2 base_cocg = tex2(pl.cocg, bc), base_ao = tex2(pl.ao, bc);
    // Reprojection :58-72
    const float* M = P.prev_vp;
    const float px4 = (M[0] * base_p.x + M[4] * base_p.y) + (M[8] * base_p.z + M[12] * 1.0f);
    const float py4 = (M[1] * base_p.x + M[5] * base_p.y) + (M[9] * base_p.z + M[13] * 1.0f);
    const float pw4 = (M[3] * base_p.x + M[7] * base_p.y) + (M[11] * base_p.z + M[15] * 1.0f);
    const float ru = (px4 / pw4) * 0.5f + 0.5f, rv = (py4 / pw4) * 0.5f + 0.5f;
    const float base_lum = tex1(pl.luma, bc);
    const int base_block = min(tex_nearest_u8_in01(pl.bid, W, H, u, v), 127);
    // (the shader's tap jitter ivec2((GradientNoise() - 0.5) * 1.0) truncates to zero for every pixel)
    const V3 to_player = origin - base_p;
    const float dist_player = sqrtf(dot3(to_player, to_player));
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
    const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;
    float total_w = 0.0f, sum_spp = 0.0f, sum_moment = 0.0f, sum_lum = 0.0f;
    float4 sum_sh = make_float4(0.f, 0.f, 0.f, 0.f);
    float2 sum_cocg = make_float2(0.f, 0.f), sum_ao = make_float2(0.f, 0.f);
    int ok = 0;
#pragma unroll 1
    for (int k = 0; k < 5; ++k) {  // Offsets[5] = (1,0) (0,1) (0,0) (-1,0) (0,-1), Weights 3/32 3/32 9/64 3/32 3/32
        const float ox = k == 0 ? 1.0f : (k == 3 ? -1.0f : 0.0f), oy = k == 1 ? 1.0f : (k == 4 ? -1.0f : 0.0f);
        const float cw = k == 2 ? 9.0f / 64.0f : 3.0f / 32.0f;
        const float su = ru + (ox + 0.0f) * tsx, sv = rv + (oy + 0.0f) * tsy;
        const float b = 0.0035f;
        if (!(su < 1.0f - b && su > b && sv < 1.0f - b && sv > b)) continue;
        const Bilinear bs = bilinear_in01(W, H, su, sv);
        const float pw = tex1(pl.prev_t, bs);
        const V3 pp = origin + normalize3(ray_direction_at(cam, su, sv)) * pw;
        const V3 diff = mk3(fabsf(base_p.x - pp.x), fabsf(base_p.y - pp.y), fabsf(base_p.z - pp.z));
        const float err = dot3(diff, diff);
        bool valid = err < tol && ((pw < 0.0f) == (base_w < 0.0f));
        if (valid && normal_w) {  // PreviousNormalAt != BaseNormal: ids above 5 all decode to (1,1,1)
            const int pn = tex_nearest_u8_in01(pl.prev_nid, W, H, su, sv);
            valid = min(pn, 6) == min(base_nid, 6);
        }
        if (valid && block_w) valid = base_block == min(tex_nearest_u8_in01(pl.prev_bid, W, H, su, sv), 127);
        if (valid) {
            const V3 ut = tex3(pl.prev_utility, bs);
            sum_sh = sum_sh + tex4(pl.prev_sh, bs) * cw;
            sum_cocg = sum_cocg + tex2(pl.prev_cocg, bs) * cw;
            sum_spp += ut.x * cw;
            sum_moment += ut.y * cw;
            sum_lum += ut.z * cw;
            sum_ao = sum_ao + tex2(pl.prev_ao, bs) * cw;
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
    float spp_inc = sum_spp + (P.be_useful ? 1.0f : 0.0f);
    if (ok <= 0) spp_inc = 0.01f;
    float blend_f = fmaxf(1.0f / spp_inc, 0.05f);
    const float moment_f = blend_f;
    if (!P.be_useful) blend_f = 0.99f;
    const float util_spp = ok <= 0 ? 0.0f : spp_inc;
    const float util_moment = (1.0f - moment_f) * sum_moment + moment_f * (base_lum * base_lum);
    const float store_luma = mixf(sum_lum, base_lum, blend_f);
    float4 o_sh = make_float4(mixf(sum_sh.x, base_sh.x, blend_f), mixf(sum_sh.y, base_sh.y, blend_f), mixf(sum_sh.z, base_sh.z, blend_f),
                              mixf(sum_sh.w, base_sh.w, blend_f));
    float2 o_cocg = make_float2(mixf(sum_cocg.x, base_cocg.x, blend_f), mixf(sum_cocg.y, base_cocg.y, blend_f));
    float2 o_ao = make_float2(mixf(sum_ao.x, base_ao.x, blend_f), mixf(sum_ao.y, base_ao.y, blend_f));
    if (ok <= 0) { o_sh = base_sh; o_cocg = base_cocg; o_ao = base_ao; }
    const size_t px = (size_t)prow * W + i;
    if (pl.o_sh) reinterpret_cast<float4*>(pl.o_sh)[px] = clamp4(o_sh, -100.0f, 100.0f);
    if (pl.o_cocg) reinterpret_cast<float2*>(pl.o_cocg)[px] = clamp2(o_cocg, -10.0f, 100.0f);
    if (pl.o_ao) reinterpret_cast<float2*>(pl.o_ao)[px] = clamp2(o_ao, 0.0f, 1.0f);
    if (pl.o_utility) {
        pl.o_utility[3 * px] = clampf(util_spp, -150.0f, 150.0f);
        pl.o_utility[3 * px + 1] = clampf(util_moment, -150.0f, 150.0f);
        pl.o_utility[3 * px + 2] = clampf(store_luma, -150.0f, 150.0f);
    }
}

// ============================================================================================= variance estimate
struct VarianceDev {
    int do_spatial, aggressive;
};
// SVGF/VarianceEstimate.glsl main() :76-192
__global__ void __launch_bounds__(VXPT_DN_VARIANCE_BOUNDS) svgf_variance_kernel(const __grid_constant__ CameraDev cam, const VarianceDev P, const SvgfPlanes pl) {
    int i, j, prow;
    if (!thread_pixel(cam, i, j, prow)) return;
    const int W = cam.width, H = cam.height;
    const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
    const Bilinear bc = bilinear_in01(W, H, u, v);
    const float base_w = tex1(pl.t, bc);
    const int base_n = tex_nearest_u8_in01(pl.nid, W, H, u, v);
    const V3 bu = tex3(pl.utility, bc);
    const float4 base_sh = tex4(pl.sh, bc);
    const float2 base_cocg = tex2(pl.cocg, bc);
    const float base_lum = sh_to_y(base_sh), frames = bu.x, base_moment = bu.y;
    const size_t px = (size_t)prow * W + i;
    if (!P.do_spatial) {  // :94-99 returns before the clamps
        if (pl.o_sh) reinterpret_cast<float4*>(pl.o_sh)[px] = base_sh;
        if (pl.o_cocg) reinterpret_cast<float2*>(pl.o_cocg)[px] = base_cocg;
        if (pl.o_variance) pl.o_variance[px] = base_moment - base_lum * base_lum;
        return;
    }
    const float thresh = P.aggressive ? 4.0f + 4.0f + 4.0f : 4.0f + 4.0f;
    float4 o_sh = base_sh;
    float2 o_cocg = base_cocg;
    float variance;
    if (frames < thresh) {
        const float color_phi = P.aggressive ? 5.0f : 5.0f * 2.0f;
        const int K = P.aggressive ? 4 : 1;
        const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;
        float tw = 0.0f, tm = 0.0f, tl = 0.0f, tw2 = 0.0f;
        float4 tsh = make_float4(0.f, 0.f, 0.f, 0.f);
        float2 tcc = make_float2(0.f, 0.f);
#pragma unroll 1
        for (int x = -K; x <= K; ++x)
#pragma unroll 1
            for (int y = -K; y <= K; ++y) {
                const float su = u + (float)x * tsx, sv = v + (float)y * tsy;
                if (!(su < 1.0f && su > 0.0f && sv < 1.0f && sv > 0.0f)) continue;
                const Bilinear bs = bilinear_in01(W, H, su, sv);
                const float sw = tex1(pl.t, bs);
                const int sn = tex_nearest_u8_in01(pl.nid, W, H, su, sv);
                const float smoment = tex3(pl.utility, bs).y;
                const float4 ssh = tex4(pl.sh, bs);
                const float2 scc = tex2(pl.cocg, bs);
                const float slum = sh_to_y(ssh);
                const float nw = normal_weight(base_n, sn, 0.0f, VXPT_POW3_16);  // pow(max(dot(base_n, sn), 0), 16)
                const float dw = sq_cr(exp_cr(-fabsf(sw - base_w)));
                const float lw = fabsf(slum - base_lum) / color_phi;
                const float w0 = exp_cr(-lw) * nw * dw;
                const float w1 = fmaxf(w0, 0.000000015f), w2 = fmaxf(w0, 0.0000000015f);
                tw += w1;
                tm += smoment * w2;
                tsh = tsh + ssh * w1;
                tcc = tcc + scc * w1;
                tl += slum * w2;
                tw2 += w2;
            }
        if (tw > 0.0f) { tm /= tw2; tl /= tw2; tcc = tcc / tw; tsh = tsh / tw; }
        o_sh = tsh;
        o_cocg = tcc;
        variance = (tm - tl * tl) * 3.0f;
    } else {
        variance = base_moment - base_lum * base_lum;
    }
    variance *= thresh / frames;
    if (pl.o_sh) reinterpret_cast<float4*>(pl.o_sh)[px] = clamp4(o_sh, -100.0f, 100.0f);
    if (pl.o_cocg) reinterpret_cast<float2*>(pl.o_cocg)[px] = clamp2(o_cocg, -10.0f, 100.0f);
    if (pl.o_variance) pl.o_variance[px] = clampf(variance, -1.0f, 50.0f);
}

// ============================================================================================= a-trous pass
struct SpatialDev {
    int step, large_kernel, do_spatial, aggressive;
    float color_phi_bias, resolution_scale;
    float noise_shift;  // mod(u_Time * 100.493850275f, 500.0f), computed on the host in fp32
};
// SVGF/SpatialFilter.glsl main() :193-354
__global__ void __launch_bounds__(VXPT_DN_SPATIAL_BOUNDS) svgf_spatial_kernel(const __grid_constant__ CameraDev cam, const SpatialDev P, const SvgfPlanes pl) {
    int i, j, prow;
    if (!thread_pixel(cam, i, j, prow)) return;
    const int W = cam.width, H = cam.height;
    const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
    const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;
    const int step = P.step;
    // GradientNoise :177-182 -> Jitter :206
    const float cx = ((float)i + 0.5f) + P.noise_shift, cy = ((float)j + 0.5f) + P.noise_shift;
    const float noise = fractf(52.9829189f * fractf(0.06711056f * cx + 0.00583715f * cy));
    const int jit = (int)((noise - 0.5f) * ((float)step * 0.8f));
    const Bilinear bc = bilinear_in01(W, H, u, v);
    const float base_depth = tex1(pl.t, bc);
    const int base_n = tex_nearest_u8_in01(pl.nid, W, H, u, v);
    const float4 base_sh = tex4(pl.sh, bc);
    const float2 base_cocg = tex2(pl.cocg, bc);
    const float base_lum = sh_to_y(base_sh);
    // GaussianVariance :98-132
    float base_var = 0.0f, vsum = 0.0f, ksum = 0.0f;
#pragma unroll
    for (int x = -1; x <= 1; ++x)
#pragma unroll
        for (int y = -1; y <= 1; ++y) {
            const float su = u + (float)x * tsx, sv = v + (float)y * tsy;
            if (!(su > 0.0f && su < 1.0f && sv > 0.0f && sv < 1.0f)) continue;
            const float kv = (x == 0 ? 0.60283f : 0.198585f) * (y == 0 ? 0.60283f : 0.198585f);
            const float V = tex1(pl.variance, bilinear_in01(W, H, su, sv));
            if (x == 0 && y == 0) base_var = V;
            vsum += V * kv;
            ksum += kv;
        }
    const float var_est = vsum / fmaxf(ksum, 0.01f);
    const float2 base_ao = tex2(pl.ao, bc);
    const size_t px = (size_t)prow * W + i;
    if (!P.do_spatial) {  // :217-223 returns before the clamps
        if (pl.o_sh) reinterpret_cast<float4*>(pl.o_sh)[px] = base_sh;
        if (pl.o_cocg) reinterpret_cast<float2*>(pl.o_cocg)[px] = base_cocg;
        if (pl.o_variance) pl.o_variance[px] = base_var;
        if (pl.o_ao) reinterpret_cast<float2*>(pl.o_ao)[px] = base_ao;
        return;
    }
    const bool filter_ao = step <= 4, filter_sky = step <= 6 || filter_ao;
    float4 tsh = base_sh;
    float2 tcc = base_cocg, tao = base_ao;
    float tw = 1.0f, tvar = base_var, taow = 1.0f;
    const bool strong = tex3(pl.utility, bc).x <= 8.0f && P.aggressive && step <= 8;
    float curve = 0.0f;
    if (var_est < 0.01f) curve = 128.0f;
    else if (var_est < 0.025f) curve = 112.0f;
    else if (var_est < 0.05f) curve = 96.0f;
    else if (var_est < 0.075f) curve = 84.0f;
    else if (var_est < 0.1f) curve = 70.0f;
    float tweaked = var_est;
    if (var_est < 0.1f) {  // TweakVariance :184-190
        const float f = clampf(var_est, 0.0f, 1.0f);
        tweaked = f * pow01_cr(1.0f - f, curve + 6.0f);
    }
    float phi = sqrtf(fmaxf(0.0f, 0.000001f + tweaked));
    phi /= fmaxf(P.color_phi_bias, 0.1f);
    const int K = P.large_kernel ? 2 : 1;
    const float add_scale = mixf(1.0f, 2.4f, P.resolution_scale);
#pragma unroll 1
    for (int x = -K; x <= K; ++x)
#pragma unroll 1
        for (int y = -K; y <= K; ++y) {
            if (x == 0 && y == 0) continue;
            const float su = u + ((((float)x * (float)step) * add_scale) + ((float)jit * 0.5f)) * tsx;
            const float sv = v + ((((float)y * (float)step) * add_scale) + ((float)jit * 0.5f)) * tsy;
            if (!(su > 0.0f && su < 1.0f && sv > 0.0f && sv < 1.0f)) continue;
            const Bilinear bs = bilinear_in01(W, H, su, sv);
            const float ddiff = fabsf(tex1(pl.t, bs) - base_depth);
            if ((base_depth < 0.0f) != (ddiff < 0.0f)) continue;  // :262
            const int sn = tex_nearest_u8_in01(pl.nid, W, H, su, sv);
            const float4 ssh = tex4(pl.sh, bs);
            const float2 scc = tex2(pl.cocg, bs);
            const float slum = sh_to_y(ssh);
            const float svar = tex1(pl.variance, bs);
            const float nw = clampf(normal_weight(base_n, sn, 0.0f, VXPT_POW3_32), 0.001f, 1.0f);  // pow(max(dot(base_n, sn), 0), 32)
            const float lw = fabsf(slum - base_lum) / phi;
            const float dw = clampf(sq_cr(exp_cr(-fmaxf(ddiff, 0.00001f))), 0.0001f, 1.0f);
            float w = strong ? (nw * dw) : (exp_cr(-lw) * nw * dw);
            w = clampf(w, 0.001f, 1.0f);
            const float xw = x == 0 ? 1.0f : ((x == 1 || x == -1) ? 2.0f / 3.0f : 1.0f / 6.0f);
            const float yw = y == 0 ? 1.0f : ((y == 1 || y == -1) ? 2.0f / 3.0f : 1.0f / 6.0f);
            w = fmaxf((xw * yw) * w, 0.00000001f);
            tsh = tsh + ssh * w;
            tcc = tcc + scc * w;
            tvar += (w * w) * svar;
            tw += w;
            if (filter_sky || filter_ao) {
                const float aw = clampf((xw * yw) * nw * dw, 0.000001f, 1.0f);
                const float2 s = tex2(pl.ao, bs);
                tao.x += s.x * aw;
                tao.y += s.y * aw;
                taow += aw;
            }
        }
    tsh = tsh / tw;
    tcc = tcc / tw;
    tvar /= (tw * tw);
    tao = tao / taow;
    if (!filter_ao) tao.x = base_ao.x;
    if (pl.o_sh) reinterpret_cast<float4*>(pl.o_sh)[px] = clamp4(tsh, -100.0f, 100.0f);
    if (pl.o_cocg) reinterpret_cast<float2*>(pl.o_cocg)[px] = clamp2(tcc, -10.0f, 100.0f);
    if (pl.o_variance) pl.o_variance[px] = clampf(tvar, -1.0f, 50.0f);
    if (pl.o_ao) reinterpret_cast<float2*>(pl.o_ao)[px] = clamp2(tao, 0.0f, 1.0f);
}

// ============================================================================================= sun-shadow filters
struct ShadowFilterPlanes {
    const float *t, *prev_t;
    const uint8_t *nid, *shadow_u8;
    const float *shadow, *transversal, *prev_shadow, *frames;
    float *o_shadow, *o_frames;
};
// u_CurrentColorTexture of the temporal pass is the raw R8 shadow plane (0 / 1): the bilinear tap on bytes
__device__ __forceinline__ float tex1_u8(const uint8_t* d, const Bilinear& b) {
    return blend(b, (float)d[b.o00], (float)d[b.o10], (float)d[b.o01], (float)d[b.o11]);
}
// ShadowTemporalFilter.glsl main() :201-298 with u_ShadowTemporal = true.  The attachments are single-channel: only the .x of the shader's
// vector arithmetic reaches an output.
__global__ void __launch_bounds__(VXPT_DN_SHADOW_BOUNDS) shadow_temporal_kernel(const __grid_constant__ CameraDev cam, const __grid_constant__ TemporalDev P,
                                                              const ShadowFilterPlanes pl) {
    int i, j, prow;
    if (!thread_pixel(cam, i, j, prow)) return;
    const int W = cam.width, H = cam.height;
    const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
    const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;
    const Bilinear bc = bilinear_in01(W, H, u, v);
    const V3 origin = ray_origin(cam);
    const float cw = tex1(pl.t, bc);
    const size_t px = (size_t)prow * W + i;
    float o_color, o_frames = 0.0f;
    if (cw > 0.0f) {
        const V3 cp = origin + normalize3(ray_direction_at(cam, u, v)) * cw;
        const float* M = P.prev_vp;
        const float px4 = (M[0] * cp.x + M[4] * cp.y) + (M[8] * cp.z + M[12] * 1.0f);
        const float py4 = (M[1] * cp.x + M[5] * cp.y) + (M[9] * cp.z + M[13] * 1.0f);
        const float pw4 = (M[3] * cp.x + M[7] * cp.y) + (M[11] * cp.z + M[15] * 1.0f);
        const float ru = (px4 / pw4) * 0.5f + 0.5f, rv = (py4 / pw4) * 0.5f + 0.5f;
        const float tr = tex1(pl.transversal, bc) * 100.0f;
        float cur_color;
        if (tr <= sqrtf(2.0f) * 2.0f) {  // GetShadowSpatial :109-155
            cur_color = 1.0f;
        } else {
            float total = tex1_u8(pl.shadow_u8, bc);
            const float base = total;
            float weight = 1.0f;
            const int bn = min(tex_nearest_u8_in01(pl.nid, W, H, u, v), 6);
#pragma unroll 1
            for (int x = -1; x <= 1; ++x)
#pragma unroll 1
                for (int y = -1; y <= 1; ++y) {
                    if (x == 0 && y == 0) continue;
                    const float su = u + (float)x * tsx, sv = v + (float)y * tsy;
                    const float b = 0.03f;
                    if (!(su > b && su < 1.0f - b && sv > b && sv < 1.0f - b)) continue;
                    const Bilinear bs = bilinear_in01(W, H, su, sv);
                    const float sd = tex1(pl.t, bs);
                    if (min(tex_nearest_u8_in01(pl.nid, W, H, su, sv), 6) == bn && fabsf(sd - cw) < 1.0f) {
                        const float smp = tex1_u8(pl.shadow_u8, bs);
                        float wa = clampf(1.0f - clampf(fabsf(smp - base) / 3.0f, 0.0f, 1.0f), 0.0f, 1.0f);
                        wa = clampf(pow01_cr(wa, 7.0f), 0.000001f, 1.0f);
                        total += smp * wa;
                        weight += wa;
                    }
                }
            cur_color = total / weight;
        }
        const Bilinear br = bilinear_at(W, H, ru, rv);
        const float prev_orig = tex1(pl.prev_shadow, br);
        float prev_color = prev_orig;
        if (tr < 1.414f * 3.0f) {  // ClipShadow :168-184, clipAABB :157-166
            float mn = 100.0f, mx = -100.0f;
#pragma unroll
            for (int s2 = 0; s2 < 5; ++s2) {  // ShadowClipOffsets (-1,0) (1,0) (0,0) (0,-1) (0,1)
                const float ox = s2 == 0 ? -1.0f : (s2 == 1 ? 1.0f : 0.0f), oy = s2 == 3 ? -1.0f : (s2 == 4 ? 1.0f : 0.0f);
                const float smp = tex1_u8(pl.shadow_u8, bilinear_at(W, H, u + ox * tsx, v + oy * tsy));
                mn = fminf(smp, mn);
                mx = fmaxf(smp, mx);
            }
            const float lo = mn - 0.025f, hi = mx + 0.025f;
            const float pc = 0.5f * (hi + lo), ec = 0.5f * (hi - lo);
            const float vc = prev_orig - pc;
            const float denom = fabsf(vc / ec);
            prev_color = denom > 1.0f ? pc + vc / denom : prev_orig;
        }
        const float pw = tex1(pl.prev_t, br);
        const V3 pp = origin + normalize3(ray_direction_at(cam, ru, rv)) * pw;
        const float bias = 0.005f;
        if (ru > 0.0f + bias && ru < 1.0f - bias && rv > 0.0f + bias && rv < 1.0f - bias) {
            const V3 dd = cp - pp;
            const float d = sqrtf(dot3(dd, dd));
            const float cc = clampf(cur_color, 0.0f, 1.0f), pc2 = clampf(prev_color, 0.0f, 1.0f);
            const float vx = (u - ru) * (float)W, vy = (v - rv) * (float)H;
            const float inc = fabsf(prev_orig - pc2) < 0.2f ? 1.0f : 0.6f;
            const float fetch = tex1(pl.frames, br);
            float blend_f = clampf((1.0f - (1.0f / (fetch + inc))) * 1.2f, 0.01f, 0.97f);
            const float vrf = clampf(exp_cr(-sqrtf(vx * vx + vy * vy)) * 0.8f + 0.6f, 0.00000001f, 1.0f);
            blend_f *= vrf;
            float depth_rej = 1.0f;
            if (d > 0.4f) {
                depth_rej = pow01_cr(exp_cr(-d), 48.0f);
                blend_f *= clampf(depth_rej, 0.0f, 1.0f);
            }
            o_color = mixf(cc, pc2, clampf(blend_f, 0.0f, 0.97f));
            const float mult = depth_rej * vrf;
            o_frames = fetch + clampf(mult * 1.1f, 0.0f, 1.0f);
            if (mult < 0.1f) o_frames = 0.0f;
            else if (mult <= 0.2f + 0.001f) o_frames = 2.0f;
            else if (mult <= 0.3f + 0.001f) o_frames = 3.25f;
        } else {
            o_color = cur_color;
        }
    } else {
        o_color = tex1_u8(pl.shadow_u8, bc);
    }
    if (pl.o_shadow) pl.o_shadow[px] = o_color;
    if (pl.o_frames) pl.o_frames[px] = clampf(o_frames, 0.0f, 256.0f);
}

// ShadowFilter.glsl ShadowSpatial :68-160
__global__ void __launch_bounds__(VXPT_DN_SHADOW_BOUNDS) shadow_filter_kernel(const __grid_constant__ CameraDev cam, const float filter_scale, const ShadowFilterPlanes pl) {
    int i, j, prow;
    if (!thread_pixel(cam, i, j, prow)) return;
    const int W = cam.width, H = cam.height;
    const float u = ((float)i + 0.5f) / (float)W, v = ((float)j + 0.5f) / (float)H;
    const float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;
    const Bilinear bc = bilinear_in01(W, H, u, v);
    const size_t px = (size_t)prow * W + i;
    const float fr = tex1(pl.frames, bc);
    const float center_w = tex1(pl.t, bc);
    const int cn = tex_nearest_u8_in01(pl.nid, W, H, u, v);
    const float center = tex1(pl.shadow, bc);
    const float tr = tex1(pl.transversal, bc) * 100.0f;
    const float cutoff = sqrtf(2.0f);
    if ((tr > 0.0f && tr < cutoff) || center_w < 0.0f) {
        pl.o_shadow[px] = center;
        return;
    }
    const int K = tr < cutoff * 1.414f ? 1 : 3;
    float scale = 1.0f;
    if (tr > 6.0f) scale = 2.0f;
    if (tr > 16.0f) scale = 2.4f;
    if (tr > 32.0f) scale = 2.6f;
    float var_est = mixf(20.0f, 6.0f, clampf(tr, 0.0f, 10.0f) / 10.0f) + (tr < 6.0f ? 5.0f : 2.0f);
    var_est = clampf(var_est - 1.75f, 0.0000001f, 64.0f);
    const float luma_mixer = fr > 7.5f ? 1.0f : mixf(0.1f, 0.5f, fr / 7.5f);
    const float luma_exp = var_est * luma_mixer * 0.9f;
    float tw = 0.0f, ts = 0.0f;
#pragma unroll 1
    for (int x = -K; x <= K; ++x)
#pragma unroll 1
        for (int y = -K; y <= K; ++y) {
            const float su = u + ((((float)x * tsx) * 1.2f) * scale) * filter_scale;
            const float sv = v + ((((float)y * tsy) * 1.2f) * scale) * filter_scale;
            const Bilinear bs = bilinear_at(W, H, su, sv);
            const float sd = tex1(pl.t, bs);
            const int sn = tex_nearest_u8(pl.nid, W, H, su, sv);
            const float dw = cube_cr(exp_cr(-(fabsf(center_w - sd))));
            const float nw = normal_weight(cn, sn, 0.0f, VXPT_POW3_32);  // pow(max(dot(cn, sn), 1e-9), 32): (1e-9)^32 rounds to 0
            const float sa = tex1(pl.shadow, bs);
            const float le = clampf(1.0f - clampf(fabsf(sa - center) / 3.0f, 0.0f, 1.0f), 0.0f, 1.0f);
            float w = 1.0f;
            w *= clampf(le == 1.0f ? 1.0f : pow_lt1_cr(le, luma_exp), 0.0f, 1.0f);
            w *= dw;
            w *= nw;
            w = clampf(w, 0.000000001f, 1.0f);
            ts += sa * w;
            tw += w;
        }
    pl.o_shadow[px] = ts / fmaxf(tw, 0.01f);
}

// ============================================================================================= launchers
static CameraDev svgf_camera(const VxCamera& cam) {
    CameraDev cd;
    for (int k = 0; k < 16; ++k) { cd.inv_view[k] = cam.inv_view[k]; cd.inv_proj[k] = cam.inv_proj[k]; }
    cd.width = cam.width; cd.height = cam.height; cd.row_begin = cam.row_begin; cd.row_end = cam.row_end;
    cd.il_n = 0; cd.il_rank = 0; cd.il_band = 1;
    return cd;
}
static dim3 svgf_grid(const VxCamera& cam) { return dim3((cam.width + 31) / 32, (cam.row_end - cam.row_begin + 7) / 8); }

int launch_svgf_initial(vxpt_ctx* c, const VxCamera& cam, const VxSvgfInitialIn& in, const VxSvgfInitialOut& out) {
    SvgfPlanes pl{};
    pl.t = in.current.t; pl.nid = in.current.normal_id;
    pl.sh = in.sh; pl.cocg = in.cocg; pl.luma = in.luma; pl.ao = in.ao_sky;
    pl.o_sh = out.sh; pl.o_cocg = out.cocg; pl.o_utility = out.luma; pl.o_ao = out.ao_sky;
    VX_LAUNCH(svgf_initial_kernel, svgf_grid(cam), 256, c->stream, svgf_camera(cam), pl);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

int launch_svgf_temporal(vxpt_ctx* c, const VxCamera& cam, const VxSvgfTemporalIn& in, const VxSvgfTemporalParams& p, const VxSvgfTemporalOut& out) {
    TemporalDev d;
    const float *A = p.prev_projection, *B = p.prev_view;  // glm mat4 * mat4: column j = ((A0*b0j + A1*b1j) + A2*b2j) + A3*b3j
    for (int jc = 0; jc < 4; ++jc)
        for (int r = 0; r < 4; ++r)
            d.prev_vp[4 * jc + r] = ((A[0 + r] * B[4 * jc + 0] + A[4 + r] * B[4 * jc + 1]) + A[8 + r] * B[4 * jc + 2]) + A[12 + r] * B[4 * jc + 3];
    d.be_useful = p.be_useful;
    SvgfPlanes pl{};
    pl.t = in.current.t; pl.nid = in.current.normal_id; pl.bid = in.current.block_id;
    pl.prev_t = in.previous.t; pl.prev_nid = in.previous.normal_id; pl.prev_bid = in.previous.block_id;
    pl.sh = in.sh; pl.cocg = in.cocg; pl.luma = in.luma; pl.ao = in.ao_sky;
    pl.prev_sh = in.prev_sh; pl.prev_cocg = in.prev_cocg; pl.prev_utility = in.prev_utility; pl.prev_ao = in.prev_ao_sky;
    pl.o_sh = out.sh; pl.o_cocg = out.cocg; pl.o_utility = out.utility; pl.o_ao = out.ao_sky;
    VX_LAUNCH(svgf_temporal_kernel, svgf_grid(cam), 256, c->stream, svgf_camera(cam), d, pl);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

int launch_svgf_variance(vxpt_ctx* c, const VxCamera& cam, const VxSvgfVarianceIn& in, const VxSvgfVarianceParams& p, const VxSvgfVarianceOut& out) {
    const VarianceDev d{p.do_spatial, p.aggressive_disocclusion};
    SvgfPlanes pl{};
    pl.t = in.current.t; pl.nid = in.current.normal_id;
    pl.sh = in.sh; pl.cocg = in.cocg; pl.utility = in.utility;
    pl.o_sh = out.sh; pl.o_cocg = out.cocg; pl.o_variance = out.variance;
    VX_LAUNCH(svgf_variance_kernel, svgf_grid(cam), 256, c->stream, svgf_camera(cam), d, pl);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

int launch_svgf_spatial(vxpt_ctx* c, const VxCamera& cam, const VxSvgfSpatialIn& in, const VxSvgfSpatialParams& p, const VxSvgfSpatialOut& out) {
    SpatialDev d;
    d.step = p.step; d.large_kernel = p.large_kernel; d.do_spatial = p.do_spatial; d.aggressive = p.aggressive_disocclusion;
    d.color_phi_bias = p.color_phi_bias; d.resolution_scale = p.resolution_scale;
    const float m = p.time * 100.493850275f;
    d.noise_shift = m - 500.0f * std::floor(m / 500.0f);  // mod(x, y) = x - y * floor(x / y)
    SvgfPlanes pl{};
    pl.t = in.current.t; pl.nid = in.current.normal_id;
    pl.sh = in.sh; pl.cocg = in.cocg; pl.variance = in.variance; pl.ao = in.ao_sky; pl.utility = in.temporal_utility;
    pl.o_sh = out.sh; pl.o_cocg = out.cocg; pl.o_variance = out.variance; pl.o_ao = out.ao_sky;
    VX_LAUNCH(svgf_spatial_kernel, svgf_grid(cam), 256, c->stream, svgf_camera(cam), d, pl);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

int launch_shadow_temporal(vxpt_ctx* c, const VxCamera& cam, const VxShadowTemporalIn& in, const VxShadowTemporalParams& p, const VxShadowTemporalOut& out) {
    TemporalDev d;
    const float *A = p.prev_projection, *B = p.prev_view;
    for (int jc = 0; jc < 4; ++jc)
        for (int r = 0; r < 4; ++r)
            d.prev_vp[4 * jc + r] = ((A[0 + r] * B[4 * jc + 0] + A[4 + r] * B[4 * jc + 1]) + A[8 + r] * B[4 * jc + 2]) + A[12 + r] * B[4 * jc + 3];
    d.be_useful = 1;
    ShadowFilterPlanes pl{};
    pl.t = in.current.t; pl.nid = in.current.normal_id; pl.prev_t = in.previous.t;
    pl.shadow_u8 = in.shadow; pl.transversal = in.transversal; pl.prev_shadow = in.prev_shadow; pl.frames = in.prev_frames;
    pl.o_shadow = out.shadow; pl.o_frames = out.frames;
    VX_LAUNCH(shadow_temporal_kernel, svgf_grid(cam), 256, c->stream, svgf_camera(cam), d, pl);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

int launch_shadow_filter(vxpt_ctx* c, const VxCamera& cam, const VxShadowFilterIn& in, const VxShadowFilterParams& p, float* out) {
    ShadowFilterPlanes pl{};
    pl.t = in.current.t; pl.nid = in.current.normal_id;
    pl.shadow = in.shadow; pl.transversal = in.transversal; pl.frames = in.frames;
    pl.o_shadow = out;
    VX_LAUNCH(shadow_filter_kernel, svgf_grid(cam), 256, c->stream, svgf_camera(cam), p.filter_scale, pl);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

}  // namespace vxpt
