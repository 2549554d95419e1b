// gbuffer.cu — G-buffer material pass: Core/Shaders/GenerateGBuffer.glsl main() :347-441 as drawn by
// Core/Pipeline.cpp:2066-2136, in the v1 parity profile of include/vxpt.h ; relief parallax mapping (u_POM) and the animated lava path are separate instantiations.
//
// One thread per pixel, the 8x4-pixel warp tiles of the trace passes.  The shader takes screen-space derivatives of the surface UV
// (GetUVDerivative :443-461) to pick the mip level: a derivative is a difference inside the pixel's 2x2 quad, so every thread
// re-derives the UV of its two quad neighbours from their G-buffer texels (three short ray set-ups instead of a shuffle: the same
// source then runs one "thread" after another in tests/host_shadow).  HBM traffic: 6 B read + 44 B written per pixel; the texel
// gathers (two mip levels of three RGBA8 arrays) hit L1/L2 — the pass is bound by its plane writes.
// Compiled with -fmad=false; log2 is the pinned correctly rounded fp32 value (double evaluation, once per pixel).
#include <cmath>

#include "gi_device.cuh"

namespace vxpt {

struct MaterialDev {
    int grass[10];
    int pom, high_quality_pom, dither_pom;
    float depth_scale;  // 0.115f * u_POMHeight
    float height_exp;   // 1.5f * u_POMExp
    float frame_term;   // fract(mod(float(u_Frame), 384.0f) * (1.0f / PHI)), computed on the host in fp32
    // ShouldUpdate :351 = update_all || the pixel's block is lava; the per-frame constants of BasicTextureDistortion :127-137 (host, fp32,
    // pinned sin / cos / pow)
    int update_all, lava_id;
    float time, lava_r, lava_sx, lava_cy, lava_t13;  // u_Time, fract(u_Time * 0.3f), sin(t * 0.25f), pow(cos(t * 0.15f), 2), t * 1.3f
};
struct MaterialOutDev {
    float* albedo;      // 3 / pixel
    float* normal;      // 3 / pixel
    float4* pbr;
    float* texture_ao;
};

__device__ __forceinline__ float log2_cr(float x) { return (float)log2((double)x); }

// texel (i, j) of level `level` of one layer's mip chain; RGBA8 -> float, c / 255 (OpenGL 4.3 section 2.3.5) read from a 256-entry table
// of exactly those quotients (an IEEE division per channel was half of the kernel's instructions, ncu r01i); the rgb of a GL_SRGB_ALPHA
// array goes through the sRGB decode table instead
__device__ __forceinline__ float4 mip_texel(const SceneDev& S, const uchar4* layer_base, int level, int i, int j, bool srgb) {
    // offset of level k = (4^9 + ... + 4^(10-k)) = (4^10 - 4^(10-k)) / 3
    const int off = ((1 << 20) - (1 << (2 * (10 - level)))) / 3;
    const int n = 512 >> level;
    const uchar4 c = layer_base[off + j * n + i];
    const float* unorm = S.srgb_lut + 256;
    const float* rgb = srgb ? S.srgb_lut : unorm;
    return make_float4(__ldg(rgb + c.x), __ldg(rgb + c.y), __ldg(rgb + c.z), __ldg(unorm + c.w));
}
__device__ __forceinline__ float4 mip_nearest(const SceneDev& S, const uchar4* layer_base, int level, float u, float v, bool srgb) {
    const int n = 512 >> level;
    const int i = ((int)floorf(u * (float)n)) & (n - 1), j = ((int)floorf(v * (float)n)) & (n - 1);
    return mip_texel(S, layer_base, level, i, j, srgb);
}
__device__ __forceinline__ float4 f4_lerp(float4 a, float4 b, float f) {
    const float g = 1.0f - f;
    return make_float4(a.x * g + b.x * f, a.y * g + b.y * f, a.z * g + b.z * f, a.w * g + b.w * f);
}
// lambda of textureGrad (include/vxpt.h: the pinned OpenGL 4.3 section 8.14 isotropic scale factor); the three arrays of a pixel are
// sampled with the same derivatives, so it is computed once
__device__ __forceinline__ float mip_lambda(float4 d) {
    const float dudx = d.x * 512.0f, dvdx = d.y * 512.0f, dudy = d.z * 512.0f, dvdy = d.w * 512.0f;
    const float rho = fmaxf(sqrtf(dudx * dudx + dvdx * dvdx), sqrtf(dudy * dudy + dvdy * dvdy));
    return log2_cr(rho);
}
// textureGrad on a block array, given lambda
__device__ __forceinline__ float4 texture_grad(const SceneDev& S, const uchar4* mips, float layer_f, float u, float v, float lambda, bool srgb,
                                               bool mag_linear) {
    int layer = (int)nearbyintf(layer_f);
    layer = min(max(layer, 0), S.n_mip_layers - 1);
    const uchar4* base = mips + (size_t)layer * VXPT_MIP_CHAIN_TEXELS;
    if (lambda <= (mag_linear ? 0.5f : 0.0f)) {  // magnification, level 0
        if (!mag_linear) return mip_nearest(S, base, 0, u, v, srgb);
        const float x = u * 512.0f - 0.5f, y = v * 512.0f - 0.5f;
        const float fx0 = floorf(x), fy0 = floorf(y);
        const float fx = x - fx0, fy = y - fy0;
        const int i0 = ((int)fx0) & 511, i1 = ((int)fx0 + 1) & 511, j0 = ((int)fy0) & 511, j1 = ((int)fy0 + 1) & 511;
        const float4 a = f4_lerp(mip_texel(S, base, 0, i0, j0, srgb), mip_texel(S, base, 0, i1, j0, srgb), fx);
        const float4 b = f4_lerp(mip_texel(S, base, 0, i0, j1, srgb), mip_texel(S, base, 0, i1, j1, srgb), fx);
        return f4_lerp(a, b, fy);
    }
    if (lambda >= 9.0f) return mip_nearest(S, base, 9, u, v, srgb);
    const float fl = floorf(lambda);
    const int d1 = (int)fl;
    return f4_lerp(mip_nearest(S, base, d1, u, v, srgb), mip_nearest(S, base, d1 + 1, u, v, srgb), lambda - fl);
}

// what a quad member hands to dFdx / dFdy: UV = fract(P) of CalculateVectors :463-530 on its own hit point and face, or nothing when
// the invocation returns before reaching it (sky :360-366, outside the frame)
template <bool LAVA>
__device__ __forceinline__ bool quad_uv(const CameraDev& cam, const MaterialDev& p, const GBufferDev& g, int i, int j, int prow, float& u, float& v,
                                        float& dist, V3& pos, int& nid) {
    if (i >= cam.width || j >= cam.height) return false;
    const size_t px = (size_t)prow * cam.width + i;
    if (LAVA && !p.update_all && min((int)g.block_id[px], 127) != p.lava_id) return false;  // discarded before it takes a derivative (:351-357)
    dist = 1.0f / g.inv_t[px];  // GetPositionAt :107-112
    if (dist < 0.0f) return false;
    nid = g.normal_id[px];
    if (nid > 5) return false;  // no face: the shader's CalculateVectors would leave its outputs unset; shaded like a miss
    const float tu = ((float)i + 0.5f) / (float)cam.width, tv = ((float)j + 0.5f) / (float)cam.height;
    pos = ray_origin(cam) + normalize3(ray_direction_at(cam, tu, tv)) * dist;
    if (nid <= 1) { u = fractf(pos.x); v = fractf(pos.y); }
    else if (nid <= 3) { u = fractf(pos.x); v = fractf(pos.z); }
    else { u = fractf(pos.z); v = fractf(pos.y); }
    return true;
}

// one RGBA8 texel of an animated lava texture -> float4 (c / 255 through the table)
__device__ __forceinline__ float4 lava_texel(const SceneDev& S, const uchar4* tex, int i, int j, int k) {
    const uchar4 c = tex[(k * VXPT_LAVA_SIZE + j) * VXPT_LAVA_SIZE + i];
    const float* unorm = S.srgb_lut + 256;
    return make_float4(__ldg(unorm + c.x), __ldg(unorm + c.y), __ldg(unorm + c.z), __ldg(unorm + c.w));
}
// texture(sampler3D, p): GL_LINEAR, GL_REPEAT on the three axes (Core/AnimatedTexture.cpp:11-15); x, then y, then z
__device__ __forceinline__ float4 lava_sample(const SceneDev& S, const uchar4* tex, float s, float t, float r) {
    const float x = s * (float)VXPT_LAVA_SIZE - 0.5f, y = t * (float)VXPT_LAVA_SIZE - 0.5f, z = r * (float)VXPT_LAVA_FRAMES - 0.5f;
    const float x0 = floorf(x), y0 = floorf(y), z0 = floorf(z);
    const float fx = x - x0, fy = y - y0, fz = z - z0;
    const int i0 = ((int)x0) & (VXPT_LAVA_SIZE - 1), i1 = ((int)x0 + 1) & (VXPT_LAVA_SIZE - 1);
    const int j0 = ((int)y0) & (VXPT_LAVA_SIZE - 1), j1 = ((int)y0 + 1) & (VXPT_LAVA_SIZE - 1);
    const int k0 = ((int)z0) & (VXPT_LAVA_FRAMES - 1), k1 = ((int)z0 + 1) & (VXPT_LAVA_FRAMES - 1);
    const float4 a = f4_lerp(f4_lerp(lava_texel(S, tex, i0, j0, k0), lava_texel(S, tex, i1, j0, k0), fx),
                             f4_lerp(lava_texel(S, tex, i0, j1, k0), lava_texel(S, tex, i1, j1, k0), fx), fy);
    const float4 b = f4_lerp(f4_lerp(lava_texel(S, tex, i0, j0, k1), lava_texel(S, tex, i1, j0, k1), fx),
                             f4_lerp(lava_texel(S, tex, i0, j1, k1), lava_texel(S, tex, i1, j1, k1), fx), fy);
    return f4_lerp(a, b, fz);
}

// POM = u_POM, LAVA = a lava block id is set: separate instantiations, so that the default pass keeps its 40 registers.
// SHFL (VXPT_OPT_MATERIAL_QUAD_SHUFFLE, default pass only): the quad partners' UV come from lanes ^1 and ^8 of the 8x4 warp tile
// instead of two more ray set-ups; the operands are the same values, so the planes are the same bits.
template <bool POM, bool LAVA, bool SHFL = false>
__global__ void __launch_bounds__(256) gbuffer_kernel(const SceneDev S, const __grid_constant__ CameraDev cam, const __grid_constant__ MaterialDev p,
                                                      const GBufferDev g, const MaterialOutDev out) {
    int i, j, prow;
    const bool in_frame = thread_pixel(cam, i, j, prow);
    if (!SHFL && !in_frame) return;
    const size_t px = (size_t)prow * cam.width + i;
    float u = 0.0f, v = 0.0f, dist;
    V3 pos;
    int nid;
#if defined(__CUDA_ARCH__) || defined(VXPT_HOST_SHADOW)
    float sx_u = 0.0f, sx_v = 0.0f, sy_u = 0.0f, sy_v = 0.0f;
    bool sx_ok = false, sy_ok = false, own_ok = false;
    if (SHFL) {  // every lane of the warp takes part; lanes outside the slab hand over "nothing", as quad_uv says of them
        own_ok = in_frame && quad_uv<LAVA>(cam, p, g, i, j, prow, u, v, dist, pos, nid);
        sx_u = __shfl_xor_sync(0xffffffffu, u, 1); sx_v = __shfl_xor_sync(0xffffffffu, v, 1);
        sx_ok = __shfl_xor_sync(0xffffffffu, (int)own_ok, 1) != 0;
        sy_u = __shfl_xor_sync(0xffffffffu, u, 8); sy_v = __shfl_xor_sync(0xffffffffu, v, 8);
        sy_ok = __shfl_xor_sync(0xffffffffu, (int)own_ok, 8) != 0;
        if (!in_frame) return;
    }
#else
    const float sx_u = 0.0f, sx_v = 0.0f, sy_u = 0.0f, sy_v = 0.0f;
    const bool sx_ok = false, sy_ok = false, own_ok = false;
#endif
    const int block_early = LAVA ? min((int)g.block_id[px], 127) : 0;  // (read here only when a lava id is set: keeps the default pass at 40 registers)
    const bool is_lava = LAVA && block_early == p.lava_id;
    if (LAVA && !p.update_all && !is_lava) return;  // :351-357 discard: the attachments keep their texels
    if (SHFL ? !own_ok : !quad_uv<LAVA>(cam, p, g, i, j, prow, u, v, dist, pos, nid)) {  // :360-366
        if (out.albedo) { out.albedo[3 * px] = 0.0f; out.albedo[3 * px + 1] = 0.0f; out.albedo[3 * px + 2] = 0.0f; }
        if (out.normal) { out.normal[3 * px] = 1.0f; out.normal[3 * px + 1] = 1.0f; out.normal[3 * px + 2] = 1.0f; }
        if (out.pbr) out.pbr[px] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (out.texture_ao) out.texture_ao[px] = 0.0f;
        return;
    }
    // GetUVDerivative :443-461 — the quad partners along x and y (the rows of a quad are neighbours in the planes too: bands are even)
    float un, vn, dn;
    V3 pn;
    int nn;
    float ax = u, ay = v, bx = u, by = v;  // x pair: a = even column, b = odd column
    if (SHFL) { un = sx_u; vn = sx_v; }
    if (SHFL ? sx_ok : quad_uv<LAVA>(cam, p, g, i ^ 1, j, prow, un, vn, dn, pn, nn)) {
        if (i & 1) { ax = un; ay = vn; } else { bx = un; by = vn; }
    }
    float cx = u, cy = v, ex = u, ey = v;  // y pair: c = even row, e = odd row
    if (SHFL) { un = sy_u; vn = sy_v; }
    if (SHFL ? sy_ok : quad_uv<LAVA>(cam, p, g, i, j ^ 1, prow ^ 1, un, vn, dn, pn, nn)) {
        if (j & 1) { cx = un; cy = vn; } else { ex = un; ey = vn; }
    }
    float4 d = make_float4(bx - ax, by - ay, ex - cx, ey - cy);
    {
        const float a2x = fractf(ax + 0.25f), a2y = fractf(ay + 0.25f), b2x = fractf(bx + 0.25f), b2y = fractf(by + 0.25f);
        const float c2x = fractf(cx + 0.25f), c2y = fractf(cy + 0.25f), e2x = fractf(ex + 0.25f), e2y = fractf(ey + 0.25f);
        const float4 d2 = make_float4(b2x - a2x, b2y - a2y, e2x - c2x, e2y - c2y);
        if ((d.x * d.x + d.y * d.y) + (d.z * d.z + d.w * d.w) > (d2.x * d2.x + d2.y * d2.y) + (d2.z * d2.z + d2.w * d2.w)) d = d2;
    }
    // GetTextureIDs :532-549
    const int block = LAVA ? block_early : min((int)g.block_id[px], 127);  // GetBlockID :95-99 (the unorm8 round trip is exact)
    float l_albedo = (float)S.materials[block], l_normal = (float)S.materials[128 + block], l_pbr = (float)S.materials[256 + block];
    const float l_emissive = (float)S.materials[384 + block];
    if (block == p.grass[0]) {
        const int o = (nid == 2) ? 1 : ((nid == 3) ? 7 : 4);
        l_albedo = (float)p.grass[o]; l_normal = (float)p.grass[o + 1]; l_pbr = (float)p.grass[o + 2];
    }
    V3 tangent, bitangent;
    if (nid <= 1) { tangent = mk3(1.f, 0.f, 0.f); bitangent = mk3(0.f, 1.f, 0.f); }
    else if (nid <= 3) { tangent = mk3(1.f, 0.f, 0.f); bitangent = mk3(0.f, 0.f, 1.f); }
    else { tangent = mk3(0.f, 0.f, -1.f); bitangent = mk3(0.f, -1.f, 0.f); }
    const V3 face = normal_from_id(nid, 1.0f);
    float lava_u = 0.0f, lava_v = 0.0f;
    if (is_lava) {  // BasicTextureDistortion :127-137 on vec3(UV, fract(u_Time * 0.3f)); liquids skip the parallax march (:394)
        float du = u + p.lava_sx, dv2 = v + p.lava_cy;
        du += cos_cr(du * 10.0f + p.time) * 0.3f;
        dv2 += sin_cr(dv2 * 5.0f + du * 4.0f + p.lava_t13) * 0.4f;
        lava_u = du * (1.0f - 0.91f) + u * 0.91f;
        lava_v = dv2 * (1.0f - 0.91f) + v * 0.91f;
        u = lava_u;
        v = lava_v;
    } else if (POM) {  // Parallax :343-353 -> ReliefParallax :153-199 (everything after its first return is dead code)
        const V3 view = normalize3(pos - ray_origin(cam));
        float bayer_steps = 0.5f;
        if (p.dither_pom) {  // bayer32(gl_FragCoord.xy): bayer2 at five octaves, coarsest first
            float cx[5], cy[5];
            cx[0] = (float)i + 0.5f; cy[0] = (float)j + 0.5f;
#pragma unroll
            for (int k = 1; k < 5; ++k) { cx[k] = 0.5f * cx[k - 1]; cy[k] = 0.5f * cy[k - 1]; }
            float b = 0.0f;
#pragma unroll
            for (int k = 4; k >= 0; --k) {
                const float ax = floorf(cx[k]), ay = floorf(cy[k]);
                const float b2 = fractf(ax * 0.5f + ay * (ay * 0.75f));
                b = k == 4 ? b2 : b * 0.25f + b2;
            }
            bayer_steps = fractf(p.frame_term + b);
        }
        const V3 tv = normalize3(mk3(dot3(view, tangent), dot3(view, bitangent), dot3(view, -face)));
        const float dv = fabsf(face.x) > 0.01f ? tv.z : -(tv.z);
        const float mdx = (tv.x / dv) * p.depth_scale, mdy = (tv.y / dv) * p.depth_scale;
        const int steps = p.high_quality_pom ? (int)mixf(64.0f, 128.0f, clampf(bayer_steps * 0.9f, 0.0f, 1.0f))
                                             : (int)mixf(32.0f, 64.0f, clampf(bayer_steps * 0.85f, 0.0f, 1.0f));
        const float step_size = 1.0f / (float)steps;
        const float su = clampf(u, 0.000001f, 1.0f), sv = clampf(v, 0.000001f, 1.0f);
        const int layer = min(max((int)nearbyintf((float)(int)l_pbr), 0), S.n_mip_layers - 1);
        const uchar4* base = S.pbr_mips + (size_t)layer * VXPT_MIP_CHAIN_TEXELS;
        const float* unorm = S.srgb_lut + 256;
        float cur_depth = 1.0f, best = 1.0f;
#pragma unroll 1
        for (int k = 0; k < steps; ++k) {
            cur_depth -= step_size;
            // texture(u_BlockPBR, ...).z inside the loop: pinned to level 0, GL_LINEAR (include/vxpt.h); only the height channel is fetched
            const float x = (su + mdx * cur_depth) * 512.0f - 0.5f, y = (sv + mdy * cur_depth) * 512.0f - 0.5f;
            const float x0 = floorf(x), y0 = floorf(y), fx = x - x0, fy = y - y0;
            const int i0 = ((int)x0) & 511, i1 = ((int)x0 + 1) & 511, j0 = ((int)y0) & 511, j1 = ((int)y0 + 1) & 511;
            const float t00 = __ldg(unorm + base[j0 * 512 + i0].z), t10 = __ldg(unorm + base[j0 * 512 + i1].z);
            const float t01 = __ldg(unorm + base[j1 * 512 + i0].z), t11 = __ldg(unorm + base[j1 * 512 + i1].z);
            const float h = (t00 * (1.0f - fx) + t10 * fx) * (1.0f - fy) + (t01 * (1.0f - fx) + t11 * fx) * fy;
            if (cur_depth >= pow_cr(h, p.height_exp)) best = cur_depth;  // MapHeight :147-149
        }
        cur_depth = best - step_size * 0.5f;
        u = su + mdx * cur_depth;
        v = sv + mdy * cur_depth;
    }
    u = 1.0f - u;  // :397
    v = 1.0f - v;
    const float lambda = mip_lambda(d);
    const float4 nm = is_lava ? lava_sample(S, S.lava_normal, lava_u, lava_v, p.lava_r) : texture_grad(S, S.normal_mips, l_normal, u, v, lambda, false, true);
    const float nx = nm.x * 2.0f - 1.0f, ny = nm.y * 2.0f - 1.0f, nz = nm.z * 2.0f - 1.0f;
    const V3 mapped = mk3((tangent.x * nx + bitangent.x * ny) + face.x * nz, (tangent.y * nx + bitangent.y * ny) + face.y * nz,
                          (tangent.z * nx + bitangent.z * ny) + face.z * nz);  // tbn * NormalMapped
    const float4 pm = texture_grad(S, S.pbr_mips, l_pbr, u, v, lambda, false, true);
    float emissivity = 0.0f;
    if (l_emissive > -0.5f) emissivity = tex_bilinear1(S.emissive, (int)l_emissive, 512, u, v);
    float4 o_pbr = make_float4(clampf(pm.x, 0.0f, 1.0f), clampf(pm.y, 0.0f, 1.0f), clampf(pm.z, 0.0f, 1.0f), clampf(emissivity, 0.0f, 1.0f));
    const float4 al = is_lava ? lava_sample(S, S.lava_albedo, lava_u, lava_v, p.lava_r) : texture_grad(S, S.albedo_mips, l_albedo, u, v, lambda, true, false);
    const float inside = (u > 0.02f && u < 1.0f - 0.02f && v > 0.02f && v < 1.0f - 0.02f) ? 1.0f : 0.0f;  // BloomLightLeakFix :432-438
    if (!is_lava) o_pbr.w *= inside;
    if (out.albedo) { out.albedo[3 * px] = al.x; out.albedo[3 * px + 1] = al.y; out.albedo[3 * px + 2] = al.z; }
    if (out.normal) { out.normal[3 * px] = mapped.x; out.normal[3 * px + 1] = mapped.y; out.normal[3 * px + 2] = mapped.z; }
    if (out.pbr) out.pbr[px] = o_pbr;
    if (out.texture_ao) out.texture_ao[px] = clampf(pm.w, 0.00000001f, 1.0f);
}

int launch_gbuffer(vxpt_ctx* c, const VxCamera& cam, const VxGBuffer& g, const VxMaterialParams& p, const VxMaterialOut& out) {
    const bool lava = p.lava_block_id >= 0;
    if (!p.update_this_frame && !p.pom && !lava) return VXPT_OK;  // :351-357: every invocation discards (ShouldUpdate = update || lava || u_POM)
    const SceneDev S = make_scene(c);
    MaterialDev d;
    for (int k = 0; k < 10; ++k) d.grass[k] = p.grass_props[k];
    d.pom = p.pom; d.high_quality_pom = p.high_quality_pom; d.dither_pom = p.dither_pom;
    d.depth_scale = 0.115f * p.pom_height;
    d.height_exp = 1.5f * p.pom_exp;
    {
        const float fr = (float)p.frame + 0.0f * 2.0f;
        const float md = fr - 384.0f * std::floor(fr / 384.0f);  // mod(x, y) = x - y * floor(x / y)
        const float t = md * (1.0f / 1.6180339f);
        d.frame_term = t - std::floor(t);
    }
    CameraDev cd;
    for (int k = 0; k < 16; ++k) { cd.inv_view[k] = cam.inv_view[k]; cd.inv_proj[k] = cam.inv_proj[k]; }
    cd.width = cam.width; cd.height = cam.height; cd.row_begin = cam.row_begin; cd.row_end = cam.row_end;
    cd.il_n = cam.interleave_n; cd.il_rank = cam.interleave_rank; cd.il_band = cam.band_rows > 0 ? cam.band_rows : 1;
    const GBufferDev gd{g.t, g.normal_id, g.block_id, g.inv_t, g.hit_voxel, c->opt_texel};
    const MaterialOutDev od{out.albedo, out.normal, reinterpret_cast<float4*>(out.pbr), out.texture_ao};
    d.update_all = (p.update_this_frame || p.pom) ? 1 : 0;
    d.lava_id = lava ? p.lava_block_id : -1;
    d.time = p.time;
    {
        const float r = p.time * 0.3f;
        d.lava_r = r - std::floor(r);
        d.lava_sx = (float)std::sin((double)(p.time * 0.25f));
        d.lava_cy = (float)std::pow((double)(float)std::cos((double)(p.time * 0.15f)), (double)2.0f);
        d.lava_t13 = p.time * 1.3f;
    }
    const dim3 grid((cam.width + 31) / 32, (cam.row_end - cam.row_begin + 7) / 8);
    if (lava) {
        if (p.pom) VX_LAUNCH((gbuffer_kernel<true, true>), grid, 256, c->stream, S, cd, d, gd, od);
        else VX_LAUNCH((gbuffer_kernel<false, true>), grid, 256, c->stream, S, cd, d, gd, od);
    } else if (p.pom) VX_LAUNCH((gbuffer_kernel<true, false>), grid, 256, c->stream, S, cd, d, gd, od);
    else if (c->opt_quad_shuffle) VX_LAUNCH_WARPSYNC((gbuffer_kernel<false, false, true>), grid, 256, c->stream, S, cd, d, gd, od);
    else VX_LAUNCH((gbuffer_kernel<false, false>), grid, 256, c->stream, S, cd, d, gd, od);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

}  // namespace vxpt
