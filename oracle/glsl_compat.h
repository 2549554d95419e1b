// glsl_compat.h — what a GLSL 4.30 shader of the reference needs to compile as C++ (test infrastructure only).
// Vector / matrix types and the built-in functions come from the reference's own vendored glm (Dependencies/glm, 0.9.8); this
// header adds the few things glm lacks: GLSL's implicit int -> float promotion in mixed vector arithmetic, texel fetches on
// GL_RED / r8 unorm8 volumes (OpenGL 4.3 core, section 2.3.5: c / 255 on load, round(clamp(f, 0, 1) * 255) on store), the compute
// built-in gl_GlobalInvocationID and inert stand-ins for the sampler types of code paths the driver never enables.
#pragma once
#define GLM_FORCE_PURE  // scalar code paths only: the arithmetic of every expression is the plain IEEE fp32 operation
#include <cmath>
#include <cstdint>

#include <glm/glm.hpp>

namespace glsl {
using namespace glm;

// ---- implicit int -> float promotions of GLSL ---------------------------------------------------------------------------
inline vec3 operator-(const vec3& a, const ivec3& b) { return a - vec3(b); }
inline vec3 operator-(const ivec3& a, const vec3& b) { return vec3(a) - b; }
inline vec3 operator+(const ivec3& a, const vec3& b) { return vec3(a) + b; }
inline vec3 operator+(const vec3& a, const ivec3& b) { return a + vec3(b); }
inline vec3 operator*(const ivec3& a, const vec3& b) { return vec3(a) * b; }
inline vec3 operator*(int a, const vec3& b) { return float(a) * b; }
inline vec3 operator*(const vec3& a, int b) { return a * float(b); }
inline vec3 operator/(float a, const vec3& b) { return vec3(a) / b; }
inline vec2 operator/(float a, const vec2& b) { return vec2(a) / b; }
inline vec2 operator*(const vec2& a, double b) { return a * float(b); }
using glm::clamp;
using glm::max;
using glm::min;
inline float min(int a, float b) { return glm::min(float(a), b); }
inline float min(float a, int b) { return glm::min(a, float(b)); }
inline float max(int a, float b) { return glm::max(float(a), b); }
inline float max(float a, int b) { return glm::max(a, float(b)); }
inline float clamp(int x, float lo, float hi) { return glm::clamp(float(x), lo, hi); }
inline float clamp(float x, int lo, int hi) { return glm::clamp(x, float(lo), float(hi)); }
inline vec2 operator+(const vec2& a, const ivec2& b) { return a + vec2(b); }
inline vec2 operator/(const ivec2& a, float b) { return vec2(a) / b; }

// ---- swizzles: oracle/glsl2cpp.py rewrites  e.xyz  as  swz_xyz(e)  (and emits the helper) and  v.xy = e;  as  swz_assign_xy(v, '=', e);
template <class V, class E> inline void swz_assign_xy(V& v, char op, const E& e) {
    vec2 cur(v.x, v.y);
    cur = op == '=' ? vec2(e) : op == '+' ? cur + e : op == '-' ? cur - e : op == '*' ? cur * e : cur / e;
    v.x = cur.x; v.y = cur.y;
}
template <class V, class E> inline void swz_assign_xyz(V& v, char op, const E& e) {
    vec3 cur(v.x, v.y, v.z);
    cur = op == '=' ? vec3(e) : op == '+' ? cur + e : op == '-' ? cur - e : op == '*' ? cur * e : cur / e;
    v.x = cur.x; v.y = cur.y; v.z = cur.z;
}
template <class V, class E> inline void swz_assign_rgb(V& v, char op, const E& e) { swz_assign_xyz(v, op, e); }

// GLSL arrays used as values (function results)
template <int N> struct farr {
    float v[N];
    float& operator[](int i) { return v[i]; }
    const float& operator[](int i) const { return v[i]; }
};
inline float rcp(float x) { return 1.0f / x; }  // accepted by the reference's (NVIDIA) GLSL compiler; SURVEY.md A.4

// ---- transcendental functions: GLSL leaves their precision open; pinned to the correctly rounded fp32 value (double evaluation,
//      one rounding) — the definition the oracle and the CUDA kernels use (oracle/vxo_oracle.cpp header) -------------------------------
//      (oracle/glsl2cpp.py renames the calls; vector arguments go to glm)
inline float pinned_sin(float x) { return (float)::sin((double)x); }
inline float pinned_cos(float x) { return (float)::cos((double)x); }
inline float pinned_tan(float x) { return (float)::tan((double)x); }
inline float pinned_pow(float x, float y) { return (float)::pow((double)x, (double)y); }
inline float pinned_log2(float x) { return (float)::log2((double)x); }
inline float pinned_acos(float x) { return (float)::acos((double)x); }
inline float pinned_exp(float x) { return (float)::exp((double)x); }  // renamed only for the shaders that ask for it (glsl2cpp.py's 4th argument)
template <class V> inline V pinned_sin(const V& v) { return glm::sin(v); }
template <class V> inline V pinned_cos(const V& v) { return glm::cos(v); }
template <class V> inline V pinned_tan(const V& v) { return glm::tan(v); }
template <class V> inline V pinned_pow(const V& a, const V& b) { return glm::pow(a, b); }
// mix(x, y, a) = x * (1 - a) + y * a, the GLSL specification's formula (glm evaluates x + a * (y - x), which rounds differently)
inline float pinned_mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline float pinned_mix(int x, int y, float a) { return (float)x * (1.0f - a) + (float)y * a; }
template <class V> inline V pinned_mix(const V& x, const V& y, float a) { return x * (1.0f - a) + y * a; }
template <class V> inline V pinned_mix(const V& x, const V& y, const V& a) { return x * (V(1.0f) - a) + y * a; }

// ---- textures and images ----------------------------------------------------------------------------------------------
struct sampler3D {  // GL_RED, GL_UNSIGNED_BYTE, NEAREST (Core/Texture3D.cpp:8-28)
    const uint8_t* data = nullptr;
    int sx = 0, sy = 0, sz = 0;
    // the animated lava textures (Core/AnimatedTexture.cpp:8-26): GL_RGBA8 [frames][n][n][4], GL_LINEAR, GL_REPEAT on the three axes
    const uint8_t* rgba = nullptr;
    int rn = 0, rframes = 0;
};
inline vec4 texelFetch(const sampler3D& s, const ivec3& p, int /*lod*/) {
    const uint8_t c = s.data[(size_t)p.x + (size_t)s.sx * ((size_t)p.y + (size_t)s.sy * (size_t)p.z)];
    return vec4(float(c) / 255.0f, 0.0f, 0.0f, 1.0f);
}
struct image3D {  // layout(r8)
    uint8_t* data = nullptr;
    int sx = 0, sy = 0, sz = 0;
};
inline vec4 imageLoad(const image3D& s, const ivec3& p) {
    const uint8_t c = s.data[(size_t)p.x + (size_t)s.sx * ((size_t)p.y + (size_t)s.sy * (size_t)p.z)];
    return vec4(float(c) / 255.0f, 0.0f, 0.0f, 1.0f);
}
inline void imageStore(image3D& s, const ivec3& p, const vec4& v) {
    float f = v.x < 0.0f ? 0.0f : (v.x > 1.0f ? 1.0f : v.x);
    s.data[(size_t)p.x + (size_t)s.sx * ((size_t)p.y + (size_t)s.sy * (size_t)p.z)] = (uint8_t)std::nearbyintf(f * 255.0f);
}
struct sampler2D {  // fp32 texels, 1..4 components.  Default: NEAREST (the G-buffer attachments point-sampled at the pixel's own centre)
    const float* data = nullptr;
    int w = 0, h = 0, comps = 1;
    // the denoising passes sample FBO attachments away from texel centres: GL_LINEAR / GL_NEAREST as Core/Pipeline.cpp:1094-1152 declares
    // each attachment, GL_REPEAT (Core/GLClasses/Framebuffer.cpp:66-67).  Pinned filter arithmetic (GL leaves it open): OpenGL 4.3
    // section 8.14.2, x = u * w - 0.5, i0 = floor(x) mod w, f = x - floor(x), weights applied as a * (1 - f) + b * f, x first.
    bool linear = false, repeat = false;
};
inline vec4 texelFetch(const sampler2D& s, const ivec2& p, int /*lod*/) {
    const float* t = s.data + ((size_t)p.y * s.w + p.x) * s.comps;
    return vec4(t[0], s.comps > 1 ? t[1] : 0.0f, s.comps > 2 ? t[2] : 0.0f, s.comps > 3 ? t[3] : 1.0f);
}
inline ivec2 textureSize(const sampler2D& s, int /*lod*/) { return ivec2(s.w, s.h); }
inline int wrap_repeat(int i, int n) { const int m = i % n; return m < 0 ? m + n : m; }
inline vec4 texture(const sampler2D& s, const vec2& uv) {
    if (s.repeat) {
        if (!s.linear) return texelFetch(s, ivec2(wrap_repeat((int)std::floor(uv.x * (float)s.w), s.w), wrap_repeat((int)std::floor(uv.y * (float)s.h), s.h)), 0);
        const float x = uv.x * (float)s.w - 0.5f, y = uv.y * (float)s.h - 0.5f;
        const float fx0 = std::floor(x), fy0 = std::floor(y);
        const float fx = x - fx0, fy = y - fy0;
        const int i0 = wrap_repeat((int)fx0, s.w), i1 = wrap_repeat((int)fx0 + 1, s.w), j0 = wrap_repeat((int)fy0, s.h), j1 = wrap_repeat((int)fy0 + 1, s.h);
        const vec4 a = texelFetch(s, ivec2(i0, j0), 0) * (1.0f - fx) + texelFetch(s, ivec2(i1, j0), 0) * fx;
        const vec4 b = texelFetch(s, ivec2(i0, j1), 0) * (1.0f - fx) + texelFetch(s, ivec2(i1, j1), 0) * fx;
        return a * (1.0f - fy) + b * fy;
    }
    int i = (int)std::floor(uv.x * (float)s.w), j = (int)std::floor(uv.y * (float)s.h);
    i = i < 0 ? 0 : (i >= s.w ? s.w - 1 : i);
    j = j < 0 ? 0 : (j >= s.h ? s.h - 1 : j);
    return texelFetch(s, ivec2(i, j), 0);
}
// Array textures and the sky cube map.  OpenGL leaves filter arithmetic (and, inside divergent control flow, even the mip level of an
// implicit-LOD texture() call) to the driver; the parity contract pins them (oracle/vxo_oracle.cpp header): the driver
// binds the texel array of the level the call addresses; textureLod() fetches the nearest texel, texture() filters bilinearly in fp32
// with the weights applied as  a * (1 - f) + b * f , x first; both wrap with GL_REPEAT.  The cube map filters inside the major-axis
// face and clamps at its edge.
struct sampler2DArray {
    const float* data = nullptr;  // [layers][n][n][comps]
    int n = 0, comps = 4;
    bool nearest_only = false;    // reflection pass: its implicit-LOD texture() calls are pinned to the nearest texel of the bound level
    // alpha-tested traversal (StopRay): the alpha channel of the whole mip chain, levels 0..8 of a 512^2 array back to back per layer
    // (unorm8, as the GL_SRGB_ALPHA storage holds it); textureLod with the shader's integer LOD reads the nearest texel of that level
    // (GL_NEAREST_MIPMAP_LINEAR with a zero fraction, Core/GLClasses/TextureArray.cpp:34)
    const uint8_t* alpha_mips = nullptr;
    int alpha_layers = 0;
    // G-buffer material pass: the whole RGBA8 mip chain of 512^2 layers, levels 0..9 back to back, [mip_layers][349525][4]; textureGrad()
    // is the only call that reads it.  srgb: GL_SRGB_ALPHA storage (albedo); mag_linear: GL_TEXTURE_MAG_FILTER (TextureArray.cpp:42)
    const uint8_t* mips = nullptr;
    int mip_layers = 0;
    bool srgb = false, mag_linear = false;
};
inline vec4 fetch_layer_texel(const sampler2DArray& s, int layer, int i, int j) {
    const float* t = s.data + (((size_t)layer * s.n + j) * s.n + i) * s.comps;
    return vec4(t[0], s.comps > 1 ? t[1] : 0.0f, s.comps > 2 ? t[2] : 0.0f, s.comps > 3 ? t[3] : 1.0f);
}
inline vec4 textureLod(const sampler2DArray& s, const vec3& p, float lod /*ignored for `data`: the bound level*/) {
    if (s.alpha_mips) {
        const int level = (int)lod, n = 512 >> level;
        size_t off = 0;
        for (int k = 0; k < level; ++k) off += (size_t)(512 >> k) * (512 >> k);
        const int i = ((int)std::floor(p.x * (float)n)) & (n - 1), j = ((int)std::floor(p.y * (float)n)) & (n - 1);
        int layer = (int)std::nearbyintf(p.z);
        layer = layer < 0 ? 0 : (layer > s.alpha_layers - 1 ? s.alpha_layers - 1 : layer);
        return vec4(1.0f, 1.0f, 1.0f, (float)s.alpha_mips[(size_t)layer * 349524 + off + (size_t)j * n + i] / 255.0f);
    }
    if (!s.data) return vec4(1.0f);
    const int i = ((int)std::floor(p.x * (float)s.n)) & (s.n - 1), j = ((int)std::floor(p.y * (float)s.n)) & (s.n - 1);
    return fetch_layer_texel(s, (int)p.z, i, j);
}
inline vec4 fetch_mip_texel(const sampler2DArray& s, int layer, int level, int i, int j);
inline vec4 texture(const sampler2DArray& s, const vec3& p) {
    if (s.mips) {  // implicit-LOD fetch on a mip-mapped block array (the parallax march of GenerateGBuffer.glsl): pinned to level 0, GL_LINEAR
        const float x = p.x * 512.0f - 0.5f, y = p.y * 512.0f - 0.5f;
        const float fx0 = std::floor(x), fy0 = std::floor(y);
        const float fx = x - fx0, fy = y - fy0;
        const int i0 = ((int)fx0) & 511, i1 = ((int)fx0 + 1) & 511, j0 = ((int)fy0) & 511, j1 = ((int)fy0 + 1) & 511;
        int l = (int)std::nearbyintf(p.z);
        l = l < 0 ? 0 : (l > s.mip_layers - 1 ? s.mip_layers - 1 : l);
        const vec4 a = fetch_mip_texel(s, l, 0, i0, j0) * (1.0f - fx) + fetch_mip_texel(s, l, 0, i1, j0) * fx;
        const vec4 b = fetch_mip_texel(s, l, 0, i0, j1) * (1.0f - fx) + fetch_mip_texel(s, l, 0, i1, j1) * fx;
        return a * (1.0f - fy) + b * fy;
    }
    if (!s.data) return vec4(1.0f);
    if (s.nearest_only) return textureLod(s, p, 0.0f);
    const float x = p.x * (float)s.n - 0.5f, y = p.y * (float)s.n - 0.5f;
    const float fx0 = std::floor(x), fy0 = std::floor(y);
    const float fx = x - fx0, fy = y - fy0;
    const int i0 = ((int)fx0) & (s.n - 1), i1 = ((int)fx0 + 1) & (s.n - 1), j0 = ((int)fy0) & (s.n - 1), j1 = ((int)fy0 + 1) & (s.n - 1);
    const int l = (int)p.z;
    const vec4 a = fetch_layer_texel(s, l, i0, j0) * (1.0f - fx) + fetch_layer_texel(s, l, i1, j0) * fx;
    const vec4 b = fetch_layer_texel(s, l, i0, j1) * (1.0f - fx) + fetch_layer_texel(s, l, i1, j1) * fx;
    return a * (1.0f - fy) + b * fy;
}
struct samplerCube {
    const float* data = nullptr;  // [6][n][n][3], faces +X -X +Y -Y +Z -Z
    int n = 0;
};
inline vec4 texture(const samplerCube& s, const vec3& d) {
    const int N = s.n;
    const float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
    int face;
    float sc, tc, ma;
    if (ax >= ay && ax >= az) { face = d.x > 0 ? 0 : 1; sc = d.x > 0 ? -d.z : d.z; tc = -d.y; ma = ax; }   // OpenGL 4.3 table 8.19
    else if (ay >= az)        { face = d.y > 0 ? 2 : 3; sc = d.x; tc = d.y > 0 ? d.z : -d.z; ma = ay; }
    else                      { face = d.z > 0 ? 4 : 5; sc = d.z > 0 ? d.x : -d.x; tc = -d.y; ma = az; }
    const float u = 0.5f * (sc / ma + 1.0f) * (float)N - 0.5f, v = 0.5f * (tc / ma + 1.0f) * (float)N - 0.5f;
    const float fu0 = std::floor(u), fv0 = std::floor(v);
    const float fu = u - fu0, fv = v - fv0;
    auto cl = [N](int i) { return i < 0 ? 0 : (i > N - 1 ? N - 1 : i); };
    const int i0 = cl((int)fu0), i1 = cl((int)fu0 + 1), j0 = cl((int)fv0), j1 = cl((int)fv0 + 1);
    const float* F = s.data + (size_t)face * N * N * 3;
    auto tx = [&](int i, int j) { const float* p = F + ((size_t)j * N + i) * 3; return vec3(p[0], p[1], p[2]); };
    const vec3 a = tx(i0, j0) * (1.0f - fu) + tx(i1, j0) * fu;
    const vec3 b = tx(i0, j1) * (1.0f - fu) + tx(i1, j1) * fu;
    return vec4(a * (1.0f - fv) + b * fv, 1.0f);
}

// ---- textureGrad on a mip-mapped block array, and the screen-space derivatives that feed it (GenerateGBuffer.glsl:404-461) ----------
// OpenGL 4.3 section 8.14 with the isotropic scale factor (the anisotropy extension leaves rho to the driver and is not modelled):
// rho = max(|dP/dx| , |dP/dy|) in texels of level 0, lambda = log2(rho) (pinned: correctly rounded); lambda <= c magnifies on level 0
// (c = 0.5 when the mag filter is LINEAR and the min filter NEAREST_MIPMAP_*, else 0), otherwise GL_NEAREST_MIPMAP_LINEAR blends the
// nearest texels of levels floor(lambda) and floor(lambda) + 1 (clamped to the last level) by fract(lambda).  Texel -> float is c / 255,
// sRGB-decoded for the rgb of a GL_SRGB_ALPHA array (section 8.23, evaluated in double).
inline float srgb_decode(uint8_t c) {
    static float lut[256];
    static bool ready = false;
    if (!ready) {
#pragma omp critical(glsl_srgb_lut)
        if (!ready) {
            for (int k = 0; k < 256; ++k) {
                const double cs = (double)k / 255.0;
                lut[k] = (float)(cs <= 0.04045 ? cs / 12.92 : ::pow((cs + 0.055) / 1.055, 2.4));
            }
            ready = true;
        }
    }
    return lut[c];
}
inline vec4 fetch_mip_texel(const sampler2DArray& s, int layer, int level, int i, int j) {
    size_t off = 0;
    for (int k = 0; k < level; ++k) off += (size_t)(512 >> k) * (512 >> k);
    const int n = 512 >> level;
    const uint8_t* c = s.mips + ((size_t)layer * 349525 + off + (size_t)j * n + i) * 4;
    if (s.srgb) return vec4(srgb_decode(c[0]), srgb_decode(c[1]), srgb_decode(c[2]), (float)c[3] / 255.0f);
    return vec4((float)c[0] / 255.0f, (float)c[1] / 255.0f, (float)c[2] / 255.0f, (float)c[3] / 255.0f);
}
inline vec4 fetch_mip_nearest(const sampler2DArray& s, int layer, int level, const vec3& p) {
    const int n = 512 >> level;
    return fetch_mip_texel(s, layer, level, ((int)std::floor(p.x * (float)n)) & (n - 1), ((int)std::floor(p.y * (float)n)) & (n - 1));
}
inline vec4 textureGrad(const sampler2DArray& s, const vec3& p, const vec2& dPdx, const vec2& dPdy) {
    if (!s.mips) return vec4(1.0f);
    int layer = (int)std::nearbyintf(p.z);
    layer = layer < 0 ? 0 : (layer > s.mip_layers - 1 ? s.mip_layers - 1 : layer);
    const float ux = dPdx.x * 512.0f, vx = dPdx.y * 512.0f, uy = dPdy.x * 512.0f, vy = dPdy.y * 512.0f;
    const float rx = std::sqrt(ux * ux + vx * vx), ry = std::sqrt(uy * uy + vy * vy);
    const float rho = rx > ry ? rx : ry;
    const float lambda = (float)::log2((double)rho);
    if (lambda <= (s.mag_linear ? 0.5f : 0.0f)) {
        if (!s.mag_linear) return fetch_mip_nearest(s, layer, 0, p);
        const float x = p.x * 512.0f - 0.5f, y = p.y * 512.0f - 0.5f;
        const float fx0 = std::floor(x), fy0 = std::floor(y);
        const float fx = x - fx0, fy = y - fy0;
        const int i0 = ((int)fx0) & 511, i1 = ((int)fx0 + 1) & 511, j0 = ((int)fy0) & 511, j1 = ((int)fy0 + 1) & 511;
        const vec4 a = fetch_mip_texel(s, layer, 0, i0, j0) * (1.0f - fx) + fetch_mip_texel(s, layer, 0, i1, j0) * fx;
        const vec4 b = fetch_mip_texel(s, layer, 0, i0, j1) * (1.0f - fx) + fetch_mip_texel(s, layer, 0, i1, j1) * fx;
        return a * (1.0f - fy) + b * fy;
    }
    if (lambda >= 9.0f) return fetch_mip_nearest(s, layer, 9, p);
    const float fl = std::floor(lambda);
    const float f = lambda - fl;
    return fetch_mip_nearest(s, layer, (int)fl, p) * (1.0f - f) + fetch_mip_nearest(s, layer, (int)fl + 1, p) * f;
}

// dFdx / dFdy: a fragment shader runs in 2x2 quads and a derivative is the difference between the two members of the quad's row /
// column (fine derivatives).  The driver runs the four invocations of a quad twice: a RECORD run, in which every dFdx / dFdy call logs
// its operand (and returns 0), then a REPLAY run in which call number k returns (odd member's k-th operand) - (even member's).  A member
// that made fewer than k+1 calls (it returned early: sky; or lies outside the frame) contributes the calling fragment's own operand.
// Valid because the number and order of derivative calls of an invocation never depend on a derivative's value.
struct QuadDerivatives {
    enum { MAX_CALLS = 32 };
    int mode = 0;  // 0 = record, 1 = replay
    int lane = 0;  // (x & 1) | ((y & 1) << 1)
    int call = 0;
    int ncalls[4] = {0, 0, 0, 0};
    vec2 operand[4][MAX_CALLS];
    vec2 diff(const vec2& v, int even_lane, int odd_lane) {
        const int k = call++;
        if (mode == 0) {
            if (k < MAX_CALLS) operand[lane][k] = v;
            ncalls[lane] = call;
            return vec2(0.0f);
        }
        const vec2 e = k < ncalls[even_lane] ? operand[even_lane][k] : v;
        const vec2 o = k < ncalls[odd_lane] ? operand[odd_lane][k] : v;
        return o - e;
    }
};
static thread_local QuadDerivatives gl_Quad;
inline vec2 dFdx(const vec2& v) { return gl_Quad.diff(v, gl_Quad.lane & 2, (gl_Quad.lane & 2) | 1); }
inline vec2 dFdy(const vec2& v) { return gl_Quad.diff(v, gl_Quad.lane & 1, (gl_Quad.lane & 1) | 2); }

// code paths the drivers never enable (light-propagation volume, clouds, screen-space reprojection): inert stand-ins
struct usampler3D {};
inline uvec4 texture(const usampler3D&, const vec3&) { return uvec4(0u); }
inline uvec4 texelFetch(const usampler3D&, const ivec3&, int) { return uvec4(0u); }
inline vec4 texture(const sampler3D& s, const vec3& p) {
    if (!s.rgba) return vec4(0.0f);
    // trilinear, OpenGL 4.3 section 8.14.2 in fp32: weights a * (1 - f) + b * f, x then y then z
    const float x = p.x * (float)s.rn - 0.5f, y = p.y * (float)s.rn - 0.5f, z = p.z * (float)s.rframes - 0.5f;
    const float x0 = std::floor(x), y0 = std::floor(y), z0 = std::floor(z);
    const float fx = x - x0, fy = y - y0, fz = z - z0;
    auto wr = [](int i, int n) { const int m = i % n; return m < 0 ? m + n : m; };
    const int i0 = wr((int)x0, s.rn), i1 = wr((int)x0 + 1, s.rn), j0 = wr((int)y0, s.rn), j1 = wr((int)y0 + 1, s.rn);
    const int k0 = wr((int)z0, s.rframes), k1 = wr((int)z0 + 1, s.rframes);
    auto tx = [&](int i, int j, int k) {
        const uint8_t* c = s.rgba + (((size_t)k * s.rn + j) * s.rn + i) * 4;
        return vec4((float)c[0] / 255.0f, (float)c[1] / 255.0f, (float)c[2] / 255.0f, (float)c[3] / 255.0f);
    };
    auto lerp = [](const vec4& a, const vec4& b, float f) { return a * (1.0f - f) + b * f; };
    const vec4 a = lerp(lerp(tx(i0, j0, k0), tx(i1, j0, k0), fx), lerp(tx(i0, j1, k0), tx(i1, j1, k0), fx), fy);
    const vec4 b = lerp(lerp(tx(i0, j0, k1), tx(i1, j0, k1), fx), lerp(tx(i0, j1, k1), tx(i1, j1, k1), fx), fy);
    return lerp(a, b, fz);
}
inline uint clamp(uint x, int lo, int hi) { return x < (uint)lo ? (uint)lo : (x > (uint)hi ? (uint)hi : x); }

// SSBO atomics: the drivers run the invocations of such shaders one after another
inline uint atomicAdd(uint& mem, uint v) { const uint old = mem; mem += v; return old; }

static thread_local uvec3 gl_GlobalInvocationID;
static thread_local vec4 gl_FragCoord;  // per invocation: the driver's OpenMP threads each run whole invocations
}  // namespace glsl
