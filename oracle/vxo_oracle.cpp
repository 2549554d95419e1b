// vxo_oracle.cpp — CPU restatement of the reference's distance-field build and DF-accelerated voxel traversal.
//
// THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load it.  The product library (voxelpathtracer_b200/csrc) never
// links, includes or calls anything in this directory and has no CPU fallback.
//
// PARITY PINNED BY THE REFERENCE ITSELF: swr06/VoxelPathTracer ships no tests, golden vectors or fixtures for this
// path and its GL application cannot run in the build container (no GL/EGL/OSMesa), but its shader source
// can be compiled: oracle/Makefile translates ManhattanDistance{X,Y,Z}.comp, InitialRayTraceFrag.glsl,
// ShadowRayTraceFrag.glsl, DiffuseRayTraceFrag.glsl and ReflectionTraceFrag.glsl where they lie (declarations only,
// oracle/glsl2cpp.py) and builds them as C++ against the reference's vendored glm (oracle/_ref/libref_shaders.so).  Every
// output of this file equals that library's bit for bit on the BASELINE worlds and frames and on live edge cases
// (tests/test_oracle_vs_reference_shaders.py; committed digests: tests/golden/ref_shader_digests.json).  Older pins kept:
// analytic known-answer tests, a brute-force definition of the distance field (vxo_df_bruteforce) and an independent
// plain-DDA cross-check (vxo_plain_dda, after the author's unused Core/Shaders/Implementations/DDA/DDA.glsl).
//
// Every function cites the reference file:line it restates (paths relative to the reference tree).
// Floating point: build with -O2 -ffp-contract=off (no FMA contraction), IEEE division and sqrt; operation
// order follows the GLSL source as written.  Where GLSL leaves a function's precision open (sin, cos, pow) the
// oracle pins the correctly-rounded fp32 value, computed in double and rounded once.
//
// Pinned definitions GL leaves to the driver:
//   dot(a,b)      = (a.x*b.x + a.y*b.y) + a.z*b.z
//   normalize(v)  = v * (1.0f / sqrt(dot(v,v)))                 (glm func_geometric.inl:94)
//   cross(a,b)    = (a.y*b.z - b.y*a.z, a.z*b.x - b.z*a.x, a.x*b.y - b.x*a.y)
//   mat4 * vec4   = (m0*x + m1*y) + (m2*z + m3*w)               (glm type_mat4x4.inl:526-537)
//   mat3 * vec3   = (m0*x + m1*y) + m2*z
//   mix(a,b,t)    = a*(1-t) + b*t ; fract(x) = x - floor(x) ; clamp(x,lo,hi) = min(max(x,lo),hi)
//   v_TexCoords at pixel (i,j) = ((i+0.5)/W, (j+0.5)/H), gl_FragCoord.xy = (i+0.5, j+0.5)
//   G-buffer reads by the secondary passes: same resolution, point sampled, fp32 t
//   textureLod(array, uvw, k) with integer k>0 = nearest texel of the pre-baked level k, REPEAT
//   texture(emissive array) = bilinear on level 0, REPEAT ; texture(cubemap) = bilinear within the major-axis
//   face, clamped to the face edge
//   float->int conversion of NaN / out-of-range positions = "outside the volume" (A.3 note 5)
//   blue-noise rankingTile index is clamped to the array (A.5)

#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/vxpt.h"

extern "C" {

typedef struct VxoScene {
    int32_t wx, wy, wz;
    const uint8_t* grid;
    const uint8_t* df;
    const int32_t* materials;  // 768, Core/BlockDataSSBO.cpp:28-35 order
    const int32_t* sobol;      // 65536
    const int32_t* scramble;   // 131072
    const int32_t* rank;       // 131072
    const float* albedo_lod3;  // [n_layers][64][64][4]
    const float* pbr_lod2;     // [n_layers][128][128][4]
    int32_t n_layers;
    const float* emissive_lod0;  // [n_emissive_layers][512][512]
    int32_t n_emissive_layers;
    const float* sky;  // [6][sky_n][sky_n][3]
    int32_t sky_n;
    const uint8_t* shadow_noise;  // [256][256][4]
    const float* normal_lod3;     // [n_normal_layers][64][64][4]
    int32_t n_normal_layers;
    const float* emissive_lod2;   // [n_emissive_layers][128][128]
    const uint8_t* alpha_mips;    // [n_alpha_layers][VXPT_ALPHA_MIP_TEXELS]: albedo alpha, mip levels 0..8 (alpha-tested traversal)
    int32_t n_alpha_layers;
    // G-buffer material pass: RGBA8 mip chains, levels 0..9 of 512^2 layers, [n_mip_layers][VXPT_MIP_CHAIN_TEXELS][4]
    const uint8_t* albedo_mips;   // GL_SRGB_ALPHA
    const uint8_t* normal_mips;   // GL_RGBA
    const uint8_t* pbr_mips;      // GL_RGBA
    int32_t n_mip_layers;
    const uint8_t* lava_albedo;   // animated lava textures, RGBA8 [VXPT_LAVA_FRAMES][VXPT_LAVA_SIZE][VXPT_LAVA_SIZE][4]
    const uint8_t* lava_normal;
} VxoScene;

typedef struct VxoStats {
    uint64_t rays, df_fetches, vox_fetches;
} VxoStats;

}  // extern "C"

namespace {

// ------------------------------------------------------------------------------------------------ vectors
struct v3 {
    float x, y, z;
    float& operator[](int i) { return i == 0 ? x : (i == 1 ? y : z); }
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline v3 V(float x, float y, float z) { return v3{x, y, z}; }
inline v3 V(float s) { return v3{s, s, s}; }
inline v3 operator+(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
inline v3 operator-(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
inline v3 operator*(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
inline v3 operator/(v3 a, v3 b) { return V(a.x / b.x, a.y / b.y, a.z / b.z); }
inline v3 operator*(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
inline v3 operator*(float s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }
inline v3 operator/(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
inline v3 operator-(v3 a) { return V(-a.x, -a.y, -a.z); }
inline float dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float length(v3 a) { return std::sqrt(dot(a, a)); }
inline v3 normalize(v3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }
inline v3 cross(v3 a, v3 b) { return V(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline float clampf(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
inline v3 clamp3(v3 a, float lo, float hi) { return V(clampf(a.x, lo, hi), clampf(a.y, lo, hi), clampf(a.z, lo, hi)); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline v3 mix3(v3 a, v3 b, float t) { return a * (1.0f - t) + b * t; }
inline float fractf(float x) { return x - std::floor(x); }
// correctly rounded fp32 transcendental pins
inline float sin_cr(float x) { return (float)std::sin((double)x); }
inline float cos_cr(float x) { return (float)std::cos((double)x); }
inline float pow_cr(float x, float y) { return (float)std::pow((double)x, (double)y); }

// column-major mat4 * vec4, glm order
inline void mat4_mul(const float* m, const float v[4], float out[4]) {
    for (int r = 0; r < 4; ++r) out[r] = (m[0 + r] * v[0] + m[4 + r] * v[1]) + (m[8 + r] * v[2] + m[12 + r] * v[3]);
}

const float PI_F = 3.14159265359f;

struct Scene {
    VxoScene s;
    inline bool in_volume_f(float fx, float fy, float fz) const {
        // IsInVolume (InitialRayTraceFrag.glsl:68-78) on floor()ed coordinates; written so NaN -> false
        return fx >= 0.0f && fy >= 0.0f && fz >= 0.0f && fx <= (float)(s.wx - 1) && fy <= (float)(s.wy - 1) &&
               fz <= (float)(s.wz - 1);
    }
    inline size_t idx(int x, int y, int z) const { return (size_t)x + (size_t)s.wx * ((size_t)y + (size_t)s.wy * (size_t)z); }
};

struct Stats {
    uint64_t rays = 0, df = 0, vox = 0;
};

struct Hit {
    float t;
    int min_idx;
    int sgn[3];
    int block;   // raw byte
    int vox[3];  // voxel of the final position (valid when t > 0)
    v3 normal;
};

inline int isign(float d) { return (d > 0.0f) - (d < 0.0f); }
inline int wrap_repeat(int i, int n) { const int m = i % n; return m < 0 ? m + n : m; }  // GL_REPEAT

// GetVoxel(ivec3(floor(p))) — InitialRayTraceFrag.glsl:80-88
inline int get_voxel_at(const Scene& S, v3 p, Stats& st, int* vox = nullptr) {
    float fx = std::floor(p.x), fy = std::floor(p.y), fz = std::floor(p.z);
    if (!S.in_volume_f(fx, fy, fz)) return 0;
    int x = (int)fx, y = (int)fy, z = (int)fz;
    if (vox) { vox[0] = x; vox[1] = y; vox[2] = z; }
    st.vox++;
    return S.s.grid[S.idx(x, y, z)];
}

// VoxelTraversalDF — InitialRayTraceFrag.glsl:307-374 (identical copies: ShadowRayTraceFrag.glsl:222-289,
// DiffuseRayTraceFrag.glsl:1043-1110, ReflectionTraceFrag.glsl:1088-1155).  SURVEY.md A.3.
float traverse_df(const Scene& S, v3 origin, v3 direction, int max_it, Hit& h, Stats& st) {
    const v3 initial_origin = origin;
    bool intersection = false;
    int min_idx = 0;
    const int sg[3] = {isign(direction.x), isign(direction.y), isign(direction.z)};
    st.rays++;
    for (int itr = 0; itr < max_it; ++itr) {
        float fx = std::floor(origin.x), fy = std::floor(origin.y), fz = std::floor(origin.z);
        if (!S.in_volume_f(fx, fy, fz)) { intersection = false; break; }
        st.df++;
        // GetDistance(Loc) * 255 : the R8 round trip is exact for 0..255 (A.1)
        float dist = (float)S.s.df[S.idx((int)fx, (int)fy, (int)fz)];
        // ToConservativeEuclidean — InitialRayTraceFrag.glsl:90-93
        int euclid = (int)std::floor(dist == 1.0f ? 1.0f : dist * 0.57735026918f);
        if (euclid == 0) break;
        if (euclid == 1) {
            int g[3] = {(int)origin.x, (int)origin.y, (int)origin.z};
            v3 w = V(origin.x - (float)g[0], origin.y - (float)g[1], origin.z - (float)g[2]);
            const int half[3] = {(1 + sg[0]) >> 1, (1 + sg[1]) >> 1, (1 + sg[2]) >> 1};
            v3 inv = V(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
            v3 f = V(((float)half[0] - w.x) * inv.x, ((float)half[1] - w.y) * inv.y, ((float)half[2] - w.z) * inv.z);
            min_idx = (f.x < f.y && sg[0] != 0) ? ((f.x < f.z || sg[2] == 0) ? 0 : 2)
                                                : ((f.y < f.z || sg[2] == 0) ? 1 : 2);
            g[min_idx] += sg[min_idx];
            float fm = f[min_idx];
            w = w + direction * fm;
            w[min_idx] = (float)(1 - half[min_idx]);
            origin = V((float)g[0] + w.x, (float)g[1] + w.y, (float)g[2] + w.z);
            origin[min_idx] += (float)sg[min_idx] * 0.0001f;
            intersection = true;
        } else {
            origin = origin + (float)(euclid - 1) * direction;
        }
    }
    h.min_idx = min_idx;
    h.sgn[0] = sg[0]; h.sgn[1] = sg[1]; h.sgn[2] = sg[2];
    h.block = 0; h.vox[0] = h.vox[1] = h.vox[2] = -1;
    h.normal = V(0.0f);
    h.t = -1.0f;
    if (intersection) {
        h.normal[min_idx] = (float)(-sg[min_idx]);
        h.block = get_voxel_at(S, origin, st, h.vox);
        h.t = h.block > 0 ? length(origin - initial_origin) : -1.0f;
        if (!(h.block > 0)) { h.vox[0] = h.vox[1] = h.vox[2] = -1; }
    }
    return h.t;
}

// ---- alpha-tested traversal (u_ShouldAlphaTest) -------------------------------------------------------------------------
// CalculateUV — InitialRayTraceFrag.glsl:470-512 / ShadowRayTraceFrag.glsl:333-386 on the normals the traversal forms: an exact
// axis vector, or all zeros when the ray's sign on that axis is 0 (no branch matches and the shader returns an uninitialised
// vec2; pinned to (0, 0)).
struct AlphaCtx {
    v3 cam_pos;      // u_InverseView[3].xyz
    float g_K;       // 1 / (tan(radians(u_FOV) / (2 * u_Dimensions.x)) * 2)   (main(), :421 / :419)
    float lod_bias;  // 0 in the primary shader (:202), 2 in the shadow shader (clamp(LOD - 2.0f, ...), :115)
    bool flip_x;     // the primary shader flips both texture coordinates (:198-199), the shadow shader only y (:112)
};
inline float alpha_g_K(float fov_degrees, float width) {
    const float radians = fov_degrees * 0.01745329251994329576923690768489f;  // glm::radians
    return 1.0f / ((float)std::tan((double)(radians / (2.0f * width))) * 2.0f);
}
// StopRay — InitialRayTraceFrag.glsl:189-203, ShadowRayTraceFrag.glsl:105-117
inline bool stop_ray(const Scene& S, const AlphaCtx& A, v3 P, int axis, int axis_sign, int block) {
    const int id = std::min(std::max(block, 0), 127);  // GetBlockID: floor(b / 255 * 255) == b for every byte
    if (S.s.materials[512 + id] == 0) return true;     // BlockTransparentData
    float u = 0.0f, v = 0.0f;
    if (axis_sign != 0) {
        if (axis == 1) { u = fractf(P.x); v = fractf(P.z); }
        else if (axis == 0) { u = fractf(P.z); v = fractf(P.y); }
        else { u = fractf(P.x); v = fractf(P.y); }
    }
    v = 1.0f - v;
    if (A.flip_x) u = 1.0f - u;
    const float D = length(A.cam_pos - P);  // distance(P, u_InverseView[3].xyz)
    const float l2 = (float)std::log2((double)(512.0f / (1.0f / D * A.g_K)));  // log2 pinned: correctly rounded fp32
    // int(): toward zero; NaN / out-of-range (D == 0) pinned to 0
    const int lod = (l2 > -2147483000.0f && l2 < 2147483000.0f) ? (int)l2 : 0;
    const int level = (int)clampf((float)lod - A.lod_bias, 0.0f, 8.0f);
    const int n = 512 >> level;
    size_t off = 0;
    for (int k = 0; k < level; ++k) off += (size_t)(512 >> k) * (512 >> k);
    const int i = ((int)std::floor(u * (float)n)) & (n - 1), j = ((int)std::floor(v * (float)n)) & (n - 1);
    int layer = S.s.materials[id];  // BlockAlbedoData; array layer = clamp(round(float(layer)), 0, layers - 1)
    layer = std::min(std::max(layer, 0), S.s.n_alpha_layers - 1);
    const float alpha = (float)S.s.alpha_mips[(size_t)layer * VXPT_ALPHA_MIP_TEXELS + off + (size_t)j * n + i] / 255.0f;
    return alpha > 0.975f;
}

// the one-voxel DDA step of VoxelTraversalDF(_AlphaTest) (:271-291, repeated in the inner loop :240-251); ivec3(origin) truncates
inline void dda_step(v3& origin, v3 direction, const int sg[3], int& min_idx) {
    int g[3] = {(int)origin.x, (int)origin.y, (int)origin.z};
    v3 w = V(origin.x - (float)g[0], origin.y - (float)g[1], origin.z - (float)g[2]);
    const int half[3] = {(1 + sg[0]) >> 1, (1 + sg[1]) >> 1, (1 + sg[2]) >> 1};
    v3 inv = V(1.0f / direction.x, 1.0f / direction.y, 1.0f / direction.z);
    v3 f = V(((float)half[0] - w.x) * inv.x, ((float)half[1] - w.y) * inv.y, ((float)half[2] - w.z) * inv.z);
    min_idx = (f.x < f.y && sg[0] != 0) ? ((f.x < f.z || sg[2] == 0) ? 0 : 2) : ((f.y < f.z || sg[2] == 0) ? 1 : 2);
    g[min_idx] += sg[min_idx];
    float fm = f[min_idx];
    w = w + direction * fm;
    w[min_idx] = (float)(1 - half[min_idx]);
    origin = V((float)g[0] + w.x, (float)g[1] + w.y, (float)g[2] + w.z);
    origin[min_idx] += (float)sg[min_idx] * 0.0001f;
}

// VoxelTraversalDF_AlphaTest — InitialRayTraceFrag.glsl:205-305 (cap u_RenderDistance), ShadowRayTraceFrag.glsl:119-220 (cap 350).
// Restated as written, including what the author calls its "known artifacts" (Pipeline.cpp:836): after a transparent texel lets the
// ray through and the four inner DDA steps find nothing to stop at, control falls into the `else` of `if (Euclidean == 1)` with
// Euclidean == 0 and the ray moves BACK by one direction vector (:294-297).  Block fetches are counted per GetVoxel call.
float traverse_df_alpha(const Scene& S, const AlphaCtx& A, v3 origin, v3 direction, int max_it, Hit& h, Stats& st) {
    const v3 initial_origin = origin;
    bool intersection = false;
    int min_idx = 0;
    const int sg[3] = {isign(direction.x), isign(direction.y), isign(direction.z)};
    st.rays++;
    h.sgn[0] = sg[0]; h.sgn[1] = sg[1]; h.sgn[2] = sg[2];
    h.block = 0; h.vox[0] = h.vox[1] = h.vox[2] = -1;
    h.normal = V(0.0f);
    h.t = -1.0f;
    auto finish = [&]() {  // :249-253 and :300-305
        h.min_idx = min_idx;
        h.normal = V(0.0f);
        h.normal[min_idx] = (float)(-sg[min_idx]);
        h.block = get_voxel_at(S, origin, st, h.vox);
        h.t = h.block > 0 ? length(origin - initial_origin) : -1.0f;
        if (!(h.block > 0)) { h.vox[0] = h.vox[1] = h.vox[2] = -1; }
        return h.t;
    };
    for (int itr = 0; itr < max_it; ++itr) {
        float fx = std::floor(origin.x), fy = std::floor(origin.y), fz = std::floor(origin.z);
        if (!S.in_volume_f(fx, fy, fz)) { intersection = false; break; }
        st.df++;
        float dist = (float)S.s.df[S.idx((int)fx, (int)fy, (int)fz)];
        int euclid = (int)std::floor(dist == 1.0f ? 1.0f : dist * 0.57735026918f);
        if (euclid == 0) {
            int bt = get_voxel_at(S, origin, st);
            if (stop_ray(S, A, origin, min_idx, sg[min_idx], bt)) break;
            for (int i = 0; i < 4; ++i) {
                dda_step(origin, direction, sg, min_idx);
                bt = get_voxel_at(S, origin, st);
                if (bt > 0) {
                    bt = get_voxel_at(S, origin, st);
                    if (stop_ray(S, A, origin, min_idx, sg[min_idx], bt)) return finish();
                }
            }
        }
        if (euclid == 1) {
            dda_step(origin, direction, sg, min_idx);
            intersection = true;
        } else {
            origin = origin + (float)(euclid - 1) * direction;
        }
    }
    h.min_idx = min_idx;
    if (intersection) return finish();
    return -1.0f;
}

// GetNormalID — InitialRayTraceFrag.glsl:143-185 : {+Z,-Z,+Y,-Y,-X,+X} -> 0..5
inline int normal_id_of(const Hit& h) {
    int s = -h.sgn[h.min_idx];  // normal component
    if (h.min_idx == 2) return s > 0 ? 0 : 1;
    if (h.min_idx == 1) return s > 0 ? 2 : 3;
    return s < 0 ? 4 : 5;
}
// GetNormalFromID — ShadowRayTraceFrag.glsl:317-328 (miss -> (1,1,1)); DiffuseRayTraceFrag.glsl:790-801 (miss -> 0.5)
inline v3 normal_from_id(int id, float miss) {
    switch (id) {
        case 0: return V(0, 0, 1);
        case 1: return V(0, 0, -1);
        case 2: return V(0, 1, 0);
        case 3: return V(0, -1, 0);
        case 4: return V(-1, 0, 0);
        case 5: return V(1, 0, 0);
        default: return V(miss);
    }
}

// GetRayDirectionAt — ShadowRayTraceFrag.glsl:303-308 / GetRayStuff InitialRayTraceFrag.glsl:410-413
inline v3 ray_direction_at(const VxCamera& cam, float u, float v) {
    float clip[4] = {u * 2.0f - 1.0f, v * 2.0f - 1.0f, -1.0f, 1.0f};
    float e4[4];
    mat4_mul(cam.inv_proj, clip, e4);
    float eye[4] = {e4[0], e4[1], -1.0f, 0.0f};
    float r4[4];
    mat4_mul(cam.inv_view, eye, r4);
    return V(r4[0], r4[1], r4[2]);
}
inline v3 ray_origin(const VxCamera& cam) { return V(cam.inv_view[12], cam.inv_view[13], cam.inv_view[14]); }

inline void pixel_uv(const VxCamera& cam, int i, int j, float& u, float& v) {
    u = ((float)i + 0.5f) / (float)cam.width;
    v = ((float)j + 0.5f) / (float)cam.height;
}

// ------------------------------------------------------------------------------------------------ textures
// samplerBlueNoiseErrorDistribution_128x128_OptimizedFor_2d2d2d2d_32spp — DiffuseRayTraceFrag.glsl:126-149
inline float blue_noise_1d(const Scene& S, int px, int py, int sample_index, int sample_dim) {
    int pi = px & 127, pj = py & 127;
    sample_index &= 255;
    sample_dim &= 255;
    int ridx = sample_dim + (pi + pj * 128) * 8;
    if (ridx > 131071) ridx = 131071;  // A.5: the shader reads past rankingTile here; pinned by clamping
    int ranked = sample_index ^ S.s.rank[ridx];
    ranked &= 255;  // table values are < 256, so this is a no-op kept for index safety
    int value = S.s.sobol[sample_dim + ranked * 256];
    value = value ^ S.s.scramble[(sample_dim % 8) + (pi + pj * 128) * 8];
    return (0.5f + (float)value) / 256.0f;
}

inline v3 sky_sample(const Scene& S, v3 d) {
    const int N = S.s.sky_n;
    float ax = std::fabs(d.x), ay = std::fabs(d.y), az = std::fabs(d.z);
    int face; float sc, tc, ma;
    if (ax >= ay && ax >= az) { face = d.x > 0 ? 0 : 1; sc = d.x > 0 ? -d.z : d.z; tc = -d.y; ma = ax; }
    else if (ay >= az)        { face = d.y > 0 ? 2 : 3; sc = d.x; tc = d.y > 0 ? d.z : -d.z; ma = ay; }
    else                      { face = d.z > 0 ? 4 : 5; sc = d.z > 0 ? d.x : -d.x; tc = -d.y; ma = az; }
    float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
    float u = s * (float)N - 0.5f, v = t * (float)N - 0.5f;
    float fu0 = std::floor(u), fv0 = std::floor(v);
    float fu = u - fu0, fv = v - fv0;
    int i0 = (int)fu0, j0 = (int)fv0, i1 = i0 + 1, j1 = j0 + 1;
    i0 = std::min(std::max(i0, 0), N - 1); i1 = std::min(std::max(i1, 0), N - 1);
    j0 = std::min(std::max(j0, 0), N - 1); j1 = std::min(std::max(j1, 0), N - 1);
    const float* F = S.s.sky + (size_t)face * N * N * 3;
    auto tx = [&](int i, int j) { const float* p = F + ((size_t)j * N + i) * 3; return V(p[0], p[1], p[2]); };
    v3 a = tx(i0, j0) * (1.0f - fu) + tx(i1, j0) * fu;
    v3 b = tx(i0, j1) * (1.0f - fu) + tx(i1, j1) * fu;
    return a * (1.0f - fv) + b * fv;
}

inline v3 tex_nearest4(const float* base, int layer, int n, float u, float v) {
    int i = ((int)std::floor(u * (float)n)) & (n - 1);
    int j = ((int)std::floor(v * (float)n)) & (n - 1);
    const float* p = base + (((size_t)std::max(layer, 0) * n + j) * n + i) * 4;  // GL clamps the array layer: -1 (id not in the block database) reads layer 0
    return V(p[0], p[1], p[2]);
}
inline float tex_bilinear1(const float* base, int layer, int n, float u, float v) {
    float x = u * (float)n - 0.5f, y = v * (float)n - 0.5f;
    float fx0 = std::floor(x), fy0 = std::floor(y);
    float fx = x - fx0, fy = y - fy0;
    int i0 = ((int)fx0) & (n - 1), i1 = ((int)fx0 + 1) & (n - 1);
    int j0 = ((int)fy0) & (n - 1), j1 = ((int)fy0 + 1) & (n - 1);
    const float* L = base + (size_t)layer * n * n;
    float a = L[(size_t)j0 * n + i0] * (1.0f - fx) + L[(size_t)j0 * n + i1] * fx;
    float b = L[(size_t)j1 * n + i0] * (1.0f - fx) + L[(size_t)j1 * n + i1] * fx;
    return a * (1.0f - fy) + b * fy;
}

// CalculateUV — DiffuseRayTraceFrag.glsl:1235-1273 (normal is an exact axis vector here)
inline void calc_uv(v3 p, int axis, float& u, float& v) {
    if (axis == 1) { u = fractf(p.x); v = fractf(p.z); }       // top / bottom : xz
    else if (axis == 0) { u = fractf(p.z); v = fractf(p.y); }  // left / right : zy
    else { u = fractf(p.x); v = fractf(p.y); }                 // front / back : xy
}

}  // namespace

// =============================================================================================== exports
extern "C" {

// ManhattanDistance{X,Y,Z}.comp executed in the order World::GenerateDistanceField dispatches them
// (Core/World.cpp:69-113).  Integer formulation; the R8 store/load round trip is exact.
void vxo_df_build(const uint8_t* grid, uint8_t* df, int wx, int wy, int wz) {
    const int max_d = std::min(254, wx + wy + wz);  // ManhattanDistanceX.comp:49
    auto at = [&](int x, int y, int z) -> size_t { return (size_t)x + (size_t)wx * ((size_t)y + (size_t)wy * (size_t)z); };
    // X pass — ManhattanDistanceX.comp:45-69
#pragma omp parallel for schedule(static)
    for (int z = 0; z < wz; ++z)
        for (int y = 0; y < wy; ++y) {
            df[at(0, y, z)] = grid[at(0, y, z)] > 0 ? 0 : (uint8_t)max_d;
            for (int x = 1; x < wx; ++x)
                df[at(x, y, z)] = grid[at(x, y, z)] > 0 ? 0 : (uint8_t)std::min(max_d, 1 + (int)df[at(x - 1, y, z)]);
            for (int x = wx - 2; x >= 0; --x)
                if (df[at(x + 1, y, z)] < df[at(x, y, z)]) df[at(x, y, z)] = (uint8_t)(1 + df[at(x + 1, y, z)]);
        }
    // Y pass — ManhattanDistanceY.comp:26-51
#pragma omp parallel for schedule(static)
    for (int z = 0; z < wz; ++z)
        for (int x = 0; x < wx; ++x) {
            for (int y = 1; y < wy; ++y)
                if (df[at(x, y - 1, z)] < df[at(x, y, z)]) df[at(x, y, z)] = (uint8_t)(1 + df[at(x, y - 1, z)]);
            for (int y = wy - 2; y >= 0; --y)
                if (df[at(x, y + 1, z)] < df[at(x, y, z)]) df[at(x, y, z)] = (uint8_t)(1 + df[at(x, y + 1, z)]);
        }
    // Z pass — ManhattanDistanceZ.comp:24-47
#pragma omp parallel for schedule(static)
    for (int y = 0; y < wy; ++y)
        for (int x = 0; x < wx; ++x) {
            for (int z = 1; z < wz; ++z)
                if (df[at(x, y, z - 1)] < df[at(x, y, z)]) df[at(x, y, z)] = (uint8_t)(1 + df[at(x, y, z - 1)]);
            for (int z = wz - 2; z >= 0; --z)
                if (df[at(x, y, z + 1)] < df[at(x, y, z)]) df[at(x, y, z)] = (uint8_t)(1 + df[at(x, y, z + 1)]);
        }
}

// Definition of the result: min(254, min over solid q of |p-q|_1).  O(voxels * solids): small grids.
void vxo_df_bruteforce(const uint8_t* grid, uint8_t* df, int wx, int wy, int wz) {
    std::vector<int> sx, sy, sz;
    for (int z = 0; z < wz; ++z)
        for (int y = 0; y < wy; ++y)
            for (int x = 0; x < wx; ++x)
                if (grid[(size_t)x + (size_t)wx * ((size_t)y + (size_t)wy * z)] > 0) { sx.push_back(x); sy.push_back(y); sz.push_back(z); }
    const int max_d = std::min(254, wx + wy + wz);
#pragma omp parallel for schedule(static)
    for (int z = 0; z < wz; ++z)
        for (int y = 0; y < wy; ++y)
            for (int x = 0; x < wx; ++x) {
                int best = max_d;
                for (size_t k = 0; k < sx.size(); ++k) {
                    int d = std::abs(x - sx[k]) + std::abs(y - sy[k]) + std::abs(z - sz[k]);
                    if (d < best) best = d;
                }
                df[(size_t)x + (size_t)wx * ((size_t)y + (size_t)wy * z)] = (uint8_t)best;
            }
}

// One VoxelTraversalDF call (known-answer tests).  out6 = {min_idx, sgn of that axis, block, vox x, y, z}.
float vxo_traverse(const VxoScene* sc, const float o[3], const float d[3], int max_it, int32_t out6[6], VxoStats* stats) {
    Scene S{*sc};
    Hit h; Stats st;
    float t = traverse_df(S, V(o[0], o[1], o[2]), V(d[0], d[1], d[2]), max_it, h, st);
    if (out6) { out6[0] = h.min_idx; out6[1] = h.sgn[h.min_idx]; out6[2] = h.block; out6[3] = h.vox[0]; out6[4] = h.vox[1]; out6[5] = h.vox[2]; }
    if (stats) { stats->rays += st.rays; stats->df_fetches += st.df; stats->vox_fetches += st.vox; }
    return t;
}

// Independent cross-check: plain one-voxel-at-a-time DDA (Amanatides-Woo) in double precision, no distance
// field.  After the author's unused Core/Shaders/Implementations/DDA/DDA.glsl:134-253 (not compiled by the
// reference app).  Returns 1 and the first solid voxel entered, 0 on a miss.  The start voxel is not tested.
int vxo_plain_dda(const VxoScene* sc, const float o[3], const float d[3], int max_steps, int32_t vox[3], int32_t* axis) {
    Scene S{*sc};
    double p[3] = {o[0], o[1], o[2]}, dir[3] = {d[0], d[1], d[2]};
    int g[3], step[3]; double tmax[3], tdelta[3];
    for (int k = 0; k < 3; ++k) {
        g[k] = (int)std::floor(p[k]);
        step[k] = dir[k] > 0 ? 1 : (dir[k] < 0 ? -1 : 0);
        if (step[k] != 0) {
            double nb = step[k] > 0 ? (g[k] + 1) : g[k];
            tmax[k] = (nb - p[k]) / dir[k];
            tdelta[k] = std::fabs(1.0 / dir[k]);
        } else { tmax[k] = INFINITY; tdelta[k] = INFINITY; }
    }
    const int dims[3] = {S.s.wx, S.s.wy, S.s.wz};
    for (int it = 0; it < max_steps; ++it) {
        int k = (tmax[0] < tmax[1]) ? ((tmax[0] < tmax[2]) ? 0 : 2) : ((tmax[1] < tmax[2]) ? 1 : 2);
        g[k] += step[k];
        tmax[k] += tdelta[k];
        if (g[0] < 0 || g[1] < 0 || g[2] < 0 || g[0] >= dims[0] || g[1] >= dims[1] || g[2] >= dims[2]) return 0;
        if (S.s.grid[S.idx(g[0], g[1], g[2])] > 0) {
            vox[0] = g[0]; vox[1] = g[1]; vox[2] = g[2]; *axis = k;
            return 1;
        }
    }
    return 0;
}

// InitialRayTraceFrag.glsl main() :417-468 with GetRayStuff :398-414.  SURVEY.md A.2.
int vxo_trace_primary(const VxoScene* sc, const VxCamera* cam, const VxPrimaryParams* prm, const VxGBuffer* out, VxoStats* stats) {
    Scene S{*sc};
    if (prm->alpha_test && (!sc->alpha_mips || !sc->materials)) return VXPT_E_STATE;
    const int W = cam->width, H = cam->height;
    const AlphaCtx A{ray_origin(*cam), alpha_g_K(prm->fov_degrees, (float)W), 0.0f, true};
    uint64_t rays = 0, dfc = 0, voxc = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : rays, dfc, voxc)
    for (int j = cam->row_begin; j < cam->row_end; ++j) {
        Stats st;
        for (int i = 0; i < W; ++i) {
            float u, v;
            pixel_uv(*cam, i, j, u, v);
            if (prm->jitter_enable) {
                float tsx = 1.0f / (float)W, tsy = 1.0f / (float)H;  // TexelSize = 1.0f / u_Dimensions
                u -= prm->jitter[0] * tsx;
                v -= prm->jitter[1] * tsy;
            }
            v3 rd = ray_direction_at(*cam, u, v);
            v3 ro = ray_origin(*cam);
            v3 dir = normalize(rd);
            Hit h;
            float t = prm->alpha_test ? traverse_df_alpha(S, A, ro, dir, prm->max_iterations, h, st)
                                      : traverse_df(S, ro, dir, prm->max_iterations, h, st);
            bool intersect = t > 0.0f && h.block > 0;
            size_t p = (size_t)j * W + i;
            if (out->t) out->t[p] = t;
            if (out->inv_t) out->inv_t[p] = 1.0f / t;
            if (out->normal_id) out->normal_id[p] = intersect ? (uint8_t)normal_id_of(h) : (uint8_t)VXPT_NORMAL_MISS;
            if (out->block_id) out->block_id[p] = intersect ? (uint8_t)h.block : 0;
            if (out->hit_voxel) {
                out->hit_voxel[3 * p + 0] = intersect ? (int16_t)h.vox[0] : -1;
                out->hit_voxel[3 * p + 1] = intersect ? (int16_t)h.vox[1] : -1;
                out->hit_voxel[3 * p + 2] = intersect ? (int16_t)h.vox[2] : -1;
            }
        }
        rays += st.rays; dfc += st.df; voxc += st.vox;
    }
    if (stats) { stats->rays += rays; stats->df_fetches += dfc; stats->vox_fetches += voxc; }
    return VXPT_OK;
}

// ShadowRayTraceFrag.glsl main() :414-513.  SURVEY.md A.7.
int vxo_trace_shadow(const VxoScene* sc, const VxCamera* cam, const VxGBuffer* g, const VxShadowParams* prm, const VxShadowOut* out, VxoStats* stats) {
    Scene S{*sc};
    if (prm->alpha_test && (!sc->alpha_mips || !sc->materials)) return VXPT_E_STATE;
    const int W = cam->width, H = cam->height;
    const AlphaCtx A{ray_origin(*cam), alpha_g_K(prm->fov_degrees, (float)W), 2.0f, false};
    uint64_t rays = 0, dfc = 0, voxc = 0;
    // per-frame blue-noise texel offset (:456-459): int products wrap like GLSL ints, the rest is fp32
    const int n = prm->frame % 1024;
    const int32_t ax = (int32_t)((uint32_t)n * 12664745u), ay = (int32_t)((uint32_t)n * 9560333u);
    const float offx = fractf((float)ax / 16777216.0f) * 1024.0f, offy = fractf((float)ay / 16777216.0f) * 1024.0f;
    const int ioffx = (int)std::floor(offx), ioffy = (int)std::floor(offy);
    const v3 L = V(prm->light_dir[0], prm->light_dir[1], prm->light_dir[2]);
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : rays, dfc, voxc)
    for (int j = cam->row_begin; j < cam->row_end; ++j) {
        Stats st;
        for (int i = 0; i < W; ++i) {
            size_t p = (size_t)j * W + i;
            float u, v;
            pixel_uv(*cam, i, j, u, v);
            u += prm->halton[0] * (1.0f / (float)W);
            v += prm->halton[1] * (1.0f / (float)H);
            float dist = g->t[p];
            if (dist < 0.0f) {
                if (out->shadow) out->shadow[p] = 0;
                if (out->transversal) out->transversal[p] = 64.0f;
                continue;
            }
            v3 pos = ray_origin(*cam) + normalize(ray_direction_at(*cam, u, v)) * dist;  // GetPositionAt :310-314
            v3 dir = L;
            if (prm->soft) {
                // ivec2(gl_FragCoord.xy + ivec2(floor(off))) % textureSize : float add, truncation, modulo
                int tx = ((int)(((float)i + 0.5f) + (float)ioffx)) % 256;
                int ty = ((int)(((float)j + 0.5f) + (float)ioffy)) % 256;
                const uint8_t* tex = S.s.shadow_noise + ((size_t)ty * 256 + tx) * 4;
                float xi_x = (float)tex[0] / 255.0f, xi_y = (float)tex[1] / 255.0f;
                v3 T = normalize(cross(L, V(0.0f, 1.0f, 1.0f)));
                v3 B = cross(T, L);
                // SampleCone :388-398
                const float cos_theta_max = 0.9999505604617f;
                float cos_theta = (1.0f - xi_x) + xi_x * cos_theta_max;
                float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
                float phi = xi_y * PI_F * 2.0f;
                v3 c = V(sin_theta * cos_cr(phi), sin_theta * sin_cr(phi), cos_theta);
                dir = (T * c.x + B * c.y) + L * c.z;  // mat3(T,B,L) * c
            }
            v3 N = normal_from_id(g->normal_id[p], 1.0f);
            float ndotl = dot(N, dir);
            if (ndotl <= 0.01f) {
                if (out->shadow) out->shadow[p] = 1;
                if (out->transversal) out->transversal[p] = 1.0f / 100.0f;
                continue;
            }
            v3 bias = N * V(0.06f);
            v3 o = pos + bias;
            int block_at = get_voxel_at(S, o, st);
            float T = -1.0f;
            if (dist > 0.0f) {
                Hit h;
                T = prm->alpha_test ? traverse_df_alpha(S, A, o, dir, 350, h, st) : traverse_df(S, o, dir, 350, h, st);
            }
            if (out->shadow) out->shadow[p] = (T > 0.0f || block_at > 0) ? 1 : 0;
            float tr = clampf(T / 100.0f, 0.00001f, 196.0f);
            if (T < 0.0f) tr = 4.25f / 100.0f;
            if (out->transversal) out->transversal[p] = tr;
        }
        rays += st.rays; dfc += st.df; voxc += st.vox;
    }
    if (stats) { stats->rays += rays; stats->df_fetches += dfc; stats->vox_fetches += voxc; }
    return VXPT_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ diffuse GI
namespace {

struct GiFrame {
    v3 light_color, stronger_dir;
    bool moon_stronger;
    float emissivity_mult;
    const VxDiffuseParams* p;
};

struct GiPixel {
    int px, py;
    int bl_sample;  // CurrentBLSample, DiffuseRayTraceFrag.glsl:809
};

// InverseSchlick — DiffuseRayTraceFrag.glsl:1305-1308
inline float inverse_schlick(float f0, float voh) {
    return 1.0f - clampf(f0 + (1.0f - f0) * pow_cr(1.0f - voh, 5.0f), 0.0f, 1.0f);
}
// DiffuseHammon — DiffuseRayTraceFrag.glsl:1311-1332 (rcp(x) restated as 1.0f/x)
inline float diffuse_hammon(v3 n, v3 view, v3 light, float rough) {
    float ndl = std::fmax(dot(n, light), 0.0f);
    if (ndl <= 0.0f) return 0.0f;
    float ndv = std::fmax(dot(n, view), 0.0f);
    float ldv = std::fmax(dot(light, view), 0.0f);
    v3 hw = normalize(view + light);
    float ndh = std::fmax(dot(n, hw), 0.0f);
    float facing = ldv * 0.5f + 0.5f;
    float single_rough = facing * (0.9f - 0.4f * facing) * ((0.5f + ndh) * (1.0f / std::fmax(ndh, 0.02f)));
    float single_smooth = 1.05f * inverse_schlick(0.0f, ndl) * inverse_schlick(0.0f, std::fmax(ndv, 0.0f));
    float single = clampf(mixf(single_smooth, single_rough, rough) * (1.0f / PI_F), 0.0f, 1.0f);
    float multi = 0.1159f * rough;
    return clampf((multi + single) * ndl, 0.0f, 1.0f);
}

// SampleBlueNoise2D + cosWeightedRandomHemisphereDirection — DiffuseRayTraceFrag.glsl:811-818, 945-967
inline v3 cos_hemisphere(const Scene& S, const GiFrame& F, GiPixel& px, v3 n) {
    int idx = F.p->frame % 128;  // u_CurrentFrameMod128
    float rx_ = blue_noise_1d(S, px.px, px.py, idx, 1 + px.bl_sample);
    float ry_ = blue_noise_1d(S, px.px, px.py, idx, 2 + px.bl_sample);
    px.bl_sample += 2;
    const float PI2 = 2.0f * PI_F;
    v3 uu = normalize(cross(n, V(0.0f, 1.0f, 1.0f)));
    v3 vv = cross(uu, n);
    float ra = std::sqrt(ry_);
    float rx = ra * cos_cr(PI2 * rx_);
    float ry = ra * sin_cr(PI2 * rx_);
    float rz = std::sqrt(1.0f - ry_);
    v3 rr = (rx * uu + ry * vv) + rz * n;
    return normalize(rr);
}

// GetSkyColorAt — DiffuseRayTraceFrag.glsl:984-988
inline v3 sky_color_at(const Scene& S, v3 rd) {
    rd.y = clampf(rd.y, 0.125f, 1.5f);
    return sky_sample(S, rd);
}

// CalculateDiffuse — DiffuseRayTraceFrag.glsl:535-664
inline void calculate_diffuse(const Scene& S, const GiFrame& F, GiPixel& px, v3 initial_origin, v3 input_normal,
                              v3& out_rad, float& out_ao, v3& odir, bool& skyhit, Stats& st) {
    skyhit = false;
    const float bias = 0.06f;
    v3 ro = initial_origin + input_normal * bias;
    v3 rd = cos_hemisphere(S, F, px, input_normal);
    float ao = 1.0f;
    v3 contrib = V(0.0f), thr = V(1.0f);
    const int32_t* M = S.s.materials;
    for (int i = 0; i < 2; ++i) {  // MAX_BOUNCE_LIMIT :17
        if (i == 0) odir = rd;
        Hit h;
        float T = traverse_df(S, ro, rd, F.p->trace_length, h, st);
        int tex_ref = std::min(std::max(h.block, 0), 127);
        bool intersect = T > 0.0f;
        v3 ipos = ro + (rd * T);
        if (intersect && h.block > 0) {
            float tu, tv;
            calc_uv(ipos, h.min_idx, tu, tv);
            int albedo_layer = M[tex_ref], emissive_layer = M[384 + tex_ref];
            v3 albedo = tex_nearest4(S.s.albedo_lod3, albedo_layer, 64, tu, tv);
            v3 pbr = tex_nearest4(S.s.pbr_lod2, albedo_layer, 128, tu, tv);  // sic: albedo layer (:578)
            float emis = 0.0f;
            if ((float)emissive_layer >= 0.0f) {
                float se = tex_bilinear1(S.s.emissive_lod0, emissive_layer, 512, tu, tv);
                emis = se * F.emissivity_mult * F.p->light_intensity;
            }
            float ndl = std::fmax(dot(h.normal, F.stronger_dir), 0.0f);
            v3 bias_shadow = h.normal * 0.045f;
            float shadow_at;
            if (F.moon_stronger) shadow_at = 1.0f;
            else if (ndl < 0.001f) shadow_at = 0.0f;
            else {  // GetShadowAt :1202-1222 (u_APPLY_PLAYER_SHADOW = false)
                Hit hs;
                float Ts = traverse_df(S, ipos + bias_shadow, F.stronger_dir, 128, hs, st);
                shadow_at = Ts > 0.0f ? 1.0f : 0.0f;
            }
            v3 emis_color = (emis * mixf(1.0f, 1.0f, F.p->sun_visibility)) * albedo;
            // SunBRDF :499-508 (CAUSTICS = false)
            v3 sunbrdf = (((albedo * diffuse_hammon(h.normal, -rd, F.stronger_dir, pbr.x)) * (F.light_color * 3.5f)) *
                          (1.0f - shadow_at)) * PI_F;
            v3 new_dir = cos_hemisphere(S, F, px, h.normal);
            float cos_theta = clampf(dot(h.normal, new_dir), 0.0f, 1.0f);
            float pdf = std::fmax(cos_theta / PI_F, 0.00001f);
            v3 atten = V(1.0f) * diffuse_hammon(h.normal, -rd, new_dir, pbr.x);  // DiffuseRayBRDF :527-531
            contrib = contrib + thr * sunbrdf;
            contrib = contrib + emis_color * thr;
            thr = thr * ((albedo * atten) / pdf);
            rd = new_dir;
            ro = ipos + h.normal * bias;
        } else {
            float x = mixf(1.0f, 1.05f, F.p->sun_visibility);
            x = clampf(x * 1.0f * F.p->gi_sky_strength, 0.0f, 5.0f);
            v3 sky = sky_color_at(S, rd) * x;
            contrib = contrib + sky * thr;
            skyhit = true;
            break;
        }
        if (i == 0) {
            const float dao = 2.0f;
            if (T < dao && T > 0.0f) ao = std::fmax(T / dao, 0.0f);
        }
    }
    out_rad = contrib;
    out_ao = ao;
}

// IrridianceToSH — DiffuseRayTraceFrag.glsl:766-784
inline void irradiance_to_sh(v3 rad, v3 dir, float out[6]) {
    float Co = rad.x - rad.z;
    float T = rad.z + Co * 0.5f;
    float Cg = rad.y - T;
    float Y = std::fmax(T + Cg * 0.5f, 0.0f);
    float L00 = 0.282095f;
    float L1_1 = 0.488603f * dir.y, L10 = 0.488603f * dir.z, L11 = 0.488603f * dir.x;
    out[0] = std::fmax(L11 * Y, -100.0f);
    out[1] = std::fmax(L1_1 * Y, -100.0f);
    out[2] = std::fmax(L10 * Y, -100.0f);
    out[3] = std::fmax(L00 * Y, -100.0f);
    out[4] = Co;
    out[5] = Cg;
}

}  // namespace

extern "C" {

// DiffuseRayTraceFrag.glsl main() :822-935.  SURVEY.md A.4.
int vxo_trace_diffuse(const VxoScene* sc, const VxCamera* cam, const VxGBuffer* g, const VxDiffuseParams* prm, const VxDiffuseOut* out, VxoStats* stats) {
    Scene S{*sc};
    if (!prm->use_blue_noise || prm->direct_sampling) return VXPT_E_UNSUPPORTED;
    const int W = cam->width, H = cam->height;
    GiFrame F;
    F.p = prm;
    const v3 sun = V(prm->sun_dir[0], prm->sun_dir[1], prm->sun_dir[2]);
    const v3 moon = V(prm->moon_dir[0], prm->moon_dir[1], prm->moon_dir[2]);
    const bool sun_stronger = -sun.y < 0.01f;
    const v3 SUN_COLOR = (V(192.0f, 216.0f, 255.0f) / 255.0f) * 16.0f;   // :184
    const v3 NIGHT_COLOR = (V(96.0f, 192.0f, 255.0f) / 255.0f) * 1.5f;  // :185
    const v3 DUSK_COLOR = (V(96.0f, 192.0f, 255.0f) / 255.0f) * 0.9f;   // :186
    float dusk = clampf(pow_cr(std::fabs(sun.y - 1.0f), 2.9f), 0.0f, 1.0f);
    v3 sun_color = mix3(SUN_COLOR, DUSK_COLOR, dusk);
    F.light_color = sun_stronger ? sun_color : NIGHT_COLOR;
    F.light_color = F.light_color * (0.4f * prm->gi_sun_strength);
    F.stronger_dir = sun_stronger ? sun : moon;
    F.moon_stronger = !sun_stronger;
    F.emissivity_mult = F.moon_stronger ? 13.0f : 12.0f;
    uint64_t rays = 0, dfc = 0, voxc = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : rays, dfc, voxc)
    for (int j = cam->row_begin; j < cam->row_end; ++j) {
        Stats st;
        for (int i = 0; i < W; ++i) {
            size_t p = (size_t)j * W + i;
            float u, v;
            pixel_uv(*cam, i, j, u, v);
            const float u0 = u, v0 = v;  // v_TexCoords / v_RayDirection are not jittered
            if (prm->supersample) {
                u += (prm->halton[0] * 0.75f) / (float)W;
                v += (prm->halton[1] * 0.75f) / (float)H;
            }
            float o_sh[4], o_cocg[2], o_util = 0.0f, o_ao[2] = {1.0f, 0.0f};
            float dist = g->t[p];
            v3 normal = normal_from_id(g->normal_id[p], 0.5f);
            if (dist < 0.0f) {
                float sh[6];
                v3 vdir = normalize(ray_direction_at(*cam, u0, v0));
                irradiance_to_sh(sky_sample(S, vdir) * 2.66f, normal, sh);
                o_sh[0] = sh[0]; o_sh[1] = sh[1]; o_sh[2] = sh[2]; o_sh[3] = sh[3];
                o_cocg[0] = sh[4]; o_cocg[1] = sh[5];
            } else {
                v3 pos = ray_origin(*cam) + normalize(ray_direction_at(*cam, u, v)) * dist;
                int spp = std::min(std::max(prm->spp, 1), 32);
                if (prm->checkerboard) {
                    // int(gl_FragCoord.x + gl_FragCoord.y) = i + j + 1
                    bool checker = ((int)(((float)i + 0.5f) + ((float)j + 0.5f))) % 2 == prm->frame % 2;
                    spp = (int)mixf((float)prm->spp, (float)prm->checker_spp, checker ? 1.0f : 0.0f);
                }
                spp = std::min(std::max(spp, 1), 32);
                if (F.moon_stronger) spp *= 2;
                GiPixel px{i, j, 0};
                float tot[4] = {0, 0, 0, 0}, cocg[2] = {0, 0}, acc_ao = 0.0f, skyhits = 0.0f;
                v3 radiance = V(0.0f);
                for (int s = 0; s < spp; ++s) {
                    v3 rad, d = V(0.0f);
                    float ao;
                    bool ss = false;
                    calculate_diffuse(S, F, px, pos, normal, rad, ao, d, ss, st);
                    rad = clamp3(rad, 0.0f, 8.0f);
                    radiance = radiance + rad;
                    acc_ao += ao;
                    float sh[6];
                    irradiance_to_sh(rad, d, sh);
                    tot[0] += sh[0]; tot[1] += sh[1]; tot[2] += sh[2]; tot[3] += sh[3];
                    cocg[0] += sh[4]; cocg[1] += sh[5];
                    skyhits += ss ? 1.0f : 0.0f;
                }
                const float fs = (float)spp;
                acc_ao /= fs;
                for (int k = 0; k < 4; ++k) tot[k] /= fs;
                cocg[0] /= fs; cocg[1] /= fs;
                radiance = radiance / fs;
                skyhits /= fs;
                float lum = dot(radiance, V(0.299f, 0.587f, 0.114f));
                o_util = std::fmax(lum, 0.01f);
                o_ao[0] = clampf(acc_ao, 0.0f, 1.0f);
                o_ao[1] = clampf(skyhits, 0.0f, 1.0f);
                for (int k = 0; k < 4; ++k) o_sh[k] = clampf(tot[k], -100.0f, 100.0f);
                o_cocg[0] = clampf(cocg[0], -100.0f, 100.0f);
                o_cocg[1] = clampf(cocg[1], -100.0f, 100.0f);
                o_util = clampf(o_util, 0.001f, 64.0f);
            }
            if (out->sh) std::memcpy(out->sh + 4 * p, o_sh, sizeof o_sh);
            if (out->cocg) std::memcpy(out->cocg + 2 * p, o_cocg, sizeof o_cocg);
            if (out->luma) out->luma[p] = o_util;
            if (out->ao_sky) std::memcpy(out->ao_sky + 2 * p, o_ao, sizeof o_ao);
        }
        rays += st.rays; dfc += st.df; voxc += st.vox;
    }
    if (stats) { stats->rays += rays; stats->df_fetches += dfc; stats->vox_fetches += voxc; }
    return VXPT_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ reflections
namespace {

inline float tex_nearest1(const float* base, int layer, int n, float u, float v) {
    int i = ((int)std::floor(u * (float)n)) & (n - 1);
    int j = ((int)std::floor(v * (float)n)) & (n - 1);
    return base[((size_t)layer * n + j) * n + i];
}
inline void tex_nearest4w(const float* base, int layer, int n, float u, float v, float out[4]) {
    int i = ((int)std::floor(u * (float)n)) & (n - 1);
    int j = ((int)std::floor(v * (float)n)) & (n - 1);
    const float* p = base + (((size_t)std::max(layer, 0) * n + j) * n + i) * 4;
    out[0] = p[0]; out[1] = p[1]; out[2] = p[2]; out[3] = p[3];
}
inline float log_cr(float x) { return (float)std::log((double)x); }

// CalculateVectors — ReflectionTraceFrag.glsl:1376-1449 (normal is an exact axis vector)
inline void calc_vectors(v3 p, int nid, v3& tangent, v3& bitangent, float& u, float& v) {
    if (nid <= 1) { u = fractf(p.x); v = fractf(p.y); tangent = V(1, 0, 0); bitangent = V(0, 1, 0); }
    else if (nid <= 3) { u = fractf(p.x); v = fractf(p.z); tangent = V(1, 0, 0); bitangent = V(0, 0, 1); }
    else { u = fractf(p.z); v = fractf(p.y); tangent = V(0, 0, -1); bitangent = V(0, -1, 0); }
}

// capIntersect — ReflectionTraceFrag.glsl:1264-1292 ; GetPlayerIntersect :1301-1307
inline float cap_intersect(v3 ro, v3 rd, v3 pa, v3 pb, float r) {
    v3 ba = pb - pa, oa = ro - pa;
    float baba = dot(ba, ba), bard = dot(ba, rd), baoa = dot(ba, oa), rdoa = dot(rd, oa), oaoa = dot(oa, oa);
    float a = baba - bard * bard;
    float b = baba * rdoa - baoa * bard;
    float c = baba * oaoa - baoa * baoa - r * r * baba;
    float h = b * b - a * c;
    if (h >= 0.0f) {
        float t = (-b - std::sqrt(h)) / a;
        float y = baoa + t * bard;
        if (y > 0.0f && y < baba) return t;
        v3 oc = (y <= 0.0f) ? oa : ro - pb;
        b = dot(rd, oc);
        c = dot(oc, oc) - r * r;
        h = b * b - c;
        if (h > 0.0f) return -b - std::sqrt(h);
    }
    return -1.0f;
}
inline bool player_intersect(v3 viewer, v3 pos, v3 d) {
    const float x = 0.4f;
    v3 vp = viewer + V(-x, -x, +x);
    return cap_intersect(pos, d, vp, vp + V(0.0f, 1.0f, 0.0f), 0.5f) > 0.0f;
}

// ImportanceSampleGGX — ReflectionTraceFrag.glsl:345-365
inline v3 importance_sample_ggx(v3 N, float roughness, float xi_x, float xi_y) {
    float alpha = roughness * roughness;
    float alpha2 = alpha * alpha;
    float phi = 2.0f * PI_F * xi_x;
    float cos_theta = std::sqrt((1.0f - xi_y) / (1.0f + (alpha2 - 1.0f) * xi_y));
    float sin_theta = std::sqrt(1.0f - cos_theta * cos_theta);
    v3 H = V(cos_cr(phi) * sin_theta, sin_cr(phi) * sin_theta, cos_theta);
    v3 up = std::fabs(N.z) < 0.999f ? V(0.0f, 0.0f, 1.0f) : V(1.0f, 0.0f, 0.0f);
    v3 tangent = normalize(cross(up, N));
    v3 bitangent = cross(N, tangent);
    v3 sv = (tangent * H.x + bitangent * H.y) + N * H.z;
    return normalize(sv);
}

// Cook-Torrance sun term — ReflectionTraceFrag.glsl:282-336 (the specular part is multiplied by radiance * 0.05 * 0)
inline v3 directional_light(v3 viewer, v3 world_pos, v3 light_dir, v3 radiance, v3 albedo, v3 normal, v3 pbr, float shadow) {
    const float Epsilon = 0.00001f;
    float Shadow = std::fmin(shadow, 1.0f);
    v3 Lo = normalize(viewer - world_pos);
    v3 N = normal;
    float cosLo = std::fmax(0.0f, dot(N, Lo));
    v3 F0 = mix3(V(0.04f), albedo, pbr.y);
    v3 Li = light_dir;
    v3 Lh = normalize(Li + Lo);
    float cosLi = std::fmax(0.0f, dot(N, Li));
    float cosLh = std::fmax(0.0f, dot(N, Lh));
    float ct = std::fmax(0.0f, dot(Lh, Lo));
    v3 F = F0 + (V(1.0f) - F0) * pow_cr(1.0f - ct, 5.0f);  // fresnelSchlick
    float alpha = pbr.x * pbr.x, alphaSq = alpha * alpha;   // ndfGGX
    float denom = (cosLh * cosLh) * (alphaSq - 1.0f) + 1.0f;
    float D = alphaSq / (PI_F * denom * denom);
    float rr = pbr.x + 1.0f;                                 // gaSchlickGGX
    float k = (rr * rr) / 8.0f;
    float G = (cosLi / (cosLi * (1.0f - k) + k)) * (cosLo / (cosLo * (1.0f - k) + k));
    v3 kd = mix3(V(1.0f) - F, V(0.0f), pbr.y);
    v3 diffuseBRDF = kd * albedo;
    v3 specularBRDF = ((F * D) * G) / std::fmax(Epsilon, 4.0f * cosLi * cosLo);
    v3 radiance_s = (radiance * 0.05f) * 0.0f;
    v3 res = ((diffuseBRDF * radiance) * cosLi) + ((specularBRDF * radiance_s) * cosLi);
    res = V(std::fmax(res.x, 0.0f), std::fmax(res.y, 0.0f), std::fmax(res.z, 0.0f));
    return res * clampf(1.0f - Shadow, 0.0f, 1.0f);
}

// TemperatureToRGB(5778) — ReflectionTraceFrag.glsl:1358-1375 with SRGBToLinear :1348-1350
inline float srgb_to_linear_f(float x) { return x > 0.04045f ? pow_cr(x * (1.0f / 1.055f) + 0.0521327f, 2.4f) : x / 12.92f; }
inline v3 temperature_to_rgb(float kelvin) {
    v3 c;
    float t = clampf(kelvin, 1000.0f, 50000.0f) / 100.0f;
    if (t <= 66.0f) {
        c.x = 1.0f;
        c.y = clampf(0.39008157876901960784f * log_cr(t) - 0.63184144378862745098f, 0.0f, 1.0f);
    } else {
        float u = t - 60.0f;
        c.x = clampf(1.29293618606274509804f * pow_cr(u, -0.1332047592f), 0.0f, 1.0f);
        c.y = clampf(1.12989086089529411765f * pow_cr(u, -0.0755148492f), 0.0f, 1.0f);
    }
    if (t >= 66.0f) c.z = 1.0f;
    else if (t <= 19.0f) c.z = 0.0f;
    else c.z = clampf(0.54320678911019607843f * log_cr(t - 10.0f) - 1.19625408914f, 0.0f, 1.0f);
    return V(srgb_to_linear_f(c.x), srgb_to_linear_f(c.y), srgb_to_linear_f(c.z));
}

}  // namespace

extern "C" {

// ReflectionTraceFrag.glsl main() :717-1038 in the v1 parity profile.
int vxo_trace_reflection(const VxoScene* sc, const VxCamera* cam, const VxGBuffer* g, const VxReflectionIn* in, const VxReflectionParams* prm,
                         const VxReflectionOut* out, VxoStats* stats) {
    Scene S{*sc};
    const int W = cam->width, H = cam->height;
    const v3 sun = V(prm->sun_dir[0], prm->sun_dir[1], prm->sun_dir[2]);
    const v3 moon = V(prm->moon_dir[0], prm->moon_dir[1], prm->moon_dir[2]);
    const v3 stronger = V(prm->stronger_dir[0], prm->stronger_dir[1], prm->stronger_dir[2]);
    const v3 viewer = V(prm->viewer_pos[0], prm->viewer_pos[1], prm->viewer_pos[2]);
    // per-frame colours (:648-668, :727-731)
    v3 sun_color = (((sky_sample(S, sun) * temperature_to_rgb(5778.0f)) * PI_F) * 2.2f) * prm->sun_strength;
    v3 moon_color = (sky_sample(S, moon) * PI_F) * prm->moon_strength;
    {
        float lum = dot(moon_color, V(0.2125f, 0.7154f, 0.0721f));
        moon_color = mix3(V(lum), moon_color, 1.3f);
        moon_color = (moon_color * 0.42525f) * prm->moon_strength;
    }
    float sun_vis = clampf(dot(sun, V(0.0f, 1.0f, 0.0f)) + 0.05f, 0.0f, 0.1f) * 12.0f;
    sun_vis = 1.0f - sun_vis;
    const v3 color_mixed = mix3(sun_color, moon_color, sun_vis);
    const int bn_index = prm->frame >= 0 ? prm->frame % 128 : 100;
    const int32_t* M = S.s.materials;
    uint64_t rays = 0, dfc = 0, voxc = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : rays, dfc, voxc)
    for (int j = cam->row_begin; j < cam->row_end; ++j) {
        Stats st;
        for (int i = 0; i < W; ++i) {
            size_t p = (size_t)j * W + i;
            float o_color[4] = {0, 0, 0, 0}, o_hit = -1.0f;
            uint8_t o_mask = 0;
            float u, v;
            pixel_uv(*cam, i, j, u, v);
            const float ju = u + (clampf(prm->halton[0], -2.0f, 2.0f) / (float)W) * 1.0f;  // u_TemporalFilterReflections = true
            const float jv = v + (clampf(prm->halton[1], -2.0f, 2.0f) / (float)H) * 1.0f;
            // texture(u_PositionTexture, JitteredUV) :766 — attachment 0 of the primary FBO is GL_LINEAR, attachment 1 (the normal id) GL_NEAREST,
            // both GL_REPEAT (Core/Pipeline.cpp:1094, Core/GLClasses/Framebuffer.cpp:64-67); filter arithmetic as pinned in glsl_compat.h
            float dist;
            int nid;
            {
                const float x = ju * (float)W - 0.5f, y = jv * (float)H - 0.5f;
                const float fx0 = std::floor(x), fy0 = std::floor(y);
                const float fx = x - fx0, fy = y - fy0;
                const int i0 = wrap_repeat((int)fx0, W), i1 = wrap_repeat((int)fx0 + 1, W), j0 = wrap_repeat((int)fy0, H), j1 = wrap_repeat((int)fy0 + 1, H);
                const float a = g->t[(size_t)j0 * W + i0] * (1.0f - fx) + g->t[(size_t)j0 * W + i1] * fx;
                const float b = g->t[(size_t)j1 * W + i0] * (1.0f - fx) + g->t[(size_t)j1 * W + i1] * fx;
                dist = a * (1.0f - fy) + b * fy;
                const int ni = wrap_repeat((int)std::floor(ju * (float)W), W), nj = wrap_repeat((int)std::floor(jv * (float)H), H);
                nid = g->normal_id[(size_t)nj * W + ni];
            }
            if (!(dist < 0.0f)) {
                int spp = std::min(std::max(prm->spp, 1), 16);
                if (prm->checkerboard) {
                    bool checker = ((int)(((float)i + 0.5f) + ((float)j + 0.5f))) % 2 == (prm->frame % 2);
                    spp = (int)mixf((float)prm->spp, (float)((prm->spp + prm->spp % 2) / 2), checker ? 1.0f : 0.0f);
                }
                spp = std::min(std::max(spp, 1), 16);
                v3 pos = ray_origin(*cam) + normalize(ray_direction_at(*cam, ju, jv)) * dist;
                const v3 face_n = normal_from_id(nid, 1.0f);
                float roughness_at, metalness_at;
                if (in->g_pbr) { roughness_at = in->g_pbr[4 * p + 0]; metalness_at = in->g_pbr[4 * p + 1]; }
                else {  // stand-in for the G-buffer material pass: level-2 PBR texel of the block at the primary hit
                    v3 tg, bt; float tu, tv;
                    calc_vectors(pos, nid, tg, bt, tu, tv);
                    tv = 1.0f - tv;
                    float t4[4];
                    tex_nearest4w(S.s.pbr_lod2, M[256 + std::min<int>(g->block_id[p], 127)], 128, tu, tv, t4);
                    roughness_at = t4[0]; metalness_at = t4[1];
                }
                v3 I = normalize(pos - viewer);
                pos = pos + face_n * 0.035f;
                v3 nmapped = in->g_normal ? V(in->g_normal[3 * p], in->g_normal[3 * p + 1], in->g_normal[3 * p + 2]) : face_n;
                float sh4[4] = {in->sh[4 * p], in->sh[4 * p + 1], in->sh[4 * p + 2], in->sh[4 * p + 3]};
                float cg[2] = {in->cocg[2 * p], in->cocg[2 * p + 1]};
                v3 base_indirect;
                {  // SHToIrradianceA :469-478
                    float Y = std::fmax(0.0f, 3.544905f * sh4[3]);
                    float sc2 = (Y * 0.282095f) / (sh4[3] + 1e-6f);
                    float c0 = cg[0] * sc2, c1 = cg[1] * sc2;
                    float T = Y - c1 * 0.5f, G = c1 + T, B = T - c0 * 0.5f, R = B + c0;
                    base_indirect = V(std::fmax(R, 0.0f), std::fmax(G, 0.0f), std::fmax(B, 0.0f));
                }
                const float rough_bias = mixf(1.0f, 0.85f, prm->roughness_bias ? 1.0f : 0.0f);
                float computed_shadow = 0.0f;
                int shadow_itr = 0, bl = 0, total_hits = 0;
                float tot[4] = {0, 0, 0, 0}, avg_hit = 0.001f, meaningful = 0.0f, mask = 0.0f;
                for (int s = 0; s < spp; ++s) {
                    v3 rn = nmapped;
                    if (prm->rough) {  // GetReflectionDirection :621-644
                        float R = std::fmax(clampf(roughness_at * rough_bias, 0.01f, 1.0f), 0.05f);
                        float nearest = -100.0f;
                        v3 best = V(0.0f);
                        for (int k = 0; k < 3; ++k) {
                            float xx = blue_noise_1d(S, i, j, bn_index, 1 + bl);
                            float xy = blue_noise_1d(S, i, j, bn_index, 2 + bl);
                            bl += 2; bl = bl % 128;
                            v3 smp = importance_sample_ggx(nmapped, R, xx * 0.9f, xy * 0.65f);
                            float d = dot(smp, nmapped);
                            if (d > nearest) { best = smp; nearest = d; }
                        }
                        rn = best;
                    }
                    v3 R = I - rn * (2.0f * dot(rn, I));  // reflect(I, N)
                    Hit h;
                    float T = traverse_df(S, pos, R, prm->trace_length, h, st);
                    v3 hit_pos = pos + (R * T);
                    if (T > 0.0f) {
                        const int hnid = normal_id_of(h);
                        v3 tg, bt; float tu, tv;
                        calc_vectors(hit_pos, hnid, tg, bt, tu, tv);
                        tv = 1.0f - tv;
                        const int ref = std::min(std::max(h.block, 0), 127);
                        int t_albedo = M[ref], t_normal = M[128 + ref], t_pbr = M[256 + ref], t_emis = M[384 + ref];
                        if (ref == prm->grass_props[0]) {
                            if (hnid == 4 || hnid == 5 || hnid == 0 || hnid == 1) { t_albedo = prm->grass_props[4]; t_normal = prm->grass_props[5]; t_pbr = prm->grass_props[6]; }
                            else if (hnid == 2) { t_albedo = prm->grass_props[1]; t_normal = prm->grass_props[2]; t_pbr = prm->grass_props[3]; }
                            else { t_albedo = prm->grass_props[7]; t_normal = prm->grass_props[8]; t_pbr = prm->grass_props[9]; }
                        }
                        v3 ambient = base_indirect;
                        v3 albedo = tex_nearest4(S.s.albedo_lod3, t_albedo, 64, tu, tv);
                        v3 radiance = color_mixed * 0.6f;
                        float pbr4[4];
                        tex_nearest4w(S.s.pbr_lod2, t_pbr, 128, tu, tv, pbr4);
                        float AO = pow_cr(pbr4[3], 2.0f);
                        bool player_shadow = player_intersect(viewer, hit_pos + h.normal * 0.035f, stronger);
                        if (shadow_itr < std::max(spp / 4, 1)) {
                            if (!player_shadow) {  // GetShadowAt :1327-1346
                                v3 so = hit_pos + h.normal * 0.055f;
                                if (player_intersect(viewer, so, stronger)) computed_shadow = 1.0f;
                                else {
                                    Hit hs;
                                    float Ts = traverse_df(S, so, stronger, 150, hs, st);
                                    computed_shadow = Ts > 0.0f ? 1.0f : 0.0f;
                                }
                            } else computed_shadow = 1.0f;
                            shadow_itr = shadow_itr + 1;
                        }
                        ambient = ((ambient * 1.0f) * clampf(AO, 0.1f, 1.0f)) * albedo;
                        v3 nm = tex_nearest4(S.s.normal_lod3, t_normal, 64, tu, tv) * 2.0f - V(1.0f);
                        v3 nmap = (tg * nm.x + bt * nm.y) + h.normal * nm.z;  // TBN * n
                        v3 direct = ambient + directional_light(viewer, hit_pos, stronger, radiance, albedo, nmap, V(pbr4[0], pbr4[1], pbr4[2]), computed_shadow);
                        if ((float)t_emis > -0.5f) {
                            float e = tex_nearest1(S.s.emissive_lod2, t_emis, 128, tu, tv);
                            if (e > 0.1f) {
                                const float lbx = 0.02501f, lby = 0.03001f;
                                e *= (tu > lbx && tu < 1.0f - lbx && tv > lby && tv < 1.0f - lby) ? 1.0f : 0.0f;
                                direct = albedo * std::fmax((e * 19.0f) * 1.0f, 2.0f);
                                mask = 1.0f;
                            }
                        }
                        tot[0] += direct.x; tot[1] += direct.y; tot[2] += direct.z; tot[3] += 1.0f;
                        avg_hit += T;
                        meaningful += 1.0f;
                    } else {
                        v3 atmo = sky_sample(S, normalize(R));
                        float m = mixf(1.0f, 1.175f, metalness_at > 0.05f ? 1.0f : 0.0f);
                        tot[0] += atmo.x * m; tot[1] += atmo.y * m; tot[2] += atmo.z * m; tot[3] += 1.0f;
                    }
                    total_hits++;
                }
                avg_hit /= std::fmax(meaningful, 0.01f);
                for (int k = 0; k < 4; ++k) tot[k] /= (float)total_hits;
                for (int k = 0; k < 4; ++k) o_color[k] = clampf(tot[k], 0.0000001f, 100.0f);
                o_hit = clampf(meaningful > 0.01f ? avg_hit : -1.0f, -10.0f, 200.0f);
                o_mask = clampf(mask, 0.0f, 1.0f) > 0.5f ? 1 : 0;
            }
            if (out->color) std::memcpy(out->color + 4 * p, o_color, sizeof o_color);
            if (out->hit_distance) out->hit_distance[p] = o_hit;
            if (out->emissive_mask) out->emissive_mask[p] = o_mask;
        }
        rays += st.rays; dfc += st.df; voxc += st.vox;
    }
    if (stats) { stats->rays += rays; stats->df_fetches += dfc; stats->vox_fetches += voxc; }
    return VXPT_OK;
}

// ------------------------------------------------------------------------------------------------ other consumers of the distance field
// A batch of VoxelTraversalDF calls on caller-supplied rays (vxpt_trace_rays).  Core/Shaders/PostProcessingVert.glsl:46-53 traces one such
// ray per frame (camera -> sun, cap 350; its copy of the function, :103-170, is InitialRayTraceFrag.glsl:307-374 word for word apart
// from the literal cap — tests/test_df_consumers.py compares the two texts when the reference tree is present).
int vxo_trace_rays(const VxoScene* sc, const float* origins, const float* directions, int n, int max_it, float* t, uint8_t* normal_id,
                   uint8_t* block_id, int16_t* hit_voxel, VxoStats* stats) {
    Scene S{*sc};
    Stats st;
    for (int k = 0; k < n; ++k) {
        Hit h;
        const float T = traverse_df(S, V(origins[3 * k], origins[3 * k + 1], origins[3 * k + 2]),
                                    V(directions[3 * k], directions[3 * k + 1], directions[3 * k + 2]), max_it, h, st);
        const bool intersect = T > 0.0f && h.block > 0;
        if (t) t[k] = T;
        if (normal_id) normal_id[k] = intersect ? (uint8_t)normal_id_of(h) : (uint8_t)VXPT_NORMAL_MISS;
        if (block_id) block_id[k] = intersect ? (uint8_t)h.block : 0;
        if (hit_voxel)
            for (int a = 0; a < 3; ++a) hit_voxel[3 * k + a] = intersect ? (int16_t)h.vox[a] : -1;
    }
    if (stats) { stats->rays += st.rays; stats->df_fetches += st.df; stats->vox_fetches += st.vox; }
    return VXPT_OK;
}

// PostProcessingVert.glsl:46-53 — v_PlayerShadowed
int vxo_player_shadowed(const VxoScene* sc, const float camera_pos[3], const float sun_dir[3]) {
    Scene S{*sc};
    const v3 sun = V(sun_dir[0], sun_dir[1], sun_dir[2]);
    const float L = length(sun);
    const v3 D = sun / L;
    Hit h; Stats st;
    return traverse_df(S, V(camera_pos[0], camera_pos[1], camera_pos[2]), D, 350, h, st) > 0.0f ? 1 : 0;
}

}  // extern "C"

namespace {
// EstimateAmbientSoundLevel.comp — HashRNG :145-153, Hash1 :160-164
inline float amb_hash1(uint32_t& seed) {
    seed ^= 2747636419u; seed *= 2654435769u;
    seed ^= seed >> 16;  seed *= 2654435769u;
    seed ^= seed >> 16;  seed *= 2654435769u;
    return (float)seed / 4294967295.0f;
}
// TraverseDistanceField :73-141: VoxelTraversalDF with a cap of 32 whose hit test reads the distance field (outside the volume
// GetDistance returns -1, which passes `D < 0.0001f`: a ray that steps out of the volume counts as a hit at that point)
float amb_traverse(const Scene& S, v3 origin, v3 direction, v3& normal, Stats& st) {
    const v3 initial_origin = origin;
    bool intersection = false;
    int min_idx = 0;
    const int sg[3] = {isign(direction.x), isign(direction.y), isign(direction.z)};
    st.rays++;
    for (int itr = 0; itr < 32; ++itr) {
        float fx = std::floor(origin.x), fy = std::floor(origin.y), fz = std::floor(origin.z);
        if (!S.in_volume_f(fx, fy, fz)) { intersection = false; break; }
        st.df++;
        float dist = (float)S.s.df[S.idx((int)fx, (int)fy, (int)fz)];
        int euclid = (int)std::floor(dist == 1.0f ? 1.0f : dist * 0.57735026918f);
        if (euclid == 0) break;
        if (euclid == 1) { dda_step(origin, direction, sg, min_idx); intersection = true; }
        else origin = origin + (float)(euclid - 1) * direction;
    }
    if (!intersection) return -1.0f;
    normal = V(0.0f);
    normal[min_idx] = (float)(-sg[min_idx]);
    float fx = std::floor(origin.x), fy = std::floor(origin.y), fz = std::floor(origin.z);
    float D = -1.0f;
    if (S.in_volume_f(fx, fy, fz)) { st.df++; D = (float)S.s.df[S.idx((int)fx, (int)fy, (int)fz)] / 255.0f; }
    return (D < 0.0001f || D == 0.0f) ? length(origin - initial_origin) : -1.0f;
}
// RaytraceAverageAmbience :210-241 with UniformHemisphere :183-190 and CosineHemisphereDirection :192-204
float amb_sample(const Scene& S, v3 player, uint32_t& seed, Stats& st) {
    const float PI = 3.141592653f;
    const v3 up = normalize(V(0.0f, 1.0f, 0.0f));
    v3 ro = player;
    const float ux = amb_hash1(seed), uy = amb_hash1(seed);  // GLSL evaluates arguments left to right
    v3 rd;
    {
        const float r = std::sqrt(1.0f - ux * ux);
        const float phi = 2.0f * PI * uy;
        const v3 B = normalize(cross(up, V(0.0f, 1.0f, 1.0f)));
        const v3 T = cross(B, up);
        rd = normalize(((r * sin_cr(phi)) * B + ux * up) + (r * cos_cr(phi)) * T);
    }
    for (int bounce = 0; bounce < 7; ++bounce) {
        v3 N = V(0.0f);
        const float T = amb_traverse(S, ro, rd, N, st);
        if (T < 0.0f) return 1.0f;
        ro = (ro + rd * T) + N * 0.05f;
        const float r1 = amb_hash1(seed), r2 = amb_hash1(seed);
        const float PI2 = 2.0f * PI;
        const v3 uu = normalize(cross(N, V(0.0f, 1.0f, 1.0f)));
        const v3 vv = cross(uu, N);
        const float ra = std::sqrt(r2);
        const float rx = ra * cos_cr(PI2 * r1), ry = ra * sin_cr(PI2 * r1), rz = std::sqrt(1.0f - r2);
        rd = normalize((rx * uu + ry * vv) + rz * N);
    }
    return 0.0f;
}
}  // namespace

extern "C" {

// EstimateAmbientSoundLevel.comp main() :247-270 over the 8 x 4 invocations Core/Pipeline.cpp:1921 dispatches
int vxo_ambient_sound(const VxoScene* sc, const float player_pos[3], int frame, uint32_t* aggregate, uint32_t* per_invocation, VxoStats* stats) {
    Scene S{*sc};
    Stats st;
    uint32_t sum = 0;
    for (int y = 0; y < 4; ++y)
        for (int x = 0; x < 8; ++x) {
            uint32_t seed = (uint32_t)((float)y * 32.0f + (float)x) + (uint32_t)(frame % 512) * 32u * 32u;  // InitRNG :155-158
            const int samples = (int)mixf(1.0f, 2.0f, frame % 2 == 0 ? 1.0f : 0.0f);
            float amount = 0.0f, weight = 0.0f;
            for (int s = 0; s < std::min(std::max(samples, 0), 3); ++s) {
                amount += amb_sample(S, V(player_pos[0], player_pos[1], player_pos[2]), seed, st);
                weight += 1.0f;
            }
            amount /= weight;
            const uint32_t mapped = (uint32_t)clampf(amount * 512.0f, 0.0f, 512.0f);
            sum += mapped;
            if (per_invocation) per_invocation[y * 8 + x] = mapped;
        }
    *aggregate = sum;
    if (stats) { stats->rays += st.rays; stats->df_fetches += st.df; stats->vox_fetches += st.vox; }
    return VXPT_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ G-buffer material pass
namespace {

struct f4 {
    float x, y, z, w;
};
inline f4 lerp4(f4 a, f4 b, float f) {
    return f4{a.x * (1.0f - f) + b.x * f, a.y * (1.0f - f) + b.y * f, a.z * (1.0f - f) + b.z * f, a.w * (1.0f - f) + b.w * f};
}

// One block array as the GL texture object holds it (Core/GLClasses/TextureArray.cpp:28-78): RGBA8 layers of 512^2 with a complete mip
// chain, GL_REPEAT, min filter GL_NEAREST_MIPMAP_LINEAR, mag filter GL_NEAREST (albedo, sRGB) or GL_LINEAR (normal, PBR).
struct MipArray {
    const uint8_t* texels;
    int layers;
    bool srgb, mag_linear;
    const float* srgb_lut;
    f4 texel(int layer, int level, int i, int j) const {
        size_t off = 0;
        for (int k = 0; k < level; ++k) off += (size_t)(512 >> k) * (size_t)(512 >> k);
        const int n = 512 >> level;
        const uint8_t* c = texels + ((size_t)layer * VXPT_MIP_CHAIN_TEXELS + off + (size_t)j * n + i) * 4;
        if (srgb) return f4{srgb_lut[c[0]], srgb_lut[c[1]], srgb_lut[c[2]], (float)c[3] / 255.0f};
        return f4{(float)c[0] / 255.0f, (float)c[1] / 255.0f, (float)c[2] / 255.0f, (float)c[3] / 255.0f};
    }
    f4 nearest(int layer, int level, float u, float v) const {
        const int n = 512 >> level;
        return texel(layer, level, ((int)std::floor(u * (float)n)) & (n - 1), ((int)std::floor(v * (float)n)) & (n - 1));
    }
    // textureGrad: OpenGL 4.3 section 8.14 with the isotropic scale factor (the pinned definition, include/vxpt.h)
    f4 grad(float layer_f, float u, float v, const float dpdx[2], const float dpdy[2]) const {
        int layer = (int)std::nearbyintf(layer_f);
        layer = std::min(std::max(layer, 0), layers - 1);
        const float ux = dpdx[0] * 512.0f, vx = dpdx[1] * 512.0f, uy = dpdy[0] * 512.0f, vy = dpdy[1] * 512.0f;
        const float rho = std::fmax(std::sqrt(ux * ux + vx * vx), std::sqrt(uy * uy + vy * vy));
        const float lambda = (float)std::log2((double)rho);
        const float c = mag_linear ? 0.5f : 0.0f;
        if (lambda <= c) {
            if (!mag_linear) return nearest(layer, 0, u, v);
            const float x = u * 512.0f - 0.5f, y = v * 512.0f - 0.5f;
            const float x0 = std::floor(x), y0 = std::floor(y);
            const float fx = x - x0, fy = y - y0;
            const int i0 = ((int)x0) & 511, i1 = ((int)x0 + 1) & 511, j0 = ((int)y0) & 511, j1 = ((int)y0 + 1) & 511;
            const f4 lo = lerp4(texel(layer, 0, i0, j0), texel(layer, 0, i1, j0), fx);
            const f4 hi = lerp4(texel(layer, 0, i0, j1), texel(layer, 0, i1, j1), fx);
            return lerp4(lo, hi, fy);
        }
        if (lambda >= 9.0f) return nearest(layer, 9, u, v);
        const int d1 = (int)std::floor(lambda);
        return lerp4(nearest(layer, d1, u, v), nearest(layer, d1 + 1, u, v), lambda - std::floor(lambda));
    }
};

// texture(u_LavaTextures[k], p): GL_LINEAR, GL_REPEAT on all axes (Core/AnimatedTexture.cpp:11-15); trilinear x, then y, then z
inline f4 lava_sample(const uint8_t* tex, float s, float t, float r) {
    const int N = VXPT_LAVA_SIZE, D = VXPT_LAVA_FRAMES;
    const float x = s * (float)N - 0.5f, y = t * (float)N - 0.5f, z = r * (float)D - 0.5f;
    const float x0 = std::floor(x), y0 = std::floor(y), z0 = std::floor(z);
    const float fx = x - x0, fy = y - y0, fz = z - z0;
    auto wr = [](int i, int n) { const int m = i % n; return m < 0 ? m + n : m; };
    const int i0 = wr((int)x0, N), i1 = wr((int)x0 + 1, N), j0 = wr((int)y0, N), j1 = wr((int)y0 + 1, N), k0 = wr((int)z0, D), k1 = wr((int)z0 + 1, D);
    auto tx = [&](int i, int j, int k) {
        const uint8_t* c = tex + (((size_t)k * N + j) * N + i) * 4;
        return f4{(float)c[0] / 255.0f, (float)c[1] / 255.0f, (float)c[2] / 255.0f, (float)c[3] / 255.0f};
    };
    const f4 a = lerp4(lerp4(tx(i0, j0, k0), tx(i1, j0, k0), fx), lerp4(tx(i0, j1, k0), tx(i1, j1, k0), fx), fy);
    const f4 b = lerp4(lerp4(tx(i0, j0, k1), tx(i1, j0, k1), fx), lerp4(tx(i0, j1, k1), tx(i1, j1, k1), fx), fy);
    return lerp4(a, b, fz);
}

// the operand a fragment hands to dFdx / dFdy in GetUVDerivative (GenerateGBuffer.glsl:443-461): UV = fract(P), P the hit point's
// coordinates in the plane of its own face (CalculateVectors :463-530).  `reached` is false for invocations that never get there.
struct QuadOperand {
    bool reached;
    float uv[2];
    float dist;
    v3 pos;
    int nid;
};
inline QuadOperand gbuffer_operand(const VxCamera& cam, const VxGBuffer& g, const VxMaterialParams& prm, int i, int j) {
    QuadOperand q{};
    if (i < 0 || j < 0 || i >= cam.width || j >= cam.height) return q;  // helper invocation outside the frame
    const size_t p = (size_t)j * cam.width + i;
    // ShouldUpdate :351-357: a fragment that discards never takes a derivative
    if (!prm.update_this_frame && !prm.pom && std::min<int>(g.block_id[p], 127) != prm.lava_block_id) return q;
    q.dist = 1.0f / g.inv_t[p];  // GetPositionAt :107-112 on u_NonLinearDepth
    if (q.dist < 0.0f) return q;  // :360-366
    q.nid = g.normal_id[p];
    if (q.nid > 5) return q;      // no face normal: CalculateVectors matches nothing (outputs unset in the shader); shaded like a miss
    float u, v;
    pixel_uv(cam, i, j, u, v);
    q.pos = ray_origin(cam) + normalize(ray_direction_at(cam, u, v)) * q.dist;
    const int axis = q.nid <= 1 ? 2 : (q.nid <= 3 ? 1 : 0);
    if (axis == 2) { q.uv[0] = fractf(q.pos.x); q.uv[1] = fractf(q.pos.y); }
    else if (axis == 1) { q.uv[0] = fractf(q.pos.x); q.uv[1] = fractf(q.pos.z); }
    else { q.uv[0] = fractf(q.pos.z); q.uv[1] = fractf(q.pos.y); }
    q.reached = true;
    return q;
}

}  // namespace

extern "C" {

// GenerateGBuffer.glsl main() :347-441 in the v1 parity profile (u_POM off, no lava), whole 2x2 quads of rows [row_begin, row_end).
// Derivatives: dFdx / dFdy = odd-minus-even member of the quad's row / column pair; a member that never reaches the derivative
// contributes the pixel's own operand.
int vxo_generate_gbuffer(const VxoScene* sc, const VxCamera* cam, const VxGBuffer* g, const VxMaterialParams* prm, const VxMaterialOut* out) {
    if (prm->lava_block_id >= 0 && (!sc->lava_albedo || !sc->lava_normal)) return VXPT_E_STATE;
    if (!prm->update_this_frame && !prm->pom && prm->lava_block_id < 0) return VXPT_OK;  // :351-357: every fragment discards
    float lut[256];
    for (int k = 0; k < 256; ++k) {
        const double cs = (double)k / 255.0;
        lut[k] = (float)(cs <= 0.04045 ? cs / 12.92 : std::pow((cs + 0.055) / 1.055, 2.4));
    }
    const MipArray albedo{sc->albedo_mips, sc->n_mip_layers, true, false, lut};
    const MipArray normals{sc->normal_mips, sc->n_mip_layers, false, true, lut};
    const MipArray pbr{sc->pbr_mips, sc->n_mip_layers, false, true, lut};
    const int W = cam->width;
    const int32_t* M = sc->materials;
#pragma omp parallel for schedule(dynamic, 4)
    for (int j = cam->row_begin; j < cam->row_end; ++j)
        for (int i = 0; i < W; ++i) {
            const size_t p = (size_t)j * W + i;
            const bool is_lava = prm->lava_block_id >= 0 && std::min<int>(g->block_id[p], 127) == prm->lava_block_id;
            if (!prm->update_this_frame && !prm->pom && !is_lava) continue;   // :351-357 discard: the planes keep their texels
            const QuadOperand me = gbuffer_operand(*cam, *g, *prm, i, j);
            if (!me.reached) {
                if (out->albedo) for (int k = 0; k < 3; ++k) out->albedo[3 * p + k] = 0.0f;
                if (out->normal) for (int k = 0; k < 3; ++k) out->normal[3 * p + k] = 1.0f;
                if (out->pbr) for (int k = 0; k < 4; ++k) out->pbr[4 * p + k] = 0.0f;
                if (out->texture_ao) out->texture_ao[p] = 0.0f;
                continue;
            }
            // the four operands of the quad's row pair and column pair
            const int i0 = i & ~1, j0 = j & ~1;
            QuadOperand row[2] = {gbuffer_operand(*cam, *g, *prm, i0, j), gbuffer_operand(*cam, *g, *prm, i0 + 1, j)};
            QuadOperand col[2] = {gbuffer_operand(*cam, *g, *prm, i, j0), gbuffer_operand(*cam, *g, *prm, i, j0 + 1)};
            for (int k = 0; k < 2; ++k) {
                if (!row[k].reached) row[k] = me;
                if (!col[k].reached) col[k] = me;
            }
            float dx[2], dy[2], dx2[2], dy2[2];
            for (int k = 0; k < 2; ++k) {
                dx[k] = row[1].uv[k] - row[0].uv[k];
                dy[k] = col[1].uv[k] - col[0].uv[k];
                dx2[k] = fractf(row[1].uv[k] + 0.25f) - fractf(row[0].uv[k] + 0.25f);
                dy2[k] = fractf(col[1].uv[k] + 0.25f) - fractf(col[0].uv[k] + 0.25f);
            }
            if ((dx[0] * dx[0] + dx[1] * dx[1]) + (dy[0] * dy[0] + dy[1] * dy[1]) > (dx2[0] * dx2[0] + dx2[1] * dx2[1]) + (dy2[0] * dy2[0] + dy2[1] * dy2[1])) {
                dx[0] = dx2[0]; dx[1] = dx2[1]; dy[0] = dy2[0]; dy[1] = dy2[1];
            }
            // GetBlockID :95-99, GetTextureIDs :532-549
            const int block = std::min(std::max((int)std::floor(((float)g->block_id[p] / 255.0f) * 255.0f), 0), 127);
            float data[4] = {(float)M[block], (float)M[128 + block], (float)M[256 + block], (float)M[384 + block]};
            if (block == prm->grass_props[0]) {
                const int base = (me.nid == 2) ? 1 : ((me.nid == 3) ? 7 : 4);  // top : bottom : the four sides
                for (int k = 0; k < 3; ++k) data[k] = (float)prm->grass_props[base + k];
            }
            v3 tangent, bitangent;
            float fu, fv;
            calc_vectors(me.pos, me.nid, tangent, bitangent, fu, fv);  // same tables as ReflectionTraceFrag's copy (:463-530 here)
            const v3 face = normal_from_id(me.nid, 1.0f);
            float lava_u = 0.0f, lava_v = 0.0f;
            const float lava_r = fractf(prm->time * 0.3f);
            if (is_lava) {  // :389-394: BasicTextureDistortion :127-137 on vec3(UV, fract(u_Time * 0.3f)); liquids skip the parallax march
                const float time = prm->time;
                float du = fu, dv = fv;
                du += sin_cr(time * 0.25f);
                dv += pow_cr(cos_cr(time * 0.15f), 2.0f);
                du += cos_cr(du * 10.0f + time) * 0.3f;
                dv += sin_cr(dv * 5.0f + du * 4.0f + time * 1.3f) * 0.4f;
                lava_u = mixf(du, fu, 0.91f);
                lava_v = mixf(dv, fv, 0.91f);
                fu = lava_u;
                fv = lava_v;
            } else if (prm->pom) {  // Parallax :343-353 -> ReliefParallax :153-199 (the code after its first return is dead)
                const v3 view = normalize(me.pos - ray_origin(*cam));
                const float depth_scale = 0.115f * prm->pom_height;
                float bayer_steps = 0.5f;
                if (prm->dither_pom) {
                    auto bayer2 = [](float ax, float ay) { ax = std::floor(ax); ay = std::floor(ay); return fractf(ax * 0.5f + ay * (ay * 0.75f)); };
                    const float cx = (float)i + 0.5f, cy = (float)j + 0.5f;   // gl_FragCoord.xy; bayer32 = bayer2 at five octaves
                    float cxs[5], cys[5];
                    cxs[0] = cx; cys[0] = cy;
                    for (int k = 1; k < 5; ++k) { cxs[k] = 0.5f * cxs[k - 1]; cys[k] = 0.5f * cys[k - 1]; }
                    float b = bayer2(cxs[4], cys[4]);
                    for (int k = 3; k >= 0; --k) b = b * 0.25f + bayer2(cxs[k], cys[k]);
                    const float fr = (float)prm->frame + 0.0f * 2.0f;
                    const float md = fr - 384.0f * std::floor(fr / 384.0f);
                    bayer_steps = fractf(fractf(md * (1.0f / 1.6180339f)) + b);
                }
                const v3 tv = normalize(V(dot(view, tangent), dot(view, bitangent), dot(view, -face)));
                float mdx = tv.x, mdy = tv.y;
                const float dv = std::fabs(face.x) > 0.01f ? tv.z : -(tv.z);
                mdx /= dv; mdy /= dv;
                mdx *= depth_scale; mdy *= depth_scale;
                const int steps = prm->high_quality_pom ? (int)mixf(64.0f, 128.0f, clampf(bayer_steps * 0.9f, 0.0f, 1.0f))
                                                        : (int)mixf(32.0f, 64.0f, clampf(bayer_steps * 0.85f, 0.0f, 1.0f));
                const float step_size = 1.0f / (float)steps;
                const float su = clampf(fu, 0.000001f, 1.0f), sv = clampf(fv, 0.000001f, 1.0f);
                float cur_depth = 1.0f, best = 1.0f;
                int layer = (int)std::nearbyintf((float)(int)data[2]);
                layer = std::min(std::max(layer, 0), pbr.layers - 1);
                for (int k = 0; k < steps; ++k) {
                    cur_depth -= step_size;
                    const float pu = su + mdx * cur_depth, pv = sv + mdy * cur_depth;
                    // texture(u_BlockPBR, ...) inside the loop: pinned to level 0, GL_LINEAR (include/vxpt.h)
                    const float x = pu * 512.0f - 0.5f, y = pv * 512.0f - 0.5f;
                    const float x0 = std::floor(x), y0 = std::floor(y), fx = x - x0, fy = y - y0;
                    const int i0 = ((int)x0) & 511, i1 = ((int)x0 + 1) & 511, j0 = ((int)y0) & 511, j1 = ((int)y0 + 1) & 511;
                    const f4 lo = lerp4(pbr.texel(layer, 0, i0, j0), pbr.texel(layer, 0, i1, j0), fx);
                    const f4 hi = lerp4(pbr.texel(layer, 0, i0, j1), pbr.texel(layer, 0, i1, j1), fx);
                    const float height = pow_cr(lerp4(lo, hi, fy).z, 1.5f * prm->pom_exp);   // MapHeight :147-149
                    if (cur_depth >= height) best = cur_depth;
                }
                cur_depth = best - step_size * 0.5f;
                fu = su + mdx * cur_depth;
                fv = sv + mdy * cur_depth;
            }
            const float U = 1.0f - fu, Vc = 1.0f - fv;  // :397
            const f4 nm = is_lava ? lava_sample(sc->lava_normal, lava_u, lava_v, lava_r) : normals.grad(data[1], U, Vc, dx, dy);
            const v3 n = V(nm.x * 2.0f - 1.0f, nm.y * 2.0f - 1.0f, nm.z * 2.0f - 1.0f);
            const v3 mapped = (tangent * n.x + bitangent * n.y) + face * n.z;  // mat3(T, B, N) * n
            const f4 pm = pbr.grad(data[2], U, Vc, dx, dy);
            float emissivity = 0.0f;
            if (data[3] > -0.5f) emissivity = tex_bilinear1(sc->emissive_lod0, (int)data[3], 512, U, Vc);
            float o_pbr[4] = {clampf(pm.x, 0.0f, 1.0f), clampf(pm.y, 0.0f, 1.0f), clampf(pm.z, 0.0f, 1.0f), clampf(emissivity, 0.0f, 1.0f)};
            const f4 al = is_lava ? lava_sample(sc->lava_albedo, lava_u, lava_v, lava_r) : albedo.grad(data[0], U, Vc, dx, dy);
            const float lb = 0.02f;
            if (!is_lava) o_pbr[3] *= (U > lb && U < 1.0f - lb && Vc > lb && Vc < 1.0f - lb) ? 1.0f : 0.0f;  // BloomLightLeakFix :432-438
            if (out->albedo) { out->albedo[3 * p] = al.x; out->albedo[3 * p + 1] = al.y; out->albedo[3 * p + 2] = al.z; }
            if (out->normal) { out->normal[3 * p] = mapped.x; out->normal[3 * p + 1] = mapped.y; out->normal[3 * p + 2] = mapped.z; }
            if (out->pbr) for (int k = 0; k < 4; ++k) out->pbr[4 * p + k] = o_pbr[k];
            if (out->texture_ao) out->texture_ao[p] = clampf(pm.w, 0.00000001f, 1.0f);
        }
    return VXPT_OK;
}

}  // extern "C"

extern "C" {

int vxo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void vxo_set_num_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#else
    (void)n;
#endif
}

}  // extern "C"
