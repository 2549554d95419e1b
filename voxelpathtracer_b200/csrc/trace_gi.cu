// trace_gi.cu — wavefront form of the diffuse-GI pass (DiffuseRayTraceFrag.glsl main() :822-935, CalculateDiffuse :535-664).
//
// The one-thread-per-pixel form (diffuse_kernel in trace.cu) runs 9 of 32 lanes per instruction on open terrain: most
// first-bounce rays escape to the sky while the few that hit geometry drag their warp through texture fetches, a shadow
// sub-ray, a second traversal and a second shadow sub-ray (ncu r01a).  Here the work of one sample is re-queued:
//
//   gi_gen_trace0  first hemisphere direction (blue-noise tables) and first traversal of every pixel, the rays of a CTA sorted by
//                  life expectancy before they are traced.  A miss is finished on the spot (sky radiance).  Hits are compacted
//                  into a queue with one warp ballot + one atomicAdd per warp (popc / prefix ranks give each lane its slot).
//   gi_continue    the rest of the sample for the queued hits (11 % of the first-bounce rays on plains), in CTA-wide stages with the
//                  rays re-distributed over the warps between the stages (shared-memory lists, counting sort):
//                    A  shade the first hit, draw the bounce direction           one thread per record (dense)
//                    B  trace the bounce rays and the sun-shadow sub-rays        warps claim groups of 32 rays from the sorted list
//                    C  shade the second hit (or the sky), queue its shadow ray  one thread per record
//                    D  trace the second shadow sub-rays                         groups of 32 rays
//                    E  finish the sample (SH projection, stores)                one thread per record
//                  r01h's one-thread-per-record kernel ran the chain texture -> shadow ray -> bounce -> texture -> shadow ray in every
//                  lane with 11.4 of 32 lanes active.
//   gi_finalize    only when some pixel takes more than one sample: per-pixel averages and clamps (:915-934).
//
// Samples of one pixel are processed one after another (launch s+1 after launch s), so the blue-noise dimension counter
// and the accumulation order are exactly those of the shader's loop; the planes are bit-identical to diffuse_kernel's:
// which thread traces a ray, and when, does not change what is computed for its pixel.
#include <algorithm>
#include <cstdlib>

#include "gi_device.cuh"

namespace vxpt {

// Development aid (build.py -DVXPT_GI_TRACE --out=...; tools/gi_timeline.py): thread 0 of every CTA logs %globaltimer at its phase
// boundaries.  Not compiled into the product library.
#ifdef VXPT_GI_TRACE
constexpr int GI_TRACE_SLOTS = 16;
__device__ unsigned long long g_gi_trace[2][4096][GI_TRACE_SLOTS];
__device__ __forceinline__ void gi_trace(int kernel, int cta, int slot) {
    if (threadIdx.x == 0 && cta < 4096 && slot < GI_TRACE_SLOTS) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_gi_trace[kernel][cta][slot] = t;
    }
}
#define GI_TRACE(kernel, cta, slot) gi_trace(kernel, cta, slot)
#else
#define GI_TRACE(kernel, cta, slot)
#endif

struct HitRec {   // 48 B
    float4 a;     // ro.xyz, T
    float4 b;     // rd.xyz, pixel index (bits)
    float4 c;     // hit code (bits): min_idx | (sgn+1) << 2 | block << 8 ; bl_sample (bits) ; unused
};
struct PixState { // 48 B, only when spp > 1 somewhere
    float4 tot;   // SH sums
    float4 rad;   // radiance sum xyz, AO sum
    float4 misc;  // CoCg sums, sky-hit sum, bl_sample (bits)
};

// SPP of a pixel — DiffuseRayTraceFrag.glsl:874-892
__device__ __forceinline__ int pixel_spp(const DiffuseDev& P, int i, int j) {
    int spp = min(max(P.spp, 1), 32);
    if (P.checkerboard) {
        const bool checker = ((int)(((float)i + 0.5f) + ((float)j + 0.5f))) % 2 == P.frame % 2;
        spp = (int)mixf((float)P.spp, (float)P.checker_spp, checker ? 1.0f : 0.0f);
    }
    spp = min(max(spp, 1), 32);
    if (P.moon_stronger) spp *= 2;
    return spp;
}

// :915-934 — averages, luminance, clamps, stores
__device__ __forceinline__ void write_final(const DiffuseOutDev& out, size_t px, V3 radiance, float acc_ao, float t0, float t1, float t2, float t3,
                                            float c0, float c1, float skyhits, int spp) {
    const float fs = (float)spp;
    acc_ao /= fs;
    t0 /= fs; t1 /= fs; t2 /= fs; t3 /= fs;
    c0 /= fs; c1 /= fs;
    radiance = radiance / fs;
    skyhits /= fs;
    const float lum = dot3(radiance, mk3(0.299f, 0.587f, 0.114f));
    float util = fmaxf(lum, 0.01f);
    util = clampf(util, 0.001f, 64.0f);
    if (out.sh) store_f4(out.sh, px, clampf(t0, -100.0f, 100.0f), clampf(t1, -100.0f, 100.0f), clampf(t2, -100.0f, 100.0f), clampf(t3, -100.0f, 100.0f), out.fmt);
    if (out.cocg) store_f2(out.cocg, px, clampf(c0, -100.0f, 100.0f), clampf(c1, -100.0f, 100.0f), out.fmt);
    if (out.luma) store_f1(out.luma, px, util, out.fmt);
    if (out.ao_sky) store_unorm2(out.ao_sky, px, clampf(acc_ao, 0.0f, 1.0f), clampf(skyhits, 0.0f, 1.0f), out.fmt);
}

// end of one sample (:898-913): clamp, SH projection, accumulate (or, when every pixel takes one sample, finish)
template <bool SPP1>
__device__ __forceinline__ void finish_sample(const DiffuseOutDev& out, PixState* state, size_t px, V3 rad, float ao, V3 odir, bool skyhit, int bl_sample) {
    rad = mk3(clampf(rad.x, 0.0f, 8.0f), clampf(rad.y, 0.0f, 8.0f), clampf(rad.z, 0.0f, 8.0f));
    float sh[6];
    irradiance_to_sh(rad, odir, sh);
    const float ss = skyhit ? 1.0f : 0.0f;
    if (SPP1) {
        // sums start at 0: 0 + x == x, and x / 1.0f == x, so the single sample goes straight to the final stage
        write_final(out, px, mk3(0.f, 0.f, 0.f) + rad, 0.0f + ao, 0.0f + sh[0], 0.0f + sh[1], 0.0f + sh[2], 0.0f + sh[3], 0.0f + sh[4], 0.0f + sh[5],
                    0.0f + ss, 1);
    } else {
        PixState st = state[px];
        st.tot.x += sh[0]; st.tot.y += sh[1]; st.tot.z += sh[2]; st.tot.w += sh[3];
        st.rad.x = st.rad.x + rad.x; st.rad.y = st.rad.y + rad.y; st.rad.z = st.rad.z + rad.z; st.rad.w += ao;
        st.misc.x += sh[4]; st.misc.y += sh[5]; st.misc.z += ss;
        st.misc.w = __int_as_float(bl_sample);
        state[px] = st;
    }
}

// sky term of a ray that leaves the scene (:625-634)
__device__ __forceinline__ V3 sky_term(const SceneDev& S, const DiffuseDev& P, V3 rd) {
    float x = mixf(1.0f, 1.05f, P.sun_visibility);
    x = clampf(x * 1.0f * P.gi_sky_strength, 0.0f, 5.0f);
    V3 sd = rd;
    sd.y = clampf(sd.y, 0.125f, 1.5f);  // GetSkyColorAt :984-988
    return sky_sample(S, sd) * x;
}

// Shading of one hit (:569-622) up to, but without, the sun-shadow sub-ray of GetShadowAt (:1202-1222), which the caller traces (or has
// traced) elsewhere.  The shader's  SunBRDF = (((albedo * hammon) * (LIGHT_COLOR * 3.5)) * (1 - shadow)) * PI  is split at the shadow
// factor: `pre` is the product in front of it, sun_brdf() below completes it once the sub-ray's verdict is known — same operations, same
// order.  shadow_state: 0 / 1 = the verdict is known without a ray (N.L < 0.001 -> lit, moon stronger -> shadowed), 2 = trace the
// sub-ray from `shadow_origin`.  emis = the emissive colour (not yet multiplied by the throughput).
struct HitShade {
    V3 ipos, hn, albedo, pre, emis, shadow_origin;
    float rough;
    int shadow_state, face;
};
__device__ __forceinline__ HitShade shade_hit(const SceneDev& S, const DiffuseDev& P, V3 ro, V3 rd, float T, int min_idx, int sgn, int block) {
    HitShade h;
    const int tex_ref = min(max(block, 0), 127);
    h.ipos = ro + (rd * T);
    const float s = (float)(-sgn);
    h.hn = mk3(min_idx == 0 ? s : 0.0f, min_idx == 1 ? s : 0.0f, min_idx == 2 ? s : 0.0f);
    h.face = face_of(min_idx, -sgn);
    float tu, tv;
    calc_uv(h.ipos, min_idx, tu, tv);
    const int albedo_layer = S.materials[tex_ref], emissive_layer = S.materials[384 + tex_ref];
    h.albedo = tex_nearest(S.albedo_lod3, albedo_layer, 64, tu, tv);
    const V3 pbr = tex_nearest(S.pbr_lod2, albedo_layer, 128, tu, tv);  // sic: albedo layer (:578)
    h.rough = pbr.x;
    float emis = 0.0f;
    if ((float)emissive_layer >= 0.0f) {
        const float se = tex_bilinear1(S.emissive, emissive_layer, 512, tu, tv);
        emis = se * P.emissivity_mult * P.light_intensity;
    }
    const float ndl = fmaxf(dot3(h.hn, P.stronger_dir), 0.0f);
    h.shadow_state = P.moon_stronger ? 1 : (ndl < 0.001f ? 0 : 2);
    h.shadow_origin = h.ipos + h.hn * 0.045f;
    h.emis = (emis * mixf(1.0f, 1.0f, P.sun_visibility)) * h.albedo;
    h.pre = (h.albedo * diffuse_hammon(h.hn, -rd, P.stronger_dir, h.rough)) * (P.light_color * 3.5f);
    return h;
}
__device__ __forceinline__ V3 sun_brdf(V3 pre, float shadow_at) { return (pre * (1.0f - shadow_at)) * PI_F; }

// gi_gen_trace0 — first hemisphere direction and first traversal of every pixel of a slab, with the rays of a CTA SORTED before they are
// traced.  How long a hemisphere ray lives is decided mostly by where it points: steep rays reach large step values within a few
// iterations and leave the volume (or hit the ground at once), grazing rays creep through E = 1..3 cells until the cap.  Neighbouring
// pixels draw unrelated directions, so an unsorted warp waits for its longest ray with half of its lanes idle (ncu r01g: 15.8 of 32
// active).  Here a CTA of 256 threads generates the rays of RPT 32x8-pixel tiles (RPT per thread), sorts them by |rd.y| (counting sort on
// an 8-bit key in shared memory) and its warps then claim groups of 32 rays from the sorted list, longest-lived first, until the list is
// empty: a warp holds rays of similar life expectancy, and short groups fill the time the warps that drew long groups are still busy
// (with one group per warp the CTA would keep its registers until its slowest warp ended).  RPT = 0 selects the plain form (one pixel per
// thread, no exchange).
// resident CTAs per SM the register allocation aims for (build.py -D... to experiment).  r02w: 5 / 4 = 48 / 64 registers, no spills
// (60 / 79 uncapped): the whole 1080p pass 0.2788 -> 0.2720 ms.  The phase timeline (tools/gi_timeline.py, profiles/r02v_gi_timeline.txt)
// shows why more resident CTAs buy so little in gi_continue: a 256-record chunk takes 40 us — 19 us in stage B and 11 us in stage D, i.e.
// the dependent iterations of its longest rays — however many chunks run beside it.
#ifndef VXPT_GI_GEN_MINB
#define VXPT_GI_GEN_MINB 5   // r03v: 6 (40 registers, 16 bytes of stack) takes the whole pass from 0.246 to 0.290 ms
#endif
#ifndef VXPT_GI_CONT_MINB
#define VXPT_GI_CONT_MINB 5   // r03t, after both sub-rays moved to stage D: 4 / 5 / 6 CTAs per SM (64 / 48 / 40 registers) = 0.258 / 0.247 / 0.264 ms
#endif
template <int LAYOUT, bool SPP1, int RPT>
__global__ void __launch_bounds__(256, VXPT_GI_GEN_MINB) gi_gen_trace0(const SceneDev S, const __grid_constant__ CameraDev cam, const DiffuseDev P, const GBufferDev g,
                                                     const DiffuseOutDev out, PixState* __restrict__ state, HitRec* __restrict__ queue,
                                                     unsigned* __restrict__ queue_count, const int sample) {
    constexpr int NR = RPT > 0 ? RPT : 1;       // pixels per thread
    constexpr bool SORT = RPT > 0;
    __shared__ float4 s_a[SORT ? 256 * NR : 1];  // ro.xyz, pixel index (bits)
    __shared__ float4 s_b[SORT ? 256 * NR : 1];  // rd.xyz, bl_sample (bits)
    __shared__ unsigned s_hist[SORT ? 258 : 1];  // 256 bins, [256] = ray count, [257] = next group
    __shared__ unsigned short s_order[SORT ? 256 * NR : 1];  // sorted position -> slot of s_a / s_b (rays stay where they were written)
    const unsigned tid = threadIdx.x, lane = tid & 31;
    Counters cnt = {0u, 0u, 0u};
    GI_TRACE(0, blockIdx.y * gridDim.x + blockIdx.x, 0);
    if (SORT) {
        s_hist[tid] = 0u;
        if (tid < 2) s_hist[256 + tid] = 0u;
        __syncthreads();
    }
    bool has_ray0 = false;                   // !SORT: the one ray of this thread stays in registers
    V3 ro0 = mk3(0.f, 0.f, 0.f), rd0 = mk3(0.f, 1.f, 0.f);
    unsigned px0 = 0;
    int bl0 = 0;
    unsigned kr[NR];                         // SORT: key << 16 | rank within the key's bin, ~0u = no ray
#pragma unroll
    for (int r = 0; r < NR; ++r) {
        kr[r] = ~0u;
        // pixel of this thread in the CTA's r-th tile: a warp covers 8x4 pixels, a tile 32x8 (thread_pixel's mapping)
        const int warp = tid >> 5;
        const int i = blockIdx.x * 32 + (warp & 3) * 8 + (lane & 7);
        const int prow = cam.row_begin + (blockIdx.y * NR + r) * 8 + (warp >> 2) * 4 + (lane >> 3);
        if (i >= cam.width || prow >= cam.row_end) continue;
        const int j = image_row(cam, prow);
        const size_t px = (size_t)prow * cam.width + i;
        float u = ((float)i + 0.5f) / (float)cam.width;
        float v = ((float)j + 0.5f) / (float)cam.height;
        const float u0 = u, v0 = v;
        if (P.supersample) {
            u += (P.hx * 0.75f) / (float)cam.width;
            v += (P.hy * 0.75f) / (float)cam.height;
        }
        const float dist = load_f1(g.t, px, g.fmt);
        const int nid = g.normal_id[px];
        const V3 normal = normal_from_id(nid, 0.5f);
        if (dist < 0.0f) {
            if (sample == 0) {  // sky pixel (:866-872)
                float sh[6];
                const V3 vdir = normalize3(ray_direction_at(cam, u0, v0));
                irradiance_to_sh(sky_sample(S, vdir) * 2.66f, normal, sh);
                if (out.sh) store_f4(out.sh, px, sh[0], sh[1], sh[2], sh[3], out.fmt);
                if (out.cocg) store_f2(out.cocg, px, sh[4], sh[5], out.fmt);
                if (out.luma) store_f1(out.luma, px, 0.0f, out.fmt);
                if (out.ao_sky) store_unorm2(out.ao_sky, px, 1.0f, 0.0f, out.fmt);
            }
        } else if (sample < pixel_spp(P, i, j)) {
            int bls = 0;
            if (!SPP1) {
                if (sample == 0) {
                    PixState z;
                    z.tot = z.rad = z.misc = make_float4(0.f, 0.f, 0.f, 0.f);
                    state[px] = z;
                } else {
                    bls = __float_as_int(state[px].misc.w);
                }
            }
            const V3 pos = ray_origin(cam) + normalize3(ray_direction_at(cam, u, v)) * dist;
            const V3 o = pos + normal * 0.06f;
            const V3 d = cos_hemisphere(S, i, j, P.frame % 128, bls, normal, nid);
            if (SORT) {
                const unsigned key = life_key(d);
                kr[r] = (key << 16) | atomicAdd(&s_hist[key], 1u);
                s_a[r * 256 + tid] = make_float4(o.x, o.y, o.z, __int_as_float((int)px));
                s_b[r * 256 + tid] = make_float4(d.x, d.y, d.z, __int_as_float(bls));
            } else {
                has_ray0 = true; ro0 = o; rd0 = d; px0 = (unsigned)px; bl0 = bls;
            }
        }
    }
    if (!SORT) {
        bool push = false;
        HitRec rec;
        if (has_ray0) {
            TraceHit h;
            const float T = traverse_df<LAYOUT>(S, ro0, rd0, P.trace_length, h, cnt);
            if (T > 0.0f && h.block > 0) {
                push = true;
                rec.a = make_float4(ro0.x, ro0.y, ro0.z, T);
                rec.b = make_float4(rd0.x, rd0.y, rd0.z, __int_as_float((int)px0));
                rec.c = make_float4(__int_as_float(h.min_idx | ((h.sgn + 1) << 2) | (h.block << 8)), __int_as_float(bl0), 0.f, 0.f);
            } else {
                const V3 contrib = mk3(0.f, 0.f, 0.f) + sky_term(S, P, rd0) * mk3(1.f, 1.f, 1.f);
                finish_sample<SPP1>(out, state, (size_t)px0, contrib, 1.0f, rd0, true, bl0);
            }
        }
        const unsigned slot = warp_push(queue_count, push);
        if (push) queue[slot] = rec;
        flush_counters(S, cnt);
        return;
    }
    __syncthreads();
    GI_TRACE(0, blockIdx.y * gridDim.x + blockIdx.x, 1);
    prefix_256(s_hist, tid);
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NR; ++r)
        if (kr[r] != ~0u) s_order[s_hist[kr[r] >> 16] + (kr[r] & 0xFFFFu)] = (unsigned short)(r * 256 + tid);
    __syncthreads();
    GI_TRACE(0, blockIdx.y * gridDim.x + blockIdx.x, 2);
    const unsigned n_rays = s_hist[256], n_groups = (n_rays + 31u) / 32u;
    while (true) {
        unsigned grp = 0;
        if (lane == 0) grp = atomicAdd(&s_hist[257], 1u);
        grp = __shfl_sync(0xffffffffu, grp, 0);
        if (grp >= n_groups) break;
        const unsigned idx = grp * 32u + lane;
        bool push = false;
        HitRec rec;
        if (idx < n_rays) {
            const unsigned slot_ab = s_order[idx];
            const float4 a = s_a[slot_ab], b = s_b[slot_ab];
            const V3 o = mk3(a.x, a.y, a.z), d = mk3(b.x, b.y, b.z);
            const size_t px = (size_t)(unsigned)__float_as_int(a.w);
            const int bls = __float_as_int(b.w);
            TraceHit h;
            const float T = traverse_df<LAYOUT>(S, o, d, P.trace_length, h, cnt);
            if (T > 0.0f && h.block > 0) {
                push = true;
                rec.a = make_float4(o.x, o.y, o.z, T);
                rec.b = make_float4(d.x, d.y, d.z, a.w);
                rec.c = make_float4(__int_as_float(h.min_idx | ((h.sgn + 1) << 2) | (h.block << 8)), b.w, 0.f, 0.f);
            } else {
                const V3 contrib = mk3(0.f, 0.f, 0.f) + sky_term(S, P, d) * mk3(1.f, 1.f, 1.f);
                finish_sample<SPP1>(out, state, px, contrib, 1.0f, d, true, bls);
            }
        }
        const unsigned slot = warp_push(queue_count, push);
        if (push) queue[slot] = rec;
    }
    GI_TRACE(0, blockIdx.y * gridDim.x + blockIdx.x, 3);   // thread 0's warp has no group left
    flush_counters(S, cnt);
}

// gi_continue — everything of a sample after its first hit, for the queued hits, in CTA-wide stages (header comment).  The hit count stays on the
// device (no host round trip): every CTA reads it and takes an even share of the queue, in equal chunks of at most 256 records.  Thread t owns record t of the chunk in the dense stages; between them its state waits in shared memory (so the
// traversal stages run with the registers of a traversal, not of the whole chain).
//
// The shader's accumulation, in its order (:583-605 per bounce, :625-634 for the sky):
//     contrib = 0;  thr = 1
//     hit 0:  contrib = contrib + thr * SunBRDF0;  contrib = contrib + emis0 * thr;  thr = thr * ((albedo0 * atten0) / pdf0)
//     hit 1:  contrib = contrib + thr * SunBRDF1;  contrib = contrib + emis1 * thr          | miss: contrib = contrib + sky * thr
// SunBRDF needs the verdict of the hit's sun-shadow sub-ray; everything else of a hit is computed in the dense stage that shades it.
constexpr int GC_THREADS = 256;
// Both sun-shadow sub-rays of a record are traced in stage D (r03q: the first hit's used to run in stage B beside the bounce ray).  Stage B
// then ends after the bounce rays' cap (trace_length iterations) instead of the sub-rays' 128, the chunk's dependent chain is
// trace_length + 128 iterations instead of 128 + 128, and the sample's sum is formed in stage E, in the shader's order: the 1080p GI pass
// went from 0.2735 to 0.2590 ms.  The struct is kept under 36.8 KB so that six CTAs fit an SM's shared memory.
struct GcShared {
    float4 ray_o[2 * GC_THREADS];   // [t] bounce ray origin, w: (stage B) min_idx | (sgn+1) << 2 | block << 8 of its hit; after stage C the second
                                    // shadow sub-ray's origin; [256 + t] first shadow sub-ray origin
    float4 ray_d[GC_THREADS];       // bounce ray direction, w: (stage B) its T
    float st[21][GC_THREADS];       // per-record state between the dense stages
    unsigned char res_sh0[GC_THREADS], res_sh1[GC_THREADS];  // shadow sub-rays: 1 = occluded
    unsigned short order[2 * GC_THREADS];
    unsigned hist[260];             // 256 bins, [256] ray count, [257] next group, [258] unused, [259] stage-D ray count
};
static_assert(sizeof(GcShared) + 1024 <= (227 * 1024) / 6, "six gi_continue CTAs fit the shared memory of an SM");
// state rows (ST_A*, ST_P*: the first hit's sun term and emission; ST_C*, ST_Q*: the second hit's, or in ST_Q* the sky term of a bounce ray that missed)
enum { ST_PX = 0, ST_BLS, ST_OD0, ST_OD1, ST_OD2, ST_AO, ST_C0, ST_C1, ST_C2, ST_T0, ST_T1, ST_T2, ST_A0, ST_A1, ST_A2, ST_P0, ST_P1, ST_P2, ST_Q0, ST_Q1, ST_Q2 };

template <int LAYOUT, bool SPP1>
__global__ void __launch_bounds__(GC_THREADS, VXPT_GI_CONT_MINB) gi_continue(const SceneDev S, const __grid_constant__ CameraDev cam, const DiffuseDev P, const DiffuseOutDev out,
                                                          PixState* __restrict__ state, const HitRec* __restrict__ queue,
                                                          unsigned* __restrict__ queue_count, const unsigned min_share) {
    extern __shared__ __align__(16) unsigned char gc_smem[];
    GcShared& sm = *reinterpret_cast<GcShared*>(gc_smem);
    const unsigned count = queue_count[0];
    Counters cnt = {0u, 0u, 0u};
    const unsigned tid = threadIdx.x, lane = tid & 31;
    // The queue is complete when this kernel starts, so the records are dealt out evenly: every CTA takes the same share, in chunks of equal
    // size (at most 256 records).  A chunk's time is set by its longest rays far more than by its record count, so what costs is a last wave
    // of chunks that fills a fraction of the SMs (r03r: 714 chunks of 256 on 592 resident CTAs = one wave and a fifth, 78 us where 52 would
    // do); with even shares every resident CTA runs the same number of chunks, and a small slab (one of 8 GPUs' rows) spreads over all SMs.
    // A queue too short to give every resident CTA a full chunk is dealt to about 2.4 CTAs per SM (r04f, rank 0's share of a 2- / 4- / 8-way
    // sharded 1080p frame with other frames' kernels running beside it: 256 / 128 / 64 records per CTA were the fastest, 355 CTAs each time —
    // fewer leave SMs idle, more hold registers the co-running kernels need); min_share > 0 overrides (VXPT_GI_MIN_SHARE).
    const unsigned spread = min((unsigned)GC_THREADS, max(32u, (count + 354u) / 355u));
    const unsigned share = max(min_share ? min_share : spread, (count + gridDim.x - 1u) / gridDim.x);
    const unsigned lo = min(count, blockIdx.x * share), hi = min(count, lo + share);
    const unsigned n_chunks = (hi - lo + GC_THREADS - 1u) / GC_THREADS, csize = n_chunks ? (hi - lo + n_chunks - 1u) / n_chunks : 0u;
    for (unsigned gc_it = 0; gc_it < n_chunks; ++gc_it) {
        __syncthreads();  // the previous chunk's stage E has read everything it needs
        GI_TRACE(1, blockIdx.x, 7 * gc_it);
        sm.hist[tid] = 0u;
        if (tid < 4) sm.hist[256 + tid] = 0u;
        __syncthreads();
        const unsigned base = lo + gc_it * csize;
        const bool valid = tid < csize && base + tid < hi;
        // ---- stage A: shade the first hit, draw the bounce ray -------------------------------------------------------------------
        unsigned kr_b = ~0u;
        bool need0 = false;  // the first hit has a sun-shadow sub-ray (traced in stage D)
        int pi = 0, pj = 0;
        if (valid) {
            const HitRec rec = queue[base + tid];
            const V3 ro = mk3(rec.a.x, rec.a.y, rec.a.z), rd = mk3(rec.b.x, rec.b.y, rec.b.z);
            const float T0 = rec.a.w;
            const unsigned px = (unsigned)__float_as_int(rec.b.w);
            pi = (int)(px % (unsigned)cam.width);
            pj = image_row(cam, (int)(px / (unsigned)cam.width));
            const int code = __float_as_int(rec.c.x);
            int bl_sample = __float_as_int(rec.c.y);
            const HitShade h = shade_hit(S, P, ro, rd, T0, code & 3, ((code >> 2) & 3) - 1, (code >> 8) & 255);
            // next direction and throughput (:607-621)
            const V3 new_dir = cos_hemisphere(S, pi, pj, P.frame % 128, bl_sample, h.hn, h.face);
            const float cos_theta = clampf(dot3(h.hn, new_dir), 0.0f, 1.0f);
            const float pdf = fmaxf(cos_theta / PI_F, 0.00001f);
            const V3 atten = mk3(1.f, 1.f, 1.f) * diffuse_hammon(h.hn, -rd, new_dir, h.rough);
            const V3 thr0 = mk3(1.f, 1.f, 1.f);
            const V3 thr1 = thr0 * ((h.albedo * atten) / pdf);
            const V3 ro1 = h.ipos + h.hn * 0.06f;
            float ao = 1.0f;
            if (T0 < 2.0f && T0 > 0.0f) ao = fmaxf(T0 / 2.0f, 0.0f);  // :637-650
            const V3 e0 = h.emis * thr0;
            sm.st[ST_PX][tid] = __int_as_float((int)px);
            sm.st[ST_BLS][tid] = __int_as_float(bl_sample);
            sm.st[ST_OD0][tid] = rd.x; sm.st[ST_OD1][tid] = rd.y; sm.st[ST_OD2][tid] = rd.z;
            sm.st[ST_AO][tid] = ao;
            sm.st[ST_T0][tid] = thr1.x; sm.st[ST_T1][tid] = thr1.y; sm.st[ST_T2][tid] = thr1.z;
            sm.st[ST_A0][tid] = h.pre.x; sm.st[ST_A1][tid] = h.pre.y; sm.st[ST_A2][tid] = h.pre.z;
            sm.st[ST_P0][tid] = e0.x; sm.st[ST_P1][tid] = e0.y; sm.st[ST_P2][tid] = e0.z;
            sm.ray_o[tid] = make_float4(ro1.x, ro1.y, ro1.z, 0.f);
            sm.ray_d[tid] = make_float4(new_dir.x, new_dir.y, new_dir.z, 0.f);
            const unsigned key = life_key(new_dir);
            kr_b = (key << 16) | atomicAdd(&sm.hist[key], 1u);
            sm.res_sh0[tid] = (unsigned char)(h.shadow_state & 1);
            if (h.shadow_state == 2) {
                sm.ray_o[GC_THREADS + tid] = make_float4(h.shadow_origin.x, h.shadow_origin.y, h.shadow_origin.z, 0.f);
                need0 = true;
            }
        }
        __syncthreads();
        GI_TRACE(1, blockIdx.x, 7 * gc_it + 1);   // stage A done
        prefix_256(sm.hist, tid);
        __syncthreads();
        if (kr_b != ~0u) sm.order[sm.hist[kr_b >> 16] + (kr_b & 0xFFFFu)] = (unsigned short)tid;
        __syncthreads();
        GI_TRACE(1, blockIdx.x, 7 * gc_it + 2);   // sorted
        // ---- stage B: bounce rays (cap trace_length), longest-lived first ---------------------------------------------------------------
        // (Tried and rejected, r02x: a warp claiming two groups at a time and tracing them interleaved, two rays per lane with both
        // step-field loads issued before either is used — bit-identical, but the pass went from 0.272 to 0.291 ms at 64 registers and
        // 0.287 ms at 78: every iteration then runs both rays' skip and DDA halves under predication.)
        {
            const unsigned n_rays = sm.hist[256], n_groups = (n_rays + 31u) / 32u;
            while (true) {
                unsigned grp = 0;
                if (lane == 0) grp = atomicAdd(&sm.hist[257], 1u);
                grp = __shfl_sync(0xffffffffu, grp, 0);
                if (grp >= n_groups) break;
                const unsigned idx = grp * 32u + lane;
                if (idx < n_rays) {
                    const unsigned slot = sm.order[idx];
                    const float4 o4 = sm.ray_o[slot], d4 = sm.ray_d[slot];
                    TraceHit h;
                    const float T = traverse_df<LAYOUT>(S, mk3(o4.x, o4.y, o4.z), mk3(d4.x, d4.y, d4.z), P.trace_length, h, cnt);
                    sm.ray_d[slot].w = T;
                    sm.ray_o[slot].w = __int_as_float(h.min_idx | ((h.sgn + 1) << 2) | (h.block << 8));
                }
            }
        }
        __syncthreads();
        GI_TRACE(1, blockIdx.x, 7 * gc_it + 3);   // stage B done
        if (tid == 0) sm.hist[257] = 0u;  // stage D claims groups from the same cursor (read again only after the next barrier)
        // ---- stage C: the bounce ray's hit (or the sky) ---------------------------------------------------------------------------
        bool done = true, skyhit = false, hit1 = false;
        if (valid) {
            const V3 thr1 = mk3(sm.st[ST_T0][tid], sm.st[ST_T1][tid], sm.st[ST_T2][tid]);
            const float4 o4 = sm.ray_o[tid], d4 = sm.ray_d[tid];
            const float T1 = d4.w;
            const int code = __float_as_int(o4.w);
            const V3 ro1 = mk3(o4.x, o4.y, o4.z), rd1 = mk3(d4.x, d4.y, d4.z);
            if (T1 > 0.0f && ((code >> 8) & 255) > 0) {
                const HitShade h = shade_hit(S, P, ro1, rd1, T1, code & 3, ((code >> 2) & 3) - 1, (code >> 8) & 255);
                sm.st[ST_BLS][tid] = __int_as_float(__float_as_int(sm.st[ST_BLS][tid]) + 2);  // the shader still draws the direction of a third segment it never traces (:607)
                const V3 e1 = h.emis * thr1;
                hit1 = true;
                sm.st[ST_C0][tid] = h.pre.x; sm.st[ST_C1][tid] = h.pre.y; sm.st[ST_C2][tid] = h.pre.z;
                sm.st[ST_Q0][tid] = e1.x; sm.st[ST_Q1][tid] = e1.y; sm.st[ST_Q2][tid] = e1.z;
                if (h.shadow_state == 2) {
                    done = false;
                    sm.ray_o[tid] = make_float4(h.shadow_origin.x, h.shadow_origin.y, h.shadow_origin.z, 0.f);  // the bounce origin has been read
                } else {
                    sm.res_sh1[tid] = h.shadow_state ? 1 : 0;
                }
            } else {
                const V3 sky = sky_term(S, P, rd1) * thr1;
                sm.st[ST_Q0][tid] = sky.x; sm.st[ST_Q1][tid] = sky.y; sm.st[ST_Q2][tid] = sky.z;
                skyhit = true;
            }
        }
        {   // stage D's rays: compacted list of slots (all share the sun direction: no sort); slot t = record t's second sub-ray
            // (origin ray_o[t]), slot 256 + t = its first (origin ray_o[256 + t])
            const unsigned m1 = __ballot_sync(0xffffffffu, !done), m0 = __ballot_sync(0xffffffffu, need0);
            const unsigned n1 = (unsigned)__popc(m1), n0 = (unsigned)__popc(m0);
            unsigned wbase = 0;
            if (lane == 0 && (n0 + n1)) wbase = atomicAdd(&sm.hist[259], n0 + n1);
            wbase = __shfl_sync(0xffffffffu, wbase, 0);
            const unsigned below = (1u << lane) - 1u;
            if (need0) sm.order[wbase + __popc(m0 & below)] = (unsigned short)(GC_THREADS + tid);
            if (!done) sm.order[wbase + n0 + __popc(m1 & below)] = (unsigned short)tid;
        }
        __syncthreads();
        GI_TRACE(1, blockIdx.x, 7 * gc_it + 4);   // stage C done
        // ---- stage D: shadow sub-rays (cap 128) --------------------------------------------------------------------------------------
        {
            const unsigned n_rays = sm.hist[259], n_groups = (n_rays + 31u) / 32u;
            while (true) {
                unsigned grp = 0;
                if (lane == 0) grp = atomicAdd(&sm.hist[257], 1u);
                grp = __shfl_sync(0xffffffffu, grp, 0);
                if (grp >= n_groups) break;
                const unsigned idx = grp * 32u + lane;
                if (idx < n_rays) {
                    const unsigned slot = sm.order[idx];
                    const float4 o4 = sm.ray_o[slot];
                    TraceHit h;
                    const float Ts = traverse_df<LAYOUT>(S, mk3(o4.x, o4.y, o4.z), P.stronger_dir, 128, h, cnt);
                    if (slot < (unsigned)GC_THREADS) sm.res_sh1[slot] = Ts > 0.0f ? 1 : 0;
                    else sm.res_sh0[slot - GC_THREADS] = Ts > 0.0f ? 1 : 0;
                }
            }
        }
        __syncthreads();
        GI_TRACE(1, blockIdx.x, 7 * gc_it + 5);   // stage D done
        // ---- stage E: end of the sample --------------------------------------------------------------------------------------------
        if (valid) {
            // the shader's sum, in its order (header comment)
            const V3 thr0 = mk3(1.f, 1.f, 1.f);
            const V3 pre0 = mk3(sm.st[ST_A0][tid], sm.st[ST_A1][tid], sm.st[ST_A2][tid]);
            V3 contrib = mk3(0.f, 0.f, 0.f);
            contrib = contrib + thr0 * sun_brdf(pre0, sm.res_sh0[tid] ? 1.0f : 0.0f);
            contrib = contrib + mk3(sm.st[ST_P0][tid], sm.st[ST_P1][tid], sm.st[ST_P2][tid]);
            if (hit1) {
                const V3 thr1 = mk3(sm.st[ST_T0][tid], sm.st[ST_T1][tid], sm.st[ST_T2][tid]);
                const V3 pre1 = mk3(sm.st[ST_C0][tid], sm.st[ST_C1][tid], sm.st[ST_C2][tid]);
                contrib = contrib + thr1 * sun_brdf(pre1, sm.res_sh1[tid] ? 1.0f : 0.0f);
            }
            contrib = contrib + mk3(sm.st[ST_Q0][tid], sm.st[ST_Q1][tid], sm.st[ST_Q2][tid]);
            finish_sample<SPP1>(out, state, (size_t)(unsigned)__float_as_int(sm.st[ST_PX][tid]), contrib, sm.st[ST_AO][tid],
                                mk3(sm.st[ST_OD0][tid], sm.st[ST_OD1][tid], sm.st[ST_OD2][tid]), skyhit, __float_as_int(sm.st[ST_BLS][tid]));
        }
        GI_TRACE(1, blockIdx.x, 7 * gc_it + 6);   // stage E done (thread 0)
    }
    flush_counters(S, cnt);
}

// (Tried and rejected, r02h / r02i: the continuation as five launches over all records of a slab — shade the first hit | trace the bounce
// rays | shade the second hit | trace every sun-shadow sub-ray | finish — each with the registers of its stage only (32..58) and the first
// hit's sub-ray moved behind the second's, so that a sample's dependent chain is 48 + 128 iterations.  Bit-identical planes, but the whole
// 1080p GI pass took 0.305 / 0.291 / 0.282 ms with 1024 / 256 / 512 bounce rays sorted per CTA against 0.278 ms with gi_continue: the three
// dense launches cost 11..15 us each (a few dependent table and texture reads per record at little parallelism) and the two traversal
// launches do not end sooner than the staged kernel's two traversal stages.)
__global__ void __launch_bounds__(256) gi_finalize(const __grid_constant__ CameraDev cam, const DiffuseDev P, const GBufferDev g, const DiffuseOutDev out,
                                                   const PixState* __restrict__ state) {
    int i, j, prow;
    if (!thread_pixel(cam, i, j, prow)) return;
    const size_t px = (size_t)prow * cam.width + i;
    if (load_f1(g.t, px, g.fmt) < 0.0f) return;
    const PixState st = state[px];
    write_final(out, px, mk3(st.rad.x, st.rad.y, st.rad.z), st.rad.w, st.tot.x, st.tot.y, st.tot.z, st.tot.w, st.misc.x, st.misc.y, st.misc.z,
                pixel_spp(P, i, j));
}

static CameraDev cam_to_dev(const VxCamera& cam) {
    CameraDev c;
    for (int k = 0; k < 16; ++k) { c.inv_view[k] = cam.inv_view[k]; c.inv_proj[k] = cam.inv_proj[k]; }
    c.width = cam.width; c.height = cam.height; c.row_begin = cam.row_begin; c.row_end = cam.row_end;
    c.il_n = cam.interleave_n; c.il_rank = cam.interleave_rank; c.il_band = cam.band_rows > 0 ? cam.band_rows : 1;
    return c;
}

// experiment knobs (environment, read once): VXPT_GI_SORT = pixels per thread of the sorted first-bounce kernel (0, 1, 2, 4),
// VXPT_GI_CTAS = gi_continue CTAs per SM
static int env_int(const char* name, int dflt) {
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

// One sample of a slab = gi_gen_trace0 (throughput-bound: 64 % of the issue slots busy) followed by gi_continue (latency-bound: three dependent
// traversals of up to 48 / 128 / 128 iterations per record, 32 % of the issue slots busy, ncu r02c).  Experiment knob VXPT_GI_SLABS = k cuts
// the slab into k row sub-slabs and runs the gi_continue of sub-slab j on a second, higher-priority stream beside the gi_gen_trace0 of
// sub-slab j + 1 (fork / join by events, still one CUDA graph).  Measured r02d, 1080p plains: 1 / 2 / 3 / 4 sub-slabs = 0.278 / 0.304 /
// 0.410 / 0.341 ms — gi_continue's duration is its longest dependency chain, not its record count, so the last sub-slab's continuation is
// exposed at full length while every sub-launch adds its own tail.  Default: 1 (off).
constexpr int GI_MAX_SLABS = 8;
constexpr int GI_COUNT_STRIDE = 8;   // words between the counter blocks of two sub-slabs (hit count, hit cursor)

template <int LAYOUT, bool SPP1>
static int run_wavefront(vxpt_ctx* c, const SceneDev& S, const CameraDev& cd, const DiffuseDev& d, const GBufferDev& g, const DiffuseOutDev& od,
                         PixState* state, HitRec* queue, unsigned* count, int max_spp) {
    const dim3 grid((cd.width + 31) / 32, (cd.row_end - cd.row_begin + 7) / 8);
    const int rows = cd.row_end - cd.row_begin;
    const size_t slab_px = (size_t)rows * cd.width;
    static const int sort_env = env_int("VXPT_GI_SORT", -1), ctas_env = env_int("VXPT_GI_CTAS", VXPT_GI_CONT_MINB), slabs_env = env_int("VXPT_GI_SLABS", -1), share_env = env_int("VXPT_GI_MIN_SHARE", 0);
    // pixels per thread of the sorted first-bounce kernel: 4 on large slabs (1024 rays sorted per CTA, 32 groups for 8 warps), 2 on
    // medium ones, plain on slabs too small to fill the GPU with such CTAs (r01g, 1080p GI pass: plain 0.339 ms, 1 / 2 / 4 pixels per
    // thread 0.364 / 0.286 / 0.283 ms)
    const int rpt = sort_env >= 0 ? sort_env : (slab_px >= (size_t)768 * 1024 ? 4 : (slab_px >= (size_t)128 * 1024 ? 2 : 0));
    const int tile_rows = 8 * std::max(rpt, 1);
    // sub-slabs: whole CTA tile rows each
    int n_slabs = slabs_env > 0 ? slabs_env : 1;
    n_slabs = std::max(1, std::min({n_slabs, GI_MAX_SLABS, (rows + tile_rows - 1) / tile_rows}));
    if (!c->gi_stream) n_slabs = 1;
    const int tiles = (rows + tile_rows - 1) / tile_rows;
    for (int s = 0; s < max_spp; ++s) {
        VX_CUDA(cudaMemsetAsync(count, 0, GI_MAX_SLABS * GI_COUNT_STRIDE * sizeof(unsigned), c->stream));  // hit count, hit cursor per sub-slab
        for (int k = 0; k < n_slabs; ++k) {
            CameraDev sc = cd;
            sc.row_begin = cd.row_begin + (tiles * k / n_slabs) * tile_rows;
            sc.row_end = k + 1 == n_slabs ? cd.row_end : cd.row_begin + (tiles * (k + 1) / n_slabs) * tile_rows;
            const int srows = sc.row_end - sc.row_begin;
            if (srows <= 0) continue;
            unsigned* cnt_k = count + k * GI_COUNT_STRIDE;
            HitRec* queue_k = queue + (size_t)(sc.row_begin - cd.row_begin) * cd.width;
            const dim3 grid_s((cd.width + 31) / 32, (srows + tile_rows - 1) / tile_rows);
            switch (rpt) {
                case 0: gi_gen_trace0<LAYOUT, SPP1, 0><<<grid_s, 256, 0, c->stream>>>(S, sc, d, g, od, state, queue_k, cnt_k, s); break;
                case 1: gi_gen_trace0<LAYOUT, SPP1, 1><<<grid_s, 256, 0, c->stream>>>(S, sc, d, g, od, state, queue_k, cnt_k, s); break;
                case 2: gi_gen_trace0<LAYOUT, SPP1, 2><<<grid_s, 256, 0, c->stream>>>(S, sc, d, g, od, state, queue_k, cnt_k, s); break;
                default: gi_gen_trace0<LAYOUT, SPP1, 4><<<grid_s, 256, 0, c->stream>>>(S, sc, d, g, od, state, queue_k, cnt_k, s); break;
            }
            // a CTA works on 256 records at a time; no more CTAs than the sub-slab can have chunks
            const unsigned max_chunks = (unsigned)(((size_t)srows * cd.width + GC_THREADS - 1) / GC_THREADS);
            const unsigned ctas = std::min<unsigned>(148u * (unsigned)std::max(ctas_env, 1), std::max(max_chunks, 1u));
            cudaStream_t ks = c->stream;
            if (n_slabs > 1) {  // fork: the continuation of this sub-slab goes to the second stream
                VX_CUDA(cudaEventRecord(c->ev_gi[k], c->stream));
                VX_CUDA(cudaStreamWaitEvent(c->gi_stream, c->ev_gi[k], 0));
                ks = c->gi_stream;
            }
            gi_continue<LAYOUT, SPP1><<<ctas, GC_THREADS, sizeof(GcShared), ks>>>(S, sc, d, od, state, queue_k, cnt_k, (unsigned)std::max(share_env, 0));
            c->launches += 2;
        }
        if (n_slabs > 1) {  // join
            VX_CUDA(cudaEventRecord(c->ev_gi[GI_MAX_SLABS], c->gi_stream));
            VX_CUDA(cudaStreamWaitEvent(c->stream, c->ev_gi[GI_MAX_SLABS], 0));
        }
    }
    if (!SPP1) {
        gi_finalize<<<grid, 256, 0, c->stream>>>(cd, d, g, od, state);
        c->launches += 1;
    }
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

// one-time, per process: gi_continue's shared memory is above the 48 KB a kernel gets without opting in
static int init_gi_kernels() {
    static bool done = false;
    if (done) return VXPT_OK;
    VX_CUDA(cudaFuncSetAttribute(gi_continue<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GcShared)));
    VX_CUDA(cudaFuncSetAttribute(gi_continue<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GcShared)));
    VX_CUDA(cudaFuncSetAttribute(gi_continue<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GcShared)));
    VX_CUDA(cudaFuncSetAttribute(gi_continue<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GcShared)));
    done = true;
    return VXPT_OK;
}

// wavefront scratch of a slab: 256 B of counters, the hit queue (one record per slab pixel at most) and, when some pixel takes several
// samples, the per-pixel sample state of the frame
size_t gi_scratch_bytes(size_t slab_px, size_t frame_px, bool spp1) { return 256 + slab_px * sizeof(HitRec) + (spp1 ? 0 : frame_px * sizeof(PixState)); }
static_assert(GI_MAX_SLABS * GI_COUNT_STRIDE * sizeof(unsigned) <= 256, "counter blocks fit the 256-byte header of the scratch");

int launch_diffuse_wavefront(vxpt_ctx* c, const VxCamera& cam, const DiffuseDev& d, const VxGBuffer& g, const VxDiffuseOut& out) {
    if (int rc = init_gi_kernels()) return rc;
    const SceneDev S = make_scene(c);
    const CameraDev cd = cam_to_dev(cam);
    const GBufferDev gd{g.t, g.normal_id, g.block_id, g.inv_t, g.hit_voxel, c->opt_texel};
    const DiffuseOutDev od{reinterpret_cast<float4*>(out.sh), reinterpret_cast<float2*>(out.cocg), out.luma, reinterpret_cast<float2*>(out.ao_sky),
                           c->opt_texel};
    // the largest per-pixel sample count (pixel_spp on the host)
    const int base = std::min(std::max(d.spp, 1), 32);
    int max_spp = base;
    if (d.checkerboard) max_spp = std::max(std::min(std::max(d.spp, 1), 32), std::min(std::max(d.checker_spp, 1), 32));
    if (d.moon_stronger) max_spp *= 2;
    const bool spp1 = max_spp == 1;
    const size_t slab_px = (size_t)(cam.row_end - cam.row_begin) * cam.width, frame_px = (size_t)cam.width * cam.height;
    // grown on demand — but never while the stream is being captured into a CUDA graph (vxpt_reserve sizes it beforehand)
    if (int rc = grow_scratch(c, &c->d_queue, &c->queue_bytes, gi_scratch_bytes(slab_px, frame_px, spp1), "the GI wavefront queue")) return rc;
    unsigned* count = static_cast<unsigned*>(c->d_queue);
    HitRec* queue = reinterpret_cast<HitRec*>(static_cast<char*>(c->d_queue) + 256);
    PixState* state = spp1 ? nullptr : reinterpret_cast<PixState*>(static_cast<char*>(c->d_queue) + 256 + slab_px * sizeof(HitRec));
    if (c->opt_layout == 1)
        return spp1 ? run_wavefront<1, true>(c, S, cd, d, gd, od, state, queue, count, max_spp)
                    : run_wavefront<1, false>(c, S, cd, d, gd, od, state, queue, count, max_spp);
    return spp1 ? run_wavefront<0, true>(c, S, cd, d, gd, od, state, queue, count, max_spp)
                : run_wavefront<0, false>(c, S, cd, d, gd, od, state, queue, count, max_spp);
}

}  // namespace vxpt

#ifdef VXPT_GI_TRACE
extern "C" __attribute__((visibility("default"))) int vxpt_debug_gi_trace(unsigned long long* out /* [2][4096][16] */) {
    return (int)cudaMemcpyFromSymbol(out, vxpt::g_gi_trace, sizeof(vxpt::g_gi_trace));
}
extern "C" __attribute__((visibility("default"))) int vxpt_debug_gi_trace_clear() {
    static unsigned long long zero[2 * 4096 * 16];
    return (int)cudaMemcpyToSymbol(vxpt::g_gi_trace, zero, sizeof(zero));
}
#endif
