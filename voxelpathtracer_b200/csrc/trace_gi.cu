// trace_gi.cu — wavefront form of the diffuse-GI pass (DiffuseRayTraceFrag.glsl main() :822-935, CalculateDiffuse :535-664).
//
// The one-thread-per-pixel form (diffuse_kernel in trace.cu) runs 9 of 32 lanes per instruction on open terrain: most
// first-bounce rays escape to the sky while the few that hit geometry drag their warp through texture fetches, a shadow
// sub-ray, a second traversal and a second shadow sub-ray (ncu r01a).  Here the work of one sample is re-queued:
//
//   gi_gen_trace0  one thread per pixel: first hemisphere direction (blue-noise tables), first traversal.  A miss is
//                  finished on the spot (sky radiance).  Hits are compacted into a queue with one warp ballot +
//                  one atomicAdd per warp (popc / prefix ranks give each lane its slot).
//   (opt-in, VXPT_OPT_GI_WAVEFRONT = 2: gi_gen -> gi_trace0 -> gi_resolve.  gi_gen queues the first-bounce rays, gi_trace0 is a
//                  persistent tracer whose lanes are refilled from the ray queue as their rays end (resumable traversal:
//                  trav_init / trav_step / trav_finish; retire + refill batched until 8 lanes need it), gi_resolve shades the
//                  misses densely and compacts the hits.  Measured slower than the fused kernel, see run_wavefront.)
//   gi_continue    one thread per queued hit (dense warps): shading of the first hit, its sun shadow sub-ray, the
//                  second traversal, shading + shadow sub-ray of the second hit, end of the sample.
//   gi_finalize    only when some pixel takes more than one sample: per-pixel averages and clamps (:915-934).
//
// Samples of one pixel are processed one after another (launch s+1 after launch s), so the blue-noise dimension counter
// and the accumulation order are exactly those of the shader's loop; the planes are bit-identical to diffuse_kernel's.
#include <algorithm>
#include <cstdlib>

#include "gi_device.cuh"

namespace vxpt {

struct RayRec {   // 32 B
    float4 a;     // ro.xyz, pixel index (bits)
    float4 b;     // rd.xyz, bl_sample (bits)
};
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

// shading of one hit (:569-622).  NEXT: also draw the next direction and update the throughput / ray.
template <int LAYOUT, bool NEXT>
__device__ __forceinline__ void shade_hit(const SceneDev& S, const DiffuseDev& P, int pi, int pj, int& bl_sample, V3& ro, V3& rd, float T, int min_idx,
                                          int sgn, int block, V3& thr, V3& contrib, Counters& cnt) {
    const float bias = 0.06f;
    const int tex_ref = min(max(block, 0), 127);
    const V3 ipos = ro + (rd * T);
    const float s = (float)(-sgn);
    const V3 hn = mk3(min_idx == 0 ? s : 0.0f, min_idx == 1 ? s : 0.0f, min_idx == 2 ? s : 0.0f);
    float tu, tv;
    calc_uv(ipos, min_idx, tu, tv);
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
    const V3 sunbrdf = (((albedo * diffuse_hammon(hn, neg_rd, P.stronger_dir, pbr.x)) * (P.light_color * 3.5f)) * (1.0f - shadow_at)) * PI_F;
    contrib = contrib + thr * sunbrdf;
    contrib = contrib + emis_color * thr;
    if (NEXT) {
        const V3 new_dir = cos_hemisphere(S, pi, pj, P.frame % 128, bl_sample, hn);
        const float cos_theta = clampf(dot3(hn, new_dir), 0.0f, 1.0f);
        const float pdf = fmaxf(cos_theta / PI_F, 0.00001f);
        const V3 atten = mk3(1.f, 1.f, 1.f) * diffuse_hammon(hn, neg_rd, new_dir, pbr.x);
        thr = thr * ((albedo * atten) / pdf);
        rd = new_dir;
        ro = ipos + hn * bias;
    } else {
        bl_sample += 2;  // the shader still draws the direction of a third segment it never traces (:607)
    }
}

// one warp ballot + one atomic per warp; returns this lane's slot (valid where pred)
__device__ __forceinline__ unsigned warp_push(unsigned* counter, bool pred) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    const unsigned lane = threadIdx.x & 31;
    unsigned base = 0;
    if (lane == 0 && mask) base = atomicAdd(counter, __popc(mask));
    base = __shfl_sync(0xffffffffu, base, 0);
    return base + __popc(mask & ((1u << lane) - 1u));
}

// gi_gen_trace0 — first hemisphere direction and first traversal of every pixel of a slab, with the rays of a CTA SORTED before they are
// traced.  How long a hemisphere ray lives is decided mostly by where it points: steep rays reach large step values within a few
// iterations and leave the volume (or hit the ground at once), grazing rays creep through E = 1..3 cells until the cap.  Neighbouring
// pixels draw unrelated directions, so an unsorted warp waits for its longest ray with half of its lanes idle (ncu r01g: 15.8 of 32
// active).  Here a CTA of 256 threads generates the rays of RPT 32x8-pixel tiles (RPT per thread), sorts them by |rd.y| (counting sort on
// an 8-bit key in shared memory) and its warps then claim groups of 32 rays from the sorted list, longest-lived first, until the list is
// empty: a warp holds rays of similar life expectancy, and short groups fill the time the warps that drew long groups are still busy
// (with one group per warp the CTA would keep its registers until its slowest warp ended).  Which thread traces a ray does not change
// what is computed for its pixel, so the planes stay bit-identical.  RPT = 0 selects the plain form (one pixel per thread, no exchange).
template <int LAYOUT, bool SPP1, int RPT>
__global__ void __launch_bounds__(256) gi_gen_trace0(const SceneDev S, const __grid_constant__ CameraDev cam, const DiffuseDev P, const GBufferDev g,
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
        const V3 normal = normal_from_id(g.normal_id[px], 0.5f);
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
            const V3 d = cos_hemisphere(S, i, j, P.frame % 128, bls, normal);
            if (SORT) {
                const unsigned key = (unsigned)min((int)(fabsf(d.y) * 255.0f), 255);  // small = grazing = long-lived
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
    if (tid < 32) {  // exclusive prefix over the 256 bins: 8 bins per lane, then a warp scan of the lane sums
        unsigned c[8], sum = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) { c[k] = s_hist[tid * 8 + k]; sum += c[k]; }
        unsigned incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned t = __shfl_up_sync(0xffffffffu, incl, d);
            if (tid >= (unsigned)d) incl += t;
        }
        unsigned base = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; ++k) { s_hist[tid * 8 + k] = base; base += c[k]; }
        if (tid == 31) s_hist[256] = incl;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < NR; ++r)
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
    flush_counters(S, cnt);
}

// first half of gi_gen_trace0: per pixel set-up, rays into the queue
template <bool SPP1>
__global__ void __launch_bounds__(256) gi_gen(const SceneDev S, const __grid_constant__ CameraDev cam, const DiffuseDev P, const GBufferDev g,
                                              const DiffuseOutDev out, PixState* __restrict__ state, RayRec* __restrict__ rays,
                                              unsigned* __restrict__ counters, const int sample) {
    int i, j, prow;
    const bool active = thread_pixel(cam, i, j, prow);
    bool push = false;
    RayRec rec;
    if (active) {
        const size_t px = (size_t)prow * cam.width + i;
        float u = ((float)i + 0.5f) / (float)cam.width;
        float v = ((float)j + 0.5f) / (float)cam.height;
        const float u0 = u, v0 = v;
        if (P.supersample) {
            u += (P.hx * 0.75f) / (float)cam.width;
            v += (P.hy * 0.75f) / (float)cam.height;
        }
        const float dist = load_f1(g.t, px, g.fmt);
        const V3 normal = normal_from_id(g.normal_id[px], 0.5f);
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
            int bl_sample = 0;
            if (!SPP1) {
                if (sample == 0) {
                    PixState z;
                    z.tot = z.rad = z.misc = make_float4(0.f, 0.f, 0.f, 0.f);
                    state[px] = z;
                } else {
                    bl_sample = __float_as_int(state[px].misc.w);
                }
            }
            const V3 pos = ray_origin(cam) + normalize3(ray_direction_at(cam, u, v)) * dist;
            const V3 ro = pos + normal * 0.06f;
            const V3 rd = cos_hemisphere(S, i, j, P.frame % 128, bl_sample, normal);
            push = true;
            rec.a = make_float4(ro.x, ro.y, ro.z, __int_as_float((int)px));
            rec.b = make_float4(rd.x, rd.y, rd.z, __int_as_float(bl_sample));
        }
    }
    const unsigned slot = warp_push(counters + 2, push);
    if (push) rays[slot] = rec;
}

// persistent first-bounce traversal with lane refill.  counters: [0] hit count, [1] hit cursor, [2] ray count, [3] ray cursor
// A lane is idle (0), tracing (1) or done (3).  The main loop only advances rays; finished rays are
// retired (final block fetch + 8-byte result) and idle lanes refilled in one maintenance step that runs when THRESH lanes need it
// (or nothing else is left to do), so the maintenance code runs with many lanes instead of one or two per iteration, and nothing
// but traversal is executed by partially filled warps.  Shading of the results happens in gi_resolve, densely.
template <int LAYOUT, int THRESH>
__global__ void __launch_bounds__(256) gi_trace0(const SceneDev S, const DiffuseDev P, const RayRec* __restrict__ rays, float2* __restrict__ results,
                                                 unsigned* __restrict__ counters) {
    const unsigned n_rays = counters[2];
    const unsigned lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    Counters cnt = {0u, 0u, 0u};
    TravState t;
    V3 origin = mk3(0.f, 0.f, 0.f);
    unsigned idx = 0;
    int phase = 0;
    bool exhausted = false;
    trav_init(t, origin, mk3(1.f, 1.f, 1.f), 0);
    while (true) {
        const unsigned done_mask = __ballot_sync(0xffffffffu, phase == 3);
        const unsigned idle_mask = __ballot_sync(0xffffffffu, phase == 0);
        const unsigned live_mask = ~(done_mask | idle_mask);
        if (live_mask == 0u || __popc(done_mask) + (exhausted ? 0 : __popc(idle_mask)) >= THRESH) {
            if (phase == 3) {  // retire
                TraceHit h;
                const float T = trav_finish(S, t, origin, h, cnt);
                results[idx] = make_float2(T, __int_as_float(h.min_idx | ((h.sgn + 1) << 2) | (h.block << 8)));
                phase = 0;
            }
            const unsigned want = done_mask | idle_mask;  // every one of these lanes is idle now
            if (!exhausted) {  // refill
                unsigned base = 0;
                if (lane == 0) base = atomicAdd(counters + 3, (unsigned)__popc(want));
                base = __shfl_sync(0xffffffffu, base, 0);
                if (base + __popc(want) >= n_rays) exhausted = true;
                if (phase == 0) {
                    const unsigned my = base + __popc(want & lt_mask);
                    if (my < n_rays) {
                        const RayRec r = rays[my];
                        origin = mk3(r.a.x, r.a.y, r.a.z);
                        trav_init(t, origin, mk3(r.b.x, r.b.y, r.b.z), P.trace_length);
                        idx = my;
                        cnt.rays++;
                        phase = 1;
                    }
                }
            }
            if (!__any_sync(0xffffffffu, phase != 0)) break;
        }
        if (phase == 1 && trav_step<LAYOUT>(S, t)) phase = 3;
    }
    flush_counters(S, cnt);
}

// one thread per traced first-bounce ray: a miss ends the sample here (sky radiance), a hit goes to the hit queue
template <bool SPP1>
__global__ void __launch_bounds__(256) gi_resolve(const SceneDev S, const DiffuseDev P, const DiffuseOutDev out, PixState* __restrict__ state,
                                                  const RayRec* __restrict__ rays, const float2* __restrict__ results, HitRec* __restrict__ hits,
                                                  unsigned* __restrict__ counters) {
    const unsigned n_rays = counters[2];
    const unsigned idx = blockIdx.x * 256u + threadIdx.x;
    if (blockIdx.x * 256u >= n_rays) return;
    bool push = false;
    HitRec rec;
    if (idx < n_rays) {
        const RayRec r = rays[idx];
        const float2 q = results[idx];
        const float T = q.x;
        const int code = __float_as_int(q.y);
        const V3 rd = mk3(r.b.x, r.b.y, r.b.z);
        if (T > 0.0f && ((code >> 8) & 255) > 0) {
            push = true;
            rec.a = make_float4(r.a.x, r.a.y, r.a.z, T);
            rec.b = make_float4(rd.x, rd.y, rd.z, r.a.w);
            rec.c = make_float4(q.y, r.b.w, 0.f, 0.f);
        } else {
            const V3 contrib = mk3(0.f, 0.f, 0.f) + sky_term(S, P, rd) * mk3(1.f, 1.f, 1.f);
            finish_sample<SPP1>(out, state, (size_t)(unsigned)__float_as_int(r.a.w), contrib, 1.0f, rd, true, __float_as_int(r.b.w));
        }
    }
    const unsigned slot = warp_push(counters + 0, push);
    if (push) hits[slot] = rec;
}

template <int LAYOUT, bool SPP1>
__global__ void __launch_bounds__(128) gi_continue(const SceneDev S, const __grid_constant__ CameraDev cam, const DiffuseDev P, const DiffuseOutDev out,
                                                   PixState* __restrict__ state, const HitRec* __restrict__ queue,
                                                   unsigned* __restrict__ queue_count) {
    const unsigned count = queue_count[0];
    unsigned* cursor = queue_count + 1;
    Counters cnt = {0u, 0u, 0u};
    const unsigned lane = threadIdx.x & 31;
    // dynamic work distribution: each warp claims the next 32 queued hits until the queue is drained, so the (device-side)
    // hit count needs no host round trip and the tail is balanced across the resident warps
    while (true) {
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(cursor, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= count) break;
        const unsigned idx = base + lane;
        if (idx >= count) continue;
        const HitRec rec = queue[idx];
        V3 ro = mk3(rec.a.x, rec.a.y, rec.a.z), rd = mk3(rec.b.x, rec.b.y, rec.b.z);
        const float T0 = rec.a.w;
        const size_t px = (size_t)(unsigned)__float_as_int(rec.b.w);
        const int pi = (int)(px % (size_t)cam.width), pj = image_row(cam, (int)(px / (size_t)cam.width));
        const int code = __float_as_int(rec.c.x);
        int bl_sample = __float_as_int(rec.c.y);
        const V3 odir = rd;
        V3 thr = mk3(1.f, 1.f, 1.f), contrib = mk3(0.f, 0.f, 0.f);
        bool skyhit = false;
        // bounce 0 hit
        shade_hit<LAYOUT, true>(S, P, pi, pj, bl_sample, ro, rd, T0, code & 3, ((code >> 2) & 3) - 1, (code >> 8) & 255, thr, contrib, cnt);
        float ao = 1.0f;
        if (T0 < 2.0f && T0 > 0.0f) ao = fmaxf(T0 / 2.0f, 0.0f);  // :637-650
        // bounce 1
        TraceHit h;
        const float T1 = traverse_df<LAYOUT>(S, ro, rd, P.trace_length, h, cnt);
        if (T1 > 0.0f && h.block > 0) {
            shade_hit<LAYOUT, false>(S, P, pi, pj, bl_sample, ro, rd, T1, h.min_idx, h.sgn, h.block, thr, contrib, cnt);
        } else {
            contrib = contrib + sky_term(S, P, rd) * thr;
            skyhit = true;
        }
        finish_sample<SPP1>(out, state, px, contrib, ao, odir, skyhit, bl_sample);
    }
    flush_counters(S, cnt);
}

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

template <int LAYOUT, bool SPP1>
static int run_wavefront(vxpt_ctx* c, const SceneDev& S, const CameraDev& cd, const DiffuseDev& d, const GBufferDev& g, const DiffuseOutDev& od,
                         PixState* state, HitRec* queue, RayRec* rays, float2* results, unsigned* count, int max_spp) {
    const dim3 grid((cd.width + 31) / 32, (cd.row_end - cd.row_begin + 7) / 8);
    const size_t slab_px = (size_t)(cd.row_end - cd.row_begin) * cd.width;
    // opt-in (VXPT_OPT_GI_WAVEFRONT = 2): measured r01g at 1080p, gi_gen 47 us + gi_trace0 188 us + gi_resolve 43 us against 240 us for the
    // fused gi_gen_trace0 — the refill keeps 25 of 32 lanes tracing, but the skip / DDA halves of an iteration still diverge (16 and 10
    // lanes) and the queue round trips cost more than the idle lanes did
    const bool persistent = c->opt_wavefront == 2;
    for (int s = 0; s < max_spp; ++s) {
        VX_CUDA(cudaMemsetAsync(count, 0, 4 * sizeof(unsigned), c->stream));  // hit count, hit cursor, ray count, ray cursor
        if (persistent) {
            gi_gen<SPP1><<<grid, 256, 0, c->stream>>>(S, cd, d, g, od, state, rays, count, s);
            static const int tune = std::getenv("VXPT_GI_TUNE") ? std::atoi(std::getenv("VXPT_GI_TUNE")) : 0;  // experiment knob
            const int ctas = 148 * (tune >= 10 ? tune / 10 : 5);  // 47 registers: 5 CTAs of 256 threads per SM
            switch (tune % 10) {
                case 1: gi_trace0<LAYOUT, 4><<<ctas, 256, 0, c->stream>>>(S, d, rays, results, count); break;
                case 2: gi_trace0<LAYOUT, 16><<<ctas, 256, 0, c->stream>>>(S, d, rays, results, count); break;
                case 3: gi_trace0<LAYOUT, 12><<<ctas, 256, 0, c->stream>>>(S, d, rays, results, count); break;
                case 4: gi_trace0<LAYOUT, 1><<<ctas, 256, 0, c->stream>>>(S, d, rays, results, count); break;
                case 5: gi_trace0<LAYOUT, 24><<<ctas, 256, 0, c->stream>>>(S, d, rays, results, count); break;
                default: gi_trace0<LAYOUT, 8><<<ctas, 256, 0, c->stream>>>(S, d, rays, results, count); break;
            }
            gi_resolve<SPP1><<<(unsigned)((slab_px + 255) / 256), 256, 0, c->stream>>>(S, d, od, state, rays, results, queue, count);
            c->launches += 3;
        } else {
            // pixels per thread of the sorted first-bounce kernel: 4 on large slabs (1024 rays sorted per CTA, 32 groups for 8 warps), 2 on
            // medium ones, plain on slabs too small to fill the GPU with such CTAs (r01g, 1080p GI pass: plain 0.339 ms, 1 / 2 / 4 pixels per
            // thread 0.364 / 0.286 / 0.283 ms); VXPT_GI_SORT (0, 1, 2, 4) overrides (experiment knob)
            static const int sort_env = std::getenv("VXPT_GI_SORT") ? std::atoi(std::getenv("VXPT_GI_SORT")) : -1;
            const int rows = cd.row_end - cd.row_begin;
            const int rpt = sort_env >= 0 ? sort_env : (slab_px >= (size_t)768 * 1024 ? 4 : (slab_px >= (size_t)128 * 1024 ? 2 : 0));
            const dim3 grid_s((cd.width + 31) / 32, (rows + 8 * std::max(rpt, 1) - 1) / (8 * std::max(rpt, 1)));
            switch (rpt) {
                case 0: gi_gen_trace0<LAYOUT, SPP1, 0><<<grid_s, 256, 0, c->stream>>>(S, cd, d, g, od, state, queue, count, s); break;
                case 1: gi_gen_trace0<LAYOUT, SPP1, 1><<<grid_s, 256, 0, c->stream>>>(S, cd, d, g, od, state, queue, count, s); break;
                case 2: gi_gen_trace0<LAYOUT, SPP1, 2><<<grid_s, 256, 0, c->stream>>>(S, cd, d, g, od, state, queue, count, s); break;
                default: gi_gen_trace0<LAYOUT, SPP1, 4><<<grid_s, 256, 0, c->stream>>>(S, cd, d, g, od, state, queue, count, s); break;
            }
            c->launches += 1;
        }
        // (tried r01h: the same shared-memory sort for the second-bounce rays — 13.9 instead of 11.4 lanes active, 50 M instead of 54 M
        // warp instructions, but the same 106 us: the kernel is bound by the dependent chain texture -> shadow ray -> bounce, not by issue)
        // (5, 6, 7 or 10 CTAs per SM: the same 0.281 ms for the pass; 4: 0.312 ms)
        gi_continue<LAYOUT, SPP1><<<148 * 6, 128, 0, c->stream>>>(S, cd, d, od, state, queue, count);
        c->launches += 1;
    }
    if (!SPP1) {
        gi_finalize<<<grid, 256, 0, c->stream>>>(cd, d, g, od, state);
        c->launches += 1;
    }
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

int launch_diffuse_wavefront(vxpt_ctx* c, const VxCamera& cam, const DiffuseDev& d, const VxGBuffer& g, const VxDiffuseOut& out) {
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
    // queues: slab-sized hit queue (+ full-frame per-pixel state when some pixel takes several samples)
    const size_t slab_px = (size_t)(cam.row_end - cam.row_begin) * cam.width, frame_px = (size_t)cam.width * cam.height;
    const size_t need = 256 + slab_px * (sizeof(HitRec) + sizeof(RayRec) + sizeof(float2)) + (spp1 ? 0 : frame_px * sizeof(PixState));
    if (need > c->queue_bytes) {
        VX_CUDA(cudaStreamSynchronize(c->stream));
        if (c->d_queue) cudaFree(c->d_queue);
        c->d_queue = nullptr;
        c->queue_bytes = 0;
        if (cudaMalloc(&c->d_queue, need) != cudaSuccess) {
            cudaGetLastError();
            set_error("wavefront queue allocation failed");
            return VXPT_E_NOMEM;
        }
        c->queue_bytes = need;
    }
    unsigned* count = static_cast<unsigned*>(c->d_queue);
    HitRec* queue = reinterpret_cast<HitRec*>(static_cast<char*>(c->d_queue) + 256);
    RayRec* rays = reinterpret_cast<RayRec*>(static_cast<char*>(c->d_queue) + 256 + slab_px * sizeof(HitRec));
    float2* results = reinterpret_cast<float2*>(static_cast<char*>(c->d_queue) + 256 + slab_px * (sizeof(HitRec) + sizeof(RayRec)));
    PixState* state =
        spp1 ? nullptr : reinterpret_cast<PixState*>(static_cast<char*>(c->d_queue) + 256 + slab_px * (sizeof(HitRec) + sizeof(RayRec) + sizeof(float2)));
    if (c->opt_layout == 1)
        return spp1 ? run_wavefront<1, true>(c, S, cd, d, gd, od, state, queue, rays, results, count, max_spp)
                    : run_wavefront<1, false>(c, S, cd, d, gd, od, state, queue, rays, results, count, max_spp);
    return spp1 ? run_wavefront<0, true>(c, S, cd, d, gd, od, state, queue, rays, results, count, max_spp)
                : run_wavefront<0, false>(c, S, cd, d, gd, od, state, queue, rays, results, count, max_spp);
}

}  // namespace vxpt
