// df_consumers.cu — the other consumers of the distance field, sm_100a.
//
//   rays_kernel     : a batch of VoxelTraversalDF calls on caller-supplied rays.  Core/Shaders/PostProcessingVert.glsl:46-53 traces
//                     ONE such ray per frame (camera -> sun, cap 350, :103-170 is the same function as InitialRayTraceFrag.glsl:307-374)
//                     to decide whether the player's eye is in shadow (v_PlayerShadowed); vxpt_player_shadowed wraps that call.
//   ambient_kernel  : Core/Shaders/EstimateAmbientSoundLevel.comp main() :247-270 as dispatched by Core/Pipeline.cpp:1908-1921
//                     (glDispatchCompute(2, 1, 1) x local size 4 x 4 = 32 invocations): hash-seeded hemisphere rays from the player,
//                     up to 7 diffuse bounces through TraverseDistanceField (:73-141: the usual loop with a cap of 32 iterations, but the
//                     hit test reads the DISTANCE FIELD: -1 outside the volume passes `D < 0.0001f`, which matters when the cap ends the loop
//                     right after a step out of the volume), the escaped fraction summed into SkyLevelAggregate with atomicAdd.
// Compiled with -fmad=false; sin / cos are the pinned correctly rounded fp32 values.
#include "gi_device.cuh"

namespace vxpt {

struct RaysDev {
    const float* origins;     // 3 n
    const float* directions;  // 3 n
    int n, max_iterations;
    float* t;
    uint8_t* normal_id;
    uint8_t* block_id;
    int16_t* hit_voxel;
};

template <int LAYOUT>
__global__ void __launch_bounds__(256) rays_kernel(const SceneDev S, const RaysDev R) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    Counters cnt = {0u, 0u, 0u};
    if (k < R.n) {
        const V3 o = mk3(R.origins[3 * k], R.origins[3 * k + 1], R.origins[3 * k + 2]);
        const V3 d = mk3(R.directions[3 * k], R.directions[3 * k + 1], R.directions[3 * k + 2]);
        TraceHit h;
        const float t = traverse_df<LAYOUT>(S, o, d, R.max_iterations, h, cnt);
        const bool intersect = t > 0.0f && h.block > 0;
        if (R.t) R.t[k] = t;
        if (R.normal_id) R.normal_id[k] = intersect ? (uint8_t)normal_id_of(h) : (uint8_t)VXPT_NORMAL_MISS;
        if (R.block_id) R.block_id[k] = intersect ? (uint8_t)h.block : (uint8_t)0;
        if (R.hit_voxel) {
            R.hit_voxel[3 * k + 0] = intersect ? (int16_t)h.vx : (int16_t)-1;
            R.hit_voxel[3 * k + 1] = intersect ? (int16_t)h.vy : (int16_t)-1;
            R.hit_voxel[3 * k + 2] = intersect ? (int16_t)h.vz : (int16_t)-1;
        }
    }
    flush_counters(S, cnt);
}

// ---- EstimateAmbientSoundLevel.comp -------------------------------------------------------------------------------------
// HashRNG / Hash1 :145-166
__device__ __forceinline__ float ambient_hash1(unsigned& seed) {
    seed ^= 2747636419u;
    seed *= 2654435769u;
    seed ^= seed >> 16;
    seed *= 2654435769u;
    seed ^= seed >> 16;
    seed *= 2654435769u;
    return (float)seed / 4294967295.0f;  // the literal is 2^32 as a float
}

// TraverseDistanceField :73-141
template <int LAYOUT>
__device__ __forceinline__ float ambient_traverse(const SceneDev& S, const V3 origin, const V3 dir, V3& normal, Counters& cnt) {
    TravState t;
    trav_init(t, origin, dir, 32);
    cnt.rays++;
    // not trav_step(): its sticky "a DDA step happened" flag stands in for Intersection only where the final test reads the block grid
    // (0 outside the volume).  Here the final test reads the distance field, so leaving the volume must clear the flag as the shader does.
    bool intersection = false;
    while (t.n < 32) {
        const float bx = floor_biased(t.ox), by = floor_biased(t.oy), bz = floor_biased(t.oz);
        const int lx = biased_to_int(bx), ly = biased_to_int(by), lz = biased_to_int(bz);
        if (!in_volume_i(lx, ly, lz)) { intersection = false; break; }
        ++t.n;
        const int euclid = fetch_step<LAYOUT>(S, lx, ly, lz);
        if (euclid == 0) break;
        if (euclid == 1) {
            dda_advance(t, biased_to_float(bx), biased_to_float(by), biased_to_float(bz));
            intersection = true;
        } else {
            const float k = (float)(euclid - 1);
            t.ox = t.ox + k * t.dx;
            t.oy = t.oy + k * t.dy;
            t.oz = t.oz + k * t.dz;
        }
    }
    cnt.df += t.n;
    if (!intersection) return -1.0f;
    const int sg = (t.min_idx == 0) ? t.sx : ((t.min_idx == 1) ? t.sy : t.sz);
    const float s = (float)(-sg);
    normal = mk3(t.min_idx == 0 ? s : 0.0f, t.min_idx == 1 ? s : 0.0f, t.min_idx == 2 ? s : 0.0f);
    // D = GetDistance(floor(origin)): -1 outside the volume (a hit: D < 0.0001f), else the field value, zero exactly on solid voxels
    const int x = biased_to_int(floor_biased(t.ox)), y = biased_to_int(floor_biased(t.oy)), z = biased_to_int(floor_biased(t.oz));
    bool hit = true;
    if (in_volume_i(x, y, z)) {
        cnt.df++;
        hit = fetch_step<LAYOUT>(S, x, y, z) == 0;  // E == 0 <=> M == 0
    }
    return hit ? length3(mk3(t.ox, t.oy, t.oz) - origin) : -1.0f;
}

// RaytraceAverageAmbience :210-241 (UniformHemisphere :183-190, CosineHemisphereDirection :192-204; PI = 3.141592653 is the same float
// as gi_device.cuh's PI_F)
template <int LAYOUT>
__device__ __forceinline__ float ambient_sample(const SceneDev& S, const V3 player, unsigned& seed, Counters& cnt) {
    const V3 up = mk3(0.0f, 1.0f, 0.0f);
    V3 ro = player;
    V3 rd;
    {
        const float ux = ambient_hash1(seed), uy = ambient_hash1(seed);  // Hash2(): left to right
        const float r = sqrtf(1.0f - ux * ux);
        const float phi = (2.0f * PI_F) * uy;
        const V3 B = normalize3(cross3(up, mk3(0.0f, 1.0f, 1.0f)));
        const V3 T = cross3(B, up);
        rd = normalize3(((r * sin_cr(phi)) * B + ux * up) + (r * cos_cr(phi)) * T);
    }
#pragma unroll 1
    for (int bounce = 0; bounce < 7; ++bounce) {
        V3 N = mk3(0.0f, 0.0f, 0.0f);
        const float T = ambient_traverse<LAYOUT>(S, ro, rd, N, cnt);
        if (T < 0.0f) return 1.0f;  // Throughput (never attenuated, :236)
        ro = (ro + rd * T) + N * 0.05f;
        const float r1 = ambient_hash1(seed), r2 = ambient_hash1(seed);
        const float PI2 = 2.0f * PI_F;
        const V3 uu = normalize3(cross3(N, mk3(0.0f, 1.0f, 1.0f)));
        const V3 vv = cross3(uu, N);
        const float ra = sqrtf(r2);
        const float rx = ra * cos_cr(PI2 * r1), ry = ra * sin_cr(PI2 * r1), rz = sqrtf(1.0f - r2);
        rd = normalize3((rx * uu + ry * vv) + rz * N);
    }
    return 0.0f;
}

template <int LAYOUT>
__global__ void __launch_bounds__(32) ambient_kernel(const SceneDev S, const V3 player, const int frame, unsigned* aggregate, unsigned* per_invocation) {
    const int ix = threadIdx.x & 7, iy = threadIdx.x >> 3;  // gl_GlobalInvocationID.xy of an 8 x 4 dispatch
    Counters cnt = {0u, 0u, 0u};
    // InitRNG(vec2(Invocation), vec2(32.0f)) :155-158
    unsigned seed = (unsigned)((float)iy * 32.0f + (float)ix) + (unsigned)(frame % 512) * 32u * 32u;
    const int samples = (frame % 2 == 0) ? 2 : 1;  // int(mix(1, 2, float(u_Frame % 2 == 0)))
    float amount = 0.0f, weight = 0.0f;
    for (int s = 0; s < samples; ++s) {
        amount += ambient_sample<LAYOUT>(S, player, seed, cnt);
        weight += 1.0f;
    }
    amount /= weight;
    const unsigned mapped = (unsigned)clampf(amount * 512.0f, 0.0f, 512.0f);
    atomicAdd(aggregate, mapped);
    if (per_invocation) per_invocation[threadIdx.x] = mapped;
    flush_counters(S, cnt);
}

// ------------------------------------------------------------------------------------------------------------------- launchers
int launch_rays(vxpt_ctx* c, const float* origins, const float* directions, int n, int max_iterations, float* t, uint8_t* normal_id,
                uint8_t* block_id, int16_t* hit_voxel) {
    const SceneDev S = make_scene(c);
    const RaysDev R{origins, directions, n, max_iterations, t, normal_id, block_id, hit_voxel};
    const dim3 grid((n + 255) / 256, 1);
    if (c->opt_layout == 1) VX_LAUNCH((rays_kernel<1>), grid, 256, c->stream, S, R);
    else VX_LAUNCH((rays_kernel<0>), grid, 256, c->stream, S, R);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

int launch_ambient(vxpt_ctx* c, const float player[3], int frame, unsigned* aggregate, unsigned* per_invocation) {
    const SceneDev S = make_scene(c);
    const V3 p{player[0], player[1], player[2]};
    const dim3 grid(1, 1);
    if (c->opt_layout == 1) VX_LAUNCH((ambient_kernel<1>), grid, 32, c->stream, S, p, frame, aggregate, per_invocation);
    else VX_LAUNCH((ambient_kernel<0>), grid, 32, c->stream, S, p, frame, aggregate, per_invocation);
    c->launches += 1;
    VX_CUDA(cudaGetLastError());
    return VXPT_OK;
}

}  // namespace vxpt
