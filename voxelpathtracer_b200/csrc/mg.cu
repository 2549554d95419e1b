// mg.cu — vxpt_mg_*: one frame over N devices from ONE host thread.
//
// The reference drives one GL context from one thread (Core/Pipeline.cpp main loop).  vxpt_mg_* keeps that shape for a caller that owns
// several GPUs in a single process: a vxpt_mg_handle is a set of ordinary vxpt handles, one per device; scene state is REPLICATED (every
// vxpt_mg_set_* / upload / edit / build is the same call on every device — the grid is 18.9 MB, an edit list is bytes, a rebuild costs
// tens of microseconds, so nothing is broadcast device to device); a frame is SHARDED as contiguous image row slabs, device k tracing
// rows [H k / n, H (k + 1) / n) rounded to 8 rows, through vxpt_render_frame_async on its own stream, so the devices run concurrently.
// The gather needs no exchange step and no collective:
//   * HOST output planes: every device copies its own rows into the caller's planes (its copy stream, overlapped with its tracing);
//   * DEVICE output planes (memory of device 0, or of any device the others can reach): peer access is enabled between the devices at
//     creation, and the kernels of device k store their rows straight into those planes over NVLink (unified addressing).
// The per-process form (one process per GPU, planes pushed into the gather root's memory through CUDA IPC mappings, flag words ordering
// the frames) is vxpt_shared_* / vxpt_signal* / vxpt_wait* in api.cu, driven by voxelpathtracer_b200/multigpu.py under torch.distributed.
#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "vxpt_internal.h"

struct vxpt_mg_ctx {
    std::vector<vxpt_handle> dev;  // one handle per device, in rank order
    std::vector<int> ids;
    bool peer_ok = true;           // every device can reach every other device's memory
};

namespace {
int mg_fail(int code, const char* msg) {
    vxpt::set_error(msg);
    return code;
}
// rows of the frame selected by `cam` that device k of n traces: contiguous slabs cut at multiples of 8 rows (the kernels' tile height)
void slab_rows(const VxCamera* cam, int k, int n, int* rb, int* re) {
    const int rows = cam->row_end - cam->row_begin;
    auto cut = [&](int i) { return i >= n ? rows : std::min(rows, ((int)((long long)rows * i / n) + 7) & ~7); };
    *rb = cam->row_begin + cut(k);
    *re = cam->row_begin + cut(k + 1);
}
}  // namespace

#define MG_EACH(call)                                   \
    do {                                                \
        if (!mg) return mg_fail(VXPT_E_INVALID, "multi-GPU handle is NULL"); \
        for (vxpt_handle h : mg->dev) {                 \
            const int rc_ = (call);                     \
            if (rc_) return rc_;                        \
        }                                               \
        return VXPT_OK;                                 \
    } while (0)

extern "C" {

int vxpt_mg_create(int n_devices, const int* device_ids, vxpt_mg_handle* out) {
    if (!out || n_devices < 1 || n_devices > 64) return mg_fail(VXPT_E_INVALID, "bad device count");
    *out = nullptr;
    vxpt_mg_ctx* mg = new (std::nothrow) vxpt_mg_ctx;
    if (!mg) return mg_fail(VXPT_E_NOMEM, "out of host memory");
    for (int k = 0; k < n_devices; ++k) {
        vxpt_handle h = nullptr;
        const int id = device_ids ? device_ids[k] : k;
        const int rc = vxpt_create(id, &h);
        if (rc) {
            const std::string why = vxpt_last_error();
            for (vxpt_handle g : mg->dev) vxpt_destroy(g);
            delete mg;
            vxpt::set_error("vxpt_mg_create: device " + std::to_string(id) + ": " + why);
            return rc;
        }
        mg->dev.push_back(h);
        mg->ids.push_back(id);
    }
    // peer access, both ways, between distinct devices: lets a kernel of device k store into planes that live on device 0
    for (int a = 0; a < n_devices; ++a)
        for (int b = 0; b < n_devices; ++b) {
            if (mg->ids[a] == mg->ids[b]) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, mg->ids[a], mg->ids[b]) != cudaSuccess || !can) {
                cudaGetLastError();
                mg->peer_ok = false;
                continue;
            }
            cudaSetDevice(mg->ids[a]);
            const cudaError_t e = cudaDeviceEnablePeerAccess(mg->ids[b], 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) mg->peer_ok = false;
            cudaGetLastError();
        }
    *out = mg;
    return VXPT_OK;
}

int vxpt_mg_destroy(vxpt_mg_handle mg) {
    if (!mg) return VXPT_OK;
    int rc = VXPT_OK;
    for (vxpt_handle h : mg->dev) {
        const int r = vxpt_destroy(h);
        if (r && !rc) rc = r;
    }
    delete mg;
    return rc;
}

int vxpt_mg_size(vxpt_mg_handle mg) { return mg ? (int)mg->dev.size() : 0; }
vxpt_handle vxpt_mg_device(vxpt_mg_handle mg, int k) { return (mg && k >= 0 && k < (int)mg->dev.size()) ? mg->dev[k] : nullptr; }

// ---- replicated scene state: the same call on every device's handle
int vxpt_mg_upload_world(vxpt_mg_handle mg, const uint8_t* blocks) { MG_EACH(vxpt_upload_world(h, blocks)); }
int vxpt_mg_set_block(vxpt_mg_handle mg, int x, int y, int z, uint8_t id) { MG_EACH(vxpt_set_block(h, x, y, z, id)); }
int vxpt_mg_set_blocks(vxpt_mg_handle mg, const int16_t* xyz, const uint8_t* ids, int n) { MG_EACH(vxpt_set_blocks(h, xyz, ids, n)); }
int vxpt_mg_build_distance_field(vxpt_mg_handle mg) { MG_EACH(vxpt_build_distance_field(h)); }
int vxpt_mg_set_materials(vxpt_mg_handle mg, const int32_t table[768]) { MG_EACH(vxpt_set_materials(h, table)); }
int vxpt_mg_set_blue_noise(vxpt_mg_handle mg, const int32_t* sobol, const int32_t* scramble, const int32_t* rank) {
    MG_EACH(vxpt_set_blue_noise(h, sobol, scramble, rank));
}
int vxpt_mg_set_material_textures(vxpt_mg_handle mg, const float* albedo_lod3, const float* pbr_lod2, int n_layers, const float* emissive_lod0,
                                  int n_emissive_layers) {
    MG_EACH(vxpt_set_material_textures(h, albedo_lod3, pbr_lod2, n_layers, emissive_lod0, n_emissive_layers));
}
int vxpt_mg_set_reflection_textures(vxpt_mg_handle mg, const float* normal_lod3, int n_normal_layers, const float* emissive_lod2, int n_emissive_layers) {
    MG_EACH(vxpt_set_reflection_textures(h, normal_lod3, n_normal_layers, emissive_lod2, n_emissive_layers));
}
int vxpt_mg_set_sky_cubemap(vxpt_mg_handle mg, const float* rgb, int n) { MG_EACH(vxpt_set_sky_cubemap(h, rgb, n)); }
int vxpt_mg_set_shadow_noise(vxpt_mg_handle mg, const uint8_t* rgba8) { MG_EACH(vxpt_set_shadow_noise(h, rgba8)); }
int vxpt_mg_set_option(vxpt_mg_handle mg, int option, int value) { MG_EACH(vxpt_set_option(h, option, value)); }
int vxpt_mg_sync(vxpt_mg_handle mg) { MG_EACH(vxpt_sync(h)); }
int vxpt_mg_reset_stats(vxpt_mg_handle mg) { MG_EACH(vxpt_reset_stats(h)); }

int vxpt_mg_slab(vxpt_mg_handle mg, const VxCamera* cam, int k, int* row_begin, int* row_end) {
    if (!mg || !cam || !row_begin || !row_end || k < 0 || k >= (int)mg->dev.size()) return mg_fail(VXPT_E_INVALID, "bad argument");
    slab_rows(cam, k, (int)mg->dev.size(), row_begin, row_end);
    return VXPT_OK;
}

// ---- a frame: every device traces its row slab of `cam` into the caller's planes
int vxpt_mg_render_frame_async(vxpt_mg_handle mg, const VxCamera* cam, const VxFrameParams* p, const VxFrameOut* out) {
    if (!mg || !cam || !p || !out) return mg_fail(VXPT_E_INVALID, "bad argument");
    if (cam->interleave_n > 1) return mg_fail(VXPT_E_INVALID, "vxpt_mg_render_frame shards contiguous row slabs itself: interleave must be off");
    const int n = (int)mg->dev.size();
    if (n > 1 && !mg->peer_ok) {
        // without peer access only host planes can be written by every device
        const void* planes[] = {out->gbuffer.t, out->gbuffer.normal_id, out->gbuffer.block_id, out->gbuffer.inv_t, out->gbuffer.hit_voxel,
                                out->shadow.shadow, out->shadow.transversal, out->diffuse.sh, out->diffuse.cocg, out->diffuse.luma, out->diffuse.ao_sky,
                                out->reflection.color, out->reflection.hit_distance, out->reflection.emissive_mask,
                                out->material.albedo, out->material.normal, out->material.pbr, out->material.texture_ao, p->g_normal, p->g_pbr};
        for (const void* q : planes) {
            if (!q) continue;
            cudaPointerAttributes a;
            if (cudaPointerGetAttributes(&a, q) == cudaSuccess && (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged))
                return mg_fail(VXPT_E_STATE, "the devices of this vxpt_mg handle cannot reach each other's memory: pass host planes");
            cudaGetLastError();
        }
    }
    for (int k = 0; k < n; ++k) {
        VxCamera ck = *cam;
        slab_rows(cam, k, n, &ck.row_begin, &ck.row_end);
        if (ck.row_end <= ck.row_begin) continue;
        const int rc = vxpt_render_frame_async(mg->dev[k], &ck, p, out);
        if (rc) return rc;
    }
    return VXPT_OK;
}

// host planes complete, device planes written (every device's stream drained)
int vxpt_mg_frame_wait(vxpt_mg_handle mg) {
    if (!mg) return mg_fail(VXPT_E_INVALID, "multi-GPU handle is NULL");
    for (vxpt_handle h : mg->dev) {
        int rc = vxpt_frame_wait(h);
        if (!rc) rc = vxpt_sync(h);
        if (rc) return rc;
    }
    return VXPT_OK;
}

int vxpt_mg_render_frame(vxpt_mg_handle mg, const VxCamera* cam, const VxFrameParams* p, const VxFrameOut* out) {
    const int rc = vxpt_mg_render_frame_async(mg, cam, p, out);
    const int rc2 = vxpt_mg_frame_wait(mg);  // also after a failed enqueue: frames already in flight own the caller's planes until they land
    return rc ? rc : rc2;
}

int vxpt_mg_get_stats(vxpt_mg_handle mg, VxStats* out) {
    if (!mg || !out) return mg_fail(VXPT_E_INVALID, "bad argument");
    VxStats tot = {};
    for (vxpt_handle h : mg->dev) {
        VxStats s;
        const int rc = vxpt_get_stats(h, &s);
        if (rc) return rc;
        tot.rays += s.rays;
        tot.df_fetches += s.df_fetches;
        tot.vox_fetches += s.vox_fetches;
        tot.last_ms = std::max(tot.last_ms, s.last_ms);  // the devices run side by side: the frame takes as long as the slowest
        tot.df_build_ms = std::max(tot.df_build_ms, s.df_build_ms);
        tot.brick_pack_ms = std::max(tot.brick_pack_ms, s.brick_pack_ms);
    }
    *out = tot;
    return VXPT_OK;
}

}  // extern "C"
