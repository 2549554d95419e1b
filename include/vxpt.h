/*
 * vxpt.h — C ABI of the B200-native voxel ray-tracing core.
 *
 * Drop-in boundary for ONE path of swr06/VoxelPathTracer: the Manhattan distance-field build over the
 * 384x128x384 uint8 block grid and the distance-field-accelerated voxel DDA traversal behind primary,
 * sun-shadow, diffuse-GI and reflection rays.  The reference has no FFI surface; its de-facto interface is
 * the set of GL objects + uniforms Core/Pipeline.cpp binds before each trace draw.  Every export below
 * names the reference call site it replaces (paths relative to the reference tree).
 *
 * Conventions
 *  - plain C, no exceptions cross the boundary; every call returns VXPT_OK (0) or a negative VXPT_E_*;
 *    vxpt_last_error() returns a thread-local description of the last failure.
 *  - a handle is NOT thread-safe (one caller thread per handle, like a GL context).
 *  - all buffers are caller-owned.  Input pointers may be host or device memory; output pointers may be host
 *    or device memory (detected with cudaPointerGetAttributes).  Host planes are staged through device memory owned
 *    by the handle and are complete when the call returns (pinned host buffers make the copies DMA-direct); device
 *    outputs are complete after vxpt_sync() (work is enqueued on the handle's private stream).
 *  - images: pixel (i, j) with j = 0 the BOTTOM row (GL convention, a_TexCoords of the full-screen quad);
 *    plane index = j * width + i.  A call renders rows [row_begin, row_end) only and touches no other row of
 *    the output planes — this is the multi-GPU row-slab contract.
 *  - matrices are column-major float[16] exactly as glm::value_ptr gives them.
 *  - there is NO CPU fallback: if no CUDA device is usable vxpt_create fails with VXPT_E_CUDA.
 */
#ifndef VXPT_H_
#define VXPT_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define VXPT_API __declspec(dllexport)
#else
#define VXPT_API __attribute__((visibility("default")))
#endif

/* Core/Macros.h:3-5 */
#define VXPT_WORLD_SIZE_X 384
#define VXPT_WORLD_SIZE_Y 128
#define VXPT_WORLD_SIZE_Z 384
#define VXPT_WORLD_VOXELS (VXPT_WORLD_SIZE_X * VXPT_WORLD_SIZE_Y * VXPT_WORLD_SIZE_Z) /* 18,874,368 */

/* normal ids, Core/Shaders/InitialRayTraceFrag.glsl:143-185 : {+Z,-Z,+Y,-Y,-X,+X} -> 0..5.
 * The reference writes id/10 into an R8 target and 1.0 on a miss, which every consumer decodes with
 * int(round(n*10)) -> 10 ("idx > 5" = no surface). */
#define VXPT_NORMAL_MISS 10

#define VXPT_OK 0
#define VXPT_E_INVALID (-1)     /* bad argument / bad state of arguments            */
#define VXPT_E_CUDA (-2)        /* CUDA runtime failure (message in last_error)     */
#define VXPT_E_NOMEM (-3)       /* allocation failed                                */
#define VXPT_E_STATE (-4)       /* call order violated (e.g. trace before DF build) */
#define VXPT_E_UNSUPPORTED (-5) /* reference option outside the v1 parity profile   */

typedef struct vxpt_ctx* vxpt_handle;

/* ---- camera / frame description --------------------------------------------------------------------------
 * replaces the uniforms u_InverseView, u_InverseProjection, u_Dimensions (Core/Pipeline.cpp:1979-1999,
 * 2203-2206, 2807-2810) and FBOVert.glsl's u_VertInverseView / u_VertInverseProjection. */
typedef struct VxCamera {
    float inv_view[16];
    float inv_proj[16];
    int32_t width, height;       /* full frame size in pixels                         */
    int32_t row_begin, row_end;  /* slab [row_begin,row_end); 0,height = whole frame  */
    /* Interleaved row bands for multi-GPU load balance (0 or 1 = off).  With interleave_n = N > 1 this handle renders the
     * bands b with b % N == interleave_rank, a band being band_rows consecutive image rows (height % (N*band_rows) == 0).
     * The handle's rows are numbered densely as VIRTUAL rows v in [0, height/N): image row j = ((v / band_rows) * N +
     * interleave_rank) * band_rows + v % band_rows.  row_begin/row_end then select virtual rows, and every plane handed to
     * the call is RANK-LOCAL: height/N rows, indexed v * width + i.  All per-pixel arithmetic uses the image row j. */
    int32_t interleave_n, interleave_rank, band_rows, reserved;
} VxCamera;

/* ---- primary rays: Core/Pipeline.cpp:1973-2016 -> InitialRayTraceFrag.glsl:417-468 ---------------------- */
typedef struct VxPrimaryParams {
    int32_t max_iterations;  /* u_RenderDistance: 475 on frame 0 then 350 (Pipeline.cpp:55,4824) */
    int32_t jitter_enable;   /* u_JitterSceneForTAA                                             */
    float jitter[2];         /* u_CurrentTAAJitter = Halton(2,3)[frame % 64] (TAAJitter.cpp)    */
    int32_t alpha_test;      /* u_ShouldAlphaTest (off by default, Pipeline.cpp:141): VoxelTraversalDF_AlphaTest,   */
                             /* InitialRayTraceFrag.glsl:189-305; needs vxpt_set_albedo_alpha_mips                  */
    float fov_degrees;       /* u_FOV (read by the alpha test's LOD only: g_K, InitialRayTraceFrag.glsl:421)        */
} VxPrimaryParams;

/* G-buffer planes (InitialTraceFBO attachments, Pipeline.cpp:1094-1095).  t is kept in fp32 (the reference's
 * R16F cannot meet the 1e-5 parity target).  Any pointer may be NULL to skip that plane.
 * As an OUTPUT of vxpt_trace_primary and as the INPUT of the secondary passes. */
typedef struct VxGBuffer {
    float* t;            /* o_HitDistance : hit distance, -1 on a miss                        */
    uint8_t* normal_id;  /* o_Normal      : 0..5, VXPT_NORMAL_MISS on a miss                  */
    uint8_t* block_id;   /* o_BlockID     : block byte, 0 on a miss                           */
    float* inv_t;        /* o_DepthNonLinear = 1/t                                            */
    int16_t* hit_voxel;  /* optional parity plane: 3 x int16 voxel coords, (-1,-1,-1) on miss */
} VxGBuffer;

/* ---- sun shadow: Core/Pipeline.cpp:2795-2852 -> ShadowRayTraceFrag.glsl:414-513 ------------------------- */
typedef struct VxShadowParams {
    float light_dir[3];      /* u_LightDirection (normalised StrongerLightDirection)              */
    int32_t frame;           /* u_CurrentFrame (blue-noise texel offset, frame % 1024)            */
    int32_t soft;            /* u_ContactHardeningShadows (cone jitter); default 1                */
    float halton[2];         /* u_Halton; zero unless supersampling (Pipeline.cpp:2823)           */
    int32_t alpha_test;      /* u_ShouldAlphaTest = ShouldAlphaTestShadows (Pipeline.cpp:2820):   */
                             /* VoxelTraversalDF_AlphaTest, ShadowRayTraceFrag.glsl:105-220       */
    float fov_degrees;       /* u_FOV (alpha test's LOD only, ShadowRayTraceFrag.glsl:419)        */
} VxShadowParams;

typedef struct VxShadowOut {
    uint8_t* shadow;     /* o_Shadow 0/1                                             */
    float* transversal;  /* o_IntersectionTransversal                                */
} VxShadowOut;

/* ---- diffuse GI: Core/Pipeline.cpp:2174-2281 -> DiffuseRayTraceFrag.glsl:822-935 ------------------------ */
typedef struct VxDiffuseParams {
    int32_t spp;             /* u_SPP (clamped 1..32)                                             */
    int32_t checker_spp;     /* u_CheckerSPP                                                      */
    int32_t checkerboard;    /* CHECKERBOARD_SPP                                                  */
    int32_t trace_length;    /* u_DiffuseTraceLength, default 48 (Pipeline.cpp:75)                */
    int32_t frame;           /* u_CurrentFrame; blue-noise index = frame % 128                    */
    int32_t use_blue_noise;  /* u_UseBlueNoise; must be 1 (hash2 path is not reproducible)        */
    int32_t supersample;     /* u_Supersample                                                     */
    int32_t direct_sampling; /* u_UseDirectSampling; must be 0                                    */
    float halton[2];         /* u_Halton                                                          */
    float sun_dir[3];        /* u_SunDirection                                                    */
    float moon_dir[3];       /* u_MoonDirection                                                   */
    float sun_visibility;    /* u_SunVisibility                                                   */
    float gi_sun_strength;   /* u_GISunStrength   (1.0)                                           */
    float gi_sky_strength;   /* u_GISkyStrength   (1.125)                                         */
    float light_intensity;   /* u_DiffuseLightIntensity (1.25)                                    */
} VxDiffuseParams;

typedef struct VxDiffuseOut {
    float* sh;      /* o_SH   4 floats / pixel   (RGBA16F in the reference)   */
    float* cocg;    /* o_CoCg 2 floats / pixel                                */
    float* luma;    /* o_Utility 1 float / pixel                              */
    float* ao_sky;  /* o_AOAndSkyLighting 2 floats / pixel                    */
} VxDiffuseOut;

/* ---- reflections: Core/Pipeline.cpp:3003-3164 -> ReflectionTraceFrag.glsl:717-1038 ----------------------
 * Parity profile v1: u_ReprojectToScreenSpace, u_LPVGI, u_CloudReflections, u_ReflectPlayer,
 * u_DeriveFromDiffuseSH = false; lava UV distortion / flicker (functions of wall-clock u_Time) never apply. */
typedef struct VxReflectionParams {
    int32_t spp;              /* u_SPP (clamped 1..16)                                            */
    int32_t trace_length;     /* u_ReflectionTraceLength, default 64                              */
    int32_t frame;            /* u_CurrentFrame; blue-noise index = frame % 128 (TEMPORAL_SPEC);  */
                              /* frame < 0 selects the TEMPORAL_SPEC = false index 100            */
    int32_t rough;            /* u_RoughReflections                                               */
    int32_t roughness_bias;   /* u_RoughnessBias (0.85x)                                          */
    int32_t checkerboard;     /* CHECKERBOARD_SPEC_SPP                                            */
    float sun_dir[3];         /* u_SunDirection                                                   */
    float moon_dir[3];        /* u_MoonDirection                                                  */
    float stronger_dir[3];    /* u_StrongerLightDirection                                         */
    float viewer_pos[3];      /* u_ViewerPosition                                                 */
    float sun_strength;       /* u_SunStrengthModifier 0.85                                       */
    float moon_strength;      /* u_MoonStrengthModifier 1.0                                       */
    float halton[2];          /* u_Halton = GetTAAJitterSecondary(frame) (Pipeline.cpp:3032), clamped to [-2, 2]: the pass reads the
                                 primary distance (GL_LINEAR) and normal id (GL_NEAREST, both GL_REPEAT) and sets up its ray at
                                 uv + halton / dims (ReflectionTraceFrag.glsl:754-777).  G-buffer rows up to ceil(|halton[1]|) + 1
                                 beyond [row_begin, row_end) must therefore be valid in VxGBuffer.t / normal_id (wrapping at the frame
                                 edge); vxpt_render_frame traces those halo rows itself.  Contiguous row slabs only (no interleave). */
    int32_t grass_props[10];  /* u_GrassBlockProps (Pipeline.cpp:3040-3049)                       */
} VxReflectionParams;

typedef struct VxReflectionIn {
    const float* g_normal;   /* u_GBufferNormals 3 floats / pixel; NULL = the face normal                        */
    const float* g_pbr;      /* u_GBufferPBR 4 floats / pixel (roughness, metalness, -, emissivity); NULL = level-2 PBR
                                texel of the block at the primary hit (needs VxGBuffer.block_id), emissivity 0  */
    const float* sh;         /* u_DiffuseSH   4 floats / pixel (this library's GI output)                        */
    const float* cocg;       /* u_DiffuseCoCg 2 floats / pixel                                                   */
} VxReflectionIn;

typedef struct VxReflectionOut {
    float* color;            /* o_Color 4 floats / pixel       */
    float* hit_distance;     /* o_HitDistance                  */
    uint8_t* emissive_mask;  /* o_EmissivityHitMask 0/1        */
} VxReflectionOut;

/* ---- traversal statistics (the reference has none; they define the roofline's algorithmic bytes) -------- */
typedef struct VxStats {
    uint64_t rays;        /* VoxelTraversalDF calls                                   */
    uint64_t df_fetches;  /* loop iterations that read the distance field (N_it)      */
    uint64_t vox_fetches; /* block-id fetches (N_vox)                                 */
    float last_ms;        /* device time of the last pass (CUDA events)               */
    float df_build_ms;    /* device time of the last distance-field build (linear field) */
    float brick_pack_ms;  /* device time of the brick re-layout that follows it           */
} VxStats;

/* ---- lifetime: new World() + World::InitializeDistanceGenerator (Core/World.cpp:48-67) ------------------ */
VXPT_API int vxpt_create(int device_id, vxpt_handle* out);
VXPT_API int vxpt_destroy(vxpt_handle h);
VXPT_API const char* vxpt_last_error(void);
VXPT_API const char* vxpt_version(void);

/* ---- world: World::Buffer (Core/World.h:167-171), glTexSubImage3D edits (Core/World.cpp:372-373,458-459) - */
VXPT_API int vxpt_upload_world(vxpt_handle h, const uint8_t* blocks /* VXPT_WORLD_VOXELS, x + 384*y + 49152*z */);
VXPT_API int vxpt_set_block(vxpt_handle h, int x, int y, int z, uint8_t id);
VXPT_API int vxpt_set_blocks(vxpt_handle h, const int16_t* xyz /* 3*n */, const uint8_t* ids, int n);
VXPT_API int vxpt_download_world(vxpt_handle h, uint8_t* blocks);

/* ---- distance field: World::GenerateDistanceField (Core/World.cpp:69-113) ------------------------------- */
VXPT_API int vxpt_build_distance_field(vxpt_handle h);
VXPT_API int vxpt_download_distance_field(vxpt_handle h, uint8_t* out /* VXPT_WORLD_VOXELS */);
/* raw device pointers of the resident grid / linear distance field (for zero-copy consumers and tests) */
VXPT_API int vxpt_device_pointers(vxpt_handle h, const uint8_t** grid, const uint8_t** df);

/* ---- tables: BlockDataSSBO::CreateBuffers (Core/BlockDataSSBO.cpp:28-39), BlueNoiseDataSSBO ctor
 *      (Core/BlueNoiseDataSSBO.cpp:19-30), texture binds (Core/Pipeline.cpp:2236-2270, 2825-2841) --------- */
VXPT_API int vxpt_set_materials(vxpt_handle h, const int32_t table[768]);
VXPT_API int vxpt_set_blue_noise(vxpt_handle h, const int32_t* sobol /*65536*/, const int32_t* scramble /*131072*/,
                                 const int32_t* rank /*131072*/);
/* pre-baked texel arrays so both sides of a parity check sample identical data:
 *  albedo_lod3  [n_layers][64][64][4]   f32 linear RGBA, mip level 3 of the 512^2 sRGB albedo array
 *  pbr_lod2     [n_layers][128][128][4] f32, mip level 2 of the PBR array
 *  emissive_lod0[n_emissive_layers][512][512] f32 (red channel)                                    */
VXPT_API int vxpt_set_material_textures(vxpt_handle h, const float* albedo_lod3, const float* pbr_lod2, int n_layers,
                                        const float* emissive_lod0, int n_emissive_layers);
/* extra texel arrays the reflection pass samples (ReflectionTraceFrag.glsl:961, :973):
 *  normal_lod3  [n_normal_layers][64][64][4]   f32, mip level 3 of the normal-map array (rgb in 0..1)
 *  emissive_lod2[n_emissive_layers][128][128]  f32, mip level 2 of the emissive array (red channel)
 * The reflection pass reads albedo at the baked level 3 and PBR (rgba) at the baked level 2 of
 * vxpt_set_material_textures (the shader's implicit-derivative texture() calls sit in divergent control flow, where GL
 * leaves the level undefined). */
VXPT_API int vxpt_set_reflection_textures(vxpt_handle h, const float* normal_lod3, int n_normal_layers, const float* emissive_lod2,
                                          int n_emissive_layers);
/* alpha channel of the albedo array's whole mip chain, for the alpha-tested traversal (StopRay, InitialRayTraceFrag.glsl:189-203,
 * ShadowRayTraceFrag.glsl:105-117: textureLod(u_AlbedoTextures, uv, LOD).w > 0.975 with a per-hit integer LOD 0..8).  The array
 * is GL_SRGB_ALPHA / GL_NEAREST_MIPMAP_LINEAR (Core/GLClasses/TextureArray.cpp:28-42), so alpha is a linear unorm8 at every
 * level and an integer LOD reads the nearest texel of exactly that level:
 *  alpha_mips [n_layers][VXPT_ALPHA_MIP_TEXELS] uint8: per layer, levels 0..8 back to back, level k = [512 >> k][512 >> k] */
#define VXPT_ALPHA_MIP_TEXELS 349524 /* 512^2 + 256^2 + ... + 2^2 */
VXPT_API int vxpt_set_albedo_alpha_mips(vxpt_handle h, const uint8_t* alpha_mips, int n_layers);
VXPT_API int vxpt_set_sky_cubemap(vxpt_handle h, const float* rgb /* [6][n][n][3], faces +X,-X,+Y,-Y,+Z,-Z */, int n);
VXPT_API int vxpt_set_shadow_noise(vxpt_handle h, const uint8_t* rgba8 /* [256][256][4] */);

/* ---- trace passes ---------------------------------------------------------------------------------------- */
VXPT_API int vxpt_trace_primary(vxpt_handle h, const VxCamera* cam, const VxPrimaryParams* p, const VxGBuffer* out);
VXPT_API int vxpt_trace_shadow(vxpt_handle h, const VxCamera* cam, const VxGBuffer* gbuf, const VxShadowParams* p,
                               const VxShadowOut* out);
VXPT_API int vxpt_trace_diffuse(vxpt_handle h, const VxCamera* cam, const VxGBuffer* gbuf, const VxDiffuseParams* p,
                                const VxDiffuseOut* out);
VXPT_API int vxpt_trace_reflection(vxpt_handle h, const VxCamera* cam, const VxGBuffer* gbuf,
                                   const VxReflectionIn* in, const VxReflectionParams* p, const VxReflectionOut* out);

/* ---- the other consumers of the distance field -----------------------------------------------------------
 * vxpt_trace_rays: a batch of n VoxelTraversalDF calls (InitialRayTraceFrag.glsl:307-374) on caller-supplied rays; directions are
 * used as given (the shaders pass normalised ones).  origins / directions: 3 n floats; outputs (each may be NULL, host or device):
 * t[n] (-1 on a miss), normal_id[n] (0..5, VXPT_NORMAL_MISS), block_id[n], hit_voxel[3 n].
 * vxpt_player_shadowed: Core/Shaders/PostProcessingVert.glsl:46-53 (u_ComputePlayerShadow): one ray from the camera along
 * u_VertSunDir / length(u_VertSunDir), cap 350; *shadowed = v_PlayerShadowed = (T > 0).
 * vxpt_estimate_ambient_sound: Core/Shaders/EstimateAmbientSoundLevel.comp as dispatched by Core/Pipeline.cpp:1908-1921
 * (32 invocations): *sky_level_aggregate = SkyLevelAggregate (0..32*512); per_invocation[32] (optional) = each invocation's
 * addend, index = gl_GlobalInvocationID.y * 8 + .x.  The host folds the aggregate into the sound volume (Pipeline.cpp:1937-1946). */
VXPT_API int vxpt_trace_rays(vxpt_handle h, const float* origins, const float* directions, int n, int max_iterations, float* t,
                             uint8_t* normal_id, uint8_t* block_id, int16_t* hit_voxel);
VXPT_API int vxpt_player_shadowed(vxpt_handle h, const float camera_pos[3], const float sun_dir[3], int* shadowed);
VXPT_API int vxpt_estimate_ambient_sound(vxpt_handle h, const float player_pos[3], int frame, uint32_t* sky_level_aggregate,
                                         uint32_t* per_invocation /* 32 or NULL */);

/* ---- G-buffer material pass: Core/Pipeline.cpp:2066-2136 -> Core/Shaders/GenerateGBuffer.glsl:347-549 --------
 * The immediate consumer of the primary hits: per pixel, the block's albedo / normal-map / PBR / emissive texels at the hit point,
 * written to the GeneratedGBuffer attachments (Pipeline.cpp:1100) whose normal and PBR planes the reflection pass reads as
 * u_GBufferNormals / u_GBufferPBR (VxReflectionIn.g_normal / g_pbr).
 * Textures: the block arrays as the GL texture objects hold them (Core/GLClasses/TextureArray.cpp:28-78): RGBA8 texels of a
 * 512^2 layer with its complete mip chain, levels 0..9 back to back (VXPT_MIP_CHAIN_TEXELS texels per layer, level k at offset
 * sum_{m<k} (512>>m)^2), [n_layers][VXPT_MIP_CHAIN_TEXELS][4] bytes.  The albedo array is GL_SRGB_ALPHA (rgb decoded to linear on
 * fetch), normal and PBR arrays are GL_RGBA.  The emissive array is the level-0 red channel vxpt_set_material_textures received.
 * Pinned sampling (GL leaves filter arithmetic to the driver): textureGrad = the isotropic OpenGL 4.3 section 8.14
 * scale factor rho = max(|dP/dx| , |dP/dy|) * 512, lambda = log2(rho); lambda <= c magnifies on level 0 (albedo: GL_NEAREST, c = 0;
 * normal / PBR: GL_LINEAR, c = 0.5); otherwise GL_NEAREST_MIPMAP_LINEAR: nearest texel of levels floor(lambda) and floor(lambda)+1
 * (clamped to 9), blended by fract(lambda).  dFdx / dFdy are differences inside the pixel's 2x2 quad (fine derivatives); a quad
 * neighbour that shades nothing (sky, outside the frame) contributes the pixel's own value.  The anisotropy extension is not modelled.
 * The height-field taps of the parallax march (texture() with implicit derivatives inside a loop: undefined LOD in GL) are pinned to
 * level 0, GL_LINEAR, like the emissive fetch.  The animated lava textures (Core/AnimatedTexture.cpp: GL_RGBA8 3-D textures of 256^2 x 8
 * frames, GL_LINEAR, GL_REPEAT on all three axes) are trilinear per OpenGL 4.3 section 8.14.2 in fp32, weights a(1-f) + bf, x then y then z;
 * sin / cos / pow of the UV distortion are the correctly rounded fp32 values. */
#define VXPT_MIP_CHAIN_TEXELS 349525 /* 512^2 + 256^2 + ... + 1 */
typedef struct VxMaterialParams {
    int32_t update_this_frame;  /* u_UpdateGBufferThisFrame: 0 = every pixel is discarded (planes untouched) unless pom is set  */
    int32_t pom;                /* u_POM (off by default, Pipeline.cpp:264): relief parallax mapping, ReliefParallax :153-199    */
    int32_t lava_block_id;      /* u_LavaBlockID: pixels of this block take the animated lava path (needs vxpt_set_lava_textures);  */
                                /* negative = no block does                                                                  */
    int32_t grass_props[10];    /* u_GrassBlockProps (Pipeline.cpp:2083-2092): block id, then albedo / normal / PBR layers of  */
                                /* the top, side and bottom faces                                                            */
    float pom_height;           /* u_POMHeight (1.0)                                                                         */
    float pom_exp;              /* u_POMExp (1.0)                                                                            */
    int32_t high_quality_pom;   /* u_HighQualityPOM (0): 64..128 march steps instead of 32..64                               */
    int32_t dither_pom;         /* u_DitherPOM (1): per-pixel step count from a Bayer pattern and u_Frame                     */
    int32_t frame;              /* u_Frame                                                                                   */
    float time;                 /* u_Time = glfwGetTime(): drives the lava UV distortion and frame blend (:127-137, :390)      */
} VxMaterialParams;
typedef struct VxMaterialOut {  /* fp32 planes in every texel format (the reflection pass reads normal / pbr as fp32) */
    float* albedo;      /* o_Albedo    3 floats / pixel (RGB16F in the reference)                             */
    float* normal;      /* o_Normal    3 floats / pixel: the normal-mapped shading normal, (1,1,1) on a miss  */
    float* pbr;         /* o_PBR       4 floats / pixel: roughness, metalness, displacement, emissivity       */
    float* texture_ao;  /* o_TextureAO 1 float / pixel                                                        */
} VxMaterialOut;
#define VXPT_LAVA_SIZE 256  /* Core/Pipeline.cpp:1475-1477: LavaAlbedo / LavaNormals .Create(path, 256, 7) -> 256 x 256 x 8 texels */
#define VXPT_LAVA_FRAMES 8
/* u_LavaTextures[0] (albedo) and [1] (normals): RGBA8 [VXPT_LAVA_FRAMES][VXPT_LAVA_SIZE][VXPT_LAVA_SIZE][4] each */
VXPT_API int vxpt_set_lava_textures(vxpt_handle h, const uint8_t* albedo_rgba8, const uint8_t* normal_rgba8);
VXPT_API int vxpt_set_gbuffer_textures(vxpt_handle h, const uint8_t* albedo_mips, const uint8_t* normal_mips, const uint8_t* pbr_mips,
                                       int n_layers);
/* reads VxGBuffer.inv_t (u_NonLinearDepth), normal_id and block_id; row_begin / row_end must be even (or the frame height): a 2x2
 * quad is shaded by one call.  Planes are device or host pointers like those of the trace passes. */
VXPT_API int vxpt_generate_gbuffer(vxpt_handle h, const VxCamera* cam, const VxGBuffer* gbuf, const VxMaterialParams* p,
                                   const VxMaterialOut* out);

/* ---- SVGF diffuse denoiser: Core/Pipeline.cpp:2284-2596 -> Core/Shaders/SVGF/{TemporalFilter,VarianceEstimate,
 *      SpatialFilter}.glsl ----------------------------------------------------------------------------------------------------------
 * The step right after the diffuse-GI pass: temporal accumulation against the previous frame (reprojected), a variance estimate, and
 * five edge-stopping a-trous passes (u_Step 16, 8, 4, 2, 1; ping-pong between two plane sets, Pipeline.cpp:2468-2596).  One export per
 * dispatch; all planes are fp32 (VXPT_OPT_TEXEL_FORMAT must be 0) and FULL-FRAME: a stencil reads rows outside [row_begin, row_end),
 * which select the OUTPUT rows only; interleaved row bands are rejected (VXPT_E_UNSUPPORTED).  Outputs must not alias inputs.
 * Pinned sampling (GL leaves filter arithmetic open): every read is texture() on an FBO attachment with GL_REPEAT and the filter
 * Pipeline.cpp:1094-1152 declares — hit distance GL_LINEAR, normal / block id GL_NEAREST, all radiance / utility / AO / variance planes
 * GL_LINEAR; linear = OpenGL 4.3 section 8.14.2 in fp32, weights applied as a(1-f) + bf, x first.  min / max / clamp return the non-NaN
 * operand (IEEE minNum / maxNum, what the GPUs the reference runs on do); exp / pow are correctly rounded fp32. */
/* pre-temporal 3x3 pass (PreTemporalSpatialPass = true by default, Pipeline.cpp:239, 2288-2330 -> Spatial3x3Initial.glsl): filters this
 * frame's raw GI planes before they enter the temporal pass; its outputs then stand in for VxSvgfTemporalIn.sh / cocg / luma / ao_sky. */
typedef struct VxSvgfInitialIn {
    VxGBuffer current;     /* t, normal_id                                         */
    const float* sh;       /* u_SH      VxDiffuseOut.sh                            */
    const float* cocg;     /* u_CoCg    VxDiffuseOut.cocg                          */
    const float* luma;     /* u_Utility VxDiffuseOut.luma (passed through)         */
    const float* ao_sky;   /* u_AO      VxDiffuseOut.ao_sky                        */
} VxSvgfInitialIn;
typedef struct VxSvgfInitialOut {
    float* sh;       /* o_SH 4       */
    float* cocg;     /* o_CoCg 2     */
    float* luma;     /* o_Utility 1  */
    float* ao_sky;   /* o_AOSky 2    */
} VxSvgfInitialOut;
VXPT_API int vxpt_svgf_initial(vxpt_handle h, const VxCamera* cam, const VxSvgfInitialIn* in, const VxSvgfInitialOut* out);

typedef struct VxSvgfTemporalIn {
    VxGBuffer current;          /* t, normal_id, block_id of this frame (InitialTraceFBO attachments 0..2)                       */
    VxGBuffer previous;         /* the same planes of the previous frame (InitialTraceFBOPrev)                                    */
    const float* sh;            /* u_CurrentSH   4 floats / pixel \                                                              */
    const float* cocg;          /* u_CurrentCoCg 2 floats / pixel  | this frame's VxDiffuseOut (DiffuseRawTraceFBO 0..3)          */
    const float* luma;          /* u_NoisyLuminosity 1 float / pixel |                                                            */
    const float* ao_sky;        /* u_CurrentAO   2 floats / pixel /                                                               */
    const float* prev_sh;       /* u_PreviousSH      \                                                                           */
    const float* prev_cocg;     /* u_PrevCoCg         | the previous frame's VxSvgfTemporalOut (PrevDiffuseTemporalFBO 0..3)      */
    const float* prev_utility;  /* u_PreviousUtility  | 3 floats / pixel                                                          */
    const float* prev_ao_sky;   /* u_PreviousAO      /                                                                            */
} VxSvgfTemporalIn;
typedef struct VxSvgfTemporalParams {
    float prev_view[16];        /* u_PrevView                                                   */
    float prev_projection[16];  /* u_PrevProjection                                             */
    int32_t be_useful;          /* u_BeUseful = DO_SVGF_TEMPORAL (1)                            */
} VxSvgfTemporalParams;
typedef struct VxSvgfTemporalOut {
    float* sh;       /* o_SH 4 floats / pixel                                                              */
    float* cocg;     /* o_CoCg 2                                                                           */
    float* utility;  /* o_Utility 3: accumulated frames, second moment of the luminance, luminance         */
    float* ao_sky;   /* o_AOAndSkyLighting 2                                                               */
} VxSvgfTemporalOut;
VXPT_API int vxpt_svgf_temporal(vxpt_handle h, const VxCamera* cam, const VxSvgfTemporalIn* in, const VxSvgfTemporalParams* p,
                                const VxSvgfTemporalOut* out);

typedef struct VxSvgfVarianceIn {
    VxGBuffer current;     /* t, normal_id                                      */
    const float* sh;       /* VxSvgfTemporalOut.sh                              */
    const float* cocg;     /* VxSvgfTemporalOut.cocg                            */
    const float* utility;  /* VxSvgfTemporalOut.utility                         */
} VxSvgfVarianceIn;
typedef struct VxSvgfVarianceParams {
    int32_t do_spatial;               /* DO_SPATIAL = DO_VARIANCE_SPATIAL (1)   */
    int32_t aggressive_disocclusion;  /* AGGRESSIVE_DISOCCLUSION_HANDLING (1)   */
} VxSvgfVarianceParams;
typedef struct VxSvgfVarianceOut {
    float* sh;        /* o_SH 4       */
    float* cocg;      /* o_CoCg 2     */
    float* variance;  /* o_Variance 1 */
} VxSvgfVarianceOut;
VXPT_API int vxpt_svgf_variance(vxpt_handle h, const VxCamera* cam, const VxSvgfVarianceIn* in, const VxSvgfVarianceParams* p,
                                const VxSvgfVarianceOut* out);

typedef struct VxSvgfSpatialIn {
    VxGBuffer current;              /* t, normal_id                                                                       */
    const float* sh;                /* u_SH   : the previous pass's sh (the variance pass's for the first)                */
    const float* cocg;              /* u_CoCg                                                                             */
    const float* variance;          /* u_VarianceTexture                                                                  */
    const float* ao_sky;            /* u_AO   : VxSvgfTemporalOut.ao_sky for the first pass, then the previous pass's     */
    const float* temporal_utility;  /* u_TemporalMoment: VxSvgfTemporalOut.utility (.x = accumulated frames)              */
} VxSvgfSpatialIn;
typedef struct VxSvgfSpatialParams {
    int32_t step;                     /* u_Step: 16, 8, 4, 2, 1 (32 .. 2 with WiderSVGF)                 */
    int32_t large_kernel;             /* u_LargeKernel = SVGF_LARGE_KERNEL (0): 5x5 instead of 3x3 taps  */
    int32_t do_spatial;               /* DO_SPATIAL = DO_SVGF_SPATIAL (1)                                */
    int32_t aggressive_disocclusion;  /* AGGRESSIVE_DISOCCLUSION_HANDLING (1)                            */
    float color_phi_bias;             /* u_ColorPhiBias = ColorPhiBias (3.325, Pipeline.cpp:85)          */
    float time;                       /* u_Time = glfwGetTime(): seeds the per-pixel tap jitter          */
    float resolution_scale;           /* u_ResolutionScale = DiffuseIndirectSuperSampleRes (0.25, :78)   */
} VxSvgfSpatialParams;
typedef struct VxSvgfSpatialOut {
    float* sh;        /* o_SH 4                 */
    float* cocg;      /* o_CoCg 2               */
    float* variance;  /* o_Variance 1           */
    float* ao_sky;    /* o_AOAndSkylighting 2   */
} VxSvgfSpatialOut;
VXPT_API int vxpt_svgf_spatial(vxpt_handle h, const VxCamera* cam, const VxSvgfSpatialIn* in, const VxSvgfSpatialParams* p,
                               const VxSvgfSpatialOut* out);

/* The whole SVGF chain of one frame in one call (Core/Pipeline.cpp:2284-2596): pre-temporal pass -> temporal -> variance -> five a-trous
 * passes, with everything between the passes — and the history the next frame needs (this frame's temporal planes, G-buffer and camera
 * matrices) — resident in device memory owned by the handle (230 B/pixel, allocated on first use).  Equivalent to the separate exports
 * above called in that order; only the GI planes come in and the denoised planes go out.  Whole frames only (row_begin = 0, row_end =
 * height); a change of resolution or reset_history = 1 starts a new history (previous planes zero, previous camera = this camera). */
typedef struct VxSvgfFrameParams {
    float view[16];                   /* u_View of this frame; u_PrevView of the next call                      */
    float projection[16];             /* u_Projection of this frame; u_PrevProjection of the next call          */
    int32_t reset_history;            /* 1 = forget the previous frame (first frame, camera cut)                */
    int32_t pre_pass;                 /* PreTemporalSpatialPass (1, Pipeline.cpp:239)                           */
    int32_t wide;                     /* WiderSVGF (0): a-trous steps 32, 16, 8, 4, 2 instead of 16, 8, 4, 2, 1 */
    int32_t large_kernel;             /* SVGF_LARGE_KERNEL (0)                                                  */
    int32_t aggressive_disocclusion;  /* AGGRESSIVE_DISOCCLUSION_HANDLING (1)                                   */
    float color_phi_bias;             /* ColorPhiBias (3.325)                                                   */
    float time;                       /* u_Time                                                                 */
    float resolution_scale;           /* DiffuseIndirectSuperSampleRes (0.25)                                   */
} VxSvgfFrameParams;
VXPT_API int vxpt_svgf_frame(vxpt_handle h, const VxCamera* cam, const VxGBuffer* gbuf /* t, normal_id, block_id */,
                             const VxDiffuseOut* diffuse /* read: this frame's GI planes */, const VxSvgfFrameParams* p,
                             const VxSvgfSpatialOut* out);

/* ---- sun-shadow filters: Core/Pipeline.cpp:2854-2944 -> Core/Shaders/ShadowTemporalFilter.glsl (u_ShadowTemporal =
 *      true), ShadowFilter.glsl ------------------------------------------------------------------------------------------------------
 * The two passes after vxpt_trace_shadow: temporal accumulation of the 0 / 1 shadow plane (history clipped to the neighbourhood of the
 * current frame, frame counter in a second plane) and an edge-stopping spatial filter whose footprint grows with the occluder distance
 * ("transversal").  Same conventions as the SVGF exports: fp32 full-frame planes, pinned texture() sampling, plain row slabs. */
typedef struct VxShadowTemporalIn {
    VxGBuffer current;          /* t, normal_id of this frame                                                         */
    VxGBuffer previous;         /* t of the previous frame (u_PreviousFramePositionTexture)                           */
    const uint8_t* shadow;      /* u_CurrentColorTexture: VxShadowOut.shadow (0 / 1, the R8 attachment)               */
    const float* transversal;   /* u_ShadowTransversals: VxShadowOut.transversal                                      */
    const float* prev_shadow;   /* u_PreviousColorTexture: the previous frame's VxShadowTemporalOut.shadow            */
    const float* prev_frames;   /* u_FrameCount: the previous frame's VxShadowTemporalOut.frames                      */
} VxShadowTemporalIn;
typedef struct VxShadowTemporalParams {
    float prev_view[16];        /* u_PrevView        */
    float prev_projection[16];  /* u_PrevProjection  */
} VxShadowTemporalParams;
typedef struct VxShadowTemporalOut {
    float* shadow;  /* o_Color.x (R8 in the reference)   */
    float* frames;  /* o_Frames  (R16F in the reference) */
} VxShadowTemporalOut;
VXPT_API int vxpt_shadow_temporal(vxpt_handle h, const VxCamera* cam, const VxShadowTemporalIn* in, const VxShadowTemporalParams* p,
                                  const VxShadowTemporalOut* out);
typedef struct VxShadowFilterIn {
    VxGBuffer current;         /* t, normal_id                                             */
    const float* shadow;       /* u_InputTexture: VxShadowTemporalOut.shadow               */
    const float* transversal;  /* u_IntersectionTransversals: VxShadowOut.transversal      */
    const float* frames;       /* u_FrameCount: VxShadowTemporalOut.frames                 */
} VxShadowFilterIn;
typedef struct VxShadowFilterParams {
    float filter_scale;        /* u_ShadowFilterScale (1.0, Pipeline.cpp:132)              */
} VxShadowFilterParams;
VXPT_API int vxpt_shadow_filter(vxpt_handle h, const VxCamera* cam, const VxShadowFilterIn* in, const VxShadowFilterParams* p,
                                float* out /* o_Color, 1 float / pixel */);

/* Both shadow filters of one frame in one call, the temporal planes and the previous frame's hit distances resident in the handle
 * (20 B/pixel): only the shadow pass's planes come in and the filtered plane goes out.  Whole frames only; reset_history / a change
 * of resolution start a new history (previous planes zero, previous camera = this camera). */
typedef struct VxShadowFrameParams {
    float view[16];         /* u_View of this frame; u_PrevView of the next call              */
    float projection[16];   /* u_Projection of this frame; u_PrevProjection of the next call  */
    int32_t reset_history;  /* 1 = forget the previous frame                                  */
    int32_t spatial;        /* DenoiseSunShadows (1): 0 returns the temporal pass's plane     */
    float filter_scale;     /* u_ShadowFilterScale (1.0)                                      */
} VxShadowFrameParams;
VXPT_API int vxpt_shadow_filter_frame(vxpt_handle h, const VxCamera* cam, const VxGBuffer* gbuf /* t, normal_id */,
                                      const VxShadowOut* shadow /* read: this frame's shadow pass */, const VxShadowFrameParams* p,
                                      float* out /* 1 float / pixel */);

/* ---- one frame of the path: the pass sequence of Core/Pipeline.cpp's render loop (:1973-2016 primary, :2795-2852 shadow,
 *      :2174-2281 diffuse GI, :3003-3164 reflections) on the rows of `cam` -------------------------------------------------
 * Equivalent to vxpt_trace_primary + vxpt_trace_shadow + vxpt_trace_diffuse (+ vxpt_trace_reflection) with the same arguments,
 * but the G-buffer (and the GI planes the reflection pass reads) stay in device memory between the passes, as the reference's
 * FBO textures do: nothing is re-uploaded.  shadow / diffuse / reflection may be NULL to skip that pass; any output plane may be
 * NULL.  HOST output planes are copied out by a second stream, slab by slab, while the following slabs and passes are still
 * tracing, and are complete when the call returns; device planes are written in place and complete after vxpt_sync(). */
typedef struct VxFrameParams {
    const VxPrimaryParams* primary;        /* required                                              */
    const VxShadowParams* shadow;          /* NULL = no shadow pass                                 */
    const VxDiffuseParams* diffuse;        /* NULL = no GI pass                                     */
    const VxReflectionParams* reflection;  /* NULL = no reflection pass (needs diffuse)             */
    const float* g_normal;                 /* VxReflectionIn.g_normal / g_pbr for the reflection    */
    const float* g_pbr;                    /* pass (device or host, may be NULL)                    */
    const VxMaterialParams* material;      /* NULL = no G-buffer material pass.  Otherwise it runs right after the primary pass
                                              (Pipeline.cpp:2066-2136) and, where g_normal / g_pbr above are NULL, the reflection
                                              pass reads ITS normal / pbr planes, as in the reference (needs even row slabs)  */
} VxFrameParams;
typedef struct VxFrameOut {
    VxGBuffer gbuffer;
    VxShadowOut shadow;
    VxDiffuseOut diffuse;
    VxReflectionOut reflection;
    VxMaterialOut material;                /* outputs of the material pass (any may be NULL)        */
} VxFrameOut;
VXPT_API int vxpt_render_frame(vxpt_handle h, const VxCamera* cam, const VxFrameParams* p, const VxFrameOut* out);
/* As vxpt_render_frame, but returns as soon as the frame is enqueued; HOST planes are complete after vxpt_frame_wait(h).
 * A handle keeps one frame in flight: any later call that needs the handle's staging memory first waits for it, so the call
 * is always safe; to overlap the copy-out of frame k with the tracing of frame k+1, alternate between two handles (and two
 * sets of host planes), as a double-buffered swap chain does.
 * The call may be recorded into a CUDA graph (stream capture on the handle's stream, no frame pending, scratch sized beforehand by an
 * eager frame or vxpt_reserve): the copy-out stream joins the capture and rejoins the handle's stream at the end, so a replay — kernels
 * and host planes — is complete when the handle's stream is (vxpt_sync); vxpt_frame_wait has nothing to wait for then. */
VXPT_API int vxpt_render_frame_async(vxpt_handle h, const VxCamera* cam, const VxFrameParams* p, const VxFrameOut* out);
VXPT_API int vxpt_frame_wait(vxpt_handle h);

/* ---- peer-to-peer gather of row slabs (multi-GPU, one process per GPU) --------------------------------------
 * The gather root allocates the slab buffer with vxpt_shared_alloc and publishes its 64-byte handle (any byte transport);
 * every other rank maps it with vxpt_shared_open and passes addresses inside the mapping as OUTPUT planes of its trace calls:
 * the kernels then store their rows straight into the root's memory over NVLink, and no separate exchange step exists.
 * vxpt_signal / vxpt_wait_all order the frames: a rank enqueues vxpt_signal(flag, k) after the passes of frame k (a
 * system-scope release: the frame's stores are visible before the flag); the root enqueues vxpt_wait_all over the N flags.
 * Both are stream-ordered (on the handle's stream, or on a caller stream of the same device, e.g. a gather stream that
 * must not stall tracing) and never block the host.  A wait gives up after timeout_ms and latches an
 * error that the next vxpt_sync / vxpt_get_stats reports (VXPT_E_STATE), so a lost peer cannot hang the device. */
/* Sequence-numbered forms for CUDA-graph capture: the value comes from a device-resident counter (device memory of this GPU)
 * that the call itself advances, so the captured node's arguments never change from frame to frame.
 *   vxpt_signal_next : v = ++*counter;  release-store v to *flag.
 *   vxpt_wait_next   : target = ++*counter - lag;  if target >= 1, wait until all n flags have reached target. */
#define VXPT_SHARED_HANDLE_BYTES 64
VXPT_API int vxpt_shared_alloc(vxpt_handle h, size_t bytes, void** dptr, uint8_t handle_out[VXPT_SHARED_HANDLE_BYTES]);
VXPT_API int vxpt_shared_open(vxpt_handle h, const uint8_t handle[VXPT_SHARED_HANDLE_BYTES], void** dptr);
VXPT_API int vxpt_shared_close(vxpt_handle h, void* dptr);  /* unmap (opened) or free (allocated) */
/* stream-ordered device-to-device copy (copy engine; dst / src may be peer mappings): pushes a locally traced slab to the root */
VXPT_API int vxpt_copy_async(vxpt_handle h, void* dst, const void* src, size_t bytes, void* cuda_stream /* NULL = the handle's stream */);
VXPT_API int vxpt_signal(vxpt_handle h, uint32_t* flag /* device, may be a peer mapping */, uint32_t value,
                         void* cuda_stream /* NULL = the handle's stream */);
VXPT_API int vxpt_signal_next(vxpt_handle h, uint32_t* flag, uint32_t* counter, void* cuda_stream);
VXPT_API int vxpt_wait_next(vxpt_handle h, const uint32_t* flags, int n, int stride_words, uint32_t* counter, int lag, int timeout_ms,
                            void* cuda_stream);
VXPT_API int vxpt_wait_all(vxpt_handle h, const uint32_t* flags /* device, may be a peer mapping */, int n, int stride_words,
                           uint32_t at_least /* wrap-around compare */, int timeout_ms, void* cuda_stream /* NULL = the handle's stream */);

/* ---- one frame over N devices from one host thread (SURVEY.md §8b "vxpt_mg_*", §8e) --------------------------------------------
 * The reference drives one GL context from one thread (Core/Pipeline.cpp main loop); vxpt_mg_* keeps that shape for a caller that owns
 * several GPUs in ONE process.  A vxpt_mg_handle is a set of ordinary handles, one per device (vxpt_mg_device(mg, k) — usable with every
 * vxpt_* call above, e.g. for a setter that has no vxpt_mg_ form).  Scene state is replicated: vxpt_mg_upload_world / set_block(s) /
 * build_distance_field / set_* are the same call on every device (the grid is 18.9 MB, an edit list is bytes, a rebuild tens of
 * microseconds: nothing is broadcast between devices).  A frame is sharded as contiguous image row slabs cut at multiples of 8 rows
 * (vxpt_mg_slab tells which): device k runs vxpt_render_frame_async on its slab, all devices side by side.  No exchange step:
 * HOST output planes are filled by every device's own copy stream; DEVICE output planes (memory of any device of the set, normally
 * device 0) are written in place by the other devices' kernels over NVLink — peer access is enabled at creation; where the devices
 * cannot reach each other, device planes are refused with VXPT_E_STATE.  With the reflection pass, neighbouring slabs both write the
 * G-buffer halo rows they share (identical values).  Same threading rule as a handle: one caller thread per vxpt_mg_handle.
 * device_ids may name a device more than once (several slabs on one GPU; how a single-GPU machine exercises this path).
 * The per-process form (one process per GPU under torch.distributed) is the vxpt_shared_ / vxpt_signal / vxpt_wait family above. */
typedef struct vxpt_mg_ctx* vxpt_mg_handle;
VXPT_API int vxpt_mg_create(int n_devices, const int* device_ids /* NULL = 0 .. n_devices-1 */, vxpt_mg_handle* out);
VXPT_API int vxpt_mg_destroy(vxpt_mg_handle mg);
VXPT_API int vxpt_mg_size(vxpt_mg_handle mg);
VXPT_API vxpt_handle vxpt_mg_device(vxpt_mg_handle mg, int k);
VXPT_API int vxpt_mg_upload_world(vxpt_mg_handle mg, const uint8_t* blocks);
VXPT_API int vxpt_mg_set_block(vxpt_mg_handle mg, int x, int y, int z, uint8_t id);
VXPT_API int vxpt_mg_set_blocks(vxpt_mg_handle mg, const int16_t* xyz, const uint8_t* ids, int n);
VXPT_API int vxpt_mg_build_distance_field(vxpt_mg_handle mg);
VXPT_API int vxpt_mg_set_materials(vxpt_mg_handle mg, const int32_t table[768]);
VXPT_API int vxpt_mg_set_blue_noise(vxpt_mg_handle mg, const int32_t* sobol, const int32_t* scramble, const int32_t* rank);
VXPT_API int vxpt_mg_set_material_textures(vxpt_mg_handle mg, const float* albedo_lod3, const float* pbr_lod2, int n_layers,
                                           const float* emissive_lod0, int n_emissive_layers);
VXPT_API int vxpt_mg_set_reflection_textures(vxpt_mg_handle mg, const float* normal_lod3, int n_normal_layers, const float* emissive_lod2,
                                             int n_emissive_layers);
VXPT_API int vxpt_mg_set_sky_cubemap(vxpt_mg_handle mg, const float* rgb, int n);
VXPT_API int vxpt_mg_set_shadow_noise(vxpt_mg_handle mg, const uint8_t* rgba8);
VXPT_API int vxpt_mg_set_option(vxpt_mg_handle mg, int option, int value);
/* rows [*row_begin, *row_end) of the frame selected by cam that device k traces */
VXPT_API int vxpt_mg_slab(vxpt_mg_handle mg, const VxCamera* cam, int k, int* row_begin, int* row_end);
/* vxpt_render_frame over all devices: returns with host planes complete and device planes written */
VXPT_API int vxpt_mg_render_frame(vxpt_mg_handle mg, const VxCamera* cam, const VxFrameParams* p, const VxFrameOut* out);
VXPT_API int vxpt_mg_render_frame_async(vxpt_mg_handle mg, const VxCamera* cam, const VxFrameParams* p, const VxFrameOut* out);
VXPT_API int vxpt_mg_frame_wait(vxpt_mg_handle mg);
VXPT_API int vxpt_mg_sync(vxpt_mg_handle mg);
VXPT_API int vxpt_mg_get_stats(vxpt_mg_handle mg, VxStats* out);  /* counters summed over the devices, times = the slowest device's */
VXPT_API int vxpt_mg_reset_stats(vxpt_mg_handle mg);

/* ---- glFinish (Core/Pipeline.cpp:4782), statistics ------------------------------------------------------- */
VXPT_API int vxpt_sync(vxpt_handle h);
VXPT_API int vxpt_get_stats(vxpt_handle h, VxStats* out);   /* totals since the last reset; syncs */
VXPT_API int vxpt_reset_stats(vxpt_handle h);
VXPT_API int vxpt_launch_count(vxpt_handle h, uint64_t* kernels_launched); /* kernels this handle launched */
/* the CUDA stream (cudaStream_t) the handle enqueues on, so callers can record events around passes */
VXPT_API int vxpt_stream(vxpt_handle h, void** cuda_stream);
/* Size the handle's device scratch up front — what the reference does when it (re)creates its FBOs for a window size
 * (Core/Pipeline.cpp:1094-1152 via Framebuffer::SetSize) instead of inside a draw.  The passes grow their scratch on demand (GI wavefront
 * queue: 48 B per slab pixel, + 48 B per frame pixel when a pixel takes several samples; staging arena for host-pointer planes), which
 * needs a stream synchronisation and a reallocation: legal any time EXCEPT while the handle's stream is being captured into a CUDA graph,
 * where such a pass returns VXPT_E_STATE.  Call this once (or run one eager frame of the same size) before capturing.
 * cam: the frame size and the row slab the handle will trace; max_gi_spp: the largest VxDiffuseParams.spp to come (0 = no GI pass);
 * staging_bytes: host-plane staging to keep (0 = none). */
VXPT_API int vxpt_reserve(vxpt_handle h, const VxCamera* cam, int max_gi_spp, size_t staging_bytes);

/* ---- tuning knobs (do not change results) ---------------------------------------------------------------- */
#define VXPT_OPT_TRAVERSAL_LAYOUT 1 /* 0 = linear distance field, 1 = brick-swizzled copy (default) */
#define VXPT_OPT_GI_WAVEFRONT 2     /* 0 = one thread per pixel, 1 = wavefront: sorted first-bounce rays, hits re-queued, the rest of a sample in
                                       CTA-wide stages (default) */
#define VXPT_OPT_DF_ALGO 3          /* 0 = one thread per grid line (reference-shaped) + a separate step-field pass, 1 = DPX sweeps whose z sweep
                                       writes the traversal's step field too (default) */
/* measurement knob: keep `value` (1..8) identical copies of the grid + step field at distinct addresses and rotate through
 * them, one per vxpt_trace_primary call (= per frame).  With 3 copies the traced inputs (132 MB) exceed the 126 MB L2, so
 * back-to-back frames cannot reuse each other's cache lines (benchmark timing rule); results are unchanged. */
#define VXPT_OPT_SCENE_REPLICAS 4
/* 1 (default): record CUDA events around every pass (VxStats.last_ms).  0: record none, so a caller may capture the
 * handle's stream into a CUDA graph (event timing is not capturable). */
#define VXPT_OPT_TIMING_EVENTS 5
/* texel format of the image planes (changes the ENCODING of results, not the results: every value is the fp32 value rounded once).
 * 0 (default): fp32 planes as the struct pointer types say.
 * 1: the reference's FBO attachment formats (Core/Pipeline.cpp:1094-1095 G-buffer, :1152 shadow, :1102 GI, :1141 reflection):
 *    VxGBuffer.t R16F (2 B), inv_t R32F (4 B), normal_id / block_id R8;  VxShadowOut.shadow R8, transversal R16F;
 *    VxDiffuseOut.sh RGBA16F (8 B), cocg RG16F (4 B), luma R16F (2 B), ao_sky RG8 unorm (2 B);
 *    VxReflectionOut.color RGBA16F, hit_distance R16F, emissive_mask R8;  VxReflectionIn.sh / cocg as VxDiffuseOut wrote them
 *    (g_normal / g_pbr stay fp32).  Halves are IEEE binary16 rounded to nearest even, unorm8 = round(v * 255).
 *    27 B/pixel for G-buffer + shadow + GI instead of 51: what a host consumer or a peer GPU has to receive. */
#define VXPT_OPT_TEXEL_FORMAT 6
/* G-buffer material pass, default instantiation (no parallax, no lava id): 0 = every thread re-derives the surface UV of its two
 * quad partners (three ray set-ups per pixel), 1 (default) = the partners' UV arrive by warp shuffle (one ray set-up per pixel).  Same
 * operands, same planes (GPU test + emulated shuffles on the host); B200, 1080p: 0.0532 ms against 0.0649 ms (profiles/r02a_material_probe.json). */
#define VXPT_OPT_REFLECTION_WAVEFRONT 8 /* reflection pass: 0 = one thread per pixel (the shader's loop as it stands), 1 = samples re-queued: reflection
                                          rays traced per pixel, hits compacted and shaded by dense warps (default); identical planes */
#define VXPT_OPT_MATERIAL_QUAD_SHUFFLE 7
VXPT_API int vxpt_set_option(vxpt_handle h, int option, int value);

/* ---- microbenchmark: resident-set random 32-byte-sector read throughput, the denominator of the
 *      traversal roofline.  Returns GB/s of sectors. ---------------------------------- */
VXPT_API int vxpt_measure_l2_sector_peak(vxpt_handle h, double* gbytes_per_s);

#ifdef __cplusplus
}
#endif
#endif /* VXPT_H_ */
