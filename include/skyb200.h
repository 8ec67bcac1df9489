/*
 * skyb200.h -- C ABI of the B200-native SkyRendering hot path (libskyb200.so).
 *
 * The reference has no FFI; its seam is the C++ class API of the render subsystems plus the
 * std140 blocks they upload.  Each entry point below replaces one
 * host call (+ the GLSL programs it dispatches); the citation names that call.
 *
 * The same declarations are compiled a second time with the prefix orc_ by the CPU oracle
 * (oracle/oracle_api.cpp) so that a test drives both through one binding; the oracle is test
 * infrastructure only and is never linked into or called from this library.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; sky_last_error() gives the text
 *     (the reference throws std::runtime_error, e.g. VolumetricCloud.cpp:169-170).
 *   - one context per GPU, externally synchronised, all work on the stream given at creation;
 *     no hidden host synchronisation except in sky_read_resource / *_host entry points.
 *   - "dev" pointers are device pointers (host pointers for the oracle build).
 *   - images: row-major, x fastest, row 0 = GL texel row 0 (bottom of the screen).
 */
#ifndef SKYB200_H
#define SKYB200_H

#include "sky_types.h"

#ifndef SKY_FN
#define SKY_FN(name) sky_##name
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct SkyContext SkyContext;

/* Context = the GL context + every GLTexture/GLBuffer the subsystems own
 * (Atmosphere.h:84-91, AtmosphereRenderer.h:118-128, VolumetricCloud.h:99-128). */
int SKY_FN(ctx_create)(int device, void* cuda_stream, SkyContext** out);
void SKY_FN(ctx_destroy)(SkyContext* ctx);
const char* SKY_FN(last_error)(SkyContext* ctx);
/* Block until all work queued on the context's stream has finished. */
int SKY_FN(sync)(SkyContext* ctx);

/* Textures::Textures blue-noise upload (src/Base/src/Textures.cpp:19-26): 64x64 R16, rows already
 * in GL order (i.e. the PNG flipped vertically, StbImage.cpp:12-17). */
int SKY_FN(set_blue_noise)(SkyContext* ctx, const uint16_t* host_texels_64x64);

/* Textures::Textures star map upload (src/Base/src/Textures.cpp:43-50): GL_SRGB8 texels, RGB8 [H][W][3], rows in GL order; K6 adds
 * star_luminance_scale * texture(star_luminance, equirect(view_direction)) to sky pixels outside the sun disc
 * (AtmosphereRenderer.glsl:326-331,427-429; sampler LinearNoMipmapClampToEdge: level 0, texels decoded to linear before
 * filtering).  Without a star map the term is absent (a black sky behind the atmosphere).  width or height 0 removes it. */
int SKY_FN(set_star_map)(SkyContext* ctx, const uint8_t* host_srgb8, int width, int height);

/* VolumetricCloud::SetViewport (VolumetricCloud.cpp:115-136): (re)allocates the per-viewport
 * targets and zero-fills the temporal histories. */
int SKY_FN(set_viewport)(SkyContext* ctx, int width, int height);

/* Atmosphere::UpdateLuts (Atmosphere.cpp:101-124): K1 transmittance + K2 multiscattering. */
int SKY_FN(atmosphere_bake)(SkyContext* ctx, const SkyAtmosphereBufferData* atmosphere);

/* AtmosphereRenderer::Render up to the environment cube (AtmosphereRenderer.cpp:164-242):
 * K3 sky-view, K4 aerial perspective, K5 environment luminance. */
int SKY_FN(atmosphere_luts)(SkyContext* ctx, const SkyAtmosphereRenderBufferData* render,
                            const SkyLutConfig* config);

/* Textures::Textures environment-BRDF LUT bake (src/Base/src/Textures.cpp:60-75, shaders/Base/EnvBRDFLut.comp, K22):
 * split-sum integral of the GGX BRDF, 1024 Hammersley samples per texel -> SKY_RES_ENV_BRDF_LUT (GL_RG16 512x512).
 * Input-free; the reference bakes it once at start-up. */
int SKY_FN(env_brdf_lut)(SkyContext* ctx);

/* The tail of AtmosphereRenderer::Render's LUT phase (AtmosphereRenderer.cpp:242-244): glGenerateTextureMipmap of the
 * environment cube (-> SKY_RES_ENVIRONMENT_MIPS) and IBL::Precompute (src/Base/src/IBL.cpp:25-45): K23 EnvRadianceSH.comp
 * (-> SKY_RES_ENV_RADIANCE_SH) and K24 PrefilterRadiance.comp, one dispatch per roughness level
 * (-> SKY_RES_PREFILTERED_RADIANCE).  Call after sky_atmosphere_luts; inputs of the object shading of the composite
 * (ComputeObjectLuminance / GetAmbient, AtmosphereRenderer.glsl:284-324, BRDF.glsl:108-130). */
int SKY_FN(ibl_precompute)(SkyContext* ctx);

/* AtmosphereRenderParameters::albedo / normal / orm (AtmosphereRenderer.h:37-41): the G-buffer the object branch of K6 reads,
 * in the formats GBuffer.cpp:19-21 allocates -- albedo GL_RGBA8 uchar4[H][W], normal GL_RGBA16_SNORM short4[H][W], orm
 * GL_RGBA16 ushort4[H][W] (occlusion, roughness, metallic) -- at the size of the following sky_composite calls.  The
 * pointers are borrowed until the next call; three null pointers unbind the G-buffer.  It is an INPUT: the meshes / ground
 * pass that would rasterise it are outside the path. */
int SKY_FN(set_gbuffer)(SkyContext* ctx, const void* albedo_dev, const void* normal_dev, const void* orm_dev);

/* Textures::Textures earth albedo upload (src/Base/src/Textures.cpp:52-58): GL_SRGB8 texels RGB8 [H][W][3], rows in GL order, +
 * glGenerateTextureMipmap (2x2 box on the decoded values, re-encoded to sRGB8, floor sizes).  An equirectangular map: u = longitude,
 * v = latitude (EarthRender.frag:22-26).  width or height 0 removes it (the ground pass then writes albedo 0). */
int SKY_FN(set_earth_albedo)(SkyContext* ctx, const uint8_t* host_srgb8, int width, int height);

/* Clear(const GBuffer&) (src/Base/include/GBuffer.h:28-34; AppWindow::Render, AppWindow.cpp:168): the three colour targets to 0, the depth
 * plane to 1.0 -- what a frame starts from before the ground pass (and the meshes, which are outside the path) fill the G-buffer. */
int SKY_FN(gbuffer_clear)(SkyContext* ctx, float* depth_dev, void* albedo_dev, void* normal_dev, void* orm_dev, int width, int height);

/* Earth::RenderToGBuffer (src/SkyRendering/Earth.cpp:46-65, shaders/SkyRendering/EarthRender.frag, K7): the analytic ground pass.  For every pixel
 * whose view ray meets the ground sphere in front of what the depth buffer already holds it writes gl_FragDepth (quantised to
 * the D24 of GBuffer.cpp:22) and the three G-buffer targets in the formats of GBuffer.cpp:19-21 -- albedo GL_RGBA8 from the earth
 * map through textureGrad with the seamless-longitude gradients of :27-35 (sampler of Earth.cpp:34-42; include/sky_texgrad.h),
 * normal GL_RGBA16_SNORM = the sphere normal, ORM GL_RGBA16 = (1, 1, 0, 1); every other pixel is left untouched (`discard`).
 * depth_dev float[H][W] is read and written; needs sky_atmosphere_bake (bottom_radius).  The G-buffer it fills is what
 * sky_set_gbuffer binds for the composite. */
int SKY_FN(earth_gbuffer)(SkyContext* ctx, const SkyEarthBufferData* earth, float* depth_dev, void* albedo_dev, void* normal_dev,
                          void* orm_dev, int width, int height);

/* AtmosphereRenderer::Render full-screen pass (AtmosphereRenderer.cpp:246-250, K6), sky / aerial
 * perspective / sun-disc branches.  depth_dev: float[H][W] in [0,1]; hdr_dev: half4[H][W] (written).
 * Object pixels (depth != 1): with a G-buffer bound (sky_set_gbuffer; needs sky_env_brdf_lut and sky_ibl_precompute) they are
 * shaded like the reference's (ComputeObjectLuminance + SampleVisibilityFromShadowMap, AtmosphereRenderer.glsl:284-343,
 * 404-410: sun through the transmittance LUT, GGX / Lambert BRDF, SH9 + prefiltered-cube ambient, mesh shadow map x cloud
 * shadow map, with PCSS soft shadows when SkyLutConfig.pcss is set); without one they receive the atmosphere in-scatter
 * only -- the reference's result on a cleared G-buffer.  Alpha is 1 in every pixel (AtmosphereRenderer.glsl:431). */
int SKY_FN(composite)(SkyContext* ctx, const float* depth_dev, void* hdr_dev, int width, int height);

/* DynamicTexture::Generate (VolumetricCloudDefaultMaterial.h:39-49): K8/K9/K10 + mip chain. */
int SKY_FN(noise_generate)(SkyContext* ctx, int kind, const SkyNoiseCreateInfo* info);

/* VolumetricCloudVoxelMaterial ctor upload (VolumetricCloudVoxelMaterial.cpp:72-75): R8 grid
 * [dz][dy][dx] (dx = vdb x, dy = vdb z, dz = vdb y) + mip chain. */
int SKY_FN(voxel_upload)(SkyContext* ctx, const uint8_t* host_voxels, int dx, int dy, int dz);

/* IVolumetricCloudMaterial::Update/Bind (IVolumetricCloudMaterial.h:17-19): material uniforms. */
int SKY_FN(set_material)(SkyContext* ctx, const SkyMaterialBlock* material);

/* VolumetricCloud::RenderShadow (VolumetricCloud.cpp:282-325): K11 -> K12 x2 -> K13. */
int SKY_FN(cloud_shadow)(SkyContext* ctx, const SkyCloudCommonBufferData* common);

/* VolumetricCloud::Render (VolumetricCloud.cpp:327-423): K14 -> K15 -> K16 -> K17 -> K18.
 * depth_dev float[H][W]; hdr_dev half4[H][W] read-modify-write. */
int SKY_FN(cloud_frame)(SkyContext* ctx, const SkyCloudCommonBufferData* common,
                        const SkyCloudBufferData* cloud, const float* depth_dev, void* hdr_dev);

/* The same frame split at the one exchange point a tile-sharded run needs:
 * _begin runs K14, K15 and K16 for quarter-res rows r with (r / band_rows) % band_count == band_index;
 * _end runs K17, K18 once every row of SKY_RES_CLOUD_RENDER / SKY_RES_CLOUD_DISTANCE is present. */
int SKY_FN(cloud_frame_begin)(SkyContext* ctx, const SkyCloudCommonBufferData* common,
                              const SkyCloudBufferData* cloud, const float* depth_dev,
                              int band_rows, int band_index, int band_count);
int SKY_FN(cloud_frame_end)(SkyContext* ctx, const float* depth_dev, void* hdr_dev);

/* Row bands of the FULL-RES passes of a tile-sharded frame: with band_count > 1 the composite (K6) and the
 * upscale (K18) only touch full-res rows r with (r / band_rows) % band_count == band_index (band_rows a multiple of 8); K17 and
 * everything else stay complete.  Every pixel of those rows is exactly what the unsharded frame computes; the caller gathers
 * the HDR rows of the ranks (skyrendering_b200/distributed.py).  band_count <= 1 restores whole frames. */
int SKY_FN(set_output_bands)(SkyContext* ctx, int band_rows, int band_index, int band_count);

/* Tile-sharded K16 fused with its exchange over peer memory (NVLink / NVSwitch), one process per GPU.
 * sky_peer_export fills the CUDA IPC handles of this context's K16 outputs and of its arrival flags;
 * the host exchanges them (any transport) and hands all ranks' handles to sky_peer_attach.  From then
 * on cloud_frame_begin with band_count == world stores every texel it renders into EVERY rank's
 * SKY_RES_CLOUD_RENDER / SKY_RES_CLOUD_DISTANCE (plain stores to mapped peer pointers inside K16), and
 * cloud_frame_end first publishes this rank's arrival flag to all peers and waits for theirs on the
 * device -- no host synchronisation, no separate all-gather. */
typedef struct SkyPeerHandles {
    unsigned char render[64];    /* cudaIpcMemHandle_t of half4[H/4][W/4] */
    unsigned char distance[64];  /* cudaIpcMemHandle_t of float[H/4][W/4] */
    unsigned char flags[64];     /* cudaIpcMemHandle_t of the arrival / done flag words */
    unsigned char hdr[64];       /* cudaIpcMemHandle_t of SKY_RES_FRAME_HDR, half4[H][W] (sky_set_output_gather) */
} SkyPeerHandles;
#define SKY_MAX_PEERS 8
int SKY_FN(peer_export)(SkyContext* ctx, SkyPeerHandles* out);
int SKY_FN(peer_attach)(SkyContext* ctx, int rank, int world_size, const SkyPeerHandles* all_ranks);
int SKY_FN(peer_detach)(SkyContext* ctx);
/* The frame target of a tile-sharded frame whose FULL-RES passes are sharded too (sky_set_output_bands): with peers attached and
 * `hdr_dev` of sky_composite / sky_cloud_frame_end == the context's own SKY_RES_FRAME_HDR (allocated and exported by sky_peer_export),
 * K18 -- the last writer of a frame -- stores every texel of its row bands straight into the frame targets of the receiving ranks over
 * NVLink, and cloud_frame_end ends with the same device-side arrival barrier as the K16 exchange: no NCCL all-gather, no host
 * synchronisation.  mode: SKY_GATHER_OFF (default), SKY_GATHER_ALL (every rank ends with the whole frame), SKY_GATHER_ROOT (rank 0 -- the
 * rank that displays -- does).  A rank's next frame releases the previous frame's target when its cloud_frame_begin is reached in stream
 * order: whatever reads the frame must be queued on the caller's stream before that. */
int SKY_FN(set_output_gather)(SkyContext* ctx, int mode);

/* cloud_frame with HOST buffers: copies depth and hdr in, runs the frame, copies hdr out and
 * synchronises -- the call a host application without device pointers makes. */
int SKY_FN(cloud_frame_host)(SkyContext* ctx, const SkyCloudCommonBufferData* common,
                             const SkyCloudBufferData* cloud, const float* depth_host, void* hdr_host);

/* VolumetricCloud::PathTracing ctor (VolumetricCloud.cpp:495-531): allocates + clears accumulators. */
int SKY_FN(pt_begin)(SkyContext* ctx, const SkyPathTracingInit* init);

/* PathTracing::Render render pass (VolumetricCloud.cpp:533-560, K19) for kFrameId in
 * [frame_begin, frame_begin+count) over kRenderRegion = {x0,y0,x1,y1}. */
int SKY_FN(pt_samples)(SkyContext* ctx, const SkyCloudCommonBufferData* common,
                       uint32_t frame_begin, uint32_t count, const int32_t region[4]);

/* Collision sampling of K19.  SKY_PT_TRACKING_REFERENCE (default) is the reference's algorithm with the reference's random
 * streams (VolumetricCloudPathTracing.comp:135-196: one global majorant kSigmaTMax inside the +-region box).
 * SKY_PT_TRACKING_MAJORANT_GRID keeps the estimator (delta tracking + ratio tracking + NEE) but samples collisions against
 * local majorants stored per 8^3-texel macro cell of the voxel texture and skips empty cells: the image has the same
 * expectation, not the same samples -- a fast mode that is validated statistically and reported separately.  Voxel material only. */
int SKY_FN(pt_set_tracking)(SkyContext* ctx, int mode);

/* PathTracing::Render display pass (VolumetricCloud.cpp:562-565, K20): hdr = hdr*avg.a + avg.rgb,
 * avg = accum / frame_count. */
int SKY_FN(pt_resolve)(SkyContext* ctx, uint32_t frame_count, void* hdr_dev);

/* HDRBuffer::DoPostProcessAndBindSdrFramebuffer's last pass (src/Base/src/HDRBuffer.cpp:96-100, shaders/Base/BloomPass2.frag:15-42)
 * without the bloom term (bloom_intensity 0: BloomPass1 + the blur pyramid are display sugar outside the path):
 * ToneMapping(luminance, exposure) (CE or ACES, HDRBuffer.h:11-22) -> pow(1 / 2.2) [-> + blue noise / 255 when dither != 0]
 * -> RGBA8 (round to nearest even, alpha 255).  hdr_dev half4[H][W] in, rgba8_dev uchar4[H][W] out. */
int SKY_FN(tonemap)(SkyContext* ctx, const void* hdr_dev, int width, int height, const SkyToneMapParams* params, void* rgba8_dev);

/* pt_samples + host read-back of the accumulation buffer (float4[H][W]). */
int SKY_FN(pt_samples_host)(SkyContext* ctx, const SkyCloudCommonBufferData* common,
                            uint32_t frame_begin, uint32_t count, const int32_t region[4],
                            float* accum_host);

/* Texture getters (Atmosphere.h:76-82, AtmosphereRenderer.h:96-108, VolumetricCloud.h:59-73). */
int SKY_FN(get_resource)(SkyContext* ctx, int resource, SkyResourceDesc* out);
int SKY_FN(read_resource)(SkyContext* ctx, int resource, void* host_dst, uint64_t bytes);
/* Overwrite a resource from host memory (used to inject LUTs / histories in tests and to
 * install the result of a collective). */
int SKY_FN(write_resource)(SkyContext* ctx, int resource, const void* host_src, uint64_t bytes);

/* Work counters for the roofline: enable != 0 switches kernels to their counting
 * variants; counters live in SKY_RES_COUNTERS and are reset here. */
int SKY_FN(counters_enable)(SkyContext* ctx, int enable);
/* Number of kernel launches this context has issued since it was created (host-side count, one per launch; memsets and copies are not
 * kernels).  bench.py reports the difference across its timed region as `gpu_launches`. */
int SKY_FN(launch_count)(SkyContext* ctx, uint64_t* launches);

/* Pin the launch shape of a kernel that has several (SkyKernelId / SkyK16Shape, sky_types.h); SKY_K16_AUTO restores the per-launch choice. */
int SKY_FN(set_launch_shape)(SkyContext* ctx, int kernel, int shape);

/* Which filtering the material textures use: 0 = exact fp32 software filtering on gathered
 * texels (default; matches the oracle), 1 = hardware linear filtering (8-bit weights). */
int SKY_FN(set_hw_filtering)(SkyContext* ctx, int enable);
/* Validation mode: 1 routes the frame kernels (K6, K11-K18) and the path tracer (K19/K20) to objects compiled from the
 * same sources WITHOUT FMA contraction and with IEEE division / square root / elementary functions, i.e. the unfused
 * fp32 arithmetic the oracle (and a GLSL compiler honouring `precise`) performs, operation by operation.  The default
 * (0) objects contract FMAs and use the hardware approximations, like a GLSL compiler is free to (frames are compared at
 * relative RMS 1e-2 because the altitude |p| - R of VolumetricCloudCommon.glsl:36-39 cancels four digits, which makes any
 * frame sensitive to contraction at the 1e-3 level).  The LUT bake and the noise kernels are always strict. */
int SKY_FN(set_strict_arithmetic)(SkyContext* ctx, int enable);
/* Arithmetic of the per-frame LUT marches K2, K3 and K4 (Atmosphere.glsl:220-295 as dispatched by Atmosphere.cpp:116 and
 * AtmosphereRenderer.cpp:222-233).  SKY_LUT_EXACT (default): one thread per march, the shader's unfused fp32 operation order with IEEE division /
 * square root and the deterministic elementary functions -- bit-identical to the oracle.  SKY_LUT_COOPERATIVE: the production march --
 * four lanes fold contiguous chunks of a march's steps into affine maps (L, T) -> (L + T A, T T_i) and compose them in order; fused
 * multiply-adds and the hardware ex2 / rcp / sqrt where the result is well conditioned, the shader's own expression for r_i.  Same quadrature,
 * same ray set-up; the LUTs agree with the exact ones to the tolerance stated in tests/test_gpu_parity.py
 * (test_cooperative_lut_bake_within_tolerance).  K1 and K5 are unaffected; with MOON_SHADOW_ENABLE or VOLUMETRIC_LIGHT_ENABLE K3 / K4 stay
 * on the exact kernel. */
#define SKY_LUT_EXACT 0
#define SKY_LUT_COOPERATIVE 1
int SKY_FN(set_lut_arithmetic)(SkyContext* ctx, int mode);

/* Opt-in overlap of the two independent halves of a frame (AppWindow::Render, AppWindow.cpp:148-175) on a second,
 * internal stream: {cloud shadow chain K11-K13, K14-K17} run beside {K3-K5, composite K6}; K18 joins them.  Results are
 * bit-identical.  While enabled, the work of cloud_shadow / cloud_frame_begin is ordered on the caller's stream at
 * cloud_frame_end and in every other entry point (they join first), not at the return of those two calls. */
int SKY_FN(set_frame_overlap)(SkyContext* ctx, int enable);

/* Opt-in pipelining of consecutive frames: the context keeps two sets of the atmosphere LUTs; sky_atmosphere_bake flips to the
 * set the frame before last used and runs the LUT phase (K1-K5, small latency-bound kernels that gate both full-machine
 * kernels of a frame) on an internal high-priority stream beside the PREVIOUS frame's K6 / K16; the caller's stream waits for
 * it where the LUTs are first read.  Results are bit-identical; getters return the newest set.  Independent of
 * sky_set_frame_overlap (they compose). */
int SKY_FN(set_frame_pipelining)(SkyContext* ctx, int enable);

/* Microbenchmark for the texture-pipe roofline: launches `iters` dependent-free trilinear R8
 * fetches per thread over the detail volume and returns texel-quads per second. */
int SKY_FN(tex_peak)(SkyContext* ctx, int mode, double* fetches_per_second);

#ifdef __cplusplus
}
#endif
#endif /* SKYB200_H */
