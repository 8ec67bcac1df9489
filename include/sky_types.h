/*
 * sky_types.h -- plain-old-data contract shared by the C-ABI library (libskyb200.so),
 * the host library (libskyhost.so) and the CPU oracle (oracle/liboracle.so).
 *
 * Every struct here is the byte-for-byte std140 uniform block the reference host fills and
 * the reference shaders read.  Keeping them verbatim is what makes the CUDA path a drop-in:
 * a maintainer hands the very same bytes to sky_* that they hand to glNamedBufferSubData.
 * Matrices are column-major (glm), m[c*4+r].
 *
 * Reference citations are relative to /root/reference.
 */
#ifndef SKY_TYPES_H
#define SKY_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* src/SkyRendering/Atmosphere.cpp:21-47, shaders/SkyRendering/Atmosphere.glsl:6-32 (128 B) */
typedef struct SkyAtmosphereBufferData {
    float solar_illuminance[3];
    float sun_angular_radius;
    float rayleigh_scattering[3];
    float inv_rayleigh_exponential_distribution;
    float mie_scattering[3];
    float inv_mie_exponential_distribution;
    float mie_absorption[3];
    float ozone_center_altitude;
    float ozone_absorption[3];
    float inv_ozone_width;
    float ground_albedo[3];
    float mie_phase_g;
    float _atmosphere_padding[3];
    float multiscattering_mask;
    float bottom_radius;
    float top_radius;
    float transmittance_steps;
    float multiscattering_steps;
} SkyAtmosphereBufferData;

/* src/SkyRendering/AtmosphereRenderer.cpp:25-50, shaders/SkyRendering/AtmosphereRenderer.glsl:25-50 (320 B) */
typedef struct SkyAtmosphereRenderBufferData {
    float sun_direction[3];
    float star_luminance_scale;
    float earth_center[3];
    float camera_earth_center_distance;
    float camera_position[3];
    float raymarching_steps;
    float up_direction[3];
    float sky_view_lut_steps;
    float right_direction[3];
    float aerial_perspective_lut_steps;
    float front_direction[3];
    float aerial_perspective_lut_max_distance;
    float moon_position[3];
    float moon_radius;
    float inv_view_projection[16];
    float light_view_projection[16];
    float padding00;
    float uInvShadowFroxelMaxDistance;
    float blocker_kernel_size_k;
    float pcss_size_k;
    float uCloudShadowMapMat[16];
} SkyAtmosphereRenderBufferData;

/* src/SkyRendering/VolumetricCloud.cpp:11-33, shaders/SkyRendering/VolumetricCloudCommon.glsl:6-28 (384 B) */
typedef struct SkyCloudCommonBufferData {
    float uInvMVP[16];
    float uReprojectMat[16];
    float uLightVP[16];
    float uInvLightVP[16];
    float uShadowMapReprojectMat[16];
    float uCameraPos[3];
    uint32_t uBaseShadingIndex;
    float uLinearDepthParam[2];
    float uBottomAltitude;
    float uTopAltitude;
    float uSunDirection[3];
    float uFrameID;
    float uInvShadowFroxelMaxDistance;
    float uAerialPerspectiveLutMaxDistance;
    float uShadowFroxelMaxDistance;
    float uEarthRadius;
} SkyCloudCommonBufferData;

/* src/SkyRendering/VolumetricCloud.cpp:35-50, shaders/SkyRendering/VolumetricCloudRender.comp:17-32 (64 B) */
typedef struct SkyCloudBufferData {
    float uSunIlluminanceScale;
    float uMaxRaymarchDistance;
    float uMaxRaymarchSteps;
    float uMaxVisibleDistance;
    float uEnvColorScale[3];
    float uShadowSteps;
    float uSunMultiscatteringSigmaScale;
    float uEnvMultiscatteringSigmaScale;
    float uShadowDistance;
    float uEnvBottomVisibility;
    float padding_[3];
    float uEnvSunHeightCurveExp;
} SkyCloudBufferData;

/* src/SkyRendering/VolumetricCloudDefaultMaterial.cpp:9-22 (64 B) */
typedef struct SkySampleInfo {
    float bias[2];
    float frequency;
    float k_lod;
} SkySampleInfo;

typedef struct SkyMaterialCommonBufferData {
    SkySampleInfo uCloudMapSampleInfo;
    SkySampleInfo uDetailSampleInfo;
    SkySampleInfo uDisplacementSampleInfo;
    float padding0[2];
    float uLodBias;
    float uDensity;
} SkyMaterialCommonBufferData;

/* src/SkyRendering/VolumetricCloudDefaultMaterial.cpp:183-187 (16 B) */
typedef struct SkyMaterial0BufferData {
    float uDetailParam[2];
    float uDisplacementScale;
    float padding1;
} SkyMaterial0BufferData;

/* src/SkyRendering/VolumetricCloudDefaultMaterial.cpp:228-237 (32 B) */
typedef struct SkyMaterial1BufferData {
    float uBaseDensityThreshold;
    float uBaseHeightHardness;
    float uBaseEdgeHardness;
    float uDetailBase;
    float uDetailScale;
    float uHeightCut;
    float uEdgeCur;
    float padding1;
} SkyMaterial1BufferData;

/* src/SkyRendering/VolumetricCloudVoxelMaterial.cpp:17-24 (32 B) */
typedef struct SkyMaterialVoxelBufferData {
    float uSampleFrequency[2];
    float uLodBias;
    float uDensity;
    float uSampleBias[2];
    float uSampleLodK;
    float voxel_material_padding;
} SkyMaterialVoxelBufferData;

/* src/SkyRendering/VolumetricCloudMinimalMaterial.cpp:5-8 (16 B) */
typedef struct SkyMaterialMinimalBufferData {
    float padding[3];
    float uDensity;
} SkyMaterialMinimalBufferData;

/* One tagged block that carries whichever material the scene uses.
 * The reference binds the common block at UBO 3 and the specific one at UBO 4
 * (VolumetricCloudDefaultMaterial.cpp:104-117,214-217); Voxel/Minimal bind theirs at UBO 3. */
enum SkyMaterialType {
    SKY_MATERIAL_DEFAULT0 = 0, /* VolumetricCloudDefaultMaterial0.glsl */
    SKY_MATERIAL_DEFAULT1 = 1, /* VolumetricCloudDefaultMaterial1.glsl */
    SKY_MATERIAL_MINIMAL = 2,  /* VolumetricCloudMaterialMinimal.glsl  */
    SKY_MATERIAL_VOXEL = 3     /* VolumetricCloudMaterialVoxel.glsl    */
};

typedef struct SkyMaterialBlock {
    int32_t type; /* SkyMaterialType */
    int32_t _pad[3];
    SkyMaterialCommonBufferData common; /* DEFAULT0 / DEFAULT1 */
    union {
        SkyMaterial0BufferData m0;
        SkyMaterial1BufferData m1;
        SkyMaterialVoxelBufferData voxel;
        SkyMaterialMinimalBufferData minimal;
    } u;
} SkyMaterialBlock;

/* src/SkyRendering/VolumetricCloudDefaultMaterial.h:57-61, shaders/SkyRendering/NoiseGen.comp:5-10 (16 B).
 * Stored as int on the C++ side, read as uint by the shader. */
typedef struct SkyNoiseCreateInfo {
    uint32_t seed;
    uint32_t base_frequency;
    float remap_min;
    float remap_max;
} SkyNoiseCreateInfo;

enum SkyNoiseKind {
    SKY_NOISE_CLOUD_MAP = 0,    /* CLOUD_MAP_GEN:    RG8 512x512,   info[0]=uDensity info[1]=uHeight */
    SKY_NOISE_DETAIL = 1,       /* DETAIL_MAP_GEN:   R8 128^3,      info[0]=uPerlin  info[1]=uWorley */
    SKY_NOISE_DISPLACEMENT = 2  /* DISPLACEMENT_GEN: RGBA8 128x128, info[0]=uPerlin                  */
};

/* Compile-time permutation flags of AtmosphereRenderer (AtmosphereRenderer.h:44-66,
 * AtmosphereRenderer.cpp:91-107) plus the LUT sizes the reference hard-codes (:15-23). */
typedef struct SkyLutConfig {
    int32_t sky_view_width;   /* reference: 128 (AtmosphereRenderer.cpp:15) */
    int32_t sky_view_height;  /* reference: 128 */
    int32_t aerial_perspective_depth; /* aerial_perspective_lut_depth; width/height are 32 */
    int32_t environment_size; /* reference: 128 (AtmosphereRenderer.cpp:23) */
    int32_t use_sky_view_lut;
    int32_t use_aerial_perspective_lut;
    int32_t sky_view_dither;            /* sky_view_lut_dither_sample_point_enable            */
    int32_t aerial_perspective_dither;  /* aerial_perspective_lut_dither_sample_point_enable  */
    int32_t raymarching_dither;         /* raymarching_dither_sample_point_enable             */
    int32_t moon_shadow;                /* moon_shadow_enable: eclipse term (Atmosphere.glsl:190-218,281-284)                 */
    int32_t volumetric_light;           /* volumetric_light_enable: mesh shadow map in the march (Atmosphere.glsl:180-188,274-277) */
    int32_t pcss;                       /* pcss_enable: percentage-closer soft shadows from the mesh shadow map on object pixels
                                           (PCSS_ENABLE, AtmosphereRenderer.cpp:99; Shadow.glsl:85-99); needs a G-buffer (sky_set_gbuffer) */
} SkyLutConfig;

/* Launch shapes a caller may pin (sky_set_launch_shape) -- the counterpart of GLReloadableComputeProgram's run-time workgroup-size
 * candidates (src/Base/include/GLReloadableProgram.h:43-59, GUI AppWindow.cpp:529-531).  Results do not depend on the shape. */
enum SkyKernelId { SKY_KERNEL_K16 = 0 };
enum SkyK16Shape {
    SKY_K16_AUTO = 0,       /* chosen per launch from the number of rays (cloud.cu) */
    SKY_K16_WAVE_8x4 = 1,   /* ray-group wavefront, 8 rays x 4 look-ahead steps per warp: the throughput shape */
    SKY_K16_WAVE_4x8 = 2,   /* 4 rays x 8 steps: the latency shape (a rank's bands of a sharded frame) */
    SKY_K16_LITERAL = 3     /* one lane = one ray, the shader's loop (what the strict objects and the counting variant always run) */
};

/* Where the row bands of a tile-sharded frame's target go (sky_set_output_gather) */
enum SkyOutputGather { SKY_GATHER_OFF = 0, SKY_GATHER_ALL = 1, SKY_GATHER_ROOT = 2 };

/* Collision sampling of the path tracer (sky_pt_set_tracking) */
enum SkyPtTracking {
    SKY_PT_TRACKING_REFERENCE = 0,     /* the reference's global majorant kSigmaTMax: identical random streams (default) */
    SKY_PT_TRACKING_MAJORANT_GRID = 1  /* local majorants on a macro-cell grid: same expectation, different streams */
};

/* src/SkyRendering/Earth.cpp:12-21, shaders/SkyRendering/EarthRender.frag:6-15: uniforms of the ground pass (176 B) */
typedef struct SkyEarthBufferData {
    float view_projection[16];
    float inv_view_projection[16];
    float camera_position[3];
    float camera_earth_center_distance;
    float earth_center[3];
    float padding;
    float up_direction[3];
    float padding1;
} SkyEarthBufferData;

/* HDRBufferParams subset used by the tone-map pass (src/Base/include/HDRBuffer.h:11-22), see sky_tonemap */
typedef struct SkyToneMapParams {
    int32_t tone_mapping;  /* 0 = CEToneMapping, 1 = ACESToneMapping (default of the reference) */
    float exposure;        /* reference default 10 */
    int32_t dither;        /* dither_color_enable */
    int32_t _pad;
} SkyToneMapParams;

/* VolumetricCloud::PathTracing::InitParam (VolumetricCloud.h:158-168) plus the compile-time
 * constants its constructor bakes into the shader text (VolumetricCloud.cpp:505-519). */
enum SkyPrng { SKY_PRNG_WANG = 0, SKY_PRNG_PCG = 1 };
enum SkyEnvLight {
    SKY_ENV_OFF = 0,
    SKY_ENV_CONST_ENVIRONMENT_MAP = 1,
    SKY_ENV_GROUND_SINGLE_BOUNCE = 2,
    SKY_ENV_GROUND_MULTI_BOUNCE = 3
};

typedef struct SkyPathTracingInit {
    int32_t sqrt_tile_count;
    int32_t max_bounces;
    float region_box_half_width;
    int32_t importance_sampling;
    float forward_phase_g;
    float back_phase_g;
    float forward_scattering_ratio;
    int32_t prng;                 /* SkyPrng */
    int32_t environment_lighting; /* SkyEnvLight */
    float sigma_t_max;            /* material->GetSigmaTMax() */
    float model_matrix3[9];       /* column-major upper 3x3 of VolumetricCloud::model_ */
    int32_t _pad;
} SkyPathTracingInit;

/* Identifiers for sky_get_resource / sky_read_resource.  Layouts are row-major, x fastest,
 * origin = GL texel (0,0); channel counts and element types as the reference allocates them. */
enum SkyResource {
    SKY_RES_TRANSMITTANCE = 0,        /* float4 [64][256]          Atmosphere.cpp:9-15          */
    SKY_RES_MULTISCATTERING = 1,      /* float4 [32][32]           Atmosphere.cpp:17-19         */
    SKY_RES_SKY_VIEW_LUMINANCE = 2,   /* float4 [H][W]             AtmosphereRenderer.cpp:15-17 */
    SKY_RES_SKY_VIEW_TRANSMITTANCE = 3,
    SKY_RES_AERIAL_LUMINANCE = 4,     /* float4 [D][32][32]        AtmosphereRenderer.cpp:19-21 */
    SKY_RES_AERIAL_TRANSMITTANCE = 5,
    SKY_RES_ENVIRONMENT = 6,          /* half4  [6][S][S]          AtmosphereRenderer.cpp:154   */
    SKY_RES_CLOUD_MAP = 7,            /* u8x2   [512][512] mip 0   VolumetricCloudDefaultMaterial.cpp:32-43 */
    SKY_RES_DETAIL = 8,               /* u8     [128][128][128]    :46-57 */
    SKY_RES_DISPLACEMENT = 9,         /* u8x4   [128][128]         :60-71 */
    SKY_RES_SHADOW_MAP_RAW = 10,      /* float2 [512][512]  shadow_maps_[0] after K11 */
    SKY_RES_SHADOW_MAP = 11,          /* float2 [512][512]  shadow_maps_[2] after K12 */
    SKY_RES_SHADOW_FROXEL = 12,       /* u16    [128][H/12][W/12]  VolumetricCloud.cpp:134-135  */
    SKY_RES_CHECKERBOARD_DEPTH = 13,  /* float  [H/2][W/2]         :121-122 */
    SKY_RES_INDEX_LINEAR_DEPTH = 14,  /* float2 [H/4][W/4]         :123-124 */
    SKY_RES_CLOUD_RENDER = 15,        /* half4  [H/4][W/4]         :125-126 */
    SKY_RES_CLOUD_DISTANCE = 16,      /* float  [H/4][W/4]         :127-128 */
    SKY_RES_RECONSTRUCT = 17,         /* half4  [H/2][W/2]  newest reconstruct_texture_ */
    SKY_RES_PT_ACCUM = 18,            /* float4 [H][W]             :498-501 */
    SKY_RES_PT_MASK = 19,             /* u8     [H][W]             :502-503 */
    SKY_RES_VOXEL = 20,               /* u8     [dz][dy][dx] mip 0 VolumetricCloudVoxelMaterial.cpp:72-74 */
    SKY_RES_CLOUD_MAP_MIPS = 21,      /* all mips >= 1, concatenated, same element type */
    SKY_RES_DETAIL_MIPS = 22,
    SKY_RES_DISPLACEMENT_MIPS = 23,
    SKY_RES_VOXEL_MIPS = 24,
    SKY_RES_COUNTERS = 25,            /* uint64 [8]: work counters, see sky_counters */
    SKY_RES_MESH_SHADOW_MAP = 26,     /* float  [2048][2048] light-space depth of the mesh shadow pass (ShadowMap.cpp:8-27,
                                         AppWindow.cpp:25,183-190), cleared to 1.0; an INPUT: written by the caller with
                                         sky_write_resource, read by K3/K4/K6 when volumetric_light is set */
    SKY_RES_ENV_BRDF_LUT = 27,        /* u16x2  [512][512] GL_RG16 environment-BRDF LUT        src/Base/src/Textures.cpp:60-75 (K22) */
    SKY_RES_ENVIRONMENT_MIPS = 28,    /* half4  levels >= 1 of SKY_RES_ENVIRONMENT, concatenated ([6][S/2][S/2], [6][S/4][S/4], ...)
                                         glGenerateTextureMipmap, AtmosphereRenderer.cpp:242 */
    SKY_RES_ENV_RADIANCE_SH = 29,     /* float4 [9] Llm, the SH9 coefficients of the environment   src/Base/src/IBL.cpp:12-13,29-34 (K23) */
    SKY_RES_PREFILTERED_RADIANCE = 30,/* half4  5 levels concatenated ([6][128][128], [6][64][64], ... [6][8][8]), level i filtered at
                                         roughness i / 4                                          src/Base/src/IBL.cpp:21-22,35-42 (K24) */
    SKY_RES_EARTH_ALBEDO = 31,        /* u8x4   GL_SRGB8 codes (RGBX) of the earth albedo map, ALL levels concatenated (level l =
                                         [max(H >> l, 1)][max(W >> l, 1)]): the upload of sky_set_earth_albedo + glGenerateTextureMipmap
                                         src/Base/src/Textures.cpp:52-58 */
    SKY_RES_FRAME_HDR = 32,           /* half4  [H][W] the context-owned frame target of a tile-sharded frame (sky_peer_export allocates it,
                                         sky_set_output_gather): pass its pointer as hdr_dev */
    SKY_RES_COUNT_
};

/* IBL constants (src/Base/include/IBL.h:10-11, src/Base/src/Textures.cpp:61-62) */
#define SKY_IBL_PREFILTERED_RESOLUTION 128
#define SKY_IBL_ROUGHNESS_COUNT 5
#define SKY_ENV_BRDF_LUT_SIZE 512

enum SkyFormat {
    SKY_FMT_F32 = 0, SKY_FMT_F16 = 1, SKY_FMT_U8 = 2, SKY_FMT_U16 = 3, SKY_FMT_U64 = 4
};

typedef struct SkyResourceDesc {
    void* ptr;          /* borrowed; device pointer for libskyb200, host pointer for the oracle */
    int32_t width, height, depth, channels;
    int32_t format;     /* SkyFormat */
    int32_t _pad;
    uint64_t bytes;
} SkyResourceDesc;

/* Slots of SKY_RES_COUNTERS (incremented only when counting is enabled). */
enum SkyCounter {
    SKY_CNT_RENDER_SIGMA_EVALS = 0,  /* SampleSigmaT calls made by K16                */
    SKY_CNT_RENDER_TEX_FETCHES = 1,  /* texture fetches those calls issued            */
    SKY_CNT_PT_PATHS = 2,            /* pixel-samples traced by K19                   */
    SKY_CNT_PT_LOOKUPS = 3,          /* SampleSigmaT calls made by K19                */
    SKY_CNT_PT_COLLISIONS = 4,       /* tentative collisions incl. the provably-empty */
    SKY_CNT_SHADOW_SIGMA_EVALS = 5   /* SampleSigmaT calls made by K11                */
};

#ifdef __cplusplus
}
#endif
#endif /* SKY_TYPES_H */
