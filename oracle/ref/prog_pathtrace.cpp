// ORACLE -- TEST INFRASTRUCTURE ONLY.  K19: the reference's path tracer (VolumetricCloudPathTracing.comp + the voxel
// material, composed by VolumetricCloud.cpp:483-531).  The constants the host writes into the shader text as #defines
// are run-time variables here; the #if / #ifdef permutations are compiled as separate namespaces.
#include "ref_common.h"

namespace ref {
struct PtConstants { float sigma_t_max, cloud_half_width, forward_g, back_g, forward_ratio; int max_bounces; mat3 model3; };
PtConstants g_pt;
}
#define kSigmaTMax (ref::g_pt.sigma_t_max)
#define kMaxBounces (ref::g_pt.max_bounces)
#define kCloudHalfWidth (ref::g_pt.cloud_half_width)
#define kForwardPhaseG (ref::g_pt.forward_g)
#define kBackPhaseG (ref::g_pt.back_g)
#define kForwardScatteringRatio (ref::g_pt.forward_ratio)
#define kModelMatrix3 (ref::g_pt.model3)
#define MATERIAL_TEXTURE_UNIT_BEGIN 7
#define LOCAL_SIZE_X 8
#define LOCAL_SIZE_Y 4

namespace ref { namespace pt_pcg_multi {
#define PRNG PCGHash
#define IMPORTANCE_SAMPLING 1
#define ENVIRONMENT_LIGHT_GROUND_MULTI_BOUNCE
#include "../_ref/gen/VolumetricCloudPathTracing.comp.inc"
#include "../_ref/gen/VolumetricCloudMaterialVoxel.glsl.inc"
#undef ENVIRONMENT_LIGHT_GROUND_MULTI_BOUNCE
#undef PRNG
#undef IMPORTANCE_SAMPLING
#include "ref_undef_guards.h"
} namespace pt_wang_single {
#define PRNG WangHash
#define IMPORTANCE_SAMPLING 1
#define ENVIRONMENT_LIGHT_GROUND_SINGLE_BOUNCE
#include "../_ref/gen/VolumetricCloudPathTracing.comp.inc"
#include "../_ref/gen/VolumetricCloudMaterialVoxel.glsl.inc"
#undef ENVIRONMENT_LIGHT_GROUND_SINGLE_BOUNCE
#undef PRNG
#undef IMPORTANCE_SAMPLING
#include "ref_undef_guards.h"
} namespace pt_pcg_envmap {
#define PRNG PCGHash
#define IMPORTANCE_SAMPLING 1
#define ENVIRONMENT_LIGHT_CONST_ENVIRONMENT_MAP
#include "../_ref/gen/VolumetricCloudPathTracing.comp.inc"
#include "../_ref/gen/VolumetricCloudMaterialVoxel.glsl.inc"
#undef ENVIRONMENT_LIGHT_CONST_ENVIRONMENT_MAP
#undef PRNG
#undef IMPORTANCE_SAMPLING
#include "ref_undef_guards.h"
} namespace pt_pcg_off_uniform {
#define PRNG PCGHash
#define IMPORTANCE_SAMPLING 0
#define ENVIRONMENT_LIGHT_OFF
#include "../_ref/gen/VolumetricCloudPathTracing.comp.inc"
#include "../_ref/gen/VolumetricCloudMaterialVoxel.glsl.inc"
#undef ENVIRONMENT_LIGHT_OFF
#undef PRNG
#undef IMPORTANCE_SAMPLING
} }

struct RefVoxelLevel { const float* rgba; int w, h, d; };   // texels as floats (value / 255 in .x)
struct RefPtIO {
    const float* transmittance;    // [64][256][4]
    const float* ap_luminance;     // [D][32][32][4]
    const float* ap_transmittance;
    int ap_depth;
    const float* froxel;           // [fd][fh][fw][4], unorm16 / 65535 in .x
    int fw, fh, fd;
    const float* environment;      // [6][E][E][4]
    int env_size;
    const RefVoxelLevel* voxel_levels;
    int voxel_level_count;
    float* accum;                  // [H][W][4], read-modify-write
    float* mask;                   // [H][W][4] scratch for the r8ui image
    float* display;                // [H][W][4] scratch: only its size is read by the render pass
    int width, height;
};

#define REF_LOAD_COMMON(c)                                                                                  \
    do {                                                                                                    \
        uInvMVP = ref::mat4((c)->uInvMVP); uReprojectMat = ref::mat4((c)->uReprojectMat); uLightVP = ref::mat4((c)->uLightVP); \
        uInvLightVP = ref::mat4((c)->uInvLightVP); uShadowMapReprojectMat = ref::mat4((c)->uShadowMapReprojectMat); \
        uCameraPos = REF_V3((c)->uCameraPos); uBaseShadingIndex = (c)->uBaseShadingIndex;                   \
        uLinearDepthParam = ref::vec2((c)->uLinearDepthParam[0], (c)->uLinearDepthParam[1]);                \
        uBottomAltitude = (c)->uBottomAltitude; uTopAltitude = (c)->uTopAltitude;                           \
        uSunDirection = REF_V3((c)->uSunDirection); uFrameID = (c)->uFrameID;                               \
        uInvShadowFroxelMaxDistance = (c)->uInvShadowFroxelMaxDistance;                                     \
        uAerialPerspectiveLutMaxDistance = (c)->uAerialPerspectiveLutMaxDistance;                           \
        uShadowFroxelMaxDistance = (c)->uShadowFroxelMaxDistance; uEarthRadius = (c)->uEarthRadius;         \
    } while (0)

#define RUN_PT(NS)                                                                                                  \
    {                                                                                                               \
        using namespace ref::NS;                                                                                    \
        REF_LOAD_ATMOSPHERE(a);                                                                                     \
        REF_LOAD_COMMON(c);                                                                                         \
        uSampleFrequency = ref::vec2(m->uSampleFrequency[0], m->uSampleFrequency[1]); uLodBias = m->uLodBias;       \
        uDensity = m->uDensity; uSampleBias = ref::vec2(m->uSampleBias[0], m->uSampleBias[1]); uSampleLodK = m->uSampleLodK; \
        ref_bind_texture(transmittance_texture, io->transmittance, 256, 64, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);    \
        ref_bind_texture(aerial_perspective_luminance_texture, io->ap_luminance, 32, 32, io->ap_depth, ref::CLAMP_TO_EDGE, ref::LINEAR); \
        ref_bind_texture(aerial_perspective_transmittance_texture, io->ap_transmittance, 32, 32, io->ap_depth, ref::CLAMP_TO_EDGE, ref::LINEAR); \
        ref_bind_texture(shadow_froxel, io->froxel, io->fw, io->fh, io->fd, ref::CLAMP_TO_EDGE, ref::LINEAR);       \
        ref_bind_texture(environment_luminance_texture, io->environment, io->env_size, io->env_size, 6, ref::CLAMP_TO_EDGE, ref::LINEAR); \
        /* VolumetricCloudVoxelMaterial.cpp:30-37: CLAMP_TO_BORDER (0), mag LINEAR, min NEAREST_MIPMAP_NEAREST */  \
        voxel.levels.clear();                                                                                       \
        for (int l = 0; l < io->voxel_level_count; ++l) {                                                           \
            ref::Image im; im.data = const_cast<float*>(io->voxel_levels[l].rgba);                                  \
            im.w = io->voxel_levels[l].w; im.h = io->voxel_levels[l].h; im.d = io->voxel_levels[l].d;               \
            voxel.levels.push_back(im);                                                                             \
        }                                                                                                           \
        voxel.wrap = ref::CLAMP_TO_BORDER; voxel.mag = ref::LINEAR; voxel.min_filter = ref::NEAREST; voxel.border = ref::vec4(0.0f); \
        ref_bind_image(accumulating_image, io->accum, io->width, io->height, 1, ref::FMT_RGBA32F);                  \
        ref_bind_image(rendered_mask_image, io->mask, io->width, io->height, 1, ref::FMT_RGBA32F);                  \
        ref_bind_image(display_image, io->display, io->width, io->height, 1, ref::FMT_RGBA16F);                     \
        kRenderRegion = ref::ivec4(region[0], region[1], region[2], region[3]);                                     \
        for (uint32_t f = 0; f < count; ++f) {                                                                      \
            kFrameId = frame_begin + f;  /* glUniform1ui(0, frame_cnt_), VolumetricCloud.cpp:555 */                \
            ref::dispatch(main, ref_ceil_div(region[2], LOCAL_SIZE_X), ref_ceil_div(region[3], LOCAL_SIZE_Y), 1, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false); \
        }                                                                                                           \
    }

extern "C" int ref_pt_samples(const SkyAtmosphereBufferData* a, const SkyCloudCommonBufferData* c, const SkyMaterialVoxelBufferData* m,
                              const SkyPathTracingInit* init, const RefPtIO* io, uint32_t frame_begin, uint32_t count, const int32_t* region) {
    ref::g_pt.sigma_t_max = init->sigma_t_max; ref::g_pt.cloud_half_width = init->region_box_half_width;
    ref::g_pt.forward_g = init->forward_phase_g; ref::g_pt.back_g = init->back_phase_g; ref::g_pt.forward_ratio = init->forward_scattering_ratio;
    ref::g_pt.max_bounces = init->max_bounces; ref::g_pt.model3 = ref::mat3(init->model_matrix3);
    const int prng = init->prng, env = init->environment_lighting, is = init->importance_sampling;
    if (prng == SKY_PRNG_PCG && env == SKY_ENV_GROUND_MULTI_BOUNCE && is) RUN_PT(pt_pcg_multi)
    else if (prng == SKY_PRNG_WANG && env == SKY_ENV_GROUND_SINGLE_BOUNCE && is) RUN_PT(pt_wang_single)
    else if (prng == SKY_PRNG_PCG && env == SKY_ENV_CONST_ENVIRONMENT_MAP && is) RUN_PT(pt_pcg_envmap)
    else if (prng == SKY_PRNG_PCG && env == SKY_ENV_OFF && !is) RUN_PT(pt_pcg_off_uniform)
    else return 2;  // permutation not compiled
    return 0;
}
