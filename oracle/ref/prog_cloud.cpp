// ORACLE -- TEST INFRASTRUCTURE ONLY.  K11-K18: the reference's cloud shadow chain and real-time cloud chain
// (VolumetricCloud{ShadowMap,ShadowMapBlur,ShadowFroxel,IndexGen,Render,Reconstruct,Upscale}.comp, CheckerboardGen.comp,
// plus the material appended by VolumetricCloud::CreateShaderPostProcess, VolumetricCloud.cpp:483-493), one pass per call
// with every input supplied by the caller, so each program is compared with the oracle on identical inputs.
#include "ref_common.h"
#define MATERIAL_TEXTURE_UNIT_BEGIN 7
#define LOCAL_SIZE_X 8
#define LOCAL_SIZE_Y 4

#define REF_MATERIAL_NS(NS, PASS_INC, MATERIAL_INC) \
    namespace ref { namespace NS {                  \
    }}

namespace ref {
namespace k11_m0 {
#include "../_ref/gen/VolumetricCloudShadowMap.comp.inc"
#include "../_ref/gen/VolumetricCloudDefaultMaterial0.glsl.inc"
#include "ref_undef_guards.h"
} namespace k11_m1 {
#include "../_ref/gen/VolumetricCloudShadowMap.comp.inc"
#include "../_ref/gen/VolumetricCloudDefaultMaterial1.glsl.inc"
#include "ref_undef_guards.h"
} namespace k11_vox {
#include "../_ref/gen/VolumetricCloudShadowMap.comp.inc"
#include "../_ref/gen/VolumetricCloudMaterialVoxel.glsl.inc"
#include "ref_undef_guards.h"
} namespace k12a {
#define VOLUMETRIC_CLOUD_SHADOW_MAP_BLUR_PASS0
#include "../_ref/gen/VolumetricCloudShadowMapBlur.comp.inc"
#undef VOLUMETRIC_CLOUD_SHADOW_MAP_BLUR_PASS0
#include "ref_undef_guards.h"
} namespace k12b {
#define VOLUMETRIC_CLOUD_SHADOW_MAP_BLUR_PASS1
#include "../_ref/gen/VolumetricCloudShadowMapBlur.comp.inc"
#undef VOLUMETRIC_CLOUD_SHADOW_MAP_BLUR_PASS1
#include "ref_undef_guards.h"
} namespace k13 {
#include "../_ref/gen/VolumetricCloudShadowFroxel.comp.inc"
#include "ref_undef_guards.h"
} namespace k14 {
#include "../_ref/gen/CheckerboardGen.comp.inc"
#include "ref_undef_guards.h"
} namespace k15 {
#include "../_ref/gen/VolumetricCloudIndexGen.comp.inc"
#include "ref_undef_guards.h"
} namespace k16_m0 {
#include "../_ref/gen/VolumetricCloudRender.comp.inc"
#include "../_ref/gen/VolumetricCloudDefaultMaterial0.glsl.inc"
#include "ref_undef_guards.h"
} namespace k16_m1 {
#include "../_ref/gen/VolumetricCloudRender.comp.inc"
#include "../_ref/gen/VolumetricCloudDefaultMaterial1.glsl.inc"
#include "ref_undef_guards.h"
} namespace k16_vox {
#include "../_ref/gen/VolumetricCloudRender.comp.inc"
#include "../_ref/gen/VolumetricCloudMaterialVoxel.glsl.inc"
#include "ref_undef_guards.h"
} namespace k17 {
#include "../_ref/gen/VolumetricCloudReconstruct.comp.inc"
#include "ref_undef_guards.h"
} namespace k18 {
#include "../_ref/gen/VolumetricCloudUpscale.comp.inc"
} }

struct RefTex { const float* rgba; int w, h, d; };  // one level, texels as float RGBA
struct RefCloudIO {
    const SkyAtmosphereBufferData* atm;
    const SkyCloudCommonBufferData* common;
    const SkyCloudBufferData* cloud;
    const SkyMaterialBlock* material;
    const RefTex* cloud_map; int cloud_map_levels;
    const RefTex* detail; int detail_levels;
    const RefTex* displacement; int displacement_levels;
    const RefTex* voxel; int voxel_levels;
    float* blue_noise;                                        // [64][64][4]
    float *shadow_prev, *shadow_raw, *shadow_tmp, *shadow_blurred; int shadow_size;   // [S][S][4]
    float* froxel; int fw, fh, fd;                            // [fd][fh][fw][4]
    float* depth; int width, height;                          // [H][W][4]
    float *checkerboard, *index_linear, *render, *cloud_distance, *reconstruct_prev, *reconstruct_out, *hdr;
    float *transmittance, *ap_luminance, *ap_transmittance; int ap_depth;
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
#define REF_LOAD_CLOUD(b)                                                                                   \
    do {                                                                                                    \
        uSunIlluminanceScale = (b)->uSunIlluminanceScale; uMaxRaymarchDistance = (b)->uMaxRaymarchDistance; \
        uMaxRaymarchSteps = (b)->uMaxRaymarchSteps; uMaxVisibleDistance = (b)->uMaxVisibleDistance;         \
        uEnvColorScale = REF_V3((b)->uEnvColorScale); uShadowSteps = (b)->uShadowSteps;                     \
        uSunMultiscatteringSigmaScale = (b)->uSunMultiscatteringSigmaScale;                                 \
        uEnvMultiscatteringSigmaScale = (b)->uEnvMultiscatteringSigmaScale; uShadowDistance = (b)->uShadowDistance; \
        uEnvBottomVisibility = (b)->uEnvBottomVisibility; uEnvSunHeightCurveExp = (b)->uEnvSunHeightCurveExp; \
    } while (0)

template <class S>
static void bind_mips(S& s, const RefTex* levels, int n, ref::Wrap wrap) {
    s.levels.clear();
    for (int l = 0; l < n; ++l) {
        ref::Image im; im.data = const_cast<float*>(levels[l].rgba); im.w = levels[l].w; im.h = levels[l].h; im.d = levels[l].d;
        s.levels.push_back(im);
    }
    // mag LINEAR, min NEAREST_MIPMAP_NEAREST (VolumetricCloudDefaultMaterial.cpp:111-116, VolumetricCloudVoxelMaterial.cpp:30-37)
    s.wrap = wrap; s.mag = ref::LINEAR; s.min_filter = ref::NEAREST; s.border = ref::vec4(0.0f);
}
#define SAMPLE_INFO(dst, src) do { dst.bias = ref::vec2((src).bias[0], (src).bias[1]); dst.frequency = (src).frequency; dst.k_lod = (src).k_lod; } while (0)
#define REF_BIND_DEFAULT_COMMON()                                                                           \
    do {                                                                                                    \
        const SkyMaterialCommonBufferData& mc = io->material->common;                                       \
        SAMPLE_INFO(uCloudMapSampleInfo, mc.uCloudMapSampleInfo); SAMPLE_INFO(uDetailSampleInfo, mc.uDetailSampleInfo); \
        SAMPLE_INFO(uDisplacementSampleInfo, mc.uDisplacementSampleInfo); uLodBias = mc.uLodBias; uDensity = mc.uDensity; \
        bind_mips(cloud_map, io->cloud_map, io->cloud_map_levels, ref::REPEAT);                             \
        bind_mips(detail_texture, io->detail, io->detail_levels, ref::REPEAT);                              \
        bind_mips(displacement_texture, io->displacement, io->displacement_levels, ref::REPEAT);            \
    } while (0)
#define REF_BIND_M0() do { REF_BIND_DEFAULT_COMMON(); const SkyMaterial0BufferData& m = io->material->u.m0;  \
        uDetailParam = ref::vec2(m.uDetailParam[0], m.uDetailParam[1]); uDisplacementScale = m.uDisplacementScale; } while (0)
#define REF_BIND_M1() do { REF_BIND_DEFAULT_COMMON(); const SkyMaterial1BufferData& m = io->material->u.m1;  \
        uBaseDensityThreshold = m.uBaseDensityThreshold; uBaseHeightHardness = m.uBaseHeightHardness; uBaseEdgeHardness = m.uBaseEdgeHardness; \
        uDetailBase = m.uDetailBase; uDetailScale = m.uDetailScale; uHeightCut = m.uHeightCut; uEdgeCur = m.uEdgeCur; } while (0)
#define REF_BIND_VOX() do { const SkyMaterialVoxelBufferData& m = io->material->u.voxel;                     \
        uSampleFrequency = ref::vec2(m.uSampleFrequency[0], m.uSampleFrequency[1]); uLodBias = m.uLodBias; uDensity = m.uDensity; \
        uSampleBias = ref::vec2(m.uSampleBias[0], m.uSampleBias[1]); uSampleLodK = m.uSampleLodK;             \
        bind_mips(voxel, io->voxel, io->voxel_levels, ref::CLAMP_TO_BORDER); } while (0)

// the cloud shadow map sampler: LINEAR, CLAMP_TO_BORDER (1e10, 1, 0, 0) (VolumetricCloud.cpp:106-112)
template <class S>
static void bind_shadow_sampler(S& s, const float* data, int n) {
    ref_bind_texture(s, data, n, n, 1, ref::CLAMP_TO_BORDER, ref::LINEAR);
    s.border = ref::vec4(1e10f, 1.0f, 0.0f, 0.0f);
}

#define RUN_K11(NS, BIND)                                                                                   \
    { using namespace ref::NS; REF_LOAD_COMMON(io->common); BIND();                                          \
      bind_shadow_sampler(pre_raw_cloud_shadow_map, io->shadow_prev, S);                                     \
      ref_bind_texture(blue_noise, io->blue_noise, 64, 64, 1, ref::REPEAT, ref::NEAREST);                    \
      ref_bind_image(raw_shadow_map_image, io->shadow_raw, S, S, 1, ref::FMT_RG32F);                         \
      ref::dispatch(main, ref_ceil_div(S, LOCAL_SIZE_X), ref_ceil_div(S, LOCAL_SIZE_Y), 1, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false); }
#define RUN_K16(NS, BIND)                                                                                   \
    { using namespace ref::NS; REF_LOAD_ATMOSPHERE(io->atm); REF_LOAD_COMMON(io->common); REF_LOAD_CLOUD(io->cloud); BIND(); \
      ref_bind_texture(checkerboard_depth, io->checkerboard, W / 2, H / 2, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);   \
      ref_bind_texture(index_linear_depth_texture, io->index_linear, W / 4, H / 4, 1, ref::CLAMP_TO_EDGE, ref::NEAREST); \
      ref_bind_texture(transmittance_texture, io->transmittance, 256, 64, 1, ref::CLAMP_TO_EDGE, ref::LINEAR); \
      ref_bind_texture(aerial_perspective_luminance_texture, io->ap_luminance, 32, 32, io->ap_depth, ref::CLAMP_TO_EDGE, ref::LINEAR); \
      ref_bind_texture(aerial_perspective_transmittance_texture, io->ap_transmittance, 32, 32, io->ap_depth, ref::CLAMP_TO_EDGE, ref::LINEAR); \
      ref_bind_texture(blue_noise, io->blue_noise, 64, 64, 1, ref::REPEAT, ref::NEAREST);                    \
      ref_bind_texture(shadow_froxel, io->froxel, io->fw, io->fh, io->fd, ref::CLAMP_TO_EDGE, ref::LINEAR);  \
      ref_bind_image(render_image, io->render, W / 4, H / 4, 1, ref::FMT_RGBA16F);                           \
      ref_bind_image(cloud_distance_image, io->cloud_distance, W / 4, H / 4, 1, ref::FMT_R32F);              \
      ref::dispatch(main, ref_ceil_div(W / 4, LOCAL_SIZE_X), ref_ceil_div(H / 4, LOCAL_SIZE_Y), 1, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false); }

extern "C" int ref_cloud_pass(int pass, const RefCloudIO* io) {
    const int S = io->shadow_size, W = io->width, H = io->height;
    const int mat = io->material ? io->material->type : -1;
    switch (pass) {
        case 11:
            if (mat == SKY_MATERIAL_DEFAULT0) RUN_K11(k11_m0, REF_BIND_M0)
            else if (mat == SKY_MATERIAL_DEFAULT1) RUN_K11(k11_m1, REF_BIND_M1)
            else if (mat == SKY_MATERIAL_VOXEL) RUN_K11(k11_vox, REF_BIND_VOX)
            else return 2;
            return 0;
        case 12: {
            { using namespace ref::k12a; bind_shadow_sampler(in_texture, io->shadow_raw, S); ref_bind_image(out_image, io->shadow_tmp, S, S, 1, ref::FMT_RG32F);
              ref::dispatch(main, ref_ceil_div(S, LOCAL_SIZE_X), ref_ceil_div(S, LOCAL_SIZE_Y), 1, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false); }
            { using namespace ref::k12b; bind_shadow_sampler(in_texture, io->shadow_tmp, S); ref_bind_image(out_image, io->shadow_blurred, S, S, 1, ref::FMT_RG32F);
              ref::dispatch(main, ref_ceil_div(S, LOCAL_SIZE_X), ref_ceil_div(S, LOCAL_SIZE_Y), 1, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false); }
            return 0;
        }
        case 13: {
            using namespace ref::k13; REF_LOAD_COMMON(io->common);
            bind_shadow_sampler(cloud_shadow_map, io->shadow_blurred, S);
            ref_bind_image(shadow_froxel_image, io->froxel, io->fw, io->fh, io->fd, ref::FMT_R16);
            ref::dispatch(main, ref_ceil_div(io->fw, LOCAL_SIZE_X), ref_ceil_div(io->fh, LOCAL_SIZE_Y), 1, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false);
            return 0;
        }
        case 14: {
            using namespace ref::k14;
            ref_bind_texture(depth_texture, io->depth, W, H, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);
            ref_bind_image(checkerboard_depth, io->checkerboard, W / 2, H / 2, 1, ref::FMT_R32F);
            ref::dispatch(main, ref_ceil_div(W / 2, LOCAL_SIZE_X), ref_ceil_div(H / 2, LOCAL_SIZE_Y), 1, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false);
            return 0;
        }
        case 15: {
            using namespace ref::k15; REF_LOAD_COMMON(io->common);
            ref_bind_texture(checkerboard_depth, io->checkerboard, W / 2, H / 2, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);
            ref_bind_image(index_linear_depth, io->index_linear, W / 4, H / 4, 1, ref::FMT_RG32F);
            ref::dispatch(main, ref_ceil_div(W / 4, LOCAL_SIZE_X), ref_ceil_div(H / 4, LOCAL_SIZE_Y), 1, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false);
            return 0;
        }
        case 16:
            if (mat == SKY_MATERIAL_DEFAULT0) RUN_K16(k16_m0, REF_BIND_M0)
            else if (mat == SKY_MATERIAL_DEFAULT1) RUN_K16(k16_m1, REF_BIND_M1)
            else if (mat == SKY_MATERIAL_VOXEL) RUN_K16(k16_vox, REF_BIND_VOX)
            else return 2;
            return 0;
        case 17: {
            using namespace ref::k17; REF_LOAD_COMMON(io->common);
            ref_bind_texture(checkerboard_depth, io->checkerboard, W / 2, H / 2, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);
            ref_bind_texture(index_linear_depth_texture, io->index_linear, W / 4, H / 4, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);
            ref_bind_texture(render_texture, io->render, W / 4, H / 4, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);
            ref_bind_texture(cloud_distance_texture, io->cloud_distance, W / 4, H / 4, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);
            ref_bind_texture(preframe_reconstruct_texture, io->reconstruct_prev, W / 2, H / 2, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);
            ref_bind_image(reconstruct_image, io->reconstruct_out, W / 2, H / 2, 1, ref::FMT_RGBA16F);
            ref::dispatch(main, ref_ceil_div(W / 2, LOCAL_SIZE_X), ref_ceil_div(H / 2, LOCAL_SIZE_Y), 1, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false);
            return 0;
        }
        case 18: {
            using namespace ref::k18; REF_LOAD_COMMON(io->common);
            ref_bind_texture(checkerboard_depth, io->checkerboard, W / 2, H / 2, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);
            ref_bind_texture(depth_texture, io->depth, W, H, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);
            ref_bind_texture(recontruct_texture, io->reconstruct_out, W / 2, H / 2, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);
            ref_bind_image(hdr_image, io->hdr, W, H, 1, ref::FMT_RGBA16F);
            ref::dispatch(main, ref_ceil_div(W, LOCAL_SIZE_X), ref_ceil_div(H, LOCAL_SIZE_Y), 1, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false);
            return 0;
        }
    }
    return 1;
}
