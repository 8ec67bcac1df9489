// ORACLE -- TEST INFRASTRUCTURE ONLY.  K3 / K4 / K5: the reference's sky-view, aerial-perspective and environment-cube
// programs (AtmosphereRenderer.glsl, composed by AtmosphereRenderer.cpp:91-148; dispatched by :219-239).  The header
// the host prepends is reproduced as #defines; the two sizes it bakes in as text become run-time variables.
#define REF_MATH_DET
#include "ref_common.h"

namespace ref { int g_sky_w = 128, g_sky_h = 128, g_ap_depth = 32; }
#define SKY_VIEW_LUT_SIZE ivec2(ref::g_sky_w, ref::g_sky_h)
#define AERIAL_PERSPECTIVE_LUT_SIZE ivec3(32, 32, ref::g_ap_depth)
#define PCSS_ENABLE 0
#define VOLUMETRIC_LIGHT_ENABLE 0
#define MOON_SHADOW_ENABLE 0
#define USE_SKY_VIEW_LUT 1
#define USE_AERIAL_PERSPECTIVE_LUT 1
#define ROUGHNESS_COUNT 5
#define LOCAL_SIZE_X 8
#define LOCAL_SIZE_Y 4
#define LOCAL_SIZE_Z 1

#define REF_LOAD_RENDER(r)                                                                                   \
    do {                                                                                                     \
        sun_direction = REF_V3((r)->sun_direction); star_luminance_scale = (r)->star_luminance_scale;         \
        earth_center = REF_V3((r)->earth_center); camera_earth_center_distance = (r)->camera_earth_center_distance; \
        camera_position = REF_V3((r)->camera_position); raymarching_steps = (r)->raymarching_steps;           \
        up_direction = REF_V3((r)->up_direction); sky_view_lut_steps = (r)->sky_view_lut_steps;               \
        right_direction = REF_V3((r)->right_direction); aerial_perspective_lut_steps = (r)->aerial_perspective_lut_steps; \
        front_direction = REF_V3((r)->front_direction);                                                       \
        aerial_perspective_lut_max_distance = (r)->aerial_perspective_lut_max_distance;                       \
        moon_position = REF_V3((r)->moon_position); moon_radius = (r)->moon_radius;                           \
        inv_view_projection = ref::mat4((r)->inv_view_projection);                                            \
        light_view_projection = ref::mat4((r)->light_view_projection);                                        \
        uInvShadowFroxelMaxDistance = (r)->uInvShadowFroxelMaxDistance;                                       \
        blocker_kernel_size_k = (r)->blocker_kernel_size_k; pcss_size_k = (r)->pcss_size_k;                   \
        uCloudShadowMapMat = ref::mat4((r)->uCloudShadowMapMat);                                              \
    } while (0)

#define REF_PROGRAM(NS, PROGRAM, DITHER)                    \
    namespace ref { namespace NS {                          \
    _Pragma("push_macro(\"DITHER_SAMPLE_POINT_ENABLE\")")   \
    } }
// (the shader text is included once per (program, dither) pair; include guards are reset in between)
#define REF_RESET_GUARDS

namespace ref { namespace k3d0 {
#define SKY_VIEW_COMPUTE_PROGRAM
#define DITHER_SAMPLE_POINT_ENABLE 0
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#undef DITHER_SAMPLE_POINT_ENABLE
#include "ref_undef_guards.h"
} namespace k3d1 {
#define DITHER_SAMPLE_POINT_ENABLE 1
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#undef DITHER_SAMPLE_POINT_ENABLE
#undef SKY_VIEW_COMPUTE_PROGRAM
#include "ref_undef_guards.h"
} namespace k4d0 {
#define AERIAL_PERSPECTIVE_COMPUTE_PROGRAM
#define DITHER_SAMPLE_POINT_ENABLE 0
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#undef DITHER_SAMPLE_POINT_ENABLE
#include "ref_undef_guards.h"
} namespace k4d1 {
#define DITHER_SAMPLE_POINT_ENABLE 1
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#undef DITHER_SAMPLE_POINT_ENABLE
#undef AERIAL_PERSPECTIVE_COMPUTE_PROGRAM
#include "ref_undef_guards.h"
}
// the optional march terms (off in the shipped scenes): moon shadow / mesh shadow map / both, no dither
#define REF_PERMUTATION(NS, PROGRAM_MACRO, VOL, MOON)
#undef VOLUMETRIC_LIGHT_ENABLE
#undef MOON_SHADOW_ENABLE
#define DITHER_SAMPLE_POINT_ENABLE 0
#define SKY_VIEW_COMPUTE_PROGRAM
#define VOLUMETRIC_LIGHT_ENABLE 0
#define MOON_SHADOW_ENABLE 1
namespace k3v0m1 {
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#include "ref_undef_guards.h"
}
#undef VOLUMETRIC_LIGHT_ENABLE
#undef MOON_SHADOW_ENABLE
#define VOLUMETRIC_LIGHT_ENABLE 1
#define MOON_SHADOW_ENABLE 0
namespace k3v1m0 {
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#include "ref_undef_guards.h"
}
#undef MOON_SHADOW_ENABLE
#define MOON_SHADOW_ENABLE 1
namespace k3v1m1 {
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#include "ref_undef_guards.h"
}
#undef SKY_VIEW_COMPUTE_PROGRAM
#define AERIAL_PERSPECTIVE_COMPUTE_PROGRAM
namespace k4v1m1 {
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#include "ref_undef_guards.h"
}
#undef VOLUMETRIC_LIGHT_ENABLE
#define VOLUMETRIC_LIGHT_ENABLE 0
namespace k4v0m1 {
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#include "ref_undef_guards.h"
}
#undef VOLUMETRIC_LIGHT_ENABLE
#undef MOON_SHADOW_ENABLE
#define VOLUMETRIC_LIGHT_ENABLE 1
#define MOON_SHADOW_ENABLE 0
namespace k4v1m0 {
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#include "ref_undef_guards.h"
}
#undef AERIAL_PERSPECTIVE_COMPUTE_PROGRAM
#undef VOLUMETRIC_LIGHT_ENABLE
#undef MOON_SHADOW_ENABLE
#undef DITHER_SAMPLE_POINT_ENABLE
#define VOLUMETRIC_LIGHT_ENABLE 0
#define MOON_SHADOW_ENABLE 0
// K6: the full-screen fragment program (AtmosphereRenderer.cpp:150-160), raymarching dither on (all shipped scenes),
// permutations by LUT use: s = USE_SKY_VIEW_LUT, a = USE_AERIAL_PERSPECTIVE_LUT
#define ATMOSPHERE_RENDER_FRAGMENT_SHADER
#define DITHER_SAMPLE_POINT_ENABLE 1
#undef USE_SKY_VIEW_LUT
#undef USE_AERIAL_PERSPECTIVE_LUT
#define USE_SKY_VIEW_LUT 1
#define USE_AERIAL_PERSPECTIVE_LUT 1
namespace k6s1a1 {
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#include "ref_undef_guards.h"
}
#undef MOON_SHADOW_ENABLE
#define MOON_SHADOW_ENABLE 1
namespace k6s1a1m1 {  // scene c1's flags + MOON_SHADOW_ENABLE: the eclipse factor on the object branch's shadow visibility (:406-408)
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#include "ref_undef_guards.h"
}
#undef MOON_SHADOW_ENABLE
#define MOON_SHADOW_ENABLE 0
#undef USE_AERIAL_PERSPECTIVE_LUT
#define USE_AERIAL_PERSPECTIVE_LUT 0
namespace k6s1a0 {
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#include "ref_undef_guards.h"
}
#undef PCSS_ENABLE
#define PCSS_ENABLE 1
namespace k6s1a0p1 {  // scene c3's flags + PCSS_ENABLE (the object branch's soft shadows, Shadow.glsl)
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#include "ref_undef_guards.h"
}
#undef PCSS_ENABLE
#define PCSS_ENABLE 0
#undef USE_SKY_VIEW_LUT
#define USE_SKY_VIEW_LUT 0
namespace k6s0a0 {
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#include "ref_undef_guards.h"
}
#undef DITHER_SAMPLE_POINT_ENABLE
#define DITHER_SAMPLE_POINT_ENABLE 0
namespace k6s0a0d0 {
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#include "ref_undef_guards.h"
}
#undef DITHER_SAMPLE_POINT_ENABLE
#undef ATMOSPHERE_RENDER_FRAGMENT_SHADER
#undef USE_SKY_VIEW_LUT
#undef USE_AERIAL_PERSPECTIVE_LUT
#define USE_SKY_VIEW_LUT 1
#define USE_AERIAL_PERSPECTIVE_LUT 1
namespace k5 {
#define ENVIRONMENT_LUMINANCE_COMPUTE_PROGRAM
#define DITHER_SAMPLE_POINT_ENABLE 1
#include "../_ref/gen/AtmosphereRenderer.glsl.inc"
#undef DITHER_SAMPLE_POINT_ENABLE
#undef ENVIRONMENT_LUMINANCE_COMPUTE_PROGRAM
} }

struct RefLutIO {
    const float* transmittance;      // [64][256][4]
    const float* multiscattering;    // [32][32][4]
    const float* blue_noise;         // [64][64][4] (R16 texels as floats in .x), may be null when no dither flag is set
    float* sky_luminance;            // [sky_h][sky_w][4]
    float* sky_transmittance;
    float* ap_luminance;             // [depth][32][32][4]
    float* ap_transmittance;
    float* environment;              // [6][size][size][4] (values rounded to fp16)
    const float* mesh_shadow_map;    // [S][S][4] (light-space depth in .x) when cfg->volumetric_light, else unused
    int mesh_shadow_size;
};

template <class F>
static void with_common_bindings(const SkyAtmosphereBufferData* a, const SkyAtmosphereRenderBufferData* r, const RefLutIO* io, F&& f) { f(); }

#define REF_BIND_COMMON()                                                                                                  \
    REF_LOAD_ATMOSPHERE(a);                                                                                                \
    REF_LOAD_RENDER(r);                                                                                                    \
    ref_bind_texture(transmittance_texture, io->transmittance, 256, 64, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);               \
    ref_bind_texture(multiscattering_texture, io->multiscattering, 32, 32, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);            \
    if (io->blue_noise) ref_bind_texture(blue_noise, io->blue_noise, 64, 64, 1, ref::REPEAT, ref::NEAREST)

extern "C" int ref_atmosphere_luts(const SkyAtmosphereBufferData* a, const SkyAtmosphereRenderBufferData* r, const SkyLutConfig* cfg,
                                   const RefLutIO* io) {
    ref::g_sky_w = cfg->sky_view_width; ref::g_sky_h = cfg->sky_view_height; ref::g_ap_depth = cfg->aerial_perspective_depth;
    const int sw = cfg->sky_view_width, sh = cfg->sky_view_height, D = cfg->aerial_perspective_depth, E = cfg->environment_size;
    if ((cfg->sky_view_dither || cfg->aerial_perspective_dither) && !io->blue_noise) return 1;
#define RUN_K3(NS)                                                                                               \
    {                                                                                                            \
        using namespace ref::NS;                                                                                 \
        REF_BIND_COMMON();                                                                                       \
        ref_bind_image(luminance_image, io->sky_luminance, sw, sh, 1, ref::FMT_RGBA32F);                         \
        ref_bind_image(transmittance_image, io->sky_transmittance, sw, sh, 1, ref::FMT_RGBA32F);                 \
        ref::dispatch(main, ref_ceil_div(sw, LOCAL_SIZE_X), ref_ceil_div(sh, LOCAL_SIZE_Y), 1, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false); \
    }
#define RUN_PERM(RUN, NS)                                                                                         \
    {                                                                                                            \
        using namespace ref::NS;                                                                                 \
        if (cfg->volumetric_light) {                                                                             \
            ref_bind_texture(ref::NS::shadow_map_texture, io->mesh_shadow_map, io->mesh_shadow_size, io->mesh_shadow_size, 1, ref::CLAMP_TO_BORDER, ref::LINEAR); \
            ref::NS::shadow_map_texture.border = ref::vec4(1.0f);                                                \
        }                                                                                                        \
    }                                                                                                            \
    RUN(NS)
    const bool perm = cfg->volumetric_light || cfg->moon_shadow;
    if (perm && (cfg->sky_view_dither || cfg->aerial_perspective_dither)) return 3;  // permutation not compiled
    if (cfg->volumetric_light && !io->mesh_shadow_map) return 4;
    if (perm) {
        if (cfg->volumetric_light && cfg->moon_shadow) { RUN_PERM(RUN_K3, k3v1m1) }
        else if (cfg->volumetric_light) { RUN_PERM(RUN_K3, k3v1m0) }
        else { RUN_PERM(RUN_K3, k3v0m1) }
    } else if (cfg->sky_view_dither) RUN_K3(k3d1) else RUN_K3(k3d0)
#define RUN_K4(NS)                                                                                               \
    {                                                                                                            \
        using namespace ref::NS;                                                                                 \
        REF_BIND_COMMON();                                                                                       \
        ref_bind_image(luminance_image, io->ap_luminance, 32, 32, D, ref::FMT_RGBA32F);                          \
        ref_bind_image(transmittance_image, io->ap_transmittance, 32, 32, D, ref::FMT_RGBA32F);                  \
        ref::dispatch(main, ref_ceil_div(32, LOCAL_SIZE_X), ref_ceil_div(32, LOCAL_SIZE_Y), D, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false); \
    }
    if (perm) {
        if (cfg->volumetric_light && cfg->moon_shadow) { RUN_PERM(RUN_K4, k4v1m1) }
        else if (cfg->volumetric_light) { RUN_PERM(RUN_K4, k4v1m0) }
        else { RUN_PERM(RUN_K4, k4v0m1) }
    } else if (cfg->aerial_perspective_dither) RUN_K4(k4d1) else RUN_K4(k4d0)
    {
        using namespace ref::k5;
        REF_BIND_COMMON();
        ref_bind_texture(sky_view_luminance_texture, io->sky_luminance, sw, sh, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);
        ref_bind_texture(sky_view_transmittance_texture, io->sky_transmittance, sw, sh, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);
        ref_bind_image(env_luminance_image, io->environment, E, E, 6, ref::FMT_RGBA16F);
        ref::dispatch(main, ref_ceil_div(E, LOCAL_SIZE_X), ref_ceil_div(E, LOCAL_SIZE_Y), 6, LOCAL_SIZE_X, LOCAL_SIZE_Y, 1, false);
    }
    return 0;
}


// ---- K6 ------------------------------------------------------------------------------------------------------------
struct RefCompositeIO {
    const float* transmittance;      // [64][256][4]
    const float* multiscattering;    // [32][32][4]
    const float* blue_noise;         // [64][64][4]
    const float* sky_luminance;      // [sky_h][sky_w][4]
    const float* sky_transmittance;
    const float* ap_luminance;       // [depth][32][32][4]
    const float* ap_transmittance;
    const float* froxel;             // [fd][fh][fw][4] (unorm16 / 65535 in .x) or null (no cloud shadow froxel yet: visibility 1)
    int fw, fh, fd;
    const float* depth;              // [h][w][4], the D24 depth in .x
    const float* star;               // [sh][sw][4] linear RGB (sRGB already decoded) or null (black)
    int star_w, star_h;
    float* out;                      // [h][w][4]: FragColor
    int width, height;
    // object shading; albedo == null: an all-zero G-buffer, which makes ComputeObjectLuminance vanish
    const float* albedo;             // [h][w][4] decoded GL_RGBA8
    const float* normal;             // [h][w][4] decoded GL_RGBA16_SNORM
    const float* orm;                // [h][w][4] decoded GL_RGBA16
    const float* env_brdf_lut;       // [512][512][4] decoded GL_RG16
    const float* prefiltered;        // 5 levels from 128^2, concatenated, [6][n][n][4]
    const float* llm;                // [9][4]
    const float* cloud_shadow_map;   // [512][512][4] (depth, transmittance) or null (unshadowed)
    const float* mesh_shadow_map;    // [S][S][4] light-space depth in .x, or null (lit): object pixels only
    int mesh_shadow_size;
};

extern "C" int ref_composite(const SkyAtmosphereBufferData* a, const SkyAtmosphereRenderBufferData* r, const SkyLutConfig* cfg,
                             const RefCompositeIO* io) {
    ref::g_sky_w = cfg->sky_view_width; ref::g_sky_h = cfg->sky_view_height; ref::g_ap_depth = cfg->aerial_perspective_depth;
    if (cfg->volumetric_light) return 3;  // permutation not compiled
    if (cfg->moon_shadow && !(cfg->use_sky_view_lut && cfg->use_aerial_perspective_lut && cfg->raymarching_dither && !cfg->pcss)) return 3;
    static const float zero4[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    static const float one4[4] = {1.0f, 1.0f, 1.0f, 1.0f};
    static float zero_cube[6 * 4] = {};
    const int W = io->width, H = io->height;
#define RUN_K6(NS)                                                                                                          \
    {                                                                                                                       \
        using namespace ref::NS;                                                                                            \
        REF_LOAD_ATMOSPHERE(a);                                                                                             \
        REF_LOAD_RENDER(r);                                                                                                 \
        ref_bind_texture(transmittance_texture, io->transmittance, 256, 64, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);            \
        ref_bind_texture(multiscattering_texture, io->multiscattering, 32, 32, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);         \
        ref_bind_texture(depth_stencil_texture, io->depth, W, H, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);                      \
        /* an all-zero G-buffer makes ComputeObjectLuminance vanish: object pixels keep the in-scatter term alone */        \
        ref_bind_texture(albedo_texture, io->albedo ? io->albedo : zero4, io->albedo ? W : 1, io->albedo ? H : 1, 1, ref::CLAMP_TO_EDGE, ref::LINEAR); \
        ref_bind_texture(normal_texture, io->albedo ? io->normal : zero4, io->albedo ? W : 1, io->albedo ? H : 1, 1, ref::CLAMP_TO_EDGE, ref::LINEAR); \
        ref_bind_texture(orm_texture, io->albedo ? io->orm : zero4, io->albedo ? W : 1, io->albedo ? H : 1, 1, ref::CLAMP_TO_EDGE, ref::LINEAR); \
        if (io->albedo && io->mesh_shadow_map) {                                                                            \
            ref_bind_texture(ref::NS::shadow_map_texture, io->mesh_shadow_map, io->mesh_shadow_size, io->mesh_shadow_size, 1, ref::CLAMP_TO_BORDER, ref::LINEAR); \
            ref::NS::shadow_map_texture.border = ref::vec4(1.0f);                                                           \
        } else ref::NS::shadow_map_texture.levels.clear();  /* no mesh shadow map: lit */                                   \
        ref_bind_texture(blue_noise, io->blue_noise, 64, 64, 1, ref::REPEAT, ref::NEAREST);                                 \
        if (io->star) ref_bind_texture(star_luminance, io->star, io->star_w, io->star_h, 1, ref::CLAMP_TO_EDGE, ref::LINEAR); \
        else ref_bind_texture(star_luminance, zero4, 1, 1, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);                             \
        ref_bind_texture(sky_view_luminance_texture, io->sky_luminance, cfg->sky_view_width, cfg->sky_view_height, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);     \
        ref_bind_texture(sky_view_transmittance_texture, io->sky_transmittance, cfg->sky_view_width, cfg->sky_view_height, 1, ref::CLAMP_TO_EDGE, ref::LINEAR); \
        ref_bind_texture(aerial_perspective_luminance_texture, io->ap_luminance, 32, 32, cfg->aerial_perspective_depth, ref::CLAMP_TO_EDGE, ref::LINEAR);    \
        ref_bind_texture(aerial_perspective_transmittance_texture, io->ap_transmittance, 32, 32, cfg->aerial_perspective_depth, ref::CLAMP_TO_EDGE, ref::LINEAR); \
        if (io->albedo && io->mesh_shadow_map) ref_bind_texture(shadow_map_depth_sampler, io->mesh_shadow_map, io->mesh_shadow_size, io->mesh_shadow_size, 1, ref::CLAMP_TO_EDGE, ref::NEAREST); \
        else ref_bind_texture(shadow_map_depth_sampler, one4, 1, 1, 1, ref::CLAMP_TO_EDGE, ref::NEAREST);                   \
        if (io->albedo && io->cloud_shadow_map) {                                                                           \
            ref_bind_texture(cloud_shadow_map, io->cloud_shadow_map, 512, 512, 1, ref::CLAMP_TO_BORDER, ref::LINEAR);       \
            ref::NS::cloud_shadow_map.border = ref::vec4(1e10f, 1.0f, 0.0f, 0.0f);  /* VolumetricCloud.cpp:106-112 */       \
        } else ref_bind_texture(cloud_shadow_map, one4, 1, 1, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);                          \
        if (io->froxel) ref_bind_texture(cloud_shadow_froxel, io->froxel, io->fw, io->fh, io->fd, ref::CLAMP_TO_EDGE, ref::LINEAR); \
        else ref_bind_texture(cloud_shadow_froxel, one4, 1, 1, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);                         \
        if (io->albedo) {                                                                                                   \
            ref::Sampler& pr = prefiltered_radiance_texture;                                                                \
            pr.levels.assign(ROUGHNESS_COUNT, ref::Image());                                                                \
            const float* lp = io->prefiltered;                                                                              \
            for (int l = 0, n = 128; l < ROUGHNESS_COUNT; ++l, n >>= 1) {                                                   \
                ref::Image& im = pr.levels[l];                                                                              \
                im.data = const_cast<float*>(lp); im.w = n; im.h = n; im.d = 6; im.fmt = ref::FMT_RGBA32F;                  \
                lp += size_t(6) * n * n * 4;                                                                                \
            }                                                                                                               \
            pr.wrap = ref::CLAMP_TO_EDGE; pr.mag = pr.min_filter = ref::LINEAR;                                             \
            ref_bind_texture(env_brdf_lut, io->env_brdf_lut, 512, 512, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);                 \
            for (int i = 0; i < 9; ++i) Llm[i] = ref::vec4(io->llm[i * 4], io->llm[i * 4 + 1], io->llm[i * 4 + 2], io->llm[i * 4 + 3]); \
        } else {                                                                                                            \
            ref_bind_texture(prefiltered_radiance_texture, zero_cube, 1, 1, 6, ref::CLAMP_TO_EDGE, ref::LINEAR);            \
            ref_bind_texture(env_brdf_lut, zero4, 1, 1, 1, ref::CLAMP_TO_EDGE, ref::LINEAR);                                \
            for (int i = 0; i < 9; ++i) Llm[i] = ref::vec4(0.0f);                                                           \
        }                                                                                                                   \
        _Pragma("omp parallel for schedule(dynamic, 4) if(!serial)")                                                        \
        for (int py = 0; py < H; ++py)                                                                                      \
            for (int px = 0; px < W; ++px) {                                                                                \
                ref::g_builtins.frag_coord = ref::vec4(float(px) + 0.5f, float(py) + 0.5f, 0.0f, 1.0f);                     \
                vTexCoord = ref::vec2((float(px) + 0.5f) / float(W), (float(py) + 0.5f) / float(H));                        \
                main();                                                                                                     \
                float* o = io->out + (size_t(py) * W + px) * 4;                                                             \
                o[0] = FragColor.x; o[1] = FragColor.y; o[2] = FragColor.z; o[3] = FragColor.w;                             \
            }                                                                                                               \
    }
    // Shadow.glsl keeps its sample table in a GLSL global (`vec2 poissonDisk[]`: one per invocation; here one per process), so the
    // PCSS permutation runs its fragments one after the other
    const bool serial = cfg->pcss != 0;
    if (cfg->pcss) {
        if (cfg->use_sky_view_lut && !cfg->use_aerial_perspective_lut && cfg->raymarching_dither) RUN_K6(k6s1a0p1)
        else return 3;
        return 0;
    }
    if (cfg->moon_shadow) { RUN_K6(k6s1a1m1) return 0; }
    if (cfg->use_sky_view_lut && cfg->use_aerial_perspective_lut && cfg->raymarching_dither) RUN_K6(k6s1a1)
    else if (cfg->use_sky_view_lut && !cfg->use_aerial_perspective_lut && cfg->raymarching_dither) RUN_K6(k6s1a0)
    else if (!cfg->use_sky_view_lut && !cfg->use_aerial_perspective_lut && cfg->raymarching_dither) RUN_K6(k6s0a0)
    else if (!cfg->use_sky_view_lut && !cfg->use_aerial_perspective_lut && !cfg->raymarching_dither) RUN_K6(k6s0a0d0)
    else return 3;
    return 0;
}
