// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h).
//
// oracle_api.cpp: exposes the CPU restatement behind the very same C ABI as libskyb200.so, with
// the prefix orc_ (include/skyb200.h compiled with SKY_FN(name) = orc_##name), so that one ctypes
// binding drives either side on identical inputs.  "dev" pointers are host pointers here.
#define SKY_FN(name) orc_##name
#include "../include/skyb200.h"

#include <cstdio>
#include <string>

#include "../include/sky_detmath.h"
#include "cloud.h"
#include "earth.h"

using namespace orc;

struct SkyContext {
    CloudScene scene;
    EarthAlbedo earth_albedo;       // sky_set_earth_albedo
    std::string error;
    std::vector<uint8_t> scratch;   // packed copies handed out by get_resource
    std::vector<uint64_t> counter_copy;
};

namespace {
int fail(SkyContext* ctx, const char* msg) { if (ctx) ctx->error = msg; return 1; }

template <int C>
void pack_unorm(const Image<C>& img, int bits, std::vector<uint8_t>& out) {
    size_t n = img.data.size();
    float m = float((1u << bits) - 1u);
    if (bits == 8) {
        out.resize(n);
        for (size_t i = 0; i < n; ++i) out[i] = uint8_t(std::nearbyint(img.data[i] * m));
    } else {
        out.resize(n * 2);
        uint16_t* p = reinterpret_cast<uint16_t*>(out.data());
        for (size_t i = 0; i < n; ++i) p[i] = uint16_t(std::nearbyint(img.data[i] * m));
    }
}
template <int C>
void pack_half(const Image<C>& img, std::vector<uint8_t>& out) {
    size_t n = img.data.size();
    out.resize(n * 2);
    uint16_t* p = reinterpret_cast<uint16_t*>(out.data());
    for (size_t i = 0; i < n; ++i) p[i] = float_to_half_bits(img.data[i]);
}
template <int C>
void pack_mips(const MipTexture<C>& t, std::vector<uint8_t>& out) {
    out.clear();
    for (size_t l = 1; l < t.levels.size(); ++l) {
        std::vector<uint8_t> tmp;
        pack_unorm(t.levels[l], 8, tmp);
        out.insert(out.end(), tmp.begin(), tmp.end());
    }
}
template <int C>
void set_desc_f32(SkyResourceDesc* d, Image<C>& img) {
    d->ptr = img.data.data(); d->width = img.w; d->height = img.h; d->depth = img.d; d->channels = C;
    d->format = SKY_FMT_F32; d->bytes = img.data.size() * 4;
}
// ShadowMap(2048, 2048) cleared to 1.0 (AppWindow.cpp:25, ShadowMap.cpp:8-27); allocated on first use
Image<1>& mesh_shadow_map(CloudScene& s) {
    if (s.mesh_shadow_map.w == 0) {
        s.mesh_shadow_map.resize(2048, 2048);
        std::fill(s.mesh_shadow_map.data.begin(), s.mesh_shadow_map.data.end(), 1.0f);
    }
    return s.mesh_shadow_map;
}
}  // namespace

extern "C" {

int orc_ctx_create(int, void*, SkyContext** out) {
    *out = new SkyContext();
    (*out)->scene.blue_noise.resize(64, 64);
    (*out)->scene.transmittance.resize(256, 64);    // Atmosphere.cpp:11-12
    (*out)->scene.multiscattering.resize(32, 32);   // Atmosphere.cpp:17-18
    return 0;
}
void orc_ctx_destroy(SkyContext* ctx) { delete ctx; }
const char* orc_last_error(SkyContext* ctx) { return ctx ? ctx->error.c_str() : "null context"; }
int orc_sync(SkyContext*) { return 0; }

int orc_set_star_map(SkyContext* ctx, const uint8_t* srgb8, int width, int height) {
    if (width <= 0 || height <= 0) { ctx->scene.star_map.resize(0, 0); return 0; }
    float decode[256];  // GL 4.6 section 8.24: sRGB -> linear, applied to each texel before filtering
    for (int c = 0; c < 256; ++c) {
        double cs = c / 255.0;
        decode[c] = float(cs <= 0.04045 ? cs / 12.92 : std::pow((cs + 0.055) / 1.055, 2.4));
    }
    Image<4>& m = ctx->scene.star_map;
    m.resize(width, height);
    for (size_t i = 0; i < size_t(width) * height; ++i) {
        for (int k = 0; k < 3; ++k) m.data[i * 4 + k] = decode[srgb8[i * 3 + k]];
        m.data[i * 4 + 3] = 1.0f;
    }
    return 0;
}

int orc_set_earth_albedo(SkyContext* ctx, const uint8_t* srgb8, int width, int height) {
    BuildEarthAlbedo(srgb8, width, height, ctx->earth_albedo);
    return 0;
}

int orc_gbuffer_clear(SkyContext* ctx, float* depth, void* albedo, void* normal, void* orm, int width, int height) {
    if (!depth || !albedo || !normal || !orm || width <= 0 || height <= 0) return fail(ctx, "gbuffer_clear: bad arguments");
    const size_t n = size_t(width) * height;   // Clear(const GBuffer&), GBuffer.h:28-34
    std::fill(depth, depth + n, 1.0f);
    std::memset(albedo, 0, n * 4); std::memset(normal, 0, n * 8); std::memset(orm, 0, n * 8);
    return 0;
}

int orc_earth_gbuffer(SkyContext* ctx, const SkyEarthBufferData* earth, float* depth, void* albedo, void* normal, void* orm, int width, int height) {
    if (!earth || !depth || !albedo || !normal || !orm || width <= 0 || height <= 0) return fail(ctx, "earth_gbuffer: bad arguments");
    EarthGBuffer(ctx->scene.atm, *earth, ctx->earth_albedo, depth, static_cast<uint8_t*>(albedo), static_cast<int16_t*>(normal), static_cast<uint16_t*>(orm), width, height);
    return 0;
}

// Test hook (oracle only, not part of skyb200.h): the deterministic fp32 functions of include/sky_detmath.h on arrays, so that
// tests/test_earth_cpu.py can check them against double precision.  fn: 0 exp, 1 sin, 2 cos, 3 acos, 4 asin, 5 atan2(x, y), 6 log2
int orc_detmath(int fn, const float* x, const float* y, float* out, int n) {
    for (int i = 0; i < n; ++i) {
        switch (fn) {
            case 0: out[i] = sky_det_expf(x[i]); break;
            case 1: out[i] = sky_det_sinf(x[i]); break;
            case 2: out[i] = sky_det_cosf(x[i]); break;
            case 3: out[i] = sky_det_acosf(x[i]); break;
            case 4: out[i] = sky_det_asinf(x[i]); break;
            case 5: out[i] = sky_det_atan2f(x[i], y[i]); break;
            case 6: out[i] = sky_det_log2f(x[i]); break;
            default: return 1;
        }
    }
    return 0;
}

int orc_set_blue_noise(SkyContext* ctx, const uint16_t* texels) {
    for (int i = 0; i < 64 * 64; ++i) ctx->scene.blue_noise.data[i] = float(texels[i]) / 65535.0f;
    return 0;
}

int orc_set_viewport(SkyContext* ctx, int w, int h) {
    if (w < 12 || h < 12) return fail(ctx, "viewport too small");
    ctx->scene.SetViewport(w, h);
    return 0;
}

int orc_atmosphere_bake(SkyContext* ctx, const SkyAtmosphereBufferData* a) {
    ctx->scene.atm.u = *a;
    ctx->scene.atm.BakeTransmittance(ctx->scene.transmittance);
    ctx->scene.atm.BakeMultiscattering(ctx->scene.transmittance, ctx->scene.multiscattering);
    return 0;
}

int orc_atmosphere_luts(SkyContext* ctx, const SkyAtmosphereRenderBufferData* r, const SkyLutConfig* cfg) {
    CloudScene& s = ctx->scene;
    s.render_u = *r;
    s.lut_cfg = *cfg;
    s.sky_lum.resize(cfg->sky_view_width, cfg->sky_view_height);
    s.sky_trans.resize(cfg->sky_view_width, cfg->sky_view_height);
    s.ap_lum.resize(32, 32, cfg->aerial_perspective_depth);  // AtmosphereRenderer.cpp:19-20
    s.ap_trans.resize(32, 32, cfg->aerial_perspective_depth);
    s.env.resize(cfg->environment_size, cfg->environment_size, 6);
    AtmosphereRenderer ar{s.atm, *r, *cfg, s.transmittance, s.multiscattering, &s.blue_noise};
    if (cfg->volumetric_light) ar.mesh_shadow_map = &mesh_shadow_map(s);
    ar.BakeSkyView(s.sky_lum, s.sky_trans);
    ar.BakeAerialPerspective(s.ap_lum, s.ap_trans);
    ar.BakeEnvironment(s.sky_lum, s.sky_trans, s.env);
    return 0;
}

int orc_env_brdf_lut(SkyContext* ctx) {
    ctx->scene.env_brdf_lut.resize(SKY_ENV_BRDF_LUT_SIZE, SKY_ENV_BRDF_LUT_SIZE);
    BakeEnvBRDFLut(ctx->scene.env_brdf_lut);
    return 0;
}

int orc_ibl_precompute(SkyContext* ctx) {
    CloudScene& s = ctx->scene;
    if (s.env.w <= 0) return fail(ctx, "ibl_precompute: the environment cube has not been baked (call atmosphere_luts first)");
    if (s.env.w & (s.env.w - 1)) return fail(ctx, "ibl_precompute: the environment size must be a power of two");
    GenerateCubeMips(s.env, s.env_chain);
    EnvRadianceSH(s.env_chain, s.env_sh);
    PrefilterRadiance(s.env_chain, SKY_IBL_PREFILTERED_RESOLUTION, SKY_IBL_ROUGHNESS_COUNT, s.prefiltered);
    return 0;
}

int orc_set_gbuffer(SkyContext* ctx, const void* albedo, const void* normal, const void* orm) {
    if ((albedo != nullptr) != (normal != nullptr) || (albedo != nullptr) != (orm != nullptr)) return fail(ctx, "set_gbuffer: bind all three targets or none");
    ctx->scene.gbuffer.albedo = static_cast<const uint8_t*>(albedo);
    ctx->scene.gbuffer.normal = static_cast<const int16_t*>(normal);
    ctx->scene.gbuffer.orm = static_cast<const uint16_t*>(orm);
    return 0;
}

int orc_composite(SkyContext* ctx, const float* depth, void* hdr, int width, int height) {
    CloudScene& s = ctx->scene;
    AtmosphereRenderer ar{s.atm, s.render_u, s.lut_cfg, s.transmittance, s.multiscattering, &s.blue_noise};
    if (s.lut_cfg.volumetric_light) ar.mesh_shadow_map = &mesh_shadow_map(s);
    ObjectShading object = s.gbuffer;
    if (object.albedo) {
        if (s.env_brdf_lut.w == 0 || s.prefiltered.levels.empty()) return fail(ctx, "composite: a G-buffer is bound but env_brdf_lut / ibl_precompute have not run");
        object.env_brdf_lut = &s.env_brdf_lut; object.prefiltered = &s.prefiltered; object.Llm = s.env_sh;
        if (s.shadow_maps[2].w > 0) object.cloud_shadow_map = &s.shadow_maps[2];
        if (s.mesh_shadow_map.w > 0) ar.mesh_shadow_map = &s.mesh_shadow_map;
        ar.object = &object;
    }
    if (s.star_map.w > 0) ar.star_map = &s.star_map;
    ar.out_band_rows = s.out_band_rows; ar.out_band_index = s.out_band_index; ar.out_band_count = s.out_band_count;
    const Image<1>* froxel = (s.shadow_froxel.w > 0) ? &s.shadow_froxel : nullptr;
    ar.Composite(s.sky_lum, s.sky_trans, s.ap_lum, s.ap_trans, froxel, depth, width, height, static_cast<uint16_t*>(hdr));
    return 0;
}

int orc_noise_generate(SkyContext* ctx, int kind, const SkyNoiseCreateInfo* info) {
    switch (kind) {
        case SKY_NOISE_CLOUD_MAP: GenerateCloudMap(info, ctx->scene.cloud_map); return 0;
        case SKY_NOISE_DETAIL: GenerateDetail(info, ctx->scene.detail); return 0;
        case SKY_NOISE_DISPLACEMENT: GenerateDisplacement(info, ctx->scene.displacement); return 0;
    }
    return fail(ctx, "unknown noise kind");
}

int orc_voxel_upload(SkyContext* ctx, const uint8_t* v, int dx, int dy, int dz) {
    MipTexture<1>& t = ctx->scene.voxel;
    t.bits = 8;
    t.levels.resize(1);
    t.levels[0].resize(dx, dy, dz);
    for (size_t i = 0; i < t.levels[0].data.size(); ++i) t.levels[0].data[i] = float(v[i]) / 255.0f;
    t.build_mips();
    return 0;
}

int orc_set_material(SkyContext* ctx, const SkyMaterialBlock* m) { ctx->scene.material = *m; return 0; }

int orc_cloud_shadow(SkyContext* ctx, const SkyCloudCommonBufferData* common) {
    CloudScene& s = ctx->scene;
    if (s.width == 0) return fail(ctx, "Volumetric cloud viewport is undefined");  // VolumetricCloud.cpp:169-170
    s.c = *common;
    s.ShadowMap();
    s.ShadowBlur();
    s.ShadowFroxel();
    return 0;
}

int orc_cloud_frame_begin(SkyContext* ctx, const SkyCloudCommonBufferData* common, const SkyCloudBufferData* cloud,
                          const float* depth, int band_rows, int band_index, int band_count) {
    CloudScene& s = ctx->scene;
    if (s.width == 0) return fail(ctx, "Volumetric cloud viewport is undefined");
    s.c = *common;
    s.b = *cloud;
    s.CheckerboardGen(depth);
    s.IndexGen();
    s.Render(band_rows, band_index, band_count);
    return 0;
}

int orc_cloud_frame_end(SkyContext* ctx, const float* depth, void* hdr) {
    CloudScene& s = ctx->scene;
    s.Reconstruct();
    s.Upscale(depth, static_cast<uint16_t*>(hdr));
    std::swap(s.reconstruct[0], s.reconstruct[1]);  // VolumetricCloud.cpp:421-422
    return 0;
}

int orc_cloud_frame(SkyContext* ctx, const SkyCloudCommonBufferData* common, const SkyCloudBufferData* cloud,
                    const float* depth, void* hdr) {
    if (int e = orc_cloud_frame_begin(ctx, common, cloud, depth, 0, 0, 1)) return e;
    return orc_cloud_frame_end(ctx, depth, hdr);
}

int orc_cloud_frame_host(SkyContext* ctx, const SkyCloudCommonBufferData* common, const SkyCloudBufferData* cloud,
                         const float* depth, void* hdr) {
    return orc_cloud_frame(ctx, common, cloud, depth, hdr);
}

int orc_pt_begin(SkyContext* ctx, const SkyPathTracingInit* init) {
    if (ctx->scene.width == 0) return fail(ctx, "Volumetric cloud viewport is undefined");
    ctx->scene.PathTraceBegin(*init);
    return 0;
}

int orc_pt_samples(SkyContext* ctx, const SkyCloudCommonBufferData* common, uint32_t frame_begin, uint32_t count,
                   const int32_t region[4]) {
    if (ctx->scene.pt_accum.w == 0) return fail(ctx, "pt_begin was not called");
    ctx->scene.c = *common;
    ctx->scene.PathTraceSamples(frame_begin, count, region);
    return 0;
}

int orc_pt_resolve(SkyContext* ctx, uint32_t frame_count, void* hdr) {
    ctx->scene.PathTraceResolve(frame_count, static_cast<uint16_t*>(hdr));
    return 0;
}

// shaders/Base/BloomPass2.frag:15-42 without the bloom term (see skyb200.h)
int orc_tonemap(SkyContext* ctx, const void* hdr, int width, int height, const SkyToneMapParams* p, void* rgba8) {
    if (p->tone_mapping != 0 && p->tone_mapping != 1) return fail(ctx, "tonemap: unknown tone mapping operator");
    const uint16_t* in = static_cast<const uint16_t*>(hdr);
    uint8_t* out = static_cast<uint8_t*>(rgba8);
    auto tone = [&](float luminance) {
        if (p->tone_mapping == 0) return 1 - std::exp(-p->exposure * luminance);
        const float k = 10.0f / 16.0f;
        const float A = 2.51f * k * k, B = 0.03f * k, C = 2.43f * k * k, D = 0.59f * k, E = 0.14f;
        luminance *= p->exposure;
        return (luminance * (A * luminance + B)) / (luminance * (C * luminance + D) + E);
    };
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            size_t pix = size_t(y) * width + x;
            for (int k = 0; k < 3; ++k) {
                float v = std::pow(tone(half_bits_to_float(in[pix * 4 + k])), 1.0f / 2.2f);
                if (p->dither) v += ctx->scene.blue_noise.data[size_t(y & 63) * 64 + (x & 63)] / 255.0f;
                v = std::fmin(std::fmax(v, 0.0f), 1.0f);
                out[pix * 4 + k] = uint8_t(std::nearbyintf(v * 255.0f));
            }
            out[pix * 4 + 3] = 255;
        }
    return 0;
}

int orc_pt_samples_host(SkyContext* ctx, const SkyCloudCommonBufferData* common, uint32_t frame_begin, uint32_t count,
                        const int32_t region[4], float* accum_host) {
    if (int e = orc_pt_samples(ctx, common, frame_begin, count, region)) return e;
    std::memcpy(accum_host, ctx->scene.pt_accum.data.data(), ctx->scene.pt_accum.data.size() * 4);
    return 0;
}

int orc_get_resource(SkyContext* ctx, int resource, SkyResourceDesc* d) {
    CloudScene& s = ctx->scene;
    std::memset(d, 0, sizeof(*d));
    auto packed = [&](int w, int h, int dep, int ch, int fmt) {
        d->ptr = ctx->scratch.data(); d->width = w; d->height = h; d->depth = dep; d->channels = ch; d->format = fmt;
        d->bytes = ctx->scratch.size();
    };
    switch (resource) {
        case SKY_RES_TRANSMITTANCE: set_desc_f32(d, s.transmittance); return 0;
        case SKY_RES_MULTISCATTERING: set_desc_f32(d, s.multiscattering); return 0;
        case SKY_RES_SKY_VIEW_LUMINANCE: set_desc_f32(d, s.sky_lum); return 0;
        case SKY_RES_SKY_VIEW_TRANSMITTANCE: set_desc_f32(d, s.sky_trans); return 0;
        case SKY_RES_AERIAL_LUMINANCE: set_desc_f32(d, s.ap_lum); return 0;
        case SKY_RES_AERIAL_TRANSMITTANCE: set_desc_f32(d, s.ap_trans); return 0;
        case SKY_RES_ENVIRONMENT: pack_half(s.env, ctx->scratch); packed(s.env.w, s.env.h, 6, 4, SKY_FMT_F16); return 0;
        case SKY_RES_CLOUD_MAP:
            if (s.cloud_map.levels.empty()) return fail(ctx, "cloud map not generated");
            pack_unorm(s.cloud_map.levels[0], 8, ctx->scratch); packed(s.cloud_map.levels[0].w, s.cloud_map.levels[0].h, 1, 2, SKY_FMT_U8); return 0;
        case SKY_RES_DETAIL:
            if (s.detail.levels.empty()) return fail(ctx, "detail not generated");
            pack_unorm(s.detail.levels[0], 8, ctx->scratch); packed(s.detail.levels[0].w, s.detail.levels[0].h, s.detail.levels[0].d, 1, SKY_FMT_U8); return 0;
        case SKY_RES_DISPLACEMENT:
            if (s.displacement.levels.empty()) return fail(ctx, "displacement not generated");
            pack_unorm(s.displacement.levels[0], 8, ctx->scratch); packed(s.displacement.levels[0].w, s.displacement.levels[0].h, 1, 4, SKY_FMT_U8); return 0;
        case SKY_RES_VOXEL:
            if (s.voxel.levels.empty()) return fail(ctx, "voxel grid not uploaded");
            pack_unorm(s.voxel.levels[0], 8, ctx->scratch); packed(s.voxel.levels[0].w, s.voxel.levels[0].h, s.voxel.levels[0].d, 1, SKY_FMT_U8); return 0;
        case SKY_RES_CLOUD_MAP_MIPS: pack_mips(s.cloud_map, ctx->scratch); packed(int(ctx->scratch.size()), 1, 1, 1, SKY_FMT_U8); return 0;
        case SKY_RES_DETAIL_MIPS: pack_mips(s.detail, ctx->scratch); packed(int(ctx->scratch.size()), 1, 1, 1, SKY_FMT_U8); return 0;
        case SKY_RES_DISPLACEMENT_MIPS: pack_mips(s.displacement, ctx->scratch); packed(int(ctx->scratch.size()), 1, 1, 1, SKY_FMT_U8); return 0;
        case SKY_RES_VOXEL_MIPS: pack_mips(s.voxel, ctx->scratch); packed(int(ctx->scratch.size()), 1, 1, 1, SKY_FMT_U8); return 0;
        case SKY_RES_SHADOW_MAP_RAW: set_desc_f32(d, s.shadow_maps[0]); return 0;
        case SKY_RES_SHADOW_MAP: set_desc_f32(d, s.shadow_maps[2]); return 0;
        case SKY_RES_SHADOW_FROXEL: pack_unorm(s.shadow_froxel, 16, ctx->scratch); packed(s.shadow_froxel.w, s.shadow_froxel.h, s.shadow_froxel.d, 1, SKY_FMT_U16); return 0;
        case SKY_RES_CHECKERBOARD_DEPTH: set_desc_f32(d, s.checkerboard_depth); return 0;
        case SKY_RES_INDEX_LINEAR_DEPTH: set_desc_f32(d, s.index_linear_depth); return 0;
        case SKY_RES_CLOUD_RENDER: pack_half(s.render_texture, ctx->scratch); packed(s.render_texture.w, s.render_texture.h, 1, 4, SKY_FMT_F16); return 0;
        case SKY_RES_CLOUD_DISTANCE: set_desc_f32(d, s.cloud_distance); return 0;
        case SKY_RES_RECONSTRUCT: pack_half(s.reconstruct[1], ctx->scratch); packed(s.reconstruct[1].w, s.reconstruct[1].h, 1, 4, SKY_FMT_F16); return 0;
        case SKY_RES_PT_ACCUM: set_desc_f32(d, s.pt_accum); return 0;
        case SKY_RES_PT_MASK:
            d->ptr = s.pt_mask.data(); d->width = s.width; d->height = s.height; d->depth = 1; d->channels = 1; d->format = SKY_FMT_U8;
            d->bytes = s.pt_mask.size(); return 0;
        case SKY_RES_MESH_SHADOW_MAP: set_desc_f32(d, mesh_shadow_map(s)); return 0;
        case SKY_RES_ENV_BRDF_LUT:
            if (s.env_brdf_lut.w == 0) return fail(ctx, "env_brdf_lut has not been baked");
            pack_unorm(s.env_brdf_lut, 16, ctx->scratch); packed(s.env_brdf_lut.w, s.env_brdf_lut.h, 1, 2, SKY_FMT_U16); return 0;
        case SKY_RES_ENVIRONMENT_MIPS: case SKY_RES_PREFILTERED_RADIANCE: {
            const CubeChain& c = resource == SKY_RES_ENVIRONMENT_MIPS ? s.env_chain : s.prefiltered;
            const size_t first = resource == SKY_RES_ENVIRONMENT_MIPS ? 1 : 0;
            if (c.levels.size() <= first) return fail(ctx, "ibl_precompute has not run");
            std::vector<uint8_t> all, one;
            for (size_t l = first; l < c.levels.size(); ++l) { pack_half(c.levels[l], one); all.insert(all.end(), one.begin(), one.end()); }
            ctx->scratch.swap(all);
            packed(int(ctx->scratch.size() / 2), 1, 1, 1, SKY_FMT_F16); return 0;  // flat, like the *_MIPS resources
        }
        case SKY_RES_EARTH_ALBEDO: {
            const EarthAlbedo& m = ctx->earth_albedo;
            if (!m.valid()) return fail(ctx, "no earth albedo map");
            ctx->scratch.clear();
            for (const std::vector<uint8_t>& lvl : m.codes)
                for (size_t i = 0; i < lvl.size(); i += 3) { ctx->scratch.push_back(lvl[i]); ctx->scratch.push_back(lvl[i + 1]); ctx->scratch.push_back(lvl[i + 2]); ctx->scratch.push_back(255); }
            packed(int(ctx->scratch.size() / 4), 1, 1, 4, SKY_FMT_U8); return 0;
        }
        case SKY_RES_ENV_RADIANCE_SH:
            d->ptr = &s.env_sh[0].x; d->width = 9; d->height = 1; d->depth = 1; d->channels = 4; d->format = SKY_FMT_F32; d->bytes = 9 * 16; return 0;
        case SKY_RES_COUNTERS:
            ctx->counter_copy.resize(8);
            for (int i = 0; i < 8; ++i) ctx->counter_copy[i] = s.counters[i].load();
            d->ptr = ctx->counter_copy.data(); d->width = 8; d->height = 1; d->depth = 1; d->channels = 1; d->format = SKY_FMT_U64;
            d->bytes = 64; return 0;
    }
    return fail(ctx, "unknown resource");
}

int orc_read_resource(SkyContext* ctx, int resource, void* dst, uint64_t bytes) {
    SkyResourceDesc d;
    if (int e = orc_get_resource(ctx, resource, &d)) return e;
    if (bytes != d.bytes) { ctx->error = "read_resource: size mismatch, expected " + std::to_string(d.bytes); return 1; }
    std::memcpy(dst, d.ptr, bytes);
    return 0;
}

int orc_write_resource(SkyContext* ctx, int resource, const void* src, uint64_t bytes) {
    CloudScene& s = ctx->scene;
    auto put_f32 = [&](std::vector<float>& v) { if (bytes != v.size() * 4) return fail(ctx, "write_resource: size mismatch"); std::memcpy(v.data(), src, bytes); return 0; };
    auto put_half = [&](std::vector<float>& v) {
        if (bytes != v.size() * 2) return fail(ctx, "write_resource: size mismatch");
        const uint16_t* p = static_cast<const uint16_t*>(src);
        for (size_t i = 0; i < v.size(); ++i) v[i] = half_bits_to_float(p[i]);
        return 0;
    };
    switch (resource) {
        case SKY_RES_TRANSMITTANCE: return put_f32(s.transmittance.data);
        case SKY_RES_MULTISCATTERING: return put_f32(s.multiscattering.data);
        case SKY_RES_SKY_VIEW_LUMINANCE: return put_f32(s.sky_lum.data);
        case SKY_RES_SKY_VIEW_TRANSMITTANCE: return put_f32(s.sky_trans.data);
        case SKY_RES_AERIAL_LUMINANCE: return put_f32(s.ap_lum.data);
        case SKY_RES_AERIAL_TRANSMITTANCE: return put_f32(s.ap_trans.data);
        case SKY_RES_ENVIRONMENT: return put_half(s.env.data);
        case SKY_RES_SHADOW_MAP_RAW: return put_f32(s.shadow_maps[0].data);
        case SKY_RES_SHADOW_MAP: return put_f32(s.shadow_maps[2].data);
        case SKY_RES_CHECKERBOARD_DEPTH: return put_f32(s.checkerboard_depth.data);
        case SKY_RES_INDEX_LINEAR_DEPTH: return put_f32(s.index_linear_depth.data);
        case SKY_RES_CLOUD_RENDER: return put_half(s.render_texture.data);
        case SKY_RES_CLOUD_DISTANCE: return put_f32(s.cloud_distance.data);
        case SKY_RES_RECONSTRUCT: return put_half(s.reconstruct[1].data);
        case SKY_RES_PT_ACCUM: return put_f32(s.pt_accum.data);
        case SKY_RES_MESH_SHADOW_MAP: return put_f32(mesh_shadow_map(s).data);
        case SKY_RES_SHADOW_FROXEL: {
            if (bytes != s.shadow_froxel.data.size() * 2) return fail(ctx, "write_resource: size mismatch");
            const uint16_t* p = static_cast<const uint16_t*>(src);
            for (size_t i = 0; i < s.shadow_froxel.data.size(); ++i) s.shadow_froxel.data[i] = float(p[i]) / 65535.0f;
            return 0;
        }
        case SKY_RES_CLOUD_MAP: case SKY_RES_DETAIL: case SKY_RES_DISPLACEMENT: {
            // install level 0 codes and rebuild the chain
            const uint8_t* p = static_cast<const uint8_t*>(src);
            auto put = [&](auto& tex) {
                if (tex.levels.empty() || bytes != tex.levels[0].data.size()) return fail(ctx, "write_resource: size mismatch");
                for (size_t i = 0; i < bytes; ++i) tex.levels[0].data[i] = float(p[i]) / 255.0f;
                tex.build_mips();
                return 0;
            };
            if (resource == SKY_RES_CLOUD_MAP) { if (s.cloud_map.levels.empty()) { s.cloud_map.levels.resize(1); s.cloud_map.levels[0].resize(512, 512); } return put(s.cloud_map); }
            if (resource == SKY_RES_DETAIL) { if (s.detail.levels.empty()) { s.detail.levels.resize(1); s.detail.levels[0].resize(128, 128, 128); } return put(s.detail); }
            if (s.displacement.levels.empty()) { s.displacement.levels.resize(1); s.displacement.levels[0].resize(128, 128); }
            return put(s.displacement);
        }
    }
    return fail(ctx, "write_resource: resource is not writable");
}

int orc_launch_count(SkyContext*, uint64_t* launches) { if (launches) *launches = 0; return 0; }  // the oracle launches nothing
int orc_counters_enable(SkyContext* ctx, int enable) {
    ctx->scene.counting = enable != 0;
    for (auto& c : ctx->scene.counters) c = 0;
    return 0;
}

int orc_peer_export(SkyContext* ctx, SkyPeerHandles*) { return fail(ctx, "peer memory is a CUDA feature"); }
int orc_peer_attach(SkyContext* ctx, int, int, const SkyPeerHandles*) { return fail(ctx, "peer memory is a CUDA feature"); }
int orc_peer_detach(SkyContext*) { return 0; }
int orc_set_output_gather(SkyContext* ctx, int mode) { return mode == SKY_GATHER_OFF ? 0 : fail(ctx, "peer memory is a CUDA feature"); }
int orc_pt_set_tracking(SkyContext* ctx, int mode) { return mode == SKY_PT_TRACKING_REFERENCE ? 0 : fail(ctx, "the oracle only implements the reference's tracking"); }
int orc_set_hw_filtering(SkyContext*, int) { return 0; }
int orc_set_launch_shape(SkyContext* ctx, int kernel, int shape) { return kernel == SKY_KERNEL_K16 && shape >= SKY_K16_AUTO && shape <= SKY_K16_LITERAL ? 0 : fail(ctx, "set_launch_shape: unknown kernel or shape"); }
int orc_set_lut_arithmetic(SkyContext*, int) { return 0; }     // ... and the exact LUT march
int orc_set_strict_arithmetic(SkyContext*, int) { return 0; }  // the oracle IS the strict arithmetic
int orc_set_output_bands(SkyContext* ctx, int band_rows, int band_index, int band_count) {
    CloudScene& s = ctx->scene;
    if (band_count <= 1) { s.out_band_rows = 0; s.out_band_index = 0; s.out_band_count = 1; return 0; }
    if (band_rows < 8 || band_rows % 8 != 0 || band_index < 0 || band_index >= band_count) return fail(ctx, "set_output_bands: band_rows must be a positive multiple of 8 and 0 <= band_index < band_count");
    s.out_band_rows = band_rows; s.out_band_index = band_index; s.out_band_count = band_count;
    return 0;
}
int orc_set_frame_overlap(SkyContext*, int) { return 0; }
int orc_set_frame_pipelining(SkyContext*, int) { return 0; }
int orc_tex_peak(SkyContext* ctx, int, double*) { return fail(ctx, "tex_peak is a GPU microbenchmark"); }

// ---- test hooks (not part of skyb200.h): expose the leaf functions to the known-answer tests ------
uint32_t orc_test_wang_hash(uint32_t x) { return WangHash(x); }
uint32_t orc_test_pcg_hash(uint32_t x) { return PCGHash(x); }
float orc_test_perlin(float x, float y, float z, uint32_t freq, uint32_t seed) { return PerlinNoise(vec3(x, y, z), freq, seed); }
float orc_test_worley(float x, float y, float z, uint32_t freq, uint32_t seed) { return WorleyNoise3(vec3(x, y, z), freq, seed); }
float orc_test_worley2(float x, float y, uint32_t freq, uint32_t seed) { return WorleyNoise2(vec2(x, y), freq, seed); }
// R8 volume [d][h][w] sampled like the material textures (mag LINEAR, min NEAREST_MIPMAP_NEAREST)
float orc_test_sample_r8(const uint8_t* texels, int w, int h, int d, float u, float v, float s, float lod, int wrap) {
    MipTexture<1> t;
    t.levels.resize(1);
    t.levels[0].resize(w, h, d);
    for (size_t i = 0; i < t.levels[0].data.size(); ++i) t.levels[0].data[i] = float(texels[i]) / 255.0f;
    t.build_mips();
    Sampler smp;
    smp.wrap = Wrap(wrap);
    return d > 1 ? t.texture_lod(vec3(u, v, s), lod, smp).x : t.texture_lod(vec2(u, v), lod, smp).x;
}
// density of the current material at a local-frame position (uses the uniforms of the last pass)
float orc_test_sigma_t(SkyContext* ctx, float x, float y, float z) {
    vec3 p(x, y, z);
    return ctx->scene.SampleSigmaT(p, ctx->scene.CalHeight01(p), 7);
}

}  // extern "C"
