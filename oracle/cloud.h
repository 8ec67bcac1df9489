// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h).
//
// cloud.h: CPU restatement of the volumetric-cloud shaders: materials
// (VolumetricCloudDefaultMaterial{Common,0,1}.glsl, VolumetricCloudMaterial{Minimal,Voxel}.glsl),
// shadow chain K11-K13, real-time chain K14-K18 and the path tracer K19/K20.
#pragma once
#include <atomic>

#include "atmosphere.h"
#include "ibl.h"
#include "noise.h"

namespace orc {

struct CloudScene {
    // ---- atmosphere inputs (owned by the API context) ----
    Atmosphere atm;
    Image<4> transmittance;      // K1
    Image<4> multiscattering;    // K2
    Image<4> sky_lum, sky_trans; // K3
    Image<4> ap_lum, ap_trans;   // K4
    Image<4> env;                // K5: 6 faces along z, fp16-rounded
    CubeChain env_chain;         // env + its mips (glGenerateTextureMipmap, AtmosphereRenderer.cpp:242), filled by ibl_precompute
    Image<2> env_brdf_lut;       // K22: RG16 512x512
    vec4 env_sh[9];              // K23
    CubeChain prefiltered;       // K24: 5 levels from 128^2
    ObjectShading gbuffer;       // sky_set_gbuffer (albedo / normal / orm only)
    SkyAtmosphereRenderBufferData render_u{};
    SkyLutConfig lut_cfg{};

    Image<1> blue_noise;  // 64x64, u16/65535 (Textures.cpp:19-26)
    int out_band_rows = 0, out_band_index = 0, out_band_count = 1;  // sky_set_output_bands: full-res rows K6 / K18 own
    bool owns_row(int y) const { return out_band_count <= 1 || (y / out_band_rows) % out_band_count == out_band_index; }
    Image<4> star_map;         // decoded (linear) star map, empty when none was set (Textures.cpp:43-50)
    Image<1> mesh_shadow_map;  // 2048x2048 light-space depth (ShadowMap.cpp:8-27), allocated on first use, cleared to 1

    // ---- material ----
    SkyMaterialBlock material{};
    MipTexture<2> cloud_map;     // RG8 512^2
    MipTexture<1> detail;        // R8 128^3
    MipTexture<4> displacement;  // RGBA8 128^2
    MipTexture<1> voxel;         // R8 dx*dy*dz

    // ---- uniforms of the current pass ----
    SkyCloudCommonBufferData c{};  // VolumetricCloudCommon.glsl:6-28
    SkyCloudBufferData b{};        // VolumetricCloudRender.comp:17-32

    // ---- shadow chain (VolumetricCloud.h:114-115) ----
    Image<2> shadow_maps[3];  // current raw, previous raw, blurred; RG32F 512^2
    Image<1> shadow_froxel;   // R16 unorm (W/12, H/12, 128)

    // ---- viewport data (VolumetricCloud.cpp:120-136) ----
    int width = 0, height = 0;
    Image<1> checkerboard_depth;    // R32F W/2 x H/2
    Image<2> index_linear_depth;    // RG32F W/4 x H/4
    Image<4> render_texture;        // RGBA16F W/4 x H/4 (values fp16-rounded)
    Image<1> cloud_distance;        // R32F W/4 x H/4
    Image<4> reconstruct[2];        // RGBA16F W/2 x H/2 ([0] = being written, [1] = previous)

    // ---- path tracer (VolumetricCloud.cpp:495-531) ----
    SkyPathTracingInit pt{};
    Image<4> pt_accum;           // RGBA32F W x H
    std::vector<uint8_t> pt_mask;

    // ---- work counters (SkyCounter) ----
    bool counting = false;
    std::atomic<uint64_t> counters[8];

    CloudScene() { for (auto& x : counters) x = 0; }

    void SetViewport(int w, int h);

    vec3 uCameraPos() const { return vec3(c.uCameraPos); }
    vec3 uSunDirection() const { return vec3(c.uSunDirection); }

    // VolumetricCloudCommon.glsl:32-39
    float DepthToLinearDepth(float depth) const { return 1.0f / (c.uLinearDepthParam[0] - c.uLinearDepthParam[1] * depth); }
    float CalHeight01(vec3 pos) const {
        float altitude = length(vec3(pos.xy(), pos.z + c.uEarthRadius)) - c.uEarthRadius;
        return clamp((altitude - c.uBottomAltitude) / (c.uTopAltitude - c.uBottomAltitude), 0.0f, 1.0f);
    }
    // VolumetricCloudCommon.glsl:42-52
    static ivec2 IndexToOffset(uint index) { return ivec2(int(((index + 1) >> 1) & 1), int(((index + 2) >> 1) & 1)); }

    // material entry point: float SampleSigmaT(vec3 pos, float height01)
    float SampleSigmaT(vec3 pos, float height01, int counter_slot);

    // VolumetricCloudCommon.glsl:73-79
    vec3 GetSunVisibility(vec3 pos) const {
        vec3 up_dir(pos.x, pos.y, pos.z + c.uEarthRadius);
        float r = length(up_dir);
        up_dir /= r;
        float mu_s = dot(uSunDirection(), up_dir);
        return atm.GetSunVisibility(transmittance, r, mu_s);
    }
    // VolumetricCloudCommon.glsl:81-97
    vec3 GetAerialPerspective(vec2 uv, float t, float r, float mu, vec3& transmittance_out) const;
    // VolumetricCloudShadowInterface.glsl:4-8, sampler = shadow_map_sampler_ (VolumetricCloud.cpp:106-112)
    float SampleCloudShadowTransmittance(const Image<2>& cloud_shadow_map, vec3 light_ndc) const;

    void ShadowMap();        // K11 VolumetricCloudShadowMap.comp:37-75 (after the swap at VolumetricCloud.cpp:284)
    void ShadowBlur();       // K12 VolumetricCloudShadowMapBlur.comp:14-42, both passes
    void ShadowFroxel();     // K13 VolumetricCloudShadowFroxel.comp:10-27
    void CheckerboardGen(const float* depth);   // K14 CheckerboardGen.comp:7-14
    void IndexGen();                            // K15 VolumetricCloudIndexGen.comp:13-42
    void Render(int band_rows = 0, int band_index = 0, int band_count = 1);  // K16 VolumetricCloudRender.comp:139-210
    void Reconstruct();                         // K17 VolumetricCloudReconstruct.comp:28-110
    void Upscale(const float* depth, uint16_t* hdr_half4);  // K18 VolumetricCloudUpscale.comp:11-55

    void PathTraceBegin(const SkyPathTracingInit& init);
    void PathTraceSamples(uint32_t frame_begin, uint32_t count, const int32_t region[4]);  // K19
    void PathTraceResolve(uint32_t frame_count, uint16_t* hdr_half4) const;                // K20
};

}  // namespace orc
