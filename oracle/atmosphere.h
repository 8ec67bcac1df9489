// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h).
//
// atmosphere.h: CPU restatement of shaders/SkyRendering/Atmosphere.glsl,
// AtmosphereInterface.glsl and the compute/fragment entry points of AtmosphereRenderer.glsl
// (K1-K6 in SURVEY.md section 2b).  Function names follow the GLSL; each cites the lines it follows.
//
// PARITY UNPINNED: the reference ships no golden vectors for this path; the
// restatement is pinned by derived known-answer tests in tests/test_oracle_kat.py instead.
#pragma once
#include "../include/sky_types.h"
#include "../include/sky_detmath.h"
#include "sampler.h"

namespace orc {

// The LUT bake (K1-K5) evaluates exp / sin / cos / acos / x^1.5 with the deterministic fp32 versions of
// include/sky_detmath.h, so that it can be compared bit for bit with the CUDA kernels.
inline vec3 lut_exp3(vec3 v) { return vec3(sky_det_expf(v.x), sky_det_expf(v.y), sky_det_expf(v.z)); }

constexpr float PI = 3.1415926535897932384626433832795f;  // shaders/Base/Common.glsl:4
constexpr float INV_PI = 1.0f / PI;

// shaders/Base/Common.glsl:7-10
inline vec3 ProjectiveMul(const mat4& m, vec3 v) {
    vec4 xyzw = m * vec4(v, 1.0f);
    return xyzw.xyz() / xyzw.w;
}

struct Atmosphere {
    SkyAtmosphereBufferData u;  // Atmosphere.glsl:6-32

    vec3 solar_illuminance() const { return vec3(u.solar_illuminance); }
    vec3 rayleigh_scattering() const { return vec3(u.rayleigh_scattering); }
    vec3 mie_scattering() const { return vec3(u.mie_scattering); }
    vec3 mie_absorption() const { return vec3(u.mie_absorption); }
    vec3 ozone_absorption() const { return vec3(u.ozone_absorption); }
    vec3 ground_albedo() const { return vec3(u.ground_albedo); }

    // Atmosphere.glsl:37-51
    static float ClampCosine(float mu) { return clamp(mu, -1.0f, 1.0f); }
    static float ClampDistance(float d) { return max(d, 0.0f); }
    static float SafeSqrt(float a) { return std::sqrt(max(a, 0.0f)); }

    // Atmosphere.glsl:53-55
    static vec2 GetTextureCoordFromUnitRange(vec2 xy, ivec2 size) {
        vec2 s = tovec2(size);
        return 0.5f / s + xy * (1.0f - 1.0f / s);
    }
    // AtmosphereInterface.glsl:15-17
    static vec3 GetTextureCoordFromUnitRange(vec3 xyz, ivec3 size) {
        vec3 s = tovec3(size);
        return 0.5f / s + xyz * (1.0f - 1.0f / s);
    }

    // Atmosphere.glsl:57-59
    bool RayIntersectsGround(float r, float mu) const {
        return mu < 0.0f && r * r * (mu * mu - 1.0f) + u.bottom_radius * u.bottom_radius >= 0.0f;
    }
    // Atmosphere.glsl:61-64
    float DistanceToTopAtmosphereBoundary(float r, float mu) const {
        float discriminant = r * r * (mu * mu - 1.0f) + u.top_radius * u.top_radius;
        return ClampDistance(-r * mu + SafeSqrt(discriminant));
    }
    // Atmosphere.glsl:66-69
    float DistanceToBottomAtmosphereBoundary(float r, float mu) const {
        float discriminant = r * r * (mu * mu - 1.0f) + u.bottom_radius * u.bottom_radius;
        return ClampDistance(-r * mu - SafeSqrt(discriminant));
    }
    // AtmosphereInterface.glsl:6-13
    bool FromSpaceIntersectTopAtmosphereBoundary(float r, float mu, float& near_distance) const {
        float discriminant = r * r * (mu * mu - 1.0f) + u.top_radius * u.top_radius;
        if (mu < 0.0f && discriminant >= 0.0f) {
            near_distance = ClampDistance(-r * mu - SafeSqrt(discriminant));
            return true;
        }
        return false;
    }

    // Atmosphere.glsl:71-88
    void GetRMuFromTransmittanceTextureIndex(ivec2 index, ivec2 size, float& r, float& mu) const {
        vec2 uv = tovec2(index) / tovec2(size - 1);
        float x_mu = uv.x, x_r = uv.y;
        float H = std::sqrt(u.top_radius * u.top_radius - u.bottom_radius * u.bottom_radius);
        float rho = H * x_r;
        r = std::sqrt(rho * rho + u.bottom_radius * u.bottom_radius);
        float d_min = u.top_radius - r;
        float d_max = rho + H;
        float d = d_min + x_mu * (d_max - d_min);
        mu = d == 0.0f ? 1.0f : (H * H - rho * rho - d * d) / (2.0f * r * d);
        mu = ClampCosine(mu);
    }

    // Atmosphere.glsl:90-103
    vec2 GetTransmittanceTextureUvFromRMu(ivec2 size, float r, float mu) const {
        float H = std::sqrt(u.top_radius * u.top_radius - u.bottom_radius * u.bottom_radius);
        float rho = SafeSqrt(r * r - u.bottom_radius * u.bottom_radius);
        float d = DistanceToTopAtmosphereBoundary(r, mu);
        float d_min = u.top_radius - r;
        float d_max = rho + H;
        float x_mu = (d - d_min) / (d_max - d_min);
        float x_r = rho / H;
        return GetTextureCoordFromUnitRange(vec2(x_mu, x_r), size);
    }

    // Atmosphere.glsl:105-108; sampler = LinearNoMipmapClampToEdge (Atmosphere.cpp:119)
    vec3 GetTransmittanceToTopAtmosphereBoundary(const Image<4>& tex, float r, float mu) const {
        vec2 uv = GetTransmittanceTextureUvFromRMu(ivec2(tex.w, tex.h), r, mu);
        return texture_linear(tex, uv, Sampler()).rgb();
    }

    // Atmosphere.glsl:110-117
    vec3 GetSunVisibility(const Image<4>& tex, float r, float mu_s) const {
        float sin_theta_h = u.bottom_radius / r;
        float cos_theta_h = -std::sqrt(max(1.0f - sin_theta_h * sin_theta_h, 0.0f));
        return GetTransmittanceToTopAtmosphereBoundary(tex, r, mu_s) *
               smoothstep(-sin_theta_h * u.sun_angular_radius, sin_theta_h * u.sun_angular_radius, mu_s - cos_theta_h);
    }

    // Atmosphere.glsl:119-132
    vec3 GetExtinction(float altitude) const {
        vec3 rayleigh_extinction =
            rayleigh_scattering() * clamp(sky_det_expf(-altitude * u.inv_rayleigh_exponential_distribution), 0.0f, 1.0f);
        vec3 mie_extinction = (mie_scattering() + mie_absorption()) *
                              clamp(sky_det_expf(-altitude * u.inv_mie_exponential_distribution), 0.0f, 1.0f);
        vec3 ozone_extinction =
            ozone_absorption() * max(0.0f, altitude < u.ozone_center_altitude
                                                   ? 1.0f + (altitude - u.ozone_center_altitude) * u.inv_ozone_width
                                                   : 1.0f - (altitude - u.ozone_center_altitude) * u.inv_ozone_width);
        return rayleigh_extinction + mie_extinction + ozone_extinction;
    }

    // Atmosphere.glsl:134-154 (MS = the MULTISCATTERING_COMPUTE_PROGRAM permutation)
    static float IsotropicPhaseFunction() { return 1.0f / (4.0f * PI); }
    template <bool MS>
    static float RayleighPhaseFunction(float cos_theta) {
        if (MS) return IsotropicPhaseFunction();
        float k = 3.0f / (16.0f * PI);
        return k * (1.0f + cos_theta * cos_theta);
    }
    template <bool MS>
    static float MiePhaseFunction(float g, float cos_theta) {
        if (MS) return IsotropicPhaseFunction();
        float k = 3.0f / (8.0f * PI) * (1.0f - g * g) / (2.0f + g * g);
        return k * (1.0f + cos_theta * cos_theta) / sky_det_pow15f(1.0f + g * g - 2.0f * g * cos_theta);
    }

    // Atmosphere.glsl:156-159
    void GetScattering(float altitude, vec3& rayleigh, vec3& mie) const {
        rayleigh = rayleigh_scattering() * clamp(sky_det_expf(-altitude * u.inv_rayleigh_exponential_distribution), 0.0f, 1.0f);
        mie = mie_scattering() * clamp(sky_det_expf(-altitude * u.inv_mie_exponential_distribution), 0.0f, 1.0f);
    }

    // Atmosphere.glsl:161-167
    void GetAltitudeMuSFromMultiscatteringTextureIndex(ivec2 index, ivec2 size, float& altitude, float& mu_s) const {
        vec2 uv = tovec2(index) / tovec2(size - 1);
        altitude = uv.y * (u.top_radius - u.bottom_radius);
        mu_s = uv.x * 2.0f - 1.0f;
    }
    // Atmosphere.glsl:169-178
    vec3 GetMultiscatteringContribution(const Image<4>& tex, float r, float mu_s) const {
        float x_mu_s = mu_s * 0.5f + 0.5f;
        float x_r = (r - u.bottom_radius) / (u.top_radius - u.bottom_radius);
        vec2 uv = GetTextureCoordFromUnitRange(vec2(x_mu_s, x_r), ivec2(tex.w, tex.h));
        return texture_linear(tex, uv, Sampler()).rgb();
    }

    // Atmosphere.glsl:180-188: sampler2DShadow lookup with Samplers::GetShadowMapSampler (Samplers.cpp:43-51): LINEAR,
    // CLAMP_TO_BORDER with border 1, compare LEQUAL -- the bilinear blend of the four comparison results (PCF).
    static float ShadowCompare(const Image<1>& shadow_map, float u, float v, float depth) {  // texture(sampler2DShadow, vec3(u, v, depth))
        float x = u * float(shadow_map.w) - 0.5f, y = v * float(shadow_map.h) - 0.5f;
        float fx = std::floor(x), fy = std::floor(y);
        float a = x - fx, b = y - fy;
        auto cmp = [&](float i, float j) {
            bool inside = i >= 0.0f && j >= 0.0f && i < float(shadow_map.w) && j < float(shadow_map.h);
            float texel = inside ? shadow_map.load(int(i), int(j)).x : 1.0f;
            return depth <= texel ? 1.0f : 0.0f;
        };
        return (1.0f - a) * (1.0f - b) * cmp(fx, fy) + a * (1.0f - b) * cmp(fx + 1.0f, fy) + (1.0f - a) * b * cmp(fx, fy + 1.0f) +
               a * b * cmp(fx + 1.0f, fy + 1.0f);
    }
    static float GetVisibilityFromShadowMap(const Image<1>& shadow_map, const mat4& light_view_projection, vec3 position) {
        vec4 xyzw = light_view_projection * vec4(position, 1.0f);
        vec3 xyz = vec3(xyzw.x, xyzw.y, xyzw.z) / xyzw.w;
        xyz = xyz * 0.5f + 0.5f;
        float depth = xyz.z;
        if (depth >= 1.0f) return 1.0f;
        return ShadowCompare(shadow_map, xyz.x, xyz.y, depth);
    }

    // Atmosphere.glsl:190-210
    static float GetVisibilityFromMoonShadow(float sun_moon_angular_distance, float sun_angular_radius, float moon_angular_radius) {
        float max_radius = sun_angular_radius + moon_angular_radius;
        float min_radius = std::fabs(sun_angular_radius - moon_angular_radius);
        float sun_r2 = sun_angular_radius * sun_angular_radius;
        float moon_r2 = moon_angular_radius * moon_angular_radius;
        if (sun_moon_angular_distance >= max_radius) return 1.0f;
        if (sun_moon_angular_distance <= min_radius) return clamp((sun_r2 - moon_r2) / sun_r2, 0.0f, 1.0f);
        float distance2 = sun_moon_angular_distance * sun_moon_angular_distance;
        float cos_half_sun = (distance2 + sun_r2 - moon_r2) / (2 * sun_moon_angular_distance * sun_angular_radius);
        float cos_half_moon = (distance2 + moon_r2 - sun_r2) / (2 * sun_moon_angular_distance * moon_angular_radius);
        float half_sun = sky_det_acosf(cos_half_sun);
        float half_moon = sky_det_acosf(cos_half_moon);
        float triangle_h = sun_angular_radius * std::sqrt(1 - cos_half_sun * cos_half_sun);
        float area_total = (PI - half_sun) * sun_r2 + (PI - half_moon) * moon_r2 + triangle_h * sun_moon_angular_distance;
        float area_uncovered = area_total - PI * moon_r2;
        return area_uncovered / (PI * sun_r2);
    }
    // Atmosphere.glsl:212-218
    float GetVisibilityFromMoonShadow(vec3 moon_vector, float moon_radius, vec3 sun_direction) const {
        float inv_moon_distance = 1.0f / std::sqrt(dot(moon_vector, moon_vector));
        vec3 moon_direction = moon_vector * inv_moon_distance;
        float moon_angular_radius = sky_det_asinf(clamp(moon_radius * inv_moon_distance, -1.0f, 1.0f));
        return GetVisibilityFromMoonShadow(sky_det_acosf(clamp(dot(sun_direction, moon_direction), -1.0f, 1.0f)), u.sun_angular_radius,
                                           moon_angular_radius);
    }

    // The two optional terms of the march (the host writes them into the shader text as #defines, AtmosphereRenderer.cpp:99-101)
    struct ScatterExtras {
        bool moon_shadow = false;          // MOON_SHADOW_ENABLE
        vec3 moon_position{0.0f};
        float moon_radius = 0.0f;
        const Image<1>* shadow_map = nullptr;  // VOLUMETRIC_LIGHT_ENABLE when non-null
        mat4 light_view_projection;
    };

    // Atmosphere.glsl:220-295
    template <bool MS>
    vec3 ComputeScatteredLuminance(const Image<4>& transmittance_texture, const Image<4>* multiscattering_texture,
                                   float start_i, vec3 earth_center, vec3 start_position, vec3 view_direction,
                                   vec3 sun_direction, float marching_distance, float steps, vec3& transmittance,
                                   vec3* L_f, const ScatterExtras* extras = nullptr) const {
        float r = length(start_position - earth_center);
        vec3 up_direction = normalize(start_position - earth_center);
        float mu = dot(view_direction, up_direction);
        float cos_sun_view = dot(view_direction, sun_direction);

        const float SAMPLE_COUNT = steps;
        float dx = marching_distance / SAMPLE_COUNT;

        transmittance = vec3(1.0f);
        vec3 luminance(0.0f);
        if (MS) { *L_f = vec3(0.0f); start_i = 0.5f; }
        float rayleigh_phase = RayleighPhaseFunction<MS>(cos_sun_view);
        float mie_phase = MiePhaseFunction<MS>(u.mie_phase_g, cos_sun_view);
        for (float i = start_i; i < SAMPLE_COUNT; ++i) {
            float d_i = i * dx;
            float r_i = std::sqrt(d_i * d_i + 2.0f * r * mu * d_i + r * r);
            vec3 position_i = start_position + view_direction * d_i;
            float altitude_i = r_i - u.bottom_radius;

            vec3 rayleigh_scattering_i, mie_scattering_i;
            GetScattering(altitude_i, rayleigh_scattering_i, mie_scattering_i);
            vec3 scattering_i = rayleigh_scattering_i + mie_scattering_i;
            vec3 scattering_with_phase_i = rayleigh_scattering_i * rayleigh_phase + mie_scattering_i * mie_phase;

            vec3 extinction_i = GetExtinction(altitude_i);
            vec3 transmittance_i = lut_exp3(-extinction_i * dx);
            vec3 up_direction_i = normalize(position_i - earth_center);
            float mu_s_i = dot(sun_direction, up_direction_i);
            vec3 luminance_i = scattering_with_phase_i * GetSunVisibility(transmittance_texture, r_i, mu_s_i);
            if (!MS) {
                if (extras && extras->shadow_map)  // :274-277
                    luminance_i *= GetVisibilityFromShadowMap(*extras->shadow_map, extras->light_view_projection, position_i);
                vec3 multiscattering_contribution = GetMultiscatteringContribution(*multiscattering_texture, r_i, mu_s_i);
                luminance_i += u.multiscattering_mask * multiscattering_contribution * scattering_i;
                if (extras && extras->moon_shadow)  // :281-284
                    luminance_i *= GetVisibilityFromMoonShadow(extras->moon_position - position_i, extras->moon_radius, sun_direction);
                luminance_i *= solar_illuminance();
            }
            luminance += transmittance * (luminance_i - luminance_i * transmittance_i) / extinction_i;
            if (MS) *L_f += transmittance * (scattering_i - scattering_i * transmittance_i) / extinction_i;
            transmittance *= transmittance_i;
        }
        return luminance;
    }

    // Atmosphere.glsl:297-306
    template <bool MS>
    vec3 ComputeGroundLuminance(const Image<4>& transmittance_texture, vec3 earth_center, vec3 position,
                                vec3 sun_direction) const {
        vec3 up_direction = normalize(position - earth_center);
        float mu_s = dot(sun_direction, up_direction);
        vec3 solar_illuminance_at_ground = GetSunVisibility(transmittance_texture, u.bottom_radius, mu_s);
        if (!MS) solar_illuminance_at_ground *= solar_illuminance();
        vec3 normal = normalize(position - earth_center);
        return INV_PI * clamp(dot(normal, sun_direction), 0.0f, 1.0f) * ground_albedo() * solar_illuminance_at_ground;
    }

    // Atmosphere.glsl:311-327
    vec3 ComputeTransmittanceToTopAtmosphereBoundary(float r, float mu) const {
        const float SAMPLE_COUNT = u.transmittance_steps;
        float dx = DistanceToTopAtmosphereBoundary(r, mu) / SAMPLE_COUNT;
        vec3 optical_length(0.0f);
        for (float i = 0.5f; i < SAMPLE_COUNT; ++i) {
            float d_i = i * dx;
            float r_i = std::sqrt(d_i * d_i + 2.0f * r * mu * d_i + r * r);
            float altitude_i = r_i - u.bottom_radius;
            optical_length += GetExtinction(altitude_i) * dx;
        }
        return lut_exp3(-optical_length);
    }

    // Atmosphere.glsl:344-355
    static vec3 GetDirectionFromLocalIndex(int index) {
        float unit_theta = (0.5f + float(index / 8)) / 8.0f;
        float unit_phi = (0.5f + float(index % 8)) / 8.0f;
        float cos_theta = 1.0f - 2.0f * unit_theta;
        float sin_theta = std::sqrt(clamp(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
        float phi = 2 * PI * unit_phi;
        return vec3(sky_det_cosf(phi) * sin_theta, cos_theta, sky_det_sinf(phi) * sin_theta);
    }

    // K1: Atmosphere.glsl:332-337 over 256x64 (Atmosphere.cpp:9-15)
    void BakeTransmittance(Image<4>& out) const;
    // K2: Atmosphere.glsl:364-438 over 32x32x64 (Atmosphere.cpp:17-19,121)
    void BakeMultiscattering(const Image<4>& transmittance, Image<4>& out) const;
};

struct CubeChain;  // ibl.h

// Inputs of the object branch of K6 (AtmosphereRenderer.glsl:284-343,404-410): the G-buffer the host binds
// (AtmosphereRenderer.h:37-41; formats GBuffer.cpp:19-21), the IBL chain (ibl.h) and the cloud shadow map (K12 output).
struct ObjectShading {
    const uint8_t* albedo = nullptr;    // GL_RGBA8        [H][W][4]
    const int16_t* normal = nullptr;    // GL_RGBA16_SNORM [H][W][4]
    const uint16_t* orm = nullptr;      // GL_RGBA16       [H][W][4]
    const Image<2>* env_brdf_lut = nullptr;
    const CubeChain* prefiltered = nullptr;
    const vec4* Llm = nullptr;
    const Image<2>* cloud_shadow_map = nullptr;  // shadow_maps_[2]; null before the first shadow pass (visibility 1)
};

// AtmosphereRenderer.glsl K3-K6.
struct AtmosphereRenderer {
    const Atmosphere& atm;
    SkyAtmosphereRenderBufferData u;  // AtmosphereRenderer.glsl:25-50
    SkyLutConfig cfg;
    const Image<4>& transmittance_texture;
    const Image<4>& multiscattering_texture;
    const Image<1>* blue_noise = nullptr;  // R16 64x64
    int out_band_rows = 0, out_band_index = 0, out_band_count = 1;  // sky_set_output_bands: rows the composite owns
    const Image<4>* star_map = nullptr;  // GL_SRGB8 star map decoded to linear RGB (Textures.cpp:43-50); null: no star term
    const Image<1>* mesh_shadow_map = nullptr;  // DEPTH32F 2048^2 (ShadowMap.cpp:8-27), used when cfg.volumetric_light
    const ObjectShading* object = nullptr;      // null: object pixels keep the in-scatter alone and alpha 0
    // AtmosphereRenderer.glsl:284-324 and :333-343
    vec3 ComputeObjectLuminance(vec3 position, vec3 view_direction, float shadow_visibility, vec2 vTexCoord, int width, int height) const;
    float SampleVisibilityFromShadowMap(vec3 position) const;
    float PCSS(const Image<1>& shadow_map, vec3 position) const;  // Shadow.glsl:85-99

    // AtmosphereRenderer.glsl:326-331 (sampler LinearNoMipmapClampToEdge, AtmosphereRenderer.cpp:204)
    vec3 GetStarLuminance(vec3 view_direction) const {
        float theta = sky_det_acosf(view_direction.y);
        float phi = std::atan2(view_direction.x, view_direction.z);
        vec2 coord(INV_PI * 0.5f * phi + 0.5f, 1.0f - theta * INV_PI);
        return u.star_luminance_scale * texture_linear(*star_map, coord, Sampler()).rgb();
    }

    Atmosphere::ScatterExtras extras() const {
        Atmosphere::ScatterExtras e;
        e.moon_shadow = cfg.moon_shadow != 0;
        e.moon_position = vec3(u.moon_position);
        e.moon_radius = u.moon_radius;
        e.shadow_map = cfg.volumetric_light ? mesh_shadow_map : nullptr;
        e.light_view_projection = mat4(u.light_view_projection);
        return e;
    }

    vec3 sun_direction() const { return vec3(u.sun_direction); }
    vec3 earth_center() const { return vec3(u.earth_center); }
    vec3 camera_position() const { return vec3(u.camera_position); }
    vec3 up_direction() const { return vec3(u.up_direction); }
    vec3 right_direction() const { return vec3(u.right_direction); }
    vec3 front_direction() const { return vec3(u.front_direction); }
    ivec2 sky_size() const { return ivec2(cfg.sky_view_width, cfg.sky_view_height); }

    // AtmosphereRenderer.glsl:56-72
    vec3 ComputeRaymarchingStartPositionAndChangeDistance(vec3 view_direction, float& marching_distance) const {
        vec3 start_position = camera_position();
        bool in_space = u.camera_earth_center_distance > atm.u.top_radius;
        if (in_space) {
            float r = u.camera_earth_center_distance;
            float mu = dot(view_direction, up_direction());
            float near_distance;
            if (atm.FromSpaceIntersectTopAtmosphereBoundary(r, mu, near_distance)) {
                start_position += near_distance * view_direction;
                marching_distance -= near_distance;
            } else {
                marching_distance = 0;
            }
        }
        return start_position;
    }
    // AtmosphereRenderer.glsl:74-78
    float GetHorizonDownAngleFromR(float r) const {
        float tangent_point_distance = std::sqrt(r * r - atm.u.bottom_radius * atm.u.bottom_radius);
        float cos_horizon_down = tangent_point_distance / r;
        return sky_det_acosf(cos_horizon_down);
    }
    // AtmosphereRenderer.glsl:81-102
    void GetCosLatLonFromSkyViewTextureIndex(ivec2 index, float r, float& cos_lat, float& cos_lon) const {
        vec2 uv = tovec2(index) / tovec2(sky_size() - 1);
        float x_cos_lon = uv.x, x_cos_lat = uv.y;
        float horizon_down_angle = GetHorizonDownAngleFromR(r);
        float horizon_up_angle = PI - horizon_down_angle;
        float lat;
        if (x_cos_lat < 0.5f) {
            float coord = 1.0f - 2.0f * x_cos_lat;
            coord = 1.0f - coord * coord;
            lat = horizon_up_angle * coord;
        } else {
            float coord = x_cos_lat * 2.0f - 1.0f;
            coord *= coord;
            lat = horizon_up_angle + horizon_down_angle * coord;
        }
        cos_lat = sky_det_cosf(lat);
        cos_lon = -(x_cos_lon * x_cos_lon * 2.0f - 1.0f);
    }
    // AtmosphereRenderer.glsl:104-111
    vec3 GetViewDirectionFromCosLatLon(float cos_lat, float cos_lon) const {
        float sin_lat = clamp(std::sqrt(1 - cos_lat * cos_lat), 0.0f, 1.0f);
        float sin_lon = clamp(std::sqrt(1 - cos_lon * cos_lon), 0.0f, 1.0f);
        return up_direction() * cos_lat + front_direction() * (sin_lat * cos_lon) + right_direction() * (sin_lat * sin_lon);
    }
    // AtmosphereRenderer.glsl:113-132
    vec2 GetSkyViewTextureUvFromCosLatLon(float r, float cos_lat, float cos_lon) const {
        float horizon_down_angle = GetHorizonDownAngleFromR(r);
        float horizon_up_angle = PI - horizon_down_angle;
        // GLSL leaves acos(|x|>1) and sqrt(x<0) undefined; rounding in dot()/normalize() can push the
        // arguments a few ulp outside, so the oracle (and the CUDA kernels) clamp them.
        float lat = sky_det_acosf(clamp(cos_lat, -1.0f, 1.0f));
        float x_cos_lat;
        if (lat < horizon_up_angle) {
            float coord = lat / horizon_up_angle;
            coord = std::sqrt(max(1 - coord, 0.0f));
            x_cos_lat = 0.5f - 0.5f * coord;
        } else {
            float coord = (lat - horizon_up_angle) / horizon_down_angle;
            coord = std::sqrt(max(coord, 0.0f));
            x_cos_lat = coord * 0.5f + 0.5f;
        }
        float x_cos_lon = std::sqrt(max(0.5f - 0.5f * cos_lon, 0.0f));
        return Atmosphere::GetTextureCoordFromUnitRange(vec2(x_cos_lon, x_cos_lat), sky_size());
    }
    // AtmosphereRenderer.glsl:134-145
    void GetCosLatLonFromViewDirection(vec3 view_direction, float& cos_lat, float& cos_lon) const {
        cos_lat = dot(up_direction(), view_direction);
        vec3 lon_direction = view_direction - up_direction() * cos_lat;
        float lon_direction_length2 = dot(lon_direction, lon_direction);
        if (lon_direction_length2 == 0) {
            cos_lon = 1;
        } else {
            lon_direction *= inversesqrt(lon_direction_length2);
            cos_lon = dot(lon_direction, front_direction());
        }
    }

    float DitherStart(bool enable, int x, int y) const {
        if (enable && blue_noise) return blue_noise->at(x & 0x3f, y & 0x3f)[0];
        return 0.5f;
    }

    // K3: AtmosphereRenderer.glsl:153-186
    void BakeSkyView(Image<4>& luminance_image, Image<4>& transmittance_image) const;
    // K4: AtmosphereRenderer.glsl:191-243
    void BakeAerialPerspective(Image<4>& luminance_image, Image<4>& transmittance_image) const;
    // K5: AtmosphereRenderer.glsl:253-273 (cube faces stacked along z; values rounded to fp16)
    void BakeEnvironment(const Image<4>& sky_lum, const Image<4>& sky_trans, Image<4>& env) const;
    // K6: AtmosphereRenderer.glsl:345-432, sky / aerial-perspective / sun-disc branches
    void Composite(const Image<4>& sky_lum, const Image<4>& sky_trans, const Image<4>& ap_lum,
                   const Image<4>& ap_trans, const Image<1>* shadow_froxel, const float* depth, int width,
                   int height, uint16_t* hdr_half4) const;
};

// shaders/Base/Common.glsl:13-30
vec3 ConvertCubUvToDir(int index, vec2 uv);
// VolumetricCloudShadowInterface.glsl:10-13; froxel = R16 unorm, LinearNoMipmapClampToEdge
float SampleRayScatterVisibility(const Image<1>& shadow_froxel, vec2 uv, float dist, float inv_max_dist);

}  // namespace orc
