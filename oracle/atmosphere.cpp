// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h).  K1-K6 entry points.
#include "atmosphere.h"

namespace orc {

// K1 -- Atmosphere.glsl:332-337
void Atmosphere::BakeTransmittance(Image<4>& out) const {
    ivec2 size(out.w, out.h);
#pragma omp parallel for schedule(dynamic)
    for (int y = 0; y < out.h; ++y)
        for (int x = 0; x < out.w; ++x) {
            float r, mu;
            GetRMuFromTransmittanceTextureIndex(ivec2(x, y), size, r, mu);
            vec3 transmittance = ComputeTransmittanceToTopAtmosphereBoundary(r, mu);
            out.store(x, y, vec4(transmittance, 1.0f));
        }
}

// K2 -- Atmosphere.glsl:364-438
void Atmosphere::BakeMultiscattering(const Image<4>& transmittance_texture, Image<4>& out) const {
    ivec2 size(out.w, out.h);
#pragma omp parallel for schedule(dynamic)
    for (int gy = 0; gy < out.h; ++gy)
        for (int gx = 0; gx < out.w; ++gx) {
            float altitude, mu_s;
            GetAltitudeMuSFromMultiscatteringTextureIndex(ivec2(gx, gy), size, altitude, mu_s);
            vec3 earth_center(0, -u.bottom_radius, 0);
            vec3 start_position(0, altitude, 0);
            vec3 sun_direction(0, mu_s, std::sqrt(1 - mu_s * mu_s));
            vec3 L_2nd_order_shared[64];
            vec3 f_ms_shared[64];
            for (int local_index = 0; local_index < 64; ++local_index) {
                vec3 view_direction = GetDirectionFromLocalIndex(local_index);
                float r = altitude + u.bottom_radius;
                float mu = view_direction.y;
                bool intersect_bottom = RayIntersectsGround(r, mu);
                float marching_distance =
                    intersect_bottom ? DistanceToBottomAtmosphereBoundary(r, mu) : DistanceToTopAtmosphereBoundary(r, mu);
                vec3 transmittance, L_f;
                vec3 luminance = ComputeScatteredLuminance<true>(transmittance_texture, nullptr, 0.5f, earth_center,
                                                                 start_position, view_direction, sun_direction,
                                                                 marching_distance, u.multiscattering_steps, transmittance, &L_f);
                if (intersect_bottom) {
                    vec3 ground_position = start_position + view_direction * marching_distance;
                    luminance += transmittance * ComputeGroundLuminance<true>(transmittance_texture, earth_center,
                                                                              ground_position, sun_direction);
                }
                L_2nd_order_shared[local_index] = luminance;
                f_ms_shared[local_index] = L_f;
            }
            // pairwise tree (i, i+32) ... (i, i+1), Atmosphere.glsl:395-425
            for (int stride = 32; stride >= 1; stride >>= 1)
                for (int i = 0; i < stride; ++i) {
                    L_2nd_order_shared[i] += L_2nd_order_shared[i + stride];
                    f_ms_shared[i] += f_ms_shared[i + stride];
                }
            vec3 L_2nd_order = L_2nd_order_shared[0] / 64.0f;
            vec3 f_ms = f_ms_shared[0] / 64.0f;
            vec3 F_ms = 1.0f / (1.0f - f_ms);
            out.store(gx, gy, vec4(L_2nd_order * F_ms, 1.0f));
        }
}

// K3 -- AtmosphereRenderer.glsl:153-186
void AtmosphereRenderer::BakeSkyView(Image<4>& luminance_image, Image<4>& transmittance_image) const {
    const Atmosphere::ScatterExtras ex = extras();
#pragma omp parallel for schedule(dynamic)
    for (int y = 0; y < cfg.sky_view_height; ++y)
        for (int x = 0; x < cfg.sky_view_width; ++x) {
            float r = u.camera_earth_center_distance;
            float cos_lat, cos_lon;
            GetCosLatLonFromSkyViewTextureIndex(ivec2(x, y), r, cos_lat, cos_lon);
            vec3 view_direction = GetViewDirectionFromCosLatLon(cos_lat, cos_lon);
            float mu = cos_lat;
            bool intersect_bottom = atm.RayIntersectsGround(r, mu);
            float marching_distance = intersect_bottom ? atm.DistanceToBottomAtmosphereBoundary(r, mu)
                                                       : atm.DistanceToTopAtmosphereBoundary(r, mu);
            vec3 start_position = ComputeRaymarchingStartPositionAndChangeDistance(view_direction, marching_distance);
            vec3 transmittance(1.0f), luminance(0.0f);
            if (marching_distance > 0) {
                float start_i = DitherStart(cfg.sky_view_dither != 0, x, y);
                luminance = atm.ComputeScatteredLuminance<false>(
                    transmittance_texture, &multiscattering_texture, start_i, earth_center(), start_position,
                    view_direction, sun_direction(), marching_distance, u.sky_view_lut_steps, transmittance, nullptr, &ex);
            }
            luminance_image.store(x, y, vec4(luminance, 0.0f));
            transmittance_image.store(x, y, vec4(transmittance, 0.0f));
        }
}

// K4 -- AtmosphereRenderer.glsl:191-243
void AtmosphereRenderer::BakeAerialPerspective(Image<4>& luminance_image, Image<4>& transmittance_image) const {
    const ivec3 size(luminance_image.w, luminance_image.h, luminance_image.d);
    const mat4 inv_view_projection(u.inv_view_projection);
    const Atmosphere::ScatterExtras ex = extras();
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int z = 0; z < size.z; ++z)
        for (int y = 0; y < size.y; ++y)
            for (int x = 0; x < size.x; ++x) {
                // GetMarchingDistanceFromAerialPerspectiveTextureIndex, :191-197
                vec3 uvw = vec3(float(x), float(y), float(z)) / vec3(float(size.x - 1), float(size.y - 1), float(size.z - 1));
                vec3 position = ProjectiveMul(inv_view_projection, vec3(uvw.xy() * 2.0f - 1.0f, 0.0f));
                vec3 view_direction = normalize(position - camera_position());
                float marching_distance = uvw.z * uvw.z * u.aerial_perspective_lut_max_distance;

                float r = u.camera_earth_center_distance;
                float mu = dot(view_direction, up_direction());
                bool intersect_bottom = atm.RayIntersectsGround(r, mu);
                float max_marching_distance = intersect_bottom ? marching_distance : atm.DistanceToTopAtmosphereBoundary(r, mu);
                vec3 start_position = ComputeRaymarchingStartPositionAndChangeDistance(view_direction, max_marching_distance);
                marching_distance = min(marching_distance, max_marching_distance);
                vec3 transmittance(1.0f), luminance(0.0f);
                if (marching_distance > 0) {
                    float start_i = DitherStart(cfg.aerial_perspective_dither != 0, x, y);
                    luminance = atm.ComputeScatteredLuminance<false>(
                        transmittance_texture, &multiscattering_texture, start_i, earth_center(), start_position,
                        view_direction, sun_direction(), marching_distance, u.aerial_perspective_lut_steps, transmittance,
                        nullptr, &ex);
                }
                luminance_image.store(x, y, z, vec4(luminance, 0.0f));
                transmittance_image.store(x, y, z, vec4(transmittance, 0.0f));
            }
}

// shaders/Base/Common.glsl:13-30
vec3 ConvertCubUvToDir(int index, vec2 uv) {
    float uc = 2.0f * uv.x - 1.0f;
    float vc = 2.0f * uv.y - 1.0f;
    vec3 dir(0.0f);
    switch (index) {
        case 0: dir = vec3(1.0f, vc, -uc); break;
        case 1: dir = vec3(-1.0f, vc, uc); break;
        case 2: dir = vec3(uc, 1.0f, -vc); break;
        case 3: dir = vec3(uc, -1.0f, vc); break;
        case 4: dir = vec3(uc, vc, 1.0f); break;
        case 5: dir = vec3(-uc, vc, -1.0f); break;
    }
    return normalize(dir);
}

// K5 -- AtmosphereRenderer.glsl:253-273
void AtmosphereRenderer::BakeEnvironment(const Image<4>& sky_lum, const Image<4>& sky_trans, Image<4>& env) const {
    const int S = env.w;
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int index = 0; index < 6; ++index)
        for (int y = 0; y < S; ++y)
            for (int x = 0; x < S; ++x) {
                vec2 face_uv = (vec2(float(x), float(y)) + vec2(0.5f)) / vec2(float(S), float(S));
                face_uv.y = 1.0f - face_uv.y;
                vec3 view_direction = ConvertCubUvToDir(index, face_uv);
                float cos_lat, cos_lon;
                GetCosLatLonFromViewDirection(view_direction, cos_lat, cos_lon);
                float r = u.camera_earth_center_distance;
                vec2 uv = GetSkyViewTextureUvFromCosLatLon(r, cos_lat, cos_lon);
                vec3 luminance = texture_linear(sky_lum, uv, Sampler()).rgb();
                vec3 transmittance = texture_linear(sky_trans, uv, Sampler()).rgb();
                float mu = cos_lat;
                if (atm.RayIntersectsGround(r, mu)) {
                    float marching_distance = atm.DistanceToBottomAtmosphereBoundary(r, mu);
                    vec3 ground_position = camera_position() + view_direction * marching_distance;
                    luminance += atm.ComputeGroundLuminance<false>(transmittance_texture, earth_center(), ground_position,
                                                                   sun_direction()) * transmittance;
                }
                // rgba16f image store
                env.store(x, y, index, vec4(to_half_and_back(luminance.x), to_half_and_back(luminance.y),
                                            to_half_and_back(luminance.z), 0.0f));
            }
}

// VolumetricCloudShadowInterface.glsl:10-13
float SampleRayScatterVisibility(const Image<1>& shadow_froxel, vec2 uv, float dist, float inv_max_dist) {
    float w = dist * inv_max_dist;
    float f = texture_linear(shadow_froxel, vec3(uv, w), Sampler()).x;
    return mix(1.0f, f, clamp(1.0f / w, 0.0f, 1.0f));
}

// K6 -- AtmosphereRenderer.glsl:345-432.  gl_FragCoord.xy = pixel + 0.5, vTexCoord = that / size.
// Object pixels (depth != 1) get the in-scatter only and alpha 0: ComputeObjectLuminance
// (:284-324) needs the G-buffer + IBL chain that SURVEY.md 8f-1 leaves for later.
void AtmosphereRenderer::Composite(const Image<4>& sky_lum, const Image<4>& sky_trans, const Image<4>& ap_lum,
                                   const Image<4>& ap_trans, const Image<1>* shadow_froxel, const float* depth_img,
                                   int width, int height, uint16_t* hdr) const {
    const mat4 inv_view_projection(u.inv_view_projection);
    const ivec3 ap_size(ap_lum.w, ap_lum.h, ap_lum.d);
    const Atmosphere::ScatterExtras ex = extras();
#pragma omp parallel for schedule(dynamic)
    for (int py = 0; py < height; ++py)
        for (int px = 0; px < width; ++px) {
            if (out_band_count > 1 && (py / out_band_rows) % out_band_count != out_band_index) continue;  // sky_set_output_bands
            vec2 vTexCoord((float(px) + 0.5f) / float(width), (float(py) + 0.5f) / float(height));
            float depth = depth_img[size_t(py) * width + px];
            vec3 fragment_position = ProjectiveMul(inv_view_projection, vec3(vTexCoord, depth) * 2.0f - 1.0f);
            vec3 view_direction = normalize(fragment_position - camera_position());
            float r = u.camera_earth_center_distance;
            float mu = dot(view_direction, up_direction());
            float marching_distance = atm.RayIntersectsGround(r, mu) ? atm.DistanceToBottomAtmosphereBoundary(r, mu)
                                                                     : atm.DistanceToTopAtmosphereBoundary(r, mu);
            bool intersect_object = false;
            if (depth != 1.0f) {
                float object_distance = length(fragment_position - camera_position());
                intersect_object = true;
                marching_distance = min(marching_distance, object_distance);
            }
            vec3 start_position = ComputeRaymarchingStartPositionAndChangeDistance(view_direction, marching_distance);
            vec3 transmittance(1.0f), luminance(0.0f);
            if (marching_distance > 0) {
                float start_i = DitherStart(cfg.raymarching_dither != 0, px, py);
                if (cfg.use_sky_view_lut && !intersect_object) {
                    float cos_lat, cos_lon;
                    GetCosLatLonFromViewDirection(view_direction, cos_lat, cos_lon);
                    vec2 uv = GetSkyViewTextureUvFromCosLatLon(r, cos_lat, cos_lon);
                    luminance = texture_linear(sky_lum, uv, Sampler()).rgb();
                    transmittance = texture_linear(sky_trans, uv, Sampler()).rgb();
                } else if (cfg.use_aerial_perspective_lut && intersect_object) {
                    // GetAerialPerspectiveTextureUvwFromTexCoordDistance, AtmosphereInterface.glsl:19-23
                    vec3 uvw = Atmosphere::GetTextureCoordFromUnitRange(
                        vec3(vTexCoord, std::sqrt(marching_distance / u.aerial_perspective_lut_max_distance)), ap_size);
                    luminance = texture_linear(ap_lum, uvw, Sampler()).rgb();
                    transmittance = texture_linear(ap_trans, uvw, Sampler()).rgb();
                } else {
                    luminance = atm.ComputeScatteredLuminance<false>(
                        transmittance_texture, &multiscattering_texture, start_i, earth_center(), start_position,
                        view_direction, sun_direction(), marching_distance, u.raymarching_steps, transmittance, nullptr, &ex);
                }
            }
            if (shadow_froxel)
                luminance *= SampleRayScatterVisibility(*shadow_froxel, vTexCoord, marching_distance, u.uInvShadowFroxelMaxDistance);

            float alpha = 1.0f;
            if (intersect_object) {
                alpha = 0.0f;
            } else if (dot(view_direction, sun_direction()) >= std::cos(atm.u.sun_angular_radius)) {
                vec3 uu(1.0f, 1.0f, 1.0f);
                vec3 a(0.397f, 0.503f, 0.652f);
                float cos_view_sun = dot(view_direction, sun_direction());
                float sin_view_sun = std::sqrt(1.0f - cos_view_sun * cos_view_sun);
                float center_to_edge = clamp(sin_view_sun / std::sin(atm.u.sun_angular_radius), 0.0f, 1.0f);
                float mu2 = std::sqrt(1.0f - center_to_edge * center_to_edge);
                vec3 factor = vec3(1.0f) - uu * (vec3(1.0f) - pow(vec3(mu2), a));
                vec3 solar_illuminance_at_eye = atm.solar_illuminance() * transmittance;
                luminance += solar_illuminance_at_eye / (PI * atm.u.sun_angular_radius * atm.u.sun_angular_radius) * factor;
            } else if (star_map) {  // :427-429
                luminance += transmittance * GetStarLuminance(view_direction);
            }
            uint16_t* o = hdr + (size_t(py) * width + px) * 4;
            o[0] = float_to_half_bits(luminance.x);
            o[1] = float_to_half_bits(luminance.y);
            o[2] = float_to_half_bits(luminance.z);
            o[3] = float_to_half_bits(alpha);
        }
}

}  // namespace orc
