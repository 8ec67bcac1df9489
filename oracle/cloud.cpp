// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h).  Materials + K11-K18.
#include "cloud.h"

namespace orc {

void CloudScene::SetViewport(int w, int h) {
    // VolumetricCloud.cpp:120-136; histories start at zero (SURVEY.md section 7 "frame-0 state")
    width = w; height = h;
    checkerboard_depth.resize(w / 2, h / 2);
    index_linear_depth.resize(w / 4, h / 4);
    render_texture.resize(w / 4, h / 4);
    cloud_distance.resize(w / 4, h / 4);
    reconstruct[0].resize(w / 2, h / 2);
    reconstruct[1].resize(w / 2, h / 2);
    shadow_froxel.resize(w / 12, h / 12, 128);
    for (auto& m : shadow_maps) m.resize(512, 512);  // VolumetricCloud.cpp:52,102-105
}

// ---------------------------------------------------------------- materials
namespace {
const Sampler kRepeat = [] { Sampler s; s.wrap = REPEAT; return s; }();
const Sampler kBorder0 = [] { Sampler s; s.wrap = CLAMP_TO_BORDER; s.border = vec4(0.0f); return s; }();

// VolumetricCloudDefaultMaterialCommon.glsl:20-24
inline vec4 GetUVWLod(vec3 pos, const SkySampleInfo& info, vec3 camera_pos, float lod_bias) {
    vec3 uvw = pos * info.frequency + vec3(info.bias[0], info.bias[1], 0.0f);
    float lod = std::log2(info.k_lod * distance(pos, camera_pos)) + lod_bias;
    return vec4(uvw, lod);
}
// VolumetricCloudDefaultMaterial0.glsl:9-16
inline float CalHeightMask(float cloud_type, float height01) {
    float height_in_type = clamp(height01 / cloud_type, 0.0f, 1.0f);
    return clamp(height_in_type * (height_in_type - 1.0f) * -4.0f, 0.0f, 1.0f);
}
inline float Remap01(float x, float x0, float x1) { return clamp((x - x0) / (x1 - x0), 0.0f, 1.0f); }
}  // namespace

float CloudScene::SampleSigmaT(vec3 pos, float height01, int slot) {
    uint64_t fetches = 0;
    float result = 0.0f;
    switch (material.type) {
        case SKY_MATERIAL_DEFAULT0: {  // VolumetricCloudDefaultMaterial0.glsl:18-32
            const auto& mc = material.common;
            const auto& m = material.u.m0;
            vec4 uvwlod = GetUVWLod(pos, mc.uCloudMapSampleInfo, uCameraPos(), mc.uLodBias);
            vec4 cloud_type = cloud_map.texture_lod(vec2(uvwlod.x, uvwlod.y), uvwlod.w, kRepeat);
            vec3 displace_vector(0.0f);
            uvwlod = GetUVWLod(pos, mc.uDisplacementSampleInfo, uCameraPos(), mc.uLodBias);
            vec4 d0 = displacement.texture_lod(vec2(uvwlod.x, uvwlod.y), uvwlod.w, kRepeat);
            displace_vector.x += d0.x; displace_vector.y += d0.y;
            vec4 d1 = displacement.texture_lod(vec2(uvwlod.x, uvwlod.z), uvwlod.w, kRepeat);
            displace_vector.x += d1.z; displace_vector.z += d1.w;
            pos += m.uDisplacementScale * displace_vector;
            uvwlod = GetUVWLod(pos, mc.uDetailSampleInfo, uCameraPos(), mc.uLodBias);
            float det = detail.texture_lod(uvwlod.xyz(), uvwlod.w, kRepeat).x;
            det = det * m.uDetailParam[0] + m.uDetailParam[1];
            result = Remap01(cloud_type.x * CalHeightMask(cloud_type.y, height01), det, 1.0f) * height01 * mc.uDensity;
            fetches = 4;
            break;
        }
        case SKY_MATERIAL_DEFAULT1: {  // VolumetricCloudDefaultMaterial1.glsl:14-29
            const auto& mc = material.common;
            const auto& m = material.u.m1;
            vec4 uvwlod = GetUVWLod(pos, mc.uCloudMapSampleInfo, uCameraPos(), mc.uLodBias);
            vec4 cloud_type = cloud_map.texture_lod(vec2(uvwlod.x, uvwlod.y), uvwlod.w, kRepeat);
            fetches = 1;
            float density = clamp((cloud_type.x - m.uBaseDensityThreshold) * m.uBaseEdgeHardness, 0.0f, 1.0f);
            density *= clamp((1 - height01) * m.uBaseHeightHardness, 0.0f, 1.0f);
            if (density == 0) { result = 0; break; }
            uvwlod = GetUVWLod(pos, mc.uDetailSampleInfo, uCameraPos(), mc.uLodBias);
            float det = detail.texture_lod(uvwlod.xyz(), uvwlod.w, kRepeat).x;
            fetches = 2;
            det = (det + m.uDetailBase) * m.uDetailScale;
            det *= max(clamp(height01 - m.uHeightCut, 0.0f, 1.0f), clamp(m.uEdgeCur - cloud_type.x, 0.0f, 1.0f));
            result = clamp(density - det, 0.0f, 1.0f) * mc.uDensity * height01;
            break;
        }
        case SKY_MATERIAL_MINIMAL:  // VolumetricCloudMaterialMinimal.glsl:6-8
            result = material.u.minimal.uDensity;
            break;
        case SKY_MATERIAL_VOXEL: {  // VolumetricCloudMaterialVoxel.glsl:12-17
            const auto& m = material.u.voxel;
            vec2 uv = pos.xy() * vec2(m.uSampleFrequency[0], m.uSampleFrequency[1]) + vec2(m.uSampleBias[0], m.uSampleBias[1]);
            float lod = std::log2(m.uSampleLodK * distance(pos, uCameraPos())) + m.uLodBias;
            float density = voxel.texture_lod(vec3(uv, height01), lod, kBorder0).x;
            result = density * m.uDensity;
            fetches = 1;
            break;
        }
    }
    if (counting) {
        counters[slot].fetch_add(1, std::memory_order_relaxed);
        if (slot == SKY_CNT_RENDER_SIGMA_EVALS) counters[SKY_CNT_RENDER_TEX_FETCHES].fetch_add(fetches, std::memory_order_relaxed);
    }
    return result;
}

// VolumetricCloudCommon.glsl:81-97; AP sampler = LinearNoMipmapClampToEdge (VolumetricCloud.cpp:373-375)
vec3 CloudScene::GetAerialPerspective(vec2 uv, float t, float r, float mu, vec3& transmittance_out) const {
    if (r > atm.u.top_radius) {
        float near_distance;
        if (atm.FromSpaceIntersectTopAtmosphereBoundary(r, mu, near_distance)) t -= near_distance;
        else t = 0;
    }
    vec3 uvw = Atmosphere::GetTextureCoordFromUnitRange(vec3(uv, std::sqrt(t / c.uAerialPerspectiveLutMaxDistance)),
                                                        ivec3(ap_lum.w, ap_lum.h, ap_lum.d));
    transmittance_out = texture_linear(ap_trans, uvw, Sampler()).rgb();
    return texture_linear(ap_lum, uvw, Sampler()).rgb();
}

float CloudScene::SampleCloudShadowTransmittance(const Image<2>& cloud_shadow_map, vec3 light_ndc) const {
    Sampler s; s.wrap = CLAMP_TO_BORDER; s.border = vec4(1e10f, 1.0f, 0.0f, 0.0f);
    const float kInvTransitionDepth = 1.0f / 0.5f;
    vec4 dt = texture_linear(cloud_shadow_map, light_ndc.xy() * 0.5f + 0.5f, s);
    return mix(dt.y, 1.0f, clamp((dt.x - light_ndc.z) * kInvTransitionDepth, 0.0f, 1.0f));
}

// ---------------------------------------------------------------- K11
void CloudScene::ShadowMap() {
    std::swap(shadow_maps[0], shadow_maps[1]);  // VolumetricCloud.cpp:284
    Image<2>& out = shadow_maps[0];
    const Image<2>& pre = shadow_maps[1];
    const mat4 uInvLightVP(c.uInvLightVP), uShadowMapReprojectMat(c.uShadowMapReprojectMat);
    Sampler pre_sampler; pre_sampler.wrap = CLAMP_TO_BORDER; pre_sampler.border = vec4(1e10f, 1.0f, 0.0f, 0.0f);
    const vec2 image_size(float(out.w), float(out.h));
#pragma omp parallel for schedule(dynamic)
    for (int gy = 0; gy < out.h; ++gy)
        for (int gx = 0; gx < out.w; ++gx) {
            vec2 uv = (vec2(float(gx), float(gy)) + 0.5f) / image_size;
            vec3 origin = ProjectiveMul(uInvLightVP, vec3(uv * 2.0f - 1.0f, 0.0f));
            vec3 dir = -uSunDirection();
            vec3 up(origin.x, origin.y, origin.z + c.uEarthRadius);
            float r = length(up);
            up /= r;
            float mu = dot(up, dir);
            // LineShellFirstIntersect, VolumetricCloudShadowMap.comp:18-35
            float t1 = 0, t2 = 0;
            {
                float bottom_radius = c.uEarthRadius + c.uBottomAltitude;
                float top_radius = c.uEarthRadius + c.uTopAltitude;
                float discriminant_top = r * r * (mu * mu - 1.0f) + top_radius * top_radius;
                if (discriminant_top > 0) {
                    float discriminant_bottom = r * r * (mu * mu - 1.0f) + bottom_radius * bottom_radius;
                    float sqrt_discriminant_top = std::sqrt(discriminant_top);
                    float sqrt_discriminant_bottom = std::sqrt(discriminant_bottom);
                    t1 = -r * mu - sqrt_discriminant_top;
                    t2 = -r * mu + (discriminant_bottom >= 0 ? -sqrt_discriminant_bottom : sqrt_discriminant_top);
                }
            }
            float dist = max(t2 - t1, 0.0f);
            dist = min(dist, 1.0f / std::cos(radians(85.0f)) * (c.uTopAltitude - c.uBottomAltitude));

            float optical_depth = 0.0f;
            if (dist > 0.0f) {
                float steps = mix(12.0f, 6.0f, std::fabs(dir.z));
                float step_size = dist / steps;
                float noise = blue_noise.at(gx & 0x3f, gy & 0x3f)[0];
                float t = t1 + step_size * fract(noise + c.uFrameID * 0.61803398875f);
                for (uint cnt = uint(steps); cnt != 0; cnt--, t += step_size) {
                    vec3 pos = origin + t * dir;
                    float height01 = CalHeight01(pos);
                    float sigma_t = SampleSigmaT(pos, height01, SKY_CNT_SHADOW_SIGMA_EVALS);
                    optical_depth += sigma_t * step_size;
                }
            }
            float transmittance = std::exp(-optical_depth);
            float depth = mix(t1, t2, 0.5f);
            vec2 res(depth, transmittance);

            vec2 pre_uv = ProjectiveMul(uShadowMapReprojectMat, vec3(uv * 2.0f - 1.0f, 0.0f)).xy() * 0.5f + 0.5f;
            vec2 lo = vec2(0.5f) / image_size, hi = (image_size - vec2(0.5f)) / image_size;
            vec2 cl = clamp(pre_uv, lo, hi);
            if (cl.x == pre_uv.x && cl.y == pre_uv.y) {
                vec4 pre_res = texture_linear(pre, pre_uv, pre_sampler);
                res = mix(vec2(pre_res.x, pre_res.y), res, 0.2f);
            }
            out.at(gx, gy)[0] = res.x;
            out.at(gx, gy)[1] = res.y;
        }
}

// ---------------------------------------------------------------- K12
void CloudScene::ShadowBlur() {
    static const float weight[5] = {0.227027f, 0.1945946f, 0.1216216f, 0.054054f, 0.016216f};
    auto pass = [&](const Image<2>& in, Image<2>& out, bool horizontal) {
#pragma omp parallel for schedule(static)
        for (int y = 0; y < out.h; ++y)
            for (int x = 0; x < out.w; ++x) {
                vec2 res = vec2(in.at(x, y)[0], in.at(x, y)[1]) * weight[0];
                for (int i = 1; i < 5; ++i) {
                    int x1 = x, y1 = y, x2 = x, y2 = y;
                    if (horizontal) { x1 = max(x - i, 0); x2 = min(x + i, in.w - 1); }
                    else { y1 = max(y - i, 0); y2 = min(y + i, in.h - 1); }
                    res += vec2(in.at(x1, y1)[0], in.at(x1, y1)[1]) * weight[i];
                    res += vec2(in.at(x2, y2)[0], in.at(x2, y2)[1]) * weight[i];
                }
                out.at(x, y)[0] = res.x;
                out.at(x, y)[1] = res.y;
            }
    };
    // VolumetricCloud.cpp:299-314: [0] -> [1] (pass 0, x), [1] -> [2] (pass 1, y)
    pass(shadow_maps[0], shadow_maps[1], true);
    pass(shadow_maps[1], shadow_maps[2], false);
}

// ---------------------------------------------------------------- K13
void CloudScene::ShadowFroxel() {
    const mat4 uInvMVP(c.uInvMVP), uLightVP(c.uLightVP);
    Image<1>& img = shadow_froxel;
#pragma omp parallel for schedule(dynamic)
    for (int gy = 0; gy < img.h; ++gy)
        for (int gx = 0; gx < img.w; ++gx) {
            vec2 uv = (vec2(float(gx), float(gy)) + 0.5f) / vec2(float(img.w), float(img.h));
            vec3 frag_pos = ProjectiveMul(uInvMVP, vec3(uv * 2.0f - 1.0f, 0.0f));
            float step_size = c.uShadowFroxelMaxDistance / float(img.d);
            vec3 dir = normalize(frag_pos - uCameraPos());
            float t = 0.5f * step_size;
            float transmittance_sum = 0.0f;
            for (uint z = 0; z < uint(img.d); z++, t += step_size) {
                vec3 pos = uCameraPos() + t * dir;
                vec3 light_ndc = ProjectiveMul(uLightVP, pos);
                transmittance_sum += SampleCloudShadowTransmittance(shadow_maps[2], light_ndc);
                float ray_scatter_visibility = transmittance_sum / float(z + 1);
                img.at(gx, gy, int(z))[0] = MipTexture<1>::quantize(ray_scatter_visibility, 16);  // r16 image store
            }
        }
}

// ---------------------------------------------------------------- K14
void CloudScene::CheckerboardGen(const float* depth) {
    Image<1>& out = checkerboard_depth;
    auto D = [&](int x, int y) { return depth[size_t(clamp(y, 0, height - 1)) * width + clamp(x, 0, width - 1)]; };
#pragma omp parallel for schedule(static)
    for (int y = 0; y < out.h; ++y)
        for (int x = 0; x < out.w; ++x) {
            // textureGather at ((index+0.5)/half_size) = the 2x2 block (2x, 2y)
            float v0 = D(2 * x, 2 * y + 1), v1 = D(2 * x + 1, 2 * y + 1), v2 = D(2 * x + 1, 2 * y), v3 = D(2 * x, 2 * y);
            bool bmax = ((x & 1) == (y & 1));
            float d = bmax ? max(max(v0, v1), max(v2, v3)) : min(min(v0, v1), min(v2, v3));
            out.at(x, y)[0] = d;
        }
}

// ---------------------------------------------------------------- K15
void CloudScene::IndexGen() {
    Image<2>& out = index_linear_depth;
    auto PosToIndex = [&](ivec2 pos) -> uint { return (c.uBaseShadingIndex + uint((pos.x + pos.y) & 1)) & 3u; };
    static const ivec2 kTileOffsets[8] = {{0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}, {-1, 0}, {-1, 1}};
#pragma omp parallel for schedule(static)
    for (int y = 0; y < out.h; ++y)
        for (int x = 0; x < out.w; ++x) {
            ivec2 pos(x, y);
            vec4 depth_tile = texture_gather(checkerboard_depth, ivec2(2 * x, 2 * y));
            uint nearest_index = 0, farthest_index = 0;
            for (uint i = 1; i < 4; ++i) {
                nearest_index = depth_tile[i] < depth_tile[nearest_index] ? i : nearest_index;
                farthest_index = depth_tile[i] > depth_tile[farthest_index] ? i : farthest_index;
            }
            float nearest_linear_depth = DepthToLinearDepth(depth_tile[nearest_index]);
            float farthest_linear_depth = DepthToLinearDepth(depth_tile[farthest_index]);
            uint close_to_nearest_count = 0, close_to_farthest_count = 0;
            for (int i = 0; i < 8; ++i) {
                ivec2 tile_lt_pos = pos + kTileOffsets[i];
                // out-of-range neighbours: the reference relies on robust texelFetch; the oracle clamps
                float depth = texel_fetch_clamp(checkerboard_depth, (tile_lt_pos << 1) + IndexToOffset(PosToIndex(tile_lt_pos))).x;
                float linear_depth = DepthToLinearDepth(depth);
                float max_delta_allowed = linear_depth * 0.25f;
                if (std::fabs(linear_depth - nearest_linear_depth) < max_delta_allowed) ++close_to_nearest_count;
                if (std::fabs(linear_depth - farthest_linear_depth) < max_delta_allowed) ++close_to_farthest_count;
            }
            uint index = close_to_nearest_count == 0 ? nearest_index : close_to_farthest_count == 0 ? farthest_index : PosToIndex(pos);
            out.at(x, y)[0] = float(index);
            out.at(x, y)[1] = DepthToLinearDepth(depth_tile[index]);
        }
}

// ---------------------------------------------------------------- K16
namespace {
struct Intersect { float t1, t2; };
struct RayMarchContext {
    float t; vec3 pos; float height01; float step_size; float transmittance; float transmittance_sum;
    float weighted_t_sum; vec2 sun_env; float cos_sun_view;
};
// VolumetricCloudCommon.glsl:58-63
inline float HenyeyGreenstein(float cos_theta, float g) {
    float a = 1.0f - g * g;
    float b = 1.0f + g * g - 2.0f * g * cos_theta;
    b *= std::sqrt(b);
    return (0.25f * INV_PI) * a / b;
}
}  // namespace

void CloudScene::Render(int band_rows, int band_index, int band_count) {
    const mat4 uInvMVP(c.uInvMVP);
    const float kMinTransmittance = 0.01f;  // VolumetricCloudCommon.glsl:30
    const int W = render_texture.w, H = render_texture.h;

    // VolumetricCloudRender.comp:39-71
    auto RayShellIntersect = [&](float r, float mu, Intersect res[2]) {
        res[0].t1 = res[0].t2 = res[1].t1 = res[1].t2 = 0.0f;
        float bottom_radius = c.uEarthRadius + c.uBottomAltitude;
        float top_radius = c.uEarthRadius + c.uTopAltitude;
        float discriminant_bottom = r * r * (mu * mu - 1.0f) + bottom_radius * bottom_radius;
        float discriminant_top = r * r * (mu * mu - 1.0f) + top_radius * top_radius;
        float sqrt_discriminant_bottom = std::sqrt(discriminant_bottom);
        float sqrt_discriminant_top = std::sqrt(discriminant_top);
        if (c.uCameraPos[2] < c.uBottomAltitude) {
            res[0].t1 = -r * mu + sqrt_discriminant_bottom;
            res[0].t2 = -r * mu + sqrt_discriminant_top;
        } else if (c.uCameraPos[2] < c.uTopAltitude) {
            if (discriminant_bottom >= 0.0f && mu < 0.0f) {
                res[0].t2 = -r * mu - sqrt_discriminant_bottom;
                res[1].t1 = -r * mu + sqrt_discriminant_bottom;
                res[1].t2 = -r * mu + sqrt_discriminant_top;
            } else {
                res[0].t2 = -r * mu + sqrt_discriminant_top;
            }
        } else {
            if (discriminant_bottom >= 0.0f && mu < 0.0f) {
                res[0].t1 = -r * mu - sqrt_discriminant_top;
                res[0].t2 = -r * mu - sqrt_discriminant_bottom;
                res[1].t1 = -r * mu + sqrt_discriminant_bottom;
                res[1].t2 = -r * mu + sqrt_discriminant_top;
            } else if (discriminant_top >= 0.0f && mu < 0.0f) {
                res[0].t1 = -r * mu - sqrt_discriminant_top;
                res[0].t2 = -r * mu + sqrt_discriminant_top;
            }
        }
    };
    // VolumetricCloudRender.comp:98-114
    auto SampleShadow = [&](vec3 pos) {
        float optical_depth = 0.0f;
        float inv_shadow_steps = 1.0f / b.uShadowSteps;
        vec3 sample_vector = b.uShadowDistance * uSunDirection();
        float previous_t = 0.0f;
        for (float t = inv_shadow_steps; t <= 1.0f; t += inv_shadow_steps) {
            float current_t = t * t;
            float delta_t = current_t - previous_t;
            vec3 sample_pos = pos + sample_vector * (previous_t + 0.5f * delta_t);
            float sample_height01 = CalHeight01(sample_pos);
            optical_depth += SampleSigmaT(sample_pos, sample_height01, SKY_CNT_RENDER_SIGMA_EVALS) * b.uShadowDistance * delta_t;
            previous_t = current_t;
        }
        return std::exp(-optical_depth);
    };
    // VolumetricCloudRender.comp:116-137
    auto RayMarchStep = [&](RayMarchContext& ctx) {
        float sigma_t = SampleSigmaT(ctx.pos, ctx.height01, SKY_CNT_RENDER_SIGMA_EVALS);
        if (sigma_t < 1e-5f) return;
        float tr = std::exp(-ctx.step_size * sigma_t);
        vec2 sun_env(0.0f);
        float transmittance_to_sun = SampleShadow(ctx.pos);
        float phase = mix(HenyeyGreenstein(ctx.cos_sun_view, -0.15f) * 2.16f, HenyeyGreenstein(ctx.cos_sun_view, 0.85f),
                          std::exp(-b.uSunMultiscatteringSigmaScale * sigma_t));
        sun_env.x = transmittance_to_sun * phase;
        sun_env.y = mix(b.uEnvBottomVisibility, 1.0f, ctx.height01);
        sun_env.y = sun_env.y - sun_env.y * std::exp(-b.uEnvMultiscatteringSigmaScale * sigma_t);
        sun_env = sun_env - sun_env * tr;
        ctx.sun_env += ctx.transmittance * sun_env;
        ctx.transmittance_sum += ctx.transmittance;
        ctx.weighted_t_sum += ctx.t * ctx.transmittance;
        ctx.transmittance *= tr;
    };

#pragma omp parallel for schedule(dynamic)
    for (int py = 0; py < H; ++py) {
        if (band_rows > 0 && (py / band_rows) % band_count != band_index) continue;
        for (int px = 0; px < W; ++px) {
            ivec2 pos(px, py);
            uint index = uint(index_linear_depth.at(px, py)[0]);
            ivec2 pos_in_checkerboard = pos * 2 + IndexToOffset(index);
            vec2 image_size(float(checkerboard_depth.w), float(checkerboard_depth.h));
            vec2 uv = (tovec2(pos_in_checkerboard) + 0.5f) / image_size;
            float depth = texel_fetch_clamp(checkerboard_depth, pos_in_checkerboard).x;
            vec3 frag_pos = ProjectiveMul(uInvMVP, vec3(uv, depth) * 2.0f - 1.0f);
            vec3 view_dir = normalize(frag_pos - uCameraPos());

            float r = c.uCameraPos[2] + c.uEarthRadius;
            float mu = view_dir.z;
            Intersect intersect[2];
            RayShellIntersect(r, mu, intersect);
            float frag_dist = distance(frag_pos, uCameraPos());
            for (int i = 0; i < 2; ++i)
                intersect[i].t2 = clamp(min(frag_dist, b.uMaxVisibleDistance), intersect[i].t1, intersect[i].t2);

            RayMarchContext ctx;
            ctx.cos_sun_view = dot(uSunDirection(), view_dir);
            float dist = intersect[0].t2 - intersect[0].t1;
            dist = min(dist, b.uMaxRaymarchDistance);
            uint num_steps = uint(max(b.uMaxRaymarchSteps * (dist / b.uMaxRaymarchDistance), 1.0f));
            ctx.step_size = dist / float(num_steps);
            ctx.transmittance = 1.0f;
            ctx.transmittance_sum = 0.0f;
            ctx.weighted_t_sum = 0.0f;
            ctx.sun_env = vec2(0.0f);
            float noise = blue_noise.at(px & 0x3f, py & 0x3f)[0];
            ctx.t = intersect[0].t1 + ctx.step_size * fract(noise + c.uFrameID * 0.61803398875f);
            for (uint cnt = num_steps; cnt != 0; cnt--, ctx.t += ctx.step_size) {
                ctx.pos = uCameraPos() + view_dir * ctx.t;  // UpdateContext, :90-93
                ctx.height01 = CalHeight01(ctx.pos);
                RayMarchStep(ctx);
                if (ctx.transmittance < kMinTransmittance) break;
            }
            float dist1 = intersect[1].t2 - intersect[1].t1;
            if (dist1 > 0) {
                dist1 = min(dist1, b.uMaxRaymarchDistance);
                uint num_steps1 = uint(max(b.uMaxRaymarchSteps * (dist1 / b.uMaxRaymarchDistance), 1.0f));
                ctx.step_size = dist1 / float(num_steps1);
                ctx.t = intersect[1].t1 + ctx.step_size * fract(noise + c.uFrameID * 0.61803398875f);
                for (uint cnt = num_steps1; cnt != 0; cnt--, ctx.t += ctx.step_size) {
                    ctx.pos = uCameraPos() + view_dir * ctx.t;
                    ctx.height01 = CalHeight01(ctx.pos);
                    RayMarchStep(ctx);
                    if (ctx.transmittance < kMinTransmittance) break;
                }
            }
            float average_t = ctx.weighted_t_sum == 0 ? frag_dist : ctx.weighted_t_sum / ctx.transmittance_sum;
            cloud_distance.at(px, py)[0] = average_t;
            vec3 average_pos = uCameraPos() + view_dir * average_t;
            // SunCosTheta, :85-88
            vec3 up = normalize(vec3(average_pos.x, average_pos.y, average_pos.z + c.uEarthRadius));
            float sun_cos_theta = clamp(dot(up, uSunDirection()), 0.0f, 1.0f);
            vec3 luminance = ctx.sun_env.x * b.uSunIlluminanceScale * GetSunVisibility(average_pos) * atm.solar_illuminance() +
                             ctx.sun_env.y * std::pow(sun_cos_theta, b.uEnvSunHeightCurveExp) * vec3(b.uEnvColorScale);
            vec3 atmosphere_transmittance;
            vec3 atmosphere_luminance = GetAerialPerspective(uv, average_t, r, mu, atmosphere_transmittance);
            atmosphere_luminance *= SampleRayScatterVisibility(shadow_froxel, uv, average_t, c.uInvShadowFroxelMaxDistance);
            luminance = luminance * atmosphere_transmittance + atmosphere_luminance * (1 - ctx.transmittance);

            luminance /= max(1e-5f, (1 - ctx.transmittance));
            float fade = smoothstep(b.uMaxVisibleDistance * 0.75f, b.uMaxVisibleDistance, intersect[0].t1);
            ctx.transmittance = mix(ctx.transmittance, 1.0f, fade);
            luminance *= 1 - ctx.transmittance;

            // rgba16f image store
            render_texture.store(px, py, vec4(to_half_and_back(luminance.x), to_half_and_back(luminance.y),
                                              to_half_and_back(luminance.z), to_half_and_back(ctx.transmittance)));
        }
    }
}

// ---------------------------------------------------------------- K17
void CloudScene::Reconstruct() {
    const mat4 uInvMVP(c.uInvMVP), uReprojectMat(c.uReprojectMat);
    Image<4>& out = reconstruct[0];
    const Image<4>& pre = reconstruct[1];
    static const ivec2 kOffsets[9] = {{0, 0}, {0, 1}, {1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}, {-1, 0}, {-1, 1}};
    auto Reinhard = [](vec4& v) { v.x = v.x / (1.0f + v.x); v.y = v.y / (1.0f + v.y); v.z = v.z / (1.0f + v.z); };
    auto InverseReinhard = [](vec4& v) { v.x = v.x / (1.0f - v.x); v.y = v.y / (1.0f - v.y); v.z = v.z / (1.0f - v.z); };
#pragma omp parallel for schedule(static)
    for (int y = 0; y < out.h; ++y)
        for (int x = 0; x < out.w; ++x) {
            ivec2 pos(x, y);
            ivec2 q = pos >> 1;
            vec2 uv = (tovec2(pos) + 0.5f) / vec2(float(out.w), float(out.h));
            float depth = checkerboard_depth.at(x, y)[0];
            float linear_depth = DepthToLinearDepth(depth);
            // gathers of the .g component at integer bases q-1 and q (:37-38)
            vec4 block_minus1_minus1 = texture_gather(index_linear_depth, q - 1, 1);
            vec4 block_0_0 = texture_gather(index_linear_depth, q, 1);
            float rendered_linear_depths[9] = {
                block_0_0.w, block_0_0.x, block_0_0.y, block_0_0.z,
                texel_fetch_clamp(index_linear_depth, q + kOffsets[4]).y,
                block_minus1_minus1.z, block_minus1_minus1.w, block_minus1_minus1.x,
                texel_fetch_clamp(index_linear_depth, q + kOffsets[8]).y,
            };
            float delta_linear_depths[9];
            float min_delta_linear_depth = 1e10f;
            int nearest_i = 0;
            for (int i = 0; i < 9; ++i) {
                delta_linear_depths[i] = std::fabs(rendered_linear_depths[i] - linear_depth);
                if (delta_linear_depths[i] < min_delta_linear_depth) {
                    min_delta_linear_depth = delta_linear_depths[i];
                    nearest_i = i;
                }
            }
            vec4 rendered_nearest = texel_fetch_clamp(render_texture, q + kOffsets[nearest_i]);
            Reinhard(rendered_nearest);
            vec4 aabb_min = rendered_nearest, aabb_max = rendered_nearest;
            for (int i = 0; i < 9; ++i) {
                vec4 rendered = texel_fetch_clamp(render_texture, q + kOffsets[i]);
                Reinhard(rendered);
                if (delta_linear_depths[i] < rendered_linear_depths[i] * 0.3f ||
                    std::fabs(rendered.w - rendered_nearest.w) / max(1e-6f, 1 - max(rendered.w, rendered_nearest.w)) < 0.2f) {
                    aabb_min = min(aabb_min, rendered);
                    aabb_max = max(aabb_max, rendered);
                }
            }
            vec4 rendered = texel_fetch_clamp(render_texture, q);
            Reinhard(rendered);
            vec3 frag_pos = ProjectiveMul(uInvMVP, vec3(uv, depth) * 2.0f - 1.0f);
            vec3 view_dir = normalize(frag_pos - uCameraPos());
            float rendered_distance = texel_fetch_clamp(cloud_distance, q).x;
            vec3 cloud_pos = uCameraPos() + view_dir * rendered_distance;
            vec2 pre_ndc = ProjectiveMul(uReprojectMat, cloud_pos).xy();
            vec2 pre_uv = pre_ndc * 0.5f + 0.5f;
            vec4 pre_frame = texture_linear(pre, pre_uv, Sampler());
            Reinhard(pre_frame);
            pre_frame = clamp(pre_frame, aabb_min, aabb_max);

            bool is_pre_out_of_screen = max(std::fabs(pre_ndc.x), std::fabs(pre_ndc.y)) > 1.0f;
            int rendered_index = int(texel_fetch_clamp(index_linear_depth, q).x);
            bool is_rendered = (pos & 1) == IndexToOffset(uint(rendered_index));
            float rendered_weight = is_pre_out_of_screen ? 1.0f : is_rendered ? 0.2f : 0.0f;
            vec4 reconstructed = mix(pre_frame, rendered, rendered_weight);
            InverseReinhard(reconstructed);
            out.store(x, y, vec4(to_half_and_back(reconstructed.x), to_half_and_back(reconstructed.y),
                                 to_half_and_back(reconstructed.z), to_half_and_back(reconstructed.w)));
        }
}

// ---------------------------------------------------------------- K18
void CloudScene::Upscale(const float* depth_img, uint16_t* hdr) {
    const float kMinTransmittance = 0.01f;
    const Image<4>& rec = reconstruct[0];
    static const ivec2 kOffsets[4] = {{-1, 1}, {1, 1}, {1, -1}, {-1, -1}};
#pragma omp parallel for schedule(static)
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            if (!owns_row(y)) continue;  // sky_set_output_bands
            ivec2 pos(x, y);
            float depth = depth_img[size_t(y) * width + x];
            float linear_depth = DepthToLinearDepth(depth);
            vec4 neighbor_depths = texture_gather(checkerboard_depth, (pos - 1) >> 1);
            vec4 reconstructed_neighbors[4];
            float min_delta_linear_depth = 1e10f;
            int nearest_i = 0;
            bool is_edge = false;
            for (int i = 0; i < 4; ++i) {
                ivec2 half_pos = (pos + kOffsets[i]) >> 1;
                reconstructed_neighbors[i] = texel_fetch_clamp(rec, half_pos);
                float neighbor_linear_depth = DepthToLinearDepth(neighbor_depths[i]);
                float delta_linear_depth = std::fabs(linear_depth - neighbor_linear_depth);
                if (delta_linear_depth < min_delta_linear_depth) {
                    nearest_i = i;
                    min_delta_linear_depth = delta_linear_depth;
                }
                if (delta_linear_depth > linear_depth * 0.1f) is_edge = true;
            }
            vec4 upscaled;
            float max_a = max(max(reconstructed_neighbors[0].w, reconstructed_neighbors[1].w),
                                   max(reconstructed_neighbors[2].w, reconstructed_neighbors[3].w));
            float min_a = min(min(reconstructed_neighbors[0].w, reconstructed_neighbors[1].w),
                                   min(reconstructed_neighbors[2].w, reconstructed_neighbors[3].w));
            if (is_edge && (max_a - min_a) / max(1e-6f, 1 - min_a) > 0.2f) {
                upscaled = reconstructed_neighbors[nearest_i];
            } else {
                upscaled = ((reconstructed_neighbors[0] + reconstructed_neighbors[1]) +
                            (reconstructed_neighbors[2] + reconstructed_neighbors[3])) * 0.25f;
            }
            vec3 luminance = upscaled.rgb();
            float transmittance = upscaled.w;
            uint16_t* o = hdr + (size_t(y) * width + x) * 4;
            vec3 color(half_bits_to_float(o[0]), half_bits_to_float(o[1]), half_bits_to_float(o[2]));
            color = color * (transmittance <= kMinTransmittance ? 0.0f : transmittance) + luminance;
            o[0] = float_to_half_bits(color.x);
            o[1] = float_to_half_bits(color.y);
            o[2] = float_to_half_bits(color.z);
        }
}

}  // namespace orc
