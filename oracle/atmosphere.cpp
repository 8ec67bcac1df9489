// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h).  K1-K6 entry points.
#include "atmosphere.h"
#include "ibl.h"

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

namespace {
// texture(sampler2D, uv) with LinearNoMipmapClampToEdge (AtmosphereRenderer.cpp:200-217) over a normalised-integer image
template <class T, class Decode>
vec4 gbuffer_texture(const T* img, int w, int h, vec2 uv, Decode decode) {
    float x = uv.x * float(w) - 0.5f, y = uv.y * float(h) - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    float a = x - fx, b = y - fy;
    int i0 = int(fx), j0 = int(fy);
    auto T4 = [&](int i, int j) {
        const T* p = img + (size_t(clamp(j, 0, h - 1)) * w + clamp(i, 0, w - 1)) * 4;
        return vec4(decode(p[0]), decode(p[1]), decode(p[2]), decode(p[3]));
    };
    return (1.0f - a) * (1.0f - b) * T4(i0, j0) + a * (1.0f - b) * T4(i0 + 1, j0) + (1.0f - a) * b * T4(i0, j0 + 1) + a * b * T4(i0 + 1, j0 + 1);
}
inline vec3 mix3(vec3 a, vec3 b, vec3 t) { return vec3(mix(a.x, b.x, t.x), mix(a.y, b.y, t.y), mix(a.z, b.z, t.z)); }
// shaders/Base/BRDF.glsl:26-49
inline float Pow5(float x) { float x2 = x * x; return x2 * x2 * x; }
inline vec3 F_Schlick(float HdotV, vec3 F0) { return F0 + (vec3(1.0f) - F0) * Pow5(1.0f - HdotV); }
inline float D_GGX(float a, float NdotH) {
    float a2 = a * a;
    float d = (NdotH * a2 - NdotH) * NdotH + 1.0f;
    return a2 / (PI * d * d);
}
inline float Vis_SmithJointApprox(float a, float NdotV, float NdotL) {
    float Vis_SmithV = NdotL * (NdotV * (1.0f - a) + a);
    float Vis_SmithL = NdotV * (NdotL * (1.0f - a) + a);
    return 0.5f / max(Vis_SmithV + Vis_SmithL, 1e-9f);
}
struct MaterialData { vec3 F0, diffuse, normal; float roughness; };
// shaders/Base/BRDF.glsl:69-78
inline vec3 BRDF(float NdotL, float NdotV, float NdotH, float HdotV, const MaterialData& data) {
    float a = data.roughness * data.roughness;
    float D = min(D_GGX(a, NdotH), 1e9f);
    float Vis = Vis_SmithJointApprox(a, NdotV, NdotL);
    vec3 F = F_Schlick(HdotV, data.F0);
    vec3 specular = vec3(D * Vis);
    vec3 diffuse = INV_PI * data.diffuse;
    return mix3(diffuse, specular, F);
}
// shaders/Base/BRDF.glsl:81-106
inline vec3 GetSHIrradiance(vec3 N, const vec4* Llm) {
    const float c1 = 0.429043f, c2 = 0.511664f, c3 = 0.743125f, c4 = 0.886227f, c5 = 0.247708f;
    vec3 L00 = Llm[0].rgb(), L1_1 = Llm[1].rgb(), L10 = Llm[2].rgb(), L11 = Llm[3].rgb(), L2_2 = Llm[4].rgb(), L2_1 = Llm[5].rgb(),
         L20 = Llm[6].rgb(), L21 = Llm[7].rgb(), L22 = Llm[8].rgb();
    float x = N.x, y = N.y, z = N.z;
    return c1 * (x * x - y * y) * L22 + c3 * (z * z) * L20 + c4 * L00 - c5 * L20
        + 2.0f * c1 * (x * y * L2_2 + x * z * L21 + y * z * L2_1)
        + 2.0f * c2 * (x * L11 + y * L1_1 + z * L10);
}
}  // namespace

// Shadow.glsl:13-99 (PCSS_ENABLE 1).  shadowMap: Samplers::GetShadowMapSampler (LINEAR, border 1, compare LEQUAL: the bilinear
// blend of four comparison results); shadow_map_depth_sampler: the same texture through NearestClampToEdge
// (AtmosphereRenderer.cpp:196,213).  sin / cos: sky_detmath.h; pow: libm.
float AtmosphereRenderer::PCSS(const Image<1>& shadow_map, vec3 position) const {
    constexpr int NUM_SAMPLES = 25, NUM_RINGS = 3;
    const float PI2 = PI * 2.0f;
    vec4 xyzw = mat4(u.light_view_projection) * vec4(position, 1.0f);
    vec3 coords = vec3(xyzw.x, xyzw.y, xyzw.z) / xyzw.w;
    coords = coords * 0.5f + 0.5f;
    if (coords.z >= 1.0f) return 1.0f;
    vec2 poissonDisk[NUM_SAMPLES];
    {   // poissonDiskSamples(coords.xy)
        float ANGLE_STEP = PI2 * float(NUM_RINGS) / float(NUM_SAMPLES);
        float INV_NUM_SAMPLES = 1.0f / float(NUM_SAMPLES);
        // rand_2to1
        const float a = 12.9898f, b = 78.233f, c = 43758.5453f;
        float dt = dot(coords.xy(), vec2(a, b)), sn = dt - PI * std::floor(dt / PI);
        float rnd = fract(sky_det_sinf(sn) * c);
        float angle = rnd * PI2;
        float radius = INV_NUM_SAMPLES;
        float radiusStep = radius;
        for (int i = 0; i < NUM_SAMPLES; i++) {
            poissonDisk[i] = vec2(sky_det_cosf(angle), sky_det_sinf(angle)) * std::pow(radius, 0.75f);
            radius += radiusStep;
            angle += ANGLE_STEP;
        }
    }
    const int W = shadow_map.w, H = shadow_map.h;
    float kernelSizeApproximate = u.blocker_kernel_size_k * coords.z;
    float avgblockerDepth;
    {   // FindBlocker
        float sum = 0.0f, cnt = 0.0f;
        for (int i = 0; i < NUM_SAMPLES; ++i) {
            vec2 samplePos = poissonDisk[i] * kernelSizeApproximate + coords.xy();
            int tx = clamp(int(std::floor(samplePos.x * float(W))), 0, W - 1), ty = clamp(int(std::floor(samplePos.y * float(H))), 0, H - 1);
            float shadow_depth = shadow_map.load(tx, ty).x;
            if (coords.z - shadow_depth > 0.0f) {
                cnt += 1.0f;
                sum += shadow_depth;
            }
        }
        avgblockerDepth = sum / max(cnt, 1e-5f);
    }
    float distanceToFragment = coords.z - avgblockerDepth;
    float penumbraSize = u.pcss_size_k * distanceToFragment;
    float sum = 0.0f;
    for (int i = 0; i < NUM_SAMPLES; ++i) {  // Filtering
        vec2 samplePos = poissonDisk[i] * penumbraSize + coords.xy();
        sum += Atmosphere::ShadowCompare(shadow_map, samplePos.x, samplePos.y, coords.z);
    }
    return sum / float(NUM_SAMPLES);
}

// AtmosphereRenderer.glsl:333-343; the mesh shadow map is an input that is all 1.0 (lit) until the caller writes
// SKY_RES_MESH_SHADOW_MAP
float AtmosphereRenderer::SampleVisibilityFromShadowMap(vec3 position) const {
    float visibility = 1.0f;
    if (mesh_shadow_map)
        visibility = cfg.pcss ? PCSS(*mesh_shadow_map, position)
                              : Atmosphere::GetVisibilityFromShadowMap(*mesh_shadow_map, mat4(u.light_view_projection), position);
    if (object->cloud_shadow_map) {
        vec3 light_ndc = ProjectiveMul(mat4(u.uCloudShadowMapMat), position);
        // SampleCloudShadowTransmittance, VolumetricCloudShadowInterface.glsl:4-8; sampler border (1e10, 1), VolumetricCloud.cpp:106-112
        Sampler s; s.wrap = CLAMP_TO_BORDER; s.border = vec4(1e10f, 1.0f, 0.0f, 0.0f);
        const float kInvTransitionDepth = 1.0f / 0.5f;
        vec4 dt = texture_linear(*object->cloud_shadow_map, light_ndc.xy() * 0.5f + 0.5f, s);
        visibility = min(visibility, mix(dt.y, 1.0f, clamp((dt.x - light_ndc.z) * kInvTransitionDepth, 0.0f, 1.0f)));
    }
    return visibility;
}

// AtmosphereRenderer.glsl:284-324 (LoadMeterialData / BRDF / GetAmbient: shaders/Base/BRDF.glsl:10-23,69-78,108-130)
vec3 AtmosphereRenderer::ComputeObjectLuminance(vec3 position, vec3 view_direction, float shadow_visibility, vec2 vTexCoord,
                                                int width, int height) const {
    float r = length(position - earth_center());
    vec3 object_up_direction = normalize(position - earth_center());
    float mu_s = dot(sun_direction(), object_up_direction);
    vec3 sun_visibility;
    if (r > atm.u.top_radius) {
        float near_distance;
        if (atm.FromSpaceIntersectTopAtmosphereBoundary(r, mu_s, near_distance)) {
            position += near_distance * sun_direction();
            r = length(position - earth_center());
            object_up_direction = normalize(position - earth_center());
            mu_s = dot(sun_direction(), object_up_direction);
            sun_visibility = atm.GetSunVisibility(transmittance_texture, r, mu_s);
        } else {
            sun_visibility = vec3(1.0f);
        }
    } else {
        sun_visibility = atm.GetSunVisibility(transmittance_texture, r, mu_s);
    }
    vec3 solar_illuminance_at_object = atm.solar_illuminance() * sun_visibility;

    vec3 albedo = gbuffer_texture(object->albedo, width, height, vTexCoord, [](uint8_t c) { return float(c) / 255.0f; }).rgb();
    vec3 normal = gbuffer_texture(object->normal, width, height, vTexCoord, [](int16_t c) { return std::fmax(float(c) / 32767.0f, -1.0f); }).rgb();
    vec3 orm = gbuffer_texture(object->orm, width, height, vTexCoord, [](uint16_t c) { return float(c) / 65535.0f; }).rgb();
    MaterialData material_data;
    float metallic = orm.z;
    material_data.F0 = vec3(0.04f) * (1.0f - metallic) + albedo * metallic;
    material_data.diffuse = albedo - albedo * metallic;
    material_data.normal = normal;
    material_data.roughness = orm.y;

    vec3 N = material_data.normal;
    vec3 L = sun_direction();
    vec3 V = -view_direction;
    vec3 H = normalize(L + V);
    float NdotL = clamp(dot(N, L), 0.0f, 1.0f);
    float NdotV = clamp(dot(N, V), 0.0f, 1.0f);
    float NdotH = clamp(dot(N, H), 0.0f, 1.0f);
    float HdotV = clamp(dot(H, V), 0.0f, 1.0f);
    vec3 brdf = BRDF(NdotL, NdotV, NdotH, HdotV, material_data);
    vec3 direct_lumiance = brdf * solar_illuminance_at_object * (NdotL * shadow_visibility);
    // GetAmbient(env_brdf_lut, ROUGHNESS_COUNT - 1, prefiltered_radiance_texture, V, material_data, Llm)
    vec3 ambient_lumiance;
    {
        const float roughness_lod_max = float(SKY_IBL_ROUGHNESS_COUNT - 1);
        vec3 F = F_Schlick(NdotV * 0.8f + 0.2f, material_data.F0);
        vec3 approx_irradiance_over_pi = GetSHIrradiance(N, object->Llm) * INV_PI;
        vec3 diffuse = (material_data.diffuse - material_data.diffuse * F) * approx_irradiance_over_pi;
        vec3 R = 2.0f * NdotV * N - V;
        vec3 prefiltered_radiance = TextureCubeLod(*object->prefiltered, R, material_data.roughness * roughness_lod_max).rgb();
        vec4 brdf_lut = texture_linear(*object->env_brdf_lut, vec2(NdotV, material_data.roughness), Sampler());
        vec3 specular = prefiltered_radiance * (material_data.F0 * brdf_lut.x + vec3(brdf_lut.y));
        ambient_lumiance = diffuse + specular;
    }
    float ambient_fade = clamp(10.0f - 0.1f * length(position - camera_position()), 0.0f, 1.0f);
    return direct_lumiance + ambient_lumiance * ambient_fade;
}

// K6 -- AtmosphereRenderer.glsl:345-432.  gl_FragCoord.xy = pixel + 0.5, vTexCoord = that / size.
// Object pixels (depth != 1) are shaded by ComputeObjectLuminance (:284-324) when a G-buffer is bound; without one they carry the
// in-scatter alone -- the program's result on a cleared G-buffer.  Alpha is 1 everywhere (:431).
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

            if (intersect_object) {
                if (object) {  // :404-410
                    float shadow_visibility = SampleVisibilityFromShadowMap(fragment_position);
                    if (ex.moon_shadow)
                        shadow_visibility *= atm.GetVisibilityFromMoonShadow(ex.moon_position - fragment_position, ex.moon_radius, sun_direction());
                    luminance += transmittance * ComputeObjectLuminance(fragment_position, view_direction, shadow_visibility, vTexCoord, width, height);
                }
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
            o[3] = float_to_half_bits(1.0f);   // FragColor = vec4(luminance, 1.0), :431
        }
}

}  // namespace orc
