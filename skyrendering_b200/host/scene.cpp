// Host-side uniform maths; see scene.h.  Each function cites the reference code it mirrors.
#include "scene.h"

#include <algorithm>

namespace skyhost {

// ---- Atmosphere.cpp:49-70 -------------------------------------------------------------------------
void AssignBufferData(const AtmosphereParameters& p, SkyAtmosphereBufferData& d) {
    std::memset(&d, 0, sizeof(d));
    auto put = [](float* dst, vec3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; };
    put(d.solar_illuminance, p.solar_illuminance);
    d.sun_angular_radius = radians(p.sun_angular_radius);
    put(d.rayleigh_scattering, p.rayleigh_scattering * p.rayleigh_scattering_scale);
    put(d.mie_scattering, p.mie_scattering * p.mie_scattering_scale);
    put(d.mie_absorption, p.mie_absorption * p.mie_absorption_scale);
    put(d.ozone_absorption, p.ozone_absorption * p.ozone_absorption_scale);
    d.inv_rayleigh_exponential_distribution = 1.0f / p.rayleigh_exponential_distribution;
    d.inv_mie_exponential_distribution = 1.0f / p.mie_exponential_distribution;
    d.ozone_center_altitude = p.ozone_center_altitude;
    d.inv_ozone_width = 1.0f / p.ozone_width;
    put(d.ground_albedo, p.ground_albedo);
    d.bottom_radius = p.bottom_radius;
    d.top_radius = p.bottom_radius + p.thickness;
    d.mie_phase_g = p.mie_phase_g;
    d.transmittance_steps = p.transmittance_steps;
    d.multiscattering_steps = p.multiscattering_steps;
    d.multiscattering_mask = p.multiscattering_mask;
}

// ---- Earth.cpp:67-77 --------------------------------------------------------------------------------
mat4 Earth::moon_model() const {
    mat4 model;
    model = translate(model, center());
    model = rotate(model, radians(moon_status.direction_phi), vec3(0, 1, 0));
    model = rotate(model, radians(moon_status.direction_theta), vec3(0, 0, -1));
    model = translate(model, vec3(0, moon_status.distance + moon_status.radius + parameters.bottom_radius, 0));
    // the trailing scale/rotate (Earth.cpp:75-76) do not move column 3, the only part the hot path reads
    return model;
}

// ---- Camera.cpp:21-37 -------------------------------------------------------------------------------
void Camera::Rotate(float dPitch, float dYaw) {
    pitch_ += dPitch;
    yaw_ += dYaw;
    if (pitch_ > 89.0f) pitch_ = 89.0f;
    else if (pitch_ < -89.0f) pitch_ = -89.0f;
    UpdateVectors();
}
void Camera::UpdateVectors() {
    const vec3 kWorldUp{0, 1, 0};
    front_ = FromThetaPhiToDirection(radians(90.f - pitch_), radians(yaw_));
    right_ = normalize(cross(front_, kWorldUp));
    up_ = normalize(cross(right_, front_));
}

// ---- VolumetricCloudDefaultMaterial.cpp:24-28 ---------------------------------------------------------
static float CalKLod(int x, int y, int z, float repeat_size, vec2 viewport, const Camera& camera) {
    float max_width = float(std::max(x, std::max(y, z)));
    float tan_half_fovy = std::tan(radians(camera.fovy) * 0.5f);
    return max_width * tan_half_fovy / (repeat_size * std::min(viewport.x, viewport.y));
}

static double fract(double x) { return x - std::floor(x); }

// ---- VolumetricCloudDefaultMaterial.cpp:78-102 ----------------------------------------------------------
void VolumetricCloudDefaultMaterialCommon::Update(vec2 viewport, const Camera& camera, dvec2 offset_from_first,
                                                  float delta_time, vec2& additional_delta,
                                                  SkyMaterialCommonBufferData& buffer) {
    for (const std::string* f : {&minfilter2d_, &minfilter3d_, &minfilter_displacement_})
        if (*f != "NEAREST_MIPMAP_NEAREST")
            throw std::runtime_error("material min filter '" + *f + "' is not supported (only NEAREST_MIPMAP_NEAREST, the value in all shipped scenes)");
    std::memset(&buffer, 0, sizeof(buffer));
    buffer.uLodBias = lod_bias_;
    buffer.uDensity = density_;
    auto gen_sample_info = [&](int x, int y, int z, float repeat_size, SkySampleInfo& info, dvec2 offset) {
        info.frequency = 1.0f / repeat_size;
        info.bias[0] = float(fract(offset.x / double(repeat_size)));
        info.bias[1] = float(fract(offset.y / double(repeat_size)));
        info.k_lod = CalKLod(x, y, z, repeat_size, viewport, camera);
    };
    const vec2 kLocalWindDirection{1.0f, 0.0f};
    additional_delta.x = kLocalWindDirection.x * wind_speed_ * delta_time;  // ImGui::GetIO().DeltaTime in the reference
    additional_delta.y = kLocalWindDirection.y * wind_speed_ * delta_time;
    dvec2 cur{offset_from_first.x + double(additional_delta.x), offset_from_first.y + double(additional_delta.y)};
    detail_offset_from_first_.x += double(additional_delta.x * detail_wind_magnify_);
    detail_offset_from_first_.y += double(additional_delta.y * detail_wind_magnify_);
    gen_sample_info(512, 512, 1, cloud_map_repeat_size, buffer.uCloudMapSampleInfo, cur);
    gen_sample_info(128, 128, 128, detail_repeat_size, buffer.uDetailSampleInfo,
                    dvec2{cur.x + detail_offset_from_first_.x, cur.y + detail_offset_from_first_.y});
    gen_sample_info(128, 128, 1, displacement_repeat_size, buffer.uDisplacementSampleInfo, cur);
}

static SkyNoiseCreateInfo to_pod(const NoiseCreateInfo& n) {
    return SkyNoiseCreateInfo{uint32_t(n.seed), uint32_t(n.base_frequency), n.remap_min, n.remap_max};
}
bool VolumetricCloudDefaultMaterialCommon::NoiseInfo(int kind, SkyNoiseCreateInfo* out) const {
    out[0] = out[1] = SkyNoiseCreateInfo{0, 1, 0, 1};
    switch (kind) {
        case SKY_NOISE_CLOUD_MAP: out[0] = to_pod(cloud_map_uDensity); out[1] = to_pod(cloud_map_uHeight); return true;
        case SKY_NOISE_DETAIL: out[0] = to_pod(detail_uPerlin); out[1] = to_pod(detail_uWorley); return true;
        case SKY_NOISE_DISPLACEMENT: out[0] = to_pod(displacement_uPerlin); return true;
    }
    return false;
}

namespace {

// VolumetricCloudDefaultMaterial.h:141-168, .cpp:183-226
struct VolumetricCloudDefaultMaterial0 : IVolumetricCloudMaterial {
    vec2 detail_param_{0.4f, 0.0f};
    float displacement_scale_ = 1.0f;
    VolumetricCloudDefaultMaterialCommon materail_common_;
    const char* TypeName() const override { return "class VolumetricCloudDefaultMaterial0"; }
    int Type() const override { return SKY_MATERIAL_DEFAULT0; }
    void Load(Archive& ar) override {
        ar("materail_common_", materail_common_); ar("detail_param_", detail_param_); ar("displacement_scale_", displacement_scale_);
    }
    void Update(vec2 viewport, const Camera& camera, dvec2 off, float dt, vec2& add, SkyMaterialBlock& out) override {
        std::memset(&out, 0, sizeof(out));
        out.type = Type();
        materail_common_.Update(viewport, camera, off, dt, add, out.common);
        out.u.m0.uDetailParam[0] = detail_param_.x;
        out.u.m0.uDetailParam[1] = detail_param_.y;
        out.u.m0.uDisplacementScale = displacement_scale_;
    }
    float GetSigmaTMax() const override { return materail_common_.density_; }
    bool NoiseInfo(int kind, SkyNoiseCreateInfo* out) const override { return materail_common_.NoiseInfo(kind, out); }
};

// VolumetricCloudDefaultMaterial.h:170-207, .cpp:228-277
struct VolumetricCloudDefaultMaterial1 : IVolumetricCloudMaterial {
    float detail_base_ = 0.67f, detail_scale_ = 1.86f, base_density_threshold_ = 0.4f, base_height_hardness_ = 6.0f;
    float base_edge_hardness_ = 6.0f, height_cut_ = 0.9f, edge_cut_ = 0.8f;
    VolumetricCloudDefaultMaterialCommon materail_common_;
    const char* TypeName() const override { return "class VolumetricCloudDefaultMaterial1"; }
    int Type() const override { return SKY_MATERIAL_DEFAULT1; }
    void Load(Archive& ar) override {
        ar("materail_common_", materail_common_); ar("base_density_threshold_", base_density_threshold_);
        ar("base_height_hardness_", base_height_hardness_); ar("base_edge_hardness_", base_edge_hardness_);
        ar("detail_base_", detail_base_); ar("detail_scale_", detail_scale_); ar("height_cut_", height_cut_);
        ar("edge_cut_", edge_cut_);
    }
    void Update(vec2 viewport, const Camera& camera, dvec2 off, float dt, vec2& add, SkyMaterialBlock& out) override {
        std::memset(&out, 0, sizeof(out));
        out.type = Type();
        materail_common_.Update(viewport, camera, off, dt, add, out.common);
        auto& b = out.u.m1;
        b.uBaseDensityThreshold = base_density_threshold_;
        b.uBaseHeightHardness = base_height_hardness_;
        b.uBaseEdgeHardness = base_edge_hardness_;
        b.uDetailBase = detail_base_;
        b.uDetailScale = detail_scale_;
        b.uHeightCut = 1.0f - height_cut_;
        b.uEdgeCur = edge_cut_;
    }
    float GetSigmaTMax() const override { return materail_common_.density_; }
    bool NoiseInfo(int kind, SkyNoiseCreateInfo* out) const override { return materail_common_.NoiseInfo(kind, out); }
};

// VolumetricCloudMinimalMaterial.{h,cpp}
struct VolumetricCloudMinimalMaterial : IVolumetricCloudMaterial {
    float density_ = 0.5f;
    const char* TypeName() const override { return "class VolumetricCloudMinimalMaterial"; }
    int Type() const override { return SKY_MATERIAL_MINIMAL; }
    void Load(Archive& ar) override { ar("density_", density_); }
    void Update(vec2, const Camera&, dvec2, float, vec2& add, SkyMaterialBlock& out) override {
        std::memset(&out, 0, sizeof(out));
        out.type = Type();
        add = vec2{};
        out.u.minimal.uDensity = density_;
    }
    float GetSigmaTMax() const override { return density_; }
};

// VolumetricCloudVoxelMaterial.{h,cpp}
struct VolumetricCloudVoxelMaterial : IVolumetricCloudMaterial {
    float lod_bias_ = 2.75f, density_ = 20.0f;
    vec2 base_{0.0f, 0.0f}, width_{2.0f, 2.0f};
    int dim_[3] = {1, 1, 1};
    const char* TypeName() const override { return "class VolumetricCloudVoxelMaterial"; }
    int Type() const override { return SKY_MATERIAL_VOXEL; }
    void Load(Archive& ar) override { ar("lod_bias_", lod_bias_); ar("density_", density_); ar("base_", base_); ar("width_", width_); }
    void SetVoxelDim(int x, int y, int z) override { dim_[0] = x; dim_[1] = y; dim_[2] = z; }
    void Update(vec2 viewport, const Camera& camera, dvec2 off, float, vec2& add, SkyMaterialBlock& out) override {  // .cpp:93-105
        std::memset(&out, 0, sizeof(out));
        out.type = Type();
        add = vec2{};
        auto& b = out.u.voxel;
        b.uLodBias = lod_bias_;
        b.uDensity = density_;
        b.uSampleFrequency[0] = 1.0f / width_.x;
        b.uSampleFrequency[1] = 1.0f / width_.y;
        b.uSampleBias[0] = float((off.x + double(base_.x)) / double(width_.x));
        b.uSampleBias[1] = float((off.y + double(base_.y)) / double(width_.y));
        float max_width = float(std::max(dim_[0], std::max(dim_[1], dim_[2])));
        float tan_half_fovy = std::tan(radians(camera.fovy) * 0.5f);
        b.uSampleLodK = max_width * tan_half_fovy / (std::min(width_.x, width_.y) * std::min(viewport.x, viewport.y));
    }
    float GetSigmaTMax() const override { return density_; }
};

// factory table, IVolumetricCloudMaterial.cpp:7-12 (registered under the MSVC typeid spellings)
std::unique_ptr<IVolumetricCloudMaterial> CreateMaterial(const std::string& type) {
    if (type == "class VolumetricCloudDefaultMaterial0") return std::make_unique<VolumetricCloudDefaultMaterial0>();
    if (type == "class VolumetricCloudDefaultMaterial1") return std::make_unique<VolumetricCloudDefaultMaterial1>();
    if (type == "class VolumetricCloudMinimalMaterial") return std::make_unique<VolumetricCloudMinimalMaterial>();
    if (type == "class VolumetricCloudVoxelMaterial") return std::make_unique<VolumetricCloudVoxelMaterial>();
    throw std::runtime_error("unknown material type '" + type + "'");
}

}  // namespace

// ---- VolumetricCloud --------------------------------------------------------------------------------
VolumetricCloud::VolumetricCloud() { material = CreateMaterial("class VolumetricCloudDefaultMaterial0"); }  // VolumetricCloud.cpp:55

template <class Ar>
static void cloud_fields(VolumetricCloud& c, Ar& ar) {
    ar("bottom_altitude_", c.bottom_altitude_); ar("thickness_", c.thickness_);
    ar("max_raymarch_distance_", c.max_raymarch_distance_); ar("max_raymarch_steps_", c.max_raymarch_steps_);
    ar("max_visible_distance_", c.max_visible_distance_); ar("env_color_", c.env_color_);
    ar("env_color_scale_", c.env_color_scale_); ar("sun_illuminance_scale_", c.sun_illuminance_scale_);
    ar("shadow_steps_", c.shadow_steps_); ar("shadow_distance_", c.shadow_distance_);
    ar("shadow_map_max_distance", c.shadow_map_max_distance); ar("shadow_froxel_max_distance", c.shadow_froxel_max_distance);
    ar("sun_multiscattering_sigma_scale", c.sun_multiscattering_sigma_scale);
    ar("env_multiscattering_sigma_scale", c.env_multiscattering_sigma_scale);
    ar("env_bottom_visibility", c.env_bottom_visibility); ar("env_sun_height_curve_exp", c.env_sun_height_curve_exp);
}

void VolumetricCloud::Load(Archive& ar) {
    // polymorphic unique_ptr: {"type": typeid-name, "data": {...}} (serialization.h:190-204)
    if (const Json* m = ar.get("material")) {
        if (m->type == Json::Object) {
            const Json* type = m->find("type");
            const Json* data = m->find("data");
            Archive::require(type && type->type == Json::String && data && data->type == Json::Object, "material");
            material = CreateMaterial(type->str);
            Archive sub{true, const_cast<Json*>(data), ar.log};
            material->Load(sub);
        } else {
            material.reset();  // null material: VolumetricCloud::Update returns early (VolumetricCloud.cpp:171-172)
        }
    }
    cloud_fields(*this, ar);
}

void VolumetricCloud::Save(Archive& ar) {
    Json& m = ar.node->set("material");
    if (material) {
        m.type = Json::Object;
        m.set("type") = Json::string(material->TypeName());
        Json& data = m.set("data");
        data.type = Json::Object;
        Archive sub{false, &data, nullptr};
        material->Load(sub);
    } else {
        m = Json();
    }
    cloud_fields(*this, ar);
}

// VolumetricCloud.cpp:138-166
static mat4 GetLightProjection(const Camera& camera, const mat4& light_view, const mat4& inv_model, float max_distance) {
    vec3 far_plane_center = camera.position_ + camera.front_ * max_distance;
    float tanHalfFovy = std::tan(radians(camera.fovy) * 0.5f);
    vec3 up = (tanHalfFovy * max_distance) * camera.up_;
    vec3 right = (camera.aspect_ * tanHalfFovy * max_distance) * camera.right_;
    vec3 vertices_world[] = {camera.position_, far_plane_center + up + right, far_plane_center + up - right,
                             far_plane_center - up + right, far_plane_center - up - right};
    float min_xy[2] = {1e10f, 1e10f}, max_xy[2] = {-1e10f, -1e10f};
    for (const vec3& vertex_world : vertices_world) {
        vec4 vertex_local = inv_model * vec4(vertex_world, 1.0f);
        vec4 vertex_light = light_view * vertex_local;
        for (int i = 0; i < 2; ++i) {
            min_xy[i] = std::min(min_xy[i], vertex_light[i]);
            max_xy[i] = std::max(max_xy[i], vertex_light[i]);
        }
    }
    const float kShadowMapResolution[2] = {512.0f, 512.0f};  // VolumetricCloud.cpp:52
    for (int i = 0; i < 2; ++i) {
        float padding = 0.5f * (max_xy[i] - min_xy[i]) / kShadowMapResolution[i];
        min_xy[i] -= padding;
        max_xy[i] += padding;
    }
    return ortho(min_xy[0], max_xy[0], min_xy[1], max_xy[1], -1.0f, 1.0f);
}

// VolumetricCloud.cpp:168-280
void VolumetricCloud::Update(const Camera& camera, const Earth& earth, vec3 sun_direction,
                             float aerial_perspective_lut_max_distance, float delta_time,
                             SkyCloudCommonBufferData& common_buffer, SkyCloudBufferData& buffer,
                             SkyMaterialBlock& material_out) {
    if (viewport_w == 0 || viewport_h == 0) throw std::runtime_error("Volumetric cloud viewport is undefined");
    if (!material) throw std::runtime_error("Volumetric cloud has no material");

    float earth_radius = earth.parameters.bottom_radius;
    vec3 earth_center = earth.center();
    vec3 camera_pos = camera.position_;
    vec3 up = normalize(camera_pos - earth_center);
    vec3 origin = earth_center + up * earth_radius;
    vec3 front = normalize(cross(up, vec3(1, 0, 0)));
    vec3 right = normalize(cross(up, front));
    mat4 model = mat4::from_columns(vec4(front, 0), vec4(right, 0), vec4(up, 0), vec4(origin, 1));
    mat4 inv_model = inverse(model);
    mat4 mvp = camera.ViewProjection() * model;

    vec3 pre_camera_pos = camera_pos_;
    mat4 pre_mvp = mvp_;
    vec3 delta_world = camera_pos - pre_camera_pos;
    vec3 delta_local = upper3(inv_model) * delta_world;

    vec2 additional_delta{};
    material->Update(vec2{float(viewport_w), float(viewport_h)}, camera,
                     dvec2{offset_from_first_.x + double(delta_local.x), offset_from_first_.y + double(delta_local.y)},
                     delta_time, additional_delta, material_out);
    delta_local.x += additional_delta.x;
    delta_local.y += additional_delta.y;
    mat4 delta_mat;
    delta_mat.at(3, 0) = delta_local.x;
    delta_mat.at(3, 1) = delta_local.y;
    mat4 additional_delta_only_mat;
    additional_delta_only_mat.at(3, 0) = additional_delta.x;
    additional_delta_only_mat.at(3, 1) = additional_delta.y;

    vec4 lc = inv_model * vec4(camera_pos, 1.0f);
    vec3 local_camera_pos(lc.x, lc.y, lc.z);
    vec3 local_sun_direction = normalize(upper3(inv_model) * sun_direction);

    vec3 light_view_up = (local_sun_direction.x == 0.0f && local_sun_direction.y == 0.0f) ? vec3(1, 0, 0) : vec3(0, 0, 1);
    mat4 light_view = lookAt(local_camera_pos, local_camera_pos - local_sun_direction, light_view_up);
    mat4 light_projection = GetLightProjection(camera, light_view, inv_model, shadow_map_max_distance);
    mat4 light_vp = light_projection * light_view;
    mat4 inv_light_vp = inverse(light_vp);
    mat4 pre_model = model_;
    mat4 pre_light_vp = light_vp_;

    std::memset(&common_buffer, 0, sizeof(common_buffer));
    inverse(mvp).store(common_buffer.uInvMVP);
    (pre_mvp * delta_mat).store(common_buffer.uReprojectMat);
    (pre_light_vp * (inverse(pre_model) * model) * additional_delta_only_mat * inv_light_vp).store(common_buffer.uShadowMapReprojectMat);
    light_vp.store(common_buffer.uLightVP);
    inv_light_vp.store(common_buffer.uInvLightVP);
    common_buffer.uCameraPos[0] = local_camera_pos.x;
    common_buffer.uCameraPos[1] = local_camera_pos.y;
    common_buffer.uCameraPos[2] = local_camera_pos.z;
    common_buffer.uBaseShadingIndex = uint32_t(frame_id_ & 0x3);
    common_buffer.uLinearDepthParam[0] = 1.0f / camera.zNear;
    common_buffer.uLinearDepthParam[1] = (camera.zFar - camera.zNear) / (camera.zFar * camera.zNear);
    common_buffer.uEarthRadius = earth_radius;
    common_buffer.uSunDirection[0] = local_sun_direction.x;
    common_buffer.uSunDirection[1] = local_sun_direction.y;
    common_buffer.uSunDirection[2] = local_sun_direction.z;
    common_buffer.uBottomAltitude = bottom_altitude_;
    common_buffer.uTopAltitude = bottom_altitude_ + thickness_;
    common_buffer.uFrameID = float(frame_id_);
    common_buffer.uShadowFroxelMaxDistance = shadow_froxel_max_distance;
    common_buffer.uAerialPerspectiveLutMaxDistance = aerial_perspective_lut_max_distance;
    common_buffer.uInvShadowFroxelMaxDistance = 1.0f / shadow_froxel_max_distance;  // GetShadowFroxel(), VolumetricCloud.h:71-73

    std::memset(&buffer, 0, sizeof(buffer));
    buffer.uMaxRaymarchDistance = max_raymarch_distance_;
    buffer.uMaxRaymarchSteps = max_raymarch_steps_;
    buffer.uMaxVisibleDistance = max_visible_distance_;
    buffer.uEnvColorScale[0] = env_color_.x * env_color_scale_;
    buffer.uEnvColorScale[1] = env_color_.y * env_color_scale_;
    buffer.uEnvColorScale[2] = env_color_.z * env_color_scale_;
    buffer.uSunIlluminanceScale = sun_illuminance_scale_;
    buffer.uShadowSteps = shadow_steps_;
    buffer.uShadowDistance = shadow_distance_;
    buffer.uSunMultiscatteringSigmaScale = sun_multiscattering_sigma_scale;
    buffer.uEnvMultiscatteringSigmaScale = env_multiscattering_sigma_scale;
    buffer.uEnvBottomVisibility = env_bottom_visibility;
    buffer.uEnvSunHeightCurveExp = env_sun_height_curve_exp;

    offset_from_first_.x += double(delta_local.x);
    offset_from_first_.y += double(delta_local.y);
    frame_id_ = (frame_id_ + 1) & 0xff;
    camera_pos_ = camera_pos;
    mvp_ = mvp;
    light_vp_inv_model_ = light_vp * inv_model;
    model_ = model;
    light_vp_ = light_vp;
}

// VolumetricCloud.cpp:505-519
void VolumetricCloud::PathTracingInit(SkyPathTracingInit& out) const {
    std::memset(&out, 0, sizeof(out));
    const PathTracingInitParam& p = path_tracing_init_param_;
    out.sqrt_tile_count = p.sqrt_tile_count;
    out.max_bounces = p.max_bounces;
    out.region_box_half_width = p.region_box_half_width;
    out.importance_sampling = p.importance_sampling ? 1 : 0;
    out.forward_phase_g = p.forward_phase_g;
    out.back_phase_g = p.back_phase_g;
    out.forward_scattering_ratio = p.forward_scattering_ratio;
    out.prng = p.prng;
    out.environment_lighting = p.environment_lighting;
    out.sigma_t_max = material ? material->GetSigmaTMax() : 0.0f;
    for (int c = 0; c < 3; ++c) for (int r = 0; r < 3; ++r) out.model_matrix3[c * 3 + r] = model_.at(c, r);
}

// VolumetricCloud.cpp:571-581
void VolumetricCloud::GetRenderRegion(int tile_index, int region[4]) const {
    int n = path_tracing_init_param_.sqrt_tile_count;
    int x = tile_index % n, y = tile_index / n;
    int vp[2] = {viewport_w, viewport_h};
    auto get_index = [&](int group_index, int axis) {
        int region_size_base = vp[axis] / n, big_region_count = vp[axis] % n;
        return std::min(group_index, big_region_count) * (region_size_base + 1) + std::max(group_index - big_region_count, 0) * region_size_base;
    };
    region[0] = get_index(x, 0); region[1] = get_index(y, 1);
    region[2] = get_index(x + 1, 0); region[3] = get_index(y + 1, 1);
}

// ---- Scene --------------------------------------------------------------------------------------------
void Scene::Load(const std::string& text) {
    root = JsonParser(text).parse();
    if (root.type != Json::Object) throw std::runtime_error("config root is not an object");
    log.clear();
    Archive ar{true, &root, &log};
    ar("earth_", earth_);
    if (const Json* c = ar.get("volumetric_cloud_")) {
        Archive sub{true, const_cast<Json*>(c), &log};
        volumetric_cloud_.Load(sub);
    }
    ar("camera_", camera_);
    ar("atmosphere_render_init_parameters_", atmosphere_render_init_parameters_);
    ar("atmosphere_render_parameters_", atmosphere_render_parameters_);
}

std::string Scene::Save() {
    Json out = root;
    out.type = Json::Object;
    Archive ar{false, &out, nullptr};
    ar("earth_", earth_);
    {
        Json& c = out.set("volumetric_cloud_");
        c = Json();
        c.type = Json::Object;
        Archive sub{false, &c, nullptr};
        volumetric_cloud_.Save(sub);
    }
    ar("camera_", camera_);
    ar("atmosphere_render_init_parameters_", atmosphere_render_init_parameters_);
    ar("atmosphere_render_parameters_", atmosphere_render_parameters_);
    std::string text;
    json_write(out, text);
    return text;
}

void Scene::LutConfig(SkyLutConfig& c) const {
    std::memset(&c, 0, sizeof(c));
    const auto& p = atmosphere_render_init_parameters_;
    c.sky_view_width = 128;   // AtmosphereRenderer.cpp:15-16
    c.sky_view_height = 128;
    c.aerial_perspective_depth = p.aerial_perspective_lut_depth;
    c.environment_size = 128;  // AtmosphereRenderer.cpp:23
    c.use_sky_view_lut = p.use_sky_view_lut;
    c.use_aerial_perspective_lut = p.use_aerial_perspective_lut;
    c.sky_view_dither = p.sky_view_lut_dither_sample_point_enable;
    c.aerial_perspective_dither = p.aerial_perspective_lut_dither_sample_point_enable;
    c.raymarching_dither = p.raymarching_dither_sample_point_enable;
    c.moon_shadow = p.moon_shadow_enable;            // MOON_SHADOW_ENABLE, AtmosphereRenderer.cpp:101
    c.volumetric_light = p.volumetric_light_enable;  // VOLUMETRIC_LIGHT_ENABLE, :100 (reads SKY_RES_MESH_SHADOW_MAP)
    c.pcss = p.pcss_enable;                          // PCSS_ENABLE, :99 (object pixels; reads SKY_RES_MESH_SHADOW_MAP)
}

// AtmosphereRenderer.cpp:52-83 and :168-174
void Scene::AtmosphereRenderBuffer(SkyAtmosphereRenderBufferData& d) {
    std::memset(&d, 0, sizeof(d));
    const AtmosphereRenderParameters& p = atmosphere_render_parameters_;
    auto put = [](float* dst, vec3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; };
    vec3 sun_direction = FromThetaPhiToDirection(radians(p.sun_direction_theta), radians(p.sun_direction_phi));
    put(d.sun_direction, sun_direction);
    d.star_luminance_scale = p.star_luminance_scale;
    vec3 camera_position = camera_.position_;  // AppWindow.cpp:208
    put(d.camera_position, camera_position);
    inverse(camera_.ViewProjection()).store(d.inv_view_projection);
    // AppWindow::Render (AppWindow.cpp:148-157) + ComputeLightMatrix (src/Base/src/ShadowMap.cpp:53-66): the ortho frustum of
    // the mesh shadow map, an 8 km square around the origin seen from 5 km along the sun direction, 0..50 km deep
    {
        vec3 light_direction = sun_direction;
        vec3 position = light_direction * 5.0f;
        vec3 right_direction = normalize(cross(light_direction, vec3(0, 1, 0)));
        vec3 up_direction = light_direction.y == 1.0f ? vec3(1, 0, 0) : normalize(cross(right_direction, light_direction));
        mat4 view_matrix = lookAt(position, position - light_direction, up_direction);
        mat4 projection_matrix = ortho(-4.0f, 4.0f, -4.0f, 4.0f, 0.0f, 5e1f);
        (projection_matrix * view_matrix).store(d.light_view_projection);
    }
    d.raymarching_steps = p.raymarching_steps;
    d.sky_view_lut_steps = p.sky_view_lut_steps;
    d.aerial_perspective_lut_steps = p.aerial_perspective_lut_steps;
    d.aerial_perspective_lut_max_distance = p.aerial_perspective_lut_max_distance;
    mat4 moon = earth_.moon_model();
    put(d.moon_position, vec3(moon.at(3, 0), moon.at(3, 1), moon.at(3, 2)));
    d.moon_radius = earth_.moon_status.radius;
    // AppWindow.cpp:149-163
    float sun_angular_radius = radians(earth_.parameters.sun_angular_radius);
    d.pcss_size_k = sun_angular_radius * (5e1f - 0.0f) / 4.0f;
    d.blocker_kernel_size_k = 2.0f * d.pcss_size_k;

    vec3 earth_center = earth_.center();
    put(d.earth_center, earth_center);
    d.camera_earth_center_distance = distance(camera_position, earth_center);
    vec3 up_direction = normalize(camera_position - earth_center);
    put(d.up_direction, up_direction);
    vec3 right_direction = cross(sun_direction, up_direction);
    float right_direction_length = length(right_direction);
    if (right_direction_length != 0.0f) {
        right_direction = right_direction / right_direction_length;
    } else {
        vec3 sun_direction_biased = FromThetaPhiToDirection(radians(p.sun_direction_theta + 90.f), radians(p.sun_direction_phi));
        right_direction = normalize(cross(sun_direction_biased, up_direction));
    }
    put(d.right_direction, right_direction);
    put(d.front_direction, cross(up_direction, right_direction));

    volumetric_cloud_.light_vp_inv_model_.store(d.uCloudShadowMapMat);
    d.uInvShadowFroxelMaxDistance = 1.0f / volumetric_cloud_.shadow_froxel_max_distance;

    sun_direction_ = sun_direction;
    aerial_perspective_lut_max_distance_ = d.aerial_perspective_lut_max_distance;
}

// Earth::RenderToGBuffer (Earth.cpp:46-53): the uniform block of the ground pass K7
void Scene::EarthBuffer(SkyEarthBufferData* out) const {
    SkyEarthBufferData b{};
    const mat4 view_projection = camera_.ViewProjection();
    view_projection.store(b.view_projection);
    inverse(view_projection).store(b.inv_view_projection);
    const vec3 camera_position = camera_.position_, earth_center = earth_.center();
    const vec3 up = normalize(camera_position - earth_center);
    b.camera_position[0] = camera_position.x; b.camera_position[1] = camera_position.y; b.camera_position[2] = camera_position.z;
    b.earth_center[0] = earth_center.x; b.earth_center[1] = earth_center.y; b.earth_center[2] = earth_center.z;
    b.camera_earth_center_distance = distance(camera_position, earth_center);
    b.up_direction[0] = up.x; b.up_direction[1] = up.y; b.up_direction[2] = up.z;
    *out = b;
}

// EarthRender.frag:40-52 + Atmosphere.glsl:57-69 evaluated in fp32 like the shader
void Scene::GroundDepth(float* depth, int width, int height) const {
    const mat4 view_projection = camera_.ViewProjection();
    const mat4 inv_view_projection = inverse(view_projection);
    const vec3 camera_position = camera_.position_;
    const vec3 earth_center = earth_.center();
    const float bottom_radius = earth_.parameters.bottom_radius;
    const float r = distance(camera_position, earth_center);
    const vec3 up_direction = normalize(camera_position - earth_center);
#pragma omp parallel for schedule(static)
    for (int py = 0; py < height; ++py)
        for (int px = 0; px < width; ++px) {
            float u = (float(px) + 0.5f) / float(width), v = (float(py) + 0.5f) / float(height);
            vec4 h = inv_view_projection * vec4(u * 2.0f - 1.0f, v * 2.0f - 1.0f, 1.0f, 1.0f);  // cleared depth = 1
            vec3 fragment_position(h.x / h.w, h.y / h.w, h.z / h.w);
            vec3 view_direction = normalize(fragment_position - camera_position);
            float mu = dot(view_direction, up_direction);
            float d = 1.0f;
            float discriminant = r * r * (mu * mu - 1.0f) + bottom_radius * bottom_radius;
            if (mu < 0.0f && discriminant >= 0.0f) {
                float dist = std::max(-r * mu - std::sqrt(std::max(discriminant, 0.0f)), 0.0f);
                if (dist < distance(fragment_position, camera_position)) {
                    vec3 ground_position = camera_position + view_direction * dist;
                    vec4 clip = view_projection * vec4(ground_position, 1.0f);
                    float z = clip.z / clip.w * 0.5f + 0.5f;
                    double q = std::floor(double(std::min(std::max(z, 0.0f), 1.0f)) * 16777215.0 + 0.5) / 16777215.0;  // D24
                    d = float(q);
                }
            }
            depth[size_t(py) * width + px] = d;
        }
}

// EarthRender.frag:40-59: what the ground pass writes into the three G-buffer targets (GBuffer.cpp:19-21) for the pixels it
// keeps -- Normal = normalize(ground_position - earth_center), ORM = (1, roughness 1, metallic 0), Albedo from the earth map
// (data/NASA/*.jpg + a seamless textureGrad: an external asset that is not shipped; here the caller's colour) -- pixels it
// discards keep the cleared value 0.
void Scene::GroundGBuffer(const float albedo_rgb[3], uint8_t* albedo, int16_t* normal, uint16_t* orm, int width, int height) const {
    const mat4 view_projection = camera_.ViewProjection();
    const mat4 inv_view_projection = inverse(view_projection);
    const vec3 camera_position = camera_.position_;
    const vec3 earth_center = earth_.center();
    const float bottom_radius = earth_.parameters.bottom_radius;
    const float r = distance(camera_position, earth_center);
    const vec3 up_direction = normalize(camera_position - earth_center);
    auto unorm = [](float v, float m) { return std::nearbyint(std::min(std::max(v, 0.0f), 1.0f) * m); };
    auto snorm16 = [](float v) { return int16_t(std::nearbyint(std::min(std::max(v, -1.0f), 1.0f) * 32767.0f)); };
#pragma omp parallel for schedule(static)
    for (int py = 0; py < height; ++py)
        for (int px = 0; px < width; ++px) {
            const size_t o = (size_t(py) * width + px) * 4;
            for (int k = 0; k < 4; ++k) { albedo[o + k] = 0; normal[o + k] = 0; orm[o + k] = 0; }
            float u = (float(px) + 0.5f) / float(width), v = (float(py) + 0.5f) / float(height);
            vec4 h = inv_view_projection * vec4(u * 2.0f - 1.0f, v * 2.0f - 1.0f, 1.0f, 1.0f);
            vec3 fragment_position(h.x / h.w, h.y / h.w, h.z / h.w);
            vec3 view_direction = normalize(fragment_position - camera_position);
            float mu = dot(view_direction, up_direction);
            float discriminant = r * r * (mu * mu - 1.0f) + bottom_radius * bottom_radius;
            if (!(mu < 0.0f && discriminant >= 0.0f)) continue;
            float dist = std::max(-r * mu - std::sqrt(std::max(discriminant, 0.0f)), 0.0f);
            if (dist >= distance(fragment_position, camera_position)) continue;
            vec3 ground_position = camera_position + view_direction * dist;
            vec3 n = normalize(ground_position - earth_center);
            for (int k = 0; k < 3; ++k) albedo[o + k] = uint8_t(unorm(albedo_rgb[k], 255.0f));
            albedo[o + 3] = 255;
            normal[o + 0] = snorm16(n.x); normal[o + 1] = snorm16(n.y); normal[o + 2] = snorm16(n.z); normal[o + 3] = 32767;
            orm[o + 0] = 65535; orm[o + 1] = 65535; orm[o + 2] = 0; orm[o + 3] = 65535;
        }
}

}  // namespace skyhost
