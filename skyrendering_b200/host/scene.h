// Host-side mirror of the reference's parameter surface and per-frame uniform maths.
// Member names are the reference's (they ARE the JSON keys: cppreflection's FIELD_DECLARE uses
// the literal member expression, src/Base/include/Serialization.h:7), so the four bin/*.json
// scenes load unchanged.  Reference citations are relative to /root/reference.
#pragma once
#include <memory>
#include <string>

#include "../../include/sky_types.h"
#include "json.h"
#include "linalg.h"

namespace skyhost {

// ---- archive: one fields(ar) per struct drives both Deserialize and Serialize --------------------
// (external/cppreflection/include/reflection/serialization.h:380-392: missing key => message on
//  stderr and the C++ default stays.)
struct Archive {
    bool loading;
    Json* node;
    std::string* log;  // missing-key messages

    const Json* get(const char* key) const {
        const Json* v = node->find(key);
        if (!v && log) { *log += "missing key: "; *log += key; *log += "\n"; }
        return v;
    }
    void operator()(const char* key, float& v) {
        if (loading) { if (const Json* j = get(key)) { require(j->type == Json::Number, key); v = float(j->num); } }
        else node->set(key) = Json::number(double(v));
    }
    void operator()(const char* key, int& v) {
        if (loading) { if (const Json* j = get(key)) { require(j->type == Json::Number, key); v = int(j->num); } }
        else node->set(key) = Json::integer(v);
    }
    void operator()(const char* key, bool& v) {
        if (loading) { if (const Json* j = get(key)) { require(j->type == Json::Bool, key); v = j->b; } }
        else node->set(key) = Json::boolean(v);
    }
    void operator()(const char* key, vec2& v) { vec(key, &v.x, 2); }
    void operator()(const char* key, vec3& v) { vec(key, &v.x, 3); }
    void operator()(const char* key, std::string& v) {
        if (loading) { if (const Json* j = get(key)) { require(j->type == Json::String, key); v = j->str; } }
        else node->set(key) = Json::string(v);
    }
    template <class T>
    auto operator()(const char* key, T& v) -> decltype(v.fields(*this), void()) {
        if (loading) {
            if (const Json* j = get(key)) {
                require(j->type == Json::Object, key);
                Archive sub{true, const_cast<Json*>(j), log};
                v.fields(sub);
            }
        } else {
            Json& j = node->set(key);
            j.type = Json::Object;
            Archive sub{false, &j, log};
            v.fields(sub);
        }
    }
    static void require(bool ok, const char* key) {  // R_ASSERT, external/cppreflection/include/reflection/util.h:9-11
        if (!ok) throw std::runtime_error(std::string("config: wrong type for key '") + key + "'");
    }

private:
    void vec(const char* key, float* p, int n) {  // serialization_ext_glm.h:9-20: glm::vec as arrays
        if (loading) {
            if (const Json* j = get(key)) {
                require(j->type == Json::Array && int(j->arr.size()) == n, key);
                for (int i = 0; i < n; ++i) p[i] = float(j->arr[i].num);
            }
        } else {
            Json& j = node->set(key);
            j.type = Json::Array;
            j.arr.clear();
            for (int i = 0; i < n; ++i) j.arr.push_back(Json::number(double(p[i])));
        }
    }
};

// ---- src/SkyRendering/Atmosphere.h:10-68 --------------------------------------------------------
struct AtmosphereParameters {
    vec3 solar_illuminance{1.f, 1.f, 1.f};
    float sun_angular_radius = 0.5334f * 0.5f;
    float bottom_radius = 6360.0f;
    float thickness = 6420.0f - 6360.0f;
    vec3 ground_albedo{0.4f, 0.4f, 0.4f};
    float rayleigh_exponential_distribution = 8.0f;
    float rayleigh_scattering_scale = 0.0331f;
    vec3 rayleigh_scattering{0.175287f, 0.409607f, 1.0f};
    float mie_exponential_distribution = 1.2f;
    float mie_phase_g = 0.8f;
    float mie_scattering_scale = 0.003996f;
    vec3 mie_scattering{1.0f, 1.0f, 1.0f};
    float mie_absorption_scale = 0.000444f;
    vec3 mie_absorption{1.0f, 1.0f, 1.0f};
    float ozone_center_altitude = 25.0f;
    float ozone_width = 15.0f;
    float ozone_absorption_scale = 0.001881f;
    vec3 ozone_absorption{0.345561f, 1.0f, 0.045189f};
    float transmittance_steps = 40.0f;
    float multiscattering_steps = 30.0f;
    float multiscattering_mask = 1.0f;

    template <class Ar> void fields(Ar& ar) {
        ar("solar_illuminance", solar_illuminance); ar("sun_angular_radius", sun_angular_radius);
        ar("bottom_radius", bottom_radius); ar("thickness", thickness); ar("ground_albedo", ground_albedo);
        ar("rayleigh_exponential_distribution", rayleigh_exponential_distribution);
        ar("rayleigh_scattering_scale", rayleigh_scattering_scale); ar("rayleigh_scattering", rayleigh_scattering);
        ar("mie_exponential_distribution", mie_exponential_distribution); ar("mie_phase_g", mie_phase_g);
        ar("mie_scattering_scale", mie_scattering_scale); ar("mie_scattering", mie_scattering);
        ar("mie_absorption_scale", mie_absorption_scale); ar("mie_absorption", mie_absorption);
        ar("ozone_center_altitude", ozone_center_altitude); ar("ozone_width", ozone_width);
        ar("ozone_absorption_scale", ozone_absorption_scale); ar("ozone_absorption", ozone_absorption);
        ar("transmittance_steps", transmittance_steps); ar("multiscattering_steps", multiscattering_steps);
        ar("multiscattering_mask", multiscattering_mask);
    }
};
// Atmosphere.cpp:49-70
void AssignBufferData(const AtmosphereParameters& parameters, SkyAtmosphereBufferData& data);

// ---- src/SkyRendering/Earth.h:14-37 -------------------------------------------------------------
struct Earth {
    AtmosphereParameters parameters;
    struct MoonStatus {
        float direction_theta = 70.0f, direction_phi = 150.0f, distance = 384401.f / 10.f, radius = 1737.f;
        template <class Ar> void fields(Ar& ar) {
            ar("direction_theta", direction_theta); ar("direction_phi", direction_phi);
            ar("distance", distance); ar("radius", radius);
        }
    } moon_status;
    template <class Ar> void fields(Ar& ar) { ar("parameters", parameters); ar("moon_status", moon_status); }

    vec3 center() const { return {0.0f, -parameters.bottom_radius, 0.0f}; }  // Earth.h:37
    mat4 moon_model() const;                                                  // Earth.cpp:67-77
};

// ---- src/Base/include/Camera.h:10-66, src/Base/src/Camera.cpp:13-37 ------------------------------
struct Camera {
    float fovy = 45.f, zNear = 1e-1f, zFar = 1e3f;
    vec3 position_{}, front_{}, right_{}, up_{};
    float yaw_ = 0, pitch_ = 0, aspect_ = 1.0f;

    template <class Ar> void fields(Ar& ar) {
        ar("fovy", fovy); ar("zNear", zNear); ar("zFar", zFar); ar("position_", position_); ar("front_", front_);
        ar("right_", right_); ar("up_", up_); ar("yaw_", yaw_); ar("pitch_", pitch_); ar("aspect_", aspect_);
    }
    mat4 ViewMatrix() const { return lookAt(position_, position_ + front_, up_); }
    mat4 ProjectionMatrix() const { return perspective(radians(fovy), aspect_, zNear, zFar); }
    mat4 ViewProjection() const { return ProjectionMatrix() * ViewMatrix(); }
    void Rotate(float dPitch, float dYaw);
    void UpdateVectors();
};

// ---- src/SkyRendering/AtmosphereRenderer.h:12-66 -------------------------------------------------
struct AtmosphereRenderParameters {
    float sun_direction_theta = 70.0f, sun_direction_phi = 180.0f, star_luminance_scale = 0.005f;
    float raymarching_steps = 40.f, sky_view_lut_steps = 40.f, aerial_perspective_lut_steps = 40.f;
    float aerial_perspective_lut_max_distance = 100.f;
    template <class Ar> void fields(Ar& ar) {
        ar("sun_direction_theta", sun_direction_theta); ar("sun_direction_phi", sun_direction_phi);
        ar("star_luminance_scale", star_luminance_scale); ar("raymarching_steps", raymarching_steps);
        ar("sky_view_lut_steps", sky_view_lut_steps); ar("aerial_perspective_lut_steps", aerial_perspective_lut_steps);
        ar("aerial_perspective_lut_max_distance", aerial_perspective_lut_max_distance);
    }
};
struct AtmosphereRenderInitParameters {
    bool pcss_enable = true, volumetric_light_enable = true, moon_shadow_enable = false;
    bool raymarching_dither_sample_point_enable = true, use_sky_view_lut = false;
    bool sky_view_lut_dither_sample_point_enable = false, use_aerial_perspective_lut = false;
    bool aerial_perspective_lut_dither_sample_point_enable = false;
    int aerial_perspective_lut_depth = 32;
    template <class Ar> void fields(Ar& ar) {
        ar("pcss_enable", pcss_enable); ar("volumetric_light_enable", volumetric_light_enable);
        ar("moon_shadow_enable", moon_shadow_enable);
        ar("raymarching_dither_sample_point_enable", raymarching_dither_sample_point_enable);
        ar("use_sky_view_lut", use_sky_view_lut);
        ar("sky_view_lut_dither_sample_point_enable", sky_view_lut_dither_sample_point_enable);
        ar("use_aerial_perspective_lut", use_aerial_perspective_lut);
        ar("aerial_perspective_lut_dither_sample_point_enable", aerial_perspective_lut_dither_sample_point_enable);
        ar("aerial_perspective_lut_depth", aerial_perspective_lut_depth);
    }
};

// ---- materials: IVolumetricCloudMaterial.h:9-30 + the four implementations ------------------------
struct NoiseCreateInfo {  // VolumetricCloudDefaultMaterial.h:57-61 (ints on the host)
    int seed, base_frequency;
    float remap_min, remap_max;
};

struct IVolumetricCloudMaterial {
    virtual ~IVolumetricCloudMaterial() = default;
    virtual const char* TypeName() const = 0;  // MSVC typeid spelling used in the JSON (reflection.h:46-48)
    virtual int Type() const = 0;              // SkyMaterialType
    virtual void Load(Archive& ar) = 0;
    // Update(viewport, camera, offset_from_first, additional_delta) + the UBO it uploads
    virtual void Update(vec2 viewport, const Camera& camera, dvec2 offset_from_first, float delta_time,
                        vec2& additional_delta, SkyMaterialBlock& out) = 0;
    virtual float GetSigmaTMax() const = 0;
    // level-0 dimensions of the density textures; for the voxel material they come from the grid
    virtual void SetVoxelDim(int, int, int) {}
    virtual bool NoiseInfo(int /*kind*/, SkyNoiseCreateInfo* /*out[2]*/) const { return false; }
};

struct VolumetricCloudDefaultMaterialCommon {  // VolumetricCloudDefaultMaterial.h:86-139
    NoiseCreateInfo cloud_map_uDensity{0, 3, 0.35f, 0.75f}, cloud_map_uHeight{0, 5, 0.8f, 0.4f};
    NoiseCreateInfo detail_uPerlin{0, 7, 0.23f, 1.0f}, detail_uWorley{0, 11, 1.0f, 0.0f};
    NoiseCreateInfo displacement_uPerlin{0, 6, 0.25f, 0.75f};
    float cloud_map_repeat_size = 18.99f, detail_repeat_size = 5.33f, displacement_repeat_size = 3.51f;
    float lod_bias_ = 2.75f, density_ = 15.0f, wind_speed_ = 0.05f, detail_wind_magnify_ = 1.0f;
    std::string minfilter2d_ = "NEAREST_MIPMAP_NEAREST", minfilter3d_ = "NEAREST_MIPMAP_NEAREST",
                minfilter_displacement_ = "NEAREST_MIPMAP_NEAREST";
    dvec2 detail_offset_from_first_{};

    template <class Ar> void fields(Ar& ar) {
#define SKY_NOISE(key, m) ar(key ".seed", m.seed); ar(key ".base_frequency", m.base_frequency); \
                          ar(key ".remap_min", m.remap_min); ar(key ".remap_max", m.remap_max);
        SKY_NOISE("cloud_map_.buffer.uDensity", cloud_map_uDensity)
        SKY_NOISE("cloud_map_.buffer.uHeight", cloud_map_uHeight)
        SKY_NOISE("detail_.buffer.uPerlin", detail_uPerlin)
        SKY_NOISE("detail_.buffer.uWorley", detail_uWorley)
        SKY_NOISE("displacement_.buffer.uPerlin", displacement_uPerlin)
#undef SKY_NOISE
        ar("cloud_map_.texture.repeat_size", cloud_map_repeat_size);
        ar("detail_.texture.repeat_size", detail_repeat_size);
        ar("displacement_.texture.repeat_size", displacement_repeat_size);
        ar("density_", density_); ar("lod_bias_", lod_bias_); ar("wind_speed_", wind_speed_);
        ar("detail_wind_magnify_", detail_wind_magnify_); ar("minfilter2d_", minfilter2d_);
        ar("minfilter3d_", minfilter3d_); ar("minfilter_displacement_", minfilter_displacement_);
    }
    // VolumetricCloudDefaultMaterial.cpp:78-102
    void Update(vec2 viewport, const Camera& camera, dvec2 offset_from_first, float delta_time, vec2& additional_delta,
                SkyMaterialCommonBufferData& buffer);
    bool NoiseInfo(int kind, SkyNoiseCreateInfo* out) const;
};

// ---- src/SkyRendering/VolumetricCloud.h:16-193 -----------------------------------------------------
struct PathTracingInitParam {  // VolumetricCloud.h:158-168 (GUI-only in the reference, not serialised)
    int sqrt_tile_count = 1, max_bounces = 128;
    float region_box_half_width = 100.0f;
    bool importance_sampling = true;
    float forward_phase_g = 0.85f, back_phase_g = -0.15f, forward_scattering_ratio = 0.7f;
    int prng = SKY_PRNG_PCG;
    int environment_lighting = SKY_ENV_GROUND_MULTI_BOUNCE;
};

struct VolumetricCloud {
    std::unique_ptr<IVolumetricCloudMaterial> material;
    float bottom_altitude_ = 2.0f, thickness_ = 2.0f, max_raymarch_distance_ = 30.0f, max_raymarch_steps_ = 128.0f;
    float max_visible_distance_ = 120.0f;
    vec3 env_color_{1, 1, 1};
    float env_color_scale_ = 0.1f, sun_illuminance_scale_ = 1.0f, shadow_steps_ = 5.0f, shadow_distance_ = 2.0f;
    float shadow_map_max_distance = 20.0f, shadow_froxel_max_distance = 20.0f;
    float sun_multiscattering_sigma_scale = 0.3f, env_multiscattering_sigma_scale = 0.5f;
    float env_bottom_visibility = 0.4f, env_sun_height_curve_exp = 1.0f;

    VolumetricCloud();
    void Load(Archive& ar);
    void Save(Archive& ar);

    // per-frame state, VolumetricCloud.h:111-142
    int viewport_w = 0, viewport_h = 0;
    vec3 camera_pos_{0.0f, 0.0f, 0.0f};
    mat4 mvp_, model_, light_vp_, light_vp_inv_model_;
    dvec2 offset_from_first_{};
    int frame_id_ = 0;

    void SetViewport(int w, int h) { viewport_w = w; viewport_h = h; }
    // VolumetricCloud.cpp:168-280.  sun_direction / aerial_perspective_lut_max_distance are what the
    // previous AtmosphereRenderer::Render stored (AtmosphereRenderer.cpp:173-174).
    void Update(const Camera& camera, const Earth& earth, vec3 sun_direction, float aerial_perspective_lut_max_distance,
                float delta_time, SkyCloudCommonBufferData& common, SkyCloudBufferData& buffer, SkyMaterialBlock& material_out);

    // PathTracing (VolumetricCloud.cpp:495-581)
    PathTracingInitParam path_tracing_init_param_;
    void PathTracingInit(SkyPathTracingInit& out) const;
    void GetRenderRegion(int tile_index, int region[4]) const;
};

// ---- src/SkyRendering/AppWindow.h: the serialised root -------------------------------------------
struct Scene {
    Earth earth_;
    Camera camera_;
    VolumetricCloud volumetric_cloud_;
    AtmosphereRenderInitParameters atmosphere_render_init_parameters_;
    AtmosphereRenderParameters atmosphere_render_parameters_;
    Json root;            // everything else (post-process, GUI flags) is carried through untouched
    std::string log;      // missing-key messages of the last load

    // state AtmosphereRenderer::Render leaves behind for VolumetricCloud::Update
    vec3 sun_direction_{};
    float aerial_perspective_lut_max_distance_ = 0.0f;

    void Load(const std::string& json_text);  // AppWindow::Init, AppWindow.cpp:30-53
    std::string Save();                       // AppWindow::SaveConfig, AppWindow.cpp:128-137

    void LutConfig(SkyLutConfig& out) const;
    // AtmosphereRenderer.cpp:52-83 + :168-174
    void AtmosphereRenderBuffer(SkyAtmosphereRenderBufferData& out);
    // analytic ground depth exactly as EarthRender.frag:40-52 writes gl_FragDepth (else 1.0),
    // quantised to the D24 depth buffer (GBuffer.cpp:22)
    void EarthBuffer(SkyEarthBufferData* out) const;
    void GroundDepth(float* depth, int width, int height) const;
    void GroundGBuffer(const float albedo_rgb[3], uint8_t* albedo, int16_t* normal, uint16_t* orm, int width, int height) const;
};

}  // namespace skyhost
