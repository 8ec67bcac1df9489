// C ABI over scene.h (include/skyhost.h).  No exceptions cross the boundary.
#include "../../include/skyhost.h"

#include <string>

#include "scene.h"
#include "vdb.h"
#include "png.h"
#include "jpeg.h"
#include <cstring>
#include <stdexcept>

using namespace skyhost;

struct SkyScene {
    Scene scene;
};
struct SkyVdbGrid {
    std::unique_ptr<VdbGrid> grid;
};

static thread_local std::string g_error;

template <class F>
static int guarded(F&& f) {
    try {
        f();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

extern "C" {

const char* skyhost_last_error(void) { return g_error.c_str(); }

int skyhost_scene_load(const char* json_text, SkyScene** out) {
    *out = nullptr;
    SkyScene* s = new SkyScene();
    int rc = guarded([&] { s->scene.Load(json_text); });
    if (rc) { delete s; return rc; }
    *out = s;
    return 0;
}
void skyhost_scene_destroy(SkyScene* s) { delete s; }
const char* skyhost_scene_log(SkyScene* s) { return s->scene.log.c_str(); }

int64_t skyhost_scene_save(SkyScene* s, char* buf, int64_t capacity) {
    std::string text;
    if (guarded([&] { text = s->scene.Save(); })) return -1;
    int64_t need = int64_t(text.size()) + 1;
    if (!buf || capacity < need) return need;
    std::memcpy(buf, text.c_str(), size_t(need));
    return 0;
}

int skyhost_atmosphere_buffer(SkyScene* s, SkyAtmosphereBufferData* out) {
    return guarded([&] { AssignBufferData(s->scene.earth_.parameters, *out); });
}
int skyhost_lut_config(SkyScene* s, SkyLutConfig* out) { return guarded([&] { s->scene.LutConfig(*out); }); }
int skyhost_atmosphere_render_buffer(SkyScene* s, SkyAtmosphereRenderBufferData* out) {
    return guarded([&] { s->scene.AtmosphereRenderBuffer(*out); });
}

int skyhost_set_viewport(SkyScene* s, int width, int height) {
    return guarded([&] {
        if (width <= 0 || height <= 0) throw std::runtime_error("viewport must be positive");
        s->scene.volumetric_cloud_.SetViewport(width, height);
        s->scene.camera_.aspect_ = float(width) / float(height);
    });
}

int skyhost_cloud_update(SkyScene* s, float dt, SkyCloudCommonBufferData* common, SkyCloudBufferData* cloud,
                         SkyMaterialBlock* material) {
    return guarded([&] {
        Scene& sc = s->scene;
        sc.volumetric_cloud_.Update(sc.camera_, sc.earth_, sc.sun_direction_, sc.aerial_perspective_lut_max_distance_, dt,
                                    *common, *cloud, *material);
    });
}

int skyhost_noise_info(SkyScene* s, int kind, SkyNoiseCreateInfo out[2], int* has) {
    return guarded([&] {
        auto& m = s->scene.volumetric_cloud_.material;
        *has = (m && m->NoiseInfo(kind, out)) ? 1 : 0;
    });
}

int skyhost_set_voxel_dim(SkyScene* s, int dx, int dy, int dz) {
    return guarded([&] {
        auto& m = s->scene.volumetric_cloud_.material;
        if (!m) throw std::runtime_error("no material");
        m->SetVoxelDim(dx, dy, dz);
    });
}

int skyhost_vdb_open(const char* path, SkyVdbGrid** out) {
    *out = nullptr;
    return guarded([&] { *out = new SkyVdbGrid{VdbGrid::open(path)}; });
}
int skyhost_vdb_parse(const void* bytes, int64_t size, SkyVdbGrid** out) {
    *out = nullptr;
    return guarded([&] { *out = new SkyVdbGrid{VdbGrid::parse(static_cast<const uint8_t*>(bytes), size_t(size))}; });
}
void skyhost_vdb_close(SkyVdbGrid* g) { delete g; }
int skyhost_vdb_info(SkyVdbGrid* g, SkyVdbInfo* out) {
    return guarded([&] {
        SkyVdbInfo info{};
        g->grid->voxel_dim(info.dim);
        g->grid->bbox(info.bbox_min, info.bbox_max);
        info.file_version = int32_t(g->grid->file_version());
        info.active_voxels = g->grid->active_voxel_count();
        if (!g->grid->metadata_int("file_voxel_count", info.file_voxel_count)) info.file_voxel_count = -1;
        info.has_file_bbox = g->grid->metadata_vec3i("file_bbox_min", info.file_bbox_min) && g->grid->metadata_vec3i("file_bbox_max", info.file_bbox_max);
        info.background = g->grid->background();
        *out = info;
    });
}
static int64_t vdb_voxels(SkyVdbGrid* g) {
    int32_t d[3];
    g->grid->voxel_dim(d);
    return int64_t(d[0]) * d[1] * d[2];
}
int skyhost_vdb_fill_r8(SkyVdbGrid* g, uint8_t* voxels, int64_t capacity) {
    return guarded([&] {
        if (capacity < vdb_voxels(g)) throw std::runtime_error("vdb: output buffer too small");
        g->grid->fill_r8(voxels);
    });
}
int skyhost_vdb_fill_float(SkyVdbGrid* g, float* voxels, int64_t capacity) {
    return guarded([&] {
        if (capacity < vdb_voxels(g)) throw std::runtime_error("vdb: output buffer too small");
        g->grid->fill_dense(voxels);
    });
}

int skyhost_material_type(SkyScene* s, int* type) {
    return guarded([&] {
        auto& m = s->scene.volumetric_cloud_.material;
        *type = m ? m->Type() : -1;
    });
}

int skyhost_pt_params(SkyScene* s, int sqrt_tile_count, int max_bounces, float half_width, int importance_sampling,
                      int prng, int environment_lighting) {
    return guarded([&] {
        if (sqrt_tile_count < 1 || max_bounces < 0) throw std::runtime_error("bad path tracing parameters");
        auto& p = s->scene.volumetric_cloud_.path_tracing_init_param_;
        p.sqrt_tile_count = sqrt_tile_count;
        p.max_bounces = max_bounces;
        p.region_box_half_width = half_width;
        p.importance_sampling = importance_sampling != 0;
        p.prng = prng;
        p.environment_lighting = environment_lighting;
    });
}
int skyhost_pt_init(SkyScene* s, SkyPathTracingInit* out) {
    return guarded([&] { s->scene.volumetric_cloud_.PathTracingInit(*out); });
}
int skyhost_pt_region(SkyScene* s, int tile_index, int32_t region[4]) {
    return guarded([&] {
        int r[4];
        s->scene.volumetric_cloud_.GetRenderRegion(tile_index, r);
        for (int i = 0; i < 4; ++i) region[i] = r[i];
    });
}

int skyhost_camera_get(SkyScene* s, float position[3], float front[3], float* fovy, float* z_near, float* z_far) {
    const Camera& c = s->scene.camera_;
    for (int i = 0; i < 3; ++i) { position[i] = c.position_[i]; front[i] = c.front_[i]; }
    *fovy = c.fovy; *z_near = c.zNear; *z_far = c.zFar;
    return 0;
}
int skyhost_camera_move(SkyScene* s, const float d[3], float d_pitch, float d_yaw) {
    return guarded([&] {
        Camera& c = s->scene.camera_;
        c.position_ = c.position_ + vec3(d[0], d[1], d[2]);
        if (d_pitch != 0.0f || d_yaw != 0.0f) c.Rotate(d_pitch, d_yaw);
    });
}
int skyhost_view_projection(SkyScene* s, float vp[16]) {
    return guarded([&] { s->scene.camera_.ViewProjection().store(vp); });
}

int skyhost_earth_buffer(SkyScene* s, SkyEarthBufferData* out) {
    return guarded([&] { s->scene.EarthBuffer(out); });
}

int skyhost_ground_depth(SkyScene* s, float* depth, int width, int height) {
    return guarded([&] { s->scene.GroundDepth(depth, width, height); });
}

int skyhost_ground_gbuffer(SkyScene* s, const float albedo_rgb[3], uint8_t* albedo, int16_t* normal, uint16_t* orm, int width, int height) {
    return guarded([&] { s->scene.GroundGBuffer(albedo_rgb, albedo, normal, orm, width, height); });
}

int skyhost_png_load(const char* path, int flip_vertically, int32_t* width, int32_t* height, int32_t* channels, int32_t* bits, void* out, int64_t out_bytes) {
    try {
        const skyhost::PngImage im = skyhost::load_png(path ? path : "");
        if (width) *width = im.width;
        if (height) *height = im.height;
        if (channels) *channels = im.channels;
        if (bits) *bits = im.bits;
        if (!out) return 0;
        const size_t bytes_per_sample = size_t(im.bits / 8), stride = size_t(im.width) * im.channels * bytes_per_sample;
        if (out_bytes != int64_t(stride * im.height)) throw std::runtime_error("png_load: the buffer must hold width * height * channels * (bits / 8) bytes");
        uint8_t* dst = static_cast<uint8_t*>(out);
        for (int y = 0; y < im.height; ++y) {
            const uint8_t* src = im.samples.data() + size_t(flip_vertically ? im.height - 1 - y : y) * stride;
            uint8_t* row = dst + size_t(y) * stride;
            if (bytes_per_sample == 1) std::memcpy(row, src, stride);
            else for (size_t i = 0; i < stride; i += 2) { row[i] = src[i + 1]; row[i + 1] = src[i]; }   // big-endian file -> little-endian host
        }
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

int skyhost_jpeg_load(const char* path, int flip_vertically, int32_t* width, int32_t* height, int32_t* channels, void* out, int64_t out_bytes) {
    try {
        static thread_local std::string cached_path;          // the two-call protocol decodes a (progressive, multi-megapixel) file once
        static thread_local skyhost::JpegImage cached;
        const std::string name = path ? path : "";
        if (name != cached_path || cached.samples.empty()) { cached = skyhost::JpegImage{}; cached_path.clear(); cached = skyhost::load_jpeg(name); cached_path = name; }
        const skyhost::JpegImage& im = cached;
        if (width) *width = im.width;
        if (height) *height = im.height;
        if (channels) *channels = im.channels;
        if (!out) return 0;
        const size_t stride = size_t(im.width) * im.channels;
        if (out_bytes != int64_t(stride * im.height)) throw std::runtime_error("jpeg_load: the buffer must hold width * height * channels bytes");
        uint8_t* dst = static_cast<uint8_t*>(out);
        for (int y = 0; y < im.height; ++y) std::memcpy(dst + size_t(y) * stride, im.samples.data() + size_t(flip_vertically ? im.height - 1 - y : y) * stride, stride);
        cached = skyhost::JpegImage{}; cached_path.clear();
        return 0;
    } catch (const std::exception& e) {
        g_error = e.what();
        return 1;
    }
}

}  // extern "C"
