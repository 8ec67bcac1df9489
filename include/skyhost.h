/*
 * skyhost.h -- C ABI of the host-side parameter surface (libskyhost.so, plain C++17, no CUDA).
 *
 * It mirrors the reference host classes that turn a scene JSON + camera into the uniform blocks
 * of sky_types.h; the blocks are then handed to libskyb200.so (skyb200.h).  Citations are
 * relative to /root/reference.  All functions return 0 on success; skyhost_last_error() has
 * the message otherwise (the reference throws std::runtime_error / R_ASSERT).
 */
#ifndef SKYHOST_H
#define SKYHOST_H

#include "sky_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct SkyScene SkyScene;

const char* skyhost_last_error(void);

/* AppWindow::Init (src/SkyRendering/AppWindow.cpp:30-53): parse (comments and trailing commas
 * allowed) and deserialise; keys that are absent keep their C++ defaults and are listed by
 * skyhost_scene_log. */
int skyhost_scene_load(const char* json_text, SkyScene** out);
void skyhost_scene_destroy(SkyScene* scene);
const char* skyhost_scene_log(SkyScene* scene);
/* AppWindow::SaveConfig (AppWindow.cpp:128-137).  Returns the required size (including NUL) when
 * buf is NULL or too small, 0 on success. */
int64_t skyhost_scene_save(SkyScene* scene, char* buf, int64_t capacity);

/* Earth::Update -> AssignBufferData (src/SkyRendering/Atmosphere.cpp:49-70). */
int skyhost_atmosphere_buffer(SkyScene* scene, SkyAtmosphereBufferData* out);
/* AtmosphereRenderInitParameters -> shader permutation flags (AtmosphereRenderer.cpp:91-107). */
int skyhost_lut_config(SkyScene* scene, SkyLutConfig* out);
/* AtmosphereRenderer::Render's AssignBufferData (AtmosphereRenderer.cpp:52-83,168-174); also
 * records sun_direction() / aerial_perspective_lut().max_distance for the next cloud update. */
int skyhost_atmosphere_render_buffer(SkyScene* scene, SkyAtmosphereRenderBufferData* out);

/* VolumetricCloud::SetViewport (VolumetricCloud.cpp:115-118); also sets camera aspect like
 * AppWindow::HandleReshapeEvent (AppWindow.cpp:559-566). */
int skyhost_set_viewport(SkyScene* scene, int width, int height);
/* VolumetricCloud::Update (VolumetricCloud.cpp:168-280) incl. material->Update; delta_time is
 * ImGui::GetIO().DeltaTime in the reference. */
int skyhost_cloud_update(SkyScene* scene, float delta_time, SkyCloudCommonBufferData* common,
                         SkyCloudBufferData* cloud, SkyMaterialBlock* material);
/* Noise parameters of the scene's material (VolumetricCloudDefaultMaterial.h:63-84); returns 1
 * in *has when the material owns that texture. */
int skyhost_noise_info(SkyScene* scene, int kind, SkyNoiseCreateInfo out[2], int* has);
/* Voxel material: level-0 grid size (VolumetricCloudVoxelMaterial.cpp:53). */
int skyhost_set_voxel_dim(SkyScene* scene, int dx, int dy, int dz);
int skyhost_material_type(SkyScene* scene, int* type);

/* VolumetricCloudVoxelMaterial ctor (VolumetricCloudVoxelMaterial.cpp:40-74): what openvdb::io::File::readGrid + the
 * dense fill + the GL_FLOAT -> GL_R8 upload produce for the first grid of an OpenVDB file.  A minimal reader of the
 * published FloatGrid ("Tree_float_5_4_3") layout, file versions 222-224 without ZIP / BLOSC (the shipped
 * data/wdas/wdas_cloud_sixteenth.vdb); anything else fails with a message.  dim = {dx, dy, dz} = vdb {x, z, y}
 * ("swap yz"); voxels [dz][dy][dx], ready for sky_voxel_upload + skyhost_set_voxel_dim. */
typedef struct SkyVdbGrid SkyVdbGrid;
typedef struct SkyVdbInfo {
    int32_t dim[3];            /* dx, dy, dz */
    int32_t bbox_min[3];       /* index-space bounding box of the active values, vdb x, y, z */
    int32_t bbox_max[3];
    int32_t file_version;
    int64_t active_voxels;
    int64_t file_voxel_count;  /* the grid's own "file_voxel_count" metadata, -1 if absent */
    int32_t file_bbox_min[3];  /* the grid's own "file_bbox_min/max" metadata (zeros if absent) */
    int32_t file_bbox_max[3];
    float background;
    int32_t has_file_bbox;
} SkyVdbInfo;
int skyhost_vdb_open(const char* path, SkyVdbGrid** out);
int skyhost_vdb_parse(const void* bytes, int64_t size, SkyVdbGrid** out);
void skyhost_vdb_close(SkyVdbGrid* grid);
int skyhost_vdb_info(SkyVdbGrid* grid, SkyVdbInfo* out);
int skyhost_vdb_fill_r8(SkyVdbGrid* grid, uint8_t* voxels, int64_t capacity);
int skyhost_vdb_fill_float(SkyVdbGrid* grid, float* voxels, int64_t capacity);

/* PathTracing ctor constants + tile schedule (VolumetricCloud.cpp:495-519,571-581). */
int skyhost_pt_params(SkyScene* scene, int sqrt_tile_count, int max_bounces, float region_box_half_width,
                      int importance_sampling, int prng, int environment_lighting);
int skyhost_pt_init(SkyScene* scene, SkyPathTracingInit* out);
int skyhost_pt_region(SkyScene* scene, int tile_index, int32_t region[4]);

/* Camera (src/Base/include/Camera.h): pose access for moving-camera tests. */
int skyhost_camera_get(SkyScene* scene, float position[3], float front[3], float* fovy, float* z_near, float* z_far);
int skyhost_camera_move(SkyScene* scene, const float delta_position[3], float d_pitch, float d_yaw);
int skyhost_view_projection(SkyScene* scene, float view_projection[16]);

/* Earth::RenderToGBuffer's uniform block (src/SkyRendering/Earth.cpp:46-53) for the ground pass K7 (sky_earth_gbuffer). */
int skyhost_earth_buffer(SkyScene* scene, SkyEarthBufferData* out);

/* Synthetic depth input: what the ground pass (shaders/SkyRendering/EarthRender.frag:40-52) writes
 * into the D24 depth buffer, 1.0 elsewhere. */
int skyhost_ground_depth(SkyScene* scene, float* depth, int width, int height);

/* Synthetic G-buffer input for sky_set_gbuffer: what the same ground pass writes into the three colour targets
 * (EarthRender.frag:53-59; formats GBuffer.cpp:19-21) for the pixels it keeps -- normal = the sphere's normal (RGBA16_SNORM),
 * ORM = (1, 1, 0) (RGBA16), albedo RGBA8 = `albedo_rgb` in place of the earth map (data/NASA, an external asset with a seamless
 * textureGrad that is not shipped) -- and 0 (the cleared value) elsewhere.  Host arrays [height][width][4]. */
int skyhost_ground_gbuffer(SkyScene* scene, const float albedo_rgb[3], uint8_t* albedo, int16_t* normal, uint16_t* orm, int width, int height);

/* stbi_load / stbi_load_16 of a PNG file with stbi_set_flip_vertically_on_load (src/Base/src/StbImage.cpp:12-17; Textures.cpp:19-26 loads the
 * 64x64 16-bit blue-noise tile this way).  Two calls: with out == NULL it returns the dimensions, channel count and bits per sample (8 / 16);
 * with a buffer of width * height * channels * (bits / 8) bytes it writes the samples in host byte order, row 0 = the BOTTOM row of the
 * image when flip_vertically != 0 (the GL texel order the reference uploads).  Non-interlaced grey / RGB / grey+alpha / RGBA, 8 or 16 bit. */
int skyhost_png_load(const char* path, int flip_vertically, int32_t* width, int32_t* height, int32_t* channels, int32_t* bits, void* out, int64_t out_bytes);

/* stbi_load of a JPEG file with stbi_set_flip_vertically_on_load (src/Base/src/StbImage.cpp:12-17; Textures.cpp:27-58 loads the earth albedo, the star
 * map and the two moon maps of data/NASA this way: progressive 4:4:4 YCbCr files).  Same two-call protocol as skyhost_png_load: out == NULL returns
 * width, height and channels (1 = grey, 3 = RGB); with a buffer of width * height * channels bytes it writes the 8-bit samples, row 0 = the BOTTOM
 * row of the image when flip_vertically != 0.  Baseline, extended-sequential and progressive Huffman DCT, restart intervals, sampling factors 1 and 2;
 * the inverse DCT, chroma upsampling and YCbCr -> RGB use stb_image's integer arithmetic, so the bytes are the ones the reference uploads
 * (host/jpeg.cpp; pinned against external/stb/stb_image.h in tests/test_jpeg.py).  The result feeds sky_set_star_map / sky_set_earth_albedo. */
int skyhost_jpeg_load(const char* path, int flip_vertically, int32_t* width, int32_t* height, int32_t* channels, void* out, int64_t out_bytes);

#ifdef __cplusplus
}
#endif
#endif /* SKYHOST_H */
