// skyrender -- headless C++ frame driver over the two C ABIs (include/skyhost.h, include/skyb200.h): what
// AppWindow::HandleDisplayEvent / Render do per frame (src/SkyRendering/AppWindow.cpp:139-181), without a window.
//
//   skyrender <scene.json> <width> <height> [--frames N] [--warmup N] [--spp N] [--vdb file.vdb] [--raw8 file dx dy dz]
//             [--hw-filtering] [--strict] [--overlap] [--pipeline] [--coop-luts] [--objects] [--earth-map file.png|file.jpg] [--out image.ppm] [--dump-rgba8 file]
// --objects shades object pixels like the reference (AtmosphereRenderer.glsl:284-343): the environment-BRDF LUT once, the IBL tail
// of the LUT phase every frame, and the G-buffer of the analytic ground pass (skyhost_ground_gbuffer) bound with sky_set_gbuffer.
// With --earth-map (an 8-bit RGB PNG or JPEG, equirectangular, read by the host library's own readers with the reference's vertical flip) the frame
// runs the reference's own order instead: Clear(gbuffer), Earth::RenderToGBuffer as the kernel K7 (sky_gbuffer_clear, sky_earth_gbuffer) into
// the depth plane and the G-buffer, then the composite with the object branch on what K7 wrote.
//
// The scene JSON is the reference's own config format (bin/config*.json); the host library deserialises it with the
// reference's defaults and computes every uniform block; the CUDA library renders.  There is no Python and no CPU fallback in
// this path.  With --spp the voxel-cloud path tracer runs (scene must use the voxel material; --vdb reads an OpenVDB file
// through host/vdb.cpp, --raw8 reads a raw uint8 [dz][dy][dx] grid), otherwise the real-time frame.  Timing: CUDA events on
// the stream around the measured frames.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/skyb200.h"
#include "../../include/skyhost.h"

namespace {

[[noreturn]] void die(const std::string& what) {
    std::fprintf(stderr, "skyrender: %s\n", what.c_str());
    std::exit(1);
}
void host_ok(int rc, const char* what) {
    if (rc) die(std::string(what) + ": " + skyhost_last_error());
}
SkyContext* g_ctx = nullptr;
void sky_ok(int rc, const char* what) {
    if (rc) die(std::string(what) + ": " + sky_last_error(g_ctx));
}
void cuda_ok(cudaError_t e, const char* what) {
    if (e != cudaSuccess) die(std::string(what) + ": " + cudaGetErrorString(e));
}
std::string read_file(const std::string& path, bool binary = false) {
    std::ifstream f(path, binary ? std::ios::binary : std::ios::in);
    if (!f) die("cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}
std::string dir_of(const std::string& path) {
    size_t p = path.find_last_of('/');
    return p == std::string::npos ? "." : path.substr(0, p);
}

struct Driver {
    SkyScene* scene = nullptr;
    SkyContext* ctx = nullptr;
    int width = 0, height = 0;
    SkyNoiseCreateInfo noise_key[3][2];
    bool noise_valid[3] = {false, false, false};
    SkyCloudCommonBufferData common{};
    SkyCloudBufferData cloud{};
    SkyMaterialBlock material{};
    bool objects = false;  // --objects: the object branch of the composite on the ground pass's G-buffer
    bool ground_pass = false;  // --earth-map: Clear(gbuffer) + Earth::RenderToGBuffer (K7) every frame
    float* gdepth = nullptr; void *gb_albedo = nullptr, *gb_normal = nullptr, *gb_orm = nullptr;

    void earth_update() {  // Earth::Update (Earth.cpp:42-44)
        SkyAtmosphereBufferData a;
        host_ok(skyhost_atmosphere_buffer(scene, &a), "atmosphere_buffer");
        sky_ok(sky_atmosphere_bake(ctx, &a), "atmosphere_bake");
    }
    void atmosphere_luts() {  // AtmosphereRenderer::Render up to the environment cube (AtmosphereRenderer.cpp:164-242)
        SkyAtmosphereRenderBufferData r;
        SkyLutConfig cfg;
        host_ok(skyhost_atmosphere_render_buffer(scene, &r), "atmosphere_render_buffer");
        host_ok(skyhost_lut_config(scene, &cfg), "lut_config");
        sky_ok(sky_atmosphere_luts(ctx, &r, &cfg), "atmosphere_luts");
        if (objects) sky_ok(sky_ibl_precompute(ctx), "ibl_precompute");  // glGenerateTextureMipmap + IBL::Precompute (AtmosphereRenderer.cpp:242-244)
    }
    void cloud_update(float dt) {  // VolumetricCloud::Update (VolumetricCloud.cpp:168-280) + DynamicTexture::GenerateIfParameterChanged
        host_ok(skyhost_cloud_update(scene, dt, &common, &cloud, &material), "cloud_update");
        const int kinds[3] = {SKY_NOISE_CLOUD_MAP, SKY_NOISE_DETAIL, SKY_NOISE_DISPLACEMENT};
        for (int i = 0; i < 3; ++i) {
            SkyNoiseCreateInfo info[2];
            int has = 0;
            host_ok(skyhost_noise_info(scene, kinds[i], info, &has), "noise_info");
            if (!has) continue;
            if (!noise_valid[i] || std::memcmp(info, noise_key[i], sizeof(info)) != 0) {
                sky_ok(sky_noise_generate(ctx, kinds[i], info), "noise_generate");
                std::memcpy(noise_key[i], info, sizeof(info));
                noise_valid[i] = true;
            }
        }
        sky_ok(sky_set_material(ctx, &material), "set_material");
    }
    void frame(const float* depth, void* hdr) {  // one HandleDisplayEvent
        earth_update();
        cloud_update(0.0f);
        sky_ok(sky_cloud_shadow(ctx, &common), "cloud_shadow");
        if (ground_pass) {  // AppWindow::Render: Clear(*gbuffer_), RenderGBuffer -> earth_.RenderToGBuffer (AppWindow.cpp:168-171,192-200)
            SkyEarthBufferData earth;
            host_ok(skyhost_earth_buffer(scene, &earth), "earth_buffer");
            sky_ok(sky_gbuffer_clear(ctx, gdepth, gb_albedo, gb_normal, gb_orm, width, height), "gbuffer_clear");
            sky_ok(sky_earth_gbuffer(ctx, &earth, gdepth, gb_albedo, gb_normal, gb_orm, width, height), "earth_gbuffer");
            depth = gdepth;
        }
        atmosphere_luts();
        sky_ok(sky_composite(ctx, depth, hdr, width, height), "composite");
        sky_ok(sky_cloud_frame(ctx, &common, &cloud, depth, hdr), "cloud_frame");
    }
};

}  // namespace

int main(int argc, char** argv) {
    if (argc < 4) die("usage: skyrender <scene.json> <width> <height> [--frames N] [--warmup N] [--spp N] [--vdb file] [--raw8 file dx dy dz] "
                      "[--hw-filtering] [--strict] [--overlap] [--pipeline] [--coop-luts] [--objects] [--earth-map file.png|file.jpg] [--out image.ppm] [--dump-rgba8 file]");
    const std::string scene_path = argv[1];
    Driver d;
    d.width = std::atoi(argv[2]);
    d.height = std::atoi(argv[3]);
    int frames = 8, warmup = 8, spp = 0, raw_dim[3] = {0, 0, 0};
    bool hw = false, strict = false, overlap = false, pipeline = false, coop_luts = false;
    std::string out_ppm, dump_rgba8, vdb_path, raw8_path, data_dir, earth_map;
    for (int i = 4; i < argc; ++i) {
        std::string a = argv[i];
        auto next = [&]() -> const char* { if (i + 1 >= argc) die("missing value after " + a); return argv[++i]; };
        if (a == "--frames") frames = std::atoi(next());
        else if (a == "--warmup") warmup = std::atoi(next());
        else if (a == "--spp") spp = std::atoi(next());
        else if (a == "--vdb") vdb_path = next();
        else if (a == "--raw8") { raw8_path = next(); for (int k = 0; k < 3; ++k) raw_dim[k] = std::atoi(next()); }
        else if (a == "--hw-filtering") hw = true;
        else if (a == "--strict") strict = true;
        else if (a == "--overlap") overlap = true;
        else if (a == "--pipeline") pipeline = true;
        else if (a == "--coop-luts") coop_luts = true;
        else if (a == "--objects") d.objects = true;
        else if (a == "--earth-map") { earth_map = next(); d.objects = true; }
        else if (a == "--out") out_ppm = next();
        else if (a == "--dump-rgba8") dump_rgba8 = next();
        else if (a == "--data") data_dir = next();
        else die("unknown option " + a);
    }
    if (d.width < 16 || d.height < 16) die("bad viewport");

    // AppWindow::Init (AppWindow.cpp:30-53)
    host_ok(skyhost_scene_load(read_file(scene_path).c_str(), &d.scene), "scene_load");
    host_ok(skyhost_set_viewport(d.scene, d.width, d.height), "set_viewport");
    if (sky_ctx_create(0, nullptr, &d.ctx)) die("sky_ctx_create failed: no CUDA device? (this driver has no CPU path)");
    g_ctx = d.ctx;
    // Textures::Textures (Textures.cpp:19-26): the 64x64 R16 blue noise, rows already in GL order
    if (data_dir.empty()) data_dir = dir_of(scene_path) + "/../skyrendering_b200/data";
    std::string bn = read_file(data_dir + "/blue_noise_64x64.u16", true);
    if (bn.size() != 64 * 64 * 2) die("blue_noise_64x64.u16 has the wrong size");
    sky_ok(sky_set_blue_noise(d.ctx, reinterpret_cast<const uint16_t*>(bn.data())), "set_blue_noise");
    sky_ok(sky_set_viewport(d.ctx, d.width, d.height), "set_viewport");
    sky_ok(sky_set_hw_filtering(d.ctx, hw), "set_hw_filtering");
    sky_ok(sky_set_lut_arithmetic(d.ctx, coop_luts ? SKY_LUT_COOPERATIVE : SKY_LUT_EXACT), "set_lut_arithmetic");
    sky_ok(sky_set_strict_arithmetic(d.ctx, strict), "set_strict_arithmetic");

    int material_type = -1;
    host_ok(skyhost_material_type(d.scene, &material_type), "material_type");
    if (material_type == SKY_MATERIAL_VOXEL) {  // VolumetricCloudVoxelMaterial ctor (VolumetricCloudVoxelMaterial.cpp:40-75)
        std::vector<uint8_t> voxels;
        int dim[3];
        if (!vdb_path.empty()) {
            SkyVdbGrid* g = nullptr;
            host_ok(skyhost_vdb_open(vdb_path.c_str(), &g), "vdb_open");
            SkyVdbInfo info;
            host_ok(skyhost_vdb_info(g, &info), "vdb_info");
            for (int k = 0; k < 3; ++k) dim[k] = info.dim[k];
            voxels.resize(size_t(dim[0]) * dim[1] * dim[2]);
            host_ok(skyhost_vdb_fill_r8(g, voxels.data(), int64_t(voxels.size())), "vdb_fill_r8");
            skyhost_vdb_close(g);
            std::fprintf(stderr, "skyrender: %s: %d x %d x %d voxels, %lld active\n", vdb_path.c_str(), dim[0], dim[1], dim[2], (long long)info.active_voxels);
        } else if (!raw8_path.empty()) {
            std::string raw = read_file(raw8_path, true);
            for (int k = 0; k < 3; ++k) dim[k] = raw_dim[k];
            if (raw.size() != size_t(dim[0]) * dim[1] * dim[2]) die("--raw8: file size does not match dx*dy*dz");
            voxels.assign(raw.begin(), raw.end());
        } else {
            die("the scene uses the voxel material: pass --vdb file.vdb or --raw8 file dx dy dz");
        }
        host_ok(skyhost_set_voxel_dim(d.scene, dim[0], dim[1], dim[2]), "set_voxel_dim");
        sky_ok(sky_voxel_upload(d.ctx, voxels.data(), dim[0], dim[1], dim[2]), "voxel_upload");
    }

    // synthetic depth: the analytic ground the reference's EarthRender pass would rasterise (EarthRender.frag:40-52)
    const size_t npix = size_t(d.width) * d.height;
    std::vector<float> depth_host(npix);
    host_ok(skyhost_ground_depth(d.scene, depth_host.data(), d.width, d.height), "ground_depth");
    float* depth = nullptr;
    void *hdr = nullptr, *rgba8 = nullptr;
    cuda_ok(cudaMalloc(&depth, npix * 4), "cudaMalloc");
    cuda_ok(cudaMalloc(&hdr, npix * 8), "cudaMalloc");
    cuda_ok(cudaMalloc(&rgba8, npix * 4), "cudaMalloc");
    cuda_ok(cudaMemcpy(depth, depth_host.data(), npix * 4, cudaMemcpyHostToDevice), "cudaMemcpy");
    cuda_ok(cudaMemset(hdr, 0, npix * 8), "cudaMemset");

    void *gb_albedo = nullptr, *gb_normal = nullptr, *gb_orm = nullptr;
    if (d.objects && !earth_map.empty()) {  // Textures::Textures: the earth albedo map (Textures.cpp:52-58), then K7 fills the G-buffer every frame
        sky_ok(sky_env_brdf_lut(d.ctx), "env_brdf_lut");
        int32_t mw = 0, mh = 0, mc = 0, mb = 8;
        unsigned char magic[2] = {0, 0};
        if (FILE* f = std::fopen(earth_map.c_str(), "rb")) { if (std::fread(magic, 1, 2, f) != 2) magic[0] = 0; std::fclose(f); }
        const bool jpeg = magic[0] == 0xff && magic[1] == 0xd8;   // the reference's own map (data/NASA/world.topo.bathy...jpg) is a progressive JPEG
        if (jpeg) host_ok(skyhost_jpeg_load(earth_map.c_str(), 1, &mw, &mh, &mc, nullptr, 0), "jpeg_load");
        else host_ok(skyhost_png_load(earth_map.c_str(), 1, &mw, &mh, &mc, &mb, nullptr, 0), "png_load");
        if (mc != 3 || mb != 8) die("--earth-map: an 8-bit RGB PNG or JPEG is expected");
        std::vector<uint8_t> texels(size_t(mw) * mh * 3);
        if (jpeg) host_ok(skyhost_jpeg_load(earth_map.c_str(), 1, nullptr, nullptr, nullptr, texels.data(), int64_t(texels.size())), "jpeg_load");
        else host_ok(skyhost_png_load(earth_map.c_str(), 1, nullptr, nullptr, nullptr, nullptr, texels.data(), int64_t(texels.size())), "png_load");
        sky_ok(sky_set_earth_albedo(d.ctx, texels.data(), mw, mh), "set_earth_albedo");
        cuda_ok(cudaMalloc(&d.gdepth, npix * 4), "cudaMalloc");
        cuda_ok(cudaMalloc(&gb_albedo, npix * 4), "cudaMalloc");
        cuda_ok(cudaMalloc(&gb_normal, npix * 8), "cudaMalloc");
        cuda_ok(cudaMalloc(&gb_orm, npix * 8), "cudaMalloc");
        d.gb_albedo = gb_albedo; d.gb_normal = gb_normal; d.gb_orm = gb_orm;
        d.ground_pass = true;
        sky_ok(sky_set_gbuffer(d.ctx, gb_albedo, gb_normal, gb_orm), "set_gbuffer");
    } else if (d.objects) {  // Textures::Textures bakes the environment-BRDF LUT once (Textures.cpp:60-75); Earth::RenderToGBuffer's targets are inputs
        sky_ok(sky_env_brdf_lut(d.ctx), "env_brdf_lut");
        std::vector<uint8_t> albedo(npix * 4);
        std::vector<int16_t> normal(npix * 4);
        std::vector<uint16_t> orm(npix * 4);
        const float grey[3] = {0.3f, 0.3f, 0.3f};
        host_ok(skyhost_ground_gbuffer(d.scene, grey, albedo.data(), normal.data(), orm.data(), d.width, d.height), "ground_gbuffer");
        cuda_ok(cudaMalloc(&gb_albedo, npix * 4), "cudaMalloc");
        cuda_ok(cudaMalloc(&gb_normal, npix * 8), "cudaMalloc");
        cuda_ok(cudaMalloc(&gb_orm, npix * 8), "cudaMalloc");
        cuda_ok(cudaMemcpy(gb_albedo, albedo.data(), npix * 4, cudaMemcpyHostToDevice), "cudaMemcpy");
        cuda_ok(cudaMemcpy(gb_normal, normal.data(), npix * 8, cudaMemcpyHostToDevice), "cudaMemcpy");
        cuda_ok(cudaMemcpy(gb_orm, orm.data(), npix * 8, cudaMemcpyHostToDevice), "cudaMemcpy");
        sky_ok(sky_set_gbuffer(d.ctx, gb_albedo, gb_normal, gb_orm), "set_gbuffer");
    }

    // frame 0 state: the atmosphere once, so that the first cloud update sees a sun direction
    d.earth_update();
    d.atmosphere_luts();
    sky_ok(sky_set_frame_overlap(d.ctx, overlap), "set_frame_overlap");
    sky_ok(sky_set_frame_pipelining(d.ctx, pipeline), "set_frame_pipelining");

    cudaEvent_t e0, e1;
    cuda_ok(cudaEventCreate(&e0), "cudaEventCreate");
    cuda_ok(cudaEventCreate(&e1), "cudaEventCreate");
    float ms = 0.0f;
    if (spp > 0) {
        if (material_type != SKY_MATERIAL_VOXEL) die("--spp needs a scene with the voxel material");
        d.cloud_update(0.0f);
        sky_ok(sky_cloud_shadow(d.ctx, &d.common), "cloud_shadow");
        d.atmosphere_luts();
        sky_ok(sky_composite(d.ctx, depth, hdr, d.width, d.height), "composite");
        SkyPathTracingInit init;
        host_ok(skyhost_pt_init(d.scene, &init), "pt_init");
        sky_ok(sky_pt_begin(d.ctx, &init), "pt_begin");
        const int32_t region[4] = {0, 0, d.width, d.height};
        cuda_ok(cudaEventRecord(e0, nullptr), "cudaEventRecord");
        sky_ok(sky_pt_samples(d.ctx, &d.common, 1, uint32_t(spp), region), "pt_samples");
        cuda_ok(cudaEventRecord(e1, nullptr), "cudaEventRecord");
        sky_ok(sky_pt_resolve(d.ctx, uint32_t(spp), hdr), "pt_resolve");
        cuda_ok(cudaEventSynchronize(e1), "cudaEventSynchronize");
        cuda_ok(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
        std::printf("{\"mode\": \"path_trace\", \"width\": %d, \"height\": %d, \"spp\": %d, \"ms\": %.3f, \"gsamples_per_s\": %.6f}\n", d.width, d.height, spp, ms,
                    double(npix) * spp / (ms * 1e-3) / 1e9);
    } else {
        for (int f = 0; f < warmup; ++f) {
            cuda_ok(cudaMemsetAsync(hdr, 0, npix * 8, nullptr), "cudaMemsetAsync");
            d.frame(depth, hdr);
        }
        cuda_ok(cudaEventRecord(e0, nullptr), "cudaEventRecord");
        for (int f = 0; f < frames; ++f) {
            cuda_ok(cudaMemsetAsync(hdr, 0, npix * 8, nullptr), "cudaMemsetAsync");
            d.frame(depth, hdr);
        }
        cuda_ok(cudaEventRecord(e1, nullptr), "cudaEventRecord");
        sky_ok(sky_sync(d.ctx), "sync");
        cuda_ok(cudaEventSynchronize(e1), "cudaEventSynchronize");
        cuda_ok(cudaEventElapsedTime(&ms, e0, e1), "cudaEventElapsedTime");
        std::printf("{\"mode\": \"frame\", \"width\": %d, \"height\": %d, \"frames\": %d, \"warmup\": %d, \"ms_per_frame\": %.4f}\n", d.width, d.height, frames, warmup,
                    frames > 0 ? ms / frames : 0.0f);
    }

    // display pass (BloomPass2.frag tone map, no bloom) and output
    SkyToneMapParams tm{1, 10.0f, 0, 0};
    sky_ok(sky_tonemap(d.ctx, hdr, d.width, d.height, &tm, rgba8), "tonemap");
    sky_ok(sky_sync(d.ctx), "sync");
    if (!out_ppm.empty() || !dump_rgba8.empty()) {
        std::vector<uint8_t> img(npix * 4);
        cuda_ok(cudaMemcpy(img.data(), rgba8, npix * 4, cudaMemcpyDeviceToHost), "cudaMemcpy");
        if (!dump_rgba8.empty()) {
            std::ofstream f(dump_rgba8, std::ios::binary);
            f.write(reinterpret_cast<const char*>(img.data()), std::streamsize(img.size()));
        }
        if (!out_ppm.empty()) {
            std::ofstream f(out_ppm, std::ios::binary);
            f << "P6\n" << d.width << " " << d.height << "\n255\n";
            for (int y = d.height - 1; y >= 0; --y)  // row 0 is the bottom of the screen
                for (int x = 0; x < d.width; ++x) f.write(reinterpret_cast<const char*>(&img[(size_t(y) * d.width + x) * 4]), 3);
        }
    }
    cudaFree(depth); cudaFree(hdr); cudaFree(rgba8);
    sky_ctx_destroy(d.ctx);
    skyhost_scene_destroy(d.scene);
    return 0;
}
