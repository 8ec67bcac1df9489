// SkyContext: everything one GL context owned in the reference (Atmosphere.h:84-91,
// AtmosphereRenderer.h:118-128, VolumetricCloud.h:99-128), as device allocations on one GPU.
#pragma once
#include <algorithm>
#include <string>

#include <nvtx3/nvToolsExt.h>

#include "common.cuh"

constexpr int kMaxMipLevels = 13;
constexpr int kDensityLutSize = 2048;  // 49 m per texel for a 100 km atmosphere: interpolation error of exp(-h / 1.2 km) 2e-4 relative
constexpr int kEarthMaxLevels = 15;  // earth albedo map: up to 16384 texels per axis
constexpr int SKY_PEER_TIMEOUT_SLOT = 16, SKY_PEER_FLAG_SLOTS = 64;  // see k_peer_flags (cloud.cu)  // up to 4096 texels per axis

// A UNORM8 texture with its full mip chain, kept twice: as linear device memory (exact fp32
// software filtering, the default) and as a CUDA mip-mapped array behind two texture objects
// (hardware filtering, opt-in; 8-bit interpolation weights).
struct MipView {
    const uint8_t* base;            // all levels, level l at base + off[l]*channels
    unsigned long long off[kMaxMipLevels];
    int w[kMaxMipLevels], h[kMaxMipLevels], d[kMaxMipLevels];
    int levels;
    int channels;
    cudaTextureObject_t tex_linear;  // LINEAR, level 0 (magnification)
    cudaTextureObject_t tex_point;   // POINT over the mip chain (NEAREST_MIPMAP_NEAREST)
    // Corner-packed copy of EVERY level for the exact path: cell (i,j[,k]) of level l holds the 4 (2-D) or 8 (3-D)
    // texels a bilinear / trilinear tap at base texel (i,j,k) needs, wrap or border already applied, so any
    // fetch -- LINEAR on level 0 or NEAREST on a mip level (corner 0 of the cell) -- is ONE aligned 8- or 16-byte
    // load at a branch-free address.  HBM is cheap on this part (9x the texel bytes), load slots and divergence are not.
    // REPEAT: cell dims = level dims, cell index = base texel.  BORDER: cell dims = (w+3, h+3, d+1) and cell
    // index = base texel + (2, 2, 1): base texels -2..w in x/y (a tap up to a quarter texel outside the half-texel
    // apron needs no range test) and -1..d-1 in z.
    const void* cells;
    unsigned long long cell_off[kMaxMipLevels];  // first cell of level l
    int cell_w[kMaxMipLevels], cell_h[kMaxMipLevels], cell_d[kMaxMipLevels];
    int pad_xy, pad_z;                            // BORDER: 2, 1; REPEAT: 0, 0
};

struct MipTextureDev {
    MipView view{};
    uint8_t* data = nullptr;
    size_t bytes = 0;
    void* cells = nullptr;  // corner-packed levels, see MipView
    size_t cell_bytes = 0;
    cudaMipmappedArray_t array = nullptr;
    bool is3d = false;
    bool border = false;  // CLAMP_TO_BORDER(0) instead of REPEAT
    bool valid = false;
};

template <class T>
struct Lut {  // plain row-major image in device memory
    T* p = nullptr;
    int w = 0, h = 0, d = 1;
    size_t bytes() const { return size_t(w) * h * d * sizeof(T); }
};

struct SkyContext {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string error;
    bool hw_filtering = false;
    bool strict_arithmetic = false;  // sky_set_strict_arithmetic: route K6, K11-K18, K19/K20 to the *_strict objects
    int lut_arithmetic = 0;          // sky_set_lut_arithmetic: SKY_LUT_EXACT (bit-faithful K2-K4) or the lane-cooperative production march
    bool counting = false;
    unsigned long long launch_count = 0;   // kernel launches issued for this context (SKY_LAUNCH_CHECK follows every launch); sky_launch_count
    int k16_group = 0;               // SKYB200_K16_GROUP=4|8: force the wavefront kernel's rays per warp (0: chosen per launch, cloud.cu)
    bool k16_literal = false;        // SKYB200_K16_LITERAL=1: the production object launches k16_render (one lane = one ray, the shader's loop) instead of k16_render_coop
    int out_band_rows = 0, out_band_index = 0, out_band_count = 1;  // sky_set_output_bands: rows K6 / K18 own

    // frame overlap (sky_set_frame_overlap): second lane of a frame, see api.cu
    bool overlap = false;
    cudaStream_t lane2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_shadow = nullptr, ev_pre_composite = nullptr, ev_lane2 = nullptr;
    bool lane2_pending = false;        // lane2 holds work the caller's stream has not been ordered after
    bool shadow_pending = false;       // ... of which the shadow chain (the composite waits for it)
    bool pre_composite_recorded = false;
    const float* pre_composite_depth = nullptr;
    bool lane2_reads_luts = false;     // K16 / K17 are queued on lane2

    // frame pipelining (sky_set_frame_pipelining): the atmosphere LUT phase of frame N+1 (K1-K5) runs on `lut_stream` into the
    // OTHER of two LUT sets while frame N's full-machine kernels (K6, K16) still read theirs, see api.cu
    bool pipelining = false;
    cudaStream_t lut_stream = nullptr;
    cudaEvent_t ev_frame_mark[2] = {nullptr, nullptr}, ev_luts_ready = nullptr;
    int mark_count = 0;                // bake calls since pipelining was enabled
    bool luts_pending = false;         // lut_stream holds work the caller's stream has not been ordered after
    bool bake_since_shadow = false;    // this frame's bake already ordered lut_stream after the frame before last
    cudaEvent_t ev_main_to_lut = nullptr;
    // Sharded frames (sky_set_output_bands, band_count > 1): a rank's share of K6 / K16 / K18 is short, so the serial chain on lut_stream --
    // K1, K2, shadow chain, K3-K5, ~0.4 ms of latency-bound kernels -- becomes the frame time; there the shadow chain gets a stream of its
    // own (same gate event, same double-buffered outputs), which takes ~0.1 ms off that chain.  On one GPU this was measured to be worse.
    cudaStream_t shadow_stream = nullptr;
    cudaEvent_t ev_shadow_ready = nullptr, frame_gate = nullptr;   // frame_gate: the frame mark lut_stream waited for at this frame's bake (borrowed)
    bool shadow_stream_pending = false, shadow_stream_last = false;
    struct LutSet {                    // the alternate copy of everything sky_atmosphere_bake / sky_atmosphere_luts write
        Lut<float4> transmittance, multiscattering, sky_lum, sky_trans, ap_lum, ap_trans;
        Lut<half4> env, transmittance_h, multiscattering_h, density_h, sky_lum_h, sky_trans_h, ap_lum_h, ap_trans_h;
        Lut<uint16_t> shadow_froxel;   // second froxel volume: the shadow chain of frame N+1 also runs on lut_stream
        Lut<float2> shadow_blurred;    // second blurred cloud shadow map (shadow_maps[2]): the object branch of frame N's K6 samples
                                       // it on the caller's stream while K12 of frame N+1 writes the other one on lut_stream
        cudaTextureObject_t density_tex = 0;
        cudaTextureObject_t transmittance_tex = 0, multiscattering_tex = 0, sky_lum_tex = 0, sky_trans_tex = 0, ap_lum_tex = 0, ap_trans_tex = 0;
        const void* lut_tex_key[4] = {nullptr, nullptr, nullptr, nullptr};
        int lut_tex_dims[4][3] = {};
        // the IBL tail of the LUT phase (sky_ibl_precompute) is double-buffered with the LUTs it follows
        half4* env_mips = nullptr; int env_mips_for = 0; size_t env_mips_texels = 0;
        Lut<float4> env_sh; half4* prefiltered = nullptr; size_t prefiltered_texels = 0; bool ibl_valid = false;
    } alt;

    // uniforms last seen
    SkyAtmosphereBufferData atm{};
    SkyAtmosphereRenderBufferData render{};
    SkyLutConfig lut_cfg{};
    SkyMaterialBlock material{};
    SkyPathTracingInit pt{};
    SkyCloudCommonBufferData last_common{};  // uniforms of the frame opened by cloud_frame_begin

    // K1-K5
    Lut<float4> transmittance, multiscattering, sky_lum, sky_trans, ap_lum, ap_trans;
    Lut<half4> env;  // [6][S][S]
    // IBL chain (ibl.cu): K22 RG16 LUT, the environment cube's mips (levels >= 1, concatenated), K23 Llm[9],
    // K24 prefiltered cube (SKY_IBL_ROUGHNESS_COUNT levels from SKY_IBL_PREFILTERED_RESOLUTION, concatenated)
    Lut<ushort2> env_brdf_lut;
    half4* env_mips = nullptr;
    int env_mips_for = 0;         // environment size the mips are allocated for
    size_t env_mips_texels = 0;
    bool ibl_valid = false;       // sky_ibl_precompute has run
    Lut<float4> env_sh;           // w = 9
    half4* prefiltered = nullptr;
    size_t prefiltered_texels = 0;
    const void *gbuffer_albedo = nullptr, *gbuffer_normal = nullptr, *gbuffer_orm = nullptr;  // sky_set_gbuffer (borrowed)
    // K6's raymarch fetches the two bake LUTs twice per step through the texture unit, which is what bounds it (ncu: L1/TEX at
    // 86 % of peak with RGBA32F texels, quarter rate); it reads RGBA16F copies (half rate), refreshed after every bake.  The
    // fp16 rounding (2^-11 relative) is far inside the frame tolerance; the strict objects filter the fp32 LUTs in software.
    Lut<half4> transmittance_h, multiscattering_h;
    cudaTextureObject_t transmittance_tex = 0, multiscattering_tex = 0;  // LINEAR views of those copies
    // K6's march also reads the three density profiles of the atmosphere model -- exp(-h / H_rayleigh), exp(-h / H_mie) and the ozone tent
    // (Atmosphere.glsl:119-132,156-159) -- from a 1-D RGBA16F table over the altitude (kDensityLutSize texels from the ground to the top
    // boundary, LINEAR + CLAMP): one TEX instead of two MUFU.EX2 and eight ALU instructions per step.  Rebuilt with every bake.
    Lut<half4> density_h;
    cudaTextureObject_t density_tex = 0;
    // LINEAR views of the per-frame LUTs for K6's look-ups: sky view as 2-D, aerial perspective as a 32 x (32 D) atlas of its slices.  They sit
    // over RGBA16F copies that K3 / K4 write beside the RGBA32F LUTs: a LUT-only composite (scenes c1 / c2) is bound by the L1/TEX pipe on 128-bit
    // texels (ncu profiles/k6c2_r02v.md: 78.6 % of its peak), 64-bit texels filter at twice the rate; K6's output is RGBA16F anyway
    Lut<half4> sky_lum_h, sky_trans_h, ap_lum_h, ap_trans_h;
    cudaTextureObject_t sky_lum_tex = 0, sky_trans_tex = 0, ap_lum_tex = 0, ap_trans_tex = 0;
    const void* lut_tex_key[4] = {nullptr, nullptr, nullptr, nullptr};
    int lut_tex_dims[4][3] = {};
    struct FroxelTex { cudaTextureObject_t tex = 0; const void* key = nullptr; int w = 0, h = 0, d = 0; };
    FroxelTex froxel_tex[2];         // LINEAR R16 views of the (double-buffered) froxel volume for K6, see froxel_texture (atmosphere.cu)
    uint16_t* blue_noise = nullptr;  // 64x64 u16

    // materials
    MipTextureDev cloud_map, detail, displacement, voxel;

    // shadow chain
    Lut<float2> shadow_maps[3];
    Lut<uint16_t> shadow_froxel;
    Lut<uchar4> star_map;        // GL_SRGB8 star map as RGBX codes (sky_set_star_map); p == nullptr: no star term
    float* srgb_decode = nullptr; // 256-entry sRGB -> linear table (device)
    // earth albedo map (sky_set_earth_albedo; earth.cu): GL_SRGB8 codes as RGBX, every mip level, level l at earth_albedo + earth_off[l]
    uchar4* earth_albedo = nullptr;
    int earth_w = 0, earth_h = 0, earth_levels = 0;
    unsigned long long earth_off[kEarthMaxLevels] = {};
    size_t earth_texels = 0;
    Lut<float> mesh_shadow_map;  // SKY_RES_MESH_SHADOW_MAP: 2048^2 light-space depth, allocated on first use, cleared to 1

    // viewport
    int width = 0, height = 0;
    Lut<float> checkerboard_depth, cloud_distance;
    Lut<float2> index_linear_depth;
    Lut<half4> render_texture, reconstruct[2];

    // K16 wavefront: per-ray records between k16_setup / k16_march / k16_resolve, and the ray queue
    void* ray_setup = nullptr;
    void* ray_raw = nullptr;
    size_t ray_records_bytes = 0;
    unsigned int* ray_job_counter = nullptr;

    // path tracer
    Lut<float4> pt_accum;
    Lut<uint8_t> pt_mask;
    void* pt_samples = nullptr;           // per-job sample slots of K19, float4[frames][region pixels]
    size_t pt_samples_bytes = 0;
    unsigned int* pt_job_counter = nullptr;
    int pt_tracking = 0;                   // SkyPtTracking (sky_pt_set_tracking)
    uint8_t* voxel_majorant = nullptr;     // majorant-grid mode: max density code per 8^3-texel macro cell of the voxel texture
    size_t voxel_majorant_cells = 0;
    bool voxel_majorant_valid = false;

    unsigned long long* counters = nullptr;  // SkyCounter slots

    // peer-memory exchange of the K16 outputs (sky_peer_attach)
    int peer_rank = 0, peer_world = 1;
    half4* peer_render[8] = {};      // [rank] -> that rank's render_texture (own entry = local pointer)
    float* peer_distance[8] = {};
    unsigned int* peer_flags[8] = {};  // [rank] -> that rank's arrival flags
    unsigned int* my_flags = nullptr;  // unsigned int[SKY_PEER_FLAG_SLOTS]: [0,8) arrival epochs, [8,16) done epochs (written by peers), [16] timeout,
                                       // [32,40) frame-target rows arrived, [40,48) frame target of the previous frame released
    Lut<half4> frame_hdr;              // SKY_RES_FRAME_HDR: the exported frame target (sky_set_output_gather)
    half4* peer_hdr[8] = {};           // [rank] -> that rank's frame_hdr (own entry = local pointer)
    int out_gather = 0;                // SkyOutputGather
    unsigned int peer_epoch = 0;
    bool peer_band_frame = false;      // the open frame rendered bands into peer memory

    // staging for *_host entry points
    float* stage_depth = nullptr;
    half4* stage_hdr = nullptr;
    size_t stage_pixels = 0;
};

// NVTX ranges named like the reference's PERF_MARKER debug groups (src/Base/include/PerformanceMarker.h:8-18; call sites
// Atmosphere.cpp:102-116, AtmosphereRenderer.cpp:222-247, VolumetricCloud.cpp:283-316,333-407, IBL.cpp:29-35): a timeline tool
// shows the same pass names the reference shows in a GL debugger.  Header-only NVTX 3: a no-op unless a tool is attached.
struct SkyPerfMarker {
    explicit SkyPerfMarker(const char* message) { nvtxRangePushA(message); }
    ~SkyPerfMarker() { nvtxRangePop(); }
    SkyPerfMarker(const SkyPerfMarker&) = delete;
    SkyPerfMarker& operator=(const SkyPerfMarker&) = delete;
};
#define SKY_PERF_CONCAT_INNER(a, b) a##b
#define SKY_PERF_CONCAT(a, b) SKY_PERF_CONCAT_INNER(a, b)
#define SKY_PERF_MARKER(message) SkyPerfMarker SKY_PERF_CONCAT(sky_perf_marker_, __LINE__)(message)

// error plumbing ------------------------------------------------------------------------------------------
int sky_fail(SkyContext* ctx, const std::string& msg);
#define SKY_CUDA(ctx, expr)                                                                         \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return sky_fail(ctx, std::string(#expr) + ": " + cudaGetErrorString(e__));               \
    } while (0)
#define SKY_LAUNCH_CHECK(ctx)                 \
    do {                                      \
        ++(ctx)->launch_count;                \
        SKY_CUDA(ctx, cudaGetLastError());    \
    } while (0)

template <class T>
int sky_alloc(SkyContext* ctx, Lut<T>& l, int w, int h, int d = 1, bool zero = true) {
    if (l.p && l.w == w && l.h == h && l.d == d) {
        if (zero) SKY_CUDA(ctx, cudaMemsetAsync(l.p, 0, l.bytes(), ctx->stream));
        return 0;
    }
    if (l.p) SKY_CUDA(ctx, cudaFree(l.p));
    l.p = nullptr;
    l.w = w; l.h = h; l.d = d;
    if (l.bytes() == 0) return 0;
    SKY_CUDA(ctx, cudaMalloc(&l.p, l.bytes()));
    if (zero) SKY_CUDA(ctx, cudaMemsetAsync(l.p, 0, l.bytes(), ctx->stream));
    return 0;
}

// per-subsystem launchers (defined in the .cu files) -----------------------------------------------------------
int ensure_mesh_shadow_map(SkyContext* ctx);                                   // api.cu
// rows of a height-h image owned under sky_set_output_bands, and the n-th owned row
inline int owned_rows(const SkyContext* ctx, int h) {
    if (ctx->out_band_count <= 1) return h;
    int n = 0;
    for (int r0 = ctx->out_band_index * ctx->out_band_rows; r0 < h; r0 += ctx->out_band_rows * ctx->out_band_count) n += std::min(ctx->out_band_rows, h - r0);
    return n;
}
int launch_lut_half_copies(SkyContext* ctx);                                  // atmosphere.cu
int launch_frame_lut_half_copies(SkyContext* ctx);                            // atmosphere.cu
int launch_atmosphere_bake(SkyContext* ctx);                                   // atmosphere.cu  K1,K2
int launch_atmosphere_luts(SkyContext* ctx);                                   // atmosphere.cu  K3,K4,K5
int launch_composite(SkyContext* ctx, const float* depth, half4* hdr, int w, int h);  // atmosphere.cu K6
int launch_env_brdf_lut(SkyContext* ctx);                                      // ibl.cu         K22
int launch_ibl_precompute(SkyContext* ctx);                                    // ibl.cu         cube mips, K23, K24
int launch_earth_albedo_mips(SkyContext* ctx, const float* thresholds_dev);    // earth.cu       glGenerateTextureMipmap (GL_SRGB8)
int launch_gbuffer_clear(SkyContext* ctx, float* depth, void* albedo, void* normal, void* orm, int width, int height);  // earth.cu
int launch_earth_gbuffer(SkyContext* ctx, const SkyEarthBufferData& e, float* depth, void* albedo, void* normal, void* orm, int width, int height);  // earth.cu K7
int launch_tonemap(SkyContext* ctx, const half4* hdr, int w, int h, const SkyToneMapParams& p, void* out);  // atmosphere.cu K21
int launch_noise(SkyContext* ctx, int kind, const SkyNoiseCreateInfo* info);   // noise.cu       K8-K10
int build_mip_texture(SkyContext* ctx, MipTextureDev& t, int w, int h, int d, int channels, bool border);
int launch_mip_chain(SkyContext* ctx, MipTextureDev& t);                       // noise.cu       glGenerateTextureMipmap
int launch_cloud_shadow(SkyContext* ctx, const SkyCloudCommonBufferData& c);   // cloud.cu       K11-K13
int launch_cloud_begin(SkyContext* ctx, const SkyCloudCommonBufferData& c, const SkyCloudBufferData& b, const float* depth,
                       int band_rows, int band_index, int band_count);         // cloud.cu       K14-K16
int launch_cloud_end(SkyContext* ctx, const SkyCloudCommonBufferData& c, const float* depth, half4* hdr, int phases = 3);  // K17 (1), K18 (2)
int launch_pt_samples(SkyContext* ctx, const SkyCloudCommonBufferData& c, uint32_t frame_begin, uint32_t count,
                      const int32_t region[4]);                                // pathtrace.cu   K19
int launch_pt_resolve(SkyContext* ctx, uint32_t frame_count, half4* hdr);      // pathtrace.cu   K20
int launch_tex_peak(SkyContext* ctx, int mode, double* fetches_per_second);    // cloud.cu
// the same entry points of the strict objects (cloud_strict.o, pathtrace_strict.o, composite_strict.o)
#ifndef SKY_STRICT_TU  // (inside those objects the plain names above ARE these, by macro)
int launch_composite_strict(SkyContext* ctx, const float* depth, half4* hdr, int w, int h);
int launch_tonemap_strict(SkyContext* ctx, const half4* hdr, int w, int h, const SkyToneMapParams& p, void* out);
int launch_cloud_shadow_strict(SkyContext* ctx, const SkyCloudCommonBufferData& c);
int launch_cloud_begin_strict(SkyContext* ctx, const SkyCloudCommonBufferData& c, const SkyCloudBufferData& b, const float* depth,
                              int band_rows, int band_index, int band_count);
int launch_cloud_end_strict(SkyContext* ctx, const SkyCloudCommonBufferData& c, const float* depth, half4* hdr, int phases);
int launch_pt_samples_strict(SkyContext* ctx, const SkyCloudCommonBufferData& c, uint32_t frame_begin, uint32_t count, const int32_t region[4]);
int launch_pt_resolve_strict(SkyContext* ctx, uint32_t frame_count, half4* hdr);
#endif
