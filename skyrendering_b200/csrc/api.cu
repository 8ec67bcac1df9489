// C ABI of libskyb200.so (include/skyb200.h).  Plain pointers and PODs only; no exceptions and no
// torch types cross this boundary.  There is no CPU fallback anywhere in this library: without a
// CUDA device sky_ctx_create fails.
#include <cmath>
#include <vector>
#include "../../include/skyb200.h"

#include <cstdlib>
#include <cstring>

#include "context.h"

int sky_fail(SkyContext* ctx, const std::string& msg) {
    if (ctx) ctx->error = msg;
    return 1;
}

namespace {

thread_local std::string g_create_error;

template <class T>
void free_lut(Lut<T>& l) {
    if (l.p) cudaFree(l.p);
    l = Lut<T>{};
}
void free_mip(MipTextureDev& t) {
    if (t.view.tex_linear) cudaDestroyTextureObject(t.view.tex_linear);
    if (t.view.tex_point) cudaDestroyTextureObject(t.view.tex_point);
    if (t.array) cudaFreeMipmappedArray(t.array);
    if (t.data) cudaFree(t.data);
    if (t.cells) cudaFree(t.cells);
    t = MipTextureDev{};
}

struct ResView {
    void* ptr = nullptr;
    int w = 0, h = 0, d = 1, ch = 1, fmt = SKY_FMT_F32;
    size_t bytes = 0;
};

template <class T>
ResView view_of(const Lut<T>& l, int ch, int fmt) {
    ResView v;
    v.ptr = l.p; v.w = l.w; v.h = l.h; v.d = l.d; v.ch = ch; v.fmt = fmt; v.bytes = l.bytes();
    return v;
}
ResView mip0_of(const MipTextureDev& t) {
    ResView v;
    if (!t.valid) return v;
    v.ptr = t.data; v.w = t.view.w[0]; v.h = t.view.h[0]; v.d = t.view.d[0]; v.ch = t.view.channels; v.fmt = SKY_FMT_U8;
    v.bytes = size_t(v.w) * v.h * v.d * v.ch;
    return v;
}
ResView mips_of(const MipTextureDev& t) {
    ResView v;
    if (!t.valid) return v;
    size_t first = size_t(t.view.off[1]) * t.view.channels;
    v.ptr = t.data + first; v.bytes = t.bytes - first; v.w = int(v.bytes); v.h = 1; v.d = 1; v.ch = 1; v.fmt = SKY_FMT_U8;
    return v;
}

bool resolve(SkyContext* ctx, int resource, ResView& v) {
    switch (resource) {
        case SKY_RES_TRANSMITTANCE: v = view_of(ctx->transmittance, 4, SKY_FMT_F32); return true;
        case SKY_RES_MULTISCATTERING: v = view_of(ctx->multiscattering, 4, SKY_FMT_F32); return true;
        case SKY_RES_SKY_VIEW_LUMINANCE: v = view_of(ctx->sky_lum, 4, SKY_FMT_F32); return true;
        case SKY_RES_SKY_VIEW_TRANSMITTANCE: v = view_of(ctx->sky_trans, 4, SKY_FMT_F32); return true;
        case SKY_RES_AERIAL_LUMINANCE: v = view_of(ctx->ap_lum, 4, SKY_FMT_F32); return true;
        case SKY_RES_AERIAL_TRANSMITTANCE: v = view_of(ctx->ap_trans, 4, SKY_FMT_F32); return true;
        case SKY_RES_ENVIRONMENT: v = view_of(ctx->env, 4, SKY_FMT_F16); return true;
        case SKY_RES_ENV_BRDF_LUT: v = view_of(ctx->env_brdf_lut, 2, SKY_FMT_U16); return true;
        case SKY_RES_ENV_RADIANCE_SH: if (ctx->ibl_valid) v = view_of(ctx->env_sh, 4, SKY_FMT_F32); return true;
        case SKY_RES_ENVIRONMENT_MIPS: case SKY_RES_PREFILTERED_RADIANCE: {  // flat, like the *_MIPS resources
            if (!ctx->ibl_valid) return true;
            const bool mips = resource == SKY_RES_ENVIRONMENT_MIPS;
            v.ptr = mips ? ctx->env_mips : ctx->prefiltered;
            v.w = int((mips ? ctx->env_mips_texels : ctx->prefiltered_texels) * 4); v.h = 1; v.d = 1; v.ch = 1; v.fmt = SKY_FMT_F16;
            v.bytes = size_t(v.w) * 2;
            return true;
        }
        case SKY_RES_CLOUD_MAP: v = mip0_of(ctx->cloud_map); return true;
        case SKY_RES_DETAIL: v = mip0_of(ctx->detail); return true;
        case SKY_RES_DISPLACEMENT: v = mip0_of(ctx->displacement); return true;
        case SKY_RES_VOXEL: v = mip0_of(ctx->voxel); return true;
        case SKY_RES_CLOUD_MAP_MIPS: v = mips_of(ctx->cloud_map); return true;
        case SKY_RES_DETAIL_MIPS: v = mips_of(ctx->detail); return true;
        case SKY_RES_DISPLACEMENT_MIPS: v = mips_of(ctx->displacement); return true;
        case SKY_RES_VOXEL_MIPS: v = mips_of(ctx->voxel); return true;
        case SKY_RES_SHADOW_MAP_RAW: v = view_of(ctx->shadow_maps[0], 2, SKY_FMT_F32); return true;
        case SKY_RES_SHADOW_MAP: v = view_of(ctx->shadow_maps[2], 2, SKY_FMT_F32); return true;
        case SKY_RES_SHADOW_FROXEL: v = view_of(ctx->shadow_froxel, 1, SKY_FMT_U16); return true;
        case SKY_RES_MESH_SHADOW_MAP: v = view_of(ctx->mesh_shadow_map, 1, SKY_FMT_F32); return true;
        case SKY_RES_CHECKERBOARD_DEPTH: v = view_of(ctx->checkerboard_depth, 1, SKY_FMT_F32); return true;
        case SKY_RES_INDEX_LINEAR_DEPTH: v = view_of(ctx->index_linear_depth, 2, SKY_FMT_F32); return true;
        case SKY_RES_CLOUD_RENDER: v = view_of(ctx->render_texture, 4, SKY_FMT_F16); return true;
        case SKY_RES_CLOUD_DISTANCE: v = view_of(ctx->cloud_distance, 1, SKY_FMT_F32); return true;
        case SKY_RES_RECONSTRUCT: v = view_of(ctx->reconstruct[1], 4, SKY_FMT_F16); return true;  // newest after the swap
        case SKY_RES_PT_ACCUM: v = view_of(ctx->pt_accum, 4, SKY_FMT_F32); return true;
        case SKY_RES_PT_MASK: v = view_of(ctx->pt_mask, 1, SKY_FMT_U8); return true;
        case SKY_RES_FRAME_HDR: v = view_of(ctx->frame_hdr, 4, SKY_FMT_F16); return true;
        case SKY_RES_EARTH_ALBEDO:
            if (!ctx->earth_albedo) return true;
            v.ptr = ctx->earth_albedo; v.w = int(ctx->earth_texels); v.h = 1; v.d = 1; v.ch = 4; v.fmt = SKY_FMT_U8; v.bytes = ctx->earth_texels * 4;
            return true;
        case SKY_RES_COUNTERS:
            v.ptr = ctx->counters; v.w = 8; v.h = 1; v.d = 1; v.ch = 1; v.fmt = SKY_FMT_U64; v.bytes = 64;
            return true;
    }
    return false;
}

int ensure_stage(SkyContext* ctx) {
    size_t px = size_t(ctx->width) * ctx->height;
    if (ctx->stage_pixels == px && ctx->stage_depth) return 0;
    if (ctx->stage_depth) cudaFree(ctx->stage_depth);
    if (ctx->stage_hdr) cudaFree(ctx->stage_hdr);
    ctx->stage_depth = nullptr; ctx->stage_hdr = nullptr; ctx->stage_pixels = 0;
    SKY_CUDA(ctx, cudaMalloc(&ctx->stage_depth, px * sizeof(float)));
    SKY_CUDA(ctx, cudaMalloc(&ctx->stage_hdr, px * sizeof(half4)));
    ctx->stage_pixels = px;
    return 0;
}

}  // namespace

extern "C" {

// The two bake LUTs (Atmosphere.cpp:11-18) of the CURRENT set + the RGBA16F copies K6 fetches through the texture unit
static int alloc_bake_luts(SkyContext* ctx) {
    int rc = 0;
    rc |= sky_alloc(ctx, ctx->transmittance, 256, 64);   // Atmosphere.cpp:11-12
    rc |= sky_alloc(ctx, ctx->multiscattering, 32, 32);  // Atmosphere.cpp:17-18
    rc |= sky_alloc(ctx, ctx->transmittance_h, 256, 64);
    rc |= sky_alloc(ctx, ctx->multiscattering_h, 32, 32);
    rc |= sky_alloc(ctx, ctx->density_h, kDensityLutSize, 1);
    if (rc) return rc;
    // GL_LINEAR + CLAMP_TO_EDGE texture views (Samplers.cpp linear_clamp_no_mipmap) over the RGBA16F copies
    auto make_tex = [](const Lut<half4>& l, cudaTextureObject_t* out) {
        cudaResourceDesc res{};
        res.resType = cudaResourceTypePitch2D;
        res.res.pitch2D.devPtr = l.p;
        res.res.pitch2D.desc = cudaCreateChannelDescHalf4();
        res.res.pitch2D.width = size_t(l.w);
        res.res.pitch2D.height = size_t(l.h);
        res.res.pitch2D.pitchInBytes = size_t(l.w) * sizeof(half4);
        cudaTextureDesc td{};
        td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 1;
        return cudaCreateTextureObject(out, &res, &td, nullptr) == cudaSuccess ? 0 : 1;
    };
    rc |= make_tex(ctx->transmittance_h, &ctx->transmittance_tex);
    rc |= make_tex(ctx->multiscattering_h, &ctx->multiscattering_tex);
    rc |= make_tex(ctx->density_h, &ctx->density_tex);
    return rc;
}

// current LUT set <-> alternate LUT set (frame pipelining)
static void swap_lut_sets(SkyContext* ctx) {
    SkyContext::LutSet& a = ctx->alt;
    std::swap(ctx->transmittance, a.transmittance); std::swap(ctx->multiscattering, a.multiscattering);
    std::swap(ctx->sky_lum, a.sky_lum); std::swap(ctx->sky_trans, a.sky_trans);
    std::swap(ctx->sky_lum_h, a.sky_lum_h); std::swap(ctx->sky_trans_h, a.sky_trans_h); std::swap(ctx->ap_lum_h, a.ap_lum_h); std::swap(ctx->ap_trans_h, a.ap_trans_h);
    std::swap(ctx->ap_lum, a.ap_lum); std::swap(ctx->ap_trans, a.ap_trans);
    std::swap(ctx->env, a.env); std::swap(ctx->transmittance_h, a.transmittance_h); std::swap(ctx->multiscattering_h, a.multiscattering_h);
    std::swap(ctx->density_h, a.density_h); std::swap(ctx->density_tex, a.density_tex);
    std::swap(ctx->transmittance_tex, a.transmittance_tex); std::swap(ctx->multiscattering_tex, a.multiscattering_tex);
    std::swap(ctx->sky_lum_tex, a.sky_lum_tex); std::swap(ctx->sky_trans_tex, a.sky_trans_tex);
    std::swap(ctx->ap_lum_tex, a.ap_lum_tex); std::swap(ctx->ap_trans_tex, a.ap_trans_tex);
    for (int i = 0; i < 4; ++i) {
        std::swap(ctx->lut_tex_key[i], a.lut_tex_key[i]);
        for (int k = 0; k < 3; ++k) std::swap(ctx->lut_tex_dims[i][k], a.lut_tex_dims[i][k]);
    }
    std::swap(ctx->env_mips, a.env_mips); std::swap(ctx->env_mips_for, a.env_mips_for); std::swap(ctx->env_mips_texels, a.env_mips_texels);
    std::swap(ctx->env_sh, a.env_sh); std::swap(ctx->prefiltered, a.prefiltered); std::swap(ctx->prefiltered_texels, a.prefiltered_texels);
    std::swap(ctx->ibl_valid, a.ibl_valid);
}

int sky_ctx_create(int device, void* cuda_stream, SkyContext** out) {
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || device < 0 || device >= count) {
        g_create_error = std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
        return 1;
    }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        g_create_error = std::string("cudaSetDevice: ") + cudaGetErrorString(e);
        return 1;
    }
    SkyContext* ctx = new SkyContext();
    ctx->device = device;
    ctx->stream = static_cast<cudaStream_t>(cuda_stream);
    if (const char* lit = std::getenv("SKYB200_K16_LITERAL")) ctx->k16_literal = lit[0] == '1';
    if (const char* grp = std::getenv("SKYB200_K16_GROUP")) ctx->k16_group = grp[0] == '4' ? 4 : grp[0] == '8' ? 8 : 0;   // A/B switch for measurements (tools/)
    int rc = 0;
    rc |= alloc_bake_luts(ctx);
    for (auto& m : ctx->shadow_maps) rc |= sky_alloc(ctx, m, 512, 512);  // VolumetricCloud.cpp:52,102-105
    if (cudaMalloc(&ctx->blue_noise, 64 * 64 * sizeof(uint16_t)) != cudaSuccess) rc = 1;
    if (cudaMalloc(&ctx->counters, 8 * sizeof(unsigned long long)) != cudaSuccess) rc = 1;
    if (!rc) {
        cudaMemsetAsync(ctx->blue_noise, 0, 64 * 64 * sizeof(uint16_t), ctx->stream);
        cudaMemsetAsync(ctx->counters, 0, 8 * sizeof(unsigned long long), ctx->stream);
    }
    if (rc) {
        g_create_error = ctx->error.empty() ? "device allocation failed" : ctx->error;
        sky_ctx_destroy(ctx);  // frees whatever was allocated before the failure
        return 1;
    }
    *out = ctx;
    return 0;
}

void sky_ctx_destroy(SkyContext* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    free_lut(ctx->transmittance_h); free_lut(ctx->multiscattering_h); free_lut(ctx->density_h);
    free_lut(ctx->transmittance); free_lut(ctx->multiscattering); free_lut(ctx->sky_lum); free_lut(ctx->sky_trans);
    free_lut(ctx->ap_lum); free_lut(ctx->ap_trans); free_lut(ctx->env);
    for (auto& m : ctx->shadow_maps) free_lut(m);
    free_lut(ctx->star_map); if (ctx->srgb_decode) cudaFree(ctx->srgb_decode);
    if (ctx->earth_albedo) cudaFree(ctx->earth_albedo);
    for (auto& e : ctx->froxel_tex) if (e.tex) cudaDestroyTextureObject(e.tex);
    free_lut(ctx->mesh_shadow_map); free_lut(ctx->shadow_froxel); free_lut(ctx->checkerboard_depth); free_lut(ctx->cloud_distance);
    free_lut(ctx->index_linear_depth); free_lut(ctx->render_texture); free_lut(ctx->reconstruct[0]); free_lut(ctx->reconstruct[1]);
    free_lut(ctx->pt_accum); free_lut(ctx->pt_mask);
    free_mip(ctx->cloud_map); free_mip(ctx->detail); free_mip(ctx->displacement); free_mip(ctx->voxel);
    if (ctx->blue_noise) cudaFree(ctx->blue_noise);
    free_lut(ctx->env_brdf_lut); free_lut(ctx->env_sh);
    if (ctx->env_mips) cudaFree(ctx->env_mips);
    if (ctx->prefiltered) cudaFree(ctx->prefiltered);
    free_lut(ctx->alt.env_sh);
    if (ctx->alt.env_mips) cudaFree(ctx->alt.env_mips);
    if (ctx->alt.prefiltered) cudaFree(ctx->alt.prefiltered);
    if (ctx->lane2) { cudaStreamSynchronize(ctx->lane2); cudaStreamDestroy(ctx->lane2); }
    for (cudaEvent_t ev : {ctx->ev_fork, ctx->ev_shadow, ctx->ev_pre_composite, ctx->ev_lane2, ctx->ev_frame_mark[0], ctx->ev_frame_mark[1], ctx->ev_luts_ready, ctx->ev_main_to_lut}) if (ev) cudaEventDestroy(ev);
    free_lut(ctx->alt.shadow_froxel);
    if (ctx->lut_stream) { cudaStreamSynchronize(ctx->lut_stream); cudaStreamDestroy(ctx->lut_stream); }
    if (ctx->shadow_stream) { cudaStreamSynchronize(ctx->shadow_stream); cudaStreamDestroy(ctx->shadow_stream); }
    if (ctx->ev_shadow_ready) cudaEventDestroy(ctx->ev_shadow_ready);
    swap_lut_sets(ctx);  // free the alternate set through the same path
    free_lut(ctx->transmittance_h); free_lut(ctx->multiscattering_h); free_lut(ctx->density_h); free_lut(ctx->transmittance); free_lut(ctx->multiscattering);
    free_lut(ctx->sky_lum); free_lut(ctx->sky_trans); free_lut(ctx->ap_lum); free_lut(ctx->ap_trans); free_lut(ctx->env);
    free_lut(ctx->sky_lum_h); free_lut(ctx->sky_trans_h); free_lut(ctx->ap_lum_h); free_lut(ctx->ap_trans_h);
    for (cudaTextureObject_t t : {ctx->density_tex, ctx->transmittance_tex, ctx->multiscattering_tex, ctx->sky_lum_tex, ctx->sky_trans_tex, ctx->ap_lum_tex, ctx->ap_trans_tex}) if (t) cudaDestroyTextureObject(t);
    swap_lut_sets(ctx);
    if (ctx->transmittance_tex) cudaDestroyTextureObject(ctx->transmittance_tex);
    if (ctx->multiscattering_tex) cudaDestroyTextureObject(ctx->multiscattering_tex);
    for (cudaTextureObject_t t : {ctx->sky_lum_tex, ctx->sky_trans_tex, ctx->ap_lum_tex, ctx->ap_trans_tex}) if (t) cudaDestroyTextureObject(t);
    if (ctx->counters) cudaFree(ctx->counters);
    if (ctx->ray_setup) cudaFree(ctx->ray_setup);
    if (ctx->ray_raw) cudaFree(ctx->ray_raw);
    if (ctx->ray_job_counter) cudaFree(ctx->ray_job_counter);
    if (ctx->stage_depth) cudaFree(ctx->stage_depth);
    if (ctx->stage_hdr) cudaFree(ctx->stage_hdr);
    if (ctx->pt_samples) cudaFree(ctx->pt_samples);
    if (ctx->pt_job_counter) cudaFree(ctx->pt_job_counter);
    if (ctx->voxel_majorant) cudaFree(ctx->voxel_majorant);
    free_lut(ctx->alt.shadow_blurred);
    sky_peer_detach(ctx);
    if (ctx->my_flags) cudaFree(ctx->my_flags);
    free_lut(ctx->frame_hdr);
    delete ctx;
}

// ---- frame overlap ------------------------------------------------------------------------------------------
// A frame has two independent halves until the upscale: {K11-K13 shadow chain, K14-K17 cloud chain} only meet
// {K1-K5 LUTs, K6 composite} at the froxels (K6 reads them), the LUTs (K16 reads them) and the HDR target (K18).  With
// overlap enabled the first half runs on `lane2`, ordered against the caller's stream with events; nothing waits on
// the host.  Every entry point outside this protocol joins the lanes first.
namespace {
struct LaneScope {  // launchers issue on ctx->stream: point it at lane2 for the scope
    SkyContext* ctx; cudaStream_t saved;
    LaneScope(SkyContext* c, cudaStream_t s) : ctx(c), saved(c->stream) { c->stream = s; }
    ~LaneScope() { ctx->stream = saved; }
};
int luts_join(SkyContext* ctx) {  // the caller's stream is ordered after everything queued on lut_stream (frame pipelining)
    if (ctx->shadow_stream_pending) {
        SKY_CUDA(ctx, cudaEventRecord(ctx->ev_shadow_ready, ctx->shadow_stream));
        SKY_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_shadow_ready, 0));
        ctx->shadow_stream_pending = false;
    }
    if (!ctx->luts_pending) return 0;
    SKY_CUDA(ctx, cudaEventRecord(ctx->ev_luts_ready, ctx->lut_stream));
    SKY_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_luts_ready, 0));
    ctx->luts_pending = false;
    return 0;
}
int lut_stream_follow_main(SkyContext* ctx) {  // lut_stream is ordered after everything queued on the caller's stream so far
    if (!ctx->pipelining) return 0;
    SKY_CUDA(ctx, cudaEventRecord(ctx->ev_main_to_lut, ctx->stream));
    SKY_CUDA(ctx, cudaStreamWaitEvent(ctx->lut_stream, ctx->ev_main_to_lut, 0));
    if (ctx->shadow_stream) SKY_CUDA(ctx, cudaStreamWaitEvent(ctx->shadow_stream, ctx->ev_main_to_lut, 0));
    return 0;
}
int lanes_join(SkyContext* ctx) {  // the caller's stream is ordered after everything queued on lane2 (and on lut_stream)
    if (int e = luts_join(ctx)) return e;
    if (!ctx->lane2_pending) return 0;
    SKY_CUDA(ctx, cudaEventRecord(ctx->ev_lane2, ctx->lane2));
    SKY_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_lane2, 0));
    ctx->lane2_pending = ctx->shadow_pending = ctx->pre_composite_recorded = ctx->lane2_reads_luts = false;
    return 0;
}
int lane2_fork(SkyContext* ctx, cudaEvent_t after) {  // lane2 is ordered after `after` (recorded on the caller's stream)
    SKY_CUDA(ctx, cudaStreamWaitEvent(ctx->lane2, after, 0));
    ctx->lane2_pending = true;
    return 0;
}
}  // namespace

int sky_set_frame_overlap(SkyContext* ctx, int enable) {
    if (int e = lanes_join(ctx)) return e;
    if (enable && !ctx->lane2) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        // lane2 (shadow chain, K14-K17) runs ABOVE the caller's stream: its kernels are the short ones of a frame, and behind K6's 32 400 queued
        // blocks they only got SM slots as those drained.  Measured at 4K, scene c3 (profiles/k16_persist_r02F.log): 1197 -> 1165 us per frame.
        // (Same log: a persistent K16 holding a fixed 3-6 blocks of every SM beside K6 -- the two instruction streams interleaved instead of
        // one kernel after the other -- is bit-identical and no faster, 1158 us: K6 already fills 89 % of the issue slots.  Removed again.)
        const char* pr = std::getenv("SKYB200_LANE2_PRIORITY");   // 0: the caller's priority (A/B measurements)
        SKY_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->lane2, cudaStreamNonBlocking, pr && pr[0] == '0' ? lo : hi));
        for (cudaEvent_t* ev : {&ctx->ev_fork, &ctx->ev_shadow, &ctx->ev_pre_composite, &ctx->ev_lane2})
            SKY_CUDA(ctx, cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
    }
    ctx->overlap = enable != 0;
    return 0;
}

// ---- frame pipelining -------------------------------------------------------------------------------------------
// Per frame the atmosphere LUT phase (K1, K2, K3+K4, K5: ~0.35 ms of small, latency-bound kernels) must finish before either
// full-machine kernel of the frame (K6, K16) can start, and nothing else is runnable meanwhile.  With pipelining enabled the
// context owns TWO LUT sets: sky_atmosphere_bake flips to the set the frame before last used and issues the LUT phase on
// `lut_stream`, where it runs beside the PREVIOUS frame's K6 / K16 (which read the other set); the caller's stream waits for
// it only where the LUTs are first read.  Same kernels, same inputs, same results (the sets are independent copies).
int sky_set_frame_pipelining(SkyContext* ctx, int enable) {
    if (int e = lanes_join(ctx)) return e;
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (enable && !ctx->lut_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        SKY_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->lut_stream, cudaStreamNonBlocking, hi));
        SKY_CUDA(ctx, cudaStreamCreateWithPriority(&ctx->shadow_stream, cudaStreamNonBlocking, hi));
        for (cudaEvent_t* ev : {&ctx->ev_frame_mark[0], &ctx->ev_frame_mark[1], &ctx->ev_luts_ready, &ctx->ev_main_to_lut, &ctx->ev_shadow_ready})
            SKY_CUDA(ctx, cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
        swap_lut_sets(ctx);                       // allocate the bake LUTs of the second set
        int rc = alloc_bake_luts(ctx);
        swap_lut_sets(ctx);
        if (rc) return sky_fail(ctx, "device allocation failed");
        SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    ctx->pipelining = enable != 0;
    ctx->mark_count = 0;
    return 0;
}

const char* sky_last_error(SkyContext* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

// a device-side wait of the tile-sharded frame's peer barrier gave up (k_peer_flags): report it once the stream has drained
static int check_peer_timeout(SkyContext* ctx) {
    if (!ctx->my_flags || ctx->peer_world <= 1) return 0;
    unsigned int flag = 0;
    SKY_CUDA(ctx, cudaMemcpy(&flag, ctx->my_flags + SKY_PEER_TIMEOUT_SLOT, sizeof(flag), cudaMemcpyDeviceToHost));
    if (!flag) return 0;
    SKY_CUDA(ctx, cudaMemset(ctx->my_flags + SKY_PEER_TIMEOUT_SLOT, 0, sizeof(flag)));
    return sky_fail(ctx, "tile-sharded frame: timed out waiting for the rows of rank " + std::to_string(flag - 1) +
                             " (a rank skipped a frame, died, or attached out of step)");
}

int sky_sync(SkyContext* ctx) {
    if (int e = lanes_join(ctx)) return e;
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return check_peer_timeout(ctx);
}

int sky_set_blue_noise(SkyContext* ctx, const uint16_t* texels) {
    SKY_CUDA(ctx, cudaMemcpyAsync(ctx->blue_noise, texels, 64 * 64 * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream));
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the host buffer may be a temporary
    return 0;
}

// GL 4.6 section 8.24: sRGB -> linear, applied to each texel before filtering
static int ensure_srgb_decode(SkyContext* ctx) {
    if (ctx->srgb_decode) return 0;
    float decode[256];
    for (int c = 0; c < 256; ++c) {
        double cs = c / 255.0;
        decode[c] = float(cs <= 0.04045 ? cs / 12.92 : std::pow((cs + 0.055) / 1.055, 2.4));
    }
    SKY_CUDA(ctx, cudaMalloc(&ctx->srgb_decode, sizeof(decode)));
    SKY_CUDA(ctx, cudaMemcpy(ctx->srgb_decode, decode, sizeof(decode), cudaMemcpyHostToDevice));
    return 0;
}

int sky_set_earth_albedo(SkyContext* ctx, const uint8_t* host_srgb8, int width, int height) {
    if (int e = lanes_join(ctx)) return e;
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->earth_albedo) { cudaFree(ctx->earth_albedo); ctx->earth_albedo = nullptr; }
    ctx->earth_w = ctx->earth_h = ctx->earth_levels = 0; ctx->earth_texels = 0;
    if (width <= 0 || height <= 0 || !host_srgb8) return 0;
    if (width > 16384 || height > 16384) return sky_fail(ctx, "earth albedo map too large");
    if (int e = ensure_srgb_decode(ctx)) return e;
    // sRGB transfer function (GL 4.6 section 17.3.7), rounded to the nearest code: monotonic, so "the code of v" is the number of
    // thresholds <= v.  thr[c] = the smallest fp32 value whose code is >= c, found by bisection on the bit pattern.
    auto code_of = [](float cl) {
        double c = cl;
        double cs = !(c > 0.0) ? 0.0 : c < 0.0031308 ? 12.92 * c : c < 1.0 ? 1.055 * std::pow(c, 0.41666) - 0.055 : 1.0;
        return int(std::floor(cs * 255.0 + 0.5));
    };
    float thr[256];
    thr[0] = -INFINITY;
    for (int c = 1; c < 256; ++c) {
        uint32_t lo = 0u, hi = 0x3f800000u;  // code_of(bits lo) < c <= code_of(bits hi)
        while (hi - lo > 1u) {
            const uint32_t mid = lo + (hi - lo) / 2u;
            float f; std::memcpy(&f, &mid, 4);
            if (code_of(f) >= c) hi = mid; else lo = mid;
        }
        std::memcpy(&thr[c], &hi, 4);
    }
    int levels = 1;
    size_t texels = size_t(width) * height;
    ctx->earth_off[0] = 0;
    for (int w = width, h = height; w > 1 || h > 1;) {
        w = std::max(w / 2, 1); h = std::max(h / 2, 1);
        ctx->earth_off[levels++] = texels;
        texels += size_t(w) * h;
    }
    SKY_CUDA(ctx, cudaMalloc(&ctx->earth_albedo, texels * sizeof(uchar4)));
    ctx->earth_w = width; ctx->earth_h = height; ctx->earth_levels = levels; ctx->earth_texels = texels;
    std::vector<uchar4> rgbx(size_t(width) * height);
    for (size_t i = 0; i < rgbx.size(); ++i) rgbx[i] = make_uchar4(host_srgb8[i * 3], host_srgb8[i * 3 + 1], host_srgb8[i * 3 + 2], 255);
    SKY_CUDA(ctx, cudaMemcpy(ctx->earth_albedo, rgbx.data(), rgbx.size() * sizeof(uchar4), cudaMemcpyHostToDevice));
    float* thr_dev = nullptr;
    SKY_CUDA(ctx, cudaMalloc(&thr_dev, sizeof(thr)));
    SKY_CUDA(ctx, cudaMemcpy(thr_dev, thr, sizeof(thr), cudaMemcpyHostToDevice));
    int rc = launch_earth_albedo_mips(ctx, thr_dev);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(thr_dev);
    return rc;
}

int sky_gbuffer_clear(SkyContext* ctx, float* depth, void* albedo, void* normal, void* orm, int width, int height) {
    if (!depth || !albedo || !normal || !orm || width <= 0 || height <= 0) return sky_fail(ctx, "gbuffer_clear: bad arguments");
    if ((reinterpret_cast<uintptr_t>(depth) & 7) || (reinterpret_cast<uintptr_t>(albedo) & 7) || (reinterpret_cast<uintptr_t>(normal) & 15) || (reinterpret_cast<uintptr_t>(orm) & 15))
        return sky_fail(ctx, "gbuffer_clear: targets must be 16-byte aligned");
    if (int e = lanes_join(ctx)) return e;  // a frame in flight on the second lane may still read the depth plane
    return launch_gbuffer_clear(ctx, depth, albedo, normal, orm, width, height);
}

int sky_earth_gbuffer(SkyContext* ctx, const SkyEarthBufferData* earth, float* depth, void* albedo, void* normal, void* orm, int width, int height) {
    if (!earth || !depth || !albedo || !normal || !orm || width <= 0 || height <= 0) return sky_fail(ctx, "earth_gbuffer: bad arguments");
    if (!ctx->transmittance.p) return sky_fail(ctx, "earth_gbuffer: the atmosphere has not been baked (bottom_radius)");
    if (int e = lanes_join(ctx)) return e;  // the depth plane may still be read by a frame in flight on the second lane
    return launch_earth_gbuffer(ctx, *earth, depth, albedo, normal, orm, width, height);
}

int sky_set_star_map(SkyContext* ctx, const uint8_t* host_srgb8, int width, int height) {
    if (int e = lanes_join(ctx)) return e;
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free_lut(ctx->star_map);
    if (width <= 0 || height <= 0 || !host_srgb8) return 0;
    if (width > 16384 || height > 16384) return sky_fail(ctx, "star map too large");
    if (int e = ensure_srgb_decode(ctx)) return e;
    if (int e = sky_alloc(ctx, ctx->star_map, width, height, 1, false)) return e;
    std::vector<uchar4> rgbx(size_t(width) * height);
    for (size_t i = 0; i < rgbx.size(); ++i) rgbx[i] = make_uchar4(host_srgb8[i * 3], host_srgb8[i * 3 + 1], host_srgb8[i * 3 + 2], 255);
    SKY_CUDA(ctx, cudaMemcpy(ctx->star_map.p, rgbx.data(), rgbx.size() * sizeof(uchar4), cudaMemcpyHostToDevice));
    return 0;
}

int sky_set_viewport(SkyContext* ctx, int w, int h) {
    if (w < 12 || h < 12) return sky_fail(ctx, "viewport too small");
    if (int e = lanes_join(ctx)) return e;
    sky_peer_detach(ctx);  // the exported buffers are about to be reallocated
    free_lut(ctx->frame_hdr);
    ctx->width = w; ctx->height = h;
    int rc = 0;  // VolumetricCloud.cpp:120-136; histories zero-filled
    rc |= sky_alloc(ctx, ctx->checkerboard_depth, w / 2, h / 2);
    rc |= sky_alloc(ctx, ctx->index_linear_depth, w / 4, h / 4);
    rc |= sky_alloc(ctx, ctx->render_texture, w / 4, h / 4);
    rc |= sky_alloc(ctx, ctx->cloud_distance, w / 4, h / 4);
    rc |= sky_alloc(ctx, ctx->reconstruct[0], w / 2, h / 2);
    rc |= sky_alloc(ctx, ctx->reconstruct[1], w / 2, h / 2);
    for (auto& e : ctx->froxel_tex) if (e.tex) { cudaDestroyTextureObject(e.tex); e = SkyContext::FroxelTex{}; }
    rc |= sky_alloc(ctx, ctx->shadow_froxel, w / 12, h / 12, 128);
    for (auto& m : ctx->shadow_maps) rc |= sky_alloc(ctx, m, 512, 512);
    if (ctx->alt.shadow_blurred.p) rc |= sky_alloc(ctx, ctx->alt.shadow_blurred, 512, 512);
    free_lut(ctx->pt_accum);
    free_lut(ctx->pt_mask);
    if (int e = lut_stream_follow_main(ctx)) return e;  // the zero-filled shadow maps are inputs of the shadow chain
    return rc;
}

int sky_atmosphere_bake(SkyContext* ctx, const SkyAtmosphereBufferData* a) {
    if (ctx->pipelining) {
        // everything queued on the caller's stream before this call -- i.e. all of the previous frame, whose lane2 work was
        // joined at its cloud_frame_end -- precedes mark[n]; the set this bake flips to was last read one frame earlier,
        // i.e. before mark[n-1], recorded at the previous bake
        const int n = ctx->mark_count & 1;
        SKY_CUDA(ctx, cudaEventRecord(ctx->ev_frame_mark[n], ctx->stream));
        swap_lut_sets(ctx);
        ctx->frame_gate = ctx->ev_frame_mark[ctx->mark_count >= 1 ? n ^ 1 : n];
        SKY_CUDA(ctx, cudaStreamWaitEvent(ctx->lut_stream, ctx->frame_gate, 0));
        ++ctx->mark_count;
        ctx->atm = *a;
        ctx->luts_pending = true;
        ctx->bake_since_shadow = true;
        LaneScope lane(ctx, ctx->lut_stream);
        return launch_atmosphere_bake(ctx);
    }
    if (ctx->lane2_reads_luts) { if (int e = lanes_join(ctx)) return e; }
    ctx->atm = *a;
    return launch_atmosphere_bake(ctx);
}

int sky_atmosphere_luts(SkyContext* ctx, const SkyAtmosphereRenderBufferData* r, const SkyLutConfig* cfg) {
    if (cfg->sky_view_width < 2 || cfg->sky_view_height < 2 || cfg->aerial_perspective_depth < 2 || cfg->environment_size < 1)
        return sky_fail(ctx, "bad LUT sizes");
    if (!ctx->pipelining && ctx->lane2_reads_luts) { if (int e = lanes_join(ctx)) return e; }  // (the shadow chain on lane2 does not touch the LUTs)
    ctx->render = *r;
    ctx->lut_cfg = *cfg;
    int rc = 0;
    rc |= sky_alloc(ctx, ctx->sky_lum, cfg->sky_view_width, cfg->sky_view_height, 1, false);
    rc |= sky_alloc(ctx, ctx->sky_trans, cfg->sky_view_width, cfg->sky_view_height, 1, false);
    rc |= sky_alloc(ctx, ctx->ap_lum, 32, 32, cfg->aerial_perspective_depth, false);  // AtmosphereRenderer.cpp:19-20
    rc |= sky_alloc(ctx, ctx->ap_trans, 32, 32, cfg->aerial_perspective_depth, false);
    rc |= sky_alloc(ctx, ctx->env, cfg->environment_size, cfg->environment_size, 6, false);
    rc |= sky_alloc(ctx, ctx->sky_lum_h, cfg->sky_view_width, cfg->sky_view_height, 1, false);
    rc |= sky_alloc(ctx, ctx->sky_trans_h, cfg->sky_view_width, cfg->sky_view_height, 1, false);
    rc |= sky_alloc(ctx, ctx->ap_lum_h, 32, 32, cfg->aerial_perspective_depth, false);
    rc |= sky_alloc(ctx, ctx->ap_trans_h, 32, 32, cfg->aerial_perspective_depth, false);
    if (rc) return rc;
    // texture views for K6 (production object): recreated when sky_alloc moved or resized a LUT
    {
        Lut<half4>* luts[4] = {&ctx->sky_lum_h, &ctx->sky_trans_h, &ctx->ap_lum_h, &ctx->ap_trans_h};   // the RGBA16F copies K3 / K4 write (context.h)
        cudaTextureObject_t* tex[4] = {&ctx->sky_lum_tex, &ctx->sky_trans_tex, &ctx->ap_lum_tex, &ctx->ap_trans_tex};
        for (int i = 0; i < 4; ++i) {
            const Lut<half4>& l = *luts[i];
            if (*tex[i] && ctx->lut_tex_key[i] == l.p && ctx->lut_tex_dims[i][0] == l.w && ctx->lut_tex_dims[i][1] == l.h && ctx->lut_tex_dims[i][2] == l.d) continue;
            if (*tex[i]) { cudaDestroyTextureObject(*tex[i]); *tex[i] = 0; }
            cudaResourceDesc res{};
            res.resType = cudaResourceTypePitch2D;
            res.res.pitch2D.devPtr = l.p;
            res.res.pitch2D.desc = cudaCreateChannelDescHalf4();
            res.res.pitch2D.width = size_t(l.w);
            res.res.pitch2D.height = size_t(l.h) * size_t(l.d);   // 3-D LUT: its slices stacked
            res.res.pitch2D.pitchInBytes = size_t(l.w) * sizeof(half4);
            cudaTextureDesc td{};
            td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
            td.filterMode = cudaFilterModeLinear;
            td.readMode = cudaReadModeElementType;
            td.normalizedCoords = 1;
            SKY_CUDA(ctx, cudaCreateTextureObject(tex[i], &res, &td, nullptr));
            ctx->lut_tex_key[i] = l.p; ctx->lut_tex_dims[i][0] = l.w; ctx->lut_tex_dims[i][1] = l.h; ctx->lut_tex_dims[i][2] = l.d;
        }
    }
    if (ctx->pipelining) {  // follows this frame's bake on lut_stream
        ctx->luts_pending = true;
        LaneScope lane(ctx, ctx->lut_stream);
        return launch_atmosphere_luts(ctx);
    }
    return launch_atmosphere_luts(ctx);
}

int sky_env_brdf_lut(SkyContext* ctx) {
    if (int e = sky_alloc(ctx, ctx->env_brdf_lut, SKY_ENV_BRDF_LUT_SIZE, SKY_ENV_BRDF_LUT_SIZE, 1, false)) return e;
    return launch_env_brdf_lut(ctx);
}

int sky_ibl_precompute(SkyContext* ctx) {
    if (!ctx->env.p) return sky_fail(ctx, "ibl_precompute: the environment cube has not been baked (call atmosphere_luts first)");
    const int n = ctx->env.w;
    if (n & (n - 1)) return sky_fail(ctx, "ibl_precompute: the environment size must be a power of two");
    if (n > 2048) return sky_fail(ctx, "ibl_precompute: environment size above 2048");
    if (ctx->env_mips_for != n) {
        if (ctx->env_mips) { SKY_CUDA(ctx, cudaFree(ctx->env_mips)); ctx->env_mips = nullptr; }
        size_t texels = 0;
        for (int w = n >> 1; w >= 1; w >>= 1) texels += size_t(6) * w * w;
        SKY_CUDA(ctx, cudaMalloc(&ctx->env_mips, std::max<size_t>(texels, 1) * sizeof(half4)));
        ctx->env_mips_texels = texels; ctx->env_mips_for = n;
    }
    if (!ctx->prefiltered) {
        size_t texels = 0;
        for (int l = 0, w = SKY_IBL_PREFILTERED_RESOLUTION; l < SKY_IBL_ROUGHNESS_COUNT; ++l, w >>= 1) texels += size_t(6) * w * w;
        SKY_CUDA(ctx, cudaMalloc(&ctx->prefiltered, texels * sizeof(half4)));
        ctx->prefiltered_texels = texels;
    }
    if (int e = sky_alloc(ctx, ctx->env_sh, 9, 1, 1, false)) return e;
    ctx->ibl_valid = true;
    if (ctx->pipelining) {  // follows this frame's K5 on lut_stream, into this frame's set of the double-buffered outputs
        ctx->luts_pending = true;
        LaneScope lane(ctx, ctx->lut_stream);
        return launch_ibl_precompute(ctx);
    }
    return launch_ibl_precompute(ctx);
}

int sky_set_gbuffer(SkyContext* ctx, const void* albedo, const void* normal, const void* orm) {
    if ((albedo != nullptr) != (normal != nullptr) || (albedo != nullptr) != (orm != nullptr)) return sky_fail(ctx, "set_gbuffer: bind all three targets or none");
    ctx->gbuffer_albedo = albedo; ctx->gbuffer_normal = normal; ctx->gbuffer_orm = orm;
    return 0;
}

int sky_composite(SkyContext* ctx, const float* depth, void* hdr, int width, int height) {
    if (!ctx->sky_lum.p) return sky_fail(ctx, "atmosphere LUTs have not been baked");
    if (ctx->gbuffer_albedo && (!ctx->env_brdf_lut.p || !ctx->ibl_valid))
        return sky_fail(ctx, "composite: a G-buffer is bound but env_brdf_lut / ibl_precompute have not run");
    if (int e = luts_join(ctx)) return e;
    if (ctx->overlap) {
        if (ctx->lane2_reads_luts) { if (int e = lanes_join(ctx)) return e; }  // out-of-order use: a cloud frame is still open
        if (ctx->shadow_pending) {  // K6 reads the god-ray froxels
            SKY_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_shadow, 0));
            ctx->shadow_pending = false;
        }
        // what the cloud chain needs from the caller's stream (depth, LUTs) is complete here; K6 is not part of it
        SKY_CUDA(ctx, cudaEventRecord(ctx->ev_pre_composite, ctx->stream));
        ctx->pre_composite_recorded = true;
        ctx->pre_composite_depth = depth;
    }
    return (ctx->strict_arithmetic ? launch_composite_strict : launch_composite)(ctx, depth, static_cast<half4*>(hdr), width, height);
}

int sky_noise_generate(SkyContext* ctx, int kind, const SkyNoiseCreateInfo* info) {
    if (int e = lanes_join(ctx)) return e;
    if (int e = launch_noise(ctx, kind, info)) return e;
    return lut_stream_follow_main(ctx);  // the material textures are inputs of the shadow chain
}

int sky_voxel_upload(SkyContext* ctx, const uint8_t* host_voxels, int dx, int dy, int dz) {
    if (dx < 1 || dy < 1 || dz < 1 || dx > 4096 || dy > 4096 || dz > 4096) return sky_fail(ctx, "voxel grid dimensions out of range");
    if (int e = lanes_join(ctx)) return e;
    ctx->voxel.valid = false;
    ctx->voxel_majorant_valid = false;
    if (int e = build_mip_texture(ctx, ctx->voxel, dx, dy, dz, 1, true)) return e;
    SKY_CUDA(ctx, cudaMemcpyAsync(ctx->voxel.data, host_voxels, size_t(dx) * dy * dz, cudaMemcpyHostToDevice, ctx->stream));
    if (int e = launch_mip_chain(ctx, ctx->voxel)) return e;
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int sky_set_material(SkyContext* ctx, const SkyMaterialBlock* m) {
    if (m->type < SKY_MATERIAL_DEFAULT0 || m->type > SKY_MATERIAL_VOXEL) return sky_fail(ctx, "unknown material type");
    ctx->material = *m;
    return 0;
}

int sky_cloud_shadow(SkyContext* ctx, const SkyCloudCommonBufferData* common) {
    if (ctx->width == 0) return sky_fail(ctx, "Volumetric cloud viewport is undefined");  // VolumetricCloud.cpp:169-170
    if (ctx->pipelining) {
        // K11-K13 run on lut_stream after this frame's bake, beside the K6 / K16 of the frame in flight.  What they WRITE and that
        // frame still READS is double-buffered: the froxels (K6, K16) and the blurred shadow map shadow_maps[2] (the object
        // branch of K6, ComputeObjectLuminance's cloud shadow).  The raw / half-blurred maps [0], [1] are only touched by the
        // chain itself (same stream).  The set written now was last read two frames ago, i.e. before the event lut_stream
        // waited for at this frame's bake.
        Lut<uint16_t>& other = ctx->alt.shadow_froxel;
        if (other.w != ctx->shadow_froxel.w || other.h != ctx->shadow_froxel.h || other.d != ctx->shadow_froxel.d) {
            if (int e = sky_alloc(ctx, other, ctx->shadow_froxel.w, ctx->shadow_froxel.h, ctx->shadow_froxel.d)) return e;
            ctx->bake_since_shadow = false;
        }
        Lut<float2>& other_map = ctx->alt.shadow_blurred;
        if (other_map.w != ctx->shadow_maps[2].w || other_map.h != ctx->shadow_maps[2].h) {
            if (int e = sky_alloc(ctx, other_map, ctx->shadow_maps[2].w, ctx->shadow_maps[2].h)) return e;
            ctx->bake_since_shadow = false;
        }
        // sharded frame: the chain runs on its own stream (context.h); switching between the two streams orders conservatively once
        const bool own_stream = ctx->out_band_count > 1 && ctx->shadow_stream != nullptr;
        if (!ctx->bake_since_shadow || own_stream != ctx->shadow_stream_last) {  // out-of-protocol call (no bake since the last shadow pass) / stream switch
            if (int e = lanes_join(ctx)) return e;
            if (int e = lut_stream_follow_main(ctx)) return e;
        }
        ctx->shadow_stream_last = own_stream;
        ctx->bake_since_shadow = false;
        std::swap(ctx->shadow_froxel, other);
        std::swap(ctx->shadow_maps[2], other_map);
        ctx->pre_composite_recorded = false;  // a new frame
        if (own_stream) {
            if (ctx->frame_gate) SKY_CUDA(ctx, cudaStreamWaitEvent(ctx->shadow_stream, ctx->frame_gate, 0));   // the write set was last read before it
            ctx->shadow_stream_pending = true;
            LaneScope lane(ctx, ctx->shadow_stream);
            return (ctx->strict_arithmetic ? launch_cloud_shadow_strict : launch_cloud_shadow)(ctx, *common);
        }
        ctx->luts_pending = true;
        LaneScope lane(ctx, ctx->lut_stream);
        return (ctx->strict_arithmetic ? launch_cloud_shadow_strict : launch_cloud_shadow)(ctx, *common);
    }
    if (!ctx->overlap) return (ctx->strict_arithmetic ? launch_cloud_shadow_strict : launch_cloud_shadow)(ctx, *common);
    if (int e = lanes_join(ctx)) return e;
    ctx->pre_composite_recorded = false;  // a new frame
    SKY_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
    if (int e = lane2_fork(ctx, ctx->ev_fork)) return e;
    {
        LaneScope lane(ctx, ctx->lane2);
        if (int e = (ctx->strict_arithmetic ? launch_cloud_shadow_strict : launch_cloud_shadow)(ctx, *common)) return e;
    }
    SKY_CUDA(ctx, cudaEventRecord(ctx->ev_shadow, ctx->lane2));
    ctx->shadow_pending = true;
    return 0;
}

int sky_cloud_frame_begin(SkyContext* ctx, const SkyCloudCommonBufferData* common, const SkyCloudBufferData* cloud,
                          const float* depth, int band_rows, int band_index, int band_count) {
    if (ctx->width == 0) return sky_fail(ctx, "Volumetric cloud viewport is undefined");
    if (band_count < 1 || band_index < 0 || band_index >= band_count || band_rows < 0) return sky_fail(ctx, "bad band arguments");
    ctx->last_common = *common;  // cloud_frame_end runs K17/K18 with the same uniforms
    if (int e = luts_join(ctx)) return e;
    if (!ctx->overlap) return (ctx->strict_arithmetic ? launch_cloud_begin_strict : launch_cloud_begin)(ctx, *common, *cloud, depth, band_rows, band_index, band_count);
    // K14-K16 need the depth and the LUTs from the caller's stream -- complete before the composite if there was one --
    // and the froxels, which are on lane2 already
    if (!ctx->pre_composite_recorded || ctx->pre_composite_depth != depth) SKY_CUDA(ctx, cudaEventRecord(ctx->ev_pre_composite, ctx->stream));
    if (int e = lane2_fork(ctx, ctx->ev_pre_composite)) return e;
    ctx->pre_composite_recorded = false;
    ctx->shadow_pending = false;  // from here on the caller's stream joins lane2 as a whole
    ctx->lane2_reads_luts = true;
    LaneScope lane(ctx, ctx->lane2);
    return (ctx->strict_arithmetic ? launch_cloud_begin_strict : launch_cloud_begin)(ctx, *common, *cloud, depth, band_rows, band_index, band_count);
}

int sky_cloud_frame(SkyContext* ctx, const SkyCloudCommonBufferData* common, const SkyCloudBufferData* cloud,
                    const float* depth, void* hdr) {
    if (int e = sky_cloud_frame_begin(ctx, common, cloud, depth, 0, 0, 1)) return e;
    return sky_cloud_frame_end(ctx, depth, hdr);
}

int sky_cloud_frame_end(SkyContext* ctx, const float* depth, void* hdr) {
    if (ctx->width == 0) return sky_fail(ctx, "Volumetric cloud viewport is undefined");
    if (!ctx->overlap || !ctx->lane2_reads_luts) {
        if (int e = lanes_join(ctx)) return e;
        return (ctx->strict_arithmetic ? launch_cloud_end_strict : launch_cloud_end)(ctx, ctx->last_common, depth, static_cast<half4*>(hdr), 3);
    }
    {
        LaneScope lane(ctx, ctx->lane2);  // K17 follows K16
        if (int e = (ctx->strict_arithmetic ? launch_cloud_end_strict : launch_cloud_end)(ctx, ctx->last_common, depth, static_cast<half4*>(hdr), 1)) return e;
    }
    if (int e = lanes_join(ctx)) return e;  // K18 composites over what K6 wrote
    return (ctx->strict_arithmetic ? launch_cloud_end_strict : launch_cloud_end)(ctx, ctx->last_common, depth, static_cast<half4*>(hdr), 2);
}

int sky_peer_detach(SkyContext* ctx) {
    if (int e = lanes_join(ctx)) return e;
    for (int k = 0; k < ctx->peer_world; ++k) {
        if (k == ctx->peer_rank) continue;
        if (ctx->peer_render[k]) cudaIpcCloseMemHandle(ctx->peer_render[k]);
        if (ctx->peer_distance[k]) cudaIpcCloseMemHandle(ctx->peer_distance[k]);
        if (ctx->peer_flags[k]) cudaIpcCloseMemHandle(ctx->peer_flags[k]);
        if (ctx->peer_hdr[k]) cudaIpcCloseMemHandle(ctx->peer_hdr[k]);
    }
    for (int k = 0; k < SKY_MAX_PEERS; ++k) { ctx->peer_render[k] = nullptr; ctx->peer_distance[k] = nullptr; ctx->peer_flags[k] = nullptr; ctx->peer_hdr[k] = nullptr; }
    ctx->peer_rank = 0; ctx->peer_world = 1; ctx->peer_band_frame = false;
    return 0;
}

int sky_peer_export(SkyContext* ctx, SkyPeerHandles* out) {
    if (int e = lanes_join(ctx)) return e;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "SkyPeerHandles carries 64-byte IPC handles");
    if (ctx->width == 0) return sky_fail(ctx, "Volumetric cloud viewport is undefined");
    if (!ctx->my_flags) {
        // [0, 8): "rows arrived" epochs, [8, 16): "finished reading" epochs, one slot per peer; [16]: a wait timed out
        SKY_CUDA(ctx, cudaMalloc(&ctx->my_flags, SKY_PEER_FLAG_SLOTS * sizeof(unsigned int)));
    }
    // Export + attach is a collective (every rank needs every rank's handles, so nobody can attach before everybody has
    // exported): all ranks restart from epoch 0 with cleared flags HERE, which keeps the per-context epochs in step even
    // when one rank re-created its context and another kept its own.
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SKY_CUDA(ctx, cudaMemset(ctx->my_flags, 0, SKY_PEER_FLAG_SLOTS * sizeof(unsigned int)));
    ctx->peer_epoch = 0;
    cudaIpcMemHandle_t h;
    SKY_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->render_texture.p));
    std::memcpy(out->render, &h, 64);
    SKY_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->cloud_distance.p));
    std::memcpy(out->distance, &h, 64);
    SKY_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->my_flags));
    std::memcpy(out->flags, &h, 64);
    if (int e = sky_alloc(ctx, ctx->frame_hdr, ctx->width, ctx->height)) return e;   // the frame target peers store their row bands into
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SKY_CUDA(ctx, cudaIpcGetMemHandle(&h, ctx->frame_hdr.p));
    std::memcpy(out->hdr, &h, 64);
    return 0;
}

int sky_peer_attach(SkyContext* ctx, int rank, int world_size, const SkyPeerHandles* all) {
    if (int e = lanes_join(ctx)) return e;
    if (world_size < 1 || world_size > SKY_MAX_PEERS || rank < 0 || rank >= world_size) return sky_fail(ctx, "bad rank / world size");
    if (!ctx->my_flags) return sky_fail(ctx, "peer_export must be called first");
    sky_peer_detach(ctx);
    ctx->peer_rank = rank; ctx->peer_world = world_size;
    for (int k = 0; k < world_size; ++k) {
        if (k == rank) {
            ctx->peer_render[k] = ctx->render_texture.p;
            ctx->peer_distance[k] = ctx->cloud_distance.p;
            ctx->peer_flags[k] = ctx->my_flags;
            ctx->peer_hdr[k] = ctx->frame_hdr.p;
            continue;
        }
        cudaIpcMemHandle_t h;
        void* p = nullptr;
        std::memcpy(&h, all[k].render, 64);
        SKY_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_render[k] = static_cast<half4*>(p);
        std::memcpy(&h, all[k].distance, 64);
        SKY_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_distance[k] = static_cast<float*>(p);
        std::memcpy(&h, all[k].flags, 64);
        SKY_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_flags[k] = static_cast<unsigned int*>(p);
        std::memcpy(&h, all[k].hdr, 64);
        SKY_CUDA(ctx, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_hdr[k] = static_cast<half4*>(p);
    }
    return 0;
}

int sky_cloud_frame_host(SkyContext* ctx, const SkyCloudCommonBufferData* common, const SkyCloudBufferData* cloud,
                         const float* depth_host, void* hdr_host) {
    if (ctx->width == 0) return sky_fail(ctx, "Volumetric cloud viewport is undefined");
    if (int e = ensure_stage(ctx)) return e;
    size_t px = ctx->stage_pixels;
    SKY_CUDA(ctx, cudaMemcpyAsync(ctx->stage_depth, depth_host, px * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    SKY_CUDA(ctx, cudaMemcpyAsync(ctx->stage_hdr, hdr_host, px * sizeof(half4), cudaMemcpyHostToDevice, ctx->stream));
    if (int e = sky_cloud_frame(ctx, common, cloud, ctx->stage_depth, ctx->stage_hdr)) return e;
    SKY_CUDA(ctx, cudaMemcpyAsync(hdr_host, ctx->stage_hdr, px * sizeof(half4), cudaMemcpyDeviceToHost, ctx->stream));
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int sky_pt_begin(SkyContext* ctx, const SkyPathTracingInit* init) {
    if (int e = lanes_join(ctx)) return e;
    if (ctx->width == 0) return sky_fail(ctx, "Volumetric cloud viewport is undefined");
    if (init->prng != SKY_PRNG_WANG && init->prng != SKY_PRNG_PCG) return sky_fail(ctx, "unknown PRNG");
    if (init->environment_lighting < SKY_ENV_OFF || init->environment_lighting > SKY_ENV_GROUND_MULTI_BOUNCE) return sky_fail(ctx, "unknown environment lighting mode");
    ctx->pt = *init;
    int rc = sky_alloc(ctx, ctx->pt_accum, ctx->width, ctx->height);  // glClearTexImage, VolumetricCloud.cpp:498-501
    rc |= sky_alloc(ctx, ctx->pt_mask, ctx->width, ctx->height);
    return rc;
}

int sky_pt_samples(SkyContext* ctx, const SkyCloudCommonBufferData* common, uint32_t frame_begin, uint32_t count,
                   const int32_t region[4]) {
    if (int e = lanes_join(ctx)) return e;
    return (ctx->strict_arithmetic ? launch_pt_samples_strict : launch_pt_samples)(ctx, *common, frame_begin, count, region);
}

int sky_pt_resolve(SkyContext* ctx, uint32_t frame_count, void* hdr) {
    if (int e = lanes_join(ctx)) return e;
    return (ctx->strict_arithmetic ? launch_pt_resolve_strict : launch_pt_resolve)(ctx, frame_count, static_cast<half4*>(hdr));
}

int sky_pt_samples_host(SkyContext* ctx, const SkyCloudCommonBufferData* common, uint32_t frame_begin, uint32_t count,
                        const int32_t region[4], float* accum_host) {
    if (int e = lanes_join(ctx)) return e;
    if (int e = (ctx->strict_arithmetic ? launch_pt_samples_strict : launch_pt_samples)(ctx, *common, frame_begin, count, region)) return e;
    SKY_CUDA(ctx, cudaMemcpyAsync(accum_host, ctx->pt_accum.p, ctx->pt_accum.bytes(), cudaMemcpyDeviceToHost, ctx->stream));
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int sky_get_resource(SkyContext* ctx, int resource, SkyResourceDesc* out) {
    std::memset(out, 0, sizeof(*out));
    ResView v;
    if (!resolve(ctx, resource, v)) return sky_fail(ctx, "unknown resource");
    if (!v.ptr) return sky_fail(ctx, "resource " + std::to_string(resource) + " has not been created yet");
    out->ptr = v.ptr; out->width = v.w; out->height = v.h; out->depth = v.d; out->channels = v.ch; out->format = v.fmt; out->bytes = v.bytes;
    return 0;
}

int sky_read_resource(SkyContext* ctx, int resource, void* host_dst, uint64_t bytes) {
    SkyResourceDesc d;
    if (int e = sky_get_resource(ctx, resource, &d)) return e;
    if (bytes != d.bytes) return sky_fail(ctx, "read_resource: size mismatch, expected " + std::to_string(d.bytes));
    if (int e = lanes_join(ctx)) return e;
    SKY_CUDA(ctx, cudaMemcpyAsync(host_dst, d.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ShadowMap(2048, 2048) + ClearBindViewport's glClearDepthf(1) (AppWindow.cpp:25, ShadowMap.cpp:8-27)
extern "C++" int ensure_mesh_shadow_map(SkyContext* ctx) {
    if (ctx->mesh_shadow_map.p) return 0;
    if (int e = sky_alloc(ctx, ctx->mesh_shadow_map, 2048, 2048, 1, false)) return e;
    std::vector<float> ones(size_t(2048) * 2048, 1.0f);
    SKY_CUDA(ctx, cudaMemcpyAsync(ctx->mesh_shadow_map.p, ones.data(), ones.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int sky_write_resource(SkyContext* ctx, int resource, const void* host_src, uint64_t bytes) {
    if (int e = lanes_join(ctx)) return e;
    if (resource == SKY_RES_MESH_SHADOW_MAP) { if (int e = ensure_mesh_shadow_map(ctx)) return e; }
    switch (resource) {
        case SKY_RES_CLOUD_MAP: if (int e = build_mip_texture(ctx, ctx->cloud_map, 512, 512, 1, 2, false)) return e; break;
        case SKY_RES_DETAIL: if (int e = build_mip_texture(ctx, ctx->detail, 128, 128, 128, 1, false)) return e; break;
        case SKY_RES_DISPLACEMENT: if (int e = build_mip_texture(ctx, ctx->displacement, 128, 128, 1, 4, false)) return e; break;
        case SKY_RES_CLOUD_MAP_MIPS: case SKY_RES_DETAIL_MIPS: case SKY_RES_DISPLACEMENT_MIPS: case SKY_RES_VOXEL_MIPS:
        case SKY_RES_VOXEL: case SKY_RES_COUNTERS: case SKY_RES_PT_MASK:
            return sky_fail(ctx, "write_resource: resource is not writable");
    }
    SkyResourceDesc d;
    if (int e = sky_get_resource(ctx, resource, &d)) return e;
    if (bytes != d.bytes) return sky_fail(ctx, "write_resource: size mismatch, expected " + std::to_string(d.bytes));
    SKY_CUDA(ctx, cudaMemcpyAsync(d.ptr, host_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    if (resource == SKY_RES_TRANSMITTANCE || resource == SKY_RES_MULTISCATTERING) { if (int e = launch_lut_half_copies(ctx)) return e; }
    if (resource == SKY_RES_SKY_VIEW_LUMINANCE || resource == SKY_RES_SKY_VIEW_TRANSMITTANCE || resource == SKY_RES_AERIAL_LUMINANCE || resource == SKY_RES_AERIAL_TRANSMITTANCE) {
        if (int e = launch_frame_lut_half_copies(ctx)) return e;
    }
    if (resource == SKY_RES_CLOUD_MAP) { if (int e = launch_mip_chain(ctx, ctx->cloud_map)) return e; }
    if (resource == SKY_RES_DETAIL) { if (int e = launch_mip_chain(ctx, ctx->detail)) return e; }
    if (resource == SKY_RES_DISPLACEMENT) { if (int e = launch_mip_chain(ctx, ctx->displacement)) return e; }
    SKY_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return 0;
}

int sky_launch_count(SkyContext* ctx, uint64_t* launches) {
    if (!ctx || !launches) return 1;
    *launches = ctx->launch_count;
    return 0;
}

int sky_counters_enable(SkyContext* ctx, int enable) {
    if (int e = lanes_join(ctx)) return e;
    ctx->counting = enable != 0;
    SKY_CUDA(ctx, cudaMemsetAsync(ctx->counters, 0, 8 * sizeof(unsigned long long), ctx->stream));
    return 0;
}

int sky_set_output_bands(SkyContext* ctx, int band_rows, int band_index, int band_count) {
    if (!ctx) return 1;
    if (band_count <= 1) { ctx->out_band_rows = 0; ctx->out_band_index = 0; ctx->out_band_count = 1; return 0; }
    if (band_rows < 8 || band_rows % 8 != 0 || band_index < 0 || band_index >= band_count) return sky_fail(ctx, "set_output_bands: band_rows must be a positive multiple of 8 and 0 <= band_index < band_count");
    ctx->out_band_rows = band_rows; ctx->out_band_index = band_index; ctx->out_band_count = band_count;
    return 0;
}

int sky_set_launch_shape(SkyContext* ctx, int kernel, int shape) {
    if (!ctx) return 1;
    if (kernel != SKY_KERNEL_K16 || shape < SKY_K16_AUTO || shape > SKY_K16_LITERAL) return sky_fail(ctx, "set_launch_shape: unknown kernel or shape");
    ctx->k16_group = shape == SKY_K16_WAVE_8x4 ? 8 : shape == SKY_K16_WAVE_4x8 ? 4 : 0;
    ctx->k16_literal = shape == SKY_K16_LITERAL;
    return 0;
}

int sky_set_output_gather(SkyContext* ctx, int mode) {
    if (!ctx) return 1;
    if (mode != SKY_GATHER_OFF && mode != SKY_GATHER_ALL && mode != SKY_GATHER_ROOT) return sky_fail(ctx, "set_output_gather: unknown mode");
    ctx->out_gather = mode;
    return 0;
}

int sky_pt_set_tracking(SkyContext* ctx, int mode) {
    if (!ctx) return 1;
    if (mode != SKY_PT_TRACKING_REFERENCE && mode != SKY_PT_TRACKING_MAJORANT_GRID) return sky_fail(ctx, "pt_set_tracking: unknown mode");
    ctx->pt_tracking = mode;
    return 0;
}

int sky_tonemap(SkyContext* ctx, const void* hdr_dev, int width, int height, const SkyToneMapParams* params, void* rgba8_dev) {
    if (!ctx || !hdr_dev || !params || !rgba8_dev || width <= 0 || height <= 0) return ctx ? sky_fail(ctx, "tonemap: bad arguments") : 1;
    if (int e = lanes_join(ctx)) return e;
    return (ctx->strict_arithmetic ? launch_tonemap_strict : launch_tonemap)(ctx, static_cast<const half4*>(hdr_dev), width, height, *params, rgba8_dev);
}

int sky_set_strict_arithmetic(SkyContext* ctx, int enable) {
    if (!ctx) return 1;
    ctx->strict_arithmetic = enable != 0;
    return 0;
}

int sky_set_lut_arithmetic(SkyContext* ctx, int mode) {
    if (!ctx) return 1;
    if (mode != SKY_LUT_EXACT && mode != SKY_LUT_COOPERATIVE) return sky_fail(ctx, "set_lut_arithmetic: unknown mode");
    ctx->lut_arithmetic = mode;
    return 0;
}

int sky_set_hw_filtering(SkyContext* ctx, int enable) {
    ctx->hw_filtering = enable != 0;
    return 0;
}

int sky_tex_peak(SkyContext* ctx, int mode, double* fetches_per_second) {
    if (int e = lanes_join(ctx)) return e;
    return launch_tex_peak(ctx, mode, fetches_per_second);
}

}  // extern "C"
