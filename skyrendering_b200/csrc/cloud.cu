// K11-K18: cloud shadow chain and the real-time quarter-resolution cloud raymarch with temporal
// reconstruction and depth-aware upscale, sm_100a.  Follows shaders/SkyRendering/
// VolumetricCloud{ShadowMap,ShadowMapBlur,ShadowFroxel,IndexGen,Render,Reconstruct,Upscale}.comp,
// CheckerboardGen.comp and VolumetricCloud{Common,ShadowInterface}.glsl.
//
// Bounds: the reference dispatches ceil-div groups with no in-shader bounds checks and relies on GL
// dropping out-of-range image stores (GLReloadableProgram.h:55-59); every kernel here guards explicitly.
// Compiled twice (Makefile): the production objects keep FMA contraction and the approximate division / sqrt / exp of
// -use_fast_math (frames are compared at relative RMS 1e-2); with -DSKY_STRICT_TU the same source is built with
// -fmad=false and IEEE division / sqrt into *_strict entry points (sky_set_strict_arithmetic), which follow the
// oracle's unfused fp32 arithmetic operation by operation -- the validation mode the bit-level parity tests use.
#ifdef SKY_STRICT_TU
#define launch_cloud_shadow launch_cloud_shadow_strict
#define launch_cloud_begin launch_cloud_begin_strict
#define launch_cloud_end launch_cloud_end_strict
#endif
#include "atmosphere_dev.cuh"
#include "context.h"
#include "material_dev.cuh"

namespace {

constexpr float kMinTransmittance = 0.01f;  // VolumetricCloudCommon.glsl:30

// K16: clear-air steps a lane may take per trip before the warp shades its pending dense steps
#ifndef SKY_K16_MARCH_STEPS
#define SKY_K16_MARCH_STEPS 4
#endif
constexpr int kMarchSteps = SKY_K16_MARCH_STEPS;
#ifndef SKY_K16_OCC
#define SKY_K16_OCC 5  // resident 128-thread blocks per SM the register budget is cut for
#endif
#ifndef SKY_K16_BLOCK
#define SKY_K16_BLOCK 128  // threads per block: 32, 64 or 128; a warp renders an 8x4 block of rays, warps pair up to 16 columns
#endif
constexpr int kK16Block = SKY_K16_BLOCK;
constexpr int kK16TileW = kK16Block >= 64 ? 16 : 8, kK16TileH = kK16Block >= 64 ? kK16Block / 16 : 4;

struct CloudParams {
    SkyCloudCommonBufferData c;  // VolumetricCloudCommon.glsl:6-28
    SkyCloudBufferData b;        // VolumetricCloudRender.comp:17-32
    AtmosphereModel atm;
    MaterialParams mat;
    LutView transmittance, ap_lum, ap_trans;
    FroxelView froxel;
    const uint16_t* blue_noise;
    unsigned long long* counters;
    // shadow chain
    const float2* shadow_prev;  // pre_raw_cloud_shadow_map
    float2* shadow_out;
    int out_band_rows, out_band_index, out_band_count;  // sky_set_output_bands (K18)
    const float2* shadow_blurred;
    uint16_t* froxel_out;
    int shadow_w, shadow_h;
    // viewport chain
    int width, height;         // full resolution
    const float* depth;        // full-res depth
    float* checkerboard;       // W/2 x H/2
    float2* index_linear;      // W/4 x H/4
    half4* render;             // W/4 x H/4
    float* cloud_distance;     // W/4 x H/4
    const half4* reconstruct_prev;
    half4* reconstruct_out;    // W/2 x H/2
    half4* hdr;
    int band_rows, band_index, band_count;
    float inv_thickness;  // 1 / (uTopAltitude - uBottomAltitude)
    // K16 output replicas in peer memory (every rank's buffers, own included); peer_count == 0: local only
    half4* peer_render[8];
    float* peer_distance[8];
    int peer_count;
    // K18: frame targets of the ranks that receive this rank's row bands (sky_set_output_gather); hdr_peer_count == 0: local only
    half4* hdr_peers[8];
    int hdr_peer_count;
};

SKY_D float DepthToLinearDepth(const SkyCloudCommonBufferData& c, float depth) {  // VolumetricCloudCommon.glsl:32-34
    return 1.0f / (c.uLinearDepthParam[0] - c.uLinearDepthParam[1] * depth);
}
SKY_D float CalHeight01(const CloudParams& P, float3 pos) {  // VolumetricCloudCommon.glsl:36-39
    const SkyCloudCommonBufferData& c = P.c;
    float altitude = length(f3(pos.x, pos.y, pos.z + c.uEarthRadius)) - c.uEarthRadius;
    return clampf((altitude - c.uBottomAltitude) * P.inv_thickness, 0.0f, 1.0f);  // the divisor is a uniform
}
SKY_D int2 IndexToOffset(uint32_t index) {  // VolumetricCloudCommon.glsl:42-52
    return make_int2(int(((index + 1) >> 1) & 1), int(((index + 2) >> 1) & 1));
}
SKY_D float HenyeyGreenstein(float cos_theta, float g) {  // VolumetricCloudCommon.glsl:58-63
    float a = 1.0f - g * g;
    float b = 1.0f + g * g - 2.0f * g * cos_theta;
    b *= sqrtf(b);
    return (0.25f * kInvPi) * a / b;
}
SKY_D float blue_noise_at(const uint16_t* bn, int x, int y) { return float(__ldg(bn + (y & 0x3f) * 64 + (x & 0x3f))) / 65535.0f; }

// GL_LINEAR + CLAMP_TO_BORDER(1e10, 1) on the RG32F cloud shadow map (VolumetricCloud.cpp:106-112)
SKY_D float2 sample_shadow_map(const float2* p, int w, int h, float u, float v) {
    float x = u * float(w) - 0.5f, y = v * float(h) - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    // keep the int conversion defined for far-away points
    int i0 = int(fminf(fmaxf(fx, -2.0f), float(w) + 1.0f)), j0 = int(fminf(fmaxf(fy, -2.0f), float(h) + 1.0f));
    auto fetch = [&](int i, int j) {
        if (i < 0 || i >= w || j < 0 || j >= h) return f2(1e10f, 1.0f);
        return __ldg(p + j * w + i);
    };
    float2 t00 = fetch(i0, j0), t10 = fetch(i0 + 1, j0), t01 = fetch(i0, j0 + 1), t11 = fetch(i0 + 1, j0 + 1);
    return (1.0f - a) * (1.0f - b) * t00 + a * (1.0f - b) * t10 + (1.0f - a) * b * t01 + a * b * t11;
}
// VolumetricCloudShadowInterface.glsl:4-8
SKY_D float SampleCloudShadowTransmittance(const float2* map, int w, int h, float3 light_ndc) {
    const float kInvTransitionDepth = 1.0f / 0.5f;
    float2 dt = sample_shadow_map(map, w, h, light_ndc.x * 0.5f + 0.5f, light_ndc.y * 0.5f + 0.5f);
    return mixf(dt.y, 1.0f, clampf((dt.x - light_ndc.z) * kInvTransitionDepth, 0.0f, 1.0f));
}

// ------------------------------------------------------------------------------------------------ K11
// VolumetricCloudShadowMap.comp:37-75: ortho light-space march through the cloud shell, blue-noise +
// golden-ratio jitter, temporal blend 0.2 with the reprojected previous raw map.
template <int MAT, bool HW, bool COUNT>
__global__ void __launch_bounds__(128) k11_shadow_map(const __grid_constant__ CloudParams P) {
    const SkyCloudCommonBufferData& c = P.c;
    int gx = blockIdx.x * 16 + (threadIdx.x & 15), gy = blockIdx.y * 8 + (threadIdx.x >> 4);
    if (gx >= P.shadow_w || gy >= P.shadow_h) return;
    float2 image_size = f2(float(P.shadow_w), float(P.shadow_h));
    float2 uv = f2((float(gx) + 0.5f) / image_size.x, (float(gy) + 0.5f) / image_size.y);
    float3 origin = projective_mul(c.uInvLightVP, f3(uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f, 0.0f));
    float3 dir = -f3(c.uSunDirection);
    float3 up = f3(origin.x, origin.y, origin.z + c.uEarthRadius);
    float r = length(up);
    up /= r;
    float mu = dot(up, dir);
    // LineShellFirstIntersect, :18-35
    float t1 = 0.0f, t2 = 0.0f;
    {
        float bottom_radius = c.uEarthRadius + c.uBottomAltitude;
        float top_radius = c.uEarthRadius + c.uTopAltitude;
        float discriminant_top = r * r * (mu * mu - 1.0f) + top_radius * top_radius;
        if (discriminant_top > 0) {
            float discriminant_bottom = r * r * (mu * mu - 1.0f) + bottom_radius * bottom_radius;
            float sqrt_discriminant_top = sqrtf(discriminant_top);
            float sqrt_discriminant_bottom = sqrtf(discriminant_bottom);
            t1 = -r * mu - sqrt_discriminant_top;
            t2 = -r * mu + (discriminant_bottom >= 0 ? -sqrt_discriminant_bottom : sqrt_discriminant_top);
        }
    }
    float dist = fmaxf(t2 - t1, 0.0f);
    dist = fminf(dist, 1.0f / cosf(85.0f * 0.01745329251994329576923690768489f) * (c.uTopAltitude - c.uBottomAltitude));
    float optical_depth = 0.0f;
    int evals = 0;
    if (dist > 0.0f) {
        float steps = mixf(12.0f, 6.0f, fabsf(dir.z));
        float step_size = dist / steps;
        float noise = blue_noise_at(P.blue_noise, gx, gy);
        float t = t1 + step_size * fractf(noise + c.uFrameID * 0.61803398875f);
        for (uint32_t cnt = uint32_t(steps); cnt != 0; cnt--, t += step_size) {
            float3 pos = origin + t * dir;
            float height01 = CalHeight01(P, pos);
            float sigma_t = SampleSigmaT<MAT, HW>(P.mat, pos, height01);
            optical_depth += sigma_t * step_size;
            if (COUNT) ++evals;
        }
    }
    float transmittance = expf(-optical_depth);
    float depth = mixf(t1, t2, 0.5f);
    float2 res = f2(depth, transmittance);
    float3 pre = projective_mul(c.uShadowMapReprojectMat, f3(uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f, 0.0f));
    float2 pre_uv = f2(pre.x * 0.5f + 0.5f, pre.y * 0.5f + 0.5f);
    float2 lo = f2(0.5f / image_size.x, 0.5f / image_size.y);
    float2 hi = f2((image_size.x - 0.5f) / image_size.x, (image_size.y - 0.5f) / image_size.y);
    if (clampf(pre_uv.x, lo.x, hi.x) == pre_uv.x && clampf(pre_uv.y, lo.y, hi.y) == pre_uv.y) {
        float2 pre_res = sample_shadow_map(P.shadow_prev, P.shadow_w, P.shadow_h, pre_uv.x, pre_uv.y);
        res = mix2(pre_res, res, 0.2f);
    }
    P.shadow_out[gy * P.shadow_w + gx] = res;
    if (COUNT && evals) atomicAdd(P.counters + SKY_CNT_SHADOW_SIGMA_EVALS, (unsigned long long)evals);
}

// ------------------------------------------------------------------------------------------------ K12
// VolumetricCloudShadowMapBlur.comp:14-42: separable 9-tap Gaussian, clamp to edge.
template <bool HORIZONTAL>
__global__ void __launch_bounds__(128) k12_blur(const float2* __restrict__ in, float2* __restrict__ out, int w, int h) {
    const float weight[5] = {0.227027f, 0.1945946f, 0.1216216f, 0.054054f, 0.016216f};
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= w) return;
    float2 res = __ldg(in + y * w + x) * weight[0];
#pragma unroll
    for (int i = 1; i < 5; ++i) {
        int x1 = x, y1 = y, x2 = x, y2 = y;
        if (HORIZONTAL) { x1 = max(x - i, 0); x2 = min(x + i, w - 1); }
        else { y1 = max(y - i, 0); y2 = min(y + i, h - 1); }
        res += __ldg(in + y1 * w + x1) * weight[i];
        res += __ldg(in + y2 * w + x2) * weight[i];
    }
    out[y * w + x] = res;
}

// ------------------------------------------------------------------------------------------------ K13
// VolumetricCloudShadowFroxel.comp:10-27: running mean of cloud-shadow transmittance along each view ray.
// Only the sum is sequential; the shadow-map taps (a projection + a bilinear RG32F fetch each, all of the work) depend on the slice index
// alone.  One thread per column walking its 128 slices -- the shader's mapping -- is a handful of warps per SM in a long dependent loop
// (ncu profiles/k13_r02v.md: 76 us at 4K, occupancy 15 %, 0.59 eligible warps; 1080p is a quarter of the columns and takes as long).
// So a block takes 32 columns x kK13Slabs slabs of slices: warp g evaluates the taps of slab g for the block's 32 columns into shared
// memory (its t is the shader's running `t += step_size`, advanced to the slab's first slice), then warp 0 -- one lane per column --
// adds them in slice order, and every warp turns its slab's sums into means and stores them.  Same operations on the same operands in the same order as the serial loop: bit-identical.
constexpr int kK13Slabs = 8, kK13Columns = 32, kK13MaxDepth = 128;
__global__ void __launch_bounds__(kK13Slabs * kK13Columns, 3) k13_shadow_froxel(const __grid_constant__ CloudParams P) {
    const SkyCloudCommonBufferData& c = P.c;
    const int FW = P.froxel.w, FH = P.froxel.h, FD = P.froxel.d;
    __shared__ float taps[kK13MaxDepth][kK13Columns];
    const int col = threadIdx.x & (kK13Columns - 1), slab = threadIdx.x / kK13Columns;
    const int gx = blockIdx.x * kK13Columns + col, gy = blockIdx.y;
    const int per_slab = (FD + kK13Slabs - 1) / kK13Slabs, z_begin = slab * per_slab, z_end = min(z_begin + per_slab, FD);
    if (gx < FW) {
        float2 uv = f2((float(gx) + 0.5f) / float(FW), (float(gy) + 0.5f) / float(FH));
        float3 camera = f3(c.uCameraPos);
        float3 frag_pos = projective_mul(c.uInvMVP, f3(uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f, 0.0f));
        float step_size = c.uShadowFroxelMaxDistance / float(FD);
        float3 dir = normalize(frag_pos - camera);
        float t = 0.5f * step_size;
        for (int z = 0; z < z_begin; ++z) t += step_size;   // the serial loop's t at this slab's first slice
#pragma unroll 4
        for (int z = z_begin; z < z_end; ++z) {
            float3 pos = camera + t * dir;
            t += step_size;
            float3 light_ndc = projective_mul(c.uLightVP, pos);
            taps[z][col] = SampleCloudShadowTransmittance(P.shadow_blurred, P.shadow_w, P.shadow_h, light_ndc);
        }
    }
    __syncthreads();
    if (slab == 0) {   // the running sum, in slice order, one lane per column; the sums replace the taps
        float transmittance_sum = 0.0f;
#pragma unroll 8
        for (int z = 0; z < FD; ++z) {
            transmittance_sum += taps[z][col];
            taps[z][col] = transmittance_sum;
        }
    }
    __syncthreads();
    if (gx >= FW) return;
    for (int z = z_begin; z < z_end; ++z) {   // mean, quantisation and store of this slab's slices: independent again
        float ray_scatter_visibility = taps[z][col] / float(z + 1);
        P.froxel_out[(size_t(z) * FH + gy) * FW + gx] = uint16_t(__float2int_rn(clampf(ray_scatter_visibility, 0.0f, 1.0f) * 65535.0f));
    }
}

// ------------------------------------------------------------------------------------------------ K14
// CheckerboardGen.comp:7-14: max / min of each 2x2 depth block in a checkerboard pattern.
__global__ void __launch_bounds__(256) k14_checkerboard(const __grid_constant__ CloudParams P) {
    const int HW_ = P.width / 2, HH = P.height / 2;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= HW_) return;
    int x0 = min(2 * x, P.width - 1), x1 = min(2 * x + 1, P.width - 1), y0 = min(2 * y, P.height - 1), y1 = min(2 * y + 1, P.height - 1);
    float v0 = __ldg(P.depth + size_t(y1) * P.width + x0), v1 = __ldg(P.depth + size_t(y1) * P.width + x1);
    float v2 = __ldg(P.depth + size_t(y0) * P.width + x1), v3 = __ldg(P.depth + size_t(y0) * P.width + x0);
    bool bmax = ((x & 1) == (y & 1));
    float d = bmax ? fmaxf(fmaxf(v0, v1), fmaxf(v2, v3)) : fminf(fminf(v0, v1), fminf(v2, v3));
    P.checkerboard[size_t(y) * HW_ + x] = d;
    (void)HH;
}

SKY_D float checker_clamped(const CloudParams& P, int x, int y) {
    const int w = P.width / 2, h = P.height / 2;
    return __ldg(P.checkerboard + size_t(clampi(y, 0, h - 1)) * w + clampi(x, 0, w - 1));
}

// ------------------------------------------------------------------------------------------------ K15
// VolumetricCloudIndexGen.comp:13-42.  Out-of-range neighbour fetches (the reference leans on robust
// texelFetch) are clamped to the edge, as the oracle defines.
__global__ void __launch_bounds__(128) k15_index_gen(const __grid_constant__ CloudParams P) {
    const SkyCloudCommonBufferData& c = P.c;
    const int QW = P.width / 4, QH = P.height / 4;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= QW || y >= QH) return;
    float depth_tile[4] = {checker_clamped(P, 2 * x, 2 * y + 1), checker_clamped(P, 2 * x + 1, 2 * y + 1),
                           checker_clamped(P, 2 * x + 1, 2 * y), checker_clamped(P, 2 * x, 2 * y)};
    uint32_t nearest_index = 0, farthest_index = 0;
#pragma unroll
    for (uint32_t i = 1; i < 4; ++i) {
        nearest_index = depth_tile[i] < depth_tile[nearest_index] ? i : nearest_index;
        farthest_index = depth_tile[i] > depth_tile[farthest_index] ? i : farthest_index;
    }
    float nearest_linear_depth = DepthToLinearDepth(c, depth_tile[nearest_index]);
    float farthest_linear_depth = DepthToLinearDepth(c, depth_tile[farthest_index]);
    uint32_t close_to_nearest_count = 0, close_to_farthest_count = 0;
    const int ox[8] = {0, 1, 1, 1, 0, -1, -1, -1}, oy[8] = {1, 1, 0, -1, -1, -1, 0, 1};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int tx = x + ox[i], ty = y + oy[i];
        uint32_t idx = (c.uBaseShadingIndex + uint32_t((tx + ty) & 1)) & 3u;  // PosToIndex, :9-11
        int2 off = IndexToOffset(idx);
        float depth = checker_clamped(P, (tx << 1) + off.x, (ty << 1) + off.y);
        float linear_depth = DepthToLinearDepth(c, depth);
        float max_delta_allowed = linear_depth * 0.25f;
        if (fabsf(linear_depth - nearest_linear_depth) < max_delta_allowed) ++close_to_nearest_count;
        if (fabsf(linear_depth - farthest_linear_depth) < max_delta_allowed) ++close_to_farthest_count;
    }
    uint32_t own = (c.uBaseShadingIndex + uint32_t((x + y) & 1)) & 3u;
    uint32_t index = close_to_nearest_count == 0 ? nearest_index : close_to_farthest_count == 0 ? farthest_index : own;
    P.index_linear[size_t(y) * QW + x] = f2(float(index), DepthToLinearDepth(c, depth_tile[index]));
}

// ------------------------------------------------------------------------------------------------ K16
// VolumetricCloudRender.comp:139-210, one thread per quarter-res ray, 16x8 tiles (screen-coherent
// rays share texture cache lines).  Tex-pipe / L1 bound; see DESIGN.md for the roofline.
struct RayMarchContext {
    float t;
    float3 pos;
    float height01;
    float step_size;
    float transmittance;
    float transmittance_sum;
    float weighted_t_sum;
    float2 sun_env;
    float cos_sun_view;
};

template <int MAT, bool HW, bool COUNT>
SKY_D float SampleShadow(const CloudParams& P, float3 pos, int& evals, int& fetches) {  // :98-114
    float optical_depth = 0.0f;
    float inv_shadow_steps = 1.0f / P.b.uShadowSteps;
    float3 sample_vector = P.b.uShadowDistance * f3(P.c.uSunDirection);
    float previous_t = 0.0f;
    for (float t = inv_shadow_steps; t <= 1.0f; t += inv_shadow_steps) {
        float current_t = t * t;
        float delta_t = current_t - previous_t;
        float3 sample_pos = pos + sample_vector * (previous_t + 0.5f * delta_t);
        float sample_height01 = CalHeight01(P, sample_pos);
        optical_depth += SampleSigmaT<MAT, HW>(P.mat, sample_pos, sample_height01, COUNT ? &fetches : nullptr) * P.b.uShadowDistance * delta_t;
        if (COUNT) ++evals;
        previous_t = current_t;
    }
    return expf(-optical_depth);
}

template <int MAT, bool HW, bool COUNT>
SKY_D void ShadeDenseStep(const CloudParams& P, RayMarchContext& ctx, float sigma_t, int& evals, int& fetches) {  // :120-136
    float tr = expf(-ctx.step_size * sigma_t);
    float transmittance_to_sun = SampleShadow<MAT, HW, COUNT>(P, ctx.pos, evals, fetches);
    float phase = mixf(HenyeyGreenstein(ctx.cos_sun_view, -0.15f) * 2.16f, HenyeyGreenstein(ctx.cos_sun_view, 0.85f),
                       expf(-P.b.uSunMultiscatteringSigmaScale * sigma_t));
    float2 sun_env;
    sun_env.x = transmittance_to_sun * phase;
    sun_env.y = mixf(P.b.uEnvBottomVisibility, 1.0f, ctx.height01);
    sun_env.y = sun_env.y - sun_env.y * expf(-P.b.uEnvMultiscatteringSigmaScale * sigma_t);
    sun_env = sun_env - sun_env * tr;
    ctx.sun_env += ctx.transmittance * sun_env;
    ctx.transmittance_sum += ctx.transmittance;
    ctx.weighted_t_sum += ctx.t * ctx.transmittance;
    ctx.transmittance *= tr;
}

// Per-ray constants of VolumetricCloudRender.comp:139-169: everything main() computes before its two step loops.
struct RaySetup {
    int px, py;
    bool valid;
    float2 uv;
    float3 view_dir;
    float r, mu, frag_dist;
    float i0t1, i0t2, i1t1, i1t2;
    float jitter, cos_sun_view;
};
SKY_D RaySetup k16_ray_setup(const CloudParams& P, int px, int row_in_band) {   // quarter-res column, row among the rows this launch renders
    const SkyCloudCommonBufferData& c = P.c;
    const SkyCloudBufferData& b = P.b;
    const int QW = P.width / 4, QH = P.height / 4, HW_ = P.width / 2, HH = P.height / 2;
    RaySetup S;
    int py = row_in_band;
    if (P.band_rows > 0) {
        // rows owned by this rank: ((py / band_rows) % band_count) == band_index
        int local_band = row_in_band / P.band_rows;
        py = (local_band * P.band_count + P.band_index) * P.band_rows + row_in_band % P.band_rows;
    }
    // lanes past the image edge march a clamped duplicate ray and skip the stores: every lane reaches the
    // warp votes of the marching loop
    S.valid = px < QW && py < QH;
    px = min(px, QW - 1);
    py = min(py, QH - 1);
    S.px = px; S.py = py;
    uint32_t index = uint32_t(__ldg(P.index_linear + size_t(py) * QW + px).x);
    int2 off = IndexToOffset(index);
    int cx = px * 2 + off.x, cy = py * 2 + off.y;
    S.uv = f2((float(cx) + 0.5f) / float(HW_), (float(cy) + 0.5f) / float(HH));
    float depth = checker_clamped(P, cx, cy);
    float3 camera = f3(c.uCameraPos);
    float3 frag_pos = projective_mul(c.uInvMVP, f3(S.uv.x * 2.0f - 1.0f, S.uv.y * 2.0f - 1.0f, depth * 2.0f - 1.0f));
    S.view_dir = normalize(frag_pos - camera);

    float r = c.uCameraPos[2] + c.uEarthRadius;
    float mu = S.view_dir.z;
    S.r = r; S.mu = mu;
    // RayShellIntersect, :39-71
    float i0t1 = 0.0f, i0t2 = 0.0f, i1t1 = 0.0f, i1t2 = 0.0f;
    {
        float bottom_radius = c.uEarthRadius + c.uBottomAltitude;
        float top_radius = c.uEarthRadius + c.uTopAltitude;
        float discriminant_bottom = r * r * (mu * mu - 1.0f) + bottom_radius * bottom_radius;
        float discriminant_top = r * r * (mu * mu - 1.0f) + top_radius * top_radius;
        float sqrt_discriminant_bottom = sqrtf(discriminant_bottom);
        float sqrt_discriminant_top = sqrtf(discriminant_top);
        if (c.uCameraPos[2] < c.uBottomAltitude) {
            i0t1 = -r * mu + sqrt_discriminant_bottom;
            i0t2 = -r * mu + sqrt_discriminant_top;
        } else if (c.uCameraPos[2] < c.uTopAltitude) {
            if (discriminant_bottom >= 0.0f && mu < 0.0f) {
                i0t2 = -r * mu - sqrt_discriminant_bottom;
                i1t1 = -r * mu + sqrt_discriminant_bottom;
                i1t2 = -r * mu + sqrt_discriminant_top;
            } else {
                i0t2 = -r * mu + sqrt_discriminant_top;
            }
        } else {
            if (discriminant_bottom >= 0.0f && mu < 0.0f) {
                i0t1 = -r * mu - sqrt_discriminant_top;
                i0t2 = -r * mu - sqrt_discriminant_bottom;
                i1t1 = -r * mu + sqrt_discriminant_bottom;
                i1t2 = -r * mu + sqrt_discriminant_top;
            } else if (discriminant_top >= 0.0f && mu < 0.0f) {
                i0t1 = -r * mu - sqrt_discriminant_top;
                i0t2 = -r * mu + sqrt_discriminant_top;
            }
        }
    }
    S.frag_dist = distance(frag_pos, camera);
    float limit = fminf(S.frag_dist, b.uMaxVisibleDistance);
    i0t2 = fminf(fmaxf(limit, i0t1), i0t2);  // clamp(x, minVal, maxVal) = min(max(x, minVal), maxVal)
    i1t2 = fminf(fmaxf(limit, i1t1), i1t2);
    S.i0t1 = i0t1; S.i0t2 = i0t2; S.i1t1 = i1t1; S.i1t2 = i1t2;
    S.cos_sun_view = dot(f3(c.uSunDirection), S.view_dir);
    float noise = blue_noise_at(P.blue_noise, px, py);
    S.jitter = fractf(noise + c.uFrameID * 0.61803398875f);
    return S;
}

// VolumetricCloudRender.comp:189-209: what main() does with the marched context, and the stores (local, or every rank's copy
// over NVLink when the frame is tile-sharded)
SKY_D void k16_ray_finish(const CloudParams& P, const RaySetup& S, RayMarchContext& ctx) {
    const SkyCloudCommonBufferData& c = P.c;
    const SkyCloudBufferData& b = P.b;
    const int QW = P.width / 4;
    const float3 camera = f3(c.uCameraPos), sun = f3(c.uSunDirection);
    float average_t = ctx.weighted_t_sum == 0 ? S.frag_dist : ctx.weighted_t_sum / ctx.transmittance_sum;
    float3 average_pos = camera + S.view_dir * average_t;
    // GetSunVisibility(pos), VolumetricCloudCommon.glsl:73-79
    float3 up_dir = f3(average_pos.x, average_pos.y, average_pos.z + c.uEarthRadius);
    float up_len = length(up_dir);
    up_dir /= up_len;
    float mu_s = dot(sun, up_dir);
    float3 sun_visibility = P.atm.GetSunVisibility(P.transmittance, up_len, mu_s);
    float sun_cos_theta = clampf(dot(normalize(f3(average_pos.x, average_pos.y, average_pos.z + c.uEarthRadius)), sun), 0.0f, 1.0f);  // SunCosTheta :85-88
    float3 luminance = ctx.sun_env.x * b.uSunIlluminanceScale * sun_visibility * P.atm.solar_illuminance() +
                       ctx.sun_env.y * powf(sun_cos_theta, b.uEnvSunHeightCurveExp) * f3(b.uEnvColorScale);
    // GetAerialPerspective, VolumetricCloudCommon.glsl:81-97
    float ap_t = average_t;
    if (S.r > P.atm.u.top_radius) {
        float near_distance;
        if (P.atm.FromSpaceIntersectTopAtmosphereBoundary(S.r, S.mu, near_distance)) ap_t -= near_distance;
        else ap_t = 0;
    }
    float3 uvw = aerial_perspective_uvw(S.uv, ap_t, c.uAerialPerspectiveLutMaxDistance, P.ap_lum.w, P.ap_lum.h, P.ap_lum.d);
    float3 atmosphere_transmittance = xyz(sample_lut3d(P.ap_trans, uvw.x, uvw.y, uvw.z));
    float3 atmosphere_luminance = xyz(sample_lut3d(P.ap_lum, uvw.x, uvw.y, uvw.z));
    atmosphere_luminance *= SampleRayScatterVisibility(P.froxel, S.uv, average_t, c.uInvShadowFroxelMaxDistance);
    luminance = luminance * atmosphere_transmittance + atmosphere_luminance * (1 - ctx.transmittance);

    luminance /= fmaxf(1e-5f, (1 - ctx.transmittance));
    float fade = smoothstepf(b.uMaxVisibleDistance * 0.75f, b.uMaxVisibleDistance, S.i0t1);
    ctx.transmittance = mixf(ctx.transmittance, 1.0f, fade);
    luminance *= 1 - ctx.transmittance;
    const half4 texel = to_half4(f4(luminance, ctx.transmittance));
    const size_t at = size_t(S.py) * QW + S.px;
    if (P.peer_count == 0) {
        P.render[at] = texel;
        P.cloud_distance[at] = average_t;
    } else {
        // the exchange step of a tile-sharded frame, fused into the producer: plain stores to every rank's
        // copy over NVLink (12 B per ray and rank); k_peer_arrive_and_wait orders them before K17
#pragma unroll 1
        for (int k = 0; k < P.peer_count; ++k) {
            P.peer_render[k][at] = texel;
            P.peer_distance[k][at] = average_t;
        }
    }
}

// The tail of RayMarchStep (:120-136) once the shadow march of the step is known
SKY_D void ShadeDenseStepTail(const CloudParams& P, RayMarchContext& ctx, float sigma_t, float transmittance_to_sun) {
    float tr = expf(-ctx.step_size * sigma_t);
    float phase = mixf(HenyeyGreenstein(ctx.cos_sun_view, -0.15f) * 2.16f, HenyeyGreenstein(ctx.cos_sun_view, 0.85f),
                       expf(-P.b.uSunMultiscatteringSigmaScale * sigma_t));
    float2 sun_env;
    sun_env.x = transmittance_to_sun * phase;
    sun_env.y = mixf(P.b.uEnvBottomVisibility, 1.0f, ctx.height01);
    sun_env.y = sun_env.y - sun_env.y * expf(-P.b.uEnvMultiscatteringSigmaScale * sigma_t);
    sun_env = sun_env - sun_env * tr;
    ctx.sun_env += ctx.transmittance * sun_env;
    ctx.transmittance_sum += ctx.transmittance;
    ctx.weighted_t_sum += ctx.t * ctx.transmittance;
    ctx.transmittance *= tr;
}

template <int MAT, bool HW, bool COUNT>
__global__ void __launch_bounds__(kK16Block, SKY_K16_OCC * 128 / kK16Block) k16_render(const __grid_constant__ CloudParams P) {
    const SkyCloudCommonBufferData& c = P.c;
    const SkyCloudBufferData& b = P.b;
    const int warp_in_block = int(threadIdx.x >> 5);
    const RaySetup S = k16_ray_setup(P, blockIdx.x * kK16TileW + (kK16Block >= 64 ? (warp_in_block & 1) * 8 : 0) + int(threadIdx.x & 7u),
                                     blockIdx.y * kK16TileH + (kK16Block >= 64 ? (warp_in_block >> 1) * 4 : 0) + int((threadIdx.x >> 3) & 3u));
    const float3 camera = f3(c.uCameraPos);
    const float3 view_dir = S.view_dir;
    int evals = 0, fetches = 0;
    RayMarchContext ctx;
    ctx.cos_sun_view = S.cos_sun_view;
    float dist = S.i0t2 - S.i0t1;
    dist = fminf(dist, b.uMaxRaymarchDistance);
    uint32_t num_steps = uint32_t(fmaxf(b.uMaxRaymarchSteps * (dist / b.uMaxRaymarchDistance), 1.0f));
    ctx.step_size = dist / float(num_steps);
    ctx.transmittance = 1.0f;
    ctx.transmittance_sum = 0.0f;
    ctx.weighted_t_sum = 0.0f;
    ctx.sun_env = f2(0.0f, 0.0f);
    const float jitter = S.jitter;
    ctx.t = S.i0t1 + ctx.step_size * jitter;
    // The two step loops of :170-188, run as two convergent phases per trip: (A) up to kMarchSteps steps
    // that find no cloud (one SampleSigmaT each), stopping at the first step with cloud in it; (B) the
    // shading of that step (5-tap shadow march + phase functions).  A lane's own sequence of operations is
    // exactly the shader's, but lanes crossing clear air no longer idle through their neighbours' shadow
    // marches, and the shadow marches of a warp run together.
    float dist1 = S.i1t2 - S.i1t1;
    uint32_t cnt = num_steps;
    int segment = 0;
    bool marching = true, dense_pending = false;
    float dense_sigma = 0.0f;
#pragma unroll 1
    while (__any_sync(0xffffffffu, marching)) {
        if (marching && !dense_pending) {
#pragma unroll 1
            for (int k = 0; k < kMarchSteps; ++k) {
                if (cnt == 0) {  // the segment's for loop ended (count exhausted or `break`)
                    if (segment == 0 && dist1 > 0) {
                        segment = 1;
                        float d1 = fminf(dist1, b.uMaxRaymarchDistance);
                        cnt = uint32_t(fmaxf(b.uMaxRaymarchSteps * (d1 / b.uMaxRaymarchDistance), 1.0f));
                        ctx.step_size = d1 / float(cnt);
                        ctx.t = S.i1t1 + ctx.step_size * jitter;
                    } else {
                        marching = false;
                        break;
                    }
                }
                ctx.pos = camera + view_dir * ctx.t;  // UpdateContext, :90-93
                ctx.height01 = CalHeight01(P, ctx.pos);
                float sigma_t = SampleSigmaT<MAT, HW>(P.mat, ctx.pos, ctx.height01, COUNT ? &fetches : nullptr);  // :117
                if (COUNT) ++evals;
                if (!(sigma_t < 1e-5f)) {  // :118-119
                    dense_pending = true;
                    dense_sigma = sigma_t;
                    break;
                }
                if (ctx.transmittance < kMinTransmittance) cnt = 0;  // :173 / :185 `break`
                else { cnt--; ctx.t += ctx.step_size; }
            }
        }
        if (marching && dense_pending) {
            dense_pending = false;
            ShadeDenseStep<MAT, HW, COUNT>(P, ctx, dense_sigma, evals, fetches);  // :120-136
            if (ctx.transmittance < kMinTransmittance) cnt = 0;
            else { cnt--; ctx.t += ctx.step_size; }
        }
    }
    if (!S.valid) return;
    k16_ray_finish(P, S, ctx);
    if (COUNT) {  // counting variant is never the timed one
        atomicAdd(P.counters + SKY_CNT_RENDER_SIGMA_EVALS, (unsigned long long)evals);
        atomicAdd(P.counters + SKY_CNT_RENDER_TEX_FETCHES, (unsigned long long)fetches);
    }
}

#ifndef SKY_STRICT_TU
// ---- K16, ray-group wavefront (the production kernel) --------------------------------------------------------------------
// What bounds k16_render (one lane = one ray) is not a throughput but the LENGTH of a ray: a ray through thin cloud is up to 143
// steps, every step with cloud in it is 6 serial SampleSigmaT evaluations (the step + 5 shadow taps), and every evaluation is two
// dependent texture round trips: ~10^6 cycles of pure latency for one ray, i.e. the whole kernel time (measured: K14-K16 308 us at
// 1920x1080 against 433 us at 3840x2160 -- four times the rays, 1.4x the time; occupancy 5 -> 8 blocks/SM and 2..5 evaluations
// in flight per lane change nothing, profiles/k16_variants_r02d.log).  But within a ray only the running transmittance product
// and the sums it weights are sequential: sigma_t, the shadow march and the phase terms of a step depend on the step's position
// alone.  So here a warp works on a GROUP of 8 of its 32 rays at a time (the 8 lowest-numbered ones still marching), and its 32 lanes
// evaluate (group ray, look-ahead step) pairs side by side:
//   stage 1: the A <= 8 group rays x K = min(32 / A, kK16LookAhead) next steps of their segment -- one SampleSigmaT per lane;
//   stage 2: the D steps that found cloud x the shadow taps (tap-major, neighbouring lanes = neighbouring rays), dealt to all 32
//            lanes in ceil(D taps / 32) passes; each step's lane then adds its taps in the shader's order and forms the step-local
//            terms exp(-step sigma_t), sun / environment in-scatter of RayMarchStep (:120-133);
//   stage 3: each group ray's own lane consumes its K steps IN ORDER exactly like the shader's loop body (:134-137, :170-188): running
//            transmittance, sums, the `break` at kMinTransmittance (steps evaluated past it are discarded: the only speculation,
//            at most K - 1 steps once per segment), step count and t.
// A ray's operations are the shader's in the shader's order, the look-ahead t included (k additions of step_size).
// The critical path of a ray drops from 12 texture round trips per step to <= 3, and lanes are full while rays of a group remain.
// The strict objects and the counting variant keep k16_render, whose lanes follow the shader's loop literally.
#ifndef SKY_K16_LOOKAHEAD
#define SKY_K16_LOOKAHEAD 4   // measured at 4K, scene c3, hardware filtering: 2 -> 585 us, 4 -> 408 us, 8 -> 436 us (8 blocks / SM)
#endif
#ifndef SKY_K16_WAVE_RAYS
#define SKY_K16_WAVE_RAYS 8   // rays a warp owns: 8 (one group) or 32 (four groups, worked on one after the other)
#endif
#ifndef SKY_K16_WAVE_OCC
#define SKY_K16_WAVE_OCC 8    // resident 128-thread blocks per SM: 5 -> 480 us, 8 -> 408 us (64 registers, 20 bytes of spill)
#endif
// GR x LA = 32: a warp owns GR rays and looks LA steps ahead.  8 x 4 is the throughput shape (4K on one GPU: instruction issue
// bounds the kernel); 4 x 8 halves a ray's critical path -- twice the warps, half the round trips per ray -- and is the latency shape,
// chosen at launch when the rays of the launch cannot fill the machine twice over (a rank's bands of a sharded 4K frame).
constexpr int kK16MaxTaps = 8, kK16LookAhead = SKY_K16_LOOKAHEAD, kK16GroupRays = 8;
template <int GR, int LA>
struct K16WaveScratch {
    float4 dir[GR];                            // group ray: view_dir.xyz, back lobe
    float4 state[GR];                          // group ray: t, step_size, steps evaluated this round (uint bits), forward lobe
    float4 out[GR][LA];                        // per group ray and look-ahead step: (tr, sun, env, state); state 0 clear, 1 cloud
    float4 pos[32];                            // steps with cloud in them, by rank: position
    float res[32 * kK16MaxTaps];               // optical-depth terms of the shadow taps [tap * D + rank]
};

// number of taps SampleShadow's float loop (:103) makes, and their (mid point, delta_t)
inline int k16_shadow_taps(float shadow_steps, float* mid, float* weight, int cap) {
    float inv_shadow_steps = 1.0f / shadow_steps, previous_t = 0.0f;
    int n = 0;
    for (float t = inv_shadow_steps; t <= 1.0f && n < 4096; t += inv_shadow_steps) {
        float current_t = t * t;
        float delta_t = current_t - previous_t;
        if (n < cap) { mid[n] = previous_t + 0.5f * delta_t; weight[n] = delta_t; }
        ++n;
        previous_t = current_t;
    }
    return n;
}
struct K16Taps {
    int count;
    float mid[kK16MaxTaps], weight[kK16MaxTaps];
};

#ifdef SKY_K16_WAVE_STATS   // experiment builds only (tools/k16_stats.py): where the lane slots of the wavefront kernel go
__device__ unsigned long long g_k16_stats[8];
#define K16_STAT(slot, value) do { const unsigned long long v__ = (unsigned long long)(value); if (lane == 0) atomicAdd(&g_k16_stats[slot], v__); } while (0)
#else
#define K16_STAT(slot, value) do { } while (0)
#endif
template <int MAT, bool HW, int GR = kK16GroupRays, int LA = kK16LookAhead>
__global__ void __launch_bounds__(kK16Block, SKY_K16_WAVE_OCC * 128 / kK16Block) k16_render_wave(const __grid_constant__ CloudParams P, const __grid_constant__ K16Taps taps) {
    const SkyCloudCommonBufferData& c = P.c;
    const SkyCloudBufferData& b = P.b;
    __shared__ K16WaveScratch<GR, LA> scratch[kK16Block / 32];
    __shared__ float tap_mid[kK16MaxTaps], tap_weight[kK16MaxTaps];   // indexed per lane: shared memory, not the constant bank
    __shared__ unsigned int recip[33];   // ceil(2^16 / n): floor(t / n) = (t * recip[n]) >> 16 exactly for t < 256, n <= 32 (task -> (tap, step) without I2F / F2I)
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt = (1u << lane) - 1u;
    K16WaveScratch<GR, LA>& W = scratch[threadIdx.x >> 5];
    const int n_taps = taps.count;
    if (threadIdx.x < unsigned(kK16MaxTaps)) { tap_mid[threadIdx.x] = taps.mid[threadIdx.x]; tap_weight[threadIdx.x] = taps.weight[threadIdx.x]; }
    if (threadIdx.x < 33u) recip[threadIdx.x] = threadIdx.x ? (65536u + threadIdx.x - 1u) / threadIdx.x : 0u;
    __syncthreads();
    const float3 camera = f3(c.uCameraPos);
    const float3 sample_vector = b.uShadowDistance * f3(c.uSunDirection);

    const int warp_in_block = int(threadIdx.x >> 5);
#if SKY_K16_WAVE_RAYS == 32
    // ---- every lane sets up one ray (the warp's 8x4 tile, like k16_render); at any time the 8 lowest-numbered rays still marching
    //      form the group the warp works on
    const RaySetup S = k16_ray_setup(P, blockIdx.x * kK16TileW + (kK16Block >= 64 ? (warp_in_block & 1) * 8 : 0) + int(threadIdx.x & 7u),
                                     blockIdx.y * kK16TileH + (kK16Block >= 64 ? (warp_in_block >> 1) * 4 : 0) + int((threadIdx.x >> 3) & 3u));
#else
    // ---- a warp owns ONE group: 8 rays (a row of 8 quarter-res texels, the block's warps are consecutive rows), held by lanes 0..7.
    //      What a warp does one after the other is what decides the kernel's critical path, so it is one group, not four.
    RaySetup S{};
    if (lane < unsigned(GR)) S = k16_ray_setup(P, int(blockIdx.x) * GR + int(lane), int(blockIdx.y) * (kK16Block / 32) + warp_in_block);
#endif
    RayMarchContext ctx;
    ctx.cos_sun_view = S.cos_sun_view;
    float dist = S.i0t2 - S.i0t1;
    dist = fminf(dist, b.uMaxRaymarchDistance);
    uint32_t cnt = uint32_t(fmaxf(b.uMaxRaymarchSteps * (dist / b.uMaxRaymarchDistance), 1.0f));
    ctx.step_size = dist / float(cnt);
    ctx.transmittance = 1.0f;
    ctx.transmittance_sum = 0.0f;
    ctx.weighted_t_sum = 0.0f;
    ctx.sun_env = f2(0.0f, 0.0f);
    ctx.t = S.i0t1 + ctx.step_size * S.jitter;
    const float dist1 = S.i1t2 - S.i1t1;
    int segment = 0;
    bool marching = S.valid;
    // the two lobes of :125-127 depend on the ray only
    const float hg_back = HenyeyGreenstein(ctx.cos_sun_view, -0.15f) * 2.16f, hg_forward = HenyeyGreenstein(ctx.cos_sun_view, 0.85f);
#pragma unroll 1
    while (true) {
        // the segment's `for` ended (count exhausted or `break`): second shell segment (:178-188), or the ray is done
        if (marching && cnt == 0) {
            if (segment == 0 && dist1 > 0) {
                segment = 1;
                float d1 = fminf(dist1, b.uMaxRaymarchDistance);
                cnt = uint32_t(fmaxf(b.uMaxRaymarchSteps * (d1 / b.uMaxRaymarchDistance), 1.0f));
                ctx.step_size = d1 / float(cnt);
                ctx.t = S.i1t1 + ctx.step_size * S.jitter;
            } else {
                marching = false;
            }
        }
        const unsigned amask = __ballot_sync(0xffffffffu, marching);
        if (amask == 0) break;
        const int rank = __popc(amask & lt);
        const bool selected = marching && rank < GR;
        const int A = min(__popc(amask), GR);
        const int K = min(32 / A, LA);
        if (selected) {
            W.dir[rank] = f4(S.view_dir, hg_back);
            W.state[rank] = f4(ctx.t, ctx.step_size, __uint_as_float(min(cnt, uint32_t(K))), hg_forward);
        }
        __syncwarp();
        // ---- stage 1: lane -> (group ray j, look-ahead step k), step-major ------------------------------------------------------
        const int k = int((lane * recip[A]) >> 16);
        const int j = int(lane) - k * A;
        const bool has_task = k < K;
        float4 rd = f4(0.0f, 0.0f, 0.0f, 0.0f), rs = rd;
        if (has_task) { rd = W.dir[j]; rs = W.state[j]; }
        const bool valid = has_task && uint32_t(k) < __float_as_uint(rs.z);
        float sigma_t = 0.0f, height01 = 0.0f;
        float3 pos = camera;
        if (valid) {
            // t of look-ahead step k: the shader's k additions `t += step_size` (:176 / :187), not t + k * step_size -- the positions, and with
            // them every texel, are then the same whatever (rays per warp, look-ahead) shape the launch uses
            float tk = rs.x;
#pragma unroll
            for (int q = 1; q < LA; ++q) tk += q <= k ? rs.y : 0.0f;
            pos = camera + f3(rd.x, rd.y, rd.z) * tk;   // UpdateContext, :90-93
            height01 = CalHeight01(P, pos);
            sigma_t = SampleSigmaT<MAT, HW>(P.mat, pos, height01);               // :117
        }
        const bool cloud = valid && !(sigma_t < 1e-5f);                           // :118-119
        K16_STAT(0, 1); K16_STAT(1, A); K16_STAT(2, __popc(__ballot_sync(0xffffffffu, valid)));
        // ---- stage 2: the shadow marches of the steps with cloud in them (:98-114), all lanes --------------------------------------
        const unsigned dmask = __ballot_sync(0xffffffffu, cloud);
        unsigned cloud_rays = 0;                                                  // group rays with cloud in one of their K steps
        if (dmask != 0) {
            cloud_rays = __reduce_or_sync(0xffffffffu, cloud ? (1u << j) : 0u);
            const int D = __popc(dmask);
            const int d = __popc(dmask & lt);
            if (cloud) W.pos[d] = f4(pos, 0.0f);
            __syncwarp();
            const int total = D * n_taps;
            const unsigned int recip_D = recip[D];
            K16_STAT(3, D); K16_STAT(4, (total + 31) / 32);
#pragma unroll 1
            for (int task = int(lane); task < total; task += 32) {
                const int tap = int((unsigned(task) * recip_D) >> 16);
                const float4 p = W.pos[task - tap * D];
                const float3 sample_pos = f3(p.x, p.y, p.z) + sample_vector * tap_mid[tap];
                const float sample_height01 = CalHeight01(P, sample_pos);
                W.res[task] = SampleSigmaT<MAT, HW>(P.mat, sample_pos, sample_height01) * b.uShadowDistance * tap_weight[tap];
            }
            __syncwarp();
            float4 o = f4(1.0f, 0.0f, 0.0f, valid ? 0.0f : -1.0f);
            if (cloud) {
                float optical_depth;
                if (n_taps == 5) {   // uShadowSteps = 5 in every shipped configuration: no loop
                    optical_depth = (((W.res[d] + W.res[D + d]) + W.res[2 * D + d]) + W.res[3 * D + d]) + W.res[4 * D + d];
                } else {
                    optical_depth = 0.0f;
                    for (int tap = 0; tap < n_taps; ++tap) optical_depth += W.res[tap * D + d];   // the shader's order
                }
                const float transmittance_to_sun = expf(-optical_depth);
                // the step-local part of RayMarchStep, :120-133
                const float tr = expf(-rs.y * sigma_t);
                const float phase = mixf(rd.w, rs.w, expf(-b.uSunMultiscatteringSigmaScale * sigma_t));
                float2 sun_env;
                sun_env.x = transmittance_to_sun * phase;
                sun_env.y = mixf(b.uEnvBottomVisibility, 1.0f, height01);
                sun_env.y = sun_env.y - sun_env.y * expf(-b.uEnvMultiscatteringSigmaScale * sigma_t);
                sun_env = sun_env - sun_env * tr;
                o = f4(tr, sun_env.x, sun_env.y, 1.0f);
            }
            if (has_task && ((cloud_rays >> j) & 1u)) W.out[j][k] = o;
            __syncwarp();
        }
        // ---- stage 3: each group ray consumes its steps in order: the loop body of :170-177 / :182-187 ------------------------------
        if (selected) {
            const uint32_t n = min(cnt, uint32_t(K));            // steps of this ray that were evaluated
#ifdef SKY_K16_WAVE_STATS
            atomicAdd(&g_k16_stats[5], (unsigned long long)n);
#endif
            if (!((cloud_rays >> rank) & 1u)) {
                // clear air all the way: the transmittance does not change, so `break` fires after the first step or never
                if (ctx.transmittance < kMinTransmittance) cnt = 0;
                else {
                    cnt -= n;
#pragma unroll
                    for (int q = 0; q < LA; ++q) ctx.t += uint32_t(q) < n ? ctx.step_size : 0.0f;   // n times `t += step_size`
                }
            } else {
#pragma unroll 1
                for (uint32_t kk = 0; kk < n; ++kk) {
                    const float4 s = W.out[rank][kk];
                    if (s.w > 0.0f) {                           // :134-137
                        ctx.sun_env += ctx.transmittance * f2(s.y, s.z);
                        ctx.transmittance_sum += ctx.transmittance;
                        ctx.weighted_t_sum += ctx.t * ctx.transmittance;
                        ctx.transmittance *= s.x;
                    }
                    if (ctx.transmittance < kMinTransmittance) { cnt = 0; break; }   // :175 / :186 `break`
                    cnt--; ctx.t += ctx.step_size;
                }
            }
        }
        __syncwarp();
    }
    if (!S.valid) return;
    k16_ray_finish(P, S, ctx);
}
#endif  // SKY_STRICT_TU

// ------------------------------------------------------------------------------------------------ K17
// VolumetricCloudReconstruct.comp:28-110: reproject last frame's half-res image through the cloud
// distance, clamp to the depth-aware 3x3 neighbourhood AABB in Reinhard space, blend 0.2.
SKY_D float4 Reinhard(float4 v) { return f4(v.x / (1.0f + v.x), v.y / (1.0f + v.y), v.z / (1.0f + v.z), v.w); }
SKY_D float4 InverseReinhard(float4 v) { return f4(v.x / (1.0f - v.x), v.y / (1.0f - v.y), v.z / (1.0f - v.z), v.w); }

SKY_D float4 sample_half4_linear_clamp(const half4* p, int w, int h, float u, float v) {
    float x = u * float(w) - 0.5f, y = v * float(h) - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    int i0 = int(fminf(fmaxf(fx, -2.0f), float(w) + 1.0f)), j0 = int(fminf(fmaxf(fy, -2.0f), float(h) + 1.0f));
    int i1 = clampi(i0 + 1, 0, w - 1), j1 = clampi(j0 + 1, 0, h - 1);
    i0 = clampi(i0, 0, w - 1); j0 = clampi(j0, 0, h - 1);
    float4 t00 = load_half4(p + size_t(j0) * w + i0), t10 = load_half4(p + size_t(j0) * w + i1);
    float4 t01 = load_half4(p + size_t(j1) * w + i0), t11 = load_half4(p + size_t(j1) * w + i1);
    return (1.0f - a) * (1.0f - b) * t00 + a * (1.0f - b) * t10 + (1.0f - a) * b * t01 + a * b * t11;
}

// One thread per QUARTER-res cell = the 2x2 half-res pixels that share it: the nine-texel neighbourhood of the quarter-res render (loads,
// Reinhard), its nine linear depths, the cell's cloud distance and shading index are the same for all four pixels, which the shader's
// one-invocation-per-pixel mapping fetches and tone-maps four times (ncu profiles/k17_r02v.md: 663 instructions per pixel, ALU pipe 50 %,
// XU 32 %, 96 registers).  Every pixel still performs the shader's operations on the same operands in the same order.
__global__ void __launch_bounds__(128) k17_reconstruct(const __grid_constant__ CloudParams P) {
    const SkyCloudCommonBufferData& c = P.c;
    const int QW = P.width / 4, QH = P.height / 4, HW_ = P.width / 2, HH = P.height / 2;
    const int qx = blockIdx.x * 16 + (threadIdx.x & 15), qy = blockIdx.y * 8 + (threadIdx.x >> 4);
    if (2 * qx >= HW_ || 2 * qy >= HH) return;
    // kOffsets, :35
    const int ox[9] = {0, 0, 1, 1, 1, 0, -1, -1, -1}, oy[9] = {0, 1, 1, 0, -1, -1, -1, 0, 1};
    float rendered_linear_depths[9];
    float4 taps[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        // the two textureGathers + two texelFetchClamps of :37-49 read exactly these clamped texels
        int tx = clampi(qx + ox[i], 0, QW - 1), ty = clampi(qy + oy[i], 0, QH - 1);
        rendered_linear_depths[i] = __ldg(P.index_linear + size_t(ty) * QW + tx).y;
        taps[i] = Reinhard(load_half4(P.render + size_t(ty) * QW + tx));
    }
    const int cqx = clampi(qx, 0, QW - 1), cqy = clampi(qy, 0, QH - 1);
    const float rendered_distance = __ldg(P.cloud_distance + size_t(cqy) * QW + cqx);
    const int rendered_index = int(__ldg(P.index_linear + size_t(cqy) * QW + cqx).x);
    const int2 roff = IndexToOffset(uint32_t(rendered_index));
    const float3 camera = f3(c.uCameraPos);
#pragma unroll 1
    for (int sub = 0; sub < 4; ++sub) {
        const int x = 2 * qx + (sub & 1), y = 2 * qy + (sub >> 1);
        if (x >= HW_ || y >= HH) continue;
        float2 uv = f2((float(x) + 0.5f) / float(HW_), (float(y) + 0.5f) / float(HH));
        float depth = __ldg(P.checkerboard + size_t(y) * HW_ + x);
        float linear_depth = DepthToLinearDepth(c, depth);
        float delta_linear_depths[9];
        float min_delta_linear_depth = 1e10f;
        int nearest_i = 0;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            delta_linear_depths[i] = fabsf(rendered_linear_depths[i] - linear_depth);
            if (delta_linear_depths[i] < min_delta_linear_depth) {
                min_delta_linear_depth = delta_linear_depths[i];
                nearest_i = i;
            }
        }
        float4 rendered_nearest = taps[0];
#pragma unroll
        for (int i = 1; i < 9; ++i) if (i == nearest_i) rendered_nearest = taps[i];
        float4 aabb_min = rendered_nearest, aabb_max = rendered_nearest;
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            float4 rendered = taps[i];
            if (delta_linear_depths[i] < rendered_linear_depths[i] * 0.3f ||
                fabsf(rendered.w - rendered_nearest.w) / fmaxf(1e-6f, 1 - fmaxf(rendered.w, rendered_nearest.w)) < 0.2f) {
                aabb_min = min4(aabb_min, rendered);
                aabb_max = max4(aabb_max, rendered);
            }
        }
        float4 rendered = taps[0];
        float3 frag_pos = projective_mul(c.uInvMVP, f3(uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f, depth * 2.0f - 1.0f));
        float3 view_dir = normalize(frag_pos - camera);
        float3 cloud_pos = camera + view_dir * rendered_distance;
        float3 pre = projective_mul(c.uReprojectMat, cloud_pos);
        float2 pre_uv = f2(pre.x * 0.5f + 0.5f, pre.y * 0.5f + 0.5f);
        float4 pre_frame = Reinhard(sample_half4_linear_clamp(P.reconstruct_prev, HW_, HH, pre_uv.x, pre_uv.y));
        pre_frame = min4(max4(pre_frame, aabb_min), aabb_max);

        bool is_pre_out_of_screen = fmaxf(fabsf(pre.x), fabsf(pre.y)) > 1.0f;
        bool is_rendered = ((x & 1) == roff.x) && ((y & 1) == roff.y);
        float rendered_weight = is_pre_out_of_screen ? 1.0f : is_rendered ? 0.2f : 0.0f;
        float4 reconstructed = InverseReinhard(mix4(pre_frame, rendered, rendered_weight));
        P.reconstruct_out[size_t(y) * HW_ + x] = to_half4(reconstructed);
    }
}

// ------------------------------------------------------------------------------------------------ K18
// VolumetricCloudUpscale.comp:11-55: depth-aware 4-tap upscale and composite over the HDR target.
// HBM-bound: 4 B depth + 8 B hdr read + 8 B hdr write per full-res pixel.  (One thread per 2x2 pixel block -- 9 shared half-res texels
// instead of 16, 16-byte HDR accesses -- was measured and is slower, 65 vs 57 us at 4K: a quarter of the threads in flight.)
__global__ void __launch_bounds__(256) k18_upscale(const __grid_constant__ CloudParams P) {
    const SkyCloudCommonBufferData& c = P.c;
    const int HW_ = P.width / 2, HH = P.height / 2;
    int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (P.out_band_count > 1) y = ((y / P.out_band_rows) * P.out_band_count + P.out_band_index) * P.out_band_rows + y % P.out_band_rows;  // owned rows (multiples of 8)
    if (x >= P.width || y >= P.height) return;
    float depth = __ldg(P.depth + size_t(y) * P.width + x);
    float linear_depth = DepthToLinearDepth(c, depth);
    const int ox[4] = {-1, 1, 1, -1}, oy[4] = {1, 1, -1, -1};  // same order as textureGather
    int bx = (x - 1) >> 1, by = (y - 1) >> 1;
    float neighbor_depths[4] = {checker_clamped(P, bx, by + 1), checker_clamped(P, bx + 1, by + 1),
                                checker_clamped(P, bx + 1, by), checker_clamped(P, bx, by)};
    float4 reconstructed_neighbors[4];
    float min_delta_linear_depth = 1e10f;
    int nearest_i = 0;
    bool is_edge = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int hx = clampi((x + ox[i]) >> 1, 0, HW_ - 1), hy = clampi((y + oy[i]) >> 1, 0, HH - 1);
        reconstructed_neighbors[i] = load_half4(P.reconstruct_out + size_t(hy) * HW_ + hx);
        float neighbor_linear_depth = DepthToLinearDepth(c, neighbor_depths[i]);
        float delta_linear_depth = fabsf(linear_depth - neighbor_linear_depth);
        if (delta_linear_depth < min_delta_linear_depth) {
            nearest_i = i;
            min_delta_linear_depth = delta_linear_depth;
        }
        if (delta_linear_depth > linear_depth * 0.1f) is_edge = true;
    }
    float max_a = fmaxf(fmaxf(reconstructed_neighbors[0].w, reconstructed_neighbors[1].w), fmaxf(reconstructed_neighbors[2].w, reconstructed_neighbors[3].w));
    float min_a = fminf(fminf(reconstructed_neighbors[0].w, reconstructed_neighbors[1].w), fminf(reconstructed_neighbors[2].w, reconstructed_neighbors[3].w));
    float4 upscaled;
    if (is_edge && (max_a - min_a) / fmaxf(1e-6f, 1 - min_a) > 0.2f) {
        upscaled = reconstructed_neighbors[0];
#pragma unroll
        for (int i = 1; i < 4; ++i) if (i == nearest_i) upscaled = reconstructed_neighbors[i];
    } else {
        upscaled = ((reconstructed_neighbors[0] + reconstructed_neighbors[1]) + (reconstructed_neighbors[2] + reconstructed_neighbors[3])) * 0.25f;
    }
    float transmittance = upscaled.w;
    half4* dst = P.hdr + size_t(y) * P.width + x;
    float4 color = load_half4(dst);
    float k = transmittance <= kMinTransmittance ? 0.0f : transmittance;
    color.x = color.x * k + upscaled.x;
    color.y = color.y * k + upscaled.y;
    color.z = color.z * k + upscaled.z;
    const half4 texel = to_half4(color);
    *dst = texel;
    for (int k = 0; k < P.hdr_peer_count; ++k) P.hdr_peers[k][size_t(y) * P.width + x] = texel;   // 8 B per pixel and receiving rank over NVLink
}

// ------------------------------------------------------------------------------------------------ tex peak
// Texture-pipe roofline microbenchmark: every thread issues `iters` independent trilinear R8 fetches
// over the (L1/L2-resident) detail volume; mode 0 = coherent (neighbouring threads, neighbouring
// texels), mode 1 = incoherent (hashed coordinates).
__global__ void __launch_bounds__(256) k_tex_peak(cudaTextureObject_t tex, int iters, int mode, float* sink) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
    float u = float(tid & 127) * (1.0f / 128.0f), v = float((tid >> 7) & 127) * (1.0f / 128.0f), w = float((tid >> 14) & 127) * (1.0f / 128.0f);
    uint32_t s = tid * 747796405u + 2891336453u;
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            if (mode == 0) {
                u += 0.00390625f; w += 0.0009765625f;
            } else {
                s = s * 1664525u + 1013904223u;
                u = float(s >> 8) * (1.0f / 16777216.0f);
                v = float((s * 2654435761u) >> 8) * (1.0f / 16777216.0f);
                w = float((s * 40503u) >> 8) * (1.0f / 16777216.0f);
            }
            acc += tex3DLod<float>(tex, u, v, w, 0.0f);
        }
    }
    if (acc == -1.0f) sink[tid] = acc;
}

// Cross-GPU arrival barrier of a tile-sharded frame.  Runs after K16 on the same stream: thread k tells rank
// k "rank `rank` has stored all its rows of frame `epoch` into your buffers" and waits for rank k's own
// message.  Epochs only grow, so a late reader never sees a stale match.
// Two flag sets per rank: [0, 8) "rows of frame e arrived", [8, 16) "I finished reading frame e" -- the second
// one keeps a fast rank from overwriting a slow rank's copy while its K17 is still reading it.
constexpr unsigned long long kPeerWaitTimeoutNs = 10ull * 1000 * 1000 * 1000;  // 10 s
struct PeerBarrierParams {
    unsigned int* peer_flags[8];
    unsigned int* my_flags;
    int rank, world;
    unsigned int epoch;
    int offset;      // 0: K16 rows arrived, 8: K17 done reading them, 32: frame-target rows arrived, 40: previous frame target released
    int signal, wait;
    unsigned int signal_mask, wait_mask;   // ranks to signal / to wait for (bit k); 0 = all
};
__global__ void __launch_bounds__(32) k_peer_flags(const __grid_constant__ PeerBarrierParams P) {
    int k = threadIdx.x;
    if (k < P.world) {
        const bool do_signal = P.signal && (P.signal_mask == 0u || ((P.signal_mask >> k) & 1u));
        const bool do_wait = P.wait && (P.wait_mask == 0u || ((P.wait_mask >> k) & 1u));
        if (do_signal) {
            __threadfence_system();  // this stream's earlier kernels (K16's peer stores / K17's reads) first
            volatile unsigned int* theirs = P.peer_flags[k] + P.offset + P.rank;
            *theirs = P.epoch;
        }
        if (do_wait) {
            volatile unsigned int* mine = P.my_flags + P.offset + k;
            // bounded: a rank that died, skipped a frame or attached out of step must not hang this GPU's stream for ever.
            // After kPeerWaitTimeoutNs the frame goes on with whatever rows arrived and slot SKY_PEER_TIMEOUT_SLOT records the
            // peer; sky_sync / sky_read_resource turn it into an error.
            unsigned long long t0 = 0, now = 0;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
            while (*mine < P.epoch) {
                __nanosleep(200);
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                if (now - t0 > kPeerWaitTimeoutNs) { P.my_flags[SKY_PEER_TIMEOUT_SLOT] = 1u + (unsigned int)k; break; }
            }
            __threadfence_system();
        }
    }
}

// mode 2: coherent bilinear fetches of the RG8 weather map (the 2-D fetches of the default materials)
__global__ void __launch_bounds__(256) k_tex_peak2d(cudaTextureObject_t tex, int iters, float* sink) {
    uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    float acc = 0.0f;
    float u = float(tid & 511) * (1.0f / 512.0f), v = float((tid >> 9) & 511) * (1.0f / 512.0f);
    for (int i = 0; i < iters; i += 4) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            u += 0.0009765625f; v += 0.000244140625f;
            float4 t = tex2DLod<float4>(tex, u, v, 0.0f);
            acc += t.x + t.y;
        }
    }
    if (acc == -1.0f) sink[tid] = acc;
}

CloudParams make_cloud_params(SkyContext* ctx, const SkyCloudCommonBufferData& c) {
    CloudParams P{};
    P.c = c;
    P.inv_thickness = 1.0f / (c.uTopAltitude - c.uBottomAltitude);
    P.atm.u = ctx->atm;
    make_material_params(ctx, c.uCameraPos, P.mat);
    P.transmittance = LutView{ctx->transmittance.p, ctx->transmittance.w, ctx->transmittance.h, 1, 0};
    P.ap_lum = LutView{ctx->ap_lum.p, ctx->ap_lum.w, ctx->ap_lum.h, ctx->ap_lum.d, 0};
    P.ap_trans = LutView{ctx->ap_trans.p, ctx->ap_trans.w, ctx->ap_trans.h, ctx->ap_trans.d, 0};
    P.froxel = FroxelView{ctx->shadow_froxel.p, ctx->shadow_froxel.w, ctx->shadow_froxel.h, ctx->shadow_froxel.d};
    P.blue_noise = ctx->blue_noise;
    P.counters = ctx->counters;
    P.shadow_w = ctx->shadow_maps[0].w; P.shadow_h = ctx->shadow_maps[0].h;
    P.width = ctx->width; P.height = ctx->height;
    P.checkerboard = ctx->checkerboard_depth.p;
    P.index_linear = ctx->index_linear_depth.p;
    P.render = ctx->render_texture.p;
    P.cloud_distance = ctx->cloud_distance.p;
    return P;
}

}  // namespace

#ifndef SKY_STRICT_TU  // host-only helper, shared with the strict objects
int make_material_params(SkyContext* ctx, const float* camera_pos, MaterialParams& M) {
    M.m = ctx->material;
    M.cloud_map = ctx->cloud_map.view;
    M.detail = ctx->detail.view;
    M.displacement = ctx->displacement.view;
    M.voxel = ctx->voxel.view;
    M.camera_pos = f3(camera_pos);
    // lambda <= 0.5  <=>  k_lod * dist <= 2^(0.5 - bias)
    auto thr2 = [](float k_lod, float bias) {
        if (!(k_lod > 0.0f)) return INFINITY;  // log2(0 * dist) = -inf: always magnified
        float t = exp2f(0.5f - bias) / k_lod;
        return t * t;
    };
    const SkyMaterialCommonBufferData& mc = M.m.common;
    M.thr2_cloud_map = thr2(mc.uCloudMapSampleInfo.k_lod, mc.uLodBias);
    M.thr2_detail = thr2(mc.uDetailSampleInfo.k_lod, mc.uLodBias);
    M.thr2_displacement = thr2(mc.uDisplacementSampleInfo.k_lod, mc.uLodBias);
    M.thr2_voxel = thr2(M.m.u.voxel.uSampleLodK, M.m.u.voxel.uLodBias);
    // 0.5 - (log2(k_lod) + bias): lambda <= 0.5  <=>  lod_h - 0.5 * log2(dist^2) >= 0
    auto lod_h = [](float k_lod, float bias) { return k_lod > 0.0f ? 0.5f - (log2f(k_lod) + bias) : INFINITY; };
    M.lod_h_cloud_map = lod_h(mc.uCloudMapSampleInfo.k_lod, mc.uLodBias);
    M.lod_h_detail = lod_h(mc.uDetailSampleInfo.k_lod, mc.uLodBias);
    M.lod_h_displacement = lod_h(mc.uDisplacementSampleInfo.k_lod, mc.uLodBias);
    M.lod_h_voxel = lod_h(M.m.u.voxel.uSampleLodK, M.m.u.voxel.uLodBias);
    const float* dp = M.m.u.m0.uDetailParam;
    M.m0_zero_base_is_zero = M.m.type == SKY_MATERIAL_DEFAULT0 && dp[0] >= 0.0f && dp[1] >= 0.0f && dp[0] + dp[1] <= 1.0f;
    return 0;
}

#endif  // SKY_STRICT_TU

static int check_material_ready(SkyContext* ctx) {
    switch (ctx->material.type) {
        case SKY_MATERIAL_DEFAULT0:
            if (!ctx->displacement.valid) return sky_fail(ctx, "displacement texture has not been generated");
            // fall through
        case SKY_MATERIAL_DEFAULT1:
            if (!ctx->cloud_map.valid || !ctx->detail.valid) return sky_fail(ctx, "cloud map / detail texture has not been generated");
            return 0;
        case SKY_MATERIAL_MINIMAL: return 0;
        case SKY_MATERIAL_VOXEL:
            if (!ctx->voxel.valid) return sky_fail(ctx, "voxel grid has not been uploaded");
            return 0;
    }
    return sky_fail(ctx, "set_material has not been called with a known material type");
}

int launch_cloud_shadow(SkyContext* ctx, const SkyCloudCommonBufferData& c) {
    if (int e = check_material_ready(ctx)) return e;
    SKY_PERF_MARKER("VolumetricCloudShadow");  // VolumetricCloud.cpp:283
    std::swap(ctx->shadow_maps[0], ctx->shadow_maps[1]);  // VolumetricCloud.cpp:284
    CloudParams P = make_cloud_params(ctx, c);
    P.shadow_prev = ctx->shadow_maps[1].p;
    P.shadow_out = ctx->shadow_maps[0].p;
    P.shadow_blurred = ctx->shadow_maps[2].p;
    P.froxel_out = ctx->shadow_froxel.p;
    const bool count = ctx->counting;
    dim3 grid(ceil_div(P.shadow_w, 16), ceil_div(P.shadow_h, 8));
    int rc = dispatch_material(ctx->material.type, ctx->hw_filtering, [&]<int MAT, bool HW>() {
        SKY_PERF_MARKER("VolumetricCloudShadowMap");  // :288
        if (count) k11_shadow_map<MAT, HW, true><<<grid, 128, 0, ctx->stream>>>(P);
        else k11_shadow_map<MAT, HW, false><<<grid, 128, 0, ctx->stream>>>(P);
        return 0;
    });
    if (rc) return sky_fail(ctx, "unknown material");
    SKY_LAUNCH_CHECK(ctx);
    dim3 bgrid(ceil_div(P.shadow_w, 128), P.shadow_h);
    nvtxRangePushA("VolumetricCloudShadowMapBlur");  // :300
    k12_blur<true><<<bgrid, 128, 0, ctx->stream>>>(ctx->shadow_maps[0].p, ctx->shadow_maps[1].p, P.shadow_w, P.shadow_h);
    SKY_LAUNCH_CHECK(ctx);
    k12_blur<false><<<bgrid, 128, 0, ctx->stream>>>(ctx->shadow_maps[1].p, ctx->shadow_maps[2].p, P.shadow_w, P.shadow_h);
    nvtxRangePop();
    SKY_LAUNCH_CHECK(ctx);
    SKY_PERF_MARKER("VolumetricCloudShadowFroxel");  // :316
    if (P.froxel.d > kK13MaxDepth) return sky_fail(ctx, "shadow froxel volume deeper than 128 slices");
    k13_shadow_froxel<<<dim3(ceil_div(P.froxel.w, kK13Columns), P.froxel.h), kK13Slabs * kK13Columns, 0, ctx->stream>>>(P);
    SKY_LAUNCH_CHECK(ctx);
    return 0;
}

int launch_cloud_begin(SkyContext* ctx, const SkyCloudCommonBufferData& c, const SkyCloudBufferData& b, const float* depth,
                       int band_rows, int band_index, int band_count) {
    if (int e = check_material_ready(ctx)) return e;
    if (!ctx->ap_lum.p || !ctx->transmittance.p) return sky_fail(ctx, "atmosphere LUTs have not been baked");
    CloudParams P = make_cloud_params(ctx, c);
    P.b = b;
    P.depth = depth;
    const int QW = P.width / 4, QH = P.height / 4, HW_ = P.width / 2, HH = P.height / 2;
    SKY_PERF_MARKER("VolumetricCloud");  // VolumetricCloud.cpp:333
    {
        SKY_PERF_MARKER("Checkerboard Depth");  // :335
        k14_checkerboard<<<dim3(ceil_div(HW_, 256), HH), 256, 0, ctx->stream>>>(P);
        SKY_LAUNCH_CHECK(ctx);
    }
    {
        SKY_PERF_MARKER("Index Generate");  // :349
        k15_index_gen<<<dim3(ceil_div(QW, 128), QH), 128, 0, ctx->stream>>>(P);
        SKY_LAUNCH_CHECK(ctx);
    }
    SKY_PERF_MARKER("Render");  // :362
    int rows = QH;
    ctx->peer_band_frame = false;
    if (band_rows > 0 && band_count > 1) {
        if (ctx->peer_world == band_count && ctx->peer_rank == band_index && ctx->my_flags) {
            // peers attached: K16 stores its rows into every rank's copy (fused exchange)
            P.peer_count = ctx->peer_world;
            for (int k = 0; k < ctx->peer_world; ++k) { P.peer_render[k] = ctx->peer_render[k]; P.peer_distance[k] = ctx->peer_distance[k]; }
            ctx->peer_band_frame = true;
            if (ctx->peer_epoch > 0) {
                // nobody may still be reading the previous frame's exchanged buffers when K16 overwrites them
                PeerBarrierParams B{};
                for (int k = 0; k < ctx->peer_world; ++k) B.peer_flags[k] = ctx->peer_flags[k];
                B.my_flags = ctx->my_flags; B.rank = ctx->peer_rank; B.world = ctx->peer_world; B.epoch = ctx->peer_epoch;
                B.offset = 8; B.signal = 0; B.wait = 1;
                k_peer_flags<<<1, 32, 0, ctx->stream>>>(B);
                SKY_LAUNCH_CHECK(ctx);
                if (ctx->out_gather != SKY_GATHER_OFF) {
                    // this point follows, in stream order, everything the caller queued to read the previous frame's target: peers may
                    // overwrite their row bands in it from now on
                    B.offset = 40; B.signal = 1; B.wait = 0;
                    k_peer_flags<<<1, 32, 0, ctx->stream>>>(B);
                    SKY_LAUNCH_CHECK(ctx);
                }
            }
            ++ctx->peer_epoch;
        }
        P.band_rows = band_rows; P.band_index = band_index; P.band_count = band_count;
        int bands_total = ceil_div(QH, band_rows);
        int my_bands = (bands_total - band_index + band_count - 1) / band_count;
        rows = my_bands * band_rows;
    }
    dim3 grid(ceil_div(QW, kK16TileW), ceil_div(rows, kK16TileH));
    const bool count = ctx->counting;
#ifndef SKY_STRICT_TU
    K16Taps taps{};
    taps.count = k16_shadow_taps(b.uShadowSteps, taps.mid, taps.weight, kK16MaxTaps);
    // the ray-group wavefront kernel: production object, not counting, a shadow march it can deal out (1..8 taps)
    const bool wave = !count && !ctx->k16_literal && taps.count >= 1 && taps.count <= kK16MaxTaps;
    // the latency shape (4 rays x 8 steps per warp) when the launch has fewer 8-ray groups than twice the warps the machine holds (measured,
    // profiles/k16_group_r02r.log: a rank's bands of an 8-way sharded 4K frame 126 -> 104 us; a whole 1080p frame 135 -> 152 us, so not there)
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device);
    const long groups8 = long(ceil_div(QW, 8)) * rows;
    const bool narrow = SKY_K16_WAVE_RAYS != 32 && (ctx->k16_group == 4 || (ctx->k16_group == 0 && groups8 < 2l * sms * SKY_K16_WAVE_OCC * (128 / 32)));
    const int group_rays = narrow ? 4 : kK16GroupRays;
    const dim3 wave_grid = SKY_K16_WAVE_RAYS == 32 ? grid : dim3(ceil_div(QW, group_rays), ceil_div(rows, kK16Block / 32));
#endif
    int rc = dispatch_material(ctx->material.type, ctx->hw_filtering, [&]<int MAT, bool HW>() {
        if (count) k16_render<MAT, HW, true><<<grid, kK16Block, 0, ctx->stream>>>(P);
#ifndef SKY_STRICT_TU
        else if (wave && narrow) k16_render_wave<MAT, HW, 4, 8><<<wave_grid, kK16Block, 0, ctx->stream>>>(P, taps);
        else if (wave) k16_render_wave<MAT, HW><<<wave_grid, kK16Block, 0, ctx->stream>>>(P, taps);
#endif
        else k16_render<MAT, HW, false><<<grid, kK16Block, 0, ctx->stream>>>(P);
        return 0;
    });
    if (rc) return sky_fail(ctx, "unknown material");
    SKY_LAUNCH_CHECK(ctx);
    return 0;
}

int launch_cloud_end(SkyContext* ctx, const SkyCloudCommonBufferData& c, const float* depth, half4* hdr, int phases) {
    CloudParams P = make_cloud_params(ctx, c);
    P.depth = depth;
    P.hdr = hdr;
    P.reconstruct_out = ctx->reconstruct[0].p;
    P.reconstruct_prev = ctx->reconstruct[1].p;
    const int HW_ = P.width / 2, HH = P.height / 2;
    SKY_PERF_MARKER("VolumetricCloud");
    if (phases & 1) {
        SKY_PERF_MARKER("Reconstruct");  // VolumetricCloud.cpp:388
        const bool peer_frame = ctx->peer_band_frame;
        PeerBarrierParams B{};
        if (peer_frame) {
            for (int k = 0; k < ctx->peer_world; ++k) B.peer_flags[k] = ctx->peer_flags[k];
            B.my_flags = ctx->my_flags; B.rank = ctx->peer_rank; B.world = ctx->peer_world; B.epoch = ctx->peer_epoch;
            B.offset = 0; B.signal = 1; B.wait = 1;  // my rows are everywhere; wait for everybody else's
            k_peer_flags<<<1, 32, 0, ctx->stream>>>(B);
            SKY_LAUNCH_CHECK(ctx);
            ctx->peer_band_frame = false;
        }
        k17_reconstruct<<<dim3(ceil_div(ceil_div(HW_, 2), 16), ceil_div(ceil_div(HH, 2), 8)), 128, 0, ctx->stream>>>(P);
        if (peer_frame) {
            B.offset = 8; B.signal = 1; B.wait = 0;  // K17 was the last reader of the exchanged buffers
            k_peer_flags<<<1, 32, 0, ctx->stream>>>(B);
            SKY_LAUNCH_CHECK(ctx);
        }
        SKY_LAUNCH_CHECK(ctx);
    }
    if (phases & 2) {
        SKY_PERF_MARKER("Upscale");  // VolumetricCloud.cpp:407
        P.out_band_rows = ctx->out_band_rows; P.out_band_index = ctx->out_band_index; P.out_band_count = ctx->out_band_count;
        const int rows = owned_rows(ctx, P.height);
        // frame target in peer memory (sky_set_output_gather): K18 is the last writer of a frame, its stores go to the receiving ranks too
        const bool gather = ctx->out_gather != SKY_GATHER_OFF && ctx->peer_world > 1 && ctx->out_band_count == ctx->peer_world && ctx->out_band_index == ctx->peer_rank &&
                            ctx->frame_hdr.p && hdr == ctx->frame_hdr.p && ctx->my_flags && ctx->peer_epoch > 0;
        PeerBarrierParams B{};
        if (gather) {
            const bool to_all = ctx->out_gather == SKY_GATHER_ALL;
            for (int k = 0; k < ctx->peer_world; ++k) {
                B.peer_flags[k] = ctx->peer_flags[k];
                if (k != ctx->peer_rank && (to_all || k == 0)) P.hdr_peers[P.hdr_peer_count++] = ctx->peer_hdr[k];
            }
            B.my_flags = ctx->my_flags; B.rank = ctx->peer_rank; B.world = ctx->peer_world;
            const unsigned int receivers = to_all ? 0u : 1u;   // mask of the ranks that receive (0 = all)
            if (P.hdr_peer_count > 0 && ctx->peer_epoch > 1) {
                // the receivers must have released the previous frame's target (flag value: the last frame they are done with)
                B.epoch = ctx->peer_epoch - 1; B.offset = 40; B.signal = 0; B.wait = 1; B.wait_mask = receivers;
                k_peer_flags<<<1, 32, 0, ctx->stream>>>(B);
                SKY_LAUNCH_CHECK(ctx);
            }
        }
        if (rows > 0) k18_upscale<<<dim3(ceil_div(P.width, 32), ceil_div(rows, 8)), 256, 0, ctx->stream>>>(P);
        SKY_LAUNCH_CHECK(ctx);
        if (gather) {
            const bool to_all = ctx->out_gather == SKY_GATHER_ALL;
            const bool receive = to_all || ctx->peer_rank == 0;
            B.epoch = ctx->peer_epoch; B.offset = 32; B.signal = 1; B.signal_mask = to_all ? 0u : 1u; B.wait = receive ? 1 : 0; B.wait_mask = 0u;
            k_peer_flags<<<1, 32, 0, ctx->stream>>>(B);   // my rows are in every receiver's target; a receiver waits for everybody's
            SKY_LAUNCH_CHECK(ctx);
        }
        std::swap(ctx->reconstruct[0], ctx->reconstruct[1]);  // VolumetricCloud.cpp:421-422
    }
    return 0;
}

#ifndef SKY_STRICT_TU
int launch_tex_peak(SkyContext* ctx, int mode, double* fetches_per_second) {
    if (!ctx->detail.valid) return sky_fail(ctx, "tex_peak needs the detail volume (noise_generate(SKY_NOISE_DETAIL))");
    const int iters = 1024, blocks = 148 * 16, threads = 256;
    cudaEvent_t e0, e1;
    SKY_CUDA(ctx, cudaEventCreate(&e0));
    SKY_CUDA(ctx, cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        SKY_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        if (mode == 2) k_tex_peak2d<<<blocks, threads, 0, ctx->stream>>>(ctx->cloud_map.view.tex_linear, iters, nullptr);
        else k_tex_peak<<<blocks, threads, 0, ctx->stream>>>(ctx->detail.view.tex_linear, iters, mode, nullptr);
        SKY_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        SKY_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0.0f;
        SKY_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *fetches_per_second = double(blocks) * threads * iters / (double(best) * 1e-3);
    return 0;
}
#endif  // SKY_STRICT_TU

#if defined(SKY_K16_WAVE_STATS) && !defined(SKY_STRICT_TU)
extern "C" __attribute__((visibility("default"))) int sky_debug_k16_stats(unsigned long long* out, int reset) {
    if (out && cudaMemcpyFromSymbol(out, g_k16_stats, sizeof(g_k16_stats)) != cudaSuccess) return 1;
    if (reset) { unsigned long long z[8] = {}; if (cudaMemcpyToSymbol(g_k16_stats, z, sizeof(z)) != cudaSuccess) return 1; }
    return 0;
}
#endif
