// K7, the analytic ground pass: shaders/SkyRendering/EarthRender.frag driven by Earth::RenderToGBuffer
// (src/SkyRendering/Earth.cpp:46-65), and the earth albedo map it samples (Textures::Textures, src/Base/src/Textures.cpp:52-58:
// GL_SRGB8 upload + glGenerateTextureMipmap; sampler Earth.cpp:34-42: REPEAT x CLAMP_TO_EDGE, LINEAR_MIPMAP_LINEAR, maximum anisotropy).
//
// Compiled with -fmad=false like the LUT bake: IEEE division / sqrt, acos / atan / log2 from include/sky_detmath.h, the shader's
// operation order, and the textureGrad rule of include/sky_texgrad.h (shared with the oracle and the reference-shader shim), so depth,
// normal, ORM, albedo and every mip level of the map are bit-identical to the oracle (tests/test_gpu_earth.py).
//
// Mapping:
//   * k7_earth_gbuffer: a fragment program takes screen-space derivatives inside 2x2 pixel quads, so a thread block is 16x8 pixels
//     laid out quad by quad -- the four pixels of a quad are four neighbouring lanes and dFdx / dFdy are two __shfl_sync each.  Lanes
//     whose pixel is discarded (or lies past the image edge) keep running as helper invocations until the derivatives are taken,
//     exactly like the rasteriser's helper pixels.  Texels are fetched as RGBX sRGB codes (one 4-byte load per texel) and decoded
//     through a 256-entry table in shared memory before filtering (GL 4.6 section 8.24).
//   * k7_albedo_mip: one level per launch, 2x2 box on the decoded values, re-encoded by a binary search over the 255 code thresholds
//     of the sRGB transfer function (the host derives them from the oracle's formula by bisection, so the codes are bit-identical).
#include "context.h"
#include "../../include/sky_detmath.h"
#include "../../include/sky_texgrad.h"
#include "atmosphere_dev.cuh"

namespace {

struct AlbedoView {
    const uchar4* texels;                 // all levels, level l at texels + off[l]
    unsigned long long off[kEarthMaxLevels];
    int w, h, levels;
};

struct EarthParams {
    SkyEarthBufferData e;
    float bottom_radius;
    AlbedoView map;
    const float* srgb_decode;             // 256 entries
    float* depth;                         // in / out
    uchar4* albedo;
    short4* normal;
    ushort4* orm;
    int width, height;
};

struct GroundSample {
    bool keep;
    float3 ground_position;
    float2 coord;
};

// EarthRender.frag:40-52 and GetEarthAlbedo :22-26 for the pixel (px, py); helper pixels past the image edge fetch the edge depth
SKY_D GroundSample ground_sample(const EarthParams& P, int px, int py) {
    const float3 camera_position = f3(P.e.camera_position), earth_center = f3(P.e.earth_center), up_direction = f3(P.e.up_direction);
    const float u = (float(px) + 0.5f) / float(P.width), v = (float(py) + 0.5f) / float(P.height);
    const float d = __ldg(P.depth + size_t(clampi(py, 0, P.height - 1)) * P.width + clampi(px, 0, P.width - 1));
    const float3 fragment_position = projective_mul(P.e.inv_view_projection, f3(u, v, d) * 2.0f - f3(1.0f));
    const float3 view_direction = normalize(fragment_position - camera_position);
    const float r = P.e.camera_earth_center_distance;
    const float mu = dot(view_direction, up_direction);
    GroundSample g;
    g.keep = mu < 0.0f && r * r * (mu * mu - 1.0f) + P.bottom_radius * P.bottom_radius >= 0.0f;       // RayIntersectsGround
    const float discriminant = r * r * (mu * mu - 1.0f) + P.bottom_radius * P.bottom_radius;
    const float dist = fmaxf(-r * mu - sqrtf(fmaxf(discriminant, 0.0f)), 0.0f);                       // DistanceToBottomAtmosphereBoundary
    if (dist >= distance(fragment_position, camera_position)) g.keep = false;
    g.ground_position = camera_position + view_direction * dist;
    const float3 direction = normalize(g.ground_position - earth_center);
    const float theta = sky_det_acosf(direction.y);
    const float phi = sky_det_atan2f(direction.x, direction.z);
    g.coord = f2(kInvPi * 0.5f * phi + 0.5f, 1.0f - theta * kInvPi);
    return g;
}

__global__ void __launch_bounds__(128) k7_earth_gbuffer(const __grid_constant__ EarthParams P) {
    __shared__ float decode[256];
    if (P.map.texels) { decode[threadIdx.x] = __ldg(P.srgb_decode + threadIdx.x); decode[threadIdx.x + 128] = __ldg(P.srgb_decode + threadIdx.x + 128); }
    __syncthreads();
    const int quad = threadIdx.x >> 2, sub = threadIdx.x & 3;
    const int px = blockIdx.x * 16 + (quad & 7) * 2 + (sub & 1), py = blockIdx.y * 8 + (quad >> 3) * 2 + (sub >> 1);
    // the depth plane is read by neighbouring quads' helper lanes while kept pixels write it: all reads of the block's footprint
    // happen in ground_sample, before any write below, and no other block reads this block's pixels (quads never straddle blocks)
    const GroundSample g = ground_sample(P, px, py);
    // fine derivatives inside the quad (:27-35): value at (x | 1) minus value at (x & ~1), same for y
    const unsigned lane = threadIdx.x & 31u, full = 0xffffffffu;
    const unsigned x0 = lane & ~1u, x1 = lane | 1u, y0 = lane & ~2u, y1 = lane | 2u;
    const float cxf = fractf(g.coord.x + 0.5f);
    const float dudx1 = __shfl_sync(full, g.coord.x, x1) - __shfl_sync(full, g.coord.x, x0);
    const float dudy1 = __shfl_sync(full, g.coord.x, y1) - __shfl_sync(full, g.coord.x, y0);
    const float dudx2 = __shfl_sync(full, cxf, x1) - __shfl_sync(full, cxf, x0);
    const float dudy2 = __shfl_sync(full, cxf, y1) - __shfl_sync(full, cxf, y0);
    const float dvdx = __shfl_sync(full, g.coord.y, x1) - __shfl_sync(full, g.coord.y, x0);
    const float dvdy = __shfl_sync(full, g.coord.y, y1) - __shfl_sync(full, g.coord.y, y0);
    if (!g.keep || px >= P.width || py >= P.height) return;   // discard / helper pixel: nothing is written
    const size_t o = size_t(py) * P.width + px;
    const float z = projective_mul(P.e.view_projection, g.ground_position).z * 0.5f + 0.5f;            // gl_FragDepth, :49
    // D24: round to nearest of z (2^24 - 1); the product and the quotient are exact in double like the oracle's
    P.depth[o] = float(floor(double(fminf(fmaxf(z, 0.0f), 1.0f)) * 16777215.0 + 0.5) / 16777215.0);
    float3 color = f3(0.0f);
    if (P.map.texels) {
        const bool first = sqrtf(dudx1 * dudx1 + dudy1 * dudy1) < sqrtf(dudx2 * dudx2 + dudy2 * dudy2);   // make the Earth seamless
        const float dudx = first ? dudx1 : dudx2, dudy = first ? dudy1 : dudy2;
        color = sky_texture_grad_2d<float3>(P.map.w, P.map.h, P.map.levels, g.coord.x, g.coord.y, dudx, dvdx, dudy, dvdy, 16.0f,
                                            [&](int l, int i, int j) {
                                                const int wl = max(P.map.w >> l, 1);
                                                const uchar4 c = __ldg(P.map.texels + P.map.off[l] + size_t(j) * wl + i);
                                                return f3(decode[c.x], decode[c.y], decode[c.z]);
                                            });
    }
    const float3 n = normalize(g.ground_position - f3(P.e.earth_center));
    auto unorm8 = [](float v) { return (unsigned char)rintf(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f); };
    auto snorm16 = [](float v) { return (short)rintf(fminf(fmaxf(v, -1.0f), 1.0f) * 32767.0f); };
    P.albedo[o] = make_uchar4(unorm8(color.x), unorm8(color.y), unorm8(color.z), 255);
    P.normal[o] = make_short4(snorm16(n.x), snorm16(n.y), snorm16(n.z), 32767);
    P.orm[o] = make_ushort4(65535, 65535, 0, 65535);                                                   // (1, roughness 1, metallic 0, 1)
}

// glGenerateTextureMipmap of a GL_SRGB8 texture, one level: 2x2 box of the decoded texels (floor sizes: an odd source dimension
// clamps its last column / row), re-encoded with round to nearest
__global__ void k7_albedo_mip(const uchar4* src, int sw, int sh, uchar4* dst, int dw, int dh, const float* srgb_decode, const float* thresholds) {
    __shared__ float decode[256], thr[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { decode[i] = __ldg(srgb_decode + i); thr[i] = __ldg(thresholds + i); }
    __syncthreads();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= dw * dh) return;
    const int x = i % dw, y = i / dw;
    const int x0 = min(2 * x, sw - 1), x1 = min(2 * x + 1, sw - 1), y0 = min(2 * y, sh - 1), y1 = min(2 * y + 1, sh - 1);
    const uchar4 t00 = __ldg(src + size_t(y0) * sw + x0), t10 = __ldg(src + size_t(y0) * sw + x1);
    const uchar4 t01 = __ldg(src + size_t(y1) * sw + x0), t11 = __ldg(src + size_t(y1) * sw + x1);
    auto encode = [&](float v) {
        // thr[c] (c = 1..255) is the smallest fp32 value whose sRGB code is >= c; thr[0] = -inf: the code is the last c with thr[c] <= v
        int lo = 0, hi = 255;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (thr[mid] <= v) lo = mid; else hi = mid - 1;
        }
        return (unsigned char)lo;
    };
    auto box = [&](unsigned char a, unsigned char b, unsigned char c, unsigned char d) { return encode(((decode[a] + decode[b]) + (decode[c] + decode[d])) * 0.25f); };
    dst[i] = make_uchar4(box(t00.x, t10.x, t01.x, t11.x), box(t00.y, t10.y, t01.y, t11.y), box(t00.z, t10.z, t01.z, t11.z), 255);
}

// Clear(const GBuffer&), GBuffer.h:28-34: one thread per pixel pair, 16-byte stores
__global__ void __launch_bounds__(256) k_gbuffer_clear(float2* depth, uint2* albedo, uint4* normal, uint4* orm, size_t pairs, float* depth_tail, unsigned* albedo_tail,
                                                       uint2* normal_tail, uint2* orm_tail) {
    const size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < pairs) {
        depth[i] = make_float2(1.0f, 1.0f);
        albedo[i] = make_uint2(0u, 0u);
        normal[i] = make_uint4(0u, 0u, 0u, 0u);
        orm[i] = make_uint4(0u, 0u, 0u, 0u);
    } else if (i == pairs && depth_tail) {   // odd pixel count
        *depth_tail = 1.0f; *albedo_tail = 0u; *normal_tail = make_uint2(0u, 0u); *orm_tail = make_uint2(0u, 0u);
    }
}

}  // namespace

int launch_gbuffer_clear(SkyContext* ctx, float* depth, void* albedo, void* normal, void* orm, int width, int height) {
    const size_t n = size_t(width) * height, pairs = n / 2;
    const bool odd = (n & 1) != 0;
    SKY_PERF_MARKER("Clear GBuffer");
    k_gbuffer_clear<<<unsigned((pairs + 1 + 255) / 256), 256, 0, ctx->stream>>>(
        reinterpret_cast<float2*>(depth), static_cast<uint2*>(albedo), static_cast<uint4*>(normal), static_cast<uint4*>(orm), pairs,
        odd ? depth + (n - 1) : nullptr, odd ? static_cast<unsigned*>(albedo) + (n - 1) : nullptr, odd ? static_cast<uint2*>(normal) + (n - 1) : nullptr,
        odd ? static_cast<uint2*>(orm) + (n - 1) : nullptr);
    SKY_LAUNCH_CHECK(ctx);
    return 0;
}

int launch_earth_albedo_mips(SkyContext* ctx, const float* thresholds_dev) {
    int w = ctx->earth_w, h = ctx->earth_h;
    for (int l = 1; l < ctx->earth_levels; ++l) {
        const int nw = std::max(w / 2, 1), nh = std::max(h / 2, 1);
        k7_albedo_mip<<<ceil_div(nw * nh, 256), 256, 0, ctx->stream>>>(ctx->earth_albedo + ctx->earth_off[l - 1], w, h, ctx->earth_albedo + ctx->earth_off[l], nw, nh,
                                                                       ctx->srgb_decode, thresholds_dev);
        SKY_LAUNCH_CHECK(ctx);
        w = nw; h = nh;
    }
    return 0;
}

int launch_earth_gbuffer(SkyContext* ctx, const SkyEarthBufferData& e, float* depth, void* albedo, void* normal, void* orm, int width, int height) {
    EarthParams P{};
    P.e = e;
    P.bottom_radius = ctx->atm.bottom_radius;
    P.map.texels = ctx->earth_albedo; P.map.w = ctx->earth_w; P.map.h = ctx->earth_h; P.map.levels = ctx->earth_levels;
    for (int l = 0; l < kEarthMaxLevels; ++l) P.map.off[l] = ctx->earth_off[l];
    P.srgb_decode = ctx->srgb_decode;
    P.depth = depth; P.albedo = static_cast<uchar4*>(albedo); P.normal = static_cast<short4*>(normal); P.orm = static_cast<ushort4*>(orm);
    P.width = width; P.height = height;
    SKY_PERF_MARKER("Earth GBuffer");
    k7_earth_gbuffer<<<dim3(ceil_div(width, 16), ceil_div(height, 8)), 128, 0, ctx->stream>>>(P);
    SKY_LAUNCH_CHECK(ctx);
    return 0;
}
