// K8-K10: tileable Perlin / Worley FBM generation (shaders/Base/Noise.glsl,
// shaders/SkyRendering/NoiseGen.comp) and the glGenerateTextureMipmap replacement, sm_100a.
// ALU-bound uint32 hashing; one thread per texel, hash chains hoisted out of the 27-cell /
// 8-corner loops (the nested WangHash(x + WangHash(y + WangHash(z))) shares its inner terms).
// Compiled with -fmad=false so the 8-bit outputs match the unfused reference arithmetic bit for bit.
#include "context.h"

namespace {

// Noise.glsl:1-8
SKY_D uint32_t WangHash(uint32_t seed) {
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}

// Noise.glsl:17-22, packed as 2-bit codes would save nothing: keep the table in constant memory
__constant__ float kPerlinGradients[16][3] = {
    {1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1}, {-1, 0, -1},
    {0, 1, 1}, {0, -1, 1}, {0, 1, -1}, {0, -1, -1}, {1, 1, 0}, {-1, 1, 0}, {0, -1, 1}, {0, -1, -1}};

SKY_D float grad_dot(uint32_t h, float x, float y, float z) {
    const float* g = kPerlinGradients[h & 0xf];
    return g[0] * x + g[1] * y + g[2] * z;  // dot(), left to right like GLSL
}

// Noise.glsl:28-60
SKY_D float PerlinNoise(float3 p, uint32_t freq, uint32_t seed) {
    p = p * float(freq);
    float3 fl = f3(floorf(p.x), floorf(p.y), floorf(p.z));
    uint32_t i0 = uint32_t(int(fl.x)) % freq, j0 = uint32_t(int(fl.y)) % freq, k0 = uint32_t(int(fl.z)) % freq;
    uint32_t i1 = uint32_t(int(ceilf(p.x))) % freq, j1 = uint32_t(int(ceilf(p.y))) % freq, k1 = uint32_t(int(ceilf(p.z))) % freq;
    float3 t = p - fl;
    float u = t.x * t.x * t.x * (t.x * (t.x * 6.0f - 15.0f) + 10.0f);
    float v = t.y * t.y * t.y * (t.y * (t.y * 6.0f - 15.0f) + 10.0f);
    float w = t.z * t.z * t.z * (t.z * (t.z * 6.0f - 15.0f) + 10.0f);
    float x0 = t.x, y0 = t.y, z0 = t.z, x1 = t.x - 1.0f, y1 = t.y - 1.0f, z1 = t.z - 1.0f;
    // GetPerlinGradients(i,j,k) = table[Wang(seed + Wang(i + Wang(j + Wang(k)))) & 15], :24-26
    uint32_t hk0 = WangHash(k0), hk1 = WangHash(k1);
    uint32_t h00 = WangHash(j0 + hk0), h10 = WangHash(j1 + hk0), h01 = WangHash(j0 + hk1), h11 = WangHash(j1 + hk1);
#define SKY_G(i, hjk) WangHash(seed + WangHash((i) + (hjk)))
    float n000 = grad_dot(SKY_G(i0, h00), x0, y0, z0), n100 = grad_dot(SKY_G(i1, h00), x1, y0, z0);
    float n010 = grad_dot(SKY_G(i0, h10), x0, y1, z0), n110 = grad_dot(SKY_G(i1, h10), x1, y1, z0);
    float n001 = grad_dot(SKY_G(i0, h01), x0, y0, z1), n101 = grad_dot(SKY_G(i1, h01), x1, y0, z1);
    float n011 = grad_dot(SKY_G(i0, h11), x0, y1, z1), n111 = grad_dot(SKY_G(i1, h11), x1, y1, z1);
#undef SKY_G
    return mixf(mixf(mixf(n000, n100, u), mixf(n010, n110, u), v), mixf(mixf(n001, n101, u), mixf(n011, n111, u), v), w);
}

// Noise.glsl:82-101 (the vec3 overload; NoiseGen.comp's WorleyFBM always calls it, also for the 2-D maps)
SKY_D float WorleyNoise(float3 p, uint32_t freq, uint32_t seed) {
    p = p * float(freq);
    uint32_t ix = uint32_t(floorf(p.x)), iy = uint32_t(floorf(p.y)), iz = uint32_t(floorf(p.z));
    p = p + f3(float(freq));
    float min_dist = 1e10f;
#pragma unroll
    for (uint32_t dk = 0; dk < 3; ++dk) {
        uint32_t gz = iz + (freq - 1 + dk);
        uint32_t sz = gz % freq;
        uint32_t hz0 = WangHash(sz), hz1 = WangHash(sz + 1), hz2 = WangHash(sz + 2);
#pragma unroll
        for (uint32_t dj = 0; dj < 3; ++dj) {
            uint32_t gy = iy + (freq - 1 + dj);
            uint32_t sy = gy % freq;
            uint32_t hy0 = WangHash(sy + hz0), hy1 = WangHash(sy + hz1), hy2 = WangHash(sy + hz2);
#pragma unroll
            for (uint32_t di = 0; di < 3; ++di) {
                uint32_t gx = ix + (freq - 1 + di);
                uint32_t sx = gx % freq;
                uint32_t rnd0 = WangHash(seed + WangHash(sx + hy0));
                uint32_t rnd1 = WangHash(seed + WangHash(sx + hy1));
                uint32_t rnd2 = WangHash(seed + WangHash(sx + hy2));
                float3 g = f3(float(gx), float(gy), float(gz)) + f3(float(rnd0), float(rnd1), float(rnd2)) / 4294967296.0f;
                min_dist = fminf(min_dist, distance(p, g));
            }
        }
    }
    return min_dist;
}

// NoiseGen.comp:12-18
SKY_D float RemapTo01(float x, float x0, float x1) { return clampf((x - x0) / (x1 - x0), 0.0f, 1.0f); }
SKY_D float RemapFrom01(float x, float y0, float y1) { return clampf(y0 + x * (y1 - y0), 0.0f, 1.0f); }

// NoiseGen.comp:20-50
template <bool WORLEY>
SKY_D float FBM(float3 p, SkyNoiseCreateInfo ci) {
    float res = 0.0f;
    uint32_t f = ci.base_frequency;
    float a = 0.5f, sum_a = 0.0f;
    for (uint32_t c = 0; c < 8; ++c) {
        float noise = WORLEY ? WorleyNoise(p, f, ci.seed) : PerlinNoise(p, f, ci.seed) * 0.5f + 0.5f;
        res += RemapTo01(noise, ci.remap_min, ci.remap_max) * a;
        sum_a += a;
        f *= 2;
        a *= 0.5f;
    }
    return res / sum_a;
}

SKY_D uint8_t unorm8(float x) { return uint8_t(__float2int_rn(clampf(x, 0.0f, 1.0f) * 255.0f)); }  // round-to-nearest-even

struct NoiseParams {
    SkyNoiseCreateInfo a, b;
    uint8_t* out;
    int size;
};

// K8 -- NoiseGen.comp:72-88: RG8 weather map
__global__ void __launch_bounds__(128) k8_cloud_map(const __grid_constant__ NoiseParams P) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.size) return;
    float3 coord = f3((float(x) + 0.5f) / float(P.size), (float(y) + 0.5f) / float(P.size), 0.0f);
    float density = FBM<false>(coord, P.a);
    float height = FBM<true>(coord, P.b);
    uchar2 v = make_uchar2(unorm8(density), unorm8(height));
    reinterpret_cast<uchar2*>(P.out)[y * P.size + x] = v;
}

// K9 -- NoiseGen.comp:90-106: R8 Perlin-Worley volume.  ~3e9 hash rounds for 128^3; each thread is
// independent, a 128-thread block covers one x-row so stores coalesce.
__global__ void __launch_bounds__(128) k9_detail(const __grid_constant__ NoiseParams P) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, z = blockIdx.z;
    if (x >= P.size) return;
    float s = float(P.size);
    float3 coord = f3((float(x) + 0.5f) / s, (float(y) + 0.5f) / s, (float(z) + 0.5f) / s);
    float perlin = FBM<false>(coord, P.a);
    float worley = FBM<true>(coord, P.b);
    P.out[(size_t(z) * P.size + y) * P.size + x] = unorm8(RemapFrom01(perlin, worley, 1.0f));
}

// K10 -- NoiseGen.comp:52-70: RGBA8 displacement map, 4 Perlin FBMs with seed+i
__global__ void __launch_bounds__(128) k10_displacement(const __grid_constant__ NoiseParams P) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.size) return;
    float3 coord = f3((float(x) + 0.5f) / float(P.size), (float(y) + 0.5f) / float(P.size), 0.0f);
    uchar4 v;
    SkyNoiseCreateInfo ci = P.a;
    v.x = unorm8(FBM<false>(coord, ci)); ci.seed += 1u;
    v.y = unorm8(FBM<false>(coord, ci)); ci.seed += 1u;
    v.z = unorm8(FBM<false>(coord, ci)); ci.seed += 1u;
    v.w = unorm8(FBM<false>(coord, ci));
    reinterpret_cast<uchar4*>(P.out)[y * P.size + x] = v;
}

// glGenerateTextureMipmap (VolumetricCloudDefaultMaterial.h:47, VolumetricCloudVoxelMaterial.cpp:75):
// 2x2(x2) box filter on the stored codes, odd trailing texels dropped, round half up.
struct MipParams {
    const uint8_t* src;
    uint8_t* dst;
    int sw, sh, sd, dw, dh, dd, channels;
};
__global__ void __launch_bounds__(256) k_mip_level(const __grid_constant__ MipParams P) {
    size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    size_t total = size_t(P.dw) * P.dh * P.dd * P.channels;
    if (idx >= total) return;
    int c = int(idx % P.channels);
    size_t t = idx / P.channels;
    int x = int(t % P.dw); t /= P.dw;
    int y = int(t % P.dh);
    int z = int(t / P.dh);
    int nx = P.sw > 1 ? 2 : 1, ny = P.sh > 1 ? 2 : 1, nz = P.sd > 1 ? 2 : 1;
    int sum = 0;
    for (int dz = 0; dz < nz; ++dz)
        for (int dy = 0; dy < ny; ++dy)
            for (int dx = 0; dx < nx; ++dx) {
                int sx = min(2 * x + dx, P.sw - 1), sy = min(2 * y + dy, P.sh - 1), sz = min(2 * z + dz, P.sd - 1);
                sum += P.src[((size_t(sz) * P.sh + sy) * P.sw + sx) * P.channels + c];
            }
    int n = nx * ny * nz;
    P.dst[idx] = uint8_t((2 * sum + n) / (2 * n));
}

// Corner packing of one level (see MipView::cells).  One thread per cell.
struct PackParams {
    const uint8_t* src;
    uint8_t* dst;
    int w, h, d, channels, border, is3d, cw, ch, cd, pad_xy, pad_z;
};
__global__ void __launch_bounds__(256) k_pack_cells(const __grid_constant__ PackParams P) {
    size_t idx = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    size_t total = size_t(P.cw) * P.ch * P.cd;
    if (idx >= total) return;
    int ci = int(idx % P.cw);
    size_t t = idx / P.cw;
    int cj = int(t % P.ch), ck = int(t / P.ch);
    const int corners = P.is3d ? 8 : 4;
    uint8_t* out = P.dst + idx * size_t(corners) * P.channels;
    for (int n = 0; n < corners; ++n) {
        int x = ci + (n & 1), y = cj + ((n >> 1) & 1), z = ck + ((n >> 2) & 1);
        bool inside = true;
        if (P.border) {
            x -= P.pad_xy; y -= P.pad_xy; if (P.is3d) z -= P.pad_z;
            inside = x >= 0 && x < P.w && y >= 0 && y < P.h && z >= 0 && z < P.d;
        } else {
            x %= P.w; y %= P.h; z %= P.d;
        }
        for (int c = 0; c < P.channels; ++c)
            out[n * P.channels + c] = inside ? P.src[((size_t(z) * P.h + y) * P.w + x) * P.channels + c] : uint8_t(0);
    }
}

}  // namespace

int build_mip_texture(SkyContext* ctx, MipTextureDev& t, int w, int h, int d, int channels, bool border) {
    if (t.valid && t.view.w[0] == w && t.view.h[0] == h && t.view.d[0] == d && t.view.channels == channels) return 0;
    if (t.view.tex_linear) cudaDestroyTextureObject(t.view.tex_linear);
    if (t.view.tex_point) cudaDestroyTextureObject(t.view.tex_point);
    if (t.array) cudaFreeMipmappedArray(t.array);
    if (t.data) cudaFree(t.data);
    if (t.cells) cudaFree(t.cells);
    t = MipTextureDev{};
    t.is3d = d > 1;
    t.border = border;
    MipView& v = t.view;
    v.channels = channels;
    size_t off = 0;
    int lw = w, lh = h, ld = d, levels = 0;
    for (;;) {
        v.w[levels] = lw; v.h[levels] = lh; v.d[levels] = ld; v.off[levels] = off;
        off += size_t(lw) * lh * ld;
        ++levels;
        if ((lw == 1 && lh == 1 && ld == 1) || levels == kMaxMipLevels) break;
        lw = max(lw / 2, 1); lh = max(lh / 2, 1); ld = max(ld / 2, 1);
    }
    v.levels = levels;
    t.bytes = off * channels;
    SKY_CUDA(ctx, cudaMalloc(&t.data, t.bytes));
    v.base = t.data;
    v.pad_xy = border ? 2 : 0;
    v.pad_z = border && t.is3d ? 1 : 0;
    size_t cells = 0;
    for (int l = 0; l < levels; ++l) {
        v.cell_off[l] = cells;
        v.cell_w[l] = border ? v.w[l] + 3 : v.w[l];
        v.cell_h[l] = border ? v.h[l] + 3 : v.h[l];
        v.cell_d[l] = t.is3d ? (border ? v.d[l] + 1 : v.d[l]) : 1;
        cells += size_t(v.cell_w[l]) * v.cell_h[l] * v.cell_d[l];
    }
    t.cell_bytes = cells * (t.is3d ? 8 : 4) * channels;
    SKY_CUDA(ctx, cudaMalloc(&t.cells, t.cell_bytes));
    v.cells = t.cells;

    cudaChannelFormatDesc desc = cudaCreateChannelDesc(8, channels >= 2 ? 8 : 0, channels >= 3 ? 8 : 0, channels >= 4 ? 8 : 0,
                                                       cudaChannelFormatKindUnsigned);
    cudaExtent ext = make_cudaExtent(size_t(w), size_t(h), t.is3d ? size_t(d) : 0);
    SKY_CUDA(ctx, cudaMallocMipmappedArray(&t.array, &desc, ext, unsigned(levels)));
    cudaResourceDesc res{};
    res.resType = cudaResourceTypeMipmappedArray;
    res.res.mipmap.mipmap = t.array;
    cudaTextureDesc td{};
    cudaTextureAddressMode mode = border ? cudaAddressModeBorder : cudaAddressModeWrap;
    td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = mode;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    td.mipmapFilterMode = cudaFilterModePoint;
    td.minMipmapLevelClamp = 0.0f;
    td.maxMipmapLevelClamp = float(levels - 1);
    td.filterMode = cudaFilterModeLinear;
    SKY_CUDA(ctx, cudaCreateTextureObject(&v.tex_linear, &res, &td, nullptr));
    td.filterMode = cudaFilterModePoint;
    SKY_CUDA(ctx, cudaCreateTextureObject(&v.tex_point, &res, &td, nullptr));
    t.valid = true;
    return 0;
}

// level 0 is in t.data already: build levels 1.. and mirror every level into the CUDA array
int launch_mip_chain(SkyContext* ctx, MipTextureDev& t) {
    MipView& v = t.view;
    for (int l = 1; l < v.levels; ++l) {
        MipParams P{t.data + v.off[l - 1] * v.channels, t.data + v.off[l] * v.channels,
                    v.w[l - 1], v.h[l - 1], v.d[l - 1], v.w[l], v.h[l], v.d[l], v.channels};
        size_t total = size_t(v.w[l]) * v.h[l] * v.d[l] * v.channels;
        k_mip_level<<<unsigned((total + 255) / 256), 256, 0, ctx->stream>>>(P);
        SKY_LAUNCH_CHECK(ctx);
    }
    const size_t cell_bytes = size_t(t.is3d ? 8 : 4) * v.channels;
    for (int l = 0; l < v.levels; ++l) {
        PackParams P{t.data + v.off[l] * v.channels, static_cast<uint8_t*>(t.cells) + v.cell_off[l] * cell_bytes, v.w[l], v.h[l], v.d[l], v.channels,
                     t.border ? 1 : 0, t.is3d ? 1 : 0, v.cell_w[l], v.cell_h[l], v.cell_d[l], v.pad_xy, v.pad_z};
        size_t total = size_t(v.cell_w[l]) * v.cell_h[l] * v.cell_d[l];
        k_pack_cells<<<unsigned((total + 255) / 256), 256, 0, ctx->stream>>>(P);
        SKY_LAUNCH_CHECK(ctx);
    }
    for (int l = 0; l < v.levels; ++l) {
        cudaArray_t level;
        SKY_CUDA(ctx, cudaGetMipmappedArrayLevel(&level, t.array, unsigned(l)));
        cudaMemcpy3DParms cp{};
        cp.srcPtr = make_cudaPitchedPtr(t.data + v.off[l] * v.channels, size_t(v.w[l]) * v.channels, size_t(v.w[l]), size_t(v.h[l]));
        cp.dstArray = level;
        cp.extent = make_cudaExtent(size_t(v.w[l]), size_t(v.h[l]), size_t(v.d[l]));
        cp.kind = cudaMemcpyDeviceToDevice;
        SKY_CUDA(ctx, cudaMemcpy3DAsync(&cp, ctx->stream));
    }
    return 0;
}

int launch_noise(SkyContext* ctx, int kind, const SkyNoiseCreateInfo* info) {
    NoiseParams P{};
    P.a = info[0];
    P.b = info[1];
    switch (kind) {
        case SKY_NOISE_CLOUD_MAP:
            if (int e = build_mip_texture(ctx, ctx->cloud_map, 512, 512, 1, 2, false)) return e;
            P.out = ctx->cloud_map.data; P.size = 512;
            k8_cloud_map<<<dim3(ceil_div(512, 128), 512), 128, 0, ctx->stream>>>(P);
            SKY_LAUNCH_CHECK(ctx);
            return launch_mip_chain(ctx, ctx->cloud_map);
        case SKY_NOISE_DETAIL:
            if (int e = build_mip_texture(ctx, ctx->detail, 128, 128, 128, 1, false)) return e;
            P.out = ctx->detail.data; P.size = 128;
            k9_detail<<<dim3(1, 128, 128), 128, 0, ctx->stream>>>(P);
            SKY_LAUNCH_CHECK(ctx);
            return launch_mip_chain(ctx, ctx->detail);
        case SKY_NOISE_DISPLACEMENT:
            if (int e = build_mip_texture(ctx, ctx->displacement, 128, 128, 1, 4, false)) return e;
            P.out = ctx->displacement.data; P.size = 128;
            k10_displacement<<<dim3(1, 128), 128, 0, ctx->stream>>>(P);
            SKY_LAUNCH_CHECK(ctx);
            return launch_mip_chain(ctx, ctx->displacement);
    }
    return sky_fail(ctx, "unknown noise kind");
}
