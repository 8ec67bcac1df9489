// Image-based-lighting chain: the inputs of the composite's object shading.
//   K22 k22_env_brdf_lut       shaders/Base/EnvBRDFLut.comp        (Textures::Textures, src/Base/src/Textures.cpp:60-75)
//   cube mips                  glGenerateTextureMipmap             (AtmosphereRenderer.cpp:242)
//   K23 EnvRadianceSH          shaders/Base/EnvRadianceSH.comp     (IBL::Precompute, src/Base/src/IBL.cpp:29-34)
//   K24 k24_prefilter_radiance shaders/Base/PrefilterRadiance.comp (IBL.cpp:35-42)
// Compiled with -fmad=false like the LUT bake: IEEE division / sqrt, sin / cos from include/sky_detmath.h, the reference's
// operation order (K23 keeps its shared-memory tree, K22 / K24 their serial sample sums), so K23, the mips and every
// level of K24 whose LOD arithmetic does not involve log2 are bit-identical to the oracle; powf / log2f are CUDA's (see
// tests/test_gpu_ibl.py for the stated tolerance).
//
// Mapping (all three are small, latency-bound kernels of the per-frame LUT phase, so the point is FEW launches):
//   * one launch does the cube mips (6 blocks, one per face, levels chained through __syncthreads -- a face's chain only
//     depends on that face) AND K23 (9 blocks of 1024 threads like the reference's 9 work groups; it only reads level 0);
//   * one launch does all five roughness levels of K24, heavy levels first (level 0 is a 1-sample copy); 8 lanes evaluate a
//     texel's samples side by side, one lane adds them in the reference's order;
//   * K22 (start-up, input-free): a block is 256 texels of one row, i.e. one roughness, so the tangent-space half vectors
//     of the 1024 Hammersley samples are computed once per block into shared memory (2 sqrt + 1 division + sin + cos per
//     sample hoisted out of every thread's loop; same functions on the same inputs, so still bit-identical).
#include "context.h"
#include "../../include/sky_detmath.h"
#include "ibl_dev.cuh"

namespace {

// shaders/Base/Noise.glsl:113-117 with Random = uvec2(0)
SKY_D float2 Hammersley0(uint32_t Index, uint32_t NumSamples) {
    float E1 = fractf(float(Index) / float(NumSamples) + float(0u & 0xffffu) / float(1 << 16));
    float E2 = float(__brev(Index) ^ 0u) * 2.3283064365386963e-10f;
    return f2(E1, E2);
}
// shaders/Base/BRDF.glsl:51-63 (tangent space)
SKY_D float3 ImportanceSampleGGX(float2 E, float a) {
    float a2 = a * a;
    float Phi = 2.0f * kPi * E.x;
    float CosTheta = sqrtf((1.0f - E.y) / (1.0f + (a2 - 1.0f) * E.y));
    float SinTheta = sqrtf(1.0f - CosTheta * CosTheta);
    return f3(SinTheta * sky_det_cosf(Phi), SinTheta * sky_det_sinf(Phi), CosTheta);
}
// shaders/Base/BRDF.glsl:36-40
SKY_D float D_GGX(float a, float NdotH) {
    float a2 = a * a;
    float d = (NdotH * a2 - NdotH) * NdotH + 1.0f;
    return a2 / (kPi * d * d);
}
// shaders/Base/Common.glsl:32-52
SKY_D void CreateOrthonormalBasis(float3 N, float3& t0, float3& t1) {
    float s = (N.z >= 0.0f ? 1.0f : -1.0f);
    float a = -1.0f / (s + N.z);
    float b = N.x * N.y * a;
    t0 = f3(1.0f + s * N.x * N.x * a, s * b, -s * N.x);
    t1 = f3(b, s + N.y * N.y * a, -N.y);
}
// shaders/Base/Common.glsl:13-30
SKY_D float3 ConvertCubUvToDir(int index, float u, float v) {
    float uc = 2.0f * u - 1.0f, vc = 2.0f * v - 1.0f;
    float3 dir = f3(0.0f);
    switch (index) {
        case 0: dir = f3(1.0f, vc, -uc); break;
        case 1: dir = f3(-1.0f, vc, uc); break;
        case 2: dir = f3(uc, 1.0f, -vc); break;
        case 3: dir = f3(uc, -1.0f, vc); break;
        case 4: dir = f3(uc, vc, 1.0f); break;
        case 5: dir = f3(-uc, vc, -1.0f); break;
    }
    return normalize(dir);
}

// ------------------------------------------------------------------------------------------------------------ K22
constexpr int kBrdfSamples = 1024;  // EnvBRDFLut.comp:20
__global__ void __launch_bounds__(256) k22_env_brdf_lut(ushort2* __restrict__ out, int W, int H) {
    __shared__ float3 sH[kBrdfSamples];
    const int y = blockIdx.y, x = blockIdx.x * 256 + threadIdx.x;
    const float Roughness = (float(y) + 0.5f) / float(H);  // uv.y
    const float a = Roughness * Roughness;
    for (int i = threadIdx.x; i < kBrdfSamples; i += 256) sH[i] = ImportanceSampleGGX(Hammersley0(uint32_t(i), kBrdfSamples), a);
    __syncthreads();
    if (x >= W) return;
    const float NoV = (float(x) + 0.5f) / float(W);  // uv.x
    const float3 V = f3(sqrtf(1.0f - NoV * NoV), 0.0f, NoV);
    float A = 0.0f, B = 0.0f;
#pragma unroll 4
    for (int i = 0; i < kBrdfSamples; ++i) {
        const float3 Hh = sH[i];
        const float VdotH = dot(V, Hh);
        const float Lz = 2.0f * VdotH * Hh.z - V.z;
        const float NoL = clampf(Lz, 0.0f, 1.0f);
        const float NoH = clampf(Hh.z, 0.0f, 1.0f);
        const float VoH = clampf(VdotH, 0.0f, 1.0f);
        if (NoL > 0.0f) {
            // Vis_SmithJointApprox, BRDF.glsl:42-46
            float Vis_SmithV = NoL * (NoV * (1.0f - a) + a);
            float Vis_SmithL = NoV * (NoL * (1.0f - a) + a);
            float Vis = 0.5f / fmaxf(Vis_SmithV + Vis_SmithL, 1e-9f);
            float NoL_Vis_PDF = NoL * Vis * (4.0f * VoH / NoH);
            float Fc = powf(1.0f - VoH, 5.0f);
            A += (1.0f - Fc) * NoL_Vis_PDF;
            B += Fc * NoL_Vis_PDF;
        }
    }
    A = A / float(kBrdfSamples);
    B = B / float(kBrdfSamples);
    // rg16 image store: round to nearest even
    out[size_t(y) * W + x] = make_ushort2((unsigned short)__float2uint_rn(clampf(A, 0.0f, 1.0f) * 65535.0f),
                                          (unsigned short)__float2uint_rn(clampf(B, 0.0f, 1.0f) * 65535.0f));
}

// ------------------------------------------------------------------------------------------- cube mips + K23 (one launch)
struct MipShParams {
    const half4* level0;  // [6][n][n]
    half4* mips;          // levels 1 .. concatenated
    int n, levels;
    float4* sh;           // Llm[9]
};

SKY_D void cube_face_mips(const MipShParams& P, int face) {
    const half4* src = P.level0 + size_t(face) * P.n * P.n;
    half4* dst_level = P.mips;
    for (int l = 1, n = P.n >> 1; l < P.levels; ++l, n >>= 1) {
        half4* dst = dst_level + size_t(face) * n * n;
        for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
            int x = t % n, y = t / n;
            const half4* r0 = src + size_t(2 * y) * (2 * n) + 2 * x;
            const half4* r1 = r0 + 2 * n;
            float4 t00 = from_half4(r0[0]), t10 = from_half4(r0[1]), t01 = from_half4(r1[0]), t11 = from_half4(r1[1]);  // plain loads: written by this block
            float4 m = ((t00 + t10) + (t01 + t11)) * 0.25f;
            dst[t] = to_half4(m);
        }
        __syncthreads();
        src = dst;
        dst_level += size_t(6) * n * n;
    }
}

// EnvRadianceSH.comp:29-85
SKY_D void env_radiance_sh(const MipShParams& P, int index, float3* Llm_local) {
    const float Y00 = 0.282095f, Y1n = 0.488603f, Y2n = 1.092548f, Y20 = 0.315392f, Y22 = 0.546274f;
    const int local_index = threadIdx.x;
    float unit_theta = (0.5f + float(local_index >> 5)) / 32.0f;
    float unit_phi = (0.5f + float(local_index & 0x1f)) / 32.0f;
    float cos_theta = 1.0f - 2.0f * unit_theta;
    float sin_theta = sqrtf(clampf(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
    float phi = 2.0f * kPi * unit_phi;
    float cos_phi = sky_det_cosf(phi);
    float sin_phi = sky_det_sinf(phi);
    float3 dir = f3(cos_phi * sin_theta, cos_theta, sin_phi * sin_theta);
    float3 radiance = xyz(TextureCubeLevel(P.level0, P.n, dir));
    float c;
    switch (index) {
        case 0: c = Y00; break;
        case 1: c = Y1n * dir.y; break;
        case 2: c = Y1n * dir.z; break;
        case 3: c = Y1n * dir.x; break;
        case 4: c = Y2n * dir.x * dir.y; break;
        case 5: c = Y2n * dir.y * dir.z; break;
        case 6: c = Y20 * (3.0f * dir.z * dir.z - 1.0f); break;
        case 7: c = Y2n * dir.x * dir.z; break;
        default: c = Y22 * (dir.x * dir.x - dir.y * dir.y); break;
    }
    Llm_local[local_index] = radiance * (c * (4.0f * kPi / 1024.0f));
    __syncthreads();
    for (int stride = 512; stride >= 1; stride >>= 1) {  // the reference's tree: [i] += [i + stride]
        if (local_index < stride) Llm_local[local_index] = Llm_local[local_index] + Llm_local[local_index + stride];
        __syncthreads();
    }
    if (local_index == 0) P.sh[index] = f4(Llm_local[0], 0.0f);
}

__global__ void __launch_bounds__(1024) k23_env_sh_and_cube_mips(const __grid_constant__ MipShParams P) {
    __shared__ float3 Llm_local[1024];
    if (blockIdx.x < 9) env_radiance_sh(P, blockIdx.x, Llm_local);
    else cube_face_mips(P, blockIdx.x - 9);
}

// ------------------------------------------------------------------------------------------------------------ K24
struct PrefilterParams {
    CubeChainView env;
    half4* out[SKY_IBL_ROUGHNESS_COUNT];
    int size;
    uint32_t num_samples[SKY_IBL_ROUGHNESS_COUNT];  // uint(mix(1, 64, pow(roughness, 0.3))), PrefilterRadiance.comp:20
    int first_block[SKY_IBL_ROUGHNESS_COUNT + 1];    // launch order: levels 1, 2, ... then 0 (heavy first)
};

// The reference sums a texel's samples serially.  What is expensive per sample -- two divisions, log2, two bilinear cube taps on
// two mip levels -- does not depend on the running sum, so kLanes lanes evaluate a texel's samples side by side into shared
// memory and ONE lane then adds them in the reference's order (same operands, same order: same bits).  The dependent chain of a
// level-4 texel drops from 64 x (tap latency) to 8 x (tap latency) + 64 additions.  Level 0 is one sample per texel: 1 lane.
constexpr int kPrefilterThreads = 128, kPrefilterLanes = 8, kPrefilterMaxSamples = 64;  // kNumSamplesMax, PrefilterRadiance.comp:19
SKY_HD int prefilter_lanes(int level) { return level == 0 ? 1 : kPrefilterLanes; }

__global__ void __launch_bounds__(kPrefilterThreads) k24_prefilter_radiance(const __grid_constant__ PrefilterParams P) {
    __shared__ float3 sH[kPrefilterMaxSamples];
    // (radiance * NoL, NoL) of sample i of the block's texel `local` at [local * NumSamples + i]; NoL == 0: sample skipped.
    // 16 texels x <= 64 samples with 8 lanes per texel, 128 texels x 1 sample on level 0: at most 1024 entries either way
    __shared__ float4 sC[kPrefilterThreads / kPrefilterLanes * kPrefilterMaxSamples];
    int slot = 0;
#pragma unroll
    for (int k = 1; k < SKY_IBL_ROUGHNESS_COUNT; ++k) slot += int(blockIdx.x) >= P.first_block[k];
    const int level = (slot + 1) % SKY_IBL_ROUGHNESS_COUNT;
    const int w = P.size >> level;
    const int lanes = prefilter_lanes(level), texels_per_block = kPrefilterThreads / lanes;
    const int local = threadIdx.x / lanes, lane = threadIdx.x % lanes;
    const int t = (int(blockIdx.x) - P.first_block[slot]) * texels_per_block + local;
    const float roughness = float(level) / float(SKY_IBL_ROUGHNESS_COUNT - 1);  // IBL.cpp:39
    const uint32_t NumSamples = P.num_samples[level];
    // a block is one roughness level: the tangent-space half vectors (sin, cos, 2 sqrt, 1 division per sample) depend on the
    // sample index alone, so they are evaluated once per block instead of once per texel and sample (same values)
    if (threadIdx.x < NumSamples) sH[threadIdx.x] = ImportanceSampleGGX(Hammersley0(threadIdx.x, NumSamples), roughness * roughness);
    __syncthreads();
    const bool valid = t < 6 * w * w;
    const int x = t % w, y = (t / w) % w, index = valid ? t / (w * w) : 0;
    float fu = (float(x) + 0.5f) / float(w), fv = (float(y) + 0.5f) / float(w);
    fv = 1.0f - fv;
    const float3 R = ConvertCubUvToDir(index, fu, fv);
    // PrefilterEnvMap, PrefilterRadiance.comp:12-40
    const float a = roughness * roughness;
    const float3 N = R, V = R;
    float3 t0, t1;
    CreateOrthonormalBasis(N, t0, t1);
    const float invSaTexel = (6.0f * float(w) * float(w)) / (4.0f * kPi);
    if (valid) {
        for (uint32_t i = lane; i < NumSamples; i += lanes) {
            float3 Hl = sH[i];
            float3 Hw = t0 * Hl.x + t1 * Hl.y + N * Hl.z;
            float3 L = 2.0f * dot(V, Hw) * Hw - V;
            float NoL = clampf(dot(N, L), 0.0f, 1.0f);
            float4 c = f4(0.0f, 0.0f, 0.0f, 0.0f);
            if (NoL > 0.0f) {
                float NoH = clampf(dot(N, Hw), 0.0f, 1.0f);
                float HoV = clampf(dot(Hw, V), 0.0f, 1.0f);
                float D = D_GGX(a, NoH);
                float pdf = fmaxf(D * NoH / (4.0f * HoV), 0.0001f);
                float saSample = 1.0f / fmaxf(float(NumSamples) * pdf, 0.00001f);
                float mipLevel = roughness == 0.0f ? 0.0f : 0.5f * log2f(saSample * invSaTexel) + 2.5f;
                c = f4(xyz(TextureCubeLod(P.env, L, mipLevel)) * NoL, NoL);
            }
            sC[local * NumSamples + i] = c;
        }
    }
    __syncthreads();
    if (!valid || lane != 0) return;
    float3 PrefilteredColor = f3(0.0f);
    float TotalWeight = 0.0f;
    for (uint32_t i = 0; i < NumSamples; i++) {
        float4 c = sC[local * NumSamples + i];
        if (c.w > 0.0f) {
            PrefilteredColor = PrefilteredColor + xyz(c);
            TotalWeight += c.w;
        }
    }
    float3 c = PrefilteredColor / TotalWeight;
    P.out[level][(size_t(index) * w + y) * w + x] = to_half4(f4(c, 0.0f));
}

}  // namespace

int launch_env_brdf_lut(SkyContext* ctx) {
    const int S = SKY_ENV_BRDF_LUT_SIZE;
    k22_env_brdf_lut<<<dim3(ceil_div(S, 256), S), 256, 0, ctx->stream>>>(ctx->env_brdf_lut.p, S, S);
    SKY_LAUNCH_CHECK(ctx);
    return 0;
}

int launch_ibl_precompute(SkyContext* ctx) {
    const int n = ctx->env.w;
    int levels = 1;
    while ((n >> levels) >= 1) ++levels;
    MipShParams M{};
    M.level0 = ctx->env.p; M.mips = ctx->env_mips; M.n = n; M.levels = levels; M.sh = ctx->env_sh.p;
    nvtxRangePushA("Environment Radiance SH");  // IBL.cpp:29 (+ glGenerateTextureMipmap of the cube, AtmosphereRenderer.cpp:243)
    k23_env_sh_and_cube_mips<<<9 + 6, 1024, 0, ctx->stream>>>(M);
    nvtxRangePop();
    SKY_LAUNCH_CHECK(ctx);

    PrefilterParams P{};
    P.env.n = n; P.env.levels = levels;
    P.env.level[0] = ctx->env.p;
    const half4* p = ctx->env_mips;
    for (int l = 1; l < levels; ++l) { P.env.level[l] = p; p += size_t(6) * (n >> l) * (n >> l); }
    P.size = SKY_IBL_PREFILTERED_RESOLUTION;
    half4* o = ctx->prefiltered;
    for (int l = 0; l < SKY_IBL_ROUGHNESS_COUNT; ++l) {
        const int w = P.size >> l;
        P.out[l] = o; o += size_t(6) * w * w;
        const float roughness = float(l) / float(SKY_IBL_ROUGHNESS_COUNT - 1);
        const float tpow = powf(roughness, 0.3f);
        P.num_samples[l] = uint32_t(1.0f * (1.0f - tpow) + 64.0f * tpow);  // mix(kNumSamplesMin, kNumSamplesMax, pow(roughness, 0.3))
    }
    int blocks = 0;
    for (int k = 0; k < SKY_IBL_ROUGHNESS_COUNT; ++k) {
        const int level = (k + 1) % SKY_IBL_ROUGHNESS_COUNT, w = P.size >> level;
        P.first_block[k] = blocks;
        blocks += ceil_div(6 * w * w, kPrefilterThreads / prefilter_lanes(level));
    }
    P.first_block[SKY_IBL_ROUGHNESS_COUNT] = blocks;
    SKY_PERF_MARKER("Prefilter Radiance");  // IBL.cpp:35
    k24_prefilter_radiance<<<blocks, kPrefilterThreads, 0, ctx->stream>>>(P);
    SKY_LAUNCH_CHECK(ctx);
    return 0;
}
