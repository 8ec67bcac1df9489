// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h and ibl.h).
#include "ibl.h"

#include "../include/sky_cubemap.h"

#include <cmath>

namespace orc {

// GL 4.6 table 8.19 (cube-map face selection) + section 8.14.2 (bilinear) with SEAMLESS filtering (section 8.14.1; the reference
// enables GL_TEXTURE_CUBE_MAP_SEAMLESS, AtmosphereRenderer.cpp:151): taps off the face come from the adjacent face
// (include/sky_cubemap.h, shared with the kernels and the reference-shader shim).
vec4 TextureCubeLevel(const Image<4>& env, vec3 dir) {
    float ax = std::fabs(dir.x), ay = std::fabs(dir.y), az = std::fabs(dir.z);
    int face; float sc, tc, ma;
    if (ax >= ay && ax >= az) { ma = ax; if (dir.x >= 0) { face = 0; sc = -dir.z; tc = -dir.y; } else { face = 1; sc = dir.z; tc = -dir.y; } }
    else if (ay >= az)        { ma = ay; if (dir.y >= 0) { face = 2; sc = dir.x; tc = dir.z; } else { face = 3; sc = dir.x; tc = -dir.z; } }
    else                      { ma = az; if (dir.z >= 0) { face = 4; sc = dir.x; tc = -dir.y; } else { face = 5; sc = -dir.x; tc = -dir.y; } }
    float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
    int n = env.w;
    float u = s * float(n) - 0.5f, v = t * float(n) - 0.5f;
    float fu = std::floor(u), fv = std::floor(v);
    int i0 = int(fu), j0 = int(fv);
    float a = u - fu, b = v - fv;
    return sky_cube_bilinear<vec4>(n, face, i0, j0, a, b, [&](int f, int i, int j) { return env.load(i, j, f); });
}

vec4 TextureCubeLod(const CubeChain& cube, vec3 dir, float lod) {
    const int q = int(cube.levels.size()) - 1;
    float l = lod < 0.0f ? 0.0f : lod > float(q) ? float(q) : lod;
    float fl = std::floor(l);
    int l0 = int(fl);
    float f = l - fl;
    vec4 t0 = TextureCubeLevel(cube.levels[l0], dir);
    if (!(f > 0.0f)) return t0;
    vec4 t1 = TextureCubeLevel(cube.levels[l0 + 1 > q ? q : l0 + 1], dir);
    return t0 * (1.0f - f) + t1 * f;
}

void GenerateCubeMips(const Image<4>& level0, CubeChain& out) {
    out.levels.clear();
    out.levels.push_back(level0);
    for (int n = level0.w / 2; n >= 1; n /= 2) {
        const Image<4>& src = out.levels.back();
        Image<4> dst;
        dst.resize(n, n, 6);
        for (int f = 0; f < 6; ++f)
            for (int y = 0; y < n; ++y)
                for (int x = 0; x < n; ++x) {
                    vec4 t00 = src.load(2 * x, 2 * y, f), t10 = src.load(2 * x + 1, 2 * y, f);
                    vec4 t01 = src.load(2 * x, 2 * y + 1, f), t11 = src.load(2 * x + 1, 2 * y + 1, f);
                    vec4 m = ((t00 + t10) + (t01 + t11)) * 0.25f;
                    dst.store(x, y, f, vec4(to_half_and_back(m.x), to_half_and_back(m.y), to_half_and_back(m.z), to_half_and_back(m.w)));
                }
        out.levels.push_back(std::move(dst));
    }
}

namespace {

// shaders/Base/Noise.glsl:103-117
inline uint ReverseBits32(uint bits) {
    bits = (bits << 16) | (bits >> 16);
    bits = ((bits & 0x00ff00ffu) << 8) | ((bits & 0xff00ff00u) >> 8);
    bits = ((bits & 0x0f0f0f0fu) << 4) | ((bits & 0xf0f0f0f0u) >> 4);
    bits = ((bits & 0x33333333u) << 2) | ((bits & 0xccccccccu) >> 2);
    bits = ((bits & 0x55555555u) << 1) | ((bits & 0xaaaaaaaau) >> 1);
    return bits;
}
inline vec2 Hammersley(uint Index, uint NumSamples, uint random_x, uint random_y) {
    float E1 = fract(float(Index) / float(NumSamples) + float(random_x & 0xffffu) / float(1 << 16));
    float E2 = float(ReverseBits32(Index) ^ random_y) * 2.3283064365386963e-10f;
    return vec2(E1, E2);
}

// shaders/Base/Common.glsl:32-52
inline void CreateOrthonormalBasis(vec3 N, vec3& t0, vec3& t1) {
    float s = (N.z >= 0.0f ? 1.0f : -1.0f);
    float a = -1.0f / (s + N.z);
    float b = N.x * N.y * a;
    t0 = vec3(1.0f + s * N.x * N.x * a, s * b, -s * N.x);
    t1 = vec3(b, s + N.y * N.y * a, -N.y);
}

// shaders/Base/BRDF.glsl:36-66
inline float D_GGX(float a, float NdotH) {
    float a2 = a * a;
    float d = (NdotH * a2 - NdotH) * NdotH + 1.0f;
    return a2 / (PI * d * d);
}
inline float Vis_SmithJointApprox(float a, float NdotV, float NdotL) {
    float Vis_SmithV = NdotL * (NdotV * (1.0f - a) + a);
    float Vis_SmithL = NdotV * (NdotL * (1.0f - a) + a);
    return 0.5f / max(Vis_SmithV + Vis_SmithL, 1e-9f);
}
inline vec3 ImportanceSampleGGX(vec2 E, float a) {
    float a2 = a * a;
    float Phi = 2.0f * PI * E.x;
    float CosTheta = std::sqrt((1.0f - E.y) / (1.0f + (a2 - 1.0f) * E.y));
    float SinTheta = std::sqrt(1.0f - CosTheta * CosTheta);
    return vec3(SinTheta * sky_det_cosf(Phi), SinTheta * sky_det_sinf(Phi), CosTheta);
}
inline vec3 ImportanceSampleGGX(vec2 E, float a, vec3 N) {
    vec3 H = ImportanceSampleGGX(E, a);
    vec3 t0, t1;
    CreateOrthonormalBasis(N, t0, t1);
    return t0 * H.x + t1 * H.y + N * H.z;
}

// shaders/Base/EnvBRDFLut.comp:10-41
vec2 IntegrateBRDF(float Roughness, float NoV) {
    vec3 V(std::sqrt(1.0f - NoV * NoV), 0.0f, NoV);
    float a = Roughness * Roughness;
    float A = 0.0f, B = 0.0f;
    const uint NumSamples = 1024;
    for (uint i = 0; i < NumSamples; i++) {
        vec2 Xi = Hammersley(i, NumSamples, 0, 0);
        vec3 H = ImportanceSampleGGX(Xi, a);
        vec3 L = 2.0f * dot(V, H) * H - V;
        float NoL = clamp(L.z, 0.0f, 1.0f);
        float NoH = clamp(H.z, 0.0f, 1.0f);
        float VoH = clamp(dot(V, H), 0.0f, 1.0f);
        if (NoL > 0.0f) {
            float Vis = Vis_SmithJointApprox(a, NoV, NoL);
            float NoL_Vis_PDF = NoL * Vis * (4.0f * VoH / NoH);
            float Fc = std::pow(1.0f - VoH, 5.0f);
            A += (1.0f - Fc) * NoL_Vis_PDF;
            B += Fc * NoL_Vis_PDF;
        }
    }
    return vec2(A, B) / float(NumSamples);
}

}  // namespace

// K22 -- shaders/Base/EnvBRDFLut.comp:43-48, image format rg16 (Textures.cpp:64)
void BakeEnvBRDFLut(Image<2>& lut) {
    const int W = lut.w, H = lut.h;
#pragma omp parallel for schedule(dynamic)
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            vec2 uv = (vec2(float(x), float(y)) + vec2(0.5f)) / vec2(float(W), float(H));
            vec2 env_brdf = IntegrateBRDF(uv.y, uv.x);
            auto unorm16 = [](float v) { return std::nearbyint(clamp(v, 0.0f, 1.0f) * 65535.0f) / 65535.0f; };
            float* p = lut.at(x, y);
            p[0] = unorm16(env_brdf.x);
            p[1] = unorm16(env_brdf.y);
        }
}

// K23 -- shaders/Base/EnvRadianceSH.comp:29-85: 9 work groups of 1024 invocations, shared-memory tree in the reference's order
void EnvRadianceSH(const CubeChain& env, vec4 Llm[9]) {
    const float Y00 = 0.282095f, Y1n = 0.488603f, Y2n = 1.092548f, Y20 = 0.315392f, Y22 = 0.546274f;
    for (int index = 0; index < 9; ++index) {
        std::vector<vec3> Llm_local(1024);
        for (int local_index = 0; local_index < 1024; ++local_index) {
            float unit_theta = (0.5f + float(local_index >> 5)) / 32.0f;
            float unit_phi = (0.5f + float(local_index & 0x1f)) / 32.0f;
            float cos_theta = 1.0f - 2.0f * unit_theta;
            float sin_theta = std::sqrt(clamp(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
            float phi = 2.0f * PI * unit_phi;
            float cos_phi = sky_det_cosf(phi);
            float sin_phi = sky_det_sinf(phi);
            vec3 dir(cos_phi * sin_theta, cos_theta, sin_phi * sin_theta);
            vec3 radiance = TextureCubeLod(env, dir, 0.0f).rgb();
            float coeff[9] = {Y00, Y1n * dir.y, Y1n * dir.z, Y1n * dir.x, Y2n * dir.x * dir.y, Y2n * dir.y * dir.z,
                              Y20 * (3.0f * dir.z * dir.z - 1.0f), Y2n * dir.x * dir.z, Y22 * (dir.x * dir.x - dir.y * dir.y)};
            Llm_local[local_index] = radiance * (coeff[index] * (4.0f * PI / 1024.0f));
        }
        for (int stride = 512; stride >= 1; stride >>= 1)
            for (int i = 0; i < stride; ++i) Llm_local[i] = Llm_local[i] + Llm_local[i + stride];
        Llm[index] = vec4(Llm_local[0], 0.0f);
    }
}

// K24 -- shaders/Base/PrefilterRadiance.comp:12-50, dispatched per level by IBL.cpp:37-41
void PrefilterRadiance(const CubeChain& env, int size, int roughness_count, CubeChain& out) {
    out.levels.assign(roughness_count, Image<4>());
    for (int level = 0, w = size; level < roughness_count; ++level, w >>= 1) {
        Image<4>& img = out.levels[level];
        img.resize(w, w, 6);
        const float roughness = float(level) / float(roughness_count - 1);
#pragma omp parallel for schedule(dynamic) collapse(2)
        for (int index = 0; index < 6; ++index)
            for (int y = 0; y < w; ++y)
                for (int x = 0; x < w; ++x) {
                    vec2 face_uv = (vec2(float(x), float(y)) + vec2(0.5f)) / vec2(float(w), float(w));
                    face_uv.y = 1.0f - face_uv.y;
                    vec3 R = ConvertCubUvToDir(index, face_uv);
                    // PrefilterEnvMap(roughness, R)
                    float a = roughness * roughness;
                    vec3 N = R, V = R;
                    vec3 PrefilteredColor(0.0f);
                    const uint NumSamples = uint(mix(1.0f, 64.0f, std::pow(roughness, 0.3f)));
                    float TotalWeight = 0.0f;
                    for (uint i = 0; i < NumSamples; i++) {
                        vec2 Xi = Hammersley(i, NumSamples, 0, 0);
                        vec3 H = ImportanceSampleGGX(Xi, a, N);
                        vec3 L = 2.0f * dot(V, H) * H - V;
                        float NoL = clamp(dot(N, L), 0.0f, 1.0f);
                        if (NoL > 0.0f) {
                            float NoH = clamp(dot(N, H), 0.0f, 1.0f);
                            float HoV = clamp(dot(H, V), 0.0f, 1.0f);
                            float D = D_GGX(a, NoH);
                            float pdf = max(D * NoH / (4.0f * HoV), 0.0001f);
                            float invSaTexel = (6.0f * float(w) * float(w)) / (4.0f * PI);
                            float saSample = 1.0f / max(float(NumSamples) * pdf, 0.00001f);
                            float mipLevel = roughness == 0.0f ? 0.0f : 0.5f * std::log2(saSample * invSaTexel) + 2.5f;
                            PrefilteredColor = PrefilteredColor + TextureCubeLod(env, L, mipLevel).rgb() * NoL;
                            TotalWeight += NoL;
                        }
                    }
                    vec3 c = PrefilteredColor / TotalWeight;
                    img.store(x, y, index, vec4(to_half_and_back(c.x), to_half_and_back(c.y), to_half_and_back(c.z), 0.0f));
                }
    }
}

}  // namespace orc
