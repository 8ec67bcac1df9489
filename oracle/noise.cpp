// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h).
//
// noise.cpp: CPU restatement of shaders/Base/Noise.glsl and shaders/SkyRendering/NoiseGen.comp
// (K8 weather map, K9 Perlin-Worley detail volume, K10 displacement map).
#include "noise.h"

namespace orc {

// Noise.glsl:1-8
uint WangHash(uint seed) {
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}

// Noise.glsl:11-15
uint PCGHash(uint seed) {
    uint state = seed * 747796405u + 2891336453u;
    uint word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}

// Noise.glsl:17-22
static const float kPerlinGradients[16][3] = {
    {1, 1, 0}, {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 0, 1}, {-1, 0, 1}, {1, 0, -1}, {-1, 0, -1},
    {0, 1, 1}, {0, -1, 1}, {0, 1, -1}, {0, -1, -1}, {1, 1, 0}, {-1, 1, 0}, {0, -1, 1}, {0, -1, -1}};

// Noise.glsl:24-26
static vec3 GetPerlinGradients(uint i, uint j, uint k, uint seed) {
    return vec3(kPerlinGradients[WangHash(seed + WangHash(i + WangHash(j + WangHash(k)))) & 0xf]);
}

// Noise.glsl:28-60
float PerlinNoise(vec3 p, uint freq, uint seed) {
    p *= float(freq);
    vec3 fl = floor(p);
    vec3 ce(std::ceil(p.x), std::ceil(p.y), std::ceil(p.z));
    uint i0 = uint(int(fl.x)) % freq, j0 = uint(int(fl.y)) % freq, k0 = uint(int(fl.z)) % freq;
    uint i1 = uint(int(ce.x)) % freq, j1 = uint(int(ce.y)) % freq, k1 = uint(int(ce.z)) % freq;

    vec3 t = p - fl;
    vec3 uvw = t * t * t * (t * (t * 6.0f - 15.0f) + 10.0f);
    float u = uvw.x, v = uvw.y, w = uvw.z;
    float x0 = t.x, y0 = t.y, z0 = t.z;
    float x1 = t.x - 1.0f, y1 = t.y - 1.0f, z1 = t.z - 1.0f;

    return mix(mix(mix(dot(GetPerlinGradients(i0, j0, k0, seed), vec3(x0, y0, z0)),
                       dot(GetPerlinGradients(i1, j0, k0, seed), vec3(x1, y0, z0)), u),
                   mix(dot(GetPerlinGradients(i0, j1, k0, seed), vec3(x0, y1, z0)),
                       dot(GetPerlinGradients(i1, j1, k0, seed), vec3(x1, y1, z0)), u), v),
               mix(mix(dot(GetPerlinGradients(i0, j0, k1, seed), vec3(x0, y0, z1)),
                       dot(GetPerlinGradients(i1, j0, k1, seed), vec3(x1, y0, z1)), u),
                   mix(dot(GetPerlinGradients(i0, j1, k1, seed), vec3(x0, y1, z1)),
                       dot(GetPerlinGradients(i1, j1, k1, seed), vec3(x1, y1, z1)), u), v), w);
}

// Noise.glsl:62-80
float WorleyNoise2(vec2 p, uint freq, uint seed) {
    p *= float(freq);
    uint ix = uint(std::floor(p.x)), iy = uint(std::floor(p.y));
    p += vec2(float(freq));
    float min_dist = 1e10f;
    for (uint di = freq - 1; di <= freq + 1; ++di)
        for (uint dj = freq - 1; dj <= freq + 1; ++dj) {
            uint gx = ix + di, gy = iy + dj;
            uint sx = gx % freq, sy = gy % freq;
            uint rnd0 = WangHash(seed + WangHash(sx + WangHash(sy)));
            uint rnd1 = WangHash(seed + WangHash(sx + WangHash(sy + 1)));
            vec2 g = vec2(float(gx), float(gy)) + vec2(float(rnd0), float(rnd1)) / 4294967296.0f;
            min_dist = min(min_dist, distance(p, g));
        }
    return min_dist;
}

// Noise.glsl:82-101
float WorleyNoise3(vec3 p, uint freq, uint seed) {
    p *= float(freq);
    uint ix = uint(std::floor(p.x)), iy = uint(std::floor(p.y)), iz = uint(std::floor(p.z));
    p += vec3(float(freq));
    float min_dist = 1e10f;
    for (uint di = freq - 1; di <= freq + 1; ++di)
        for (uint dj = freq - 1; dj <= freq + 1; ++dj)
            for (uint dk = freq - 1; dk <= freq + 1; ++dk) {
                uint gx = ix + di, gy = iy + dj, gz = iz + dk;
                uint sx = gx % freq, sy = gy % freq, sz = gz % freq;
                uint rnd0 = WangHash(seed + WangHash(sx + WangHash(sy + WangHash(sz))));
                uint rnd1 = WangHash(seed + WangHash(sx + WangHash(sy + WangHash(sz + 1))));
                uint rnd2 = WangHash(seed + WangHash(sx + WangHash(sy + WangHash(sz + 2))));
                vec3 g = vec3(float(gx), float(gy), float(gz)) + vec3(float(rnd0), float(rnd1), float(rnd2)) / 4294967296.0f;
                min_dist = min(min_dist, distance(p, g));
            }
    return min_dist;
}

// NoiseGen.comp:12-18
static float RemapTo01(float x, float x0, float x1) { return clamp((x - x0) / (x1 - x0), 0.0f, 1.0f); }
static float RemapFrom01(float x, float y0, float y1) { return clamp(y0 + x * (y1 - y0), 0.0f, 1.0f); }

// NoiseGen.comp:20-34
float PerlinFBM(vec3 p, SkyNoiseCreateInfo ci) {
    float res = 0.0f;
    uint f = ci.base_frequency;
    float a = 0.5f, sum_a = 0.0f;
    for (uint c = 0; c < 8; ++c) {
        float noise = PerlinNoise(p, f, ci.seed) * 0.5f + 0.5f;
        res += RemapTo01(noise, ci.remap_min, ci.remap_max) * a;
        sum_a += a;
        f *= 2;
        a *= 0.5f;
    }
    return res / sum_a;
}

// NoiseGen.comp:36-50.  WorleyFBM(vec3 p) resolves to the vec3 overload even for the 2-D maps:
// CLOUD_MAP_GEN passes vec3(coord, 0.0) (NoiseGen.comp:84).
float WorleyFBM(vec3 p, SkyNoiseCreateInfo ci) {
    float res = 0.0f;
    uint f = ci.base_frequency;
    float a = 0.5f, sum_a = 0.0f;
    for (uint c = 0; c < 8; ++c) {
        float noise = WorleyNoise3(p, f, ci.seed);
        res += RemapTo01(noise, ci.remap_min, ci.remap_max) * a;
        sum_a += a;
        f *= 2;
        a *= 0.5f;
    }
    return res / sum_a;
}

static float unorm8(float x) { return MipTexture<1>::quantize(x, 8); }

// K8 -- NoiseGen.comp:72-88, rg8 image store
void GenerateCloudMap(const SkyNoiseCreateInfo info[2], MipTexture<2>& tex, int size) {
    tex.bits = 8;
    tex.levels.resize(1);
    Image<2>& img = tex.levels[0];
    img.resize(size, size);
#pragma omp parallel for schedule(dynamic)
    for (int y = 0; y < size; ++y)
        for (int x = 0; x < size; ++x) {
            vec2 coord = (vec2(float(x), float(y)) + 0.5f) / vec2(float(size), float(size));
            float density = PerlinFBM(vec3(coord, 0.0f), info[0]);
            float height = WorleyFBM(vec3(coord, 0.0f), info[1]);
            img.at(x, y)[0] = unorm8(density);
            img.at(x, y)[1] = unorm8(height);
        }
    tex.build_mips();
}

// K9 -- NoiseGen.comp:90-106, r8 image store
void GenerateDetail(const SkyNoiseCreateInfo info[2], MipTexture<1>& tex, int size) {
    tex.bits = 8;
    tex.levels.resize(1);
    Image<1>& img = tex.levels[0];
    img.resize(size, size, size);
#pragma omp parallel for schedule(dynamic) collapse(2)
    for (int z = 0; z < size; ++z)
        for (int y = 0; y < size; ++y)
            for (int x = 0; x < size; ++x) {
                vec3 coord = (vec3(float(x), float(y), float(z)) + 0.5f) / vec3(float(size), float(size), float(size));
                float perlin = PerlinFBM(coord, info[0]);
                float worley = WorleyFBM(coord, info[1]);
                float perlin_worley = RemapFrom01(perlin, worley, 1.0f);
                img.at(x, y, z)[0] = unorm8(perlin_worley);
            }
    tex.build_mips();
}

// K10 -- NoiseGen.comp:52-70, rgba8 image store
void GenerateDisplacement(const SkyNoiseCreateInfo info[1], MipTexture<4>& tex, int size) {
    tex.bits = 8;
    tex.levels.resize(1);
    Image<4>& img = tex.levels[0];
    img.resize(size, size);
#pragma omp parallel for schedule(dynamic)
    for (int y = 0; y < size; ++y)
        for (int x = 0; x < size; ++x) {
            vec2 coord = (vec2(float(x), float(y)) + 0.5f) / vec2(float(size), float(size));
            for (int i = 0; i < 4; ++i) {
                SkyNoiseCreateInfo ci = info[0];
                ci.seed += uint(i);
                img.at(x, y)[i] = unorm8(PerlinFBM(vec3(coord, 0.0f), ci));
            }
        }
    tex.build_mips();
}

}  // namespace orc
