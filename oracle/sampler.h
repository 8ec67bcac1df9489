// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h).
//
// sampler.h: software restatement of the OpenGL 4.6 sampling rules the reference shaders rely on
// (spec section 8.14 "Texture Minification", 8.15 "Texture Magnification"); the reference's own
// sampler objects are in src/Base/src/Samplers.cpp:3-52, VolumetricCloud.cpp:106-112 and
// VolumetricCloudVoxelMaterial.cpp:30-37.  The GL driver is third-party arithmetic that is not
// in /root/reference, so the conventions that the spec leaves open are fixed here and documented:
//   * filtering weights are exact fp32 (hardware uses >= 8 fractional bits),
//   * mip chains are 2x2(x2) box filters, odd trailing texels are dropped (floor convention),
//   * UNORM stores round to nearest-even.
#pragma once
#include <vector>

#include "glsl.h"

namespace orc {
using namespace glsl;

enum Wrap { CLAMP_TO_EDGE, REPEAT, CLAMP_TO_BORDER };

inline int wrap_index(int i, int n, Wrap w, bool& border) {
    border = false;
    switch (w) {
        case CLAMP_TO_EDGE: return clamp(i, 0, n - 1);
        case REPEAT: { int m = i % n; return m < 0 ? m + n : m; }
        default: if (i < 0 || i >= n) { border = true; return 0; } return i;
    }
}

// A float image with C channels; storage quantisation is applied by the writers.
template <int C>
struct Image {
    int w = 0, h = 0, d = 1;
    std::vector<float> data;
    void resize(int w_, int h_, int d_ = 1) { w = w_; h = h_; d = d_; data.assign(size_t(w) * h * d * C, 0.0f); }
    float* at(int x, int y, int z = 0) { return &data[(size_t(z) * h + y) * w * C + size_t(x) * C]; }
    const float* at(int x, int y, int z = 0) const { return &data[(size_t(z) * h + y) * w * C + size_t(x) * C]; }
    vec4 load(int x, int y, int z = 0) const {
        const float* p = at(x, y, z);
        vec4 r(0, 0, 0, 1);
        for (int c = 0; c < C; ++c) r[c] = p[c];
        return r;
    }
    void store(int x, int y, int z, vec4 v) { float* p = at(x, y, z); for (int c = 0; c < C; ++c) p[c] = v[c]; }
    void store(int x, int y, vec4 v) { store(x, y, 0, v); }
};

struct Sampler {
    Wrap wrap = CLAMP_TO_EDGE;
    vec4 border = vec4(0.0f);
};

template <int C>
inline vec4 fetch_wrapped(const Image<C>& img, int x, int y, int z, const Sampler& s) {
    bool bx, by, bz = false;
    int ix = wrap_index(x, img.w, s.wrap, bx);
    int iy = wrap_index(y, img.h, s.wrap, by);
    int iz = img.d > 1 ? wrap_index(z, img.d, s.wrap, bz) : 0;
    if (bx || by || bz) return s.border;
    return img.load(ix, iy, iz);
}

// texelFetch with coordinates clamped into the image (texelFetchClamp,
// VolumetricCloudCommon.glsl:54-56; also used where the reference fetches out of range and
// relies on robust access -- the oracle defines clamp, SURVEY.md section 7 "hard parts").
template <int C>
inline vec4 texel_fetch_clamp(const Image<C>& img, ivec2 p) {
    return img.load(clamp(p.x, 0, img.w - 1), clamp(p.y, 0, img.h - 1));
}

// GL_LINEAR 2-D: u = s*w - 0.5, i0 = floor(u), weights frac(u)  (spec eq. 8.13 / 8.14).
template <int C>
inline vec4 texture_linear(const Image<C>& img, vec2 uv, const Sampler& s) {
    float u = uv.x * float(img.w) - 0.5f, v = uv.y * float(img.h) - 0.5f;
    float fu = std::floor(u), fv = std::floor(v);
    int i0 = int(fu), j0 = int(fv);
    float a = u - fu, b = v - fv;
    vec4 t00 = fetch_wrapped(img, i0, j0, 0, s), t10 = fetch_wrapped(img, i0 + 1, j0, 0, s);
    vec4 t01 = fetch_wrapped(img, i0, j0 + 1, 0, s), t11 = fetch_wrapped(img, i0 + 1, j0 + 1, 0, s);
    return (1.0f - a) * (1.0f - b) * t00 + a * (1.0f - b) * t10 + (1.0f - a) * b * t01 + a * b * t11;
}

// GL_LINEAR 3-D.
template <int C>
inline vec4 texture_linear(const Image<C>& img, vec3 uvw, const Sampler& s) {
    float u = uvw.x * float(img.w) - 0.5f, v = uvw.y * float(img.h) - 0.5f, w = uvw.z * float(img.d) - 0.5f;
    float fu = std::floor(u), fv = std::floor(v), fw = std::floor(w);
    int i0 = int(fu), j0 = int(fv), k0 = int(fw);
    float a = u - fu, b = v - fv, c = w - fw;
    vec4 r(0.0f);
    for (int dk = 0; dk < 2; ++dk)
        for (int dj = 0; dj < 2; ++dj)
            for (int di = 0; di < 2; ++di) {
                float wt = (di ? a : 1.0f - a) * (dj ? b : 1.0f - b) * (dk ? c : 1.0f - c);
                r += wt * fetch_wrapped(img, i0 + di, j0 + dj, k0 + dk, s);
            }
    return r;
}

// GL_NEAREST: i = floor(s*w).
template <int C>
inline vec4 texture_nearest(const Image<C>& img, vec2 uv, const Sampler& s) {
    return fetch_wrapped(img, int(std::floor(uv.x * float(img.w))), int(std::floor(uv.y * float(img.h))), 0, s);
}
template <int C>
inline vec4 texture_nearest(const Image<C>& img, vec3 uvw, const Sampler& s) {
    return fetch_wrapped(img, int(std::floor(uvw.x * float(img.w))), int(std::floor(uvw.y * float(img.h))),
                         int(std::floor(uvw.z * float(img.d))), s);
}

// textureGather with the base texel given as an integer: returns (i0,j1),(i1,j1),(i1,j0),(i0,j0)
// of component `comp`, clamp-to-edge.  Every gather in the reference lands exactly on a 2x2
// block corner (CheckerboardGen.comp:9-10, VolumetricCloudIndexGen.comp:15-17,
// VolumetricCloudReconstruct.comp:37-38, VolumetricCloudUpscale.comp:18), so the oracle takes
// the intended integer base instead of re-deriving it from a float that sits on a rounding edge.
template <int C>
inline vec4 texture_gather(const Image<C>& img, ivec2 base, int comp = 0) {
    auto f = [&](int x, int y) { return img.at(clamp(x, 0, img.w - 1), clamp(y, 0, img.h - 1))[comp]; };
    return vec4(f(base.x, base.y + 1), f(base.x + 1, base.y + 1), f(base.x + 1, base.y), f(base.x, base.y));
}

// A mip-mapped UNORM texture sampled with mag = LINEAR, min = NEAREST_MIPMAP_NEAREST and an
// explicit LOD (textureLod): the material samplers, VolumetricCloudDefaultMaterial.cpp:111-116,
// VolumetricCloudVoxelMaterial.cpp:30-37.
template <int C>
struct MipTexture {
    std::vector<Image<C>> levels;
    int bits = 8;  // UNORM storage width

    static float quantize(float x, int bits) {
        float m = float((1u << bits) - 1u);
        return std::nearbyint(clamp(x, 0.0f, 1.0f) * m) / m;
    }
    // glGenerateTextureMipmap: box filter, floor convention, re-quantised per level.
    void build_mips() {
        levels.resize(1);
        while (true) {
            const Image<C>& src = levels.back();
            if (src.w == 1 && src.h == 1 && src.d == 1) break;
            Image<C> dst;
            dst.resize(max(src.w / 2, 1), max(src.h / 2, 1), max(src.d / 2, 1));
            float m = float((1u << bits) - 1u);
            for (int z = 0; z < dst.d; ++z)
                for (int y = 0; y < dst.h; ++y)
                    for (int x = 0; x < dst.w; ++x)
                        for (int c = 0; c < C; ++c) {
                            // integer sum of the stored codes keeps the average exact
                            int sum = 0, n = 0;
                            for (int dz = 0; dz < (src.d > 1 ? 2 : 1); ++dz)
                                for (int dy = 0; dy < (src.h > 1 ? 2 : 1); ++dy)
                                    for (int dx = 0; dx < (src.w > 1 ? 2 : 1); ++dx) {
                                        int sx = min(2 * x + dx, src.w - 1), sy = min(2 * y + dy, src.h - 1),
                                            sz = min(2 * z + dz, src.d - 1);
                                        sum += int(std::nearbyint(src.at(sx, sy, sz)[c] * m));
                                        ++n;
                                    }
                            // round-half-up on the integer average (ties are representable exactly)
                            dst.at(x, y, z)[c] = float((2 * sum + n) / (2 * n)) / m;
                        }
            levels.push_back(std::move(dst));
        }
    }
    // spec 8.14.3: lambda <= 0.5 -> magnification (LINEAR, base level);
    // otherwise level d = ceil(lambda + 0.5) - 1 clamped to the last level, NEAREST.
    template <class V>
    vec4 texture_lod(V coord, float lod, const Sampler& s) const {
        int q = int(levels.size()) - 1;
        if (!(lod > 0.5f)) return texture_linear(levels[0], coord, s);
        int dlevel = (lod <= float(q) + 0.5f) ? int(std::ceil(lod + 0.5f)) - 1 : q;
        dlevel = clamp(dlevel, 0, q);
        return texture_nearest(levels[dlevel], coord, s);
    }
};

}  // namespace orc
