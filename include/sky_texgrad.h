/*
 * sky_texgrad.h -- textureGrad() on a mip-mapped 2-D texture with an anisotropic LINEAR_MIPMAP_LINEAR sampler.
 *
 * The reference's ground pass (shaders/SkyRendering/EarthRender.frag:22-38) samples the equirectangular earth albedo map with
 * textureGrad through the sampler of src/SkyRendering/Earth.cpp:34-42: WRAP_S = REPEAT, WRAP_T = CLAMP_TO_EDGE, MAG = LINEAR,
 * MIN = LINEAR_MIPMAP_LINEAR, TEXTURE_MAX_ANISOTROPY = the driver's maximum (16 on every GL 4.6 implementation).  What the GL
 * driver does with that is implementation-defined in its details; this header fixes the rule the GL 4.6 specification itself
 * spells out (section 8.14.1 "Scale Factor and Level of Detail", with the anisotropic sample pattern the specification gives as
 * its example) and is shared by the kernel (device), the oracle and the reference-shader shim (host):
 *
 *   Px = |(du/dx w, dv/dx h)|, Py = |(du/dy w, dv/dy h)|  (texels per pixel along the two screen axes),
 *   N  = min(ceil(Pmax / Pmin), max_anisotropy),  lambda = log2(Pmax / N), clamped to [0, q];
 *   lambda <= 0: one LINEAR sample of level 0 (magnification);
 *   otherwise N probes spread along the major axis, P + dP_major (i / (N + 1) - 1/2), i = 1..N, each a LINEAR_MIPMAP_LINEAR sample
 *   (bilinear on levels floor(lambda) and floor(lambda) + 1, blended with the fraction), averaged.
 */
#ifndef SKY_TEXGRAD_H
#define SKY_TEXGRAD_H

#include <math.h>

#include "sky_detmath.h"

#ifdef __CUDACC__
#define SKY_TEXGRAD_FN __host__ __device__ inline
#else
#define SKY_TEXGRAD_FN inline
#endif

/* GL 4.6 section 8.14.2: LINEAR on level `l` of a w0 x h0 texture (level size max(size >> l, 1)), REPEAT in s, CLAMP_TO_EDGE in t.
 * load(l, i, j) returns the (decoded) texel; V needs V + V and V * float. */
template <class V, class Load>
SKY_TEXGRAD_FN V sky_texgrad_bilinear(int w0, int h0, int l, float u, float v, Load load) {
    const int w = (w0 >> l) > 1 ? (w0 >> l) : 1, h = (h0 >> l) > 1 ? (h0 >> l) : 1;
    const float x = u * float(w) - 0.5f, y = v * float(h) - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float a = x - fx, b = y - fy;
    /* |u| stays far below 2^24 / w for texture coordinates: the conversions are exact */
    int i0 = int(fx) % w, i1;
    if (i0 < 0) i0 += w;
    i1 = i0 + 1 == w ? 0 : i0 + 1;
    int j0 = int(fy), j1 = j0 + 1;
    j0 = j0 < 0 ? 0 : j0 > h - 1 ? h - 1 : j0;
    j1 = j1 < 0 ? 0 : j1 > h - 1 ? h - 1 : j1;
    return load(l, i0, j0) * ((1.0f - a) * (1.0f - b)) + load(l, i1, j0) * (a * (1.0f - b)) + load(l, i0, j1) * ((1.0f - a) * b) + load(l, i1, j1) * (a * b);
}

template <class V, class Load>
SKY_TEXGRAD_FN V sky_texture_grad_2d(int w0, int h0, int levels, float u, float v, float dudx, float dvdx, float dudy, float dvdy, float max_anisotropy,
                                     Load load) {
    const float ax = dudx * float(w0), bx = dvdx * float(h0), ay = dudy * float(w0), by = dvdy * float(h0);
    const float Px = sqrtf(ax * ax + bx * bx), Py = sqrtf(ay * ay + by * by);
    const bool x_major = Px > Py;
    const float Pmax = x_major ? Px : Py, Pmin = x_major ? Py : Px;
    float N = 1.0f;
    if (Pmax > 0.0f) {
        N = Pmin > 0.0f ? ceilf(Pmax / Pmin) : max_anisotropy;
        N = N < 1.0f ? 1.0f : N > max_anisotropy ? max_anisotropy : N;
    }
    const float rho = Pmax / N;
    const float q = float(levels - 1);
    if (!(rho > 1.0f)) return sky_texgrad_bilinear<V>(w0, h0, 0, u, v, load);   /* lambda <= 0 (also rho == 0 and NaN): magnification */
    float lambda = sky_det_log2f(rho);   /* deterministic fp32 (include/sky_detmath.h): the same bits on the host and on the device */
    lambda = lambda > q ? q : lambda;
    const float fl = floorf(lambda), frac = lambda - fl;
    const int d1 = int(fl), d2 = d1 + 1 > levels - 1 ? levels - 1 : d1 + 1;
    const float du = x_major ? dudx : dudy, dv = x_major ? dvdx : dvdy;
    const int n = int(N);
    V sum = V();
    for (int i = 1; i <= n; ++i) {
        const float t = float(i) / float(n + 1) - 0.5f;
        const float pu = u + du * t, pv = v + dv * t;
        V tau = sky_texgrad_bilinear<V>(w0, h0, d1, pu, pv, load);
        if (frac > 0.0f) tau = tau * (1.0f - frac) + sky_texgrad_bilinear<V>(w0, h0, d2, pu, pv, load) * frac;
        sum = i == 1 ? tau : sum + tau;
    }
    return sum * (1.0f / float(n));
}

#endif /* SKY_TEXGRAD_H */
