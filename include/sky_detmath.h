/*
 * sky_detmath.h -- deterministic fp32 elementary functions for the atmosphere LUT bake.
 *
 * GLSL leaves the accuracy of exp/sin/cos/acos to the driver, and the LUT arithmetic amplifies a
 * one-ulp difference without bound.  These versions are written with plain IEEE
 * fp32 +, -, *, /, sqrt only, in a fixed order, so that a CPU build with -ffp-contract=off and a CUDA
 * build with -fmad=false return bit-identical results: the LUT kernels and the oracle can then be
 * compared bit for bit.  Accuracy is the usual 1-2 ulp of a minimax polynomial after Cody-Waite
 * range reduction (checked against double precision in tests/test_oracle_kat.py).
 * This header is neutral infrastructure: it contains no part of the rendering algorithm.
 */
#ifndef SKY_DETMATH_H
#define SKY_DETMATH_H

#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define SKY_DM __host__ __device__ __forceinline__
#else
#define SKY_DM inline
#include <math.h>
#endif

SKY_DM float sky_bits_to_float(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
SKY_DM uint32_t sky_float_to_bits(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}

/* exp(x): k = round(x / ln2), r = x - k ln2 in two pieces, Pade-style kernel, scale by 2^k. */
SKY_DM float sky_det_expf(float x) {
    if (x != x) return x;
    if (x > 88.72f) return sky_bits_to_float(0x7f800000u);
    if (x < -103.9f) return 0.0f;
    const float ln2hi = 6.9314575195e-01f, ln2lo = 1.4286067653e-06f, invln2 = 1.4426950216e+00f;
    const float P1 = 1.6666625440e-01f, P2 = -2.7667332906e-03f;
    int k = (int)(invln2 * x + (x < 0.0f ? -0.5f : 0.5f));
    float fk = (float)k;
    float hi = x - fk * ln2hi;
    float lo = fk * ln2lo;
    float r = hi - lo;
    float rr = r * r;
    float c = r - rr * (P1 + rr * P2);
    float y = 1.0f + (r * c / (2.0f - c) - lo + hi);
    if (k >= -125) return y * sky_bits_to_float((uint32_t)(k + 127) << 23);
    /* subnormal results: two-step scaling */
    return y * sky_bits_to_float((uint32_t)(k + 64 + 127) << 23) * sky_bits_to_float((uint32_t)(127 - 64) << 23);
}

/* sin and cos together: quadrant n = round(x * 2/pi), r = x - n pi/2 (two-piece), kernels on [-pi/4, pi/4].
 * Meant for the angles of the bake (|x| up to a few pi). */
SKY_DM void sky_det_sincosf(float x, float* s_out, float* c_out) {
    const float two_over_pi = 6.3661977237e-01f, pio2_1 = 1.5707855225e+00f, pio2_1t = 1.0804334124e-05f;
    int n = (int)(x * two_over_pi + (x < 0.0f ? -0.5f : 0.5f));
    float fn = (float)n;
    float r = (x - fn * pio2_1) - fn * pio2_1t;
    float z = r * r;
    const float S1 = -1.6666667163e-01f, S2 = 8.3333337680e-03f, S3 = -1.9841270114e-04f, S4 = 2.7557314297e-06f, S5 = -2.5050759689e-08f;
    const float C1 = 4.1666667908e-02f, C2 = -1.3888889225e-03f, C3 = 2.4801587642e-05f, C4 = -2.7557314297e-07f, C5 = 2.0875723372e-09f;
    float s = r + r * z * (S1 + z * (S2 + z * (S3 + z * (S4 + z * S5))));
    float c = 1.0f - (0.5f * z - z * z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * C5)))));
    switch (n & 3) {
        case 0: *s_out = s; *c_out = c; break;
        case 1: *s_out = c; *c_out = -s; break;
        case 2: *s_out = -s; *c_out = -c; break;
        default: *s_out = -c; *c_out = s; break;
    }
}
SKY_DM float sky_det_sinf(float x) { float s, c; sky_det_sincosf(x, &s, &c); return s; }
SKY_DM float sky_det_cosf(float x) { float s, c; sky_det_sincosf(x, &s, &c); return c; }

/* acos(x), |x| <= 1 (callers clamp): rational kernel R(z) on z = x^2 or (1 -+ x)/2, as in the classic
 * Sun libm split into |x| < 0.5, x < -0.5, x > 0.5. */
SKY_DM float sky_det_acos_r(float z) {
    const float pS0 = 1.6666586697e-01f, pS1 = -4.2743422091e-02f, pS2 = -8.6563630030e-03f, qS1 = -7.0662963390e-01f;
    float p = z * (pS0 + z * (pS1 + z * pS2));
    float q = 1.0f + z * qS1;
    return p / q;
}
SKY_DM float sky_det_acosf(float x) {
    const float pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    if (x >= 1.0f) return 0.0f;
    if (x <= -1.0f) return 2.0f * pio2_hi;
    float ax = x < 0.0f ? -x : x;
    if (ax < 0.5f) {
        if (ax < 3.7252903e-09f) return pio2_hi;
        return pio2_hi - (x - (pio2_lo - x * sky_det_acos_r(x * x)));
    }
    if (x < 0.0f) {
        float z = (1.0f + x) * 0.5f;
        float s = sqrtf(z);
        float w = sky_det_acos_r(z) * s - pio2_lo;
        return 2.0f * (pio2_hi - (s + w));
    }
    float z = (1.0f - x) * 0.5f;
    float s = sqrtf(z);
    float df = sky_bits_to_float(sky_float_to_bits(s) & 0xfffff000u);
    float c = (z - df * df) / (s + df);
    float w = sky_det_acos_r(z) * s + c;
    return 2.0f * (df + w);
}

/* asin(x), |x| <= 1 (callers clamp): the same rational kernel; |x| < 0.5: x + x R(x^2); else pi/2 - 2 (s + s R(z)),
 * z = (1 - |x|)/2, s = sqrt(z)  (GetVisibilityFromMoonShadow, Atmosphere.glsl:213). */
SKY_DM float sky_det_asinf(float x) {
    const float pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
    float ax = x < 0.0f ? -x : x;
    if (ax >= 1.0f) return x < 0.0f ? -(pio2_hi + pio2_lo) : (pio2_hi + pio2_lo);
    if (ax < 0.5f) {
        if (ax < 2.4414062e-04f) return x;
        return x + x * sky_det_acos_r(x * x);
    }
    float z = (1.0f - ax) * 0.5f;
    float s = sqrtf(z);
    float t = pio2_hi - (2.0f * (s + s * sky_det_acos_r(z)) - pio2_lo);
    return x < 0.0f ? -t : t;
}

/* atan(x), any finite x: argument reduction to |t| < 0.4375 around atan(0.5), atan(1), atan(1.5), atan(inf) and an odd polynomial
 * of degree 23 (the classic four-interval scheme); atan2(y, x) from it with the usual quadrant fix-up (GetEarthAlbedo,
 * EarthRender.frag:25; atan2(0, 0) = 0 here, GLSL: undefined). */
SKY_DM float sky_det_atanf(float x) {
    const float hi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
    const float lo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
    float ax = x < 0.0f ? -x : x;
    if (!(ax < 6.7108864e7f)) return x != x ? x : (x < 0.0f ? -(hi[3] + lo[3]) : (hi[3] + lo[3]));
    int id;
    float t;
    if (ax < 0.4375f) {
        if (ax < 2.4414062e-04f) return x;
        id = -1; t = x;
    } else if (ax < 1.1875f) {
        if (ax < 0.6875f) { id = 0; t = (2.0f * ax - 1.0f) / (2.0f + ax); }
        else { id = 1; t = (ax - 1.0f) / (ax + 1.0f); }
    } else if (ax < 2.4375f) { id = 2; t = (ax - 1.5f) / (1.0f + 1.5f * ax); }
    else { id = 3; t = -1.0f / ax; }
    float z = t * t, w = z * z;
    float s1 = z * (3.3333334327e-01f + w * (1.4285714924e-01f + w * (9.0908870101e-02f + w * (6.6610731184e-02f + w * (4.9768779427e-02f + w * 1.6285819933e-02f)))));
    float s2 = w * (-2.0000000298e-01f + w * (-1.1111110449e-01f + w * (-7.6918758452e-02f + w * (-5.8335702866e-02f + w * -3.6531571299e-02f))));
    if (id < 0) return t - t * (s1 + s2);
    float r = hi[id] - ((t * (s1 + s2) - lo[id]) - t);
    return x < 0.0f ? -r : r;
}
SKY_DM float sky_det_atan2f(float y, float x) {
    const float pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f, pio2 = 1.5707963705e+00f;
    if (x != x || y != y) return x + y;
    if (y == 0.0f) return x < 0.0f || (x == 0.0f && sky_float_to_bits(x) >> 31) ? (sky_float_to_bits(y) >> 31 ? -pi : pi) : y;
    if (x == 0.0f) return y < 0.0f ? -pio2 : pio2;
    float q = y / x;
    float z = sky_det_atanf(q < 0.0f ? -q : q);   /* [0, pi/2] */
    if (x > 0.0f) return y < 0.0f ? -z : z;
    return y < 0.0f ? (z - pi_lo) - pi : pi - (z - pi_lo);
}

/* log2(x) for positive, finite, normal x (the LOD of sky_texgrad.h): x = m 2^e with m in [sqrt(1/2), sqrt(2)), log(m) from the
 * classic s = f / (2 + f) series, then e + log(m) / ln 2. */
SKY_DM float sky_det_log2f(float x) {
    uint32_t ix = sky_float_to_bits(x);
    ix += 0x3f800000u - 0x3f3504f3u;                       /* m in [sqrt(1/2), sqrt(2)) */
    const int e = int(ix >> 23) - 127;
    const float m = sky_bits_to_float((ix & 0x007fffffu) + 0x3f3504f3u);
    const float f = m - 1.0f;
    const float s = f / (2.0f + f);
    const float z = s * s, w = z * z;
    const float t1 = w * (0.40000972152f + w * 0.24279078841f);
    const float t2 = z * (0.66666662693f + w * 0.28498786688f);
    const float hfsq = 0.5f * f * f;
    const float lg = f - (hfsq - s * (hfsq + (t2 + t1)));  /* log(m) */
    return float(e) + lg * 1.4426950216e+00f + lg * 1.9259629891e-08f;
}

/* x^1.5 for x >= 0 (MiePhaseFunction, Atmosphere.glsl:150): x * sqrt(x) */
SKY_DM float sky_det_pow15f(float x) { return x * sqrtf(x); }

#endif /* SKY_DETMATH_H */
