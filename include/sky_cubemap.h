/*
 * sky_cubemap.h -- seamless cube-map filtering (GL 4.6 section 8.14.1, "Seamless Cube Map Filtering").
 *
 * The reference enables GL_TEXTURE_CUBE_MAP_SEAMLESS globally before it creates the environment cube
 * (src/SkyRendering/AtmosphereRenderer.cpp:151), so every LINEAR fetch of a cube in K23 / K24 (EnvRadianceSH.comp,
 * PrefilterRadiance.comp), in the object branch of K6 (BRDF.glsl GetAmbient) and in the path tracer's SampleEnvironment
 * (VolumetricCloudPathTracing.comp:206,215) takes the taps that fall off a face from the ADJACENT face, and a tap that falls
 * off in both directions (a cube corner) is the average of the three taps that exist.  This header is that rule, shared by the
 * kernels (device), the oracle and the reference-shader shim (host) so that the three cannot drift apart.
 *
 * Face order and (sc, tc) orientation: GL 4.6 table 8.19 (+X, -X, +Y, -Y, +Z, -Z).
 */
#ifndef SKY_CUBEMAP_H
#define SKY_CUBEMAP_H

#ifdef __CUDACC__
#define SKY_CUBE_FN __host__ __device__ inline
#else
#define SKY_CUBE_FN inline
#endif

/* (face, i, j) with exactly ONE of i, j outside [0, n) by one texel -> the texel across that edge on the neighbouring face.
 * Integer arithmetic on the cube [-n, n]^3 in units of 1 / n (texel centres are at odd - n offsets): the tap is moved onto the
 * shared edge and half a texel down the neighbouring face, then re-read through table 8.19. */
SKY_CUBE_FN void sky_cube_adjacent(int n, int* face, int* i, int* j) {
    int sc = 2 * *i + 1 - n, tc = 2 * *j + 1 - n, ma = n;
    if (*i < 0) { sc = -n; ma = n - 1; } else if (*i >= n) { sc = n; ma = n - 1; }
    if (*j < 0) { tc = -n; ma = n - 1; } else if (*j >= n) { tc = n; ma = n - 1; }
    int x, y, z;
    switch (*face) {
        case 0: x = ma; y = -tc; z = -sc; break;
        case 1: x = -ma; y = -tc; z = sc; break;
        case 2: x = sc; y = ma; z = tc; break;
        case 3: x = sc; y = -ma; z = -tc; break;
        case 4: x = sc; y = -tc; z = ma; break;
        default: x = -sc; y = -tc; z = -ma; break;
    }
    int g, s, t;
    if (x == n) { g = 0; s = -z; t = -y; }
    else if (x == -n) { g = 1; s = z; t = -y; }
    else if (y == n) { g = 2; s = x; t = z; }
    else if (y == -n) { g = 3; s = x; t = -z; }
    else if (z == n) { g = 4; s = x; t = -y; }
    else { g = 5; s = -x; t = -y; }
    *face = g;
    *i = (s + n - 1) / 2;
    *j = (t + n - 1) / 2;
}

/* Bilinear blend of the 2x2 footprint at base texel (i0, j0) of `face`, weights (a, b), i0 / j0 in [-1, n - 1]:
 * load(face, i, j) reads an in-range texel of any face.  V needs V + V and V * float. */
template <class V, class Load>
SKY_CUBE_FN V sky_cube_bilinear(int n, int face, int i0, int j0, float a, float b, Load load) {
    const int i1 = i0 + 1, j1 = j0 + 1;
    const bool out_i0 = i0 < 0, out_i1 = i1 >= n, out_j0 = j0 < 0, out_j1 = j1 >= n;
    V t00, t10, t01, t11;
    if (!(out_i0 || out_i1 || out_j0 || out_j1)) {
        t00 = load(face, i0, j0); t10 = load(face, i1, j0); t01 = load(face, i0, j1); t11 = load(face, i1, j1);
    } else {
        auto tap = [&](int i, int j, bool out_i, bool out_j, bool* corner) {
            *corner = out_i && out_j;
            if (*corner) return load(face, out_i ? (i < 0 ? 0 : n - 1) : i, out_j ? (j < 0 ? 0 : n - 1) : j);   // placeholder, replaced below
            int f = face;
            if (out_i || out_j) sky_cube_adjacent(n, &f, &i, &j);
            return load(f, i, j);
        };
        bool c00, c10, c01, c11;
        t00 = tap(i0, j0, out_i0, out_j0, &c00); t10 = tap(i1, j0, out_i1, out_j0, &c10);
        t01 = tap(i0, j1, out_i0, out_j1, &c01); t11 = tap(i1, j1, out_i1, out_j1, &c11);
        /* a cube corner has no fourth texel: the average of the three that exist (the spec's recommended construction) */
        const float third = 1.0f / 3.0f;
        if (c00) t00 = ((t11 + t10) + t01) * third;
        else if (c10) t10 = ((t01 + t00) + t11) * third;
        else if (c01) t01 = ((t10 + t00) + t11) * third;
        else if (c11) t11 = ((t00 + t10) + t01) * third;
    }
    return t00 * ((1.0f - a) * (1.0f - b)) + t10 * (a * (1.0f - b)) + t01 * ((1.0f - a) * b) + t11 * (a * b);
}

#endif /* SKY_CUBEMAP_H */
