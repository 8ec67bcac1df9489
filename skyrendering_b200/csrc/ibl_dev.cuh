// Cube-map sampling shared by the IBL kernels (ibl.cu) and the object branch of the composite (atmosphere.cu, K6).
// Conventions of oracle/ibl.h: GL 4.6 table 8.19 face selection, bilinear inside the face clamped at its edge,
// LINEAR_MIPMAP_LINEAR blend t0 * (1 - f) + t1 * f with the exact fp32 fraction of the clamped LOD.
#pragma once
#include "common.cuh"

namespace {

struct CubeChainView {
    const half4* level[12];  // level l: [6][n >> l][n >> l]
    int n;
    int levels;
};

// GL 4.6 table 8.19 face selection + bilinear inside the face, clamped at its edge (oracle/ibl.cpp TextureCubeLevel)
SKY_D float4 TextureCubeLevel(const half4* lvl, int n, float3 dir) {
    float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    int face; float sc, tc, ma;
    if (ax >= ay && ax >= az) { ma = ax; if (dir.x >= 0) { face = 0; sc = -dir.z; tc = -dir.y; } else { face = 1; sc = dir.z; tc = -dir.y; } }
    else if (ay >= az)        { ma = ay; if (dir.y >= 0) { face = 2; sc = dir.x; tc = dir.z; } else { face = 3; sc = dir.x; tc = -dir.z; } }
    else                      { ma = az; if (dir.z >= 0) { face = 4; sc = dir.x; tc = -dir.y; } else { face = 5; sc = -dir.x; tc = -dir.y; } }
    float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
    float u = s * float(n) - 0.5f, v = t * float(n) - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    int i0 = int(fu), j0 = int(fv);
    float a = u - fu, b = v - fv;
    const half4* f = lvl + size_t(face) * n * n;
    int x0 = clampi(i0, 0, n - 1), x1 = clampi(i0 + 1, 0, n - 1), y0 = clampi(j0, 0, n - 1), y1 = clampi(j0 + 1, 0, n - 1);
    float4 t00 = load_half4(f + y0 * n + x0), t10 = load_half4(f + y0 * n + x1), t01 = load_half4(f + y1 * n + x0), t11 = load_half4(f + y1 * n + x1);
    return (1.0f - a) * (1.0f - b) * t00 + a * (1.0f - b) * t10 + (1.0f - a) * b * t01 + a * b * t11;
}

// textureLod(samplerCube, dir, lod), LINEAR_MIPMAP_LINEAR (oracle/ibl.cpp TextureCubeLod)
SKY_D float4 TextureCubeLod(const CubeChainView& c, float3 dir, float lod) {
    const int q = c.levels - 1;
    float l = lod < 0.0f ? 0.0f : lod > float(q) ? float(q) : lod;
    float fl = floorf(l);
    int l0 = int(fl);
    float f = l - fl;
    float4 t0 = TextureCubeLevel(c.level[l0], c.n >> l0, dir);
    if (!(f > 0.0f)) return t0;
    int l1 = l0 + 1 > q ? q : l0 + 1;
    float4 t1 = TextureCubeLevel(c.level[l1], c.n >> l1, dir);
    return t0 * (1.0f - f) + t1 * f;
}

}  // namespace
