// Cube-map sampling shared by the IBL kernels (ibl.cu) and the object branch of the composite (atmosphere.cu, K6).
// Conventions of oracle/ibl.h: GL 4.6 table 8.19 face selection, SEAMLESS bilinear filtering (section 8.14.1; the reference enables
// GL_TEXTURE_CUBE_MAP_SEAMLESS, AtmosphereRenderer.cpp:151: taps off a face come from the adjacent face, include/sky_cubemap.h),
// LINEAR_MIPMAP_LINEAR blend t0 * (1 - f) + t1 * f with the exact fp32 fraction of the clamped LOD.
#pragma once
#include "../../include/sky_cubemap.h"
#include "common.cuh"

namespace {

struct CubeChainView {
    const half4* level[12];  // level l: [6][n >> l][n >> l]
    int n;
    int levels;
};

// GL 4.6 table 8.19 face selection + seamless bilinear filtering (oracle/ibl.cpp TextureCubeLevel)
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
    return sky_cube_bilinear<float4>(n, face, i0, j0, a, b, [&](int f, int i, int j) { return load_half4(lvl + (size_t(f) * n + j) * n + i); });
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
