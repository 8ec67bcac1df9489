// Cloud density materials: float SampleSigmaT(vec3 pos, float height01), the inner cost of K11, K16
// and K19.  Follows VolumetricCloudDefaultMaterial{Common,0,1}.glsl and
// VolumetricCloudMaterial{Minimal,Voxel}.glsl.
//
// Sampler semantics (VolumetricCloudDefaultMaterial.cpp:111-116, VolumetricCloudVoxelMaterial.cpp:30-37):
// mag = LINEAR, min = NEAREST_MIPMAP_NEAREST, explicit LOD.  GL 4.6 section 8.14.3: lambda <= 0.5 ->
// LINEAR on level 0; otherwise NEAREST on level ceil(lambda + 0.5) - 1.
//   HW == false: exact fp32 weights on texels read from device memory: one 8/16-byte load of a
//                corner-packed cell (MipView::cells) per LINEAR fetch, one texel load per NEAREST fetch.
//   HW == true : the texture unit filters (8-bit weights); one TEX per fetch.
// lambda = log2(k_lod * dist) + bias is only needed to pick the level: lambda <= 0.5 is decided by
// comparing dist^2 against a per-texture threshold, so the common (magnified) case costs no log2 / sqrt.
#pragma once
#include "common.cuh"
#include "context.h"

struct MaterialParams {
    SkyMaterialBlock m;
    MipView cloud_map, detail, displacement, voxel;
    float3 camera_pos;  // uCameraPos
    // (2^(0.5 - lod_bias) / k_lod)^2 per texture: dist^2 <= thr2  <=>  lambda <= 0.5
    float thr2_cloud_map, thr2_detail, thr2_displacement, thr2_voxel;
};

// spec 8.14.3 level selection; returns -1 for magnification
SKY_D int select_mip_level(float lod, int levels) {
    if (!(lod > 0.5f)) return -1;
    int q = levels - 1;
    int d = (lod <= float(q) + 0.5f) ? int(ceilf(lod + 0.5f)) - 1 : q;
    return clampi(d, 0, q);
}
// level for lambda = log2(k_lod * sqrt(d2)) + bias (GetUVWLod, VolumetricCloudDefaultMaterialCommon.glsl:20-24)
// Minified case: log2(k * sqrt(d2)) = 0.5 * log2(k^2 * d2), one MUFU.LG2 (2^-22 relative error moves a level
// boundary by less than a millimetre of camera distance).
SKY_D int level_from_distance2(float d2, float k_lod, float lod_bias, float thr2, int levels) {
    if (d2 <= thr2) return -1;
    return select_mip_level(0.5f * __log2f(k_lod * k_lod * d2) + lod_bias, levels);
}

// ---- exact path ---------------------------------------------------------------------------------------
// Conversion-free arithmetic.  I2F / F2I / FRND run on the quarter-rate conversion pipe, and a software trilinear
// fetch would need 14 of them (ncu: that pipe, not the FMA pipe, bounded K16 and K19).  Two exact replacements:
//   floor: for |x| < 2^22, round-down(x + 1.5*2^23) = 1.5*2^23 + floor(x) exactly, and the low mantissa bits of
//          that sum ARE floor(x) as an integer -> one FADD.RM, one FADD, one IADD;
//   byte -> float: PRMT drops byte n of a word into the mantissa of 2^23: the float 2^23 + byte.  Differences of
//          two such values are the exact byte differences, so a lerp needs one un-biasing FADD per pair.
// Results are bit-identical to floorf() / float(uint8_t).
constexpr float kFloorMagic = 12582912.0f;       // 1.5 * 2^23 = 0x4B400000
constexpr float kByteBias = 8388608.0f;          // 2^23       = 0x4B000000
SKY_D float floor_small(float x, int& i) {       // |x| < 2^22
    float m = __fadd_rd(x, kFloorMagic);
    i = __float_as_int(m) - 0x4B400000;
    return m - kFloorMagic;
}
SKY_D float byte_biased(uint32_t word, int n) { return __uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650u + n)); }  // 2^23 + byte n
SKY_D float lerp_biased(float t0b, float t1b, float a) { return (t0b - kByteBias) + a * (t1b - t0b); }  // t0 + a * (t1 - t0)
SKY_D float byte_to_float(uint32_t word, int n) { return byte_biased(word, n) - kByteBias; }

template <int C>
SKY_D void load_texel(const MipView& t, int level, int x, int y, int z, float* out) {
    const uint8_t* p = t.base + (t.off[level] + (size_t(z) * t.h[level] + y) * t.w[level] + x) * C;
    if (C == 1) {
        out[0] = byte_to_float(__ldg(p), 0) * (1.0f / 255.0f);
    } else if (C == 2) {
        uint32_t v = __ldg(reinterpret_cast<const unsigned short*>(p));
        out[0] = byte_to_float(v, 0) * (1.0f / 255.0f); out[1] = byte_to_float(v, 1) * (1.0f / 255.0f);
    } else {
        uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(p));
        out[0] = byte_to_float(v, 0) * (1.0f / 255.0f); out[1] = byte_to_float(v, 1) * (1.0f / 255.0f);
        out[2] = byte_to_float(v, 2) * (1.0f / 255.0f); out[3] = byte_to_float(v, 3) * (1.0f / 255.0f);
    }
}

// REPEAT textures have power-of-two sizes (512 / 128, fixed by the reference)
template <int C>
SKY_D void sample2d_repeat_exact(const MipView& t, float u, float v, int level, float* out) {
    if (level < 0) {
        int w = t.w[0], h = t.h[0];
        float x = u * float(w) - 0.5f, y = v * float(h) - 0.5f;
        int i0, j0;
        float fx = floor_small(x, i0), fy = floor_small(y, j0);
        float a = x - fx, b = y - fy;
        i0 &= w - 1; j0 &= h - 1;
        // one load brings the four corners: C == 2 -> 8 bytes {c00 c10 c01 c11} x {r,g}; C == 4 -> 16 bytes
        if (C == 2) {
            uint2 cell = __ldg(reinterpret_cast<const uint2*>(t.cells) + size_t(j0) * w + i0);
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                float r0 = lerp_biased(byte_biased(cell.x, c), byte_biased(cell.x, 2 + c), a);
                float r1 = lerp_biased(byte_biased(cell.y, c), byte_biased(cell.y, 2 + c), a);
                out[c] = (r0 + b * (r1 - r0)) * (1.0f / 255.0f);
            }
        } else {
            uint4 cell = __ldg(reinterpret_cast<const uint4*>(t.cells) + size_t(j0) * w + i0);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                float r0 = lerp_biased(byte_biased(cell.x, c), byte_biased(cell.y, c), a);
                float r1 = lerp_biased(byte_biased(cell.z, c), byte_biased(cell.w, c), a);
                out[c] = (r0 + b * (r1 - r0)) * (1.0f / 255.0f);
            }
        }
    } else {
        int w = t.w[level], h = t.h[level];
        int i, j;
        floor_small(u * float(w), i); floor_small(v * float(h), j);
        load_texel<C>(t, level, i & (w - 1), j & (h - 1), 0, out);
    }
}

// trilinear blend of the 8 corner bytes of a packed cell (byte n = di + 2*dj + 4*dk), x first
SKY_D float blend_cell(uint2 cell, float a, float b, float c) {
    float x00 = lerp_biased(byte_biased(cell.x, 0), byte_biased(cell.x, 1), a), x10 = lerp_biased(byte_biased(cell.x, 2), byte_biased(cell.x, 3), a);
    float x01 = lerp_biased(byte_biased(cell.y, 0), byte_biased(cell.y, 1), a), x11 = lerp_biased(byte_biased(cell.y, 2), byte_biased(cell.y, 3), a);
    float y0 = x00 + b * (x10 - x00), y1 = x01 + b * (x11 - x01);
    return (y0 + c * (y1 - y0)) * (1.0f / 255.0f);
}

SKY_D float sample3d_repeat_exact(const MipView& t, float u, float v, float w_, int level) {
    if (level < 0) {
        int w = t.w[0], h = t.h[0], d = t.d[0];
        float x = u * float(w) - 0.5f, y = v * float(h) - 0.5f, z = w_ * float(d) - 0.5f;
        int i0, j0, k0;
        float fx = floor_small(x, i0), fy = floor_small(y, j0), fz = floor_small(z, k0);
        i0 &= w - 1; j0 &= h - 1; k0 &= d - 1;
        uint2 cell = __ldg(reinterpret_cast<const uint2*>(t.cells) + (size_t(k0) * h + j0) * w + i0);
        return blend_cell(cell, x - fx, y - fy, z - fz);
    }
    int w = t.w[level], h = t.h[level], d = t.d[level];
    int i, j, k;
    floor_small(u * float(w), i); floor_small(v * float(h), j); floor_small(w_ * float(d), k);
    float out;
    load_texel<1>(t, level, i & (w - 1), j & (h - 1), k & (d - 1), &out);
    return out;
}

// CLAMP_TO_BORDER with border colour 0, any size (voxel grid).  The border test runs on the float floor, which
// stays far outside the grid for coordinates beyond the exact range of floor_small, so its integer is only used in range.
SKY_D float sample3d_border_exact(const MipView& t, float u, float v, float w_, int level) {
    if (level < 0) {
        int w = t.w[0], h = t.h[0], d = t.d[0];
        float x = u * float(w) - 0.5f, y = v * float(h) - 0.5f, z = w_ * float(d) - 0.5f;
        int i0, j0, k0;
        float fx = floor_small(x, i0), fy = floor_small(y, j0), fz = floor_small(z, k0);
        // cell index = base texel + 1; outside [0, w] x [0, h] x [0, d] every corner is the border
        float cx = fx + 1.0f, cy = fy + 1.0f, cz = fz + 1.0f;
        if (!(cx >= 0.0f && cx <= float(w) && cy >= 0.0f && cy <= float(h) && cz >= 0.0f && cz <= float(d))) return 0.0f;
        uint2 cell = __ldg(reinterpret_cast<const uint2*>(t.cells) + (size_t(k0 + 1) * t.cell_h + (j0 + 1)) * t.cell_w + (i0 + 1));
        return blend_cell(cell, x - fx, y - fy, z - fz);
    }
    int w = t.w[level], h = t.h[level], d = t.d[level];
    float fu = u * float(w), fv = v * float(h), fw = w_ * float(d);
    float out = 0.0f;
    if (fu >= 0.0f && fu < float(w) && fv >= 0.0f && fv < float(h) && fw >= 0.0f && fw < float(d)) {
        int i, j, k;
        floor_small(fu, i); floor_small(fv, j); floor_small(fw, k);
        load_texel<1>(t, level, i, j, k, &out);
    }
    return out;
}

// Two-phase form of the level-0 LINEAR lookup above (same arithmetic): address and weights first, the
// 8-byte load second, the blend third, so a caller can keep several independent lookups in flight per lane.
struct VoxelTap {
    long long cell;  // index into MipView::cells; kVoxelTapBorder: every corner is the border
    float a, b, c;
};
constexpr long long kVoxelTapBorder = -1;
SKY_D VoxelTap voxel_tap(const MipView& t, float u, float v, float w_) {
    int w = t.w[0], h = t.h[0], d = t.d[0];
    float x = u * float(w) - 0.5f, y = v * float(h) - 0.5f, z = w_ * float(d) - 0.5f;
    int i0, j0, k0;
    float fx = floor_small(x, i0), fy = floor_small(y, j0), fz = floor_small(z, k0);
    float cx = fx + 1.0f, cy = fy + 1.0f, cz = fz + 1.0f;
    VoxelTap tap;
    tap.a = x - fx; tap.b = y - fy; tap.c = z - fz;
    bool inside = cx >= 0.0f && cx <= float(w) && cy >= 0.0f && cy <= float(h) && cz >= 0.0f && cz <= float(d);
    tap.cell = inside ? ((long long)(k0 + 1) * t.cell_h + (j0 + 1)) * t.cell_w + (i0 + 1) : kVoxelTapBorder;
    return tap;
}
SKY_D uint2 voxel_tap_load(const MipView& t, const VoxelTap& tap) {  // always a valid address: cell 0 stands in for the border
    return __ldg(reinterpret_cast<const uint2*>(t.cells) + (tap.cell == kVoxelTapBorder ? 0ll : tap.cell));
}
SKY_D float voxel_tap_blend(const VoxelTap& tap, uint2 cell) {
    return tap.cell == kVoxelTapBorder ? 0.0f : blend_cell(cell, tap.a, tap.b, tap.c);
}

// ---- hardware path ----------------------------------------------------------------------------------
SKY_D float4 sample2d_hw(const MipView& t, float u, float v, int level) {
    return level < 0 ? tex2DLod<float4>(t.tex_linear, u, v, 0.0f) : tex2DLod<float4>(t.tex_point, u, v, float(level));
}
SKY_D float sample3d_hw(const MipView& t, float u, float v, float w, int level) {
    return level < 0 ? tex3DLod<float>(t.tex_linear, u, v, w, 0.0f) : tex3DLod<float>(t.tex_point, u, v, w, float(level));
}

// ---- SampleSigmaT ------------------------------------------------------------------------------------
// VolumetricCloudDefaultMaterial0.glsl:9-16
SKY_D float CalHeightMask(float cloud_type, float height01) {
    float height_in_type = clampf(__fdividef(height01, cloud_type), 0.0f, 1.0f);  // same inf / NaN cases as '/'
    return clampf(height_in_type * (height_in_type - 1.0f) * -4.0f, 0.0f, 1.0f);
}
SKY_D float Remap01(float x, float x0, float x1) { return clampf(__fdividef(x - x0, x1 - x0), 0.0f, 1.0f); }
SKY_D float distance2(float3 a, float3 b) { float3 d = a - b; return dot(d, d); }

// MAT: SkyMaterialType.  `fetches` (optional) receives the number of texture fetches issued.
template <int MAT, bool HW>
SKY_D float SampleSigmaT(const MaterialParams& M, float3 pos, float height01, int* fetches = nullptr) {
    if (MAT == SKY_MATERIAL_DEFAULT0) {  // VolumetricCloudDefaultMaterial0.glsl:18-32
        const SkyMaterialCommonBufferData& mc = M.m.common;
        const SkyMaterial0BufferData& m = M.m.u.m0;
        // GetUVWLod (VolumetricCloudDefaultMaterialCommon.glsl:20-24) for the cloud map and the displacement map: same pos
        float d2 = distance2(pos, M.camera_pos);
        const SkySampleInfo& ci = mc.uCloudMapSampleInfo;
        const SkySampleInfo& di = mc.uDisplacementSampleInfo;
        int lc = level_from_distance2(d2, ci.k_lod, mc.uLodBias, M.thr2_cloud_map, M.cloud_map.levels);
        int ld = level_from_distance2(d2, di.k_lod, mc.uLodBias, M.thr2_displacement, M.displacement.levels);
        float cu = pos.x * ci.frequency + ci.bias[0], cv = pos.y * ci.frequency + ci.bias[1];
        float du = pos.x * di.frequency + di.bias[0], dv = pos.y * di.frequency + di.bias[1], dw = pos.z * di.frequency;
        float cloud_type[2];
        float d0[4], d1[4];
        if (HW) {
            float4 c = sample2d_hw(M.cloud_map, cu, cv, lc);
            cloud_type[0] = c.x; cloud_type[1] = c.y;
            float4 a = sample2d_hw(M.displacement, du, dv, ld);
            float4 b = sample2d_hw(M.displacement, du, dw, ld);
            d0[0] = a.x; d0[1] = a.y; d1[2] = b.z; d1[3] = b.w;
        } else {
            sample2d_repeat_exact<2>(M.cloud_map, cu, cv, lc, cloud_type);
            sample2d_repeat_exact<4>(M.displacement, du, dv, ld, d0);
            sample2d_repeat_exact<4>(M.displacement, du, dw, ld, d1);
        }
        float3 displace_vector = f3(0.0f + d0[0] + d1[2], 0.0f + d0[1], 0.0f + d1[3]);
        pos = pos + m.uDisplacementScale * displace_vector;
        const SkySampleInfo& ti = mc.uDetailSampleInfo;
        int lt = level_from_distance2(distance2(pos, M.camera_pos), ti.k_lod, mc.uLodBias, M.thr2_detail, M.detail.levels);
        float tu = pos.x * ti.frequency + ti.bias[0], tv = pos.y * ti.frequency + ti.bias[1], tw = pos.z * ti.frequency;
        float detail = HW ? sample3d_hw(M.detail, tu, tv, tw, lt) : sample3d_repeat_exact(M.detail, tu, tv, tw, lt);
        detail = detail * m.uDetailParam[0] + m.uDetailParam[1];
        if (fetches) *fetches += 4;
        return Remap01(cloud_type[0] * CalHeightMask(cloud_type[1], height01), detail, 1.0f) * height01 * mc.uDensity;
    } else if (MAT == SKY_MATERIAL_DEFAULT1) {  // VolumetricCloudDefaultMaterial1.glsl:14-29
        const SkyMaterialCommonBufferData& mc = M.m.common;
        const SkyMaterial1BufferData& m = M.m.u.m1;
        float d2 = distance2(pos, M.camera_pos);
        const SkySampleInfo& ci = mc.uCloudMapSampleInfo;
        int lc = level_from_distance2(d2, ci.k_lod, mc.uLodBias, M.thr2_cloud_map, M.cloud_map.levels);
        float cu = pos.x * ci.frequency + ci.bias[0], cv = pos.y * ci.frequency + ci.bias[1];
        float cloud_type[2];
        if (HW) {
            float4 c = sample2d_hw(M.cloud_map, cu, cv, lc);
            cloud_type[0] = c.x; cloud_type[1] = c.y;
        } else {
            sample2d_repeat_exact<2>(M.cloud_map, cu, cv, lc, cloud_type);
        }
        if (fetches) *fetches += 1;
        float density = clampf((cloud_type[0] - m.uBaseDensityThreshold) * m.uBaseEdgeHardness, 0.0f, 1.0f);
        density *= clampf((1 - height01) * m.uBaseHeightHardness, 0.0f, 1.0f);
        if (density == 0) return 0.0f;
        const SkySampleInfo& ti = mc.uDetailSampleInfo;
        int lt = level_from_distance2(d2, ti.k_lod, mc.uLodBias, M.thr2_detail, M.detail.levels);
        float tu = pos.x * ti.frequency + ti.bias[0], tv = pos.y * ti.frequency + ti.bias[1], tw = pos.z * ti.frequency;
        float detail = HW ? sample3d_hw(M.detail, tu, tv, tw, lt) : sample3d_repeat_exact(M.detail, tu, tv, tw, lt);
        if (fetches) *fetches += 1;
        detail = (detail + m.uDetailBase) * m.uDetailScale;
        detail *= fmaxf(clampf(height01 - m.uHeightCut, 0.0f, 1.0f), clampf(m.uEdgeCur - cloud_type[0], 0.0f, 1.0f));
        return clampf(density - detail, 0.0f, 1.0f) * mc.uDensity * height01;
    } else if (MAT == SKY_MATERIAL_MINIMAL) {  // VolumetricCloudMaterialMinimal.glsl:6-8
        return M.m.u.minimal.uDensity;
    } else {  // VolumetricCloudMaterialVoxel.glsl:12-17
        const SkyMaterialVoxelBufferData& m = M.m.u.voxel;
        float u = pos.x * m.uSampleFrequency[0] + m.uSampleBias[0];
        float v = pos.y * m.uSampleFrequency[1] + m.uSampleBias[1];
        int level = level_from_distance2(distance2(pos, M.camera_pos), m.uSampleLodK, m.uLodBias, M.thr2_voxel, M.voxel.levels);
        float density = HW ? sample3d_hw(M.voxel, u, v, height01, level) : sample3d_border_exact(M.voxel, u, v, height01, level);
        if (fetches) *fetches += 1;
        return density * m.uDensity;
    }
}

// Call `f.template operator()<MAT, HW>()` for the runtime (material type, filtering) pair.
template <class F>
inline int dispatch_material(int type, bool hw, F&& f) {
#define SKY_CASE(T)                                              \
    case T: return hw ? f.template operator()<T, true>() : f.template operator()<T, false>();
    switch (type) {
        SKY_CASE(SKY_MATERIAL_DEFAULT0)
        SKY_CASE(SKY_MATERIAL_DEFAULT1)
        SKY_CASE(SKY_MATERIAL_MINIMAL)
        SKY_CASE(SKY_MATERIAL_VOXEL)
    }
#undef SKY_CASE
    return -1;
}

int make_material_params(SkyContext* ctx, const float* camera_pos, MaterialParams& M);
