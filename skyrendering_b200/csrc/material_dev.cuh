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

// Approximate division / log2 in the production objects, IEEE ones in the strict objects (see cloud.cu)
#ifdef SKY_STRICT_TU
#define SKY_FDIV(a, b) ((a) / (b))
#define SKY_LOG2(x) log2f(x)
#else
#define SKY_FDIV(a, b) __fdividef(a, b)
#define SKY_LOG2(x) __log2f(x)
#endif

struct MaterialParams {
    SkyMaterialBlock m;
    MipView cloud_map, detail, displacement, voxel;
    float3 camera_pos;  // uCameraPos
    // (2^(0.5 - lod_bias) / k_lod)^2 per texture: dist^2 <= thr2  <=>  lambda <= 0.5
    float thr2_cloud_map, thr2_detail, thr2_displacement, thr2_voxel;
    // 0.5 - (log2(k_lod) + lod_bias) per texture: lambda <= 0.5  <=>  lod_h - 0.5 * log2(dist^2) >= 0
    float lod_h_cloud_map, lod_h_detail, lod_h_displacement, lod_h_voxel;
    // DEFAULT0: 0 <= uDetailParam.x * detail + uDetailParam.y <= 1 for every texel value, see SigmaEval::advance
    bool m0_zero_base_is_zero;
};

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

// GetUVWLod (VolumetricCloudDefaultMaterialCommon.glsl:20-24) + the GL level rule from L = log2(dist^2), which the
// textures sampled at one position share: lambda = 0.5 * L + log2(k_lod) + bias.  x = 0.5 - lambda >= 0: LINEAR on
// level 0 (-1); otherwise NEAREST on level min(ceil(lambda + 0.5) - 1, q) = min(ceil(-x), q) = min(-floor(x), q).
// (The 2^-22 relative error of MUFU.LG2 moves a level boundary by less than a millimetre of camera distance.)
SKY_D int level_from_log2(float log2_d2, float lod_h, int levels) {
    float x = lod_h - 0.5f * log2_d2;
    int d = 0x4B400000 - __float_as_int(__fadd_rd(x, kFloorMagic));  // -floor(x); garbage-but-large for |x| >= 2^22
    return !(x < 0.0f) ? -1 : min(d, levels - 1);
}

// Hardware path: the same decision without the integer level -- `mag` (x >= 0: LINEAR on level 0) and the LOD as the float -floor(x),
// which the POINT-mip texture object clamps to its last level (maxMipmapLevelClamp, noise.cu) -- 4 instructions instead of 8 + I2F.
struct HwLod { bool mag; float lod; };
SKY_D HwLod hw_lod_from_log2(float log2_d2, float lod_h) {
    const float x = lod_h - 0.5f * log2_d2;
    HwLod r;
    r.mag = !(x < 0.0f);
    r.lod = kFloorMagic - __fadd_rd(x, kFloorMagic);   // -floor(x), exact for |x| < 2^22
    return r;
}

// ---- unified fetches ------------------------------------------------------------------------------------
// One fetch = one tap: cell index + weights.  `level` < 0: LINEAR on level 0 (GL magnification); otherwise NEAREST
// on that mip level, expressed as the same cell load with zero weights (corner 0 of cell (i,j,k) is texel (i,j,k)),
// so lanes of a warp that sit on different sides of the LOD threshold run the same instructions.
struct Tap2 { unsigned int cell; float a, b; };
struct Tap3 { unsigned int cell; float a, b, c; };
struct TapB { long long cell; float a, b, c; };  // BORDER volume (any size); cell < 0: every corner is the border

// REPEAT textures have power-of-two sizes (512 / 128, fixed by the reference): level dims are shifts
SKY_D Tap2 tap2_repeat(const MipView& t, float u, float v, int level) {
    const bool mag = level < 0;
    const int l = mag ? 0 : level;
    const int w = max(t.w[0] >> l, 1), h = max(t.h[0] >> l, 1);
    const float half = mag ? 0.5f : 0.0f;
    float x = u * float(w) - half, y = v * float(h) - half;
    int i0, j0;
    float fx = floor_small(x, i0), fy = floor_small(y, j0);
    Tap2 tap;
    tap.a = mag ? x - fx : 0.0f;
    tap.b = mag ? y - fy : 0.0f;
    tap.cell = (unsigned int)t.cell_off[l] + (unsigned int)((j0 & (h - 1)) * w + (i0 & (w - 1)));
    return tap;
}
SKY_D Tap3 tap3_repeat(const MipView& t, float u, float v, float w_, int level) {
    const bool mag = level < 0;
    const int l = mag ? 0 : level;
    const int w = max(t.w[0] >> l, 1), h = max(t.h[0] >> l, 1), d = max(t.d[0] >> l, 1);
    const float half = mag ? 0.5f : 0.0f;
    float x = u * float(w) - half, y = v * float(h) - half, z = w_ * float(d) - half;
    int i0, j0, k0;
    float fx = floor_small(x, i0), fy = floor_small(y, j0), fz = floor_small(z, k0);
    Tap3 tap;
    tap.a = mag ? x - fx : 0.0f;
    tap.b = mag ? y - fy : 0.0f;
    tap.c = mag ? z - fz : 0.0f;
    tap.cell = (unsigned int)t.cell_off[l] + (unsigned int)(((k0 & (d - 1)) * h + (j0 & (h - 1))) * w + (i0 & (w - 1)));
    return tap;
}
// CLAMP_TO_BORDER with border colour 0, any size (voxel grid).  The range test runs on the float floor, which stays
// far outside the grid for coordinates beyond the exact range of floor_small, so its integer is only used in range.
SKY_D TapB tapb_border(const MipView& t, float u, float v, float w_, int level) {
    const bool mag = level < 0;
    const int l = mag ? 0 : level;
    const int w = max(t.w[0] >> l, 1), h = max(t.h[0] >> l, 1), d = max(t.d[0] >> l, 1);
    const float half = mag ? 0.5f : 0.0f;
    float x = u * float(w) - half, y = v * float(h) - half, z = w_ * float(d) - half;
    int i0, j0, k0;
    float fx = floor_small(x, i0), fy = floor_small(y, j0), fz = floor_small(z, k0);
    TapB tap;
    tap.a = mag ? x - fx : 0.0f;
    tap.b = mag ? y - fy : 0.0f;
    tap.c = mag ? z - fz : 0.0f;
    // LINEAR: base texel -1 .. size-1 touches the texture; NEAREST: texel 0 .. size-1
    const float lo = mag ? -1.0f : 0.0f;
    bool inside = fx >= lo && fx <= float(w - 1) && fy >= lo && fy <= float(h - 1) && fz >= lo && fz <= float(d - 1);
    tap.cell = inside ? (long long)t.cell_off[l] + ((long long)(k0 + t.pad_z) * t.cell_h[l] + (j0 + t.pad_xy)) * t.cell_w[l] + (i0 + t.pad_xy) : -1ll;
    return tap;
}
SKY_D uint2 load_cell8(const MipView& t, unsigned int cell) { return __ldg(reinterpret_cast<const uint2*>(t.cells) + cell); }
SKY_D uint4 load_cell16(const MipView& t, unsigned int cell) { return __ldg(reinterpret_cast<const uint4*>(t.cells) + cell); }
SKY_D uint2 load_cellb(const MipView& t, const TapB& tap) {  // always a valid address: cell 0 stands in for the border
    return __ldg(reinterpret_cast<const uint2*>(t.cells) + (tap.cell < 0 ? 0ll : tap.cell));
}

// bilinear blend of channel c of an RG8 cell (8 bytes: {c00 c10 | c01 c11} x {r,g}) / an RGBA8 cell (16 bytes: one word per corner)
SKY_D float blend_rg8(uint2 cell, int c, float a, float b) {
    float r0 = lerp_biased(byte_biased(cell.x, c), byte_biased(cell.x, 2 + c), a);
    float r1 = lerp_biased(byte_biased(cell.y, c), byte_biased(cell.y, 2 + c), a);
    return (r0 + b * (r1 - r0)) * (1.0f / 255.0f);
}
SKY_D float blend_rgba8(uint4 cell, int c, float a, float b) {
    float r0 = lerp_biased(byte_biased(cell.x, c), byte_biased(cell.y, c), a);
    float r1 = lerp_biased(byte_biased(cell.z, c), byte_biased(cell.w, c), a);
    return (r0 + b * (r1 - r0)) * (1.0f / 255.0f);
}
// trilinear blend of the 8 corner bytes of a packed cell (byte n = di + 2*dj + 4*dk), x first
SKY_D float blend_cell(uint2 cell, float a, float b, float c) {
    float x00 = lerp_biased(byte_biased(cell.x, 0), byte_biased(cell.x, 1), a), x10 = lerp_biased(byte_biased(cell.x, 2), byte_biased(cell.x, 3), a);
    float x01 = lerp_biased(byte_biased(cell.y, 0), byte_biased(cell.y, 1), a), x11 = lerp_biased(byte_biased(cell.y, 2), byte_biased(cell.y, 3), a);
    float y0 = x00 + b * (x10 - x00), y1 = x01 + b * (x11 - x01);
    return (y0 + c * (y1 - y0)) * (1.0f / 255.0f);
}

SKY_D float blend_cell_raw(uint2 cell, float a, float b, float c) {  // the same blend on the byte scale (0..255): the caller folds 1/255 into its own factor
    float x00 = lerp_biased(byte_biased(cell.x, 0), byte_biased(cell.x, 1), a), x10 = lerp_biased(byte_biased(cell.x, 2), byte_biased(cell.x, 3), a);
    float x01 = lerp_biased(byte_biased(cell.y, 0), byte_biased(cell.y, 1), a), x11 = lerp_biased(byte_biased(cell.y, 2), byte_biased(cell.y, 3), a);
    float y0 = x00 + b * (x10 - x00), y1 = x01 + b * (x11 - x01);
    return y0 + c * (y1 - y0);
}

// K19's fast path: level-0 LINEAR lookup of the voxel grid for a position the caller has already bounded to the
// footprint apron (u, v at most a quarter texel outside [-1/2w, 1 + 1/2w], w_ in [0, 1]): the padded cell layout needs
// no range test.  Same arithmetic as tapb_border.
struct VoxelTap { unsigned int xy; int k; float a, b, c; };
SKY_D VoxelTap voxel_tap(const MipView& t, float u, float v, float w_) {
    float x = u * float(t.w[0]) - 0.5f, y = v * float(t.h[0]) - 0.5f, z = w_ * float(t.d[0]) - 0.5f;
    int i0, j0, k0;
    float fx = floor_small(x, i0), fy = floor_small(y, j0), fz = floor_small(z, k0);
    VoxelTap tap;
    tap.a = x - fx; tap.b = y - fy; tap.c = z - fz;
    tap.xy = (unsigned int)((j0 + 2) * t.cell_w[0] + (i0 + 2));
    tap.k = k0 + 1;
    return tap;
}
SKY_D VoxelTap voxel_tap_texel(const MipView& t, float x, float y, float z) {  // the same tap from level-0 texel coordinates (u * w - 0.5, ...)
    int i0, j0, k0;
    float fx = floor_small(x, i0), fy = floor_small(y, j0), fz = floor_small(z, k0);
    VoxelTap tap;
    tap.a = x - fx; tap.b = y - fy; tap.c = z - fz;
    tap.xy = (unsigned int)((j0 + 2) * t.cell_w[0] + (i0 + 2));
    tap.k = k0 + 1;
    return tap;
}
SKY_D uint2 voxel_tap_load(const MipView& t, const VoxelTap& tap, bool live) {  // !live: any valid address
    size_t cell = live ? size_t((unsigned int)tap.k) * (unsigned int)(t.cell_w[0] * t.cell_h[0]) + tap.xy : size_t(0);
    return __ldg(reinterpret_cast<const uint2*>(t.cells) + cell);
}

// ---- hardware path ----------------------------------------------------------------------------------
// LINEAR on level 0 and POINT over the mip chain are two texture objects (a CUDA texture object has one filter mode)
// (one TEX with the handle selected per lane measured 20 % slower than the branch: 534 vs 445 us for K16 at 4K)
SKY_D float4 sample2d_hw(const MipView& t, float u, float v, int level) {
    return level < 0 ? tex2DLod<float4>(t.tex_linear, u, v, 0.0f) : tex2DLod<float4>(t.tex_point, u, v, float(level));
}
SKY_D float sample3d_hw(const MipView& t, float u, float v, float w, int level) {
    return level < 0 ? tex3DLod<float>(t.tex_linear, u, v, w, 0.0f) : tex3DLod<float>(t.tex_point, u, v, w, float(level));
}
SKY_D float4 sample2d_hw(const MipView& t, float u, float v, HwLod l) {
    return l.mag ? tex2DLod<float4>(t.tex_linear, u, v, 0.0f) : tex2DLod<float4>(t.tex_point, u, v, l.lod);
}
SKY_D float sample3d_hw(const MipView& t, float u, float v, float w, HwLod l) {
    return l.mag ? tex3DLod<float>(t.tex_linear, u, v, w, 0.0f) : tex3DLod<float>(t.tex_point, u, v, w, l.lod);
}
// Production objects, DEFAULT0: fetch the two displacement taps only for evaluations that pass the weather-map early-out (one more
// dependent texture round trip, ~35 fewer instructions for every evaluation in clear air).  The kernels that run this material are
// bound by instruction issue as much as by latency; measured, the eager fetch (0) wins.
#ifndef SKY_M0_LATE_DISPLACEMENT
#define SKY_M0_LATE_DISPLACEMENT 0   // measured at 4K, scene c3 (profiles/k16_variants_r02n.log): late 383 us, eager 367 us -- the round trip costs more than the instructions
#endif
constexpr bool kLateDisplacement = SKY_M0_LATE_DISPLACEMENT != 0;

// ---- SampleSigmaT ------------------------------------------------------------------------------------
// VolumetricCloudDefaultMaterial0.glsl:9-16
SKY_D float CalHeightMask(float cloud_type, float height01) {
    float height_in_type = clampf(SKY_FDIV(height01, cloud_type), 0.0f, 1.0f);  // same inf / NaN cases as '/'
    return clampf(height_in_type * (height_in_type - 1.0f) * -4.0f, 0.0f, 1.0f);
}
SKY_D float Remap01(float x, float x0, float x1) { return clampf(SKY_FDIV(x - x0, x1 - x0), 0.0f, 1.0f); }
SKY_D float distance2(float3 a, float3 b) { float3 d = a - b; return dot(d, d); }

// One SampleSigmaT(pos, height01) evaluation split at its texture fetches, so a caller can keep several evaluations
// in flight: issue() starts the position-only fetches, advance() consumes them and starts the dependent fetch,
// finish() returns sigma_t.  SampleSigmaT() below is the three in a row.  `fetches` counts texture
// fetches as the reference issues them.
//
// Exact early-out of DEFAULT0: sigma_t = Remap01(base, detail', 1) * height01 * density with base = cloud_map.r *
// CalHeightMask(cloud_map.g, height01).  Where the weather map leaves no cloud (base == 0) the result is
// clamp(-detail' / (1 - detail'), 0, 1) = 0 for every detail' in [0, 1], i.e. for every texel when uDetailParam keeps
// detail' in that range (checked on the host: MaterialParams::m0_zero_base_is_zero).  Such evaluations -- most of the
// clear-air steps of a march -- end before the displacement blend and the detail (3-D) fetch.  DEFAULT1 has the same
// early-out in the shader itself (:22).
template <int MAT, bool HW>
struct SigmaEval {
    float3 pos;
    float height01;
    bool need;  // the dependent fetches are needed
    // DEFAULT0 / DEFAULT1
    Tap2 t_cm, t_d0, t_d1;
    Tap3 t_dt;
    uint2 c_cm, c_dt;
    uint4 c_d0, c_d1;
    float cloud_type[2], disp[4], detail, log2_d2, density;
    // VOXEL
    TapB t_vx;
    uint2 c_vx;

    SKY_D void issue(const MaterialParams& M, float3 p, float h01) {
        pos = p; height01 = h01; need = true;
        if (MAT == SKY_MATERIAL_DEFAULT0 || MAT == SKY_MATERIAL_DEFAULT1) {
            const SkyMaterialCommonBufferData& mc = M.m.common;
            log2_d2 = SKY_LOG2(distance2(pos, M.camera_pos));
            const SkySampleInfo& ci = mc.uCloudMapSampleInfo;
            float cu = pos.x * ci.frequency + ci.bias[0], cv = pos.y * ci.frequency + ci.bias[1];
            if (HW) {
                float4 c = sample2d_hw(M.cloud_map, cu, cv, hw_lod_from_log2(log2_d2, M.lod_h_cloud_map));
                cloud_type[0] = c.x; cloud_type[1] = c.y;
            } else {
                int lc = level_from_log2(log2_d2, M.lod_h_cloud_map, M.cloud_map.levels);
                t_cm = tap2_repeat(M.cloud_map, cu, cv, lc);
                c_cm = load_cell8(M.cloud_map, t_cm.cell);
            }
            if (MAT == SKY_MATERIAL_DEFAULT0 && !(HW && kLateDisplacement)) {
                // the displacement fetches do not depend on the weather map: issued with it (one latency, not two),
                // even though an evaluation that ends at the early-out below drops them
                const SkySampleInfo& di = mc.uDisplacementSampleInfo;
                float du = pos.x * di.frequency + di.bias[0], dv = pos.y * di.frequency + di.bias[1], dw = pos.z * di.frequency;
                if (HW) {
                    const HwLod ld = hw_lod_from_log2(log2_d2, M.lod_h_displacement);
                    float4 a = sample2d_hw(M.displacement, du, dv, ld);
                    float4 b = sample2d_hw(M.displacement, du, dw, ld);
                    disp[0] = a.x; disp[1] = a.y; disp[2] = b.z; disp[3] = b.w;
                } else {
                    int ld = level_from_log2(log2_d2, M.lod_h_displacement, M.displacement.levels);
                    t_d0 = tap2_repeat(M.displacement, du, dv, ld);
                    t_d1 = tap2_repeat(M.displacement, du, dw, ld);
                    c_d0 = load_cell16(M.displacement, t_d0.cell);
                    c_d1 = load_cell16(M.displacement, t_d1.cell);
                }
            }
        } else if (MAT == SKY_MATERIAL_VOXEL) {
            const SkyMaterialVoxelBufferData& m = M.m.u.voxel;
            float u = pos.x * m.uSampleFrequency[0] + m.uSampleBias[0];
            float v = pos.y * m.uSampleFrequency[1] + m.uSampleBias[1];
            if (HW) {
                detail = sample3d_hw(M.voxel, u, v, height01, hw_lod_from_log2(SKY_LOG2(distance2(pos, M.camera_pos)), M.lod_h_voxel));
            } else {
                int level = level_from_log2(SKY_LOG2(distance2(pos, M.camera_pos)), M.lod_h_voxel, M.voxel.levels);
                t_vx = tapb_border(M.voxel, u, v, height01, level);
                c_vx = load_cellb(M.voxel, t_vx);
            }
        }
    }

    SKY_D void advance(const MaterialParams& M) {
        if (MAT == SKY_MATERIAL_DEFAULT0) {  // VolumetricCloudDefaultMaterial0.glsl:18-32
            const SkyMaterialCommonBufferData& mc = M.m.common;
            const SkyMaterial0BufferData& m = M.m.u.m0;
            if (!HW) { cloud_type[0] = blend_rg8(c_cm, 0, t_cm.a, t_cm.b); cloud_type[1] = blend_rg8(c_cm, 1, t_cm.a, t_cm.b); }
            density = cloud_type[0] * CalHeightMask(cloud_type[1], height01);  // `base`
            need = !(M.m0_zero_base_is_zero && density == 0.0f);
#ifdef SKY_EXPERIMENT_FORCE_SKIP  // timing experiment only (wrong images): what an evaluation costs when it ends at the early-out
            need = false;
#endif
            if (need) {
                if (HW && kLateDisplacement) {
                    const SkySampleInfo& di = mc.uDisplacementSampleInfo;
                    const float du = pos.x * di.frequency + di.bias[0], dv = pos.y * di.frequency + di.bias[1], dw = pos.z * di.frequency;
                    const HwLod ld = hw_lod_from_log2(log2_d2, M.lod_h_displacement);
                    const float4 a = sample2d_hw(M.displacement, du, dv, ld), b = sample2d_hw(M.displacement, du, dw, ld);
                    disp[0] = a.x; disp[1] = a.y; disp[2] = b.z; disp[3] = b.w;
                }
                if (!HW) {
                    disp[0] = blend_rgba8(c_d0, 0, t_d0.a, t_d0.b); disp[1] = blend_rgba8(c_d0, 1, t_d0.a, t_d0.b);
                    disp[2] = blend_rgba8(c_d1, 2, t_d1.a, t_d1.b); disp[3] = blend_rgba8(c_d1, 3, t_d1.a, t_d1.b);
                }
#ifdef SKY_STRICT_TU   // the shader's literal sum (:24-26), signed zeros included
                float3 displace_vector = f3(0.0f + disp[0] + disp[2], 0.0f + disp[1], 0.0f + disp[3]);
#else
                float3 displace_vector = f3(disp[0] + disp[2], disp[1], disp[3]);
#endif
                float3 p = pos + m.uDisplacementScale * displace_vector;
                const SkySampleInfo& ti = mc.uDetailSampleInfo;
                float tu = p.x * ti.frequency + ti.bias[0], tv = p.y * ti.frequency + ti.bias[1], tw = p.z * ti.frequency;
                if (HW) {
                    detail = sample3d_hw(M.detail, tu, tv, tw, hw_lod_from_log2(SKY_LOG2(distance2(p, M.camera_pos)), M.lod_h_detail));
                } else {
                    int lt = level_from_log2(SKY_LOG2(distance2(p, M.camera_pos)), M.lod_h_detail, M.detail.levels);
                    t_dt = tap3_repeat(M.detail, tu, tv, tw, lt);
                    c_dt = load_cell8(M.detail, t_dt.cell);
                }
            }
        } else if (MAT == SKY_MATERIAL_DEFAULT1) {  // VolumetricCloudDefaultMaterial1.glsl:14-29
            const SkyMaterialCommonBufferData& mc = M.m.common;
            const SkyMaterial1BufferData& m = M.m.u.m1;
            if (!HW) { cloud_type[0] = blend_rg8(c_cm, 0, t_cm.a, t_cm.b); cloud_type[1] = blend_rg8(c_cm, 1, t_cm.a, t_cm.b); }
            density = clampf((cloud_type[0] - m.uBaseDensityThreshold) * m.uBaseEdgeHardness, 0.0f, 1.0f);
            density *= clampf((1 - height01) * m.uBaseHeightHardness, 0.0f, 1.0f);
            need = !(density == 0);  // :22 returns before the detail fetch
            if (need) {
                const SkySampleInfo& ti = mc.uDetailSampleInfo;
                float tu = pos.x * ti.frequency + ti.bias[0], tv = pos.y * ti.frequency + ti.bias[1], tw = pos.z * ti.frequency;
                if (HW) {
                    detail = sample3d_hw(M.detail, tu, tv, tw, hw_lod_from_log2(log2_d2, M.lod_h_detail));
                } else {
                    int lt = level_from_log2(log2_d2, M.lod_h_detail, M.detail.levels);
                    t_dt = tap3_repeat(M.detail, tu, tv, tw, lt);
                    c_dt = load_cell8(M.detail, t_dt.cell);
                }
            }
        }
    }

    SKY_D float finish(const MaterialParams& M, int* fetches = nullptr) {
        if (MAT == SKY_MATERIAL_DEFAULT0) {
            const SkyMaterialCommonBufferData& mc = M.m.common;
            const SkyMaterial0BufferData& m = M.m.u.m0;
            if (fetches) *fetches += 4;
            if (!need) return 0.0f;
            if (!HW) detail = blend_cell(c_dt, t_dt.a, t_dt.b, t_dt.c);
            float dd = detail * m.uDetailParam[0] + m.uDetailParam[1];
            return Remap01(density, dd, 1.0f) * height01 * mc.uDensity;
        } else if (MAT == SKY_MATERIAL_DEFAULT1) {
            const SkyMaterialCommonBufferData& mc = M.m.common;
            const SkyMaterial1BufferData& m = M.m.u.m1;
            if (fetches) *fetches += need ? 2 : 1;
            if (!need) return 0.0f;
            if (!HW) detail = blend_cell(c_dt, t_dt.a, t_dt.b, t_dt.c);
            float dd = (detail + m.uDetailBase) * m.uDetailScale;
            dd *= fmaxf(clampf(height01 - m.uHeightCut, 0.0f, 1.0f), clampf(m.uEdgeCur - cloud_type[0], 0.0f, 1.0f));
            return clampf(density - dd, 0.0f, 1.0f) * mc.uDensity * height01;
        } else if (MAT == SKY_MATERIAL_MINIMAL) {  // VolumetricCloudMaterialMinimal.glsl:6-8
            return M.m.u.minimal.uDensity;
        } else {  // VolumetricCloudMaterialVoxel.glsl:12-17
            if (fetches) *fetches += 1;
            if (!HW) detail = t_vx.cell < 0 ? 0.0f : blend_cell(c_vx, t_vx.a, t_vx.b, t_vx.c);
            return detail * M.m.u.voxel.uDensity;
        }
    }
};

// MAT: SkyMaterialType.  `fetches` (optional) receives the number of texture fetches the reference issues.
template <int MAT, bool HW>
SKY_D float SampleSigmaT(const MaterialParams& M, float3 pos, float height01, int* fetches = nullptr) {
    SigmaEval<MAT, HW> e;
    e.issue(M, pos, height01);
    e.advance(M);
    return e.finish(M, fetches);
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
