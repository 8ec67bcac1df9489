// Cloud density materials: float SampleSigmaT(vec3 pos, float height01), the inner cost of K11, K16
// and K19.  Follows VolumetricCloudDefaultMaterial{Common,0,1}.glsl and
// VolumetricCloudMaterial{Minimal,Voxel}.glsl.
//
// Sampler semantics (VolumetricCloudDefaultMaterial.cpp:111-116, VolumetricCloudVoxelMaterial.cpp:30-37):
// mag = LINEAR, min = NEAREST_MIPMAP_NEAREST, explicit LOD.  GL 4.6 section 8.14.3: lambda <= 0.5 ->
// LINEAR on level 0; otherwise NEAREST on level ceil(lambda + 0.5) - 1.
//   HW == false: exact fp32 weights on texels read straight from linear device memory (L1/L2 hits);
//                bit-compatible with the oracle's software sampler.
//   HW == true : the texture unit filters (8-bit weights); one TEX per fetch instead of 4-8 loads.
#pragma once
#include "common.cuh"
#include "context.h"

struct MaterialParams {
    SkyMaterialBlock m;
    MipView cloud_map, detail, displacement, voxel;
    float3 camera_pos;  // uCameraPos
};

// spec 8.14.3 level selection; returns -1 for magnification
SKY_D int select_mip_level(float lod, int levels) {
    if (!(lod > 0.5f)) return -1;
    int q = levels - 1;
    int d = (lod <= float(q) + 0.5f) ? int(ceilf(lod + 0.5f)) - 1 : q;
    return clampi(d, 0, q);
}

// ---- exact path: REPEAT textures have power-of-two sizes (512 / 128, fixed by the reference) --------
template <int C>
SKY_D void load_texel(const MipView& t, int level, int x, int y, int z, float* out) {
    const uint8_t* p = t.base + (t.off[level] + (size_t(z) * t.h[level] + y) * t.w[level] + x) * C;
    if (C == 1) {
        out[0] = float(__ldg(p)) * (1.0f / 255.0f);
    } else if (C == 2) {
        uchar2 v = __ldg(reinterpret_cast<const uchar2*>(p));
        out[0] = float(v.x) * (1.0f / 255.0f); out[1] = float(v.y) * (1.0f / 255.0f);
    } else {
        uchar4 v = __ldg(reinterpret_cast<const uchar4*>(p));
        out[0] = float(v.x) * (1.0f / 255.0f); out[1] = float(v.y) * (1.0f / 255.0f);
        out[2] = float(v.z) * (1.0f / 255.0f); out[3] = float(v.w) * (1.0f / 255.0f);
    }
}

SKY_D float byte_of(uint32_t word, int n) { return float((word >> (8 * n)) & 0xffu) * (1.0f / 255.0f); }

template <int C>
SKY_D void sample2d_repeat_exact(const MipView& t, float u, float v, float lod, float* out) {
    int level = select_mip_level(lod, t.levels);
    if (level < 0) {
        int w = t.w[0], h = t.h[0];
        float x = u * float(w) - 0.5f, y = v * float(h) - 0.5f;
        float fx = floorf(x), fy = floorf(y);
        float a = x - fx, b = y - fy;
        int i0 = int(fx) & (w - 1), j0 = int(fy) & (h - 1);
        float w00 = (1.0f - a) * (1.0f - b), w10 = a * (1.0f - b), w01 = (1.0f - a) * b, w11 = a * b;
        // one load brings the four corners: C == 2 -> 8 bytes {c00 c10 c01 c11} x {r,g}; C == 4 -> 16 bytes
        if (C == 2) {
            uint2 cell = __ldg(reinterpret_cast<const uint2*>(t.cells) + size_t(j0) * w + i0);
#pragma unroll
            for (int c = 0; c < 2; ++c)
                out[c] = w00 * byte_of(cell.x, c) + w10 * byte_of(cell.x, 2 + c) + w01 * byte_of(cell.y, c) + w11 * byte_of(cell.y, 2 + c);
        } else {
            uint4 cell = __ldg(reinterpret_cast<const uint4*>(t.cells) + size_t(j0) * w + i0);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                out[c] = w00 * byte_of(cell.x, c) + w10 * byte_of(cell.y, c) + w01 * byte_of(cell.z, c) + w11 * byte_of(cell.w, c);
        }
    } else {
        int w = t.w[level], h = t.h[level];
        int i = int(floorf(u * float(w))) & (w - 1), j = int(floorf(v * float(h))) & (h - 1);
        load_texel<C>(t, level, i, j, 0, out);
    }
}

SKY_D float sample3d_repeat_exact(const MipView& t, float u, float v, float w_, float lod) {
    int level = select_mip_level(lod, t.levels);
    float out;
    if (level < 0) {
        int w = t.w[0], h = t.h[0], d = t.d[0];
        float x = u * float(w) - 0.5f, y = v * float(h) - 0.5f, z = w_ * float(d) - 0.5f;
        float fx = floorf(x), fy = floorf(y), fz = floorf(z);
        float a = x - fx, b = y - fy, c = z - fz;
        int i0 = int(fx) & (w - 1), j0 = int(fy) & (h - 1), k0 = int(fz) & (d - 1);
        // 8 corners in one 8-byte load: byte n = di + 2*dj + 4*dk
        uint2 cell = __ldg(reinterpret_cast<const uint2*>(t.cells) + (size_t(k0) * h + j0) * w + i0);
        float r = 0.0f;
#pragma unroll
        for (int dk = 0; dk < 2; ++dk)
#pragma unroll
            for (int dj = 0; dj < 2; ++dj)
#pragma unroll
                for (int di = 0; di < 2; ++di) {
                    float wt = (di ? a : 1.0f - a) * (dj ? b : 1.0f - b) * (dk ? c : 1.0f - c);
                    r += wt * byte_of(dk ? cell.y : cell.x, di + 2 * dj);
                }
        out = r;
    } else {
        int w = t.w[level], h = t.h[level], d = t.d[level];
        int i = int(floorf(u * float(w))) & (w - 1), j = int(floorf(v * float(h))) & (h - 1), k = int(floorf(w_ * float(d))) & (d - 1);
        load_texel<1>(t, level, i, j, k, &out);
    }
    return out;
}

// CLAMP_TO_BORDER with border colour 0, any size (voxel grid)
SKY_D float sample3d_border_exact(const MipView& t, float u, float v, float w_, float lod) {
    int level = select_mip_level(lod, t.levels);
    if (level < 0) {
        int w = t.w[0], h = t.h[0], d = t.d[0];
        float x = u * float(w) - 0.5f, y = v * float(h) - 0.5f, z = w_ * float(d) - 0.5f;
        float fx = floorf(x), fy = floorf(y), fz = floorf(z);
        float a = x - fx, b = y - fy, c = z - fz;
        // cell index = base texel + 1; outside [0, w] x [0, h] x [0, d] every corner is the border
        float cx = fx + 1.0f, cy = fy + 1.0f, cz = fz + 1.0f;
        if (!(cx >= 0.0f && cx <= float(w) && cy >= 0.0f && cy <= float(h) && cz >= 0.0f && cz <= float(d))) return 0.0f;
        uint2 cell = __ldg(reinterpret_cast<const uint2*>(t.cells) + (size_t(int(cz)) * t.cell_h + int(cy)) * t.cell_w + int(cx));
        float r = 0.0f;
#pragma unroll
        for (int dk = 0; dk < 2; ++dk)
#pragma unroll
            for (int dj = 0; dj < 2; ++dj)
#pragma unroll
                for (int di = 0; di < 2; ++di) {
                    float wt = (di ? a : 1.0f - a) * (dj ? b : 1.0f - b) * (dk ? c : 1.0f - c);
                    r += wt * byte_of(dk ? cell.y : cell.x, di + 2 * dj);
                }
        return r;
    }
    int w = t.w[level], h = t.h[level], d = t.d[level];
    int i = int(floorf(u * float(w))), j = int(floorf(v * float(h))), k = int(floorf(w_ * float(d)));
    float out = 0.0f;
    if (i >= 0 && i < w && j >= 0 && j < h && k >= 0 && k < d) load_texel<1>(t, level, i, j, k, &out);
    return out;
}

// ---- hardware path ----------------------------------------------------------------------------------
SKY_D float4 sample2d_hw(const MipView& t, float u, float v, float lod) {
    int level = select_mip_level(lod, t.levels);
    return level < 0 ? tex2DLod<float4>(t.tex_linear, u, v, 0.0f) : tex2DLod<float4>(t.tex_point, u, v, float(level));
}
SKY_D float sample3d_hw(const MipView& t, float u, float v, float w, float lod) {
    int level = select_mip_level(lod, t.levels);
    return level < 0 ? tex3DLod<float>(t.tex_linear, u, v, w, 0.0f) : tex3DLod<float>(t.tex_point, u, v, w, float(level));
}

// ---- SampleSigmaT ------------------------------------------------------------------------------------
// VolumetricCloudDefaultMaterialCommon.glsl:20-24
SKY_D float4 GetUVWLod(float3 pos, const SkySampleInfo& info, float3 camera_pos, float lod_bias) {
    float lod = log2f(info.k_lod * distance(pos, camera_pos)) + lod_bias;
    return f4(pos.x * info.frequency + info.bias[0], pos.y * info.frequency + info.bias[1], pos.z * info.frequency, lod);
}
// VolumetricCloudDefaultMaterial0.glsl:9-16
SKY_D float CalHeightMask(float cloud_type, float height01) {
    float height_in_type = clampf(height01 / cloud_type, 0.0f, 1.0f);
    return clampf(height_in_type * (height_in_type - 1.0f) * -4.0f, 0.0f, 1.0f);
}
SKY_D float Remap01(float x, float x0, float x1) { return clampf((x - x0) / (x1 - x0), 0.0f, 1.0f); }

// MAT: SkyMaterialType.  `fetches` (optional) receives the number of texture fetches issued.
template <int MAT, bool HW>
SKY_D float SampleSigmaT(const MaterialParams& M, float3 pos, float height01, int* fetches = nullptr) {
    if (MAT == SKY_MATERIAL_DEFAULT0) {  // VolumetricCloudDefaultMaterial0.glsl:18-32
        const SkyMaterialCommonBufferData& mc = M.m.common;
        const SkyMaterial0BufferData& m = M.m.u.m0;
        float4 uvwlod = GetUVWLod(pos, mc.uCloudMapSampleInfo, M.camera_pos, mc.uLodBias);
        float cloud_type[2];
        float d0[4], d1[4];
        if (HW) {
            float4 c = sample2d_hw(M.cloud_map, uvwlod.x, uvwlod.y, uvwlod.w);
            cloud_type[0] = c.x; cloud_type[1] = c.y;
        } else {
            sample2d_repeat_exact<2>(M.cloud_map, uvwlod.x, uvwlod.y, uvwlod.w, cloud_type);
        }
        uvwlod = GetUVWLod(pos, mc.uDisplacementSampleInfo, M.camera_pos, mc.uLodBias);
        if (HW) {
            float4 a = sample2d_hw(M.displacement, uvwlod.x, uvwlod.y, uvwlod.w);
            float4 b = sample2d_hw(M.displacement, uvwlod.x, uvwlod.z, uvwlod.w);
            d0[0] = a.x; d0[1] = a.y; d1[2] = b.z; d1[3] = b.w;
        } else {
            sample2d_repeat_exact<4>(M.displacement, uvwlod.x, uvwlod.y, uvwlod.w, d0);
            sample2d_repeat_exact<4>(M.displacement, uvwlod.x, uvwlod.z, uvwlod.w, d1);
        }
        float3 displace_vector = f3(0.0f + d0[0] + d1[2], 0.0f + d0[1], 0.0f + d1[3]);
        pos = pos + m.uDisplacementScale * displace_vector;
        uvwlod = GetUVWLod(pos, mc.uDetailSampleInfo, M.camera_pos, mc.uLodBias);
        float detail = HW ? sample3d_hw(M.detail, uvwlod.x, uvwlod.y, uvwlod.z, uvwlod.w)
                          : sample3d_repeat_exact(M.detail, uvwlod.x, uvwlod.y, uvwlod.z, uvwlod.w);
        detail = detail * m.uDetailParam[0] + m.uDetailParam[1];
        if (fetches) *fetches += 4;
        return Remap01(cloud_type[0] * CalHeightMask(cloud_type[1], height01), detail, 1.0f) * height01 * mc.uDensity;
    } else if (MAT == SKY_MATERIAL_DEFAULT1) {  // VolumetricCloudDefaultMaterial1.glsl:14-29
        const SkyMaterialCommonBufferData& mc = M.m.common;
        const SkyMaterial1BufferData& m = M.m.u.m1;
        float4 uvwlod = GetUVWLod(pos, mc.uCloudMapSampleInfo, M.camera_pos, mc.uLodBias);
        float cloud_type[2];
        if (HW) {
            float4 c = sample2d_hw(M.cloud_map, uvwlod.x, uvwlod.y, uvwlod.w);
            cloud_type[0] = c.x; cloud_type[1] = c.y;
        } else {
            sample2d_repeat_exact<2>(M.cloud_map, uvwlod.x, uvwlod.y, uvwlod.w, cloud_type);
        }
        if (fetches) *fetches += 1;
        float density = clampf((cloud_type[0] - m.uBaseDensityThreshold) * m.uBaseEdgeHardness, 0.0f, 1.0f);
        density *= clampf((1 - height01) * m.uBaseHeightHardness, 0.0f, 1.0f);
        if (density == 0) return 0.0f;
        uvwlod = GetUVWLod(pos, mc.uDetailSampleInfo, M.camera_pos, mc.uLodBias);
        float detail = HW ? sample3d_hw(M.detail, uvwlod.x, uvwlod.y, uvwlod.z, uvwlod.w)
                          : sample3d_repeat_exact(M.detail, uvwlod.x, uvwlod.y, uvwlod.z, uvwlod.w);
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
        float lod = log2f(m.uSampleLodK * distance(pos, M.camera_pos)) + m.uLodBias;
        float density = HW ? sample3d_hw(M.voxel, u, v, height01, lod) : sample3d_border_exact(M.voxel, u, v, height01, lod);
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
