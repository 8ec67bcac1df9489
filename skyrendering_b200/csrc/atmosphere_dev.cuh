// Device functions of the atmosphere model, shared by the LUT kernels (atmosphere.cu), the cloud
// chain (cloud.cu) and the path tracer (pathtrace.cu).  They follow
// shaders/SkyRendering/Atmosphere.glsl and AtmosphereInterface.glsl; LUT reads are exact fp32
// software bilinear/trilinear fetches (the LUTs are RGBA32F and L1/L2 resident, so the 8-bit
// weights of the hardware filter are not worth their error here).
#pragma once
#include "common.cuh"

struct LutView {
    const float4* p;
    int w, h, d;
    cudaTextureObject_t tex;  // optional (2-D LUTs): LINEAR + CLAMP_TO_EDGE over the same memory, see sample_lut2d_hw
};

// GL_LINEAR + CLAMP_TO_EDGE on an RGBA32F image: u*w - 0.5, floor, fract (GL 4.6 section 8.14.2)
SKY_D float4 sample_lut2d(const LutView& t, float u, float v) {
    float x = u * float(t.w) - 0.5f, y = v * float(t.h) - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    int i0 = int(fx), j0 = int(fy);
    int i1 = clampi(i0 + 1, 0, t.w - 1), j1 = clampi(j0 + 1, 0, t.h - 1);
    i0 = clampi(i0, 0, t.w - 1); j0 = clampi(j0, 0, t.h - 1);
    float4 t00 = __ldg(t.p + j0 * t.w + i0), t10 = __ldg(t.p + j0 * t.w + i1);
    float4 t01 = __ldg(t.p + j1 * t.w + i0), t11 = __ldg(t.p + j1 * t.w + i1);
    return (1.0f - a) * (1.0f - b) * t00 + a * (1.0f - b) * t10 + (1.0f - a) * b * t01 + a * b * t11;
}

// The same fetch through the texture unit (8-bit interpolation weights: an error of at most 1/512 of the difference
// between neighbouring texels, ~1e-4 relative on these smooth LUTs).  Only the per-pixel raymarch of the composite
// K6 -- a frame, tolerance 1e-2 relative RMS, 40 steps x 2 LUT fetches per ground pixel -- uses it.
template <bool TEXLUT>
SKY_D float4 sample_lut2d_sel(const LutView& t, float u, float v) {
    if (TEXLUT) return tex2D<float4>(t.tex, u, v);
    return sample_lut2d(t, u, v);
}

// 3-D LUT through the texture unit: the slices of the LUT are stacked in one 2-D texture (LutView::tex over the same memory);
// two bilinear fetches (8-bit weights) + the slice blend in fp32.  v is clamped to the texel centres of its slice, which is
// what CLAMP_TO_EDGE does and keeps the bilinear footprint out of the neighbouring slice.
template <bool TEXLUT>
SKY_D float4 sample_lut3d_sel(const LutView& t, float u, float v, float w);
SKY_D float4 sample_lut3d_atlas(const LutView& t, float u, float v, float w) {
    float z = w * float(t.d) - 0.5f;
    float fz = floorf(z), c = z - fz;
    int k0 = clampi(int(fz), 0, t.d - 1), k1 = clampi(int(fz) + 1, 0, t.d - 1);
    float hv = 0.5f / float(t.h);
    float vc = fminf(fmaxf(v, hv), 1.0f - hv);
    float inv_d = 1.0f / float(t.d);
    float4 a = tex2D<float4>(t.tex, u, (float(k0) + vc) * inv_d), b = tex2D<float4>(t.tex, u, (float(k1) + vc) * inv_d);
    return f4(a.x + c * (b.x - a.x), a.y + c * (b.y - a.y), a.z + c * (b.z - a.z), a.w + c * (b.w - a.w));
}
SKY_D float4 sample_lut3d(const LutView& t, float u, float v, float w) {
    float x = u * float(t.w) - 0.5f, y = v * float(t.h) - 0.5f, z = w * float(t.d) - 0.5f;
    float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    float a = x - fx, b = y - fy, c = z - fz;
    int i[2], j[2], k[2];
    i[0] = clampi(int(fx), 0, t.w - 1); i[1] = clampi(int(fx) + 1, 0, t.w - 1);
    j[0] = clampi(int(fy), 0, t.h - 1); j[1] = clampi(int(fy) + 1, 0, t.h - 1);
    k[0] = clampi(int(fz), 0, t.d - 1); k[1] = clampi(int(fz) + 1, 0, t.d - 1);
    float4 r = f4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int dk = 0; dk < 2; ++dk)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int di = 0; di < 2; ++di) {
                float wt = (di ? a : 1.0f - a) * (dj ? b : 1.0f - b) * (dk ? c : 1.0f - c);
                r += wt * __ldg(t.p + (size_t(k[dk]) * t.h + j[dj]) * t.w + i[di]);
            }
    return r;
}

struct AtmosphereModel {
    SkyAtmosphereBufferData u;  // Atmosphere.glsl:6-32

    SKY_D float3 solar_illuminance() const { return f3(u.solar_illuminance); }
    SKY_D float3 ground_albedo() const { return f3(u.ground_albedo); }

    // Atmosphere.glsl:41-51
    SKY_D static float ClampDistance(float d) { return fmaxf(d, 0.0f); }
    SKY_D static float SafeSqrt(float a) { return sqrtf(fmaxf(a, 0.0f)); }

    // Atmosphere.glsl:57-59
    SKY_D bool RayIntersectsGround(float r, float mu) const {
        return mu < 0.0f && r * r * (mu * mu - 1.0f) + u.bottom_radius * u.bottom_radius >= 0.0f;
    }
    // Atmosphere.glsl:61-64
    SKY_D float DistanceToTopAtmosphereBoundary(float r, float mu) const {
        float discriminant = r * r * (mu * mu - 1.0f) + u.top_radius * u.top_radius;
        return ClampDistance(-r * mu + SafeSqrt(discriminant));
    }
    // Atmosphere.glsl:66-69
    SKY_D float DistanceToBottomAtmosphereBoundary(float r, float mu) const {
        float discriminant = r * r * (mu * mu - 1.0f) + u.bottom_radius * u.bottom_radius;
        return ClampDistance(-r * mu - SafeSqrt(discriminant));
    }
    // AtmosphereInterface.glsl:6-13
    SKY_D bool FromSpaceIntersectTopAtmosphereBoundary(float r, float mu, float& near_distance) const {
        float discriminant = r * r * (mu * mu - 1.0f) + u.top_radius * u.top_radius;
        if (mu < 0.0f && discriminant >= 0.0f) {
            near_distance = ClampDistance(-r * mu - SafeSqrt(discriminant));
            return true;
        }
        return false;
    }
    // Atmosphere.glsl:90-108; GetTextureCoordFromUnitRange :53-55
    template <bool TEXLUT = false>
    SKY_D float3 GetTransmittanceToTopAtmosphereBoundary(const LutView& tex, float r, float mu) const {
        float H = sqrtf(u.top_radius * u.top_radius - u.bottom_radius * u.bottom_radius);
        float rho = SafeSqrt(r * r - u.bottom_radius * u.bottom_radius);
        float d = DistanceToTopAtmosphereBoundary(r, mu);
        float d_min = u.top_radius - r;
        float d_max = rho + H;
        float x_mu = (d - d_min) / (d_max - d_min);
        float x_r = rho / H;
        float uu = 0.5f / float(tex.w) + x_mu * (1.0f - 1.0f / float(tex.w));
        float vv = 0.5f / float(tex.h) + x_r * (1.0f - 1.0f / float(tex.h));
        return xyz(sample_lut2d_sel<TEXLUT>(tex, uu, vv));
    }
    // Atmosphere.glsl:110-117
    template <bool TEXLUT = false>
    SKY_D float3 GetSunVisibility(const LutView& tex, float r, float mu_s) const {
        float sin_theta_h = u.bottom_radius / r;
        float cos_theta_h = -sqrtf(fmaxf(1.0f - sin_theta_h * sin_theta_h, 0.0f));
        return GetTransmittanceToTopAtmosphereBoundary<TEXLUT>(tex, r, mu_s) *
               smoothstepf(-sin_theta_h * u.sun_angular_radius, sin_theta_h * u.sun_angular_radius, mu_s - cos_theta_h);
    }
};

// AtmosphereInterface.glsl:15-23: uvw of the aerial-perspective froxel for (screen uv, distance)
SKY_D float3 aerial_perspective_uvw(float2 uv, float marching_distance, float max_distance, int w, int h, int d) {
    float z = sqrtf(marching_distance / max_distance);
    return f3(0.5f / float(w) + uv.x * (1.0f - 1.0f / float(w)), 0.5f / float(h) + uv.y * (1.0f - 1.0f / float(h)),
              0.5f / float(d) + z * (1.0f - 1.0f / float(d)));
}

// R16 unorm froxel volume, GL_LINEAR + CLAMP_TO_EDGE, exact weights
struct FroxelView {
    const uint16_t* p;
    int w, h, d;
    cudaTextureObject_t tex;   // optional (K6, production object): the slices stacked in one R16 2-D texture over the same memory, 0 = none
};
// The same fetch through the texture unit: two bilinear R16-unorm fetches (8-bit weights) + the slice blend in fp32, instead of eight loads,
// eight conversions and their weights.  v is clamped to the texel centres of its slice (CLAMP_TO_EDGE; keeps the footprint out of the next slice).
SKY_D float sample_froxel_atlas(const FroxelView& t, float u, float v, float w) {
    const float z = fminf(fmaxf(w * float(t.d) - 0.5f, -1.0f), float(t.d));   // dist >> froxel range: clamp before the floor
    const float fz = floorf(z), c = z - fz;
    const int k0 = clampi(int(fz), 0, t.d - 1), k1 = clampi(int(fz) + 1, 0, t.d - 1);
    const float hv = 0.5f / float(t.h);
    const float vc = fminf(fmaxf(v, hv), 1.0f - hv);
    const float inv_d = 1.0f / float(t.d);
    const float a = tex2D<float>(t.tex, u, (float(k0) + vc) * inv_d), b = tex2D<float>(t.tex, u, (float(k1) + vc) * inv_d);
    return a + c * (b - a);
}
SKY_D float sample_froxel(const FroxelView& t, float u, float v, float w) {
    float x = u * float(t.w) - 0.5f, y = v * float(t.h) - 0.5f, z = w * float(t.d) - 0.5f;
    float fx = floorf(x), fy = floorf(y), fz = floorf(z);
    float a = x - fx, b = y - fy, c = z - fz;
    int i[2], j[2], k[2];
    i[0] = clampi(int(fx), 0, t.w - 1); i[1] = clampi(int(fx) + 1, 0, t.w - 1);
    j[0] = clampi(int(fy), 0, t.h - 1); j[1] = clampi(int(fy) + 1, 0, t.h - 1);
    // z can be far outside (dist >> froxel range): clamp before the int conversion
    int kz = int(fminf(fmaxf(fz, -1.0f), float(t.d)));
    k[0] = clampi(kz, 0, t.d - 1); k[1] = clampi(kz + 1, 0, t.d - 1);
    float r = 0.0f;
#pragma unroll
    for (int dk = 0; dk < 2; ++dk)
#pragma unroll
        for (int dj = 0; dj < 2; ++dj)
#pragma unroll
            for (int di = 0; di < 2; ++di) {
                float wt = (di ? a : 1.0f - a) * (dj ? b : 1.0f - b) * (dk ? c : 1.0f - c);
                r += wt * (float(__ldg(t.p + (size_t(k[dk]) * t.h + j[dj]) * t.w + i[di])) / 65535.0f);
            }
    return r;
}
// VolumetricCloudShadowInterface.glsl:10-13
SKY_D float SampleRayScatterVisibility(const FroxelView& froxel, float2 uv, float dist, float inv_max_dist) {
    float w = dist * inv_max_dist;
    return mixf(1.0f, sample_froxel(froxel, uv.x, uv.y, w), clampf(1.0f / w, 0.0f, 1.0f));
}
template <bool TEXLUT>
SKY_D float SampleRayScatterVisibilitySel(const FroxelView& froxel, float2 uv, float dist, float inv_max_dist) {
    if (!TEXLUT || !froxel.tex) return SampleRayScatterVisibility(froxel, uv, dist, inv_max_dist);
    float w = dist * inv_max_dist;
    return mixf(1.0f, sample_froxel_atlas(froxel, uv.x, uv.y, w), clampf(1.0f / w, 0.0f, 1.0f));
}

template <bool TEXLUT>
SKY_D float4 sample_lut3d_sel(const LutView& t, float u, float v, float w) {
    if (TEXLUT) return sample_lut3d_atlas(t, u, v, w);
    return sample_lut3d(t, u, v, w);
}
