// K1-K6: atmosphere LUT bake and the sky / aerial-perspective composite, sm_100a.
// Follows shaders/SkyRendering/Atmosphere.glsl and AtmosphereRenderer.glsl (line references on
// each function).  Compiled with -fmad=false: these kernels are microsecond-scale and
// latency-bound, and the altitude r_i - bottom_radius (Atmosphere.glsl:258-260) cancels ~7 digits,
// so keeping the reference's unfused fp32 operation order is worth more than the FMAs.
#include <cstdlib>
#include "../../include/sky_detmath.h"
#include "../../include/skyb200.h"
#include "atmosphere_dev.cuh"
#include "ibl_dev.cuh"
#include "context.h"

// The LUT kernels use the deterministic fp32 elementary functions shared with the oracle (bit-for-bit
// comparable results); the composite translation unit uses the hardware intrinsics.
#if defined(SKY_COMPOSITE_TU) && !defined(SKY_STRICT_TU)
#define LUT_EXP(x) __expf(x)
#define LUT_COS(x) __cosf(x)
#define LUT_SIN(x) __sinf(x)
#define LUT_ACOS(x) acosf(x)
#define LUT_ASIN(x) asinf(x)
#else
#define LUT_EXP(x) sky_det_expf(x)
#define LUT_COS(x) sky_det_cosf(x)
#define LUT_SIN(x) sky_det_sinf(x)
#define LUT_ACOS(x) sky_det_acosf(x)
#define LUT_ASIN(x) sky_det_asinf(x)
#endif

namespace {

SKY_D float3 lut_exp3(float3 a) { return f3(LUT_EXP(a.x), LUT_EXP(a.y), LUT_EXP(a.z)); }

struct BakeParams {
    AtmosphereModel atm;
    LutView transmittance;
    float4* transmittance_out;
    float4* multiscattering_out;
    int ms_w, ms_h;
};

// Atmosphere.glsl:119-132
SKY_D float3 GetExtinction(const SkyAtmosphereBufferData& u, float altitude) {
    float3 rayleigh_extinction = f3(u.rayleigh_scattering) * clampf(LUT_EXP(-altitude * u.inv_rayleigh_exponential_distribution), 0.0f, 1.0f);
    float3 mie_extinction = (f3(u.mie_scattering) + f3(u.mie_absorption)) *
                            clampf(LUT_EXP(-altitude * u.inv_mie_exponential_distribution), 0.0f, 1.0f);
    float3 ozone_extinction = f3(u.ozone_absorption) *
                              fmaxf(0.0f, altitude < u.ozone_center_altitude ? 1.0f + (altitude - u.ozone_center_altitude) * u.inv_ozone_width
                                                                            : 1.0f - (altitude - u.ozone_center_altitude) * u.inv_ozone_width);
    return rayleigh_extinction + mie_extinction + ozone_extinction;
}

// Atmosphere.glsl:220-295.  MS = MULTISCATTERING_COMPUTE_PROGRAM permutation.
// K6's per-pixel raymarch fetches the bake LUTs through texture objects (8-bit filter weights) in the production object,
// with exact fp32 software filtering in the strict one
#ifdef SKY_STRICT_TU
constexpr bool kCompositeTexLut = false;
#else
constexpr bool kCompositeTexLut = true;
#endif
// The two optional terms of the march: MOON_SHADOW_ENABLE (Atmosphere.glsl:190-218,281-284) and VOLUMETRIC_LIGHT_ENABLE
// (:180-188,274-277); the host writes them into the shader text, here they are a template flag (EXTRA) + this block.
struct ScatterExtras {
    int moon_shadow;
    float moon_radius;
    float moon_position[3];
    int shadow_size;             // 0: VOLUMETRIC_LIGHT_ENABLE off
    const float* shadow_map;     // float[S][S] light-space depth (ShadowMap.cpp:8-27)
    float light_view_projection[16];
};

// Atmosphere.glsl:180-188 through Samplers::GetShadowMapSampler (Samplers.cpp:43-51): LINEAR, CLAMP_TO_BORDER (1),
// compare LEQUAL -- the bilinear blend of the four comparison results
SKY_D float GetVisibilityFromShadowMap(const ScatterExtras& e, float3 position) {
    const float* m = e.light_view_projection;
    float X = m[0] * position.x + m[4] * position.y + m[8] * position.z + m[12] * 1.0f;
    float Y = m[1] * position.x + m[5] * position.y + m[9] * position.z + m[13] * 1.0f;
    float Z = m[2] * position.x + m[6] * position.y + m[10] * position.z + m[14] * 1.0f;
    float Wc = m[3] * position.x + m[7] * position.y + m[11] * position.z + m[15] * 1.0f;
    float sx = X / Wc * 0.5f + 0.5f, sy = Y / Wc * 0.5f + 0.5f, depth = Z / Wc * 0.5f + 0.5f;
    if (depth >= 1.0f) return 1.0f;
    const float S = float(e.shadow_size);
    float x = sx * S - 0.5f, y = sy * S - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    auto cmp = [&](float i, float j) {
        bool inside = i >= 0.0f && j >= 0.0f && i < S && j < S;
        float texel = inside ? __ldg(e.shadow_map + size_t(int(j)) * e.shadow_size + int(i)) : 1.0f;
        return depth <= texel ? 1.0f : 0.0f;
    };
    return (1.0f - a) * (1.0f - b) * cmp(fx, fy) + a * (1.0f - b) * cmp(fx + 1.0f, fy) + (1.0f - a) * b * cmp(fx, fy + 1.0f) +
           a * b * cmp(fx + 1.0f, fy + 1.0f);
}

// Atmosphere.glsl:190-210
SKY_D float GetVisibilityFromMoonShadow(float sun_moon_angular_distance, float sun_angular_radius, float moon_angular_radius) {
    float max_radius = sun_angular_radius + moon_angular_radius;
    float min_radius = fabsf(sun_angular_radius - moon_angular_radius);
    float sun_r2 = sun_angular_radius * sun_angular_radius;
    float moon_r2 = moon_angular_radius * moon_angular_radius;
    if (sun_moon_angular_distance >= max_radius) return 1.0f;
    if (sun_moon_angular_distance <= min_radius) return clampf((sun_r2 - moon_r2) / sun_r2, 0.0f, 1.0f);
    float distance2 = sun_moon_angular_distance * sun_moon_angular_distance;
    float cos_half_sun = (distance2 + sun_r2 - moon_r2) / (2 * sun_moon_angular_distance * sun_angular_radius);
    float cos_half_moon = (distance2 + moon_r2 - sun_r2) / (2 * sun_moon_angular_distance * moon_angular_radius);
    float half_sun = LUT_ACOS(cos_half_sun);
    float half_moon = LUT_ACOS(cos_half_moon);
    float triangle_h = sun_angular_radius * sqrtf(1 - cos_half_sun * cos_half_sun);
    float area_total = (kPi - half_sun) * sun_r2 + (kPi - half_moon) * moon_r2 + triangle_h * sun_moon_angular_distance;
    float area_uncovered = area_total - kPi * moon_r2;
    return area_uncovered / (kPi * sun_r2);
}
// Atmosphere.glsl:212-218
SKY_D float GetVisibilityFromMoonShadow(float3 moon_vector, float moon_radius, float3 sun_direction, float sun_angular_radius) {
    float inv_moon_distance = 1.0f / sqrtf(dot(moon_vector, moon_vector));
    float3 moon_direction = moon_vector * inv_moon_distance;
    float moon_angular_radius = LUT_ASIN(clampf(moon_radius * inv_moon_distance, -1.0f, 1.0f));
    return GetVisibilityFromMoonShadow(LUT_ACOS(clampf(dot(sun_direction, moon_direction), -1.0f, 1.0f)), sun_angular_radius, moon_angular_radius);
}

// One step of the march of Atmosphere.glsl:244-290: everything that depends on the step alone (not on the running
// transmittance).  A = luminance_i - luminance_i * transmittance_i, B = scattering_i - scattering_i * transmittance_i.
struct MarchSetup {
    float r, mu, dx, rayleigh_phase, mie_phase;
};
struct MarchStep {
    float3 A, B, extinction, transmittance;
};
template <bool MS>
SKY_D MarchSetup march_setup(const AtmosphereModel& atm, float3 earth_center, float3 start_position, float3 view_direction,
                             float3 sun_direction, float marching_distance, float steps) {
    const SkyAtmosphereBufferData& u = atm.u;
    MarchSetup m;
    m.r = length(start_position - earth_center);
    float3 up_direction = normalize(start_position - earth_center);
    m.mu = dot(view_direction, up_direction);
    float cos_sun_view = dot(view_direction, sun_direction);
    m.dx = marching_distance / steps;
    if (MS) {
        m.rayleigh_phase = m.mie_phase = 1.0f / (4.0f * kPi);  // IsotropicPhaseFunction, :134-136
    } else {
        // RayleighPhaseFunction / MiePhaseFunction, :138-154
        m.rayleigh_phase = (3.0f / (16.0f * kPi)) * (1.0f + cos_sun_view * cos_sun_view);
        float g = u.mie_phase_g;
        float k = 3.0f / (8.0f * kPi) * (1.0f - g * g) / (2.0f + g * g);
        m.mie_phase = k * (1.0f + cos_sun_view * cos_sun_view) / sky_det_pow15f(1.0f + g * g - 2.0f * g * cos_sun_view);
    }
    return m;
}
template <bool MS, bool TEXLUT, bool EXTRA>
SKY_D MarchStep march_step(const AtmosphereModel& atm, const LutView& transmittance_texture, const LutView& multiscattering_texture,
                           const MarchSetup& m, float i, float3 earth_center, float3 start_position, float3 view_direction,
                           float3 sun_direction, const ScatterExtras* extras) {
    const SkyAtmosphereBufferData& u = atm.u;
    const float r = m.r, mu = m.mu, dx = m.dx;
    float d_i = i * dx;
    float r_i = sqrtf(d_i * d_i + 2.0f * r * mu * d_i + r * r);
    float3 position_i = start_position + view_direction * d_i;
    float altitude_i = r_i - u.bottom_radius;

    // GetScattering, :156-159
    float3 rayleigh_scattering_i = f3(u.rayleigh_scattering) * clampf(LUT_EXP(-altitude_i * u.inv_rayleigh_exponential_distribution), 0.0f, 1.0f);
    float3 mie_scattering_i = f3(u.mie_scattering) * clampf(LUT_EXP(-altitude_i * u.inv_mie_exponential_distribution), 0.0f, 1.0f);
    float3 scattering_i = rayleigh_scattering_i + mie_scattering_i;
    float3 scattering_with_phase_i = rayleigh_scattering_i * m.rayleigh_phase + mie_scattering_i * m.mie_phase;

    float3 extinction_i = GetExtinction(u, altitude_i);
    float3 transmittance_i = lut_exp3(-extinction_i * dx);
    float3 up_direction_i = normalize(position_i - earth_center);
    float mu_s_i = dot(sun_direction, up_direction_i);
    float3 luminance_i = scattering_with_phase_i * atm.template GetSunVisibility<TEXLUT>(transmittance_texture, r_i, mu_s_i);
    if (!MS) {
        if (EXTRA && extras->shadow_size > 0) luminance_i *= GetVisibilityFromShadowMap(*extras, position_i);  // :274-277
        // GetMultiscatteringContribution, :169-178
        float x_mu_s = mu_s_i * 0.5f + 0.5f;
        float x_r = (r_i - u.bottom_radius) / (u.top_radius - u.bottom_radius);
        float uu = 0.5f / float(multiscattering_texture.w) + x_mu_s * (1.0f - 1.0f / float(multiscattering_texture.w));
        float vv = 0.5f / float(multiscattering_texture.h) + x_r * (1.0f - 1.0f / float(multiscattering_texture.h));
        float3 multiscattering_contribution = xyz(sample_lut2d_sel<TEXLUT>(multiscattering_texture, uu, vv));
        luminance_i += u.multiscattering_mask * multiscattering_contribution * scattering_i;
        if (EXTRA && extras->moon_shadow)  // :281-284
            luminance_i *= GetVisibilityFromMoonShadow(f3(extras->moon_position) - position_i, extras->moon_radius, sun_direction, u.sun_angular_radius);
        luminance_i *= f3(u.solar_illuminance);
    }
    MarchStep s;
    s.A = luminance_i - luminance_i * transmittance_i;
    s.B = MS ? scattering_i - scattering_i * transmittance_i : f3(0.0f);
    s.extinction = extinction_i;
    s.transmittance = transmittance_i;
    return s;
}

// Steps evaluated side by side per loop trip (bit-identical for any value, see the loop).  Measured on B200 (profiles/lut_variants_r02t.log,
// scene c1 / c3, K3-K5): 1 -> 137 / 167 us, 2 -> 163 / 247 us, 4 -> 200 / 277 us -- the extra independent work does not pay for the
// predicated tail evaluations and the registers, so the shader's own one-step loop stays.
#ifndef SKY_MARCH_STEPS
#define SKY_MARCH_STEPS 1
#endif
// Atmosphere.glsl:220-295, one thread per march.  MS = MULTISCATTERING_COMPUTE_PROGRAM permutation.
template <bool MS, bool TEXLUT = false, bool EXTRA = false>
SKY_D float3 ComputeScatteredLuminance(const AtmosphereModel& atm, const LutView& transmittance_texture,
                                       const LutView& multiscattering_texture, float start_i, float3 earth_center,
                                       float3 start_position, float3 view_direction, float3 sun_direction,
                                       float marching_distance, float steps, float3& transmittance, float3& L_f,
                                       const ScatterExtras* extras = nullptr) {
    const float SAMPLE_COUNT = steps;
    const MarchSetup m = march_setup<MS>(atm, earth_center, start_position, view_direction, sun_direction, marching_distance, steps);
    transmittance = f3(1.0f);
    float3 luminance = f3(0.0f);
    if (MS) { L_f = f3(0.0f); start_i = 0.5f; }
#if SKY_MARCH_STEPS > 1
    // SKY_MARCH_STEPS steps per trip.  A step's evaluation (march_step: the exp / sqrt / LUT-fetch chain, ~95 % of the work) depends on
    // nothing the loop carries -- only the three accumulations below do -- so evaluating steps i, i + 1, ... side by side multiplies the
    // independent instructions a thread can have in flight.  The LUT kernels are a few warps per SM working through long dependent marches
    // (K2: two warps per texel, 30 steps), so this is what they are short of.  Same operations on the same operands in the same order:
    // the same bits.
    for (float i = start_i; i < SAMPLE_COUNT;) {
        float idx[SKY_MARCH_STEPS];
        bool live[SKY_MARCH_STEPS];
        MarchStep st[SKY_MARCH_STEPS];
#pragma unroll
        for (int q = 0; q < SKY_MARCH_STEPS; ++q) {
            idx[q] = q == 0 ? i : idx[q - 1] + 1.0f;    // the loop's own `++i`
            live[q] = idx[q] < SAMPLE_COUNT;
        }
#pragma unroll
        for (int q = 0; q < SKY_MARCH_STEPS; ++q)
            st[q] = march_step<MS, TEXLUT, EXTRA>(atm, transmittance_texture, multiscattering_texture, m, live[q] ? idx[q] : i, earth_center, start_position,
                                                  view_direction, sun_direction, extras);
#pragma unroll
        for (int q = 0; q < SKY_MARCH_STEPS; ++q) {
            if (live[q]) {
                luminance += transmittance * st[q].A / st[q].extinction;
                if (MS) L_f += transmittance * st[q].B / st[q].extinction;
                transmittance *= st[q].transmittance;
            }
        }
        i = idx[SKY_MARCH_STEPS - 1] + 1.0f;
    }
#else
    for (float i = start_i; i < SAMPLE_COUNT; ++i) {
        const MarchStep s = march_step<MS, TEXLUT, EXTRA>(atm, transmittance_texture, multiscattering_texture, m, i, earth_center, start_position,
                                                          view_direction, sun_direction, extras);
        luminance += transmittance * s.A / s.extinction;
        if (MS) L_f += transmittance * s.B / s.extinction;
        transmittance *= s.transmittance;
    }
#endif
    return luminance;
}

#if defined(SKY_COMPOSITE_TU) && !defined(SKY_STRICT_TU)
// The production march of K6 (the composite translation unit, -use_fast_math, frame tolerance 1e-2 relative RMS): the SAME
// quadrature as ComputeScatteredLuminance<false, true, EXTRA> above -- same steps, same midpoint rule, same two LUT fetches per
// step, same analytic in-segment integral -- with the per-step algebra reduced to what actually depends on the step.  ncu on the
// statement-by-statement version (profiles/k6_r01i.md): 148 warp instructions and 17 MUFU per step, 80 % issue-active with the
// XU pipe at ~73 % -- instruction-bound.  Here (96 instructions, 12 MUFU per step):
//   * |position_i - earth_center| IS r_i, and dot(sun, position_i - earth_center) is linear in d_i: mu_s_i = (a_s + b_s d_i) / r_i
//     without building or normalising position_i (it is only built for the optional EXTRA terms);
//   * r_i^2 = dd + r^2 with dd = d_i (d_i + 2 r mu); rho^2 = r_i^2 - bottom^2 = dd + (r^2 - bottom^2) and the discriminant of
//     DistanceToTopAtmosphereBoundary = (r_i mu_s)^2 + (top^2 - r^2) - dd: the ray constants carry the large cancelling terms once,
//     in fp32 but from (r - bottom)(r + bottom), which is MORE accurate than the shader's r_i * r_i - bottom * bottom;
//   * one MUFU.RSQ gives r_i and 1 / r_i; cos_theta_h = -rho / r_i and the smoothstep's 1 / (e1 - e0) = r_i / (2 bottom alpha), so
//     GetSunVisibility's edge term is ((r_i mu_s + rho) / (2 bottom alpha) + 1/2) saturated: no division, no extra square root;
//   * every uniform-only factor (phase x scattering x solar illuminance, log2(e) / scale height, LUT texel insets) is folded once per ray.
// The strict object keeps the shader's statement order (bit-level parity); this one is checked against the oracle at frame tolerance
// (tests/test_gpu_parity.py: C2 / C3 at 1920x1080, C4 at 3840x2160).
//   * the solar illuminance multiplies every term of a step: it is applied once to the finished sum; the multiscattering mask is folded
//     into the RGBA16F copy of the multiscattering LUT when the copy is made (k_luts_to_half); a grey Mie term (rgb-equal scattering and
//     absorption coefficients, as in every shipped scene -- checked per launch, template flag GREY) is one scalar per step, not three.
template <bool EXTRA, bool GREY>
SKY_D float3 ComputeScatteredLuminanceFast(const AtmosphereModel& atm, const LutView& transmittance_texture, const LutView& multiscattering_texture,
                                           cudaTextureObject_t density_texture, float start_i, float3 earth_center, float3 start_position, float3 view_direction, float3 sun_direction,
                                           float marching_distance, float steps, float3& transmittance, const ScatterExtras* extras) {
    const SkyAtmosphereBufferData& u = atm.u;
    const float kLog2e = 1.4426950408889634f;
    const float3 rel = start_position - earth_center;
    const float r2 = dot(rel, rel);
    const float two_rmu = 2.0f * dot(view_direction, rel);               // 2 r mu
    const float a_s = dot(sun_direction, rel), b_s = dot(sun_direction, view_direction);  // r_i mu_s_i = a_s + b_s d_i
    const float cos_sun_view = b_s;
    const float dx = marching_distance / steps;
    // phase functions (Atmosphere.glsl:138-154) x scattering coefficients x solar illuminance
    const float rayleigh_phase = (3.0f / (16.0f * kPi)) * (1.0f + cos_sun_view * cos_sun_view);
    const float g = u.mie_phase_g;
    const float mie_k = 3.0f / (8.0f * kPi) * (1.0f - g * g) / (2.0f + g * g);
    const float mie_b = 1.0f + g * g - 2.0f * g * cos_sun_view;
    const float mie_phase = mie_k * (1.0f + cos_sun_view * cos_sun_view) / (mie_b * sqrtf(mie_b));
    const float3 solar = f3(u.solar_illuminance);
    const float3 Rs = f3(u.rayleigh_scattering), Ms = f3(u.mie_scattering), Ma = f3(u.mie_absorption), Oz = f3(u.ozone_absorption);
    const float3 RsP = Rs * rayleigh_phase, MsP = Ms * mie_phase;   // single scattering with the phase folded (x solar at the end)
    const float3 Me = Ms + Ma;                                          // Mie extinction
    // both altitude tables are addressed with r_i directly: the `- bottom` of the altitude sits in the offsets
    const float dn_a = (1.0f - 1.0f / float(kDensityLutSize)) / (u.top_radius - u.bottom_radius), dn_b = 0.5f / float(kDensityLutSize) - u.bottom_radius * dn_a;
    const float k_t = -dx * kLog2e;
    // ray constants of the (r, mu_s) -> transmittance-LUT mapping (Atmosphere.glsl:90-108)
    const float bottom = u.bottom_radius, top = u.top_radius;
    const float H = sqrtf((top - bottom) * (top + bottom));
    const float r0 = sqrtf(r2);
    const float r2_minus_b2 = (r0 - bottom) * (r0 + bottom), t2_minus_r2 = (top - r0) * (top + r0);
    const float tw = float(transmittance_texture.w), th = float(transmittance_texture.h);
    const float uu_a = 1.0f - 1.0f / tw, uu_b = 0.5f / tw, vv_a = (1.0f - 1.0f / th) / H, vv_b = 0.5f / th;
    const float H_minus_top = H - top;
    const float edge_k = 0.5f / (bottom * u.sun_angular_radius);
    // multiscattering LUT (Atmosphere.glsl:169-178): u from mu_s, v from the altitude
    const float mw = float(multiscattering_texture.w), mh = float(multiscattering_texture.h);
    const float mu_a = 0.5f * (1.0f - 1.0f / mw), mu_b = 0.5f / mw + mu_a;
    const float mv_a = (1.0f - 1.0f / mh) / (top - bottom), mv_b = 0.5f / mh - bottom * mv_a;

    float3 T = f3(1.0f), L = f3(0.0f);
    // the shader's `for (float i = start_i; i < SAMPLE_COUNT; ++i)` with an integer trip count (start_i is in [0, 1): the number of
    // i = start_i + k below `steps` is ceil(steps - start_i)), so that the compiler may unroll it
    const int trip = max(int(ceilf(steps - start_i)), 0);
    float i = start_i;
#ifndef SKY_K6_UNROLL
#define SKY_K6_UNROLL 4   // measured at 4K, scene c3 (profiles/k6_variants_r02A.log): 1 -> 642 us, 2 -> 613 us, 4 -> 600 us; 4 blocks/SM (64 registers): 629-669 us
#endif
    constexpr int kUnroll = SKY_K6_UNROLL;
#pragma unroll kUnroll
    for (int k = 0; k < trip; ++k, i += 1.0f) {
        const float d = i * dx;
        const float dd = d * (d + two_rmu);
        const float ri2 = dd + r2;
        const float inv_r = rsqrtf(ri2);
        const float r_i = ri2 * inv_r;
        const float rms = a_s + b_s * d;            // r_i mu_s_i
        const float mu_s = rms * inv_r;
        // densities (GetScattering / GetExtinction, :119-132,156-159): one fetch of the altitude table
        const float4 dens = tex2D<float4>(density_texture, dn_b + r_i * dn_a, 0.5f);
        const float dR = dens.x, dM = dens.y, dO = dens.z;
        float3 scattering, extinction, single;
        if (GREY) {
            const float sM = Ms.x * dM, eM = Me.x * dM, pM = MsP.x * dM;
            scattering = f3(Rs.x * dR + sM, Rs.y * dR + sM, Rs.z * dR + sM);
            extinction = f3(Rs.x * dR + eM, Rs.y * dR + eM, Rs.z * dR + eM) + Oz * dO;
            single = f3(RsP.x * dR + pM, RsP.y * dR + pM, RsP.z * dR + pM);
        } else {
            scattering = Rs * dR + Ms * dM;
            extinction = scattering + Ma * dM + Oz * dO;
            single = f3(RsP.x * dR + MsP.x * dM, RsP.y * dR + MsP.y * dM, RsP.z * dR + MsP.z * dM);
        }
        const float3 T_i = f3(exp2f(extinction.x * k_t), exp2f(extinction.y * k_t), exp2f(extinction.z * k_t));
        // GetSunVisibility (:110-117) = LUT(r_i, mu_s) x smoothstep around the geometric horizon
        const float rho = sqrtf(fmaxf(dd + r2_minus_b2, 0.0f));
        const float disc = fmaxf(rms * rms + (t2_minus_r2 - dd), 0.0f);
        const float d_top = fmaxf(sqrtf(disc) - rms, 0.0f);
        const float d_min = top - r_i;
        const float x_mu = (d_top - d_min) / (rho + r_i + H_minus_top);   // (d - d_min) / (d_max - d_min)
        const float4 t_sun = tex2D<float4>(transmittance_texture.tex, uu_b + x_mu * uu_a, vv_b + rho * vv_a);
        const float4 ms = tex2D<float4>(multiscattering_texture.tex, mu_b + mu_s * mu_a, mv_b + r_i * mv_a);   // x multiscattering_mask already
        const float e = __saturatef((rms + rho) * edge_k + 0.5f);
        float vis = e * e * (3.0f - 2.0f * e);
        float3 position_i;
        if (EXTRA) {
            position_i = start_position + view_direction * d;
            if (extras->shadow_size > 0) vis *= GetVisibilityFromShadowMap(*extras, position_i);  // :274-277 (single scattering only)
        }
        float3 L_i = single * (f3(t_sun.x, t_sun.y, t_sun.z) * vis) + f3(ms.x, ms.y, ms.z) * scattering;
        if (EXTRA && extras->moon_shadow)  // :281-284
            L_i *= GetVisibilityFromMoonShadow(f3(extras->moon_position) - position_i, extras->moon_radius, sun_direction, u.sun_angular_radius);
        // analytic integral over the segment (:288): (L_i - L_i T_i) / extinction, attenuated by the transmittance so far
        // (one MUFU.RCP for the three channels: 1 / x = y z / (x y z); the products stay far inside the fp32 range for extinction
        // coefficients per km, and a vanishing extinction gives the same 0 x inf = NaN the shader's own division produces)
#ifdef SKY_K6_RCP3   // experiment: three MUFU.RCP instead of one + six multiplications
        L += T * (L_i - L_i * T_i) * f3(1.0f / extinction.x, 1.0f / extinction.y, 1.0f / extinction.z);
#else
        const float xy = extinction.x * extinction.y;
        const float inv_xyz = 1.0f / (xy * extinction.z);
        const float inv_z = xy * inv_xyz, inv_xy_z = extinction.z * inv_xyz;
        L += T * (L_i - L_i * T_i) * f3(extinction.y * inv_xy_z, extinction.x * inv_xy_z, inv_z);
#endif
        T *= T_i;
    }
    transmittance = T;
    return L * solar;
}
#endif

// (Measured and removed: the same march spread over the 32 lanes of a warp -- lane l evaluates steps l, l + 32, ..., then every
// lane replays the running transmittance product / luminance sum over the broadcast records in the reference's order.  It is
// bit-identical, but K2 went 78 -> 520 us and K3-K5 230 -> 470 us: these kernels already issue at ~50 % of the machine with one
// thread per march, and the replicated serial part costs 5x the instructions.)

#ifndef SKY_COMPOSITE_TU
// ---- the production LUT march (sky_set_lut_arithmetic(ctx, SKY_LUT_COOPERATIVE)) ----------------------------------------------
// What differs from the march above, and why it may.  The bit-exact kernels are one thread per march: a serial chain of 30-51 steps
// of ~450 instructions each (IEEE division / square root, the deterministic exp), i.e. the LATENCY of one thread, which no
// amount of sharding shortens (K2 78 us, K3+K4 80 us on B200).  Only two things in a march are sequential, though: the running
// transmittance product and the sum it weights.  A step is the affine map (L, T) -> (L + T A_i, T T_i), affine maps compose
// associatively, so LANES lanes each fold a contiguous chunk of the steps into one map and a shuffle tree composes the chunks in
// order.  The per-step arithmetic uses FMAs and the hardware ex2 / rcp / sqrt approximations where the result is well conditioned,
// and keeps the shader's own unfused expression for the one quantity that is not: r_i (Atmosphere.glsl:258), whose altitude
// r_i - bottom_radius cancels seven digits (one ulp of r_i is 0.5 m, 4e-4 of the Mie density).  The ray set-up (directions,
// boundary distances, phase functions) is the exact code above, so every lane group marches the reference's own segment.
// Tolerance instead of bit-exactness: tests/test_gpu_parity.py::test_cooperative_lut_bake_within_tolerance.
#ifndef SKY_LUT_COOP_THREADS   // resident threads per SM the cooperative kernels are compiled for (register budget = 65536 / this)
#define SKY_LUT_COOP_THREADS 768
#endif
#ifndef SKY_LUT_COOP_UNROLL
#define SKY_LUT_COOP_UNROLL 1
#endif
SKY_D float ap_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
SKY_D float ap_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
SKY_D float ap_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
SKY_D float3 fma3(float3 a, float3 b, float3 c) { return f3(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y), fmaf(a.z, b.z, c.z)); }
SKY_D float3 fma3(float3 a, float s, float3 c) { return f3(fmaf(a.x, s, c.x), fmaf(a.y, s, c.y), fmaf(a.z, s, c.z)); }
SKY_D float3 shfl_down3(unsigned mask, float3 v, int offset, int width) {
    return f3(__shfl_down_sync(mask, v.x, offset, width), __shfl_down_sync(mask, v.y, offset, width), __shfl_down_sync(mask, v.z, offset, width));
}
// (1 - exp(-x)) / sigma with x = sigma dx: the segment weight of Atmosphere.glsl:288.  For x < 1/32 the alternating series
// dx (1 - x/2 + x^2/6 - x^3/24 + x^4/120) instead of the cancelling difference (truncation < 5e-11): measured against the oracle
// (profiles/lut_coop_r02D.log, mode 2 vs mode 1) it halves the count of texels beyond 1e-4 at the same speed.
SKY_D float segment_weight(float sigma, float x, float T_i, float dx) {
    const float w = (1.0f - T_i) * ap_rcp(sigma);
    const float p = fmaf(x, fmaf(x, fmaf(x, fmaf(x, 1.0f / 120.0f, -1.0f / 24.0f), 1.0f / 6.0f), -0.5f), 1.0f);
    return x < 0.03125f ? dx * p : w;
}

// Atmosphere.glsl:220-295 for the lane group `group_mask` (LANES consecutive lanes, this lane is number `lane` of it); the result is
// valid in lane 0 of the group.  Every lane of the group must call it with the same ray.
template <bool MS, int LANES>
SKY_D float3 ComputeScatteredLuminanceCooperative(const AtmosphereModel& atm, const LutView& transmittance_texture, const LutView& multiscattering_texture,
                                                  float start_i, float3 earth_center, float3 start_position, float3 view_direction, float3 sun_direction,
                                                  float marching_distance, float steps, int lane, unsigned group_mask, float3& transmittance, float3& L_f) {
    const SkyAtmosphereBufferData& u = atm.u;
    const MarchSetup m = march_setup<MS>(atm, earth_center, start_position, view_direction, sun_direction, marching_distance, steps);
    if (MS) start_i = 0.5f;
    // the shader's `for (float i = start_i; i < SAMPLE_COUNT; ++i)`: start_i is in [0, 1), so it runs ceil(steps - start_i) times
    const int trip = max(int(ceilf(steps - start_i)), 0);
    const int per = (trip + LANES - 1) / LANES;
    const int k_begin = lane * per, k_end = min(k_begin + per, trip);

    const float kLog2e = 1.4426950408889634f;
    const float bottom = u.bottom_radius, top = u.top_radius;
    const float dx = m.dx, r2 = m.r * m.r, two_rmu = 2.0f * m.r * m.mu;      // the shader's own terms of r_i
    const float3 rel = start_position - earth_center;
    const float a_s = dot(sun_direction, rel), b_s = dot(sun_direction, view_direction);   // r_i mu_s_i = a_s + b_s d_i
    const float kR = -u.inv_rayleigh_exponential_distribution * kLog2e, kM = -u.inv_mie_exponential_distribution * kLog2e;
    const float3 Rs = f3(u.rayleigh_scattering), Ms = f3(u.mie_scattering), Ma = f3(u.mie_absorption), Oz = f3(u.ozone_absorption);
    const float3 RsP = Rs * m.rayleigh_phase, MsP = Ms * m.mie_phase;
    const float k_t = -dx * kLog2e;
    // (r, mu_s) -> transmittance LUT coordinates (Atmosphere.glsl:90-108), ray-invariant factors
    const float H = sqrtf((top - bottom) * (top + bottom));
    const float tw = float(transmittance_texture.w), th = float(transmittance_texture.h);
    const float uu_a = 1.0f - 1.0f / tw, uu_b = 0.5f / tw, vv_a = (1.0f - 1.0f / th) / H, vv_b = 0.5f / th;
    const float edge_k = 0.5f / (bottom * u.sun_angular_radius);
    // multiscattering LUT coordinates (:169-178)
    const float mw = float(multiscattering_texture.w), mh = float(multiscattering_texture.h);
    const float mu_a = 0.5f * (1.0f - 1.0f / mw), mu_b = 0.5f / mw + mu_a;
    const float mv_a = (1.0f - 1.0f / mh) / (top - bottom), mv_b = 0.5f / mh;

    float3 T = f3(1.0f), L = f3(0.0f), Lf = f3(0.0f);
    constexpr int kCoopUnroll = SKY_LUT_COOP_UNROLL;
#pragma unroll kCoopUnroll
    for (int k = k_begin; k < k_end; ++k) {
        const float d = (start_i + float(k)) * dx;
        const float r_i = sqrtf(d * d + two_rmu * d + r2);          // unfused, IEEE: the shader's expression (:258)
        const float altitude = r_i - bottom;
        const float dR = __saturatef(ap_ex2(altitude * kR)), dM = __saturatef(ap_ex2(altitude * kM));
        const float dO = fmaxf(0.0f, fmaf(-fabsf(altitude - u.ozone_center_altitude), u.inv_ozone_width, 1.0f));
        const float3 scattering = fma3(Ms, dM, Rs * dR);
        const float3 extinction = fma3(Oz, dO, fma3(Ma, dM, scattering));
        const float3 x = extinction * dx;
        const float3 T_i = f3(ap_ex2(extinction.x * k_t), ap_ex2(extinction.y * k_t), ap_ex2(extinction.z * k_t));
        // GetSunVisibility (:110-117): rho^2 = r_i^2 - bottom^2 = altitude (r_i + bottom), the top-boundary discriminant
        // r_i^2 (mu_s^2 - 1) + top^2 = (r_i mu_s)^2 + (top - r_i)(top + r_i): no cancelling squares
        const float rms = fmaf(b_s, d, a_s);
        const float rho = ap_sqrt(fmaxf(altitude * (r_i + bottom), 0.0f));
        const float d_min = top - r_i;
        const float d_top = fmaxf(ap_sqrt(fmaxf(fmaf(rms, rms, d_min * (top + r_i)), 0.0f)) - rms, 0.0f);
        const float x_mu = (d_top - d_min) * ap_rcp(rho + H - d_min);
        const float3 t_sun = xyz(sample_lut2d(transmittance_texture, fmaf(x_mu, uu_a, uu_b), fmaf(rho, vv_a, vv_b)));
        const float e = __saturatef(fmaf(rms + rho, edge_k, 0.5f));   // the smoothstep around the geometric horizon, cos_theta_h = -rho / r_i
        const float vis = e * e * fmaf(-2.0f, e, 3.0f);
        float3 L_i;
        if (MS) {
            L_i = scattering * (t_sun * (vis * m.rayleigh_phase));     // isotropic phase for both species (:134-136)
        } else {
            const float mu_s = rms * ap_rcp(r_i);
            const float3 ms = xyz(sample_lut2d(multiscattering_texture, fmaf(mu_s, mu_a, mu_b), fmaf(altitude, mv_a, mv_b)));
            L_i = fma3(fma3(MsP, dM, RsP * dR), t_sun * vis, (ms * u.multiscattering_mask) * scattering);
        }
        const float3 w = f3(segment_weight(extinction.x, x.x, T_i.x, dx), segment_weight(extinction.y, x.y, T_i.y, dx),
                            segment_weight(extinction.z, x.z, T_i.z, dx));
        L = fma3(T, L_i * w, L);
        if (MS) Lf = fma3(T, scattering * w, Lf);
        T = T * T_i;
    }
    // compose the chunks in order: lane l takes (L, T) o (L', T') = (L + T L', T T') with the map `offset` lanes up
#pragma unroll
    for (int offset = 1; offset < LANES; offset <<= 1) {
        const float3 Lr = shfl_down3(group_mask, L, offset, LANES), Tr = shfl_down3(group_mask, T, offset, LANES);
        L = fma3(T, Lr, L);
        if (MS) Lf = fma3(T, shfl_down3(group_mask, Lf, offset, LANES), Lf);
        T = T * Tr;
    }
    transmittance = T;
    L_f = Lf;
    return MS ? L : L * f3(u.solar_illuminance);
}
#endif  // !SKY_COMPOSITE_TU

// Atmosphere.glsl:297-306
template <bool MS>
SKY_D float3 ComputeGroundLuminance(const AtmosphereModel& atm, const LutView& transmittance_texture, float3 earth_center,
                                    float3 position, float3 sun_direction) {
    float3 up_direction = normalize(position - earth_center);
    float mu_s = dot(sun_direction, up_direction);
    float3 solar_illuminance_at_ground = atm.GetSunVisibility(transmittance_texture, atm.u.bottom_radius, mu_s);
    if (!MS) solar_illuminance_at_ground *= atm.solar_illuminance();
    float3 normal = normalize(position - earth_center);
    return kInvPi * clampf(dot(normal, sun_direction), 0.0f, 1.0f) * atm.ground_albedo() * solar_illuminance_at_ground;
}

// ------------------------------------------------------------------------------------------------- K1
// Atmosphere.glsl:311-337 (+ GetRMuFromTransmittanceTextureIndex :71-88); one thread per texel.
__global__ void __launch_bounds__(128) k1_transmittance(const __grid_constant__ BakeParams P) {
    const SkyAtmosphereBufferData& u = P.atm.u;
    const int W = P.transmittance.w, Hh = P.transmittance.h;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= W || y >= Hh) return;
    float x_mu = float(x) / float(W - 1), x_r = float(y) / float(Hh - 1);
    float H = sqrtf(u.top_radius * u.top_radius - u.bottom_radius * u.bottom_radius);
    float rho = H * x_r;
    float r = sqrtf(rho * rho + u.bottom_radius * u.bottom_radius);
    float d_min = u.top_radius - r;
    float d_max = rho + H;
    float d = d_min + x_mu * (d_max - d_min);
    float mu = d == 0.0f ? 1.0f : (H * H - rho * rho - d * d) / (2.0f * r * d);
    mu = clampf(mu, -1.0f, 1.0f);

    const float SAMPLE_COUNT = u.transmittance_steps;
    float dx = P.atm.DistanceToTopAtmosphereBoundary(r, mu) / SAMPLE_COUNT;
    float3 optical_length = f3(0.0f);
    for (float i = 0.5f; i < SAMPLE_COUNT; ++i) {
        float d_i = i * dx;
        float r_i = sqrtf(d_i * d_i + 2.0f * r * mu * d_i + r * r);
        float altitude_i = r_i - u.bottom_radius;
        optical_length += GetExtinction(u, altitude_i) * dx;
    }
    P.transmittance_out[y * W + x] = f4(lut_exp3(-optical_length), 1.0f);
}

// ------------------------------------------------------------------------------------------------- K2
// Atmosphere.glsl:344-438: one 64-thread block per texel, one sphere direction per thread, pairwise
// tree (i, i+32) ... (i, i+1) kept in the reference's order (:392-425) via smem + shuffles.
__global__ void __launch_bounds__(64) k2_multiscattering(const __grid_constant__ BakeParams P) {
    const SkyAtmosphereBufferData& u = P.atm.u;
    const int gx = blockIdx.x, gy = blockIdx.y;
    const int local_index = threadIdx.x;
    // GetAltitudeMuSFromMultiscatteringTextureIndex, :161-167
    float x_mu_s = float(gx) / float(P.ms_w - 1), x_altitude = float(gy) / float(P.ms_h - 1);
    float altitude = x_altitude * (u.top_radius - u.bottom_radius);
    float mu_s = x_mu_s * 2.0f - 1.0f;
    float3 earth_center = f3(0.0f, -u.bottom_radius, 0.0f);
    float3 start_position = f3(0.0f, altitude, 0.0f);
    float3 sun_direction = f3(0.0f, mu_s, sqrtf(1 - mu_s * mu_s));
    // GetDirectionFromLocalIndex, :344-355
    float unit_theta = (0.5f + float(local_index / 8)) / 8.0f;
    float unit_phi = (0.5f + float(local_index % 8)) / 8.0f;
    float cos_theta = 1.0f - 2.0f * unit_theta;
    float sin_theta = sqrtf(clampf(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
    float phi = 2 * kPi * unit_phi;
    float3 view_direction = f3(LUT_COS(phi) * sin_theta, cos_theta, LUT_SIN(phi) * sin_theta);

    float r = altitude + u.bottom_radius;
    float mu = view_direction.y;
    bool intersect_bottom = P.atm.RayIntersectsGround(r, mu);
    float marching_distance = intersect_bottom ? P.atm.DistanceToBottomAtmosphereBoundary(r, mu) : P.atm.DistanceToTopAtmosphereBoundary(r, mu);
    float3 transmittance, L_f;
    float3 luminance = ComputeScatteredLuminance<true>(P.atm, P.transmittance, P.transmittance, 0.5f, earth_center, start_position,
                                                       view_direction, sun_direction, marching_distance, u.multiscattering_steps,
                                                       transmittance, L_f);
    if (intersect_bottom) {
        float3 ground_position = start_position + view_direction * marching_distance;
        luminance += transmittance * ComputeGroundLuminance<true>(P.atm, P.transmittance, earth_center, ground_position, sun_direction);
    }
    __shared__ float sh[32][6];
    if (local_index >= 32) {
        float* s = sh[local_index - 32];
        s[0] = luminance.x; s[1] = luminance.y; s[2] = luminance.z; s[3] = L_f.x; s[4] = L_f.y; s[5] = L_f.z;
    }
    __syncthreads();
    if (local_index >= 32) return;
    float v[6] = {luminance.x, luminance.y, luminance.z, L_f.x, L_f.y, L_f.z};
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] += sh[local_index][k];
#pragma unroll
    for (int stride = 16; stride >= 1; stride >>= 1)
#pragma unroll
        for (int k = 0; k < 6; ++k) v[k] += __shfl_down_sync(0xffffffffu, v[k], stride);
    if (local_index == 0) {
        float3 L_2nd_order = f3(v[0], v[1], v[2]) / 64.0f;
        float3 f_ms = f3(v[3], v[4], v[5]) / 64.0f;
        float3 F_ms = 1.0f / (f3(1.0f) - f_ms);
        P.multiscattering_out[gy * P.ms_w + gx] = f4(L_2nd_order * F_ms, 1.0f);
    }
}

#ifndef SKY_COMPOSITE_TU
// K2, production arithmetic: one block per texel as above, but LANES lanes per sphere direction (64 x LANES threads), each folding a
// chunk of the direction's 30 steps (ComputeScatteredLuminanceCooperative); the 64 directions are then added in the reference's
// pairwise order like above.
template <int LANES>
__global__ void __launch_bounds__(64 * LANES, (SKY_LUT_COOP_THREADS / (64 * LANES)) > 0 ? SKY_LUT_COOP_THREADS / (64 * LANES) : 1) k2_multiscattering_cooperative(const __grid_constant__ BakeParams P) {
    const SkyAtmosphereBufferData& u = P.atm.u;
    const int gx = blockIdx.x, gy = blockIdx.y;
    const int local_index = int(threadIdx.x) / LANES, lane = int(threadIdx.x) % LANES;
    const unsigned group_mask = (LANES == 32 ? 0xffffffffu : ((1u << LANES) - 1u)) << ((threadIdx.x & 31u) / LANES * LANES);
    float x_mu_s = float(gx) / float(P.ms_w - 1), x_altitude = float(gy) / float(P.ms_h - 1);
    float altitude = x_altitude * (u.top_radius - u.bottom_radius);
    float mu_s = x_mu_s * 2.0f - 1.0f;
    float3 earth_center = f3(0.0f, -u.bottom_radius, 0.0f);
    float3 start_position = f3(0.0f, altitude, 0.0f);
    float3 sun_direction = f3(0.0f, mu_s, sqrtf(1 - mu_s * mu_s));
    float unit_theta = (0.5f + float(local_index / 8)) / 8.0f;
    float unit_phi = (0.5f + float(local_index % 8)) / 8.0f;
    float cos_theta = 1.0f - 2.0f * unit_theta;
    float sin_theta = sqrtf(clampf(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
    float phi = 2 * kPi * unit_phi;
    float3 view_direction = f3(LUT_COS(phi) * sin_theta, cos_theta, LUT_SIN(phi) * sin_theta);
    float r = altitude + u.bottom_radius;
    float mu = view_direction.y;
    bool intersect_bottom = P.atm.RayIntersectsGround(r, mu);
    float marching_distance = intersect_bottom ? P.atm.DistanceToBottomAtmosphereBoundary(r, mu) : P.atm.DistanceToTopAtmosphereBoundary(r, mu);
    float3 transmittance, L_f;
    float3 luminance = ComputeScatteredLuminanceCooperative<true, LANES>(P.atm, P.transmittance, P.transmittance, 0.5f, earth_center, start_position,
                                                                                  view_direction, sun_direction, marching_distance, u.multiscattering_steps,
                                                                                  lane, group_mask, transmittance, L_f);
    __shared__ float sh[64][6];
    if (lane == 0) {
        if (intersect_bottom) {
            float3 ground_position = start_position + view_direction * marching_distance;
            luminance += transmittance * ComputeGroundLuminance<true>(P.atm, P.transmittance, earth_center, ground_position, sun_direction);
        }
        float* s = sh[local_index];
        s[0] = luminance.x; s[1] = luminance.y; s[2] = luminance.z; s[3] = L_f.x; s[4] = L_f.y; s[5] = L_f.z;
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    float v[6];
#pragma unroll
    for (int k = 0; k < 6; ++k) v[k] = sh[threadIdx.x][k] + sh[threadIdx.x + 32][k];
#pragma unroll
    for (int stride = 16; stride >= 1; stride >>= 1)
#pragma unroll
        for (int k = 0; k < 6; ++k) v[k] += __shfl_down_sync(0xffffffffu, v[k], stride);
    if (threadIdx.x == 0) {
        float3 L_2nd_order = f3(v[0], v[1], v[2]) / 64.0f;
        float3 f_ms = f3(v[3], v[4], v[5]) / 64.0f;
        float3 F_ms = 1.0f / (f3(1.0f) - f_ms);
        P.multiscattering_out[gy * P.ms_w + gx] = f4(L_2nd_order * F_ms, 1.0f);
    }
}
#endif

// RGBA16F copies of the two bake LUTs + the density table of K6's march (context.h): texel i of the table is the altitude
// i / (nd - 1) x (top - bottom) and holds (rayleigh density, mie density, ozone density, 0) -- GetScattering / GetExtinction's three
// altitude profiles (Atmosphere.glsl:119-132,156-159), with the same clamps
__global__ void __launch_bounds__(256) k_luts_to_half(const float4* __restrict__ a, half4* __restrict__ ah, int na, const float4* __restrict__ b,
                                                      half4* __restrict__ bh, int nb, half4* __restrict__ density, int nd, SkyAtmosphereBufferData u) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < na) ah[i] = to_half4(a[i]);
    else if (i - na < nb) {   // (K6's march takes the multiscattering mask from here: one multiplication per texel instead of three per step)
        const float4 v = b[i - na];
        bh[i - na] = to_half4(f4(v.x * u.multiscattering_mask, v.y * u.multiscattering_mask, v.z * u.multiscattering_mask, v.w));
    } else if (i - na - nb < nd) {
        const int k = i - na - nb;
        const float altitude = float(k) / float(nd - 1) * (u.top_radius - u.bottom_radius);
        const float dR = clampf(LUT_EXP(-altitude * u.inv_rayleigh_exponential_distribution), 0.0f, 1.0f);
        const float dM = clampf(LUT_EXP(-altitude * u.inv_mie_exponential_distribution), 0.0f, 1.0f);
        const float dO = fmaxf(0.0f, altitude < u.ozone_center_altitude ? 1.0f + (altitude - u.ozone_center_altitude) * u.inv_ozone_width
                                                                        : 1.0f - (altitude - u.ozone_center_altitude) * u.inv_ozone_width);
        density[k] = to_half4(f4(dR, dM, dO, 0.0f));
    }
}

// ------------------------------------------------------------------------------------------- K3 / K4 / K5 / K6
// Inputs of K6's object branch (AtmosphereRenderer.glsl:284-343,404-410): G-buffer (sky_set_gbuffer), IBL chain
// (ibl.cu), cloud shadow map (K12) and the mesh shadow map (an input, all 1.0 until written)
struct ObjectParams {
    const uchar4* albedo;        // GL_RGBA8
    const short4* normal;        // GL_RGBA16_SNORM
    const ushort4* orm;          // GL_RGBA16
    const ushort2* env_brdf_lut; // GL_RG16 [S][S]
    int env_brdf_size;
    CubeChainView prefiltered;   // SKY_IBL_ROUGHNESS_COUNT levels
    const float4* Llm;           // [9]
    const float2* cloud_shadow_map;  // shadow_maps_[2]: [512][512] (depth, transmittance), zero until the first shadow pass like the reference's
    int cloud_shadow_size;
    ScatterExtras mesh;          // shadow_size == 0: no mesh shadow map allocated (lit)
};

struct RenderParams {
    AtmosphereModel atm;
    SkyAtmosphereRenderBufferData r;  // AtmosphereRenderer.glsl:25-50
    SkyLutConfig cfg;
    LutView transmittance, multiscattering, sky_lum, sky_trans, ap_lum, ap_trans;
    cudaTextureObject_t density_tex;  // 1-D altitude -> (rayleigh, mie, ozone) density table, kDensityLutSize texels (context.h)
    FroxelView froxel;  // p == nullptr: no cloud shadow froxel yet (visibility 1)
    ScatterExtras extras;
    const uchar4* star_map;      // GL_SRGB8 codes (RGBX), nullptr: no star term
    const float* srgb_decode;    // 256 entries
    int star_w, star_h;
    const uint16_t* blue_noise;
    float4 *sky_lum_out, *sky_trans_out, *ap_lum_out, *ap_trans_out;
    half4 *sky_lum_h_out, *sky_trans_h_out, *ap_lum_h_out, *ap_trans_h_out;   // RGBA16F copies for K6's texture fetches (context.h)
    half4* env_out;
    const float* depth;
    half4* hdr;
    int width, height;
    int band_rows, band_index, band_count;  // sky_set_output_bands: blockIdx.y counts OWNED rows (band_count <= 1: all rows)
    ObjectParams object;  // K6 object branch (template flag OBJECT)
};

// AtmosphereRenderer.glsl:56-72
SKY_D float3 ComputeRaymarchingStartPositionAndChangeDistance(const RenderParams& P, float3 view_direction, float& marching_distance) {
    float3 start_position = f3(P.r.camera_position);
    bool in_space = P.r.camera_earth_center_distance > P.atm.u.top_radius;
    if (in_space) {
        float r = P.r.camera_earth_center_distance;
        float mu = dot(view_direction, f3(P.r.up_direction));
        float near_distance;
        if (P.atm.FromSpaceIntersectTopAtmosphereBoundary(r, mu, near_distance)) {
            start_position += near_distance * view_direction;
            marching_distance -= near_distance;
        } else {
            marching_distance = 0;
        }
    }
    return start_position;
}
// AtmosphereRenderer.glsl:74-78
SKY_D float GetHorizonDownAngleFromR(const RenderParams& P, float r) {
    float tangent_point_distance = sqrtf(r * r - P.atm.u.bottom_radius * P.atm.u.bottom_radius);
    return LUT_ACOS(tangent_point_distance / r);
}
// AtmosphereRenderer.glsl:113-132.  acos/sqrt arguments are clamped into their domains: GLSL leaves
// them undefined a few ulp outside, which rounding in dot()/normalize() does produce.
SKY_D float2 GetSkyViewTextureUvFromCosLatLon(const RenderParams& P, float r, float cos_lat, float cos_lon) {
    float horizon_down_angle = GetHorizonDownAngleFromR(P, r);
    float horizon_up_angle = kPi - horizon_down_angle;
    float lat = LUT_ACOS(clampf(cos_lat, -1.0f, 1.0f));
    float x_cos_lat;
    if (lat < horizon_up_angle) {
        float coord = lat / horizon_up_angle;
        coord = sqrtf(fmaxf(1 - coord, 0.0f));
        x_cos_lat = 0.5f - 0.5f * coord;
    } else {
        float coord = (lat - horizon_up_angle) / horizon_down_angle;
        coord = sqrtf(fmaxf(coord, 0.0f));
        x_cos_lat = coord * 0.5f + 0.5f;
    }
    float x_cos_lon = sqrtf(fmaxf(0.5f - 0.5f * cos_lon, 0.0f));
    float w = float(P.cfg.sky_view_width), h = float(P.cfg.sky_view_height);
    return f2(0.5f / w + x_cos_lon * (1.0f - 1.0f / w), 0.5f / h + x_cos_lat * (1.0f - 1.0f / h));
}
// AtmosphereRenderer.glsl:134-145
SKY_D void GetCosLatLonFromViewDirection(const RenderParams& P, float3 view_direction, float& cos_lat, float& cos_lon) {
    float3 up = f3(P.r.up_direction);
    cos_lat = dot(up, view_direction);
    float3 lon_direction = view_direction - up * cos_lat;
    float lon_direction_length2 = dot(lon_direction, lon_direction);
    if (lon_direction_length2 == 0) {
        cos_lon = 1;
    } else {
        lon_direction *= (1.0f / sqrtf(lon_direction_length2));
        cos_lon = dot(lon_direction, f3(P.r.front_direction));
    }
}
SKY_D float DitherStart(const RenderParams& P, int enable, int x, int y) {
    if (enable) return float(__ldg(P.blue_noise + (y & 0x3f) * 64 + (x & 0x3f))) / 65535.0f;
    return 0.5f;
}

// K3 -- AtmosphereRenderer.glsl:153-186 (+ :81-111)
// LANES > 1: the production arithmetic -- LANES consecutive lanes share the texel's march (lane `lane` of the group `group_mask`)
template <bool EXTRA, int LANES = 1>
SKY_D void k3_sky_view_texel(const RenderParams& P, int x, int y, int lane = 0, unsigned group_mask = 0) {
    const int W = P.cfg.sky_view_width, H = P.cfg.sky_view_height;
    if (x >= W || y >= H) return;
    float r = P.r.camera_earth_center_distance;
    // GetCosLatLonFromSkyViewTextureIndex
    float x_cos_lon = float(x) / float(W - 1), x_cos_lat = float(y) / float(H - 1);
    float horizon_down_angle = GetHorizonDownAngleFromR(P, r);
    float horizon_up_angle = kPi - horizon_down_angle;
    float lat;
    if (x_cos_lat < 0.5f) {
        float coord = 1.0f - 2.0f * x_cos_lat;
        coord = 1.0f - coord * coord;
        lat = horizon_up_angle * coord;
    } else {
        float coord = x_cos_lat * 2.0f - 1.0f;
        coord *= coord;
        lat = horizon_up_angle + horizon_down_angle * coord;
    }
    float cos_lat = LUT_COS(lat);
    float cos_lon = -(x_cos_lon * x_cos_lon * 2.0f - 1.0f);
    // GetViewDirectionFromCosLatLon
    float sin_lat = clampf(sqrtf(1 - cos_lat * cos_lat), 0.0f, 1.0f);
    float sin_lon = clampf(sqrtf(1 - cos_lon * cos_lon), 0.0f, 1.0f);
    float3 view_direction = f3(P.r.up_direction) * cos_lat + f3(P.r.front_direction) * (sin_lat * cos_lon) +
                            f3(P.r.right_direction) * (sin_lat * sin_lon);
    float mu = cos_lat;
    bool intersect_bottom = P.atm.RayIntersectsGround(r, mu);
    float marching_distance = intersect_bottom ? P.atm.DistanceToBottomAtmosphereBoundary(r, mu) : P.atm.DistanceToTopAtmosphereBoundary(r, mu);
    float3 start_position = ComputeRaymarchingStartPositionAndChangeDistance(P, view_direction, marching_distance);
    float3 transmittance = f3(1.0f), luminance = f3(0.0f), unused;
    if (marching_distance > 0) {
        float start_i = DitherStart(P, P.cfg.sky_view_dither, x, y);
#ifndef SKY_COMPOSITE_TU
        if constexpr (LANES > 1)
            luminance = ComputeScatteredLuminanceCooperative<false, LANES>(P.atm, P.transmittance, P.multiscattering, start_i, f3(P.r.earth_center),
                                                                                    start_position, view_direction, f3(P.r.sun_direction), marching_distance,
                                                                                    P.r.sky_view_lut_steps, lane, group_mask, transmittance, unused);
        else
#endif
        luminance = ComputeScatteredLuminance<false, false, EXTRA>(P.atm, P.transmittance, P.multiscattering, start_i, f3(P.r.earth_center), start_position,
                                                                   view_direction, f3(P.r.sun_direction), marching_distance, P.r.sky_view_lut_steps,
                                                                   transmittance, unused, &P.extras);
    }
    if (lane != 0) return;
    P.sky_lum_out[y * W + x] = f4(luminance, 0.0f);
    P.sky_trans_out[y * W + x] = f4(transmittance, 0.0f);
    P.sky_lum_h_out[y * W + x] = to_half4(f4(luminance, 0.0f));
    P.sky_trans_h_out[y * W + x] = to_half4(f4(transmittance, 0.0f));
}

// K4 -- AtmosphereRenderer.glsl:191-243
template <bool EXTRA, int LANES = 1>
SKY_D void k4_aerial_perspective_froxel(const RenderParams& P, int x, int y, int z, int lane = 0, unsigned group_mask = 0) {
    const int W = P.ap_lum.w, H = P.ap_lum.h, D = P.ap_lum.d;
    if (x >= W || y >= H || z >= D) return;
    float3 uvw = f3(float(x) / float(W - 1), float(y) / float(H - 1), float(z) / float(D - 1));
    float3 position = projective_mul(P.r.inv_view_projection, f3(uvw.x * 2.0f - 1.0f, uvw.y * 2.0f - 1.0f, 0.0f));
    float3 view_direction = normalize(position - f3(P.r.camera_position));
    float marching_distance = uvw.z * uvw.z * P.r.aerial_perspective_lut_max_distance;

    float r = P.r.camera_earth_center_distance;
    float mu = dot(view_direction, f3(P.r.up_direction));
    bool intersect_bottom = P.atm.RayIntersectsGround(r, mu);
    float max_marching_distance = intersect_bottom ? marching_distance : P.atm.DistanceToTopAtmosphereBoundary(r, mu);
    float3 start_position = ComputeRaymarchingStartPositionAndChangeDistance(P, view_direction, max_marching_distance);
    marching_distance = fminf(marching_distance, max_marching_distance);
    float3 transmittance = f3(1.0f), luminance = f3(0.0f), unused;
    if (marching_distance > 0) {
        float start_i = DitherStart(P, P.cfg.aerial_perspective_dither, x, y);
#ifndef SKY_COMPOSITE_TU
        if constexpr (LANES > 1)
            luminance = ComputeScatteredLuminanceCooperative<false, LANES>(P.atm, P.transmittance, P.multiscattering, start_i, f3(P.r.earth_center),
                                                                                    start_position, view_direction, f3(P.r.sun_direction), marching_distance,
                                                                                    P.r.aerial_perspective_lut_steps, lane, group_mask, transmittance, unused);
        else
#endif
        luminance = ComputeScatteredLuminance<false, false, EXTRA>(P.atm, P.transmittance, P.multiscattering, start_i, f3(P.r.earth_center), start_position,
                                                                   view_direction, f3(P.r.sun_direction), marching_distance,
                                                                   P.r.aerial_perspective_lut_steps, transmittance, unused, &P.extras);
    }
    if (lane != 0) return;
    size_t o = (size_t(z) * H + y) * W + x;
    P.ap_lum_out[o] = f4(luminance, 0.0f);
    P.ap_trans_out[o] = f4(transmittance, 0.0f);
    P.ap_lum_h_out[o] = to_half4(f4(luminance, 0.0f));
    P.ap_trans_h_out[o] = to_half4(f4(transmittance, 0.0f));
}

// K3 and K4 are independent (both read only the bake LUTs) and each is a few hundred 64-thread blocks of long dependent
// marches -- far too little to fill 148 SMs alone -- so they share ONE launch: the first n3 blocks are K3's texel rows, the
// rest K4's froxel rows.  Same per-thread code, same results; the LUT phase of a frame costs max(K3, K4) instead of K3 + K4.
template <bool EXTRA>
__global__ void __launch_bounds__(64) k34_sky_view_and_aerial_perspective(const __grid_constant__ RenderParams P, int n3, int g3x, int g4x) {
    if (int(blockIdx.x) < n3) {
        k3_sky_view_texel<EXTRA>(P, int(blockIdx.x % g3x) * 64 + int(threadIdx.x), int(blockIdx.x / g3x));
    } else {
        const int b = int(blockIdx.x) - n3, W = P.ap_lum.w;
        k4_aerial_perspective_froxel<EXTRA>(P, int(threadIdx.x) % W, (b % g4x) * (64 / W) + int(threadIdx.x) / W, b / g4x);
    }
}

#ifndef SKY_COMPOSITE_TU
// K3 + K4, production arithmetic: the same shared launch, 256-thread blocks = 256 / LANES marches x LANES lanes.  K3 blocks take a run of
// texels of a sky-view row, K4 blocks a run of froxels of a row of a slice.
template <int LANES>
__global__ void __launch_bounds__(256, SKY_LUT_COOP_THREADS / 256) k34_cooperative(const __grid_constant__ RenderParams P, int n3, int g3x, int g4x) {
    constexpr int kMarches = 256 / LANES;
    const int march = int(threadIdx.x) / LANES, lane = int(threadIdx.x) % LANES;
    const unsigned group_mask = (LANES == 32 ? 0xffffffffu : ((1u << LANES) - 1u)) << ((threadIdx.x & 31u) / LANES * LANES);
    if (int(blockIdx.x) < n3) {
        k3_sky_view_texel<false, LANES>(P, int(blockIdx.x % g3x) * kMarches + march, int(blockIdx.x / g3x), lane, group_mask);
    } else {
        const int b = int(blockIdx.x) - n3, H = P.ap_lum.h;
        k4_aerial_perspective_froxel<false, LANES>(P, (b % g4x) * kMarches + march, (b / g4x) % H, b / (g4x * H), lane, group_mask);
    }
}
#endif

// shaders/Base/Common.glsl:13-30
SKY_D float3 ConvertCubUvToDir(int index, float u, float v) {
    float uc = 2.0f * u - 1.0f, vc = 2.0f * v - 1.0f;
    float3 dir = f3(0.0f);
    switch (index) {
        case 0: dir = f3(1.0f, vc, -uc); break;
        case 1: dir = f3(-1.0f, vc, uc); break;
        case 2: dir = f3(uc, 1.0f, -vc); break;
        case 3: dir = f3(uc, -1.0f, vc); break;
        case 4: dir = f3(uc, vc, 1.0f); break;
        case 5: dir = f3(-uc, vc, -1.0f); break;
    }
    return normalize(dir);
}

// K5 -- AtmosphereRenderer.glsl:253-273
__global__ void __launch_bounds__(128) k5_environment(const __grid_constant__ RenderParams P) {
    const int S = P.cfg.environment_size;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, index = blockIdx.z;
    if (x >= S) return;
    float fu = (float(x) + 0.5f) / float(S), fv = (float(y) + 0.5f) / float(S);
    fv = 1.0f - fv;
    float3 view_direction = ConvertCubUvToDir(index, fu, fv);
    float cos_lat, cos_lon;
    GetCosLatLonFromViewDirection(P, view_direction, cos_lat, cos_lon);
    float r = P.r.camera_earth_center_distance;
    float2 uv = GetSkyViewTextureUvFromCosLatLon(P, r, cos_lat, cos_lon);
    float3 luminance = xyz(sample_lut2d(P.sky_lum, uv.x, uv.y));
    float3 transmittance = xyz(sample_lut2d(P.sky_trans, uv.x, uv.y));
    float mu = cos_lat;
    if (P.atm.RayIntersectsGround(r, mu)) {
        float marching_distance = P.atm.DistanceToBottomAtmosphereBoundary(r, mu);
        float3 ground_position = f3(P.r.camera_position) + view_direction * marching_distance;
        luminance += ComputeGroundLuminance<false>(P.atm, P.transmittance, f3(P.r.earth_center), ground_position, f3(P.r.sun_direction)) * transmittance;
    }
    P.env_out[(size_t(index) * S + y) * S + x] = to_half4(f4(luminance, 0.0f));
}

// ---- object branch of K6 ----------------------------------------------------------------------------------------------
// texture(sampler2D, uv), LinearNoMipmapClampToEdge, over a normalised-integer G-buffer target (exact fp32 weights)
template <class T4, class Decode>
SKY_D float3 gbuffer_texture(const T4* img, int w, int h, float2 uv, Decode decode) {
    float x = uv.x * float(w) - 0.5f, y = uv.y * float(h) - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    int i0 = clampi(int(fx), 0, w - 1), i1 = clampi(int(fx) + 1, 0, w - 1), j0 = clampi(int(fy), 0, h - 1), j1 = clampi(int(fy) + 1, 0, h - 1);
    auto T = [&](int i, int j) { T4 c = __ldg(img + size_t(j) * w + i); return f3(decode(c.x), decode(c.y), decode(c.z)); };
    return (1.0f - a) * (1.0f - b) * T(i0, j0) + a * (1.0f - b) * T(i1, j0) + (1.0f - a) * b * T(i0, j1) + a * b * T(i1, j1);
}
// shaders/Base/BRDF.glsl:26-49
SKY_D float Pow5(float x) { float x2 = x * x; return x2 * x2 * x; }
SKY_D float3 F_Schlick(float HdotV, float3 F0) { return F0 + (f3(1.0f) - F0) * Pow5(1.0f - HdotV); }
SKY_D float D_GGX(float a, float NdotH) {
    float a2 = a * a;
    float d = (NdotH * a2 - NdotH) * NdotH + 1.0f;
    return a2 / (kPi * d * d);
}
SKY_D float Vis_SmithJointApprox(float a, float NdotV, float NdotL) {
    float Vis_SmithV = NdotL * (NdotV * (1.0f - a) + a);
    float Vis_SmithL = NdotV * (NdotL * (1.0f - a) + a);
    return 0.5f / fmaxf(Vis_SmithV + Vis_SmithL, 1e-9f);
}
SKY_D float3 mix3v(float3 a, float3 b, float3 t) { return f3(mixf(a.x, b.x, t.x), mixf(a.y, b.y, t.y), mixf(a.z, b.z, t.z)); }
// shaders/Base/BRDF.glsl:81-106
SKY_D float3 GetSHIrradiance(float3 N, const float4* Llm) {
    const float c1 = 0.429043f, c2 = 0.511664f, c3 = 0.743125f, c4 = 0.886227f, c5 = 0.247708f;
    float3 L00 = xyz(__ldg(Llm + 0)), L1_1 = xyz(__ldg(Llm + 1)), L10 = xyz(__ldg(Llm + 2)), L11 = xyz(__ldg(Llm + 3)), L2_2 = xyz(__ldg(Llm + 4)),
           L2_1 = xyz(__ldg(Llm + 5)), L20 = xyz(__ldg(Llm + 6)), L21 = xyz(__ldg(Llm + 7)), L22 = xyz(__ldg(Llm + 8));
    float x = N.x, y = N.y, z = N.z;
    return c1 * (x * x - y * y) * L22 + c3 * (z * z) * L20 + c4 * L00 - c5 * L20
        + 2.0f * c1 * (x * y * L2_2 + x * z * L21 + y * z * L2_1)
        + 2.0f * c2 * (x * L11 + y * L1_1 + z * L10);
}
// texture(sampler2DShadow, vec3(u, v, depth)) through Samplers::GetShadowMapSampler: bilinear blend of four LEQUAL comparisons, border 1
SKY_D float ShadowCompare(const ScatterExtras& e, float u, float v, float depth) {
    const float S = float(e.shadow_size);
    float x = u * S - 0.5f, y = v * S - 0.5f;
    float fx = floorf(x), fy = floorf(y);
    float a = x - fx, b = y - fy;
    auto cmp = [&](float i, float j) {
        bool inside = i >= 0.0f && j >= 0.0f && i < S && j < S;
        float texel = inside ? __ldg(e.shadow_map + size_t(int(j)) * e.shadow_size + int(i)) : 1.0f;
        return depth <= texel ? 1.0f : 0.0f;
    };
    return (1.0f - a) * (1.0f - b) * cmp(fx, fy) + a * (1.0f - b) * cmp(fx + 1.0f, fy) + (1.0f - a) * b * cmp(fx, fy + 1.0f) +
           a * b * cmp(fx + 1.0f, fy + 1.0f);
}
// Shadow.glsl:13-99 (PCSS_ENABLE 1): per-pixel Poisson disc, blocker search through the NEAREST / CLAMP_TO_EDGE view of the mesh
// shadow map (AtmosphereRenderer.cpp:196,213), 25-tap percentage-closer filter.  The strict object takes sin / cos from
// sky_detmath.h like the oracle; the production object the hardware's -- except for the ONE sin per pixel behind the disc's
// random rotation, fract(sin(x) * 43758.5), which no two sin implementations agree on: it is sky_detmath.h's in both objects,
// so that the production penumbra is the oracle's penumbra and not just statistically alike
SKY_D float PCSS(const RenderParams& P, const ScatterExtras& e, float3 position) {
    constexpr int NUM_SAMPLES = 25, NUM_RINGS = 3;
    const float PI2 = kPi * 2.0f;
    const float* m = e.light_view_projection;
    float X = m[0] * position.x + m[4] * position.y + m[8] * position.z + m[12] * 1.0f;
    float Y = m[1] * position.x + m[5] * position.y + m[9] * position.z + m[13] * 1.0f;
    float Z = m[2] * position.x + m[6] * position.y + m[10] * position.z + m[14] * 1.0f;
    float Wc = m[3] * position.x + m[7] * position.y + m[11] * position.z + m[15] * 1.0f;
    float cx = X / Wc * 0.5f + 0.5f, cy = Y / Wc * 0.5f + 0.5f, cz = Z / Wc * 0.5f + 0.5f;
    if (cz >= 1.0f) return 1.0f;
#ifdef SKY_STRICT_TU
#define PCSS_SIN sky_det_sinf
#define PCSS_COS sky_det_cosf
#else
#define PCSS_SIN sinf
#define PCSS_COS cosf
#endif
    float2 poissonDisk[NUM_SAMPLES];
    {
        float ANGLE_STEP = PI2 * float(NUM_RINGS) / float(NUM_SAMPLES);
        float INV_NUM_SAMPLES = 1.0f / float(NUM_SAMPLES);
        const float a = 12.9898f, b = 78.233f, c = 43758.5453f;
        float dt = cx * a + cy * b, sn = dt - kPi * floorf(dt / kPi);
        float rnd = fractf(sky_det_sinf(sn) * c);  // both objects: see above
        float angle = rnd * PI2;
        float radius = INV_NUM_SAMPLES;
        float radiusStep = radius;
#pragma unroll
        for (int i = 0; i < NUM_SAMPLES; i++) {
            float pr = powf(radius, 0.75f);
            poissonDisk[i] = f2(PCSS_COS(angle) * pr, PCSS_SIN(angle) * pr);
            radius += radiusStep;
            angle += ANGLE_STEP;
        }
    }
    const int S = e.shadow_size;
    float kernelSizeApproximate = P.r.blocker_kernel_size_k * cz;
    float sum = 0.0f, cnt = 0.0f;
#pragma unroll
    for (int i = 0; i < NUM_SAMPLES; ++i) {  // FindBlocker
        float sx = poissonDisk[i].x * kernelSizeApproximate + cx, sy = poissonDisk[i].y * kernelSizeApproximate + cy;
        int tx = clampi(int(floorf(sx * float(S))), 0, S - 1), ty = clampi(int(floorf(sy * float(S))), 0, S - 1);
        float shadow_depth = __ldg(e.shadow_map + size_t(ty) * S + tx);
        if (cz - shadow_depth > 0.0f) {
            cnt += 1.0f;
            sum += shadow_depth;
        }
    }
    float avgblockerDepth = sum / fmaxf(cnt, 1e-5f);
    float distanceToFragment = cz - avgblockerDepth;
    float penumbraSize = P.r.pcss_size_k * distanceToFragment;
    float vis = 0.0f;
#pragma unroll
    for (int i = 0; i < NUM_SAMPLES; ++i)  // Filtering
        vis += ShadowCompare(e, poissonDisk[i].x * penumbraSize + cx, poissonDisk[i].y * penumbraSize + cy, cz);
    return vis / float(NUM_SAMPLES);
}
// AtmosphereRenderer.glsl:333-343; SampleCloudShadowTransmittance (VolumetricCloudShadowInterface.glsl:4-8)
// through the cloud shadow sampler: LINEAR, CLAMP_TO_BORDER (1e10, 1) (VolumetricCloud.cpp:106-112)
template <bool PCSS_ON>  // PCSS_ENABLE is a template flag like the host's #define: as a run-time branch its 25-entry sample table cost
                         // the plain object kernel 0.5 ms at 4K (1.14 -> 1.64 ms, measured)
SKY_D float SampleVisibilityFromShadowMap(const RenderParams& P, float3 position) {
    const ObjectParams& O = P.object;
    float visibility = 1.0f;
    if (O.mesh.shadow_size > 0) visibility = PCSS_ON ? PCSS(P, O.mesh, position) : GetVisibilityFromShadowMap(O.mesh, position);
    if (O.cloud_shadow_map) {
        float3 light_ndc = projective_mul(P.r.uCloudShadowMapMat, position);
        const int S = O.cloud_shadow_size;
        float x = (light_ndc.x * 0.5f + 0.5f) * float(S) - 0.5f, y = (light_ndc.y * 0.5f + 0.5f) * float(S) - 0.5f;
        float fx = floorf(x), fy = floorf(y);
        float a = x - fx, b = y - fy;
        auto T = [&](float i, float j) {
            bool inside = i >= 0.0f && j >= 0.0f && i < float(S) && j < float(S);
            return inside ? __ldg(O.cloud_shadow_map + size_t(int(j)) * S + int(i)) : f2(1e10f, 1.0f);
        };
        float2 dt = (1.0f - a) * (1.0f - b) * T(fx, fy) + a * (1.0f - b) * T(fx + 1.0f, fy) + (1.0f - a) * b * T(fx, fy + 1.0f) + a * b * T(fx + 1.0f, fy + 1.0f);
        const float kInvTransitionDepth = 1.0f / 0.5f;
        visibility = fminf(visibility, mixf(dt.y, 1.0f, clampf((dt.x - light_ndc.z) * kInvTransitionDepth, 0.0f, 1.0f)));
    }
    return visibility;
}
// AtmosphereRenderer.glsl:284-324 (LoadMeterialData / BRDF / GetAmbient: shaders/Base/BRDF.glsl:10-23,69-78,108-130)
SKY_D float3 ComputeObjectLuminance(const RenderParams& P, float3 position, float3 view_direction, float shadow_visibility, float2 vTexCoord) {
    const ObjectParams& O = P.object;
    const float3 earth_center = f3(P.r.earth_center), sun_direction = f3(P.r.sun_direction);
    float r = length(position - earth_center);
    float3 object_up_direction = normalize(position - earth_center);
    float mu_s = dot(sun_direction, object_up_direction);
    float3 sun_visibility;
    if (r > P.atm.u.top_radius) {
        float near_distance;
        if (P.atm.FromSpaceIntersectTopAtmosphereBoundary(r, mu_s, near_distance)) {
            position += near_distance * sun_direction;
            r = length(position - earth_center);
            object_up_direction = normalize(position - earth_center);
            mu_s = dot(sun_direction, object_up_direction);
            sun_visibility = P.atm.GetSunVisibility(P.transmittance, r, mu_s);
        } else {
            sun_visibility = f3(1.0f);
        }
    } else {
        sun_visibility = P.atm.GetSunVisibility(P.transmittance, r, mu_s);
    }
    float3 solar_illuminance_at_object = P.atm.solar_illuminance() * sun_visibility;

    float3 albedo = gbuffer_texture(O.albedo, P.width, P.height, vTexCoord, [](unsigned char c) { return float(c) / 255.0f; });
    float3 normal = gbuffer_texture(O.normal, P.width, P.height, vTexCoord, [](short c) { return fmaxf(float(c) / 32767.0f, -1.0f); });
    float3 orm = gbuffer_texture(O.orm, P.width, P.height, vTexCoord, [](unsigned short c) { return float(c) / 65535.0f; });
    float metallic = orm.z, roughness = orm.y;
    float3 F0 = f3(0.04f) * (1.0f - metallic) + albedo * metallic;
    float3 mdiffuse = albedo - albedo * metallic;

    float3 N = normal, L = sun_direction, V = -view_direction;
    float3 H = normalize(L + V);
    float NdotL = clampf(dot(N, L), 0.0f, 1.0f), NdotV = clampf(dot(N, V), 0.0f, 1.0f);
    float NdotH = clampf(dot(N, H), 0.0f, 1.0f), HdotV = clampf(dot(H, V), 0.0f, 1.0f);
    float3 brdf;
    {
        float a = roughness * roughness;
        float D = fminf(D_GGX(a, NdotH), 1e9f);
        float Vis = Vis_SmithJointApprox(a, NdotV, NdotL);
        float3 F = F_Schlick(HdotV, F0);
        brdf = mix3v(kInvPi * mdiffuse, f3(D * Vis), F);
    }
    float3 direct_lumiance = brdf * solar_illuminance_at_object * (NdotL * shadow_visibility);
    float3 ambient_lumiance;
    {
        const float roughness_lod_max = float(SKY_IBL_ROUGHNESS_COUNT - 1);
        float3 F = F_Schlick(NdotV * 0.8f + 0.2f, F0);
        float3 approx_irradiance_over_pi = GetSHIrradiance(N, O.Llm) * kInvPi;
        float3 diffuse = (mdiffuse - mdiffuse * F) * approx_irradiance_over_pi;
        float3 R = 2.0f * NdotV * N - V;
        float3 prefiltered_radiance = xyz(TextureCubeLod(O.prefiltered, R, roughness * roughness_lod_max));
        // texture(env_brdf_lut, vec2(NdotV, roughness)).rg, LinearNoMipmapClampToEdge over GL_RG16
        const int S = O.env_brdf_size;
        float x = NdotV * float(S) - 0.5f, y = roughness * float(S) - 0.5f;
        float fx = floorf(x), fy = floorf(y);
        float a = x - fx, b = y - fy;
        int i0 = clampi(int(fx), 0, S - 1), i1 = clampi(int(fx) + 1, 0, S - 1), j0 = clampi(int(fy), 0, S - 1), j1 = clampi(int(fy) + 1, 0, S - 1);
        auto T = [&](int i, int j) { ushort2 c = __ldg(O.env_brdf_lut + size_t(j) * S + i); return f2(float(c.x) / 65535.0f, float(c.y) / 65535.0f); };
        float2 lut = (1.0f - a) * (1.0f - b) * T(i0, j0) + a * (1.0f - b) * T(i1, j0) + (1.0f - a) * b * T(i0, j1) + a * b * T(i1, j1);
        float3 specular = prefiltered_radiance * (F0 * lut.x + f3(lut.y));
        ambient_lumiance = diffuse + specular;
    }
    float ambient_fade = clampf(10.0f - 0.1f * length(position - f3(P.r.camera_position)), 0.0f, 1.0f);
    return direct_lumiance + ambient_lumiance * ambient_fade;
}

// K6 -- AtmosphereRenderer.glsl:345-432: sky-view LUT / aerial-perspective LUT / per-pixel raymarch,
// x cloud-shadow froxel, + sun disc with limb darkening, + the star map on sky pixels (:427-429).  Object pixels (depth != 1) are
// shaded by ComputeObjectLuminance when a G-buffer is bound (template flag OBJECT); without one they carry the in-scatter alone,
// which is what the reference's program computes on a cleared (all-zero) G-buffer.  Alpha is 1 everywhere (:431).
// HBM-bound: 4 B depth in + 8 B hdr out per pixel; LUTs and froxels are L2-resident.
#ifndef SKY_K6_OCC
#define SKY_K6_OCC 3   // resident 256-thread blocks per SM of the plain composite (80 registers)
#endif
#ifndef SKY_K6_LUT_OCC
#define SKY_K6_LUT_OCC 6
#endif
// LUTONLY: the launch's configuration has both USE_SKY_VIEW_LUT and USE_AERIAL_PERSPECTIVE_LUT (scenes c1 / c2): no pixel marches, so the
// march is compiled out and the kernel -- a latency-bound chain of depth load, LUT and froxel fetches -- runs at twice the occupancy
template <bool EXTRA, bool OBJECT, bool PCSS_ON = false, bool LUTONLY = false, bool GREY = false>
__global__ void __launch_bounds__(256, PCSS_ON ? 1 : OBJECT ? 2 : LUTONLY ? SKY_K6_LUT_OCC : SKY_K6_OCC) k6_composite(const __grid_constant__ RenderParams P) {
    int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y;
    if (P.band_count > 1) py = ((py / P.band_rows) * P.band_count + P.band_index) * P.band_rows + py % P.band_rows;
    if (px >= P.width || py >= P.height) return;
    float2 vTexCoord = f2((float(px) + 0.5f) / float(P.width), (float(py) + 0.5f) / float(P.height));
    float depth = __ldg(P.depth + size_t(py) * P.width + px);
    float3 camera_position = f3(P.r.camera_position);
    float3 fragment_position = projective_mul(P.r.inv_view_projection, f3(vTexCoord.x * 2.0f - 1.0f, vTexCoord.y * 2.0f - 1.0f, depth * 2.0f - 1.0f));
    float3 view_direction = normalize(fragment_position - camera_position);
    float r = P.r.camera_earth_center_distance;
    float mu = dot(view_direction, f3(P.r.up_direction));
    float marching_distance = P.atm.RayIntersectsGround(r, mu) ? P.atm.DistanceToBottomAtmosphereBoundary(r, mu) : P.atm.DistanceToTopAtmosphereBoundary(r, mu);
    bool intersect_object = false;
    if (depth != 1.0f) {
        float object_distance = length(fragment_position - camera_position);
        intersect_object = true;
        marching_distance = fminf(marching_distance, object_distance);
    }
    float3 start_position = ComputeRaymarchingStartPositionAndChangeDistance(P, view_direction, marching_distance);
    float3 transmittance = f3(1.0f), luminance = f3(0.0f), unused;
    float3 sun_direction = f3(P.r.sun_direction);
    if (marching_distance > 0) {
        if (P.cfg.use_sky_view_lut && !intersect_object) {
            float cos_lat, cos_lon;
            GetCosLatLonFromViewDirection(P, view_direction, cos_lat, cos_lon);
            float2 uv = GetSkyViewTextureUvFromCosLatLon(P, r, cos_lat, cos_lon);
            luminance = xyz(sample_lut2d_sel<kCompositeTexLut>(P.sky_lum, uv.x, uv.y));
            transmittance = xyz(sample_lut2d_sel<kCompositeTexLut>(P.sky_trans, uv.x, uv.y));
        } else if (P.cfg.use_aerial_perspective_lut && intersect_object) {
            float3 uvw = aerial_perspective_uvw(vTexCoord, marching_distance, P.r.aerial_perspective_lut_max_distance, P.ap_lum.w, P.ap_lum.h, P.ap_lum.d);
            luminance = xyz(sample_lut3d_sel<kCompositeTexLut>(P.ap_lum, uvw.x, uvw.y, uvw.z));
            transmittance = xyz(sample_lut3d_sel<kCompositeTexLut>(P.ap_trans, uvw.x, uvw.y, uvw.z));
        } else if (!LUTONLY) {
            float start_i = DitherStart(P, P.cfg.raymarching_dither, px, py);
#if defined(SKY_COMPOSITE_TU) && !defined(SKY_STRICT_TU)
            luminance = ComputeScatteredLuminanceFast<EXTRA, GREY>(P.atm, P.transmittance, P.multiscattering, P.density_tex, start_i, f3(P.r.earth_center), start_position, view_direction,
                                                             sun_direction, marching_distance, P.r.raymarching_steps, transmittance, &P.extras);
#else
            luminance = ComputeScatteredLuminance<false, kCompositeTexLut, EXTRA>(P.atm, P.transmittance, P.multiscattering, start_i, f3(P.r.earth_center), start_position,
                                                                                  view_direction, sun_direction, marching_distance, P.r.raymarching_steps, transmittance, unused,
                                                                                  &P.extras);
#endif
        }
    }
    if (P.froxel.p) luminance *= SampleRayScatterVisibilitySel<kCompositeTexLut>(P.froxel, vTexCoord, marching_distance, P.r.uInvShadowFroxelMaxDistance);

    if (intersect_object) {
        if (OBJECT) {  // :404-410
            float shadow_visibility = SampleVisibilityFromShadowMap<PCSS_ON>(P, fragment_position);
            if (EXTRA && P.extras.moon_shadow)
                shadow_visibility *= GetVisibilityFromMoonShadow(f3(P.extras.moon_position) - fragment_position, P.extras.moon_radius, sun_direction, P.atm.u.sun_angular_radius);
            luminance += transmittance * ComputeObjectLuminance(P, fragment_position, view_direction, shadow_visibility, vTexCoord);
        }
    } else if (dot(view_direction, sun_direction) >= cosf(P.atm.u.sun_angular_radius)) {
        const float3 a = f3(0.397f, 0.503f, 0.652f);
        float cos_view_sun = dot(view_direction, sun_direction);
        float sin_view_sun = sqrtf(1.0f - cos_view_sun * cos_view_sun);
        float center_to_edge = clampf(sin_view_sun / sinf(P.atm.u.sun_angular_radius), 0.0f, 1.0f);
        float mu2 = sqrtf(1.0f - center_to_edge * center_to_edge);
        float3 factor = f3(1.0f) - f3(1.0f) * (f3(1.0f) - f3(powf(mu2, a.x), powf(mu2, a.y), powf(mu2, a.z)));
        float3 solar_illuminance_at_eye = P.atm.solar_illuminance() * transmittance;
        luminance += solar_illuminance_at_eye / (kPi * P.atm.u.sun_angular_radius * P.atm.u.sun_angular_radius) * factor;
    } else if (P.star_map) {
        // GetStarLuminance, :326-331 and :427-429: equirectangular look-up, LinearNoMipmapClampToEdge, texels decoded before filtering
        float theta = LUT_ACOS(fminf(fmaxf(view_direction.y, -1.0f), 1.0f));
        float phi = atan2f(view_direction.x, view_direction.z);
        float u = kInvPi * 0.5f * phi + 0.5f, v = 1.0f - theta * kInvPi;
        float x = u * float(P.star_w) - 0.5f, y = v * float(P.star_h) - 0.5f;
        float fx = floorf(x), fy = floorf(y);
        float a = x - fx, b = y - fy;
        int i0 = clampi(int(fx), 0, P.star_w - 1), i1 = clampi(int(fx) + 1, 0, P.star_w - 1);
        int j0 = clampi(int(fy), 0, P.star_h - 1), j1 = clampi(int(fy) + 1, 0, P.star_h - 1);
        auto texel = [&](int i, int j) {
            uchar4 c = __ldg(P.star_map + size_t(j) * P.star_w + i);
            return f3(__ldg(P.srgb_decode + c.x), __ldg(P.srgb_decode + c.y), __ldg(P.srgb_decode + c.z));
        };
        float3 star = (1.0f - a) * (1.0f - b) * texel(i0, j0) + a * (1.0f - b) * texel(i1, j0) + (1.0f - a) * b * texel(i0, j1) + a * b * texel(i1, j1);
        luminance += transmittance * (P.r.star_luminance_scale * star);
    }
    P.hdr[size_t(py) * P.width + px] = to_half4(f4(luminance, 1.0f));   // FragColor = vec4(luminance, 1.0), :431
}

// R16-unorm LINEAR view of the current froxel volume (its slices stacked), for K6's production object.  The volume is double-buffered under
// frame pipelining: two cached views keyed by pointer.  0 when the row pitch does not meet the texture alignment (odd viewports): software path.
cudaTextureObject_t froxel_texture(SkyContext* ctx) {
    const Lut<uint16_t>& f = ctx->shadow_froxel;
    if (!f.p || (size_t(f.w) * sizeof(uint16_t)) % 32 != 0 || size_t(f.h) * f.d > 65536) return 0;
    for (auto& e : ctx->froxel_tex)
        if (e.tex && e.key == f.p && e.w == f.w && e.h == f.h && e.d == f.d) return e.tex;
    SkyContext::FroxelTex* slot = &ctx->froxel_tex[0];
    for (auto& e : ctx->froxel_tex) if (!e.tex) { slot = &e; break; }
    if (slot->tex) {   // both slots hold other volumes (a resize): rebuild both lazily
        for (auto& e : ctx->froxel_tex) { cudaDestroyTextureObject(e.tex); e = SkyContext::FroxelTex{}; }
        slot = &ctx->froxel_tex[0];
    }
    cudaResourceDesc res{};
    res.resType = cudaResourceTypePitch2D;
    res.res.pitch2D.devPtr = f.p;
    res.res.pitch2D.desc = cudaCreateChannelDesc(16, 0, 0, 0, cudaChannelFormatKindUnsigned);
    res.res.pitch2D.width = size_t(f.w);
    res.res.pitch2D.height = size_t(f.h) * size_t(f.d);
    res.res.pitch2D.pitchInBytes = size_t(f.w) * sizeof(uint16_t);
    cudaTextureDesc td{};
    td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp;
    td.filterMode = cudaFilterModeLinear;
    td.readMode = cudaReadModeNormalizedFloat;
    td.normalizedCoords = 1;
    if (cudaCreateTextureObject(&slot->tex, &res, &td, nullptr) != cudaSuccess) { cudaGetLastError(); slot->tex = 0; return 0; }
    slot->key = f.p; slot->w = f.w; slot->h = f.h; slot->d = f.d;
    return slot->tex;
}

RenderParams make_render_params(SkyContext* ctx) {
    RenderParams P{};
    P.atm.u = ctx->atm;
    P.r = ctx->render;
    P.cfg = ctx->lut_cfg;
    P.transmittance = LutView{ctx->transmittance.p, ctx->transmittance.w, ctx->transmittance.h, 1, ctx->transmittance_tex};
    P.multiscattering = LutView{ctx->multiscattering.p, ctx->multiscattering.w, ctx->multiscattering.h, 1, ctx->multiscattering_tex};
    P.sky_lum = LutView{ctx->sky_lum.p, ctx->sky_lum.w, ctx->sky_lum.h, 1, ctx->sky_lum_tex};
    P.sky_trans = LutView{ctx->sky_trans.p, ctx->sky_trans.w, ctx->sky_trans.h, 1, ctx->sky_trans_tex};
    P.ap_lum = LutView{ctx->ap_lum.p, ctx->ap_lum.w, ctx->ap_lum.h, ctx->ap_lum.d, ctx->ap_lum_tex};
    P.ap_trans = LutView{ctx->ap_trans.p, ctx->ap_trans.w, ctx->ap_trans.h, ctx->ap_trans.d, ctx->ap_trans_tex};
    P.density_tex = ctx->density_tex;
    P.froxel = FroxelView{ctx->shadow_froxel.p, ctx->shadow_froxel.w, ctx->shadow_froxel.h, ctx->shadow_froxel.d, froxel_texture(ctx)};
    P.blue_noise = ctx->blue_noise;
    P.star_map = ctx->star_map.p; P.srgb_decode = ctx->srgb_decode; P.star_w = ctx->star_map.w; P.star_h = ctx->star_map.h;
    P.sky_lum_out = ctx->sky_lum.p; P.sky_trans_out = ctx->sky_trans.p;
    P.ap_lum_out = ctx->ap_lum.p; P.ap_trans_out = ctx->ap_trans.p;
    P.sky_lum_h_out = ctx->sky_lum_h.p; P.sky_trans_h_out = ctx->sky_trans_h.p; P.ap_lum_h_out = ctx->ap_lum_h.p; P.ap_trans_h_out = ctx->ap_trans_h.p;
    P.env_out = ctx->env.p;
    P.extras.moon_shadow = P.cfg.moon_shadow;
    P.extras.moon_radius = P.r.moon_radius;
    for (int i = 0; i < 3; ++i) P.extras.moon_position[i] = P.r.moon_position[i];
    P.extras.shadow_size = P.cfg.volumetric_light ? ctx->mesh_shadow_map.w : 0;
    P.extras.shadow_map = ctx->mesh_shadow_map.p;
    for (int i = 0; i < 16; ++i) P.extras.light_view_projection[i] = P.r.light_view_projection[i];
    ObjectParams& O = P.object;
    O.albedo = static_cast<const uchar4*>(ctx->gbuffer_albedo);
    O.normal = static_cast<const short4*>(ctx->gbuffer_normal);
    O.orm = static_cast<const ushort4*>(ctx->gbuffer_orm);
    O.env_brdf_lut = ctx->env_brdf_lut.p; O.env_brdf_size = ctx->env_brdf_lut.w;
    O.prefiltered.n = SKY_IBL_PREFILTERED_RESOLUTION; O.prefiltered.levels = SKY_IBL_ROUGHNESS_COUNT;
    const half4* pl = ctx->prefiltered;
    for (int l = 0; l < SKY_IBL_ROUGHNESS_COUNT && pl; ++l) { O.prefiltered.level[l] = pl; pl += size_t(6) * (O.prefiltered.n >> l) * (O.prefiltered.n >> l); }
    O.Llm = ctx->env_sh.p;
    O.cloud_shadow_map = ctx->shadow_maps[2].p; O.cloud_shadow_size = ctx->shadow_maps[2].w;
    O.mesh = P.extras;
    O.mesh.shadow_size = ctx->mesh_shadow_map.p ? ctx->mesh_shadow_map.w : 0;
    return P;
}

}  // namespace

// This file is compiled twice: once with -fmad=false for the LUT kernels (K1-K5, bit-faithful to the
// reference's unfused arithmetic) and once with -DSKY_COMPOSITE_TU -use_fast_math for the full-screen
// composite K6, which is a frame (tolerance: relative RMS 1e-2) and ALU-bound on its per-pixel raymarch.
#ifndef SKY_COMPOSITE_TU
// RGBA16F copies of the two bake LUTs for K6's texture fetches (context.h)
int launch_lut_half_copies(SkyContext* ctx) {
    const int na = ctx->transmittance.w * ctx->transmittance.h, nb = ctx->multiscattering.w * ctx->multiscattering.h, nd = ctx->density_h.w;
    k_luts_to_half<<<ceil_div(na + nb + nd, 256), 256, 0, ctx->stream>>>(ctx->transmittance.p, ctx->transmittance_h.p, na, ctx->multiscattering.p,
                                                                          ctx->multiscattering_h.p, nb, ctx->density_h.p, nd, ctx->atm);
    SKY_LAUNCH_CHECK(ctx);
    return 0;
}

// ... and of the four per-frame LUTs after sky_write_resource replaced one of them (K3 / K4 write their copies themselves)
int launch_frame_lut_half_copies(SkyContext* ctx) {
    struct { const Lut<float4>* src; const Lut<half4>* dst; } pairs[4] = {{&ctx->sky_lum, &ctx->sky_lum_h}, {&ctx->sky_trans, &ctx->sky_trans_h},
                                                                           {&ctx->ap_lum, &ctx->ap_lum_h}, {&ctx->ap_trans, &ctx->ap_trans_h}};
    for (auto& pr : pairs) {
        if (!pr.src->p || !pr.dst->p) continue;
        const int n = pr.src->w * pr.src->h * pr.src->d;
        k_luts_to_half<<<ceil_div(n, 256), 256, 0, ctx->stream>>>(pr.src->p, pr.dst->p, n, nullptr, nullptr, 0, nullptr, 0, ctx->atm);
        SKY_LAUNCH_CHECK(ctx);
    }
    return 0;
}

// lanes per march of the cooperative kernels: 4 unless SKYB200_LUT_LANES says 2, 8 or 16.  Measured on B200, scene c3, back-to-back launches
// (profiles/lut_coop_variants_r02E.log): K1 + K2 + copies 153 us exact -> 46.5 / 46.5 / 61 / 77 us with 2 / 4 / 8 / 16 lanes, K3 + K4 + K5
// 158 us -> 66 / 60 / 70 / 86 us.  Few lanes win: the exact ray set-up is replicated per lane, and 4 lanes already fill the machine 1.7 times over.
// 64 instead of 80 registers, and the step loop unrolled by two, are within 2 us of this either way (same log).
static int lut_lanes() {
    const char* e = getenv("SKYB200_LUT_LANES");
    const int n = e ? atoi(e) : 4;
    return n == 2 || n == 8 || n == 16 ? n : 4;
}

int launch_atmosphere_bake(SkyContext* ctx) {
    BakeParams P{};
    P.atm.u = ctx->atm;
    P.transmittance = LutView{ctx->transmittance.p, ctx->transmittance.w, ctx->transmittance.h, 1, 0};
    P.transmittance_out = ctx->transmittance.p;
    P.multiscattering_out = ctx->multiscattering.p;
    P.ms_w = ctx->multiscattering.w; P.ms_h = ctx->multiscattering.h;
    SKY_PERF_MARKER("UpdateLuts");  // Atmosphere.cpp:102
    nvtxRangePushA("Transmittance");  // :112
    k1_transmittance<<<dim3(ceil_div(P.transmittance.w, 128), P.transmittance.h), 128, 0, ctx->stream>>>(P);
    nvtxRangePop();
    SKY_LAUNCH_CHECK(ctx);
    nvtxRangePushA("Multiscattering");  // :116
    if (ctx->lut_arithmetic == SKY_LUT_COOPERATIVE) {
        switch (lut_lanes()) {
            case 2: k2_multiscattering_cooperative<2><<<dim3(P.ms_w, P.ms_h), 128, 0, ctx->stream>>>(P); break;
            case 8: k2_multiscattering_cooperative<8><<<dim3(P.ms_w, P.ms_h), 512, 0, ctx->stream>>>(P); break;
            case 16: k2_multiscattering_cooperative<16><<<dim3(P.ms_w, P.ms_h), 1024, 0, ctx->stream>>>(P); break;
            default: k2_multiscattering_cooperative<4><<<dim3(P.ms_w, P.ms_h), 256, 0, ctx->stream>>>(P); break;
        }
    } else k2_multiscattering<<<dim3(P.ms_w, P.ms_h), 64, 0, ctx->stream>>>(P);
    nvtxRangePop();
    SKY_LAUNCH_CHECK(ctx);
    return launch_lut_half_copies(ctx);
}

int launch_atmosphere_luts(SkyContext* ctx) {
    RenderParams P = make_render_params(ctx);
    if (P.cfg.volumetric_light) {
        if (int e = ensure_mesh_shadow_map(ctx)) return e;
        P = make_render_params(ctx);
    }
    const bool extra = P.cfg.moon_shadow || P.cfg.volumetric_light;
    // K3: 64 texels of a row per block; K4: 64 threads = two rows of the 32-wide froxel slice
    const int g3x = ceil_div(P.cfg.sky_view_width, 64), n3 = g3x * P.cfg.sky_view_height;
    const int g4x = ceil_div(P.ap_lum.h, 64 / P.ap_lum.w), n4 = g4x * P.ap_lum.d;
    nvtxRangePushA("SkyViewLut + AerialPerspective");  // AtmosphereRenderer.cpp:222,229 (one launch here)
    if (!extra && ctx->lut_arithmetic != SKY_LUT_EXACT) {   // (the two optional march terms stay on the exact kernel)
        const int lanes = lut_lanes(), marches = 256 / lanes;
        const int c3x = ceil_div(P.cfg.sky_view_width, marches), c3 = c3x * P.cfg.sky_view_height;
        const int c4x = ceil_div(P.ap_lum.w, marches), c4 = c4x * P.ap_lum.h * P.ap_lum.d;
        switch (lanes) {
            case 2: k34_cooperative<2><<<c3 + c4, 256, 0, ctx->stream>>>(P, c3, c3x, c4x); break;
            case 8: k34_cooperative<8><<<c3 + c4, 256, 0, ctx->stream>>>(P, c3, c3x, c4x); break;
            case 16: k34_cooperative<16><<<c3 + c4, 256, 0, ctx->stream>>>(P, c3, c3x, c4x); break;
            default: k34_cooperative<4><<<c3 + c4, 256, 0, ctx->stream>>>(P, c3, c3x, c4x); break;
        }
    } else if (extra) k34_sky_view_and_aerial_perspective<true><<<n3 + n4, 64, 0, ctx->stream>>>(P, n3, g3x, g4x);
    else k34_sky_view_and_aerial_perspective<false><<<n3 + n4, 64, 0, ctx->stream>>>(P, n3, g3x, g4x);
    nvtxRangePop();
    SKY_LAUNCH_CHECK(ctx);
    SKY_PERF_MARKER("EnvironmentLuminance");  // :236
    k5_environment<<<dim3(ceil_div(P.cfg.environment_size, 128), P.cfg.environment_size, 6), 128, 0, ctx->stream>>>(P);
    SKY_LAUNCH_CHECK(ctx);
    return 0;
}

#else   // SKY_COMPOSITE_TU
#ifdef SKY_STRICT_TU
#define launch_composite launch_composite_strict
#define launch_tonemap launch_tonemap_strict
#endif
namespace {
// shaders/Base/BloomPass2.frag:15-42 without the bloom term
struct ToneMapKernelParams {
    SkyToneMapParams p;
    const half4* hdr;
    uchar4* out;
    const uint16_t* blue_noise;
    int width, height;
};
SKY_D float ToneMapping(float luminance, float exposure, int mode) {
    if (mode == 0) return 1 - expf(-exposure * luminance);
    const float k = 10.0f / 16.0f;
    const float A = 2.51f * k * k, B = 0.03f * k, C = 2.43f * k * k, D = 0.59f * k, E = 0.14f;
    luminance *= exposure;
    return (luminance * (A * luminance + B)) / (luminance * (C * luminance + D) + E);
}
__global__ void __launch_bounds__(256) k21_tonemap(const __grid_constant__ ToneMapKernelParams P) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.width) return;
    float4 c = load_half4(P.hdr + size_t(y) * P.width + x);
    float rgb[3] = {c.x, c.y, c.z};
    unsigned char o[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float v = powf(ToneMapping(rgb[k], P.p.exposure, P.p.tone_mapping), 1.0f / 2.2f);
        if (P.p.dither) v += (float(__ldg(P.blue_noise + (y & 63) * 64 + (x & 63))) / 65535.0f) / 255.0f;
        v = fminf(fmaxf(v, 0.0f), 1.0f);   // the SDR framebuffer clamps (NaN -> 0, fmaxf)
        o[k] = (unsigned char)__float2int_rn(v * 255.0f);
    }
    P.out[size_t(y) * P.width + x] = make_uchar4(o[0], o[1], o[2], 255);
}
}  // namespace
int launch_tonemap(SkyContext* ctx, const half4* hdr, int w, int h, const SkyToneMapParams& p, void* out) {
    if (p.tone_mapping != 0 && p.tone_mapping != 1) return sky_fail(ctx, "tonemap: unknown tone mapping operator");
    SKY_PERF_MARKER("PostProcess");  // HDRBuffer.cpp:42
    ToneMapKernelParams P{p, hdr, static_cast<uchar4*>(out), ctx->blue_noise, w, h};
    k21_tonemap<<<dim3(ceil_div(w, 256), h), 256, 0, ctx->stream>>>(P);
    SKY_LAUNCH_CHECK(ctx);
    return 0;
}

int launch_composite(SkyContext* ctx, const float* depth, half4* hdr, int w, int h) {
    if (ctx->lut_cfg.volumetric_light) { if (int e = ensure_mesh_shadow_map(ctx)) return e; }
    RenderParams P = make_render_params(ctx);
    P.depth = depth; P.hdr = hdr; P.width = w; P.height = h;
    P.band_rows = ctx->out_band_rows; P.band_index = ctx->out_band_index; P.band_count = ctx->out_band_count;
    const int rows = owned_rows(ctx, h);
    if (rows <= 0) return 0;
    const bool extra = P.cfg.moon_shadow || P.cfg.volumetric_light;
    const SkyAtmosphereBufferData& a = ctx->atm;   // grey Mie term (all shipped scenes): the march's scalar-Mie instantiation
    const bool grey = a.mie_scattering[0] == a.mie_scattering[1] && a.mie_scattering[1] == a.mie_scattering[2] &&
                      a.mie_absorption[0] == a.mie_absorption[1] && a.mie_absorption[1] == a.mie_absorption[2];
    SKY_PERF_MARKER("Render");  // AtmosphereRenderer.cpp:247
    const dim3 grid(ceil_div(w, 256), rows);
    if (ctx->gbuffer_albedo) {  // object branch (sky_set_gbuffer); api.cu has checked that the IBL chain exists
        if (P.cfg.pcss) {
            if (extra) k6_composite<true, true, true><<<grid, 256, 0, ctx->stream>>>(P);
            else k6_composite<false, true, true><<<grid, 256, 0, ctx->stream>>>(P);
        } else if (extra) k6_composite<true, true><<<grid, 256, 0, ctx->stream>>>(P);
        else if (grey) k6_composite<false, true, false, false, true><<<grid, 256, 0, ctx->stream>>>(P);
        else k6_composite<false, true><<<grid, 256, 0, ctx->stream>>>(P);
    } else if (extra) k6_composite<true, false><<<grid, 256, 0, ctx->stream>>>(P);
    else if (P.cfg.use_sky_view_lut && P.cfg.use_aerial_perspective_lut) k6_composite<false, false, false, true><<<grid, 256, 0, ctx->stream>>>(P);
    else if (grey) k6_composite<false, false, false, false, true><<<grid, 256, 0, ctx->stream>>>(P);
    else k6_composite<false, false><<<grid, 256, 0, ctx->stream>>>(P);
    SKY_LAUNCH_CHECK(ctx);
    return 0;
}
#endif  // SKY_COMPOSITE_TU
