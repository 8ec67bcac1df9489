// ORACLE -- TEST INFRASTRUCTURE ONLY (see glsl.h).
//
// pathtrace.cpp: CPU restatement of shaders/SkyRendering/VolumetricCloudPathTracing.comp (K19
// render pass, K20 display pass) with the compile-time constants VolumetricCloud.cpp:505-519
// bakes into the shader text taken from SkyPathTracingInit.  The reference's quirks are kept on
// purpose: Random01 returns then advances; TransmittanceEstimation takes
// the context BY VALUE so the shadow ray's random numbers are replayed by the next bounce;
// scattered_t is measured to the segment origin; CloudRegionIntersect divides by zero components.
#include "../include/sky_cubemap.h"
#include "cloud.h"

namespace orc {
namespace {

struct Ray { vec3 o, d; };
struct Context { uint seed; Ray ray; float sigma_t_max; };

inline float HenyeyGreenstein(float cos_theta, float g) {  // VolumetricCloudCommon.glsl:58-63
    float a = 1.0f - g * g;
    float b = 1.0f + g * g - 2.0f * g * cos_theta;
    b *= std::sqrt(b);
    return (0.25f * INV_PI) * a / b;
}
inline float HenyeyGreensteinInvertcdf(float xi, float g) {  // VolumetricCloudCommon.glsl:65-71
    float one_plus_g2 = 1.0f + g * g;
    float one_minus_g2 = 1.0f - g * g;
    float one_over_2g = 0.5f / g;
    float t = (one_minus_g2) / (1.0f - g + 2.0f * g * xi);
    return one_over_2g * (one_plus_g2 - t * t);
}
inline void CreateOrthonormalBasis(vec3 N, vec3& t0, vec3& t1) {  // shaders/Base/Common.glsl:32-52
    float s = (N.z >= 0.0f ? 1.0f : -1.0f);
    float a = -1.0f / (s + N.z);
    float b = N.x * N.y * a;
    t0 = vec3(1.0f + s * N.x * N.x * a, s * b, -s * N.x);
    t1 = vec3(b, s + N.y * N.y * a, -N.y);
}

struct Tracer {
    CloudScene& S;
    const SkyPathTracingInit& P;
    vec3 kCloudAABBMin, kCloudAABBMax;  // VolumetricCloudPathTracing.comp:19-21
    mat3 kModelMatrix3;

    Tracer(CloudScene& s)
        : S(s), P(s.pt),
          kCloudAABBMin(-s.pt.region_box_half_width, -s.pt.region_box_half_width, s.c.uBottomAltitude),
          kCloudAABBMax(s.pt.region_box_half_width, s.pt.region_box_half_width, s.c.uTopAltitude),
          kModelMatrix3(s.pt.model_matrix3) {}

    uint PRNG(uint x) const { return P.prng == SKY_PRNG_WANG ? WangHash(x) : PCGHash(x); }

    float Random01(Context& ctx) const {  // :44-48
        float res = float(ctx.seed) / 4294967296.0f;
        ctx.seed = PRNG(ctx.seed);
        return res;
    }
    vec3 UniformSphereSample(Context& ctx) const {  // :50-55
        float phi = 2.0f * PI * Random01(ctx);
        float cos_theta = 1.0f - 2.0f * Random01(ctx);
        float sin_theta = std::sqrt(clamp(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
        return vec3(std::cos(phi) * sin_theta, std::sin(phi) * sin_theta, cos_theta);
    }
    vec2 CloudRegionIntersect(const Ray& ray) const {  // :57-75
        vec2 t(0.0f, 1e7f);
        for (int i = 0; i < 3; ++i) {
            float t1 = (kCloudAABBMin[i] - ray.o[i]) / ray.d[i];
            float t2 = (kCloudAABBMax[i] - ray.o[i]) / ray.d[i];
            float tmin = min(t1, t2);  // minNum/maxNum semantics, see glsl.h
            float tmax = max(t1, t2);
            t.x = max(t.x, tmin);
            t.y = min(t.y, tmax);
        }
        return t;
    }
    static float InfiniteTransmittanceIS(float sigma_t, float zeta) { return -std::log(1.0f - zeta) / sigma_t; }  // :82-84
    float SampleSigmaT(vec3 Pp) const {  // :89-91
        return S.SampleSigmaT(Pp, clamp((Pp.z - S.c.uBottomAltitude) / (S.c.uTopAltitude - S.c.uBottomAltitude), 0.0f, 1.0f),
                              SKY_CNT_PT_LOOKUPS);
    }
    float GetPhase(float cos_theta) const {  // :93-96
        return mix(HenyeyGreenstein(cos_theta, P.back_phase_g), HenyeyGreenstein(cos_theta, P.forward_phase_g),
                   P.forward_scattering_ratio);
    }
    void GenerateHGSample(Context& ctx, vec3& direction, float& value_over_pdf) const {  // :98-117
        if (P.importance_sampling) {
            float g = Random01(ctx) < P.forward_scattering_ratio ? P.forward_phase_g : P.back_phase_g;
            float cos_theta = HenyeyGreensteinInvertcdf(Random01(ctx), g);
            float sin_theta = std::sqrt(clamp(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
            vec3 t0, t1;
            CreateOrthonormalBasis(ctx.ray.d, t0, t1);
            float phi = 2.0f * PI * Random01(ctx);
            direction = sin_theta * std::sin(phi) * t0 + sin_theta * std::cos(phi) * t1 + cos_theta * ctx.ray.d;
            value_over_pdf = 1.0f;
        } else {
            direction = UniformSphereSample(ctx);
            float value = GetPhase(dot(ctx.ray.d, direction));
            float pdf = 1.0f / (4.0f * PI);
            value_over_pdf = value / pdf;
        }
    }
    void GenerateLambertSample(Context& ctx, vec3 N, vec3 albedo, vec3& direction, vec3& value_over_pdf) const {  // :119-129
        float sin_theta = std::sqrt(Random01(ctx));
        float cos_theta = std::sqrt(clamp(1.0f - sin_theta * sin_theta, 0.0f, 1.0f));
        vec3 t0, t1;
        CreateOrthonormalBasis(N, t0, t1);
        float phi = 2.0f * PI * Random01(ctx);
        direction = sin_theta * std::sin(phi) * t0 + sin_theta * std::cos(phi) * t1 + cos_theta * N;
        value_over_pdf = albedo;
    }
    vec3 GetSunIlluminance(vec3 pos) const { return S.GetSunVisibility(pos) * S.atm.solar_illuminance(); }  // :131-133

    float TransmittanceEstimation(Context ctx /* by value! */, Ray ray) const {  // :135-151
        float transmittance = 1.0f;
        vec2 inter_t = CloudRegionIntersect(ray);
        if (inter_t.x >= inter_t.y) return transmittance;
        float t = inter_t.x;
        while (true) {
            t += InfiniteTransmittanceIS(ctx.sigma_t_max, Random01(ctx));
            if (t > inter_t.y) break;
            float sigma_t = SampleSigmaT(ray.o + ray.d * t);
            transmittance *= 1.0f - max(0.0f, sigma_t / ctx.sigma_t_max);
            if (S.counting) S.counters[SKY_CNT_PT_COLLISIONS].fetch_add(1, std::memory_order_relaxed);
        }
        return clamp(transmittance, 0.0f, 1.0f);
    }
    vec3 SampleLuminanceFromLight(Context& ctx, vec3 pos, vec3 bsdf_with_cosine) const {  // :153-158
        const float kSunSolidAngle = 1.0f;
        vec3 light_luminance = GetSunIlluminance(pos) / kSunSolidAngle;
        float pdf = 1.0f / kSunSolidAngle;
        return TransmittanceEstimation(ctx, Ray{pos, S.uSunDirection()}) * light_luminance * bsdf_with_cosine / pdf;
    }
    // environment_luminance_texture lookup, :206,215.  The cube is sampled with implicit LOD inside a
    // compute shader (derivatives undefined): the oracle defines LOD 0.  Face selection / (s,t) follow the GL cube-map table
    // (spec 8.13); bilinear filtering is SEAMLESS (GL 4.6 8.14.1; glEnable(GL_TEXTURE_CUBE_MAP_SEAMLESS), AtmosphereRenderer.cpp:151).
    vec3 SampleEnvironment(vec3 dir) const {
        float ax = std::fabs(dir.x), ay = std::fabs(dir.y), az = std::fabs(dir.z);
        int face; float sc, tc, ma;
        if (ax >= ay && ax >= az) { ma = ax; if (dir.x >= 0) { face = 0; sc = -dir.z; tc = -dir.y; } else { face = 1; sc = dir.z; tc = -dir.y; } }
        else if (ay >= az)        { ma = ay; if (dir.y >= 0) { face = 2; sc = dir.x; tc = dir.z; } else { face = 3; sc = dir.x; tc = -dir.z; } }
        else                      { ma = az; if (dir.z >= 0) { face = 4; sc = dir.x; tc = -dir.y; } else { face = 5; sc = -dir.x; tc = -dir.y; } }
        float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
        const Image<4>& env = S.env;
        int n = env.w;
        float u = s * float(n) - 0.5f, v = t * float(n) - 0.5f;
        float fu = std::floor(u), fv = std::floor(v);
        int i0 = int(fu), j0 = int(fv);
        float a = u - fu, b = v - fv;
        return sky_cube_bilinear<vec3>(n, face, i0, j0, a, b, [&](int f, int i, int j) { return env.load(i, j, f).rgb(); });
    }

    vec4 Trace(Context& ctx, vec3 view_dir, bool& has_scattered, float& scattered_t) const {  // :160-249
        vec3 L(0.0f);
        vec3 throughput(1.0f);
        has_scattered = false;
        ctx.ray = Ray{S.uCameraPos(), view_dir};
        vec2 camera_inter_t = CloudRegionIntersect(ctx.ray);
        if (camera_inter_t.x >= camera_inter_t.y) return vec4(L, throughput.x);  // vec4(vec3, vec3) keeps 4 components

        ctx.ray.o += camera_inter_t.x * ctx.ray.d;
        int istep = 0;
        while (istep < P.max_bounces && max(throughput.x, max(throughput.y, throughput.z)) > 0.0f) {
            vec2 inter_t = CloudRegionIntersect(ctx.ray);
            if (inter_t.x >= inter_t.y) break;
            float t_max = inter_t.y;
            float t = inter_t.x;
            bool event_scatter = false;
            while (true) {
                if (ctx.sigma_t_max <= 0) break;
                t += InfiniteTransmittanceIS(ctx.sigma_t_max, Random01(ctx));
                if (t > t_max) break;
                vec3 Pp = ctx.ray.o + ctx.ray.d * t;
                float sigma_t = SampleSigmaT(Pp);
                if (S.counting) S.counters[SKY_CNT_PT_COLLISIONS].fetch_add(1, std::memory_order_relaxed);
                float xi = Random01(ctx);
                if (xi < sigma_t / ctx.sigma_t_max) { event_scatter = true; break; }
            }
            if (!event_scatter) {
                if (P.environment_lighting == SKY_ENV_OFF) break;
                if (!has_scattered) break;
                if (P.environment_lighting == SKY_ENV_CONST_ENVIRONMENT_MAP) {
                    L += throughput * SampleEnvironment(kModelMatrix3 * ctx.ray.d);
                    break;
                }
                vec3 up_dir(ctx.ray.o.x, ctx.ray.o.y, ctx.ray.o.z + S.c.uEarthRadius);
                float r = length(up_dir);
                up_dir /= r;
                float mu = dot(ctx.ray.d, up_dir);
                if (!S.atm.RayIntersectsGround(r, mu)) {
                    L += throughput * SampleEnvironment(kModelMatrix3 * ctx.ray.d);
                    break;
                }
                ctx.ray.o += ctx.ray.d * S.atm.DistanceToBottomAtmosphereBoundary(r, mu);
                vec3 ground_normal = normalize(vec3(ctx.ray.o.x, ctx.ray.o.y, ctx.ray.o.z + S.c.uEarthRadius));
                vec3 light_bsdf = INV_PI * S.atm.ground_albedo();
                float NdotL = dot(ground_normal, S.uSunDirection());
                L += throughput * SampleLuminanceFromLight(ctx, ctx.ray.o, light_bsdf * NdotL);
                if (P.environment_lighting == SKY_ENV_GROUND_SINGLE_BOUNCE) break;
                vec3 bsdf_with_cosine_over_pdf;
                GenerateLambertSample(ctx, ground_normal, S.atm.ground_albedo(), ctx.ray.d, bsdf_with_cosine_over_pdf);
                throughput *= bsdf_with_cosine_over_pdf;
            } else {
                if (!has_scattered) scattered_t = distance(S.uCameraPos(), ctx.ray.o);
                has_scattered = true;
                ctx.ray.o += ctx.ray.d * t;
                float light_bsdf = GetPhase(dot(ctx.ray.d, S.uSunDirection()));
                L += throughput * SampleLuminanceFromLight(ctx, ctx.ray.o, vec3(light_bsdf));
                float bsdf_over_pdf;
                GenerateHGSample(ctx, ctx.ray.d, bsdf_over_pdf);
                throughput *= bsdf_over_pdf;
            }
            ++istep;
        }
        return vec4(L, has_scattered ? 0.0f : 1.0f);
    }
};

}  // namespace

void CloudScene::PathTraceBegin(const SkyPathTracingInit& init) {
    pt = init;
    pt_accum.resize(width, height);
    pt_mask.assign(size_t(width) * height, 0);
}

// K19 -- VolumetricCloudPathTracing.comp:255-284 for kFrameId = frame_begin .. frame_begin+count-1
void CloudScene::PathTraceSamples(uint32_t frame_begin, uint32_t count, const int32_t region[4]) {
    Tracer T(*this);
    const mat4 uInvMVP(c.uInvMVP);
    for (uint32_t f = 0; f < count; ++f) {
        const uint kFrameId = frame_begin + f;
#pragma omp parallel for schedule(dynamic)
        for (int py = region[1]; py < min(region[3], height); ++py)
            for (int px = region[0]; px < min(region[2], width); ++px) {
                Context ctx;
                ctx.seed = T.PRNG(T.PRNG(T.PRNG(uint(px)) + uint(py)) + kFrameId);
                ctx.sigma_t_max = pt.sigma_t_max;
                vec2 uv = (vec2(float(px), float(py)) + 0.5f) / vec2(float(width), float(height));
                vec3 frag_pos = ProjectiveMul(uInvMVP, vec3(uv, 1.0f) * 2.0f - 1.0f);
                vec3 view_dir = normalize(frag_pos - uCameraPos());
                bool has_scattered;
                float scattered_t = 0.0f;
                vec4 this_res = T.Trace(ctx, view_dir, has_scattered, scattered_t);
                if (has_scattered) {
                    float r = c.uCameraPos[2] + c.uEarthRadius;
                    float mu = view_dir.z;
                    vec3 atmosphere_transmittance;
                    vec3 atmosphere_luminance = GetAerialPerspective(uv, scattered_t, r, mu, atmosphere_transmittance);
                    atmosphere_luminance *= SampleRayScatterVisibility(shadow_froxel, uv, scattered_t, c.uInvShadowFroxelMaxDistance);
                    vec3 rgb = this_res.rgb() * atmosphere_transmittance + atmosphere_luminance;
                    this_res = vec4(rgb, this_res.w);
                }
                float* a = pt_accum.at(px, py);
                for (int k = 0; k < 4; ++k) a[k] = a[k] + this_res[k];
                pt_mask[size_t(py) * width + px] = 1;
                if (counting) counters[SKY_CNT_PT_PATHS].fetch_add(1, std::memory_order_relaxed);
            }
    }
}

// K20 -- VolumetricCloudPathTracing.comp:288-296
void CloudScene::PathTraceResolve(uint32_t kFrameId, uint16_t* hdr) const {
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            const float* a = pt_accum.at(x, y);
            bool is_rendered = pt_mask[size_t(y) * width + x] != 0;
            float div = float(is_rendered ? kFrameId : kFrameId - 1);
            uint16_t* o = hdr + (size_t(y) * width + x) * 4;
            for (int k = 0; k < 3; ++k) {
                float color = half_bits_to_float(o[k]);
                o[k] = float_to_half_bits(color * (a[3] / div) + a[k] / div);
            }
        }
}

}  // namespace orc
