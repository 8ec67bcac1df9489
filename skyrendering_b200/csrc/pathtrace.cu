// K19 / K20: delta-tracking / ratio-tracking voxel-cloud path tracer, sm_100a.
// Follows shaders/SkyRendering/VolumetricCloudPathTracing.comp; the constants its host bakes into the
// shader text (VolumetricCloud.cpp:505-519) arrive in SkyPathTracingInit.
//
// Stream-exact mode: the random-number stream of every (pixel, kFrameId) is the reference's
// (seed = PRNG(PRNG(PRNG(x)+y)+frame), Random01 returns-then-advances, the shadow ray replays the
// parent's stream because TransmittanceEstimation takes its Context by value, :135).  The only
// liberty taken is exact: tentative collisions that provably land on zero density (outside the voxel
// footprint, where CLAMP_TO_BORDER returns 0 at every mip level) skip the texture fetch but still
// consume their random numbers.
#include "atmosphere_dev.cuh"
#include "context.h"
#include "material_dev.cuh"

namespace {

struct PtParams {
    SkyCloudCommonBufferData c;
    SkyPathTracingInit pt;
    AtmosphereModel atm;
    MaterialParams mat;
    LutView transmittance, ap_lum, ap_trans;
    FroxelView froxel;
    const half4* env;  // [6][S][S]
    int env_size;
    float4* accum;
    uint8_t* mask;
    half4* hdr;
    unsigned long long* counters;
    int width, height;
    int x0, y0, x1, y1;  // kRenderRegion
    uint32_t frame_begin, frame_count;
};

SKY_D uint32_t WangHash(uint32_t seed) {  // shaders/Base/Noise.glsl:1-8
    seed = (seed ^ 61u) ^ (seed >> 16);
    seed *= 9u;
    seed = seed ^ (seed >> 4);
    seed *= 0x27d4eb2du;
    seed = seed ^ (seed >> 15);
    return seed;
}
SKY_D uint32_t PCGHash(uint32_t seed) {  // shaders/Base/Noise.glsl:11-15
    uint32_t state = seed * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (word >> 22u) ^ word;
}
template <int PRNG_KIND>
SKY_D uint32_t PRNG(uint32_t x) { return PRNG_KIND == SKY_PRNG_WANG ? WangHash(x) : PCGHash(x); }

struct Ray { float3 o, d; };

template <int PRNG_KIND>
SKY_D float Random01(uint32_t& seed) {  // :44-48 (returns, then advances; can round to 1.0)
    float res = float(seed) / 4294967296.0f;
    seed = PRNG<PRNG_KIND>(seed);
    return res;
}

SKY_D float HenyeyGreenstein(float cos_theta, float g) {  // VolumetricCloudCommon.glsl:58-63
    float a = 1.0f - g * g;
    float b = 1.0f + g * g - 2.0f * g * cos_theta;
    b *= sqrtf(b);
    return (0.25f * kInvPi) * a / b;
}
SKY_D float HenyeyGreensteinInvertcdf(float xi, float g) {  // VolumetricCloudCommon.glsl:65-71
    float one_plus_g2 = 1.0f + g * g;
    float one_minus_g2 = 1.0f - g * g;
    float one_over_2g = 0.5f / g;
    float t = (one_minus_g2) / (1.0f - g + 2.0f * g * xi);
    return one_over_2g * (one_plus_g2 - t * t);
}
SKY_D void CreateOrthonormalBasis(float3 N, float3& t0, float3& t1) {  // shaders/Base/Common.glsl:32-52
    float s = (N.z >= 0.0f ? 1.0f : -1.0f);
    float a = -1.0f / (s + N.z);
    float b = N.x * N.y * a;
    t0 = f3(1.0f + s * N.x * N.x * a, s * b, -s * N.x);
    t1 = f3(b, s + N.y * N.y * a, -N.y);
}

// :57-75.  IEEE division keeps the reference's inf/NaN behaviour for zero direction components;
// min/max are the hardware FMNMX (minNum/maxNum), which is what GLSL min/max compile to.
SKY_D float2 CloudRegionIntersect(const PtParams& P, const Ray& ray) {
    const float hw = P.pt.region_box_half_width;
    const float bmin[3] = {-hw, -hw, P.c.uBottomAltitude}, bmax[3] = {hw, hw, P.c.uTopAltitude};
    const float o[3] = {ray.o.x, ray.o.y, ray.o.z}, d[3] = {ray.d.x, ray.d.y, ray.d.z};
    float2 t = f2(0.0f, 1e7f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float t1 = (bmin[i] - o[i]) / d[i];
        float t2 = (bmax[i] - o[i]) / d[i];
        float tmin = fminf(t1, t2);
        float tmax = fmaxf(t1, t2);
        t.x = fmaxf(t.x, tmin);
        t.y = fminf(t.y, tmax);
    }
    return t;
}

SKY_D float InfiniteTransmittanceIS(float sigma_t, float zeta) { return -logf(1.0f - zeta) / sigma_t; }  // :82-84

template <int MAT, bool HW, bool COUNT>
SKY_D float SampleSigmaTAt(const PtParams& P, float3 pos, int& lookups) {  // :89-91
    float height01 = clampf((pos.z - P.c.uBottomAltitude) / (P.c.uTopAltitude - P.c.uBottomAltitude), 0.0f, 1.0f);
    if (MAT == SKY_MATERIAL_VOXEL) {
        // exact empty-space skip: with CLAMP_TO_BORDER(0) every tap of every level is the border here
        const SkyMaterialVoxelBufferData& m = P.mat.m.u.voxel;
        float u = pos.x * m.uSampleFrequency[0] + m.uSampleBias[0];
        float v = pos.y * m.uSampleFrequency[1] + m.uSampleBias[1];
        float hu = 0.5f / float(P.mat.voxel.w[0]), hv = 0.5f / float(P.mat.voxel.h[0]);
        if (u < -hu || u > 1.0f + hu || v < -hv || v > 1.0f + hv) return 0.0f;
    }
    if (COUNT) ++lookups;
    return SampleSigmaT<MAT, HW>(P.mat, pos, height01);
}

SKY_D float GetPhase(const PtParams& P, float cos_theta) {  // :93-96
    return mixf(HenyeyGreenstein(cos_theta, P.pt.back_phase_g), HenyeyGreenstein(cos_theta, P.pt.forward_phase_g),
                P.pt.forward_scattering_ratio);
}

// environment_luminance_texture (:206,215): LOD 0, bilinear inside the selected face (the oracle's
// definition; implicit derivatives are undefined in a compute shader).  Face table: GL 4.6 section 8.13.
SKY_D float3 SampleEnvironment(const PtParams& P, float3 dir) {
    float ax = fabsf(dir.x), ay = fabsf(dir.y), az = fabsf(dir.z);
    int face; float sc, tc, ma;
    if (ax >= ay && ax >= az) { ma = ax; if (dir.x >= 0) { face = 0; sc = -dir.z; tc = -dir.y; } else { face = 1; sc = dir.z; tc = -dir.y; } }
    else if (ay >= az)        { ma = ay; if (dir.y >= 0) { face = 2; sc = dir.x; tc = dir.z; } else { face = 3; sc = dir.x; tc = -dir.z; } }
    else                      { ma = az; if (dir.z >= 0) { face = 4; sc = dir.x; tc = -dir.y; } else { face = 5; sc = -dir.x; tc = -dir.y; } }
    float s = 0.5f * (sc / ma + 1.0f), t = 0.5f * (tc / ma + 1.0f);
    const int n = P.env_size;
    float u = s * float(n) - 0.5f, v = t * float(n) - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    int i0 = int(fu), j0 = int(fv);
    float a = u - fu, b = v - fv;
    const half4* base = P.env + size_t(face) * n * n;
    auto L = [&](int i, int j) { return xyz(load_half4(base + clampi(j, 0, n - 1) * n + clampi(i, 0, n - 1))); };
    return (1.0f - a) * (1.0f - b) * L(i0, j0) + a * (1.0f - b) * L(i0 + 1, j0) + (1.0f - a) * b * L(i0, j0 + 1) + a * b * L(i0 + 1, j0 + 1);
}

SKY_D float3 GetSunIlluminance(const PtParams& P, float3 pos) {  // :131-133 + VolumetricCloudCommon.glsl:73-79
    float3 up_dir = f3(pos.x, pos.y, pos.z + P.c.uEarthRadius);
    float r = length(up_dir);
    up_dir /= r;
    float mu_s = dot(f3(P.c.uSunDirection), up_dir);
    return P.atm.GetSunVisibility(P.transmittance, r, mu_s) * P.atm.solar_illuminance();
}

// :135-151; `seed` is a COPY of the path's stream
template <int MAT, bool HW, int PRNG_KIND, bool COUNT>
SKY_D float TransmittanceEstimation(const PtParams& P, uint32_t seed, const Ray& ray, int& lookups, int& collisions) {
    float transmittance = 1.0f;
    float2 inter_t = CloudRegionIntersect(P, ray);
    if (inter_t.x >= inter_t.y) return transmittance;
    const float sigma_t_max = P.pt.sigma_t_max;
    float t = inter_t.x;
    while (true) {
        t += InfiniteTransmittanceIS(sigma_t_max, Random01<PRNG_KIND>(seed));
        if (t > inter_t.y) break;
        float sigma_t = SampleSigmaTAt<MAT, HW, COUNT>(P, ray.o + ray.d * t, lookups);
        transmittance *= 1.0f - fmaxf(0.0f, sigma_t / sigma_t_max);
        if (COUNT) ++collisions;
    }
    return clampf(transmittance, 0.0f, 1.0f);
}

template <int MAT, bool HW, int PRNG_KIND, bool COUNT>
SKY_D float3 SampleLuminanceFromLight(const PtParams& P, uint32_t seed, float3 pos, float3 bsdf_with_cosine, int& lookups, int& collisions) {  // :153-158
    float3 light_luminance = GetSunIlluminance(P, pos);
    Ray ray{pos, f3(P.c.uSunDirection)};
    return TransmittanceEstimation<MAT, HW, PRNG_KIND, COUNT>(P, seed, ray, lookups, collisions) * light_luminance * bsdf_with_cosine;
}

// :160-249
template <int MAT, bool HW, int PRNG_KIND, bool COUNT>
SKY_D float4 Trace(const PtParams& P, uint32_t& seed, float3 view_dir, bool& has_scattered, float& scattered_t, int& lookups, int& collisions) {
    float3 L = f3(0.0f);
    float3 throughput = f3(1.0f);
    has_scattered = false;
    const float3 camera = f3(P.c.uCameraPos);
    const float3 sun = f3(P.c.uSunDirection);
    const float sigma_t_max = P.pt.sigma_t_max;
    Ray ray{camera, view_dir};
    float2 camera_inter_t = CloudRegionIntersect(P, ray);
    if (camera_inter_t.x >= camera_inter_t.y) return f4(L, throughput.x);

    ray.o += camera_inter_t.x * ray.d;
    int istep = 0;
    while (istep < P.pt.max_bounces && fmaxf(throughput.x, fmaxf(throughput.y, throughput.z)) > 0.0f) {
        float2 inter_t = CloudRegionIntersect(P, ray);
        if (inter_t.x >= inter_t.y) break;
        float t_max = inter_t.y;
        float t = inter_t.x;
        bool event_scatter = false;
        while (true) {
            if (sigma_t_max <= 0) break;
            t += InfiniteTransmittanceIS(sigma_t_max, Random01<PRNG_KIND>(seed));
            if (t > t_max) break;
            float3 Pp = ray.o + ray.d * t;
            float sigma_t = SampleSigmaTAt<MAT, HW, COUNT>(P, Pp, lookups);
            if (COUNT) ++collisions;
            float xi = Random01<PRNG_KIND>(seed);
            if (xi < sigma_t / sigma_t_max) { event_scatter = true; break; }
        }
        if (!event_scatter) {
            if (P.pt.environment_lighting == SKY_ENV_OFF) break;
            if (!has_scattered) break;
            const float* mm = P.pt.model_matrix3;
            float3 world_dir = f3(mm[0] * ray.d.x + mm[3] * ray.d.y + mm[6] * ray.d.z, mm[1] * ray.d.x + mm[4] * ray.d.y + mm[7] * ray.d.z,
                                  mm[2] * ray.d.x + mm[5] * ray.d.y + mm[8] * ray.d.z);
            if (P.pt.environment_lighting == SKY_ENV_CONST_ENVIRONMENT_MAP) {
                L += throughput * SampleEnvironment(P, world_dir);
                break;
            }
            float3 up_dir = f3(ray.o.x, ray.o.y, ray.o.z + P.c.uEarthRadius);
            float r = length(up_dir);
            up_dir /= r;
            float mu = dot(ray.d, up_dir);
            if (!P.atm.RayIntersectsGround(r, mu)) {
                L += throughput * SampleEnvironment(P, world_dir);
                break;
            }
            ray.o += ray.d * P.atm.DistanceToBottomAtmosphereBoundary(r, mu);
            float3 ground_normal = normalize(f3(ray.o.x, ray.o.y, ray.o.z + P.c.uEarthRadius));
            float3 light_bsdf = kInvPi * P.atm.ground_albedo();
            float NdotL = dot(ground_normal, sun);
            L += throughput * SampleLuminanceFromLight<MAT, HW, PRNG_KIND, COUNT>(P, seed, ray.o, light_bsdf * NdotL, lookups, collisions);
            if (P.pt.environment_lighting == SKY_ENV_GROUND_SINGLE_BOUNCE) break;
            // GenerateLambertSample, :119-129
            float sin_theta = sqrtf(Random01<PRNG_KIND>(seed));
            float cos_theta = sqrtf(clampf(1.0f - sin_theta * sin_theta, 0.0f, 1.0f));
            float3 t0, t1;
            CreateOrthonormalBasis(ground_normal, t0, t1);
            float phi = 2.0f * kPi * Random01<PRNG_KIND>(seed);
            ray.d = sin_theta * sinf(phi) * t0 + sin_theta * cosf(phi) * t1 + cos_theta * ground_normal;
            throughput *= P.atm.ground_albedo();
        } else {
            if (!has_scattered) scattered_t = distance(camera, ray.o);
            has_scattered = true;
            ray.o += ray.d * t;
            float light_bsdf = GetPhase(P, dot(ray.d, sun));
            L += throughput * SampleLuminanceFromLight<MAT, HW, PRNG_KIND, COUNT>(P, seed, ray.o, f3(light_bsdf), lookups, collisions);
            // GenerateHGSample, :98-117
            if (P.pt.importance_sampling) {
                float g = Random01<PRNG_KIND>(seed) < P.pt.forward_scattering_ratio ? P.pt.forward_phase_g : P.pt.back_phase_g;
                float cos_theta = HenyeyGreensteinInvertcdf(Random01<PRNG_KIND>(seed), g);
                float sin_theta = sqrtf(clampf(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
                float3 t0, t1;
                CreateOrthonormalBasis(ray.d, t0, t1);
                float phi = 2.0f * kPi * Random01<PRNG_KIND>(seed);
                ray.d = sin_theta * sinf(phi) * t0 + sin_theta * cosf(phi) * t1 + cos_theta * ray.d;
            } else {
                // UniformSphereSample, :50-55
                float phi = 2.0f * kPi * Random01<PRNG_KIND>(seed);
                float cos_theta = 1.0f - 2.0f * Random01<PRNG_KIND>(seed);
                float sin_theta = sqrtf(clampf(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
                float3 direction = f3(cosf(phi) * sin_theta, sinf(phi) * sin_theta, cos_theta);
                float value = GetPhase(P, dot(ray.d, direction));
                ray.d = direction;
                throughput *= value / (1.0f / (4.0f * kPi));
            }
        }
        ++istep;
    }
    return f4(L, has_scattered ? 0.0f : 1.0f);
}

// K19 -- :255-284.  One thread per pixel of kRenderRegion; the thread walks kFrameId =
// frame_begin .. frame_begin+frame_count-1 and adds each sample to the accumulator in frame order
// (the reference's order of fp32 additions), with one read-modify-write of the RGBA32F texel.
template <int MAT, bool HW, int PRNG_KIND, bool COUNT>
__global__ void __launch_bounds__(128) k19_path_trace(const __grid_constant__ PtParams P) {
    int px = P.x0 + blockIdx.x * 16 + (threadIdx.x & 15), py = P.y0 + blockIdx.y * 8 + (threadIdx.x >> 4);
    if (px >= P.x1 || py >= P.y1 || px >= P.width || py >= P.height) return;
    const size_t pix = size_t(py) * P.width + px;
    float2 uv = f2((float(px) + 0.5f) / float(P.width), (float(py) + 0.5f) / float(P.height));
    const float3 camera = f3(P.c.uCameraPos);
    float3 frag_pos = projective_mul(P.c.uInvMVP, f3(uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f, 1.0f));
    float3 view_dir = normalize(frag_pos - camera);
    const uint32_t pixel_seed = PRNG<PRNG_KIND>(PRNG<PRNG_KIND>(uint32_t(px)) + uint32_t(py));
    float4 accumulated = P.accum[pix];
    int lookups = 0, collisions = 0;
    for (uint32_t f = 0; f < P.frame_count; ++f) {
        uint32_t seed = PRNG<PRNG_KIND>(pixel_seed + (P.frame_begin + f));
        bool has_scattered;
        float scattered_t = 0.0f;
        float4 this_res = Trace<MAT, HW, PRNG_KIND, COUNT>(P, seed, view_dir, has_scattered, scattered_t, lookups, collisions);
        if (has_scattered) {
            float r = P.c.uCameraPos[2] + P.c.uEarthRadius;
            float mu = view_dir.z;
            float ap_t = scattered_t;
            if (r > P.atm.u.top_radius) {  // GetAerialPerspective, VolumetricCloudCommon.glsl:81-97
                float near_distance;
                if (P.atm.FromSpaceIntersectTopAtmosphereBoundary(r, mu, near_distance)) ap_t -= near_distance;
                else ap_t = 0;
            }
            float3 uvw = aerial_perspective_uvw(uv, ap_t, P.c.uAerialPerspectiveLutMaxDistance, P.ap_lum.w, P.ap_lum.h, P.ap_lum.d);
            float3 atmosphere_transmittance = xyz(sample_lut3d(P.ap_trans, uvw.x, uvw.y, uvw.z));
            float3 atmosphere_luminance = xyz(sample_lut3d(P.ap_lum, uvw.x, uvw.y, uvw.z));
            atmosphere_luminance *= SampleRayScatterVisibility(P.froxel, uv, scattered_t, P.c.uInvShadowFroxelMaxDistance);
            float3 rgb = xyz(this_res) * atmosphere_transmittance + atmosphere_luminance;
            this_res = f4(rgb, this_res.w);
        }
        accumulated = accumulated + this_res;
    }
    P.accum[pix] = accumulated;
    P.mask[pix] = 1;
    if (COUNT) {
        atomicAdd(P.counters + SKY_CNT_PT_PATHS, (unsigned long long)P.frame_count);
        atomicAdd(P.counters + SKY_CNT_PT_LOOKUPS, (unsigned long long)lookups);
        atomicAdd(P.counters + SKY_CNT_PT_COLLISIONS, (unsigned long long)collisions);
    }
}

// K20 -- :288-296
__global__ void __launch_bounds__(256) k20_display(const __grid_constant__ PtParams P) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= P.width) return;
    size_t pix = size_t(y) * P.width + x;
    float4 accumulated = P.accum[pix];
    float4 color = load_half4(P.hdr + pix);
    bool is_rendered = P.mask[pix] != 0;
    float div = float(is_rendered ? P.frame_begin : P.frame_begin - 1);
    float4 avg = f4(accumulated.x / div, accumulated.y / div, accumulated.z / div, accumulated.w / div);
    color.x = color.x * avg.w + avg.x;
    color.y = color.y * avg.w + avg.y;
    color.z = color.z * avg.w + avg.z;
    P.hdr[pix] = to_half4(color);
}

PtParams make_pt_params(SkyContext* ctx, const SkyCloudCommonBufferData& c) {
    PtParams P{};
    P.c = c;
    P.pt = ctx->pt;
    P.atm.u = ctx->atm;
    make_material_params(ctx, c.uCameraPos, P.mat);
    P.transmittance = LutView{ctx->transmittance.p, ctx->transmittance.w, ctx->transmittance.h, 1};
    P.ap_lum = LutView{ctx->ap_lum.p, ctx->ap_lum.w, ctx->ap_lum.h, ctx->ap_lum.d};
    P.ap_trans = LutView{ctx->ap_trans.p, ctx->ap_trans.w, ctx->ap_trans.h, ctx->ap_trans.d};
    P.froxel = FroxelView{ctx->shadow_froxel.p, ctx->shadow_froxel.w, ctx->shadow_froxel.h, ctx->shadow_froxel.d};
    P.env = ctx->env.p;
    P.env_size = ctx->env.w;
    P.accum = ctx->pt_accum.p;
    P.mask = ctx->pt_mask.p;
    P.counters = ctx->counters;
    P.width = ctx->width; P.height = ctx->height;
    return P;
}

}  // namespace

int launch_pt_samples(SkyContext* ctx, const SkyCloudCommonBufferData& c, uint32_t frame_begin, uint32_t count, const int32_t region[4]) {
    if (!ctx->pt_accum.p) return sky_fail(ctx, "pt_begin was not called");
    if (!ctx->ap_lum.p || !ctx->env.p) return sky_fail(ctx, "atmosphere LUTs have not been baked");
    if (ctx->material.type == SKY_MATERIAL_VOXEL && !ctx->voxel.valid) return sky_fail(ctx, "voxel grid has not been uploaded");
    if ((ctx->material.type == SKY_MATERIAL_DEFAULT0 || ctx->material.type == SKY_MATERIAL_DEFAULT1) && (!ctx->cloud_map.valid || !ctx->detail.valid))
        return sky_fail(ctx, "cloud map / detail texture has not been generated");
    if (count == 0) return 0;
    PtParams P = make_pt_params(ctx, c);
    P.x0 = region[0]; P.y0 = region[1]; P.x1 = region[2]; P.y1 = region[3];
    P.frame_begin = frame_begin; P.frame_count = count;
    int rw = P.x1 - P.x0, rh = P.y1 - P.y0;
    if (rw <= 0 || rh <= 0) return 0;
    dim3 grid(ceil_div(rw, 16), ceil_div(rh, 8));
    const bool count_on = ctx->counting;
    const int prng = ctx->pt.prng;
    int rc = dispatch_material(ctx->material.type, ctx->hw_filtering, [&]<int MAT, bool HW>() {
        if (prng == SKY_PRNG_WANG) {
            if (count_on) k19_path_trace<MAT, HW, SKY_PRNG_WANG, true><<<grid, 128, 0, ctx->stream>>>(P);
            else k19_path_trace<MAT, HW, SKY_PRNG_WANG, false><<<grid, 128, 0, ctx->stream>>>(P);
        } else {
            if (count_on) k19_path_trace<MAT, HW, SKY_PRNG_PCG, true><<<grid, 128, 0, ctx->stream>>>(P);
            else k19_path_trace<MAT, HW, SKY_PRNG_PCG, false><<<grid, 128, 0, ctx->stream>>>(P);
        }
        return 0;
    });
    if (rc) return sky_fail(ctx, "unknown material");
    SKY_LAUNCH_CHECK(ctx);
    return 0;
}

int launch_pt_resolve(SkyContext* ctx, uint32_t frame_count, half4* hdr) {
    if (!ctx->pt_accum.p) return sky_fail(ctx, "pt_begin was not called");
    SkyCloudCommonBufferData c{};
    PtParams P = make_pt_params(ctx, c);
    P.hdr = hdr;
    P.frame_begin = frame_count;
    k20_display<<<dim3(ceil_div(P.width, 256), P.height), 256, 0, ctx->stream>>>(P);
    SKY_LAUNCH_CHECK(ctx);
    return 0;
}
