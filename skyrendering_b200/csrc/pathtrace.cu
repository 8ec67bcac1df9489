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
//
// Execution model (why this is not the shader's loop nest).  A path is ~5e3 tentative collisions, with
// the count varying by 1000x between neighbouring pixels, split between the free-flight loop (:182-196)
// and the shadow-ray loop (:142-149) of up to 128 bounces.  Run as written, lanes of a warp sit in
// different loops and ncu shows 2 of 32 lanes active.  Here every lane is a small state machine inside ONE
// loop whose hot block -- a single tentative collision, shared by free-flight and shadow tracking --
// is executed convergently by all tracking lanes; the rare transitions (scatter, shadow end, ground
// hit, path end) are side blocks.  Lanes are persistent and pull (pixel, frame) jobs from a global
// counter, so a finished path never idles its lane.  Each job writes its sample to its own slot and
// a second kernel adds the slots to the RGBA32F accumulator in kFrameId order, which keeps the
// reference's order of fp32 additions and makes the result independent of scheduling.
// Compiled twice like cloud.cu: -DSKY_STRICT_TU builds the unfused, IEEE-division objects behind sky_set_strict_arithmetic.
#ifdef SKY_STRICT_TU
#define launch_pt_samples launch_pt_samples_strict
#define launch_pt_resolve launch_pt_resolve_strict
#define SKY_K19_PRECISE_LOG
#endif
#include "../../include/sky_cubemap.h"
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
    float4* samples;             // [frame_count][region pixels]
    unsigned int* job_counter;
    unsigned long long* counters;
    int width, height;
    int x0, y0, x1, y1;  // kRenderRegion
    uint32_t frame_begin, frame_count;
    // voxel material, production objects: local position -> level-0 texel coordinate as one affine map per axis (make_pt_params)
    float vx_scale, vx_bias, vy_scale, vy_bias, vz_scale, vz_bias, vz_depth, v_collision_scale;
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

template <int PRNG_KIND>
SKY_D float Random01(uint32_t& seed) {  // :44-48 (returns, then advances; can round to 1.0)
    float res = float(seed) * (1.0f / 4294967296.0f);  // exact: power-of-two scale
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
SKY_D float2 CloudRegionIntersect(const PtParams& P, float3 ro, float3 rd) {
    const float hw = P.pt.region_box_half_width;
    const float bmin[3] = {-hw, -hw, P.c.uBottomAltitude}, bmax[3] = {hw, hw, P.c.uTopAltitude};
    const float o[3] = {ro.x, ro.y, ro.z}, d[3] = {rd.x, rd.y, rd.z};
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

// Exact empty-space test: for the voxel material with CLAMP_TO_BORDER(0), every tap of every mip level is
// the border when (u, v) is more than half a level-0 texel outside [0, 1], so sigma_t == 0 without a fetch.
template <int MAT>
SKY_D bool ProvablyEmpty(const PtParams& P, float3 pos) {
    if (MAT != SKY_MATERIAL_VOXEL) return false;
    const SkyMaterialVoxelBufferData& m = P.mat.m.u.voxel;
    float u = pos.x * m.uSampleFrequency[0] + m.uSampleBias[0];
    float v = pos.y * m.uSampleFrequency[1] + m.uSampleBias[1];
    float hu = 0.5f / float(P.mat.voxel.w[0]), hv = 0.5f / float(P.mat.voxel.h[0]);
    return u < -hu || u > 1.0f + hu || v < -hv || v > 1.0f + hv;
}

template <int MAT, bool HW>
SKY_D float SampleSigmaTAt(const PtParams& P, float3 pos, float inv_thickness) {  // :89-91
    float height01 = clampf((pos.z - P.c.uBottomAltitude) * inv_thickness, 0.0f, 1.0f);
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
    // seamless filtering at the face edges (GL 4.6 8.14.1; AtmosphereRenderer.cpp:151): include/sky_cubemap.h
    return sky_cube_bilinear<float3>(n, face, i0, j0, a, b, [&](int f, int i, int j) { return xyz(load_half4(P.env + (size_t(f) * n + j) * n + i)); });
}

SKY_D float3 GetSunIlluminance(const PtParams& P, float3 pos) {  // :131-133 + VolumetricCloudCommon.glsl:73-79
    float3 up_dir = f3(pos.x, pos.y, pos.z + P.c.uEarthRadius);
    float r = length(up_dir);
    up_dir /= r;
    float mu_s = dot(f3(P.c.uSunDirection), up_dir);
    return P.atm.GetSunVisibility(P.transmittance, r, mu_s) * P.atm.solar_illuminance();
}

// lane states
enum : int {
    ST_FETCH = 0,     // needs a (pixel, frame) job
    ST_SEGMENT,       // top of `while (istep < kMaxBounces && throughput > 0)`, :172
    ST_BEGIN_TRACK,   // a tracking loop is about to start: bound the part of the ray that can see the voxel footprint
    ST_TRACK,         // inside a tracking loop (free flight :182-196 or shadow ray :142-149)
    ST_EXIT_PRIMARY,  // free flight left the box without a collision, :198-230
    ST_SCATTER,       // real collision, :231-243
    ST_SHADOW_END,    // shadow ray finished: add the light sample, then sample the next direction
    ST_FINISH,        // path complete: aerial perspective, write the sample
    ST_IDLE           // no jobs left
};

// tuning knobs of the tracking loop (measured on B200, see profiles/): batches per trip around the state machine,
// collisions per batch, resident blocks per SM the register budget is cut for
#ifndef SKY_K19_TRACK_ROUNDS
#define SKY_K19_TRACK_ROUNDS 8
#endif
#ifndef SKY_K19_MIN_TRACKING
#define SKY_K19_MIN_TRACKING 1
#endif
#ifndef SKY_K19_BATCH
#define SKY_K19_BATCH 4
#endif
// Production objects: the texel coordinate of a collision is ONE fma per axis in the ray parameter -- the per-ray coefficients are formed once per
// trip of kBatch collisions -- instead of position -> (u, v, height01) -> texel (12 instructions -> 5), and the footprint interval and the
// box exit are one compare pair against [t_in, min(t_out, t_max)].  Same arithmetic up to rounding (the stream test's bounds hold); the
// strict objects keep the oracle's operation order.  Level 2 also carries sigma_t / sigma_t_max through the batch (one multiplication by
// uDensity / 255 / sigma_t_max instead of three).
#ifndef SKY_K19_FOLD
#define SKY_K19_FOLD 2   // measured (profiles/k19_fold_r02H.log, 1280x720 x 64 kFrameIds): 0 -> 83.7, 1 -> 88.2, 2 -> 90.0 Msamples/s; parity tests unchanged
#endif
#if defined(SKY_STRICT_TU) && SKY_K19_FOLD
#undef SKY_K19_FOLD
#define SKY_K19_FOLD 0
#endif
#ifndef SKY_K19_BRANCHLESS   // 1: resolution of a batch as selects; 2: the blends of a batch too
#define SKY_K19_BRANCHLESS 2   // measured (profiles/k19_branchless_r02K.log, 64 kFrameIds): 0 -> 91.9, 1 -> 98.7, 2 -> 100.0 Msamples/s, accumulators byte-identical
#endif
#if defined(SKY_STRICT_TU) && SKY_K19_BRANCHLESS
#undef SKY_K19_BRANCHLESS
#define SKY_K19_BRANCHLESS 0
#endif
#ifndef SKY_K19_OCC   // resident 128-thread blocks per SM.  With the folded hot block the production kernel fits 80 registers without spilling:
#ifdef SKY_STRICT_TU  // measured (profiles/k19_occ_r02J.log, 64 kFrameIds): 4 -> 82.4, 5 -> 89.8, 6 -> 91.6 Msamples/s
#define SKY_K19_OCC 5
#else
#define SKY_K19_OCC 6
#endif
#endif
constexpr int kTrackRounds = SKY_K19_TRACK_ROUNDS, kBatch = SKY_K19_BATCH;
// free-flight logarithm: logf (<= 1 ulp) or the MUFU.LG2-based __logf (what GLSL's log() compiles to on this hardware)
// Measured: 20.8 -> 23.9 Msamples/s with __logf, parity against the oracle unchanged (relative RMS of the 16-spp
// accumulator 2.16e-3 vs 2.13e-3, same 71 % bit-identical pixels): the default.
#ifdef SKY_K19_PRECISE_LOG
#define SKY_K19_LOG(x) logf(x)
#else
#define SKY_K19_LOG(x) __logf(x)
#endif

// K19 -- :160-284 as a per-lane state machine; see the header of this file.
template <int MAT, bool HW, int PRNG_KIND, bool COUNT>
__global__ void __launch_bounds__(128, SKY_K19_OCC) k19_path_trace(const __grid_constant__ PtParams P) {
    const int rw = P.x1 - P.x0, rh = P.y1 - P.y0;
    const int tiles_x = (rw + 7) >> 3, tiles_y = (rh + 3) >> 2;  // jobs walk the region in 8x4 pixel tiles so a warp starts on one tile
    const unsigned int npix = (unsigned int)rw * (unsigned int)rh;                  // sample-slot stride per frame
    const unsigned int npix_padded = (unsigned int)tiles_x * (unsigned int)tiles_y * 32u;  // job stride per frame
    const unsigned int njobs = npix_padded * P.frame_count;
    const float3 camera = f3(P.c.uCameraPos);
    const float3 sun = f3(P.c.uSunDirection);
    const float sigma_t_max = P.pt.sigma_t_max;
    const float inv_sigma_t_max = 1.0f / sigma_t_max;
    const float inv_thickness = 1.0f / (P.c.uTopAltitude - P.c.uBottomAltitude);
    const unsigned int lane = threadIdx.x & 31u;
    // voxel footprint in local x/y, enlarged (see ST_BEGIN_TRACK); other materials fill all of space
    float fp_x_lo = -INFINITY, fp_x_hi = INFINITY, fp_y_lo = -INFINITY, fp_y_hi = INFINITY;
    if (MAT == SKY_MATERIAL_VOXEL) {
        const SkyMaterialVoxelBufferData& vm = P.mat.m.u.voxel;
        // 0.2 texel beyond the half-texel apron: far more than the fp32 error of a position along the ray, and inside the
        // quarter texel the padded cell layout allows without a range test (MipView::cells, voxel_tap)
        float hu = 0.7f / float(P.mat.voxel.w[0]), hv = 0.7f / float(P.mat.voxel.h[0]);
        fp_x_lo = (-hu - vm.uSampleBias[0]) / vm.uSampleFrequency[0];
        fp_x_hi = (1.0f + hu - vm.uSampleBias[0]) / vm.uSampleFrequency[0];
        fp_y_lo = (-hv - vm.uSampleBias[1]) / vm.uSampleFrequency[1];
        fp_y_hi = (1.0f + hv - vm.uSampleBias[1]) / vm.uSampleFrequency[1];
    }
    float t_in = -INFINITY, t_out = INFINITY;
    bool ray_magnified = false;  // every lookup of the current tracking ray is a level-0 LINEAR fetch (lambda <= 0.5)

    int state = ST_FETCH;
    unsigned int job = 0;
    int px = 0, py = 0;
    uint32_t seed = 0, saved_seed = 0;
    float3 ro = f3(0.0f), rd = f3(0.0f, 0.0f, 1.0f);  // ctx.ray
    float3 L = f3(0.0f), throughput = f3(1.0f), light = f3(0.0f), bsdf = f3(0.0f);
    float t = 0.0f, t_max = 0.0f, transmittance = 1.0f, scattered_t = 0.0f;
    int istep = 0;
    bool has_scattered = false, in_shadow = false, after_ground = false;
    int lookups = 0, collisions = 0, paths = 0;
#ifdef SKY_K19_PROBE  // experiment build only: per-path collision / cycle maxima and the drain phase of the launch
    unsigned long long probe_coll = 0, probe_t0 = 0;
    auto probe_now = [] { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; };
    if (threadIdx.x == 0 && blockIdx.x == 0) atomicMax(P.counters + 5, (1ull << 62) - probe_now());
#endif

    for (;;) {
        // ---------------------------------------------------------------- job fetch (warp-aggregated)
        {
            unsigned int need = __ballot_sync(0xffffffffu, state == ST_FETCH);
            if (need) {
                unsigned int base = 0;
                int leader = __ffs(need) - 1;
                if (lane == (unsigned int)leader) base = atomicAdd(P.job_counter, (unsigned int)__popc(need));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (state == ST_FETCH) {
                    job = base + __popc(need & ((1u << lane) - 1u));
                    if (job >= njobs) {
                        state = ST_IDLE;
#ifdef SKY_K19_PROBE
                        atomicMax(P.counters + 0, (1ull << 62) - probe_now());
#endif
                    } else {
                        unsigned int frame_index = job / npix_padded, p = job - frame_index * npix_padded;
                        // 8x4 tile order inside the region
                        unsigned int tile = p >> 5, in_tile = p & 31u;
                        int tx = int(tile % (unsigned int)tiles_x), ty = int(tile / (unsigned int)tiles_x);
                        px = P.x0 + tx * 8 + int(in_tile & 7u);
                        py = P.y0 + ty * 4 + int(in_tile >> 3);
                        if (px >= P.x1 || py >= P.y1) {
                            state = ST_FETCH;  // padding of the tile grid: take another job next round
                        } else {
                            seed = PRNG<PRNG_KIND>(PRNG<PRNG_KIND>(PRNG<PRNG_KIND>(uint32_t(px)) + uint32_t(py)) + (P.frame_begin + frame_index));  // :260
                            float2 uv = f2((float(px) + 0.5f) / float(P.width), (float(py) + 0.5f) / float(P.height));
                            float3 frag_pos = projective_mul(P.c.uInvMVP, f3(uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f, 1.0f));
                            rd = normalize(frag_pos - camera);
                            ro = camera;
                            L = f3(0.0f); throughput = f3(1.0f);
                            has_scattered = false; in_shadow = false; istep = 0; scattered_t = 0.0f;
                            if (COUNT) ++paths;
#ifdef SKY_K19_PROBE
                            probe_coll = 0; probe_t0 = probe_now();
#endif
                            float2 camera_inter_t = CloudRegionIntersect(P, ro, rd);  // :166-170
                            if (camera_inter_t.x >= camera_inter_t.y) {
                                state = ST_FINISH;
                            } else {
                                ro += camera_inter_t.x * rd;
                                state = ST_SEGMENT;
                            }
                        }
                    }
                }
            }
        }
        if (__all_sync(0xffffffffu, state == ST_IDLE)) break;

        // ---------------------------------------------------------------- :172-181
        if (state == ST_SEGMENT) {
            if (!(istep < P.pt.max_bounces && fmaxf(throughput.x, fmaxf(throughput.y, throughput.z)) > 0.0f)) {
                state = ST_FINISH;
            } else {
                float2 inter_t = CloudRegionIntersect(P, ro, rd);
                if (inter_t.x >= inter_t.y) {
                    state = ST_FINISH;
                } else {
                    t = inter_t.x; t_max = inter_t.y;
                    in_shadow = false;
                    state = sigma_t_max <= 0 ? ST_EXIT_PRIMARY : ST_BEGIN_TRACK;  // :183
                }
            }
        }

        // ---------------------------------------------------------------- footprint interval of the new tracking ray
        // Conservative t-range in which (u, v) can lie inside the voxel footprint enlarged by 1e-3 (the exact test
        // needs half a texel; fp32 error of the ray parameter is ~1e-6): outside it a collision is provably empty
        // from two compares, without forming the position.
        if (state == ST_BEGIN_TRACK) {
            if (MAT == SKY_MATERIAL_VOXEL) {
                const float3 d = in_shadow ? sun : rd;
                float ix = 1.0f / d.x, iy = 1.0f / d.y;
                float ta = (fp_x_lo - ro.x) * ix, tb = (fp_x_hi - ro.x) * ix;
                float tc = (fp_y_lo - ro.y) * iy, td = (fp_y_hi - ro.y) * iy;
                t_in = fmaxf(fminf(ta, tb), fminf(tc, td));
                t_out = fminf(fmaxf(ta, tb), fmaxf(tc, td));
                // Dead-stream cut (exact): past t_out every remaining collision of this ray is a null collision, so
                // the tail can only matter through the random numbers it consumes.  A shadow ray works on a COPY of
                // the stream (:135) and its transmittance is final; a free flight that will leave the box towards the
                // sky (or before any scattering) ends the path without drawing again (:198-216).  In both cases the
                // first collision past t_out may end the ray.  Only a free flight that will reach the ground keeps its
                // full chain: the Lambert bounce continues the stream.  The exit test reads the segment origin and
                // direction only (:207-212), so it is known here.  The counting variant keeps every collision: its
                // totals are the reference algorithm's.
                if (!COUNT) {
                    bool stream_is_dead = in_shadow;
                    if (!in_shadow) {
                        bool ground_bounce = false;
                        if (P.pt.environment_lighting != SKY_ENV_OFF && P.pt.environment_lighting != SKY_ENV_CONST_ENVIRONMENT_MAP && has_scattered) {
                            float3 up_dir = f3(ro.x, ro.y, ro.z + P.c.uEarthRadius);
                            float r = length(up_dir);
                            up_dir /= r;
                            ground_bounce = P.atm.RayIntersectsGround(r, dot(rd, up_dir));
                        }
                        stream_is_dead = !ground_bounce;
                    }
                    if (stream_is_dead) t_max = fminf(t_max, t_out);  // NaN t_out (axis-parallel ray on a footprint edge) leaves t_max alone
                }
                // lambda <= 0.5 <=> dist(pos, camera)^2 <= thr2; the distance to a point is convex along a ray, so the two
                // ends of the part of the ray that can see the grid decide for all of it (0.01 % margin for rounding)
                float ta_ = fmaxf(t, t_in), tb_ = fminf(t_max, t_out);
                float thr2 = P.mat.thr2_voxel * 0.9999f;
                ray_magnified = distance2(ro + d * ta_, P.mat.camera_pos) <= thr2 && distance2(ro + d * tb_, P.mat.camera_pos) <= thr2;
            }
            state = ST_TRACK;
        }

        // ---------------------------------------------------------------- hot block: tentative collisions
        // kTrackRounds rounds of one BATCH of up to kBatch collisions per tracking lane, in three convergent phases:
        //  (1) the collision distances.  They depend on the random stream only -- never on the density -- so the
        //      batch is generated ahead of its lookups: a shadow ray never branches on sigma_t (:148), a free flight
        //      only ends on a real collision (:191-195), and then the speculated rest of the batch is dropped and the
        //      stream rewound to that collision;
        //  (2) the density lookups of the batch, issued together (kBatch independent 8-byte loads in flight per lane
        //      instead of a load -> blend -> compare -> next load chain; this is what bounds a lone long path);
        //  (3) the in-order resolution: transmittance product (reference order) or the scatter test.
        // Collisions that provably land on zero density (outside the footprint interval) resolve with sigma_t = 0
        // without a lookup.  Divisions by the constant majorant are multiplications by its reciprocal.
#pragma unroll 1
        for (int round = 0; round < kTrackRounds; ++round) {
#if SKY_K19_MIN_TRACKING > 1
            // leave the hot block once fewer than SKY_K19_MIN_TRACKING lanes still track: the others idle here until the state
            // transitions below run (results do not depend on it: every path consumes its own stream)
            if (__popc(__ballot_sync(0xffffffffu, state == ST_TRACK)) < (round == 0 ? 1 : SKY_K19_MIN_TRACKING)) break;
#else
            if (!__any_sync(0xffffffffu, state == ST_TRACK)) break;
#endif
            if (state == ST_TRACK) {
                const float3 dir = in_shadow ? sun : rd;
                // (1) branch-free: the next kBatch collision distances and stream positions.  Entries past the end of
                //     the ray (tk > t_max) are computed and ignored.
                float tk[kBatch];
                uint32_t sk[kBatch];  // stream position after the free-flight draw of collision k: xi_k = float(sk[k]) / 2^32
                uint32_t s = seed;
#pragma unroll
                for (int k = 0; k < kBatch; ++k) {
                    float step = -SKY_K19_LOG(1.0f - Random01<PRNG_KIND>(s)) * inv_sigma_t_max;  // InfiniteTransmittanceIS, :82-84
                    tk[k] = (k ? tk[k - 1] : t) + step;
                    sk[k] = s;
                    if (!in_shadow) s = PRNG<PRNG_KIND>(s);  // the xi draw of :191
                }
                // (2) the lookups
                float sig[kBatch];
                bool live[kBatch];
#if SKY_K19_FOLD
                // NaN t_out (ray inside a footprint edge plane): nothing is live; fminf alone would drop the NaN
                const float t_hi = t_out == t_out ? fminf(t_out, t_max) : -INFINITY;
#pragma unroll
                for (int k = 0; k < kBatch; ++k) live[k] = tk[k] >= t_in && tk[k] <= t_hi;
#else
#pragma unroll
                for (int k = 0; k < kBatch; ++k) live[k] = tk[k] >= t_in && tk[k] <= t_out && tk[k] <= t_max;  // NaN bounds (ray inside a footprint edge plane): not live, and sigma_t is 0 there
#endif
                if (MAT == SKY_MATERIAL_VOXEL && !HW && ray_magnified) {
                    const SkyMaterialVoxelBufferData& vm = P.mat.m.u.voxel;
                    VoxelTap tap[kBatch];
#if SKY_K19_FOLD
                    const float x0 = fmaf(ro.x, P.vx_scale, P.vx_bias), dx = dir.x * P.vx_scale;
                    const float y0 = fmaf(ro.y, P.vy_scale, P.vy_bias), dy = dir.y * P.vy_scale;
                    const float h0 = fmaf(ro.z, P.vz_scale, P.vz_bias), dh = dir.z * P.vz_scale;
#pragma unroll
                    for (int k = 0; k < kBatch; ++k)
                        tap[k] = voxel_tap_texel(P.mat.voxel, fmaf(dx, tk[k], x0), fmaf(dy, tk[k], y0), fmaf(__saturatef(fmaf(dh, tk[k], h0)), P.vz_depth, -0.5f));
#else
#pragma unroll
                    for (int k = 0; k < kBatch; ++k) {
                        float3 pos = ro + dir * tk[k];
                        float height01 = clampf((pos.z - P.c.uBottomAltitude) * inv_thickness, 0.0f, 1.0f);
                        float u = pos.x * vm.uSampleFrequency[0] + vm.uSampleBias[0];
                        float v = pos.y * vm.uSampleFrequency[1] + vm.uSampleBias[1];
                        tap[k] = voxel_tap(P.mat.voxel, u, v, height01);  // live: inside the padded cell range, no test
                    }
#endif
                    uint2 cell[kBatch];
#pragma unroll
                    for (int k = 0; k < kBatch; ++k) cell[k] = voxel_tap_load(P.mat.voxel, tap[k], live[k]);
#pragma unroll
#if SKY_K19_FOLD >= 2   // sig[] holds sigma_t / sigma_t_max directly: one multiplication for (1/255) * uDensity / sigma_t_max
#if SKY_K19_BRANCHLESS >= 2
                    // every lane blends the cell it loaded (a lane that is not live loaded cell 0) and the result is selected: four reconvergence
                    // regions fewer per batch; a warp in which NO lane is live at position k is rare once paths have desynchronised
                    for (int k = 0; k < kBatch; ++k) {
                        float blended = blend_cell_raw(cell[k], tap[k].a, tap[k].b, tap[k].c) * P.v_collision_scale;
                        asm volatile("" : "+f"(blended));   // keep the blend out of a conditional arm
                        sig[k] = live[k] ? blended : 0.0f;
                    }
#else
                    for (int k = 0; k < kBatch; ++k) sig[k] = live[k] ? blend_cell_raw(cell[k], tap[k].a, tap[k].b, tap[k].c) * P.v_collision_scale : 0.0f;
#endif
#else
                    for (int k = 0; k < kBatch; ++k) sig[k] = live[k] ? blend_cell(cell[k], tap[k].a, tap[k].b, tap[k].c) * vm.uDensity : 0.0f;
#endif
                } else {  // other materials, hardware filtering, rays reaching minified (NEAREST mip level) distances: the general sampler
#pragma unroll
                    for (int k = 0; k < kBatch; ++k) {
                        float3 pos = ro + dir * tk[k];
                        live[k] = live[k] && !ProvablyEmpty<MAT>(P, pos);
                        sig[k] = live[k] ? SampleSigmaTAt<MAT, HW>(P, pos, inv_thickness) : 0.0f;
#if SKY_K19_FOLD >= 2
                        sig[k] *= inv_sigma_t_max;
#endif
                    }
                }
#if SKY_K19_FOLD >= 2
                constexpr bool kSigIsProbability = true;
#else
                constexpr bool kSigIsProbability = false;
#endif
                const float to_probability = kSigIsProbability ? 1.0f : inv_sigma_t_max;
                // (3) in-order resolution
#if SKY_K19_BRANCHLESS && !defined(SKY_K19_PROBE)
                // Lanes of a warp are a mix of shadow rays and free flights almost all the time, so the nested `if`s below cost a warp BOTH arms for
                // every collision (~30 instructions and four reconvergence regions per collision).  The same decisions as selects: every collision
                // updates the running product under a predicate and remembers the first event (box exit or real collision) of the batch; the one
                // divergent block that acts on the event runs once per batch.  Same operations on the same operands: the accumulator is bit-identical.
                if (!COUNT) {
                    bool alive = true, event_seen = false, event_is_exit = false;
                    uint32_t event_s = s;
                    float event_t = tk[kBatch - 1];
#pragma unroll
                    for (int k = 0; k < kBatch; ++k) {
                        const float pk = kSigIsProbability ? sig[k] : sig[k] * to_probability;
                        const bool out = tk[k] > t_max;
                        const bool hit = float(sk[k]) * (1.0f / 4294967296.0f) < pk;                      // :191-195
                        const float factor = 1.0f - fmaxf(0.0f, pk);                                     // :148
                        // (bitwise, not short-circuit, operators: the compiler must not turn the tests back into branches)
                        transmittance = (alive & !out & in_shadow) ? transmittance * factor : transmittance;
                        const bool event = alive & (out | (!in_shadow & hit));
                        event_is_exit = event ? out : event_is_exit;
                        event_s = event ? sk[k] : event_s;
                        event_t = event ? tk[k] : event_t;
                        event_seen = event_seen | event;
                        alive = alive & !event;
                    }
                    if (!event_seen) {
                        t = tk[kBatch - 1]; seed = s;
                    } else if (event_is_exit) {   // the ray left the box; the stream stands after this step's draw
                        state = in_shadow ? ST_SHADOW_END : ST_EXIT_PRIMARY;
                        seed = event_s;
                    } else {
                        state = ST_SCATTER;
                        t = event_t;
                        seed = PRNG<PRNG_KIND>(event_s);
                    }
                } else
#endif
                {
#pragma unroll
                for (int k = 0; k < kBatch; ++k) {
                    if (state == ST_TRACK) {
                        if (tk[k] > t_max) {  // the ray left the box; the stream stands after this step's draw
                            state = in_shadow ? ST_SHADOW_END : ST_EXIT_PRIMARY;
                            seed = sk[k];
                        } else {
                            if (COUNT) { ++collisions; lookups += live[k] ? 1 : 0; }
#ifdef SKY_K19_PROBE
                            ++probe_coll;
#endif
                            if (in_shadow) {
                                transmittance *= 1.0f - fmaxf(0.0f, kSigIsProbability ? sig[k] : sig[k] * to_probability);  // :148 (sigma_t == 0: times one, exactly)
                            } else if (float(sk[k]) * (1.0f / 4294967296.0f) < (kSigIsProbability ? sig[k] : sig[k] * to_probability)) {  // :191-195
                                state = ST_SCATTER;
                                t = tk[k];
                                seed = PRNG<PRNG_KIND>(sk[k]);
                            }
                        }
                    }
                }
                if (state == ST_TRACK) { t = tk[kBatch - 1]; seed = s; }
                }
#ifndef SKY_K19_NO_ZERO_CUT
                // Zero-transmittance cut (exact): once the running product of a shadow ray is exactly 0 -- a texel at the
                // majorant, or the underflow of ~10^2 collisions deep inside the cloud -- every further factor multiplies
                // zero and the stream the ray consumes is a copy (:135), so the ray may end here.  (The counting variant
                // walks the whole chain: its totals are the reference algorithm's.)
                if (!COUNT && state == ST_TRACK && in_shadow && transmittance == 0.0f) state = ST_SHADOW_END;
#endif
            }
        }

        // ---------------------------------------------------------------- :198-230
        if (state == ST_EXIT_PRIMARY) {
            state = ST_FINISH;
            if (P.pt.environment_lighting != SKY_ENV_OFF && has_scattered) {
                const float* mm = P.pt.model_matrix3;
                float3 world_dir = f3(mm[0] * rd.x + mm[3] * rd.y + mm[6] * rd.z, mm[1] * rd.x + mm[4] * rd.y + mm[7] * rd.z,
                                      mm[2] * rd.x + mm[5] * rd.y + mm[8] * rd.z);
                if (P.pt.environment_lighting == SKY_ENV_CONST_ENVIRONMENT_MAP) {
                    L += throughput * SampleEnvironment(P, world_dir);
                } else {
                    float3 up_dir = f3(ro.x, ro.y, ro.z + P.c.uEarthRadius);
                    float r = length(up_dir);
                    up_dir /= r;
                    float mu = dot(rd, up_dir);
                    if (!P.atm.RayIntersectsGround(r, mu)) {
                        L += throughput * SampleEnvironment(P, world_dir);
                    } else {
                        ro += rd * P.atm.DistanceToBottomAtmosphereBoundary(r, mu);
                        float3 ground_normal = normalize(f3(ro.x, ro.y, ro.z + P.c.uEarthRadius));
                        float NdotL = dot(ground_normal, sun);
                        light = GetSunIlluminance(P, ro);                     // :131-133
                        bsdf = (kInvPi * P.atm.ground_albedo()) * NdotL;       // :220-222
                        after_ground = true;
                        // TransmittanceEstimation(ctx BY VALUE, ray(pos, sun)), :135-141
                        saved_seed = seed;
                        transmittance = 1.0f;
                        float2 st = CloudRegionIntersect(P, ro, sun);
                        if (st.x >= st.y) {
                            state = ST_SHADOW_END;
                        } else {
                            t = st.x; t_max = st.y; in_shadow = true;
                            state = ST_BEGIN_TRACK;
                        }
                    }
                }
            }
        }

        // ---------------------------------------------------------------- :231-238
        if (state == ST_SCATTER) {
            if (!has_scattered) scattered_t = distance(camera, ro);
            has_scattered = true;
            ro += rd * t;
            light = GetSunIlluminance(P, ro);
            bsdf = f3(GetPhase(P, dot(rd, sun)));  // :237
            after_ground = false;
            saved_seed = seed;
            transmittance = 1.0f;
            float2 st = CloudRegionIntersect(P, ro, sun);
            if (st.x >= st.y) {
                state = ST_SHADOW_END;
            } else {
                t = st.x; t_max = st.y; in_shadow = true;
                state = ST_BEGIN_TRACK;
            }
        }

        // ---------------------------------------------------------------- :150, :157, then :224-230 or :240-242
        if (state == ST_SHADOW_END) {
            seed = saved_seed;  // the shadow ray consumed a copy of the stream
            in_shadow = false;
            L += throughput * (clampf(transmittance, 0.0f, 1.0f) * light * bsdf);  // :157, same association as the shader
            state = ST_SEGMENT;
            if (after_ground) {
                if (P.pt.environment_lighting == SKY_ENV_GROUND_SINGLE_BOUNCE) {
                    state = ST_FINISH;
                } else {
                    // GenerateLambertSample, :119-129
                    float3 ground_normal = normalize(f3(ro.x, ro.y, ro.z + P.c.uEarthRadius));
                    float sin_theta = sqrtf(Random01<PRNG_KIND>(seed));
                    float cos_theta = sqrtf(clampf(1.0f - sin_theta * sin_theta, 0.0f, 1.0f));
                    float3 t0, t1;
                    CreateOrthonormalBasis(ground_normal, t0, t1);
                    float phi = 2.0f * kPi * Random01<PRNG_KIND>(seed);
                    rd = sin_theta * sinf(phi) * t0 + sin_theta * cosf(phi) * t1 + cos_theta * ground_normal;
                    throughput *= P.atm.ground_albedo();
                }
            } else if (P.pt.importance_sampling) {
                // GenerateHGSample, :98-111
                float g = Random01<PRNG_KIND>(seed) < P.pt.forward_scattering_ratio ? P.pt.forward_phase_g : P.pt.back_phase_g;
                float cos_theta = HenyeyGreensteinInvertcdf(Random01<PRNG_KIND>(seed), g);
                float sin_theta = sqrtf(clampf(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
                float3 t0, t1;
                CreateOrthonormalBasis(rd, t0, t1);
                float phi = 2.0f * kPi * Random01<PRNG_KIND>(seed);
                rd = sin_theta * sinf(phi) * t0 + sin_theta * cosf(phi) * t1 + cos_theta * rd;
            } else {
                // UniformSphereSample, :50-55, :112-115
                float phi = 2.0f * kPi * Random01<PRNG_KIND>(seed);
                float cos_theta = 1.0f - 2.0f * Random01<PRNG_KIND>(seed);
                float sin_theta = sqrtf(clampf(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
                float3 direction = f3(cosf(phi) * sin_theta, sinf(phi) * sin_theta, cos_theta);
                float value = GetPhase(P, dot(rd, direction));
                rd = direction;
                throughput *= value / (1.0f / (4.0f * kPi));
            }
            ++istep;
        }

        // ---------------------------------------------------------------- :248, :263-283
        if (state == ST_FINISH) {
            float4 this_res = f4(L, has_scattered ? 0.0f : 1.0f);
            if (has_scattered) {
                float2 uv = f2((float(px) + 0.5f) / float(P.width), (float(py) + 0.5f) / float(P.height));
                float3 frag_pos = projective_mul(P.c.uInvMVP, f3(uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f, 1.0f));
                float3 view_dir = normalize(frag_pos - camera);
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
                this_res = f4(xyz(this_res) * atmosphere_transmittance + atmosphere_luminance, this_res.w);
            }
            unsigned int frame_index = job / npix_padded;
            P.samples[size_t(frame_index) * npix + size_t(py - P.y0) * rw + (px - P.x0)] = this_res;
            state = ST_FETCH;
#ifdef SKY_K19_PROBE
            {
                unsigned long long dur_us = (probe_now() - probe_t0) / 1000ull;
                unsigned long long coll = probe_coll < (1ull << 28) ? probe_coll : (1ull << 28) - 1;
                atomicMax(P.counters + 6, (dur_us << 40) | (coll << 12) | ((unsigned long long)(istep & 0xff) << 4));
                atomicMax(P.counters + 7, (dur_us << 40) | ((unsigned long long)px << 20) | (unsigned long long)py);
                atomicMax(P.counters + 2, (coll << 32) | ((unsigned long long)px << 16) | (unsigned long long)py);
                atomicAdd(P.counters + 4, probe_coll);
            }
#endif
        }
    }
#ifdef SKY_K19_PROBE
    if (lane == 0) atomicMax(P.counters + 1, probe_now());
#endif
    if (COUNT) {
        atomicAdd(P.counters + SKY_CNT_PT_PATHS, (unsigned long long)paths);
        atomicAdd(P.counters + SKY_CNT_PT_LOOKUPS, (unsigned long long)lookups);
        atomicAdd(P.counters + SKY_CNT_PT_COLLISIONS, (unsigned long long)collisions);
    }
}

// ------------------------------------------------------------------------------------------------ majorant-grid mode
// the same estimator (delta tracking for the free flight, ratio tracking for the sun transmittance, NEE at
// every vertex, VolumetricCloudPathTracing.comp:135-249) with LOCAL majorants instead of the reference's single global
// kSigmaTMax over the +-100 km box: a coarse grid over the voxel texture holds, per 8^3-texel macro cell, the largest
// density any lookup inside the cell can return (level 0 with its bilinear apron and every mip level the LOD rule may pick),
// and tracking walks that grid with a DDA, restarting the exponential at each cell boundary (memorylessness keeps it
// unbiased).  Cells with majorant 0 -- 3/4 of the data set, and all of space outside the footprint -- are crossed without a
// lookup.  The image has the same expectation but NOT the reference's random streams: it is validated statistically against
// the stream-exact kernel (tests) and reported separately (bench.py: "majorant_grid").
constexpr int kMacro = 8;

__global__ void __launch_bounds__(128) k_majorant_build(const MipView vox, uint8_t* __restrict__ out, int gw, int gh, int gd) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= gw * gh * gd) return;
    const int cx = c % gw, cy = (c / gw) % gh, cz = c / (gw * gh);
    const float W0 = float(vox.w[0]), H0 = float(vox.h[0]), D0 = float(vox.d[0]);
    // normalised extent of the macro cell (the last one may be partial)
    const float u0 = float(cx * kMacro) / W0, u1 = fminf(float(cx * kMacro + kMacro) / W0, 1.0f);
    const float v0 = float(cy * kMacro) / H0, v1 = fminf(float(cy * kMacro + kMacro) / H0, 1.0f);
    const float w0 = float(cz * kMacro) / D0, w1 = fminf(float(cz * kMacro + kMacro) / D0, 1.0f);
    int m = 0;
    for (int l = 0; l < vox.levels; ++l) {
        const int wl = vox.w[l], hl = vox.h[l], dl = vox.d[l];
        // texels a lookup with coordinates inside the cell can read at this level: LINEAR reaches half a texel beyond,
        // NEAREST stays inside; one more texel on each side absorbs the fp32 rounding of u * size
        const int i0 = max(int(floorf(u0 * float(wl) - 0.5f)) - 1, 0), i1 = min(int(floorf(u1 * float(wl) - 0.5f)) + 2, wl - 1);
        const int j0 = max(int(floorf(v0 * float(hl) - 0.5f)) - 1, 0), j1 = min(int(floorf(v1 * float(hl) - 0.5f)) + 2, hl - 1);
        const int k0 = max(int(floorf(w0 * float(dl) - 0.5f)) - 1, 0), k1 = min(int(floorf(w1 * float(dl) - 0.5f)) + 2, dl - 1);
        const uint8_t* base = vox.base + vox.off[l];
        for (int k = k0; k <= k1; ++k)
            for (int j = j0; j <= j1; ++j)
                for (int i = i0; i <= i1; ++i) m = max(m, int(base[(size_t(k) * hl + j) * wl + i]));
    }
    out[c] = uint8_t(m);
}

struct MajorantView {
    const uint8_t* p;
    int w, h, d;
};

// One tracking ray through the majorant grid over [t0, t1] (already inside the region box).  RATIO: ratio tracking, returns
// the transmittance estimate in `transmittance`, never "hits"; otherwise delta tracking, returns true with the collision
// distance in t_hit.
template <bool RATIO, int PRNG_KIND>
SKY_D bool TrackMajorantGrid(const PtParams& P, const MajorantView& G, float3 ro, float3 rd, float t0, float t1, float inv_thickness,
                             uint32_t& seed, float& t_hit, float& transmittance, int& lookups) {
    transmittance = 1.0f;
    const SkyMaterialVoxelBufferData& vm = P.mat.m.u.voxel;
    const float W0 = float(P.mat.voxel.w[0]), H0 = float(P.mat.voxel.h[0]), D0 = float(P.mat.voxel.d[0]);
    // (u, v, w)(t) = a + b t; outside u, v in [-1/2W, 1 + 1/2W] every tap of every level is the border (sigma_t = 0)
    const float au = ro.x * vm.uSampleFrequency[0] + vm.uSampleBias[0], bu = rd.x * vm.uSampleFrequency[0];
    const float av = ro.y * vm.uSampleFrequency[1] + vm.uSampleBias[1], bv = rd.y * vm.uSampleFrequency[1];
    const float aw = (ro.z - P.c.uBottomAltitude) * inv_thickness, bw = rd.z * inv_thickness;
    {
        const float hu = 0.75f / W0, hv = 0.75f / H0;
        float iu = 1.0f / bu, iv = 1.0f / bv;
        float ta = (-hu - au) * iu, tb = (1.0f + hu - au) * iu, tc = (-hv - av) * iv, td = (1.0f + hv - av) * iv;
        if (!(fabsf(bu) > 1e-30f)) { bool in = au >= -hu && au <= 1.0f + hu; ta = in ? -INFINITY : INFINITY; tb = in ? INFINITY : -INFINITY; }
        if (!(fabsf(bv) > 1e-30f)) { bool in = av >= -hv && av <= 1.0f + hv; tc = in ? -INFINITY : INFINITY; td = in ? INFINITY : -INFINITY; }
        t0 = fmaxf(t0, fmaxf(fminf(ta, tb), fminf(tc, td)));
        t1 = fminf(t1, fminf(fmaxf(ta, tb), fmaxf(tc, td)));
    }
    if (!(t0 < t1)) return false;
    // macro-grid coordinates g(t) = g0 + gd t
    const float sx = W0 / float(kMacro), sy = H0 / float(kMacro), sz = D0 / float(kMacro);
    const float g0x = au * sx, g0y = av * sy, g0z = aw * sz, gdx = bu * sx, gdy = bv * sy, gdz = bw * sz;
    const float igx = fabsf(gdx) > 1e-30f ? 1.0f / gdx : 0.0f, igy = fabsf(gdy) > 1e-30f ? 1.0f / gdy : 0.0f, igz = fabsf(gdz) > 1e-30f ? 1.0f / gdz : 0.0f;
    const int stx = gdx > 0.0f ? 1 : -1, sty = gdy > 0.0f ? 1 : -1, stz = gdz > 0.0f ? 1 : -1;
    float t = t0;
    // the first cell is found from a point just inside the interval
    const float te = fminf(t0 + 1e-5f * fmaxf(1.0f, fabsf(t0)), t1);
    int cx = clampi(int(floorf(g0x + gdx * te)), 0, G.w - 1), cy = clampi(int(floorf(g0y + gdy * te)), 0, G.h - 1),
        cz = clampi(int(floorf(g0z + gdz * te)), 0, G.d - 1);
    const float density = vm.uDensity * (1.0f / 255.0f);
    for (int guard = 0; guard < 4096; ++guard) {
        // exit of this macro cell: the next grid plane on each axis, unless that plane is the outside of the grid (the apron
        // beyond the last cell belongs to it; the ray then ends at t1)
        const int nx = cx + stx, ny = cy + sty, nz = cz + stz;
        float tx = (igx != 0.0f && nx >= 0 && nx < G.w) ? (float(stx > 0 ? cx + 1 : cx) - g0x) * igx : INFINITY;
        float ty = (igy != 0.0f && ny >= 0 && ny < G.h) ? (float(sty > 0 ? cy + 1 : cy) - g0y) * igy : INFINITY;
        float tz = (igz != 0.0f && nz >= 0 && nz < G.d) ? (float(stz > 0 ? cz + 1 : cz) - g0z) * igz : INFINITY;
        float t_exit = fminf(fminf(tx, ty), fminf(tz, t1));
        const float mu = float(__ldg(G.p + (size_t(cz) * G.h + cy) * G.w + cx)) * density;
        if (mu > 0.0f) {
            const float inv_mu = 1.0f / mu;
            for (;;) {
                t += -SKY_K19_LOG(1.0f - Random01<PRNG_KIND>(seed)) * inv_mu;   // InfiniteTransmittanceIS, :82-84
                if (!(t < t_exit)) break;
                float sigma_t = SampleSigmaTAt<SKY_MATERIAL_VOXEL, false>(P, ro + rd * t, inv_thickness);
                ++lookups;
                if (RATIO) {
                    transmittance *= 1.0f - fmaxf(0.0f, sigma_t * inv_mu);   // :148
                    if (transmittance <= 0.0f) return false;
                } else if (Random01<PRNG_KIND>(seed) < sigma_t * inv_mu) {          // :193
                    t_hit = t;
                    return true;
                }
            }
        }
        if (!(t_exit < t1)) return false;
        t = t_exit;
        if (tx <= ty && tx <= tz) cx = nx; else if (ty <= tz) cy = ny; else cz = nz;
    }
    return false;
}

// K19, majorant-grid mode: one lane per (pixel, block of 8 kFrameIds); VolumetricCloudPathTracing.comp:160-284 as written
// (the loop nest is fine here: a path is ~10^2 lookups, not ~5 10^3).
template <int PRNG_KIND, bool COUNT>
__global__ void __launch_bounds__(128) k19_majorant_grid(const __grid_constant__ PtParams P, const MajorantView G) {
    const int rw = P.x1 - P.x0, rh = P.y1 - P.y0;
    const int tiles_x = (rw + 7) >> 3, tiles_y = (rh + 3) >> 2;
    const unsigned int npix = (unsigned int)rw * (unsigned int)rh;
    const unsigned int npix_padded = (unsigned int)tiles_x * (unsigned int)tiles_y * 32u;
    constexpr unsigned int kFramesPerJob = 8;
    const unsigned int frame_blocks = (P.frame_count + kFramesPerJob - 1) / kFramesPerJob;
    const unsigned int njobs = npix_padded * frame_blocks;
    const float3 camera = f3(P.c.uCameraPos), sun = f3(P.c.uSunDirection);
    const float inv_thickness = 1.0f / (P.c.uTopAltitude - P.c.uBottomAltitude);
    int lookups = 0, paths = 0;
    for (;;) {
        // warp-aggregated job fetch among the lanes that arrive together
        unsigned int job;
        {
            const unsigned int m = __activemask();
            const int leader = __ffs(m) - 1;
            unsigned int base = 0;
            if ((threadIdx.x & 31u) == (unsigned int)leader) base = atomicAdd(P.job_counter, (unsigned int)__popc(m));
            base = __shfl_sync(m, base, leader);
            job = base + __popc(m & ((1u << (threadIdx.x & 31u)) - 1u));
        }
        if (job >= njobs) break;
        const unsigned int fb = job / npix_padded, p = job - fb * npix_padded;
        const unsigned int tile = p >> 5, in_tile = p & 31u;
        const int px = P.x0 + int(tile % (unsigned int)tiles_x) * 8 + int(in_tile & 7u);
        const int py = P.y0 + int(tile / (unsigned int)tiles_x) * 4 + int(in_tile >> 3);
        if (px >= P.x1 || py >= P.y1) continue;
        const float2 uv = f2((float(px) + 0.5f) / float(P.width), (float(py) + 0.5f) / float(P.height));
        const float3 frag_pos = projective_mul(P.c.uInvMVP, f3(uv.x * 2.0f - 1.0f, uv.y * 2.0f - 1.0f, 1.0f));
        const float3 view_dir = normalize(frag_pos - camera);
        const unsigned int f_end = min(fb * kFramesPerJob + kFramesPerJob, P.frame_count);
        for (unsigned int frame_index = fb * kFramesPerJob; frame_index < f_end; ++frame_index) {
            uint32_t seed = PRNG<PRNG_KIND>(PRNG<PRNG_KIND>(PRNG<PRNG_KIND>(uint32_t(px)) + uint32_t(py)) + (P.frame_begin + frame_index));  // :260
            // ---- Trace, :160-249
            float3 L = f3(0.0f), throughput = f3(1.0f);
            bool has_scattered = false;
            float scattered_t = 0.0f;
            float3 ro = camera, rd = view_dir;
            if (COUNT) ++paths;
            float2 camera_inter_t = CloudRegionIntersect(P, ro, rd);
            if (!(camera_inter_t.x >= camera_inter_t.y)) {
                ro += camera_inter_t.x * rd;
                int istep = 0;
                while (istep < P.pt.max_bounces && fmaxf(throughput.x, fmaxf(throughput.y, throughput.z)) > 0.0f) {
                    float2 inter_t = CloudRegionIntersect(P, ro, rd);
                    if (inter_t.x >= inter_t.y) break;
                    float t_hit = 0.0f, unused_tr;
                    bool event_scatter = P.pt.sigma_t_max > 0 &&
                                         TrackMajorantGrid<false, PRNG_KIND>(P, G, ro, rd, inter_t.x, inter_t.y, inv_thickness, seed, t_hit, unused_tr, lookups);
                    float3 nee_pos, nee_bsdf;
                    bool nee = false;
                    if (!event_scatter) {  // :198-230
                        if (P.pt.environment_lighting == SKY_ENV_OFF) break;
                        if (!has_scattered) break;
                        const float* M = P.pt.model_matrix3;
                        float3 env_dir = f3(M[0] * rd.x + M[3] * rd.y + M[6] * rd.z, M[1] * rd.x + M[4] * rd.y + M[7] * rd.z, M[2] * rd.x + M[5] * rd.y + M[8] * rd.z);
                        if (P.pt.environment_lighting == SKY_ENV_CONST_ENVIRONMENT_MAP) { L += throughput * SampleEnvironment(P, env_dir); break; }
                        float3 up_dir = f3(ro.x, ro.y, ro.z + P.c.uEarthRadius);
                        float r = length(up_dir);
                        up_dir /= r;
                        float mu = dot(rd, up_dir);
                        if (!P.atm.RayIntersectsGround(r, mu)) { L += throughput * SampleEnvironment(P, env_dir); break; }
                        ro += rd * P.atm.DistanceToBottomAtmosphereBoundary(r, mu);
                        float3 ground_normal = normalize(f3(ro.x, ro.y, ro.z + P.c.uEarthRadius));
                        nee = true; nee_pos = ro; nee_bsdf = kInvPi * P.atm.ground_albedo() * dot(ground_normal, sun);
                    } else {               // :231-243
                        if (!has_scattered) scattered_t = distance(camera, ro);
                        has_scattered = true;
                        ro += rd * t_hit;
                        nee = true; nee_pos = ro; nee_bsdf = f3(GetPhase(P, dot(rd, sun)));
                    }
                    if (nee) {  // SampleLuminanceFromLight, :153-158 (the estimate of the sun transmittance uses its own draws here)
                        float tr = 1.0f, th;
                        float2 it = CloudRegionIntersect(P, nee_pos, sun);
                        if (!(it.x >= it.y)) TrackMajorantGrid<true, PRNG_KIND>(P, G, nee_pos, sun, it.x, it.y, inv_thickness, seed, th, tr, lookups);
                        L += throughput * (clampf(tr, 0.0f, 1.0f) * GetSunIlluminance(P, nee_pos) * nee_bsdf);
                    }
                    if (!event_scatter) {
                        if (P.pt.environment_lighting == SKY_ENV_GROUND_SINGLE_BOUNCE) break;
                        // GenerateLambertSample, :117-129 (cosine-weighted hemisphere about the ground normal)
                        float3 N = normalize(f3(ro.x, ro.y, ro.z + P.c.uEarthRadius));
                        float3 b0, b1;
                        CreateOrthonormalBasis(N, b0, b1);
                        float sin_theta = sqrtf(Random01<PRNG_KIND>(seed));
                        float cos_theta = sqrtf(clampf(1.0f - sin_theta * sin_theta, 0.0f, 1.0f));
                        float phi = 2.0f * kPi * Random01<PRNG_KIND>(seed);
                        rd = sin_theta * sinf(phi) * b0 + sin_theta * cosf(phi) * b1 + cos_theta * N;
                        throughput *= P.atm.ground_albedo();
                    } else if (P.pt.importance_sampling) {
                        // GenerateHGSample with importance sampling, :98-111
                        float g = Random01<PRNG_KIND>(seed) < P.pt.forward_scattering_ratio ? P.pt.forward_phase_g : P.pt.back_phase_g;
                        float cos_theta = HenyeyGreensteinInvertcdf(Random01<PRNG_KIND>(seed), g);
                        float sin_theta = sqrtf(clampf(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
                        float3 b0, b1;
                        CreateOrthonormalBasis(rd, b0, b1);
                        float phi = 2.0f * kPi * Random01<PRNG_KIND>(seed);
                        rd = sin_theta * sinf(phi) * b0 + sin_theta * cosf(phi) * b1 + cos_theta * rd;
                    } else {
                        float phi = 2.0f * kPi * Random01<PRNG_KIND>(seed);
                        float cos_theta = 1.0f - 2.0f * Random01<PRNG_KIND>(seed);
                        float sin_theta = sqrtf(clampf(1.0f - cos_theta * cos_theta, 0.0f, 1.0f));
                        float3 direction = f3(cosf(phi) * sin_theta, sinf(phi) * sin_theta, cos_theta);
                        throughput *= GetPhase(P, dot(rd, direction)) / (1.0f / (4.0f * kPi));
                        rd = direction;
                    }
                    ++istep;
                }
            }
            // ---- :248, :263-283
            float4 this_res = f4(L, has_scattered ? 0.0f : 1.0f);
            if (has_scattered) {
                float r = P.c.uCameraPos[2] + P.c.uEarthRadius;
                float mu = view_dir.z;
                float ap_t = scattered_t;
                if (r > P.atm.u.top_radius) {
                    float near_distance;
                    if (P.atm.FromSpaceIntersectTopAtmosphereBoundary(r, mu, near_distance)) ap_t -= near_distance;
                    else ap_t = 0;
                }
                float3 uvw = aerial_perspective_uvw(uv, ap_t, P.c.uAerialPerspectiveLutMaxDistance, P.ap_lum.w, P.ap_lum.h, P.ap_lum.d);
                float3 atmosphere_transmittance = xyz(sample_lut3d(P.ap_trans, uvw.x, uvw.y, uvw.z));
                float3 atmosphere_luminance = xyz(sample_lut3d(P.ap_lum, uvw.x, uvw.y, uvw.z));
                atmosphere_luminance *= SampleRayScatterVisibility(P.froxel, uv, scattered_t, P.c.uInvShadowFroxelMaxDistance);
                this_res = f4(xyz(this_res) * atmosphere_transmittance + atmosphere_luminance, this_res.w);
            }
            P.samples[size_t(frame_index) * npix + size_t(py - P.y0) * rw + (px - P.x0)] = this_res;
        }
    }
    if (COUNT) {
        atomicAdd(P.counters + SKY_CNT_PT_PATHS, (unsigned long long)paths);
        atomicAdd(P.counters + SKY_CNT_PT_LOOKUPS, (unsigned long long)lookups);
    }
}

// second half of K19 (:281-283): accumulated += this_res, one sample per kFrameId, in frame order.
__global__ void __launch_bounds__(256) k19_accumulate(const __grid_constant__ PtParams P) {
    const int rw = P.x1 - P.x0, rh = P.y1 - P.y0;
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= rw || y >= rh) return;
    const size_t npix = size_t(rw) * rh, local = size_t(y) * rw + x;
    const size_t pix = size_t(P.y0 + y) * P.width + (P.x0 + x);
    float4 accumulated = P.accum[pix];
    for (uint32_t f = 0; f < P.frame_count; ++f) accumulated = accumulated + __ldcs(P.samples + size_t(f) * npix + local);
    P.accum[pix] = accumulated;
    P.mask[pix] = 1;
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
    P.transmittance = LutView{ctx->transmittance.p, ctx->transmittance.w, ctx->transmittance.h, 1, 0};
    P.ap_lum = LutView{ctx->ap_lum.p, ctx->ap_lum.w, ctx->ap_lum.h, ctx->ap_lum.d, 0};
    P.ap_trans = LutView{ctx->ap_trans.p, ctx->ap_trans.w, ctx->ap_trans.h, ctx->ap_trans.d, 0};
    P.froxel = FroxelView{ctx->shadow_froxel.p, ctx->shadow_froxel.w, ctx->shadow_froxel.h, ctx->shadow_froxel.d};
    P.env = ctx->env.p;
    P.env_size = ctx->env.w;
    P.accum = ctx->pt_accum.p;
    P.mask = ctx->pt_mask.p;
    P.counters = ctx->counters;
    P.width = ctx->width; P.height = ctx->height;
    if (ctx->material.type == SKY_MATERIAL_VOXEL && ctx->voxel.valid) {
        // texel = (pos * frequency + bias) * size - 0.5 (x, y);  saturate((pos.z - bottom) / thickness) * depth - 0.5 (z)
        const SkyMaterialVoxelBufferData& vm = P.mat.m.u.voxel;
        const float w = float(P.mat.voxel.w[0]), h = float(P.mat.voxel.h[0]);
        const float inv_thickness = 1.0f / (c.uTopAltitude - c.uBottomAltitude);
        P.vx_scale = vm.uSampleFrequency[0] * w; P.vx_bias = vm.uSampleBias[0] * w - 0.5f;
        P.vy_scale = vm.uSampleFrequency[1] * h; P.vy_bias = vm.uSampleBias[1] * h - 0.5f;
        P.vz_scale = inv_thickness; P.vz_bias = -c.uBottomAltitude * inv_thickness;
        P.vz_depth = float(P.mat.voxel.d[0]);
        P.v_collision_scale = vm.uDensity * (1.0f / 255.0f) / ctx->pt.sigma_t_max;
    }
    return P;
}

}  // namespace

int launch_pt_samples(SkyContext* ctx, const SkyCloudCommonBufferData& c, uint32_t frame_begin, uint32_t count, const int32_t region[4]) {
    SKY_PERF_MARKER("PathTracing");  // PathTracing::Render, VolumetricCloud.cpp:533-560 (the reference sets no marker of its own here)
    if (!ctx->pt_accum.p) return sky_fail(ctx, "pt_begin was not called");
    if (!ctx->ap_lum.p || !ctx->env.p) return sky_fail(ctx, "atmosphere LUTs have not been baked");
    if (ctx->material.type == SKY_MATERIAL_VOXEL && !ctx->voxel.valid) return sky_fail(ctx, "voxel grid has not been uploaded");
    if ((ctx->material.type == SKY_MATERIAL_DEFAULT0 || ctx->material.type == SKY_MATERIAL_DEFAULT1) && (!ctx->cloud_map.valid || !ctx->detail.valid))
        return sky_fail(ctx, "cloud map / detail texture has not been generated");
    if (count == 0) return 0;
    PtParams P = make_pt_params(ctx, c);
    P.x0 = std::max(region[0], 0); P.y0 = std::max(region[1], 0);
    P.x1 = std::min(region[2], ctx->width); P.y1 = std::min(region[3], ctx->height);
    const int rw = P.x1 - P.x0, rh = P.y1 - P.y0;
    if (rw <= 0 || rh <= 0) return 0;
    // tile-padded job space: ceil(rw/8) x ceil(rh/4) tiles of 32 pixels
    const size_t padded = size_t((rw + 7) / 8) * size_t((rh + 3) / 4) * 32;
    // sample slots: at most ~4 GiB per launch (HBM is plentiful; every launch ends with a drain phase in which a few lanes finish
    // the longest paths, DESIGN.md section 7, so fewer, longer launches are better), longer jobs run in chunks of frames
    uint32_t frames_per_launch = uint32_t(std::max<size_t>(1, std::min<size_t>(count, (size_t(1) << 28) / padded)));
    if (padded * frames_per_launch >= (size_t(1) << 32)) return sky_fail(ctx, "region too large for one launch");
    const size_t need = padded * frames_per_launch * sizeof(float4);
    if (ctx->pt_samples_bytes < need) {
        if (ctx->pt_samples) SKY_CUDA(ctx, cudaFree(ctx->pt_samples));
        ctx->pt_samples = nullptr; ctx->pt_samples_bytes = 0;
        SKY_CUDA(ctx, cudaMalloc(&ctx->pt_samples, need));
        ctx->pt_samples_bytes = need;
    }
    if (!ctx->pt_job_counter) SKY_CUDA(ctx, cudaMalloc(&ctx->pt_job_counter, sizeof(unsigned int)));
    P.samples = static_cast<float4*>(ctx->pt_samples);
    P.job_counter = ctx->pt_job_counter;

    int sm_count = 148;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, ctx->device);
    const bool count_on = ctx->counting;
    const int prng = ctx->pt.prng;
    // majorant-grid mode (sky_pt_set_tracking): build the macro-cell majorants of the current voxel texture on first use
    const bool majorant_mode = ctx->pt_tracking == SKY_PT_TRACKING_MAJORANT_GRID;
    MajorantView G{};
    if (majorant_mode) {
        if (ctx->material.type != SKY_MATERIAL_VOXEL) return sky_fail(ctx, "majorant-grid tracking needs the voxel material");
        const MipView& v = ctx->voxel.view;
        G.w = ceil_div(v.w[0], kMacro); G.h = ceil_div(v.h[0], kMacro); G.d = ceil_div(v.d[0], kMacro);
        const size_t cells = size_t(G.w) * G.h * G.d;
        if (!ctx->voxel_majorant || ctx->voxel_majorant_cells != cells || !ctx->voxel_majorant_valid) {
            if (ctx->voxel_majorant) SKY_CUDA(ctx, cudaFree(ctx->voxel_majorant));
            ctx->voxel_majorant = nullptr;
            SKY_CUDA(ctx, cudaMalloc(&ctx->voxel_majorant, cells));
            ctx->voxel_majorant_cells = cells;
            k_majorant_build<<<unsigned((cells + 127) / 128), 128, 0, ctx->stream>>>(v, ctx->voxel_majorant, G.w, G.h, G.d);
            SKY_LAUNCH_CHECK(ctx);
            ctx->voxel_majorant_valid = true;
        }
        G.p = ctx->voxel_majorant;
    }
    for (uint32_t done = 0; done < count; done += frames_per_launch) {
        P.frame_begin = frame_begin + done;
        P.frame_count = std::min(frames_per_launch, count - done);
        SKY_CUDA(ctx, cudaMemsetAsync(ctx->pt_job_counter, 0, sizeof(unsigned int), ctx->stream));
        // NOTE: the job space is the tile-padded pixel count, the slots are indexed by real pixels
        PtParams Q = P;
        if (majorant_mode) {
            auto launch = [&](auto kernel) {
                int per_sm = 1;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0);
                size_t jobs = padded * ((Q.frame_count + 7) / 8);
                unsigned blocks = unsigned(std::min<size_t>(size_t(sm_count) * std::max(per_sm, 1), (jobs + 127) / 128));
                kernel<<<blocks, 128, 0, ctx->stream>>>(Q, G);
            };
            if (prng == SKY_PRNG_WANG) { if (count_on) launch(k19_majorant_grid<SKY_PRNG_WANG, true>); else launch(k19_majorant_grid<SKY_PRNG_WANG, false>); }
            else { if (count_on) launch(k19_majorant_grid<SKY_PRNG_PCG, true>); else launch(k19_majorant_grid<SKY_PRNG_PCG, false>); }
            SKY_LAUNCH_CHECK(ctx);
            k19_accumulate<<<dim3(ceil_div(rw, 256), rh), 256, 0, ctx->stream>>>(Q);
            SKY_LAUNCH_CHECK(ctx);
            continue;
        }
        int rc = dispatch_material(ctx->material.type, ctx->hw_filtering, [&]<int MAT, bool HW>() {
            auto launch = [&](auto kernel) {
                int per_sm = 1;
                cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 128, 0);
                size_t jobs = padded * Q.frame_count;
                unsigned blocks = unsigned(std::min<size_t>(size_t(sm_count) * std::max(per_sm, 1), (jobs + 127) / 128));
                kernel<<<blocks, 128, 0, ctx->stream>>>(Q);
            };
            if (prng == SKY_PRNG_WANG) {
                if (count_on) launch(k19_path_trace<MAT, HW, SKY_PRNG_WANG, true>);
                else launch(k19_path_trace<MAT, HW, SKY_PRNG_WANG, false>);
            } else {
                if (count_on) launch(k19_path_trace<MAT, HW, SKY_PRNG_PCG, true>);
                else launch(k19_path_trace<MAT, HW, SKY_PRNG_PCG, false>);
            }
            return 0;
        });
        if (rc) return sky_fail(ctx, "unknown material");
        SKY_LAUNCH_CHECK(ctx);
        k19_accumulate<<<dim3(ceil_div(rw, 256), rh), 256, 0, ctx->stream>>>(Q);
        SKY_LAUNCH_CHECK(ctx);
    }
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
