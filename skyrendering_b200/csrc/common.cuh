// Device-side vector helpers and launch plumbing shared by the sm_100a kernels.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sky_types.h"

#define SKY_HD __host__ __device__ __forceinline__
#define SKY_D __device__ __forceinline__

constexpr float kPi = 3.1415926535897932384626433832795f;  // shaders/Base/Common.glsl:4
constexpr float kInvPi = 1.0f / kPi;

// ---- float3 / float4 arithmetic --------------------------------------------------------------------
SKY_HD float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
SKY_HD float3 f3(float s) { return make_float3(s, s, s); }
SKY_HD float3 f3(const float* p) { return make_float3(p[0], p[1], p[2]); }
SKY_HD float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
SKY_HD float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
SKY_HD float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
SKY_HD float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }
SKY_HD float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
SKY_HD float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
SKY_HD float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
SKY_HD float3 operator/(float s, float3 a) { return f3(s / a.x, s / a.y, s / a.z); }
SKY_HD float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
SKY_HD float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
SKY_HD float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
SKY_HD float3& operator*=(float3& a, float s) { a = a * s; return a; }
SKY_HD float3& operator/=(float3& a, float s) { a = a / s; return a; }
SKY_HD float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
SKY_HD float length(float3 a) { return sqrtf(dot(a, a)); }
SKY_HD float3 normalize(float3 a) { return a / length(a); }
SKY_HD float distance(float3 a, float3 b) { return length(a - b); }
SKY_HD float3 exp3(float3 a) { return f3(expf(a.x), expf(a.y), expf(a.z)); }

SKY_HD float4 f4(float x, float y, float z, float w) { return make_float4(x, y, z, w); }
SKY_HD float4 f4(float3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
SKY_HD float4 operator+(float4 a, float4 b) { return f4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
SKY_HD float4 operator-(float4 a, float4 b) { return f4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
SKY_HD float4 operator*(float4 a, float s) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
SKY_HD float4 operator*(float s, float4 a) { return f4(a.x * s, a.y * s, a.z * s, a.w * s); }
SKY_HD float4& operator+=(float4& a, float4 b) { a = a + b; return a; }
SKY_HD float3 xyz(float4 a) { return f3(a.x, a.y, a.z); }
SKY_HD float4 min4(float4 a, float4 b) { return f4(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z), fminf(a.w, b.w)); }
SKY_HD float4 max4(float4 a, float4 b) { return f4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)); }

SKY_HD float2 f2(float x, float y) { return make_float2(x, y); }
SKY_HD float2 operator+(float2 a, float2 b) { return f2(a.x + b.x, a.y + b.y); }
SKY_HD float2 operator-(float2 a, float2 b) { return f2(a.x - b.x, a.y - b.y); }
SKY_HD float2 operator*(float2 a, float s) { return f2(a.x * s, a.y * s); }
SKY_HD float2 operator*(float s, float2 a) { return f2(a.x * s, a.y * s); }
SKY_HD float2& operator+=(float2& a, float2 b) { a = a + b; return a; }

// GLSL built-ins with GLSL's definitions
SKY_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
SKY_HD int clampi(int x, int lo, int hi) { return min(max(x, lo), hi); }
SKY_HD float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
SKY_HD float2 mix2(float2 a, float2 b, float t) { return a * (1.0f - t) + b * t; }
SKY_HD float4 mix4(float4 a, float4 b, float t) { return a * (1.0f - t) + b * t; }
SKY_HD float fractf(float x) { return x - floorf(x); }
SKY_HD float smoothstepf(float e0, float e1, float x) {
    float t = clampf((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}

// column-major mat4 (m[c*4+r]) * (v,1), then perspective divide: ProjectiveMul, shaders/Base/Common.glsl:7-10
SKY_HD float3 projective_mul(const float* m, float3 v) {
    float x = m[0] * v.x + m[4] * v.y + m[8] * v.z + m[12] * 1.0f;
    float y = m[1] * v.x + m[5] * v.y + m[9] * v.z + m[13] * 1.0f;
    float z = m[2] * v.x + m[6] * v.y + m[10] * v.z + m[14] * 1.0f;
    float w = m[3] * v.x + m[7] * v.y + m[11] * v.z + m[15] * 1.0f;
    return f3(x / w, y / w, z / w);
}

// RGBA16F texel <-> float4 (round-to-nearest-even like a GL rgba16f image store)
struct __align__(8) half4 { __half x, y, z, w; };
SKY_D half4 to_half4(float4 v) {
    half4 h;
    h.x = __float2half_rn(v.x); h.y = __float2half_rn(v.y); h.z = __float2half_rn(v.z); h.w = __float2half_rn(v.w);
    return h;
}
SKY_D float4 from_half4(half4 h) { return f4(__half2float(h.x), __half2float(h.y), __half2float(h.z), __half2float(h.w)); }
SKY_D float4 load_half4(const half4* p) {
    uint2 raw = __ldg(reinterpret_cast<const uint2*>(p));
    half4 h = *reinterpret_cast<half4*>(&raw);
    return from_half4(h);
}

SKY_HD int ceil_div(int a, int b) { return (a + b - 1) / b; }
