// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or executed from the product
// path (skyrendering_b200/, libskyb200.so); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference leg may use it.
//
// glsl.h: a minimal fp32 GLSL-flavoured vector layer so the CPU restatement of the reference
// shaders (shaders/SkyRendering/*.glsl|*.comp) reads like the GLSL it follows.  Built with
// -ffp-contract=off so no FMA contraction sneaks in.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl {

typedef uint32_t uint;

struct vec2 {
    float x, y;
    vec2() : x(0), y(0) {}
    explicit vec2(float s) : x(s), y(s) {}
    vec2(float x_, float y_) : x(x_), y(y_) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct vec3 {
    float x, y, z;
    vec3() : x(0), y(0), z(0) {}
    explicit vec3(float s) : x(s), y(s), z(s) {}
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    vec3(vec2 v, float z_) : x(v.x), y(v.y), z(z_) {}
    explicit vec3(const float* p) : x(p[0]), y(p[1]), z(p[2]) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    vec2 xy() const { return vec2(x, y); }
    vec2 xz() const { return vec2(x, z); }
};
struct vec4 {
    float x, y, z, w;
    vec4() : x(0), y(0), z(0), w(0) {}
    explicit vec4(float s) : x(s), y(s), z(s), w(s) {}
    vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    vec4(vec3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
    vec3 xyz() const { return vec3(x, y, z); }
    vec3 rgb() const { return vec3(x, y, z); }
};
struct ivec2 {
    int x, y;
    ivec2() : x(0), y(0) {}
    ivec2(int x_, int y_) : x(x_), y(y_) {}
};
struct ivec3 {
    int x, y, z;
    ivec3() : x(0), y(0), z(0) {}
    ivec3(int x_, int y_, int z_) : x(x_), y(y_), z(z_) {}
};

#define GLSL_BINOP(T, N, op)                                                   \
    inline T operator op(T a, T b) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] op b[i]; return r; } \
    inline T operator op(T a, float b) { T r; for (int i = 0; i < N; ++i) r[i] = a[i] op b; return r; } \
    inline T operator op(float a, T b) { T r; for (int i = 0; i < N; ++i) r[i] = a op b[i]; return r; } \
    inline T& operator op##=(T& a, T b) { for (int i = 0; i < N; ++i) a[i] op##= b[i]; return a; }      \
    inline T& operator op##=(T& a, float b) { for (int i = 0; i < N; ++i) a[i] op##= b; return a; }
#define GLSL_OPS(T, N) GLSL_BINOP(T, N, +) GLSL_BINOP(T, N, -) GLSL_BINOP(T, N, *) GLSL_BINOP(T, N, /) \
    inline T operator-(T a) { T r; for (int i = 0; i < N; ++i) r[i] = -a[i]; return r; }
GLSL_OPS(vec2, 2)
GLSL_OPS(vec3, 3)
GLSL_OPS(vec4, 4)

inline ivec2 operator+(ivec2 a, ivec2 b) { return ivec2(a.x + b.x, a.y + b.y); }
inline ivec2 operator-(ivec2 a, int b) { return ivec2(a.x - b, a.y - b); }
inline ivec2 operator*(ivec2 a, int b) { return ivec2(a.x * b, a.y * b); }
inline ivec2 operator>>(ivec2 a, int b) { return ivec2(a.x >> b, a.y >> b); }
inline ivec2 operator<<(ivec2 a, int b) { return ivec2(a.x << b, a.y << b); }
inline ivec2 operator&(ivec2 a, int b) { return ivec2(a.x & b, a.y & b); }
inline bool operator==(ivec2 a, ivec2 b) { return a.x == b.x && a.y == b.y; }
inline vec2 tovec2(ivec2 a) { return vec2(float(a.x), float(a.y)); }
inline vec3 tovec3(ivec3 a) { return vec3(float(a.x), float(a.y), float(a.z)); }

// GLSL leaves min/max/clamp of NaN undefined; GPUs implement them with the IEEE-754 minNum/maxNum
// instruction (NVIDIA FMNMX), which returns the non-NaN operand.  The oracle does the same, so e.g.
// clamp(0.0/0.0, 0, 1) == 0 exactly as in the shaders' RemapTo01 with remap_min == remap_max
// (bin/config.json detail_.buffer.uPerlin).
inline float max(float a, float b) { return std::fmax(a, b); }
inline float min(float a, float b) { return std::fmin(a, b); }
inline int max(int a, int b) { return a < b ? b : a; }
inline int min(int a, int b) { return b < a ? b : a; }
inline float clamp(float x, float lo, float hi) { return std::fmin(std::fmax(x, lo), hi); }
inline int clamp(int x, int lo, int hi) { return min(max(x, lo), hi); }
inline float mix(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline float fract(float x) { return x - std::floor(x); }
inline float radians(float d) { return d * 0.01745329251994329576923690768489f; }
inline float smoothstep(float e0, float e1, float x) {
    float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }

#define GLSL_MAP1(T, N, name, expr) \
    inline T name(T a) { T r; for (int i = 0; i < N; ++i) { float v = a[i]; r[i] = (expr); } return r; }
#define GLSL_MAPS(T, N)                                     \
    GLSL_MAP1(T, N, exp, std::exp(v))                        \
    GLSL_MAP1(T, N, sqrt, std::sqrt(v))                      \
    GLSL_MAP1(T, N, abs, std::fabs(v))                       \
    GLSL_MAP1(T, N, floor, std::floor(v))                    \
    GLSL_MAP1(T, N, fract, v - std::floor(v))                \
    inline T min(T a, T b) { T r; for (int i = 0; i < N; ++i) r[i] = std::fmin(a[i], b[i]); return r; } \
    inline T max(T a, T b) { T r; for (int i = 0; i < N; ++i) r[i] = std::fmax(a[i], b[i]); return r; } \
    inline T clamp(T a, T lo, T hi) { return min(max(a, lo), hi); }                                    \
    inline T clamp(T a, float lo, float hi) { T r; for (int i = 0; i < N; ++i) r[i] = clamp(a[i], lo, hi); return r; } \
    inline T mix(T a, T b, float t) { return a * (1.0f - t) + b * t; }                                 \
    inline float dot(T a, T b) { float s = a[0] * b[0]; for (int i = 1; i < N; ++i) s += a[i] * b[i]; return s; } \
    inline float length(T a) { return std::sqrt(dot(a, a)); }                                          \
    inline float distance(T a, T b) { return length(a - b); }                                          \
    inline T normalize(T a) { return a / length(a); }
GLSL_MAPS(vec2, 2)
GLSL_MAPS(vec3, 3)
GLSL_MAPS(vec4, 4)

inline vec3 pow(vec3 a, vec3 b) { return vec3(std::pow(a.x, b.x), std::pow(a.y, b.y), std::pow(a.z, b.z)); }
inline vec3 cross(vec3 a, vec3 b) {
    return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}

// column-major 4x4, m[c*4 + r] (glm / GLSL layout)
struct mat4 {
    float m[16];
    mat4() { std::memset(m, 0, sizeof(m)); }
    explicit mat4(const float* p) { std::memcpy(m, p, sizeof(m)); }
};
inline vec4 operator*(const mat4& a, vec4 v) {
    vec4 r;
    for (int i = 0; i < 4; ++i)
        r[i] = a.m[0 + i] * v.x + a.m[4 + i] * v.y + a.m[8 + i] * v.z + a.m[12 + i] * v.w;
    return r;
}
struct mat3 {
    float m[9];
    explicit mat3(const float* p) { std::memcpy(m, p, sizeof(m)); }
};
inline vec3 operator*(const mat3& a, vec3 v) {
    vec3 r;
    for (int i = 0; i < 3; ++i) r[i] = a.m[0 + i] * v.x + a.m[3 + i] * v.y + a.m[6 + i] * v.z;
    return r;
}

// fp16 storage round trip (RGBA16F images), round-to-nearest-even like the GL image store.
inline float to_half_and_back(float x) { return (float)(_Float16)x; }
inline uint16_t float_to_half_bits(float x) {
    _Float16 h = (_Float16)x;
    uint16_t b;
    std::memcpy(&b, &h, 2);
    return b;
}
inline float half_bits_to_float(uint16_t b) {
    _Float16 h;
    std::memcpy(&h, &b, 2);
    return (float)h;
}

}  // namespace glsl
