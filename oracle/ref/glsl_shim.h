// ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/glsl.h).  Never linked into or executed from the product path.
//
// glsl_shim.h: enough of GLSL 4.60 in C++17 that the reference's OWN shader text (shaders/SkyRendering/*.glsl|*.comp,
// shaders/Base/*.glsl under /root/reference) compiles with g++ after the mechanical rewrites of glsl2cpp.py -- so the
// hand-written restatement in oracle/*.cpp can be checked against the real thing ("oracle/_ref", built only where the
// reference tree is present; nothing of the reference is copied into this repository).
//
// What the GLSL specification leaves to the driver is supplied here with the SAME conventions the oracle and the CUDA
// kernels use: fp32 throughout, no FMA contraction (-ffp-contract=off), IEEE sqrt and division,
// exact fp32 filter weights, round-to-nearest-even UNORM / fp16 image stores, minNum / maxNum for min / max / clamp, and
// -- when REF_MATH_DET is defined, as for the LUT programs -- exp / sin / cos / acos / x^1.5 from include/sky_detmath.h.
#pragma once
#include "../../include/sky_cubemap.h"
#include "../../include/sky_texgrad.h"
#include <algorithm>
#include <array>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#include "../../include/sky_detmath.h"

namespace ref {

typedef uint32_t uint;

// ---- vectors with the swizzles the shaders use (.xy .rg .xz .ba .xyz .rgb) -----------------------------------------
template <class V, class S, int N, int A, int B>
struct Swz2 {
    S d[N];
    operator V() const { return V(d[A], d[B]); }
    Swz2& operator=(const V& v) { d[A] = v.x; d[B] = v.y; return *this; }
    Swz2& operator+=(const V& v) { d[A] += v.x; d[B] += v.y; return *this; }
    Swz2& operator*=(const V& v) { d[A] *= v.x; d[B] *= v.y; return *this; }
};
template <class V, class S, int N, int A, int B, int C>
struct Swz3 {
    S d[N];
    operator V() const { return V(d[A], d[B], d[C]); }
    Swz3& operator=(const V& v) { d[A] = v.x; d[B] = v.y; d[C] = v.z; return *this; }
    Swz3& operator+=(const V& v) { d[A] += v.x; d[B] += v.y; d[C] += v.z; return *this; }
    Swz3& operator*=(const V& v) { d[A] *= v.x; d[B] *= v.y; d[C] *= v.z; return *this; }
};

struct vec2; struct vec3; struct vec4; struct ivec2; struct ivec3; struct ivec4; struct uvec2; struct uvec3; struct uvec4;
#define REF_VEC2(NAME, S, O1, O2)                                                                              \
    struct NAME {                                                                                      \
        union {                                                                                        \
            struct { S x, y; };                                                                        \
            struct { S r, g; };                                                                        \
            S d[2];                                                                                    \
            Swz2<NAME, S, 2, 0, 1> xy, rg;                                                             \
        };                                                                                             \
        NAME() : x(0), y(0) {}                                                                         \
        NAME(const NAME& o) : x(o.x), y(o.y) {}                                                        \
        explicit NAME(const O1& o); explicit NAME(const O2& o);                                        \
        NAME& operator=(const NAME& o) { x = o.x; y = o.y; return *this; }                             \
        template <class T, class = std::enable_if_t<std::is_arithmetic_v<T>>> NAME(T s) : x(S(s)), y(S(s)) {} \
        template <class T, class U, class = std::enable_if_t<std::is_arithmetic_v<T> && std::is_arithmetic_v<U>>> \
        NAME(T a, U b) : x(S(a)), y(S(b)) {}                                                           \
        S& operator[](int i) { return d[i]; }                                                          \
        S operator[](int i) const { return d[i]; }                                                     \
    };
#define REF_VEC3(NAME, V2, S, O1, O2, A2, B2)                                                                          \
    struct NAME {                                                                                      \
        union {                                                                                        \
            struct { S x, y, z; };                                                                     \
            struct { S r, g, b; };                                                                     \
            S d[3];                                                                                    \
            Swz2<V2, S, 3, 0, 1> xy, rg;                                                               \
            Swz2<V2, S, 3, 0, 2> xz;                                                                   \
            Swz3<NAME, S, 3, 0, 1, 2> xyz, rgb;                                                        \
        };                                                                                             \
        NAME() : x(0), y(0), z(0) {}                                                                   \
        NAME(const NAME& o) : x(o.x), y(o.y), z(o.z) {}                                                \
        explicit NAME(const O1& o); explicit NAME(const O2& o);                                        \
        NAME& operator=(const NAME& o) { x = o.x; y = o.y; z = o.z; return *this; }                    \
        template <class T, class = std::enable_if_t<std::is_arithmetic_v<T>>> NAME(T s) : x(S(s)), y(S(s)), z(S(s)) {} \
        template <class T, class U, class W, class = std::enable_if_t<std::is_arithmetic_v<T> && std::is_arithmetic_v<U> && std::is_arithmetic_v<W>>> \
        NAME(T a, U b, W c) : x(S(a)), y(S(b)), z(S(c)) {}                                             \
        template <class T, class = std::enable_if_t<std::is_arithmetic_v<T>>> NAME(const V2& v, T c) : x(v.x), y(v.y), z(S(c)) {} \
        template <class T, class = std::enable_if_t<std::is_arithmetic_v<T>>> NAME(const A2& v, T c); \
        template <class T, class = std::enable_if_t<std::is_arithmetic_v<T>>> NAME(const B2& v, T c); \
        template <class T, class = std::enable_if_t<std::is_arithmetic_v<T>>> NAME(T a, const V2& v) : x(S(a)), y(v.x), z(v.y) {} \
        S& operator[](int i) { return d[i]; }                                                          \
        S operator[](int i) const { return d[i]; }                                                     \
    };
#define REF_VEC4(NAME, V2, V3, S, O1, O2)                                                                      \
    struct NAME {                                                                                      \
        union {                                                                                        \
            struct { S x, y, z, w; };                                                                  \
            struct { S r, g, b, a; };                                                                  \
            S d[4];                                                                                    \
            Swz2<V2, S, 4, 0, 1> xy, rg;                                                               \
            Swz2<V2, S, 4, 0, 2> xz;                                                                   \
            Swz2<V2, S, 4, 2, 3> ba, zw;                                                               \
            Swz3<V3, S, 4, 0, 1, 2> xyz, rgb;                                                          \
        };                                                                                             \
        NAME() : x(0), y(0), z(0), w(0) {}                                                             \
        NAME(const NAME& o) : x(o.x), y(o.y), z(o.z), w(o.w) {}                                        \
        explicit NAME(const O1& o); explicit NAME(const O2& o);                                        \
        NAME& operator=(const NAME& o) { x = o.x; y = o.y; z = o.z; w = o.w; return *this; }           \
        template <class T, class = std::enable_if_t<std::is_arithmetic_v<T>>> NAME(T s) : x(S(s)), y(S(s)), z(S(s)), w(S(s)) {} \
        template <class T, class U, class W, class X, class = std::enable_if_t<std::is_arithmetic_v<T> && std::is_arithmetic_v<U> && std::is_arithmetic_v<W> && std::is_arithmetic_v<X>>> \
        NAME(T a_, U b_, W c_, X d_) : x(S(a_)), y(S(b_)), z(S(c_)), w(S(d_)) {}                       \
        template <class T, class = std::enable_if_t<std::is_arithmetic_v<T>>> NAME(const V3& v, T d_) : x(v.x), y(v.y), z(v.z), w(S(d_)) {} \
        template <class T, class U, class = std::enable_if_t<std::is_arithmetic_v<T> && std::is_arithmetic_v<U>>> \
        NAME(const V2& v, T c_, U d_) : x(v.x), y(v.y), z(S(c_)), w(S(d_)) {}                          \
        NAME(const V2& p, const V2& q) : x(p.x), y(p.y), z(q.x), w(q.y) {}                             \
        NAME(const V3& p, const V3& q) : x(p.x), y(p.y), z(p.z), w(q.x) {} /* GLSL drops the unused tail of the last argument */ \
        S& operator[](int i) { return d[i]; }                                                          \
        S operator[](int i) const { return d[i]; }                                                     \
    };

REF_VEC2(vec2, float, ivec2, uvec2) REF_VEC3(vec3, vec2, float, ivec3, uvec3, ivec2, uvec2) REF_VEC4(vec4, vec2, vec3, float, ivec4, uvec4)
REF_VEC2(ivec2, int, vec2, uvec2) REF_VEC3(ivec3, ivec2, int, vec3, uvec3, vec2, uvec2) REF_VEC4(ivec4, ivec2, ivec3, int, vec4, uvec4)
REF_VEC2(uvec2, uint, vec2, ivec2) REF_VEC3(uvec3, uvec2, uint, vec3, ivec3, vec2, ivec2) REF_VEC4(uvec4, uvec2, uvec3, uint, vec4, ivec4)
// casts between the float / int / uint families (explicit in GLSL as well): component-wise C++ conversions
#define REF_CONV21(T, A2) template <class U, class> inline T::T(const A2& v, U c) : x(decltype(x)(v.x)), y(decltype(y)(v.y)), z(decltype(z)(c)) {}
REF_CONV21(vec3, ivec2) REF_CONV21(vec3, uvec2) REF_CONV21(ivec3, vec2) REF_CONV21(ivec3, uvec2) REF_CONV21(uvec3, vec2) REF_CONV21(uvec3, ivec2)
#define REF_CONV2(T, O) inline T::T(const O& o) : x(decltype(x)(o.x)), y(decltype(y)(o.y)) {}
#define REF_CONV3(T, O) inline T::T(const O& o) : x(decltype(x)(o.x)), y(decltype(y)(o.y)), z(decltype(z)(o.z)) {}
#define REF_CONV4(T, O) inline T::T(const O& o) : x(decltype(x)(o.x)), y(decltype(y)(o.y)), z(decltype(z)(o.z)), w(decltype(w)(o.w)) {}
REF_CONV2(vec2, ivec2) REF_CONV2(vec2, uvec2) REF_CONV2(ivec2, vec2) REF_CONV2(ivec2, uvec2) REF_CONV2(uvec2, vec2) REF_CONV2(uvec2, ivec2)
REF_CONV3(vec3, ivec3) REF_CONV3(vec3, uvec3) REF_CONV3(ivec3, vec3) REF_CONV3(ivec3, uvec3) REF_CONV3(uvec3, vec3) REF_CONV3(uvec3, ivec3)
REF_CONV4(vec4, ivec4) REF_CONV4(vec4, uvec4) REF_CONV4(ivec4, vec4) REF_CONV4(ivec4, uvec4) REF_CONV4(uvec4, vec4) REF_CONV4(uvec4, ivec4)

// ---- operators (component-wise; scalars are float / int / uint: the literal rewrite makes every real literal a float) --
#define REF_BINOP(T, S, N, op)                                                                              \
    inline T operator op(const T& a, const T& b) { T r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] op b.d[i]; return r; } \
    inline T operator op(const T& a, S b) { T r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] op b; return r; }             \
    inline T operator op(S a, const T& b) { T r; for (int i = 0; i < N; ++i) r.d[i] = a op b.d[i]; return r; }             \
    inline T& operator op##=(T& a, const T& b) { for (int i = 0; i < N; ++i) a.d[i] = a.d[i] op b.d[i]; return a; }        \
    inline T& operator op##=(T& a, S b) { for (int i = 0; i < N; ++i) a.d[i] = a.d[i] op b; return a; }
#define REF_ARITH(T, S, N) REF_BINOP(T, S, N, +) REF_BINOP(T, S, N, -) REF_BINOP(T, S, N, *) REF_BINOP(T, S, N, /) \
    inline T operator-(const T& a) { T r; for (int i = 0; i < N; ++i) r.d[i] = -a.d[i]; return r; }                \
    inline bool operator==(const T& a, const T& b) { for (int i = 0; i < N; ++i) if (!(a.d[i] == b.d[i])) return false; return true; } \
    inline bool operator!=(const T& a, const T& b) { return !(a == b); }
#define REF_INTOPS(T, S, N) REF_BINOP(T, S, N, %) REF_BINOP(T, S, N, &) REF_BINOP(T, S, N, |) REF_BINOP(T, S, N, ^) \
    REF_BINOP(T, S, N, >>) REF_BINOP(T, S, N, <<)
REF_ARITH(vec2, float, 2) REF_ARITH(vec3, float, 3) REF_ARITH(vec4, float, 4)
REF_ARITH(ivec2, int, 2) REF_ARITH(ivec3, int, 3) REF_ARITH(ivec4, int, 4)
REF_ARITH(uvec2, uint, 2) REF_ARITH(uvec3, uint, 3) REF_ARITH(uvec4, uint, 4)
REF_INTOPS(ivec2, int, 2) REF_INTOPS(ivec3, int, 3) REF_INTOPS(ivec4, int, 4)
REF_INTOPS(uvec2, uint, 2) REF_INTOPS(uvec3, uint, 3) REF_INTOPS(uvec4, uint, 4)

// GLSL converts ivec -> vec implicitly in mixed arithmetic (vec2 / ivec2 ...)
#define REF_MIXED(V, I, N, op) \
    inline V operator op(const V& a, const I& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = a.d[i] op float(b.d[i]); return r; } \
    inline V operator op(const I& a, const V& b) { V r; for (int i = 0; i < N; ++i) r.d[i] = float(a.d[i]) op b.d[i]; return r; }
REF_MIXED(vec2, ivec2, 2, +) REF_MIXED(vec2, ivec2, 2, -) REF_MIXED(vec2, ivec2, 2, *) REF_MIXED(vec2, ivec2, 2, /)
REF_MIXED(vec3, ivec3, 3, +) REF_MIXED(vec3, ivec3, 3, -) REF_MIXED(vec3, ivec3, 3, *) REF_MIXED(vec3, ivec3, 3, /)

// ---- elementary functions -----------------------------------------------------------------------------------
#ifdef REF_MATH_DET
inline float exp(float x) { return sky_det_expf(x); }
inline float sin(float x) { return sky_det_sinf(x); }
inline float cos(float x) { return sky_det_cosf(x); }
inline float acos(float x) { return sky_det_acosf(x); }
inline float asin(float x) { return sky_det_asinf(x); }
inline float pow(float x, float y) { return y == 1.5f ? sky_det_pow15f(x) : std::pow(x, y); }
#else
inline float exp(float x) { return std::exp(x); }
inline float sin(float x) { return std::sin(x); }
inline float cos(float x) { return std::cos(x); }
inline float acos(float x) { return std::acos(x); }
inline float asin(float x) { return std::asin(x); }
inline float pow(float x, float y) { return std::pow(x, y); }
#endif
inline float sqrt(float x) { return std::sqrt(x); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
inline float log(float x) { return std::log(x); }
inline float log2(float x) { return std::log2(x); }
inline float exp2(float x) { return std::exp2(x); }
inline float tan(float x) { return std::tan(x); }
#ifdef REF_ATAN_DET   // the ground pass (prog_earth.cpp): atan from include/sky_detmath.h like the oracle and the kernel
inline float atan(float y, float x) { return sky_det_atan2f(y, x); }
#else
inline float atan(float y, float x) { return std::atan2(y, x); }
#endif
inline float abs(float x) { return std::fabs(x); }
inline int abs(int x) { return x < 0 ? -x : x; }
inline float floor(float x) { return std::floor(x); }
inline float ceil(float x) { return std::ceil(x); }
inline float round(float x) { return std::nearbyint(x); }
inline float fract(float x) { return x - std::floor(x); }
inline float sign(float x) { return x > 0.0f ? 1.0f : x < 0.0f ? -1.0f : 0.0f; }
inline float radians(float d) { return d * 0.01745329251994329576923690768489f; }
// min / max / clamp: IEEE minNum / maxNum like the GPU's FMNMX (see oracle/glsl.h)
// (templates: the shaders mix int literals into float calls -- clamp(x, 0, 1), max(d, 0) -- which GLSL converts implicitly)
template <class A, class B> constexpr bool both_arith = std::is_arithmetic_v<A> && std::is_arithmetic_v<B>;
template <class A, class B> using promoted = std::conditional_t<std::is_floating_point_v<A> || std::is_floating_point_v<B>, float,
                                             std::conditional_t<std::is_unsigned_v<A> && std::is_unsigned_v<B>, uint, int>>;
template <class A, class B, class = std::enable_if_t<both_arith<A, B>>> promoted<A, B> max(A a, B b) {
    using P = promoted<A, B>;
    if constexpr (std::is_same_v<P, float>) return std::fmax(float(a), float(b)); else return P(a) < P(b) ? P(b) : P(a);
}
template <class A, class B, class = std::enable_if_t<both_arith<A, B>>> promoted<A, B> min(A a, B b) {
    using P = promoted<A, B>;
    if constexpr (std::is_same_v<P, float>) return std::fmin(float(a), float(b)); else return P(b) < P(a) ? P(b) : P(a);
}
template <class A, class B, class C, class = std::enable_if_t<both_arith<A, B> && std::is_arithmetic_v<C>>>
promoted<promoted<A, B>, C> clamp(A x, B lo, C hi) { return min(max(x, lo), hi); }
template <class A, class B, class C, class = std::enable_if_t<both_arith<A, B> && std::is_arithmetic_v<C>>>
float mix(A a, B b, C t) { return float(a) * (1.0f - float(t)) + float(b) * float(t); }
inline float step(float edge, float x) { return x < edge ? 0.0f : 1.0f; }
inline float smoothstep(float e0, float e1, float x) {
    float t = clamp((x - e0) / (e1 - e0), 0.0f, 1.0f);
    return t * t * (3.0f - 2.0f * t);
}
inline bool all(bool b) { return b; }
inline bool any(bool b) { return b; }
inline float mod(float x, float y) { return x - y * std::floor(x / y); }

#define REF_MAP1(T, N, name) inline T name(const T& a) { T r; for (int i = 0; i < N; ++i) r.d[i] = name(a.d[i]); return r; }
#define REF_FLOATFUNCS(T, N)                                                                                    \
    REF_MAP1(T, N, exp) REF_MAP1(T, N, sqrt) REF_MAP1(T, N, abs) REF_MAP1(T, N, floor) REF_MAP1(T, N, ceil)   \
    REF_MAP1(T, N, fract) REF_MAP1(T, N, sin) REF_MAP1(T, N, cos) REF_MAP1(T, N, log) REF_MAP1(T, N, exp2)    \
    inline T min(const T& a, const T& b) { T r; for (int i = 0; i < N; ++i) r.d[i] = min(a.d[i], b.d[i]); return r; } \
    inline T max(const T& a, const T& b) { T r; for (int i = 0; i < N; ++i) r.d[i] = max(a.d[i], b.d[i]); return r; } \
    inline T min(const T& a, float b) { T r; for (int i = 0; i < N; ++i) r.d[i] = min(a.d[i], b); return r; }  \
    inline T max(const T& a, float b) { T r; for (int i = 0; i < N; ++i) r.d[i] = max(a.d[i], b); return r; }  \
    inline T clamp(const T& a, const T& lo, const T& hi) { return min(max(a, lo), hi); }                        \
    inline T clamp(const T& a, float lo, float hi) { T r; for (int i = 0; i < N; ++i) r.d[i] = clamp(a.d[i], lo, hi); return r; } \
    inline T mix(const T& a, const T& b, float t) { return a * (1.0f - t) + b * t; }                            \
    inline T mix(const T& a, const T& b, const T& t) { T r; for (int i = 0; i < N; ++i) r.d[i] = mix(a.d[i], b.d[i], t.d[i]); return r; } \
    inline T pow(const T& a, const T& b) { T r; for (int i = 0; i < N; ++i) r.d[i] = pow(a.d[i], b.d[i]); return r; } \
    inline float dot(const T& a, const T& b) { float s = a.d[0] * b.d[0]; for (int i = 1; i < N; ++i) s += a.d[i] * b.d[i]; return s; } \
    inline float length(const T& a) { return std::sqrt(dot(a, a)); }                                            \
    inline float distance(const T& a, const T& b) { return length(a - b); }                                     \
    inline T normalize(const T& a) { return a / length(a); }
REF_FLOATFUNCS(vec2, 2) REF_FLOATFUNCS(vec3, 3) REF_FLOATFUNCS(vec4, 4)
inline vec3 cross(const vec3& a, const vec3& b) { return vec3(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y); }
inline ivec2 min(const ivec2& a, const ivec2& b) { return ivec2(min(a.x, b.x), min(a.y, b.y)); }
inline ivec2 max(const ivec2& a, const ivec2& b) { return ivec2(max(a.x, b.x), max(a.y, b.y)); }
inline ivec2 clamp(const ivec2& a, const ivec2& lo, const ivec2& hi) { return min(max(a, lo), hi); }
inline ivec3 min(const ivec3& a, const ivec3& b) { return ivec3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
inline ivec3 max(const ivec3& a, const ivec3& b) { return ivec3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
inline ivec3 clamp(const ivec3& a, const ivec3& lo, const ivec3& hi) { return min(max(a, lo), hi); }

// ---- matrices: column-major, m[c][r] (GLSL / glm layout) -------------------------------------------------------
struct mat4 {
    vec4 c[4];
    mat4() {}
    explicit mat4(const float* p) { for (int i = 0; i < 4; ++i) c[i] = vec4(p[4 * i], p[4 * i + 1], p[4 * i + 2], p[4 * i + 3]); }
    vec4& operator[](int i) { return c[i]; }
    const vec4& operator[](int i) const { return c[i]; }
};
inline vec4 operator*(const mat4& m, const vec4& v) {  // GLSL: sum over columns, left to right
    vec4 r;
    for (int i = 0; i < 4; ++i) r.d[i] = m.c[0].d[i] * v.x + m.c[1].d[i] * v.y + m.c[2].d[i] * v.z + m.c[3].d[i] * v.w;
    return r;
}
struct mat3 {
    vec3 c[3];
    mat3() {}
    explicit mat3(const float* p) { for (int i = 0; i < 3; ++i) c[i] = vec3(p[3 * i], p[3 * i + 1], p[3 * i + 2]); }
    vec3& operator[](int i) { return c[i]; }
    const vec3& operator[](int i) const { return c[i]; }
};
inline vec3 operator*(const mat3& m, const vec3& v) {
    vec3 r;
    for (int i = 0; i < 3; ++i) r.d[i] = m.c[0].d[i] * v.x + m.c[1].d[i] * v.y + m.c[2].d[i] * v.z;
    return r;
}

// ---- images and samplers ------------------------------------------------------------------------------------
enum ImageFormat { FMT_RGBA32F, FMT_RG32F, FMT_R32F, FMT_RGBA16F, FMT_RGBA8, FMT_RG8, FMT_R8, FMT_R16, FMT_RG16 };
inline float unorm16_round(float x);
inline float half_round(float x) { return (float)(_Float16)x; }
inline float unorm8_round(float x) { return std::nearbyint(clamp(x, 0.0f, 1.0f) * 255.0f) / 255.0f; }
inline float unorm16_round(float x) { return std::nearbyint(clamp(x, 0.0f, 1.0f) * 65535.0f) / 65535.0f; }

// One level of texel storage, always 4 floats per texel in memory (the format decides the rounding of a store).
struct Image {
    float* data = nullptr;  // [d][h][w][4]
    int w = 0, h = 0, d = 1;
    ImageFormat fmt = FMT_RGBA32F;
    vec4 load(int x, int y, int z = 0) const {
        const float* p = data + ((size_t(z) * h + y) * w + x) * 4;
        return vec4(p[0], p[1], p[2], p[3]);
    }
    void store(int x, int y, int z, vec4 v) {
        if (x < 0 || y < 0 || z < 0 || x >= w || y >= h || z >= d) return;  // GL drops out-of-range image stores
        switch (fmt) {
            case FMT_RGBA16F: for (int i = 0; i < 4; ++i) v.d[i] = half_round(v.d[i]); break;
            case FMT_RGBA8: for (int i = 0; i < 4; ++i) v.d[i] = unorm8_round(v.d[i]); break;
            case FMT_RG8: v = vec4(unorm8_round(v.x), unorm8_round(v.y), 0.0f, 1.0f); break;
            case FMT_R8: v = vec4(unorm8_round(v.x), 0.0f, 0.0f, 1.0f); break;
            case FMT_R16: v = vec4(unorm16_round(v.x), 0.0f, 0.0f, 1.0f); break;
            case FMT_RG16: v = vec4(unorm16_round(v.x), unorm16_round(v.y), 0.0f, 1.0f); break;
            case FMT_RG32F: v = vec4(v.x, v.y, 0.0f, 1.0f); break;
            case FMT_R32F: v = vec4(v.x, 0.0f, 0.0f, 1.0f); break;
            default: break;
        }
        float* p = data + ((size_t(z) * h + y) * w + x) * 4;
        p[0] = v.x; p[1] = v.y; p[2] = v.z; p[3] = v.w;
    }
};
struct image2D : Image {};
struct image3D : Image {};
struct imageCube : Image {};  // d = 6 faces
struct uimage2D : Image {};   // r8ui: the integer lives in .x
inline ivec2 imageSize(const image2D& im) { return ivec2(im.w, im.h); }
inline ivec3 imageSize(const image3D& im) { return ivec3(im.w, im.h, im.d); }
inline ivec2 imageSize(const imageCube& im) { return ivec2(im.w, im.h); }
inline ivec2 imageSize(const uimage2D& im) { return ivec2(im.w, im.h); }
inline void imageStore(Image& im, const ivec2& p, const vec4& v) { im.store(p.x, p.y, 0, v); }
inline void imageStore(Image& im, const ivec3& p, const vec4& v) { im.store(p.x, p.y, p.z, v); }
inline vec4 imageLoad(const image2D& im, const ivec2& p) { return im.load(p.x, p.y, 0); }
inline uvec4 imageLoad(const uimage2D& im, const ivec2& p) { vec4 v = im.load(p.x, p.y, 0); return uvec4(uint(v.x), uint(v.y), uint(v.z), uint(v.w)); }
inline void imageStore(uimage2D& im, const ivec2& p, const uvec4& v) { im.store(p.x, p.y, 0, vec4(float(v.x), float(v.y), float(v.z), float(v.w))); }
inline vec4 imageLoad(const image3D& im, const ivec3& p) { return im.load(p.x, p.y, p.z); }

enum Wrap { CLAMP_TO_EDGE, REPEAT, CLAMP_TO_BORDER };
enum Filter { NEAREST, LINEAR };
// A texture = mip chain of Images + sampler state (src/Base/src/Samplers.cpp supplies the states the host binds)
struct Sampler {
    std::vector<Image> levels;
    Wrap wrap = CLAMP_TO_EDGE;
    Filter mag = LINEAR, min_filter = LINEAR;  // min_filter NEAREST here means NEAREST_MIPMAP_NEAREST when there are mips
    vec4 border = vec4(0.0f);
    int wrap_index(int i, int n, bool& is_border) const {
        if (wrap == REPEAT) { int m = i % n; return m < 0 ? m + n : m; }
        if (wrap == CLAMP_TO_BORDER && (i < 0 || i >= n)) { is_border = true; return 0; }
        return i < 0 ? 0 : i >= n ? n - 1 : i;
    }
    vec4 texel(const Image& im, int x, int y, int z) const {
        bool b = false;
        x = wrap_index(x, im.w, b); y = wrap_index(y, im.h, b); z = im.d > 1 ? wrap_index(z, im.d, b) : 0;
        return b ? border : im.load(x, y, z);
    }
    // GL 4.6 section 8.14.2: u*size - 0.5, floor, fract; exact fp32 weights, x first then y then z
    vec4 linear(const Image& im, float u, float v, float w) const {
        float x = u * float(im.w) - 0.5f, y = v * float(im.h) - 0.5f;
        float fx = std::floor(x), fy = std::floor(y);
        float a = x - fx, b = y - fy;
        int i0 = int(fx), j0 = int(fy);
        if (im.d > 1) {
            float z = w * float(im.d) - 0.5f, fz = std::floor(z), c = z - fz;
            int k0 = int(fz);
            vec4 acc(0.0f);
            for (int dk = 0; dk < 2; ++dk) for (int dj = 0; dj < 2; ++dj) for (int di = 0; di < 2; ++di) {
                float wt = (di ? a : 1.0f - a) * (dj ? b : 1.0f - b) * (dk ? c : 1.0f - c);
                acc += wt * texel(im, i0 + di, j0 + dj, k0 + dk);
            }
            return acc;
        }
        vec4 t00 = texel(im, i0, j0, 0), t10 = texel(im, i0 + 1, j0, 0), t01 = texel(im, i0, j0 + 1, 0), t11 = texel(im, i0 + 1, j0 + 1, 0);
        return (1.0f - a) * (1.0f - b) * t00 + a * (1.0f - b) * t10 + (1.0f - a) * b * t01 + a * b * t11;
    }
    vec4 nearest(const Image& im, float u, float v, float w) const {
        return texel(im, int(std::floor(u * float(im.w))), int(std::floor(v * float(im.h))), im.d > 1 ? int(std::floor(w * float(im.d))) : 0);
    }
    vec4 sample(float u, float v, float w, float lod) const {  // section 8.14.3 for (mag LINEAR, min NEAREST_MIPMAP_NEAREST) or no mips
        if (levels.size() == 1 || !(lod > 0.5f)) return mag == LINEAR ? linear(levels[0], u, v, w) : nearest(levels[0], u, v, w);
        int q = int(levels.size()) - 1;
        int d = (lod <= float(q) + 0.5f) ? int(std::ceil(lod + 0.5f)) - 1 : q;
        d = d < 0 ? 0 : d > q ? q : d;
        return min_filter == LINEAR ? linear(levels[d], u, v, w) : nearest(levels[d], u, v, w);
    }
};
struct sampler2D : Sampler {};
struct sampler3D : Sampler {};
struct samplerCube : Sampler {};      // levels[l].d = 6 faces; see textureCubeLod0
struct sampler2DShadow : Sampler {};  // mesh shadow map (VOLUMETRIC_LIGHT_ENABLE permutation)
// GL 4.6 section 8.23 with COMPARE_REF_TO_TEXTURE / LEQUAL and LINEAR filtering: the comparison is made per texel and the
// 0/1 results are blended with the bilinear weights (what every implementation does; the spec leaves it open)
inline float texture(const sampler2DShadow& s, const vec3& p) {
    if (s.levels.empty()) return 1.0f;
    const Image& im = s.levels[0];
    float x = p.x * float(im.w) - 0.5f, y = p.y * float(im.h) - 0.5f;
    float fx = std::floor(x), fy = std::floor(y);
    float a = x - fx, b = y - fy;
    int i0 = int(fx), j0 = int(fy);
    auto cmp = [&](int i, int j) { return p.z <= s.texel(im, i, j, 0).x ? 1.0f : 0.0f; };
    return (1.0f - a) * (1.0f - b) * cmp(i0, j0) + a * (1.0f - b) * cmp(i0 + 1, j0) + (1.0f - a) * b * cmp(i0, j0 + 1) + a * b * cmp(i0 + 1, j0 + 1);
}
inline vec4 texture(const Sampler& s, const vec2& uv) { return s.sample(uv.x, uv.y, 0.0f, 0.0f); }
inline vec4 texture2D(const Sampler& s, const vec2& uv) { return s.sample(uv.x, uv.y, 0.0f, 0.0f); }
inline vec4 texture(const Sampler& s, const vec3& uvw) { return s.sample(uvw.x, uvw.y, uvw.z, 0.0f); }
inline vec4 textureLod(const Sampler& s, const vec2& uv, float lod) { return s.sample(uv.x, uv.y, 0.0f, lod); }
inline vec4 textureLod(const Sampler& s, const vec3& uvw, float lod) { return s.sample(uvw.x, uvw.y, uvw.z, lod); }
// textureGather (spec 8.14.2 / 8.15): the 2x2 footprint LINEAR filtering would use; (i0,j1) (i1,j1) (i1,j0) (i0,j0).
// VolumetricCloudReconstruct.comp:38-39 and VolumetricCloudUpscale.comp:17 aim the coordinate EXACTLY at a texel corner
// ((q +- 0.5) / size), where fp32 rounding of u * size - 0.5 would pick the block at random.  The texture unit works on
// a fixed-point coordinate with 8 sub-texel bits (round to nearest), which makes the intended block the answer; so does
// this shim (and the oracle and the kernels, which index that block directly).
inline vec4 textureGather(const Sampler& s, const vec2& uv, int comp = 0) {
    const Image& im = s.levels[0];
    auto base = [](float u, int n) { return int(std::floor(std::nearbyint((u * float(n) - 0.5f) * 256.0f) / 256.0f)); };
    int i0 = base(uv.x, im.w), j0 = base(uv.y, im.h);
    return vec4(s.texel(im, i0, j0 + 1, 0).d[comp], s.texel(im, i0 + 1, j0 + 1, 0).d[comp], s.texel(im, i0 + 1, j0, 0).d[comp], s.texel(im, i0, j0, 0).d[comp]);
}
// texelFetch outside the image is undefined in GL (zero with robust access); convention here and in the oracle / kernels:
// the edge texel (only VolumetricCloudIndexGen.comp:31 gets there, for the tiles on the image border)
inline vec4 texelFetch(const Sampler& s, const ivec2& p, int lod) {
    const Image& im = s.levels[lod];
    return im.load(clamp(p.x, 0, im.w - 1), clamp(p.y, 0, im.h - 1), 0);
}
inline vec4 texelFetch(const Sampler& s, const ivec3& p, int lod) {
    const Image& im = s.levels[lod];
    return im.load(clamp(p.x, 0, im.w - 1), clamp(p.y, 0, im.h - 1), clamp(p.z, 0, im.d - 1));
}
// samplerCube inside a compute shader: implicit derivatives are undefined, so the LOD is the driver's choice.  Convention
// (same as the oracle and the kernels): level 0, bilinear inside the face the GL cube-map table
// (spec 8.13) selects, with SEAMLESS filtering at the face edges (GL 4.6 8.14.1: the reference calls
// glEnable(GL_TEXTURE_CUBE_MAP_SEAMLESS), AtmosphereRenderer.cpp:151) -- include/sky_cubemap.h, the rule the oracle and the kernels share.
inline vec4 textureCubeLevel(const Image& env, const vec3& dir) {
    float ax = std::fabs(dir.x), ay = std::fabs(dir.y), az = std::fabs(dir.z);
    int face; float sc, tc, ma;
    if (ax >= ay && ax >= az) { ma = ax; if (dir.x >= 0) { face = 0; sc = -dir.z; tc = -dir.y; } else { face = 1; sc = dir.z; tc = -dir.y; } }
    else if (ay >= az)        { ma = ay; if (dir.y >= 0) { face = 2; sc = dir.x; tc = dir.z; } else { face = 3; sc = dir.x; tc = -dir.z; } }
    else                      { ma = az; if (dir.z >= 0) { face = 4; sc = dir.x; tc = -dir.y; } else { face = 5; sc = -dir.x; tc = -dir.y; } }
    float ss = 0.5f * (sc / ma + 1.0f), tt = 0.5f * (tc / ma + 1.0f);
    int n = env.w;
    float u = ss * float(n) - 0.5f, v = tt * float(n) - 0.5f;
    float fu = std::floor(u), fv = std::floor(v);
    int i0 = int(fu), j0 = int(fv);
    float a = u - fu, b = v - fv;
    return sky_cube_bilinear<vec4>(n, face, i0, j0, a, b, [&](int f, int i, int j) { return env.load(i, j, f); });
}
inline vec4 texture(const samplerCube& s, const vec3& dir) { return textureCubeLevel(s.levels[0], dir); }
// textureLod on a cube with a LINEAR_MIPMAP_LINEAR sampler (Samplers::GetAnisotropySampler, src/Base/src/Samplers.cpp:33-41; an
// explicit LOD has no footprint, so the anisotropy limit is moot): GL 4.6 section 8.14.3 -- the LOD is clamped to
// [0, q], levels floor(lod) and floor(lod) + 1 are filtered bilinearly inside the selected face and blended with the exact fp32
// fraction, mix(t0, t1, frac).  A fraction of exactly 0 takes the lower level alone.
inline vec4 textureLod(const samplerCube& s, const vec3& dir, float lod) {
    const int q = int(s.levels.size()) - 1;
    float l = lod < 0.0f ? 0.0f : lod > float(q) ? float(q) : lod;
    float fl = std::floor(l);
    int l0 = int(fl);
    float f = l - fl;
    vec4 t0 = textureCubeLevel(s.levels[l0], dir);
    if (!(f > 0.0f)) return t0;
    vec4 t1 = textureCubeLevel(s.levels[l0 + 1 > q ? q : l0 + 1], dir);
    return t0 * (1.0f - f) + t1 * f;
}
inline ivec2 textureSize(const sampler2D& s, int lod) { return ivec2(s.levels[lod].w, s.levels[lod].h); }
inline ivec3 textureSize(const sampler3D& s, int lod) { return ivec3(s.levels[lod].w, s.levels[lod].h, s.levels[lod].d); }

// ---- fragment programs: discard, gl_FragDepth, derivatives, textureGrad -----------------------------------------------------
// A fragment program runs per 2x2 pixel quad (GL 4.6 section 15.1; helper invocations keep running after `discard` so that
// their neighbours' derivatives stay defined).  Here the quad is evaluated twice: a RECORD pass, in which the k-th dFdx / dFdy
// call of each of the four pixels stores its argument, and a REPLAY pass, in which the same call returns the fine difference of
// the recorded values inside the quad row / column.  Valid while no derivative argument depends on an earlier derivative and the
// four pixels take the same path to every derivative call -- true for EarthRender.frag, the one user.
struct QuadState {
    int mode = 0;       // 0 record, 1 replay
    int pixel = 0;      // (y & 1) << 1 | (x & 1)
    int call = 0;
    bool discarded = false;
    float frag_depth = 0.0f;
    std::vector<std::array<float, 4>> values;
};
inline thread_local QuadState g_quad;
#define gl_FragDepth (ref::g_quad.frag_depth)
#define REF_DISCARD_HELPER do { ref::g_quad.discarded = true; } while (0)
inline float quad_derivative(float v, int lo, int hi) {
    QuadState& q = g_quad;
    const size_t k = size_t(q.call++);
    if (q.mode == 0) {
        if (q.values.size() <= k) q.values.resize(k + 1);
        q.values[k][q.pixel] = v;
        return 0.0f;
    }
    return q.values[k][hi] - q.values[k][lo];
}
inline float dFdx(float v) { const int row = g_quad.pixel & 2; return quad_derivative(v, row, row | 1); }
inline float dFdy(float v) { const int col = g_quad.pixel & 1; return quad_derivative(v, col, col | 2); }
// textureGrad on a mip-mapped 2-D texture with the anisotropic LINEAR_MIPMAP_LINEAR sampler of Earth.cpp:34-42 (REPEAT in s,
// CLAMP_TO_EDGE in t): include/sky_texgrad.h, the rule the oracle and the kernel share.  The levels hold decoded (linear) texels.
inline vec4 textureGrad(const sampler2D& s, const vec2& P, const vec2& dPdx, const vec2& dPdy) {
    const Image& l0 = s.levels[0];
    return sky_texture_grad_2d<vec4>(l0.w, l0.h, int(s.levels.size()), P.x, P.y, dPdx.x, dPdx.y, dPdy.x, dPdy.y, 16.0f,
                                     [&](int l, int i, int j) { return s.levels[l].load(i, j, 0); });
}

// ---- compute dispatch ------------------------------------------------------------------------------------------
struct Builtins {
    uvec3 global_id, local_id, group_id;
    uint local_index = 0;
    vec4 frag_coord;
    std::barrier<>* group_barrier = nullptr;
};
inline thread_local Builtins g_builtins;
#define gl_GlobalInvocationID (ref::g_builtins.global_id)
#define gl_LocalInvocationID (ref::g_builtins.local_id)
#define gl_WorkGroupID (ref::g_builtins.group_id)
#define gl_LocalInvocationIndex (ref::g_builtins.local_index)
#define gl_FragCoord (ref::g_builtins.frag_coord)
inline void barrier() { if (g_builtins.group_barrier) g_builtins.group_barrier->arrive_and_wait(); }

// glDispatchCompute(gx, gy, gz) of a program whose local size is (lx, ly, lz).  Programs without barrier() run their
// invocations one after the other; programs with barriers (`uses_barrier`) run one OS thread per invocation of a work group,
// work groups one after the other (`shared` variables are plain statics, so two groups must never overlap).
inline void dispatch(const std::function<void()>& shader_main, int gx, int gy, int gz, int lx, int ly, int lz, bool uses_barrier) {
    const int local = lx * ly * lz;
    auto run = [&](int wx, int wy, int wz, int l, std::barrier<>* bar) {
        int ix = l % lx, iy = (l / lx) % ly, iz = l / (lx * ly);
        g_builtins.group_id = uvec3(wx, wy, wz);
        g_builtins.local_id = uvec3(ix, iy, iz);
        g_builtins.local_index = uint(l);
        g_builtins.global_id = uvec3(wx * lx + ix, wy * ly + iy, wz * lz + iz);
        g_builtins.group_barrier = bar;
        shader_main();
        if (bar) bar->arrive_and_drop();
    };
    if (!uses_barrier) {
        // invocations are independent (no shared variables, disjoint image stores): work groups across the host cores
        const long groups = long(gx) * gy * gz;
#pragma omp parallel for schedule(dynamic, 4)
        for (long g = 0; g < groups; ++g) {
            int wx = int(g % gx), wy = int((g / gx) % gy), wz = int(g / (long(gx) * gy));
            for (int l = 0; l < local; ++l) run(wx, wy, wz, l, nullptr);
        }
        return;
    }
    for (int wz = 0; wz < gz; ++wz) for (int wy = 0; wy < gy; ++wy) for (int wx = 0; wx < gx; ++wx) {
        std::barrier<> bar(local);
        std::vector<std::thread> pool;
        for (int l = 0; l < local; ++l) pool.emplace_back(run, wx, wy, wz, l, &bar);
        for (auto& t : pool) t.join();
    }
}

}  // namespace ref
