// Small fp32 vector/matrix layer for the host-side uniform maths (the reference uses glm 0.9.9;
// the conventions below are glm's defaults: column-major, right-handed, clip z in [-1, 1]).
#pragma once
#include <cmath>
#include <cstring>

namespace skyhost {

struct vec2 { float x = 0, y = 0; };
struct dvec2 { double x = 0, y = 0; };
struct vec3 {
    float x = 0, y = 0, z = 0;
    vec3() = default;
    vec3(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};
struct vec4 {
    float x = 0, y = 0, z = 0, w = 0;
    vec4() = default;
    vec4(float x_, float y_, float z_, float w_) : x(x_), y(y_), z(z_), w(w_) {}
    vec4(vec3 v, float w_) : x(v.x), y(v.y), z(v.z), w(w_) {}
    float& operator[](int i) { return (&x)[i]; }
    float operator[](int i) const { return (&x)[i]; }
};

inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return a * s; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a * (1.0f / std::sqrt(dot(a, a))); }  // glm: v * inversesqrt(dot(v, v))
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }
inline float distance(vec3 a, vec3 b) { return length(a - b); }
inline float radians(float deg) { return deg * 0.01745329251994329576923690768489f; }

struct mat3 {
    float m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};  // m[c*3 + r]
};
inline vec3 operator*(const mat3& a, vec3 v) {
    return {a.m[0] * v.x + a.m[3] * v.y + a.m[6] * v.z, a.m[1] * v.x + a.m[4] * v.y + a.m[7] * v.z,
            a.m[2] * v.x + a.m[5] * v.y + a.m[8] * v.z};
}

struct mat4 {
    float m[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};  // m[c*4 + r]
    float& at(int c, int r) { return m[c * 4 + r]; }
    float at(int c, int r) const { return m[c * 4 + r]; }
    static mat4 from_columns(vec4 c0, vec4 c1, vec4 c2, vec4 c3) {
        mat4 r;
        for (int i = 0; i < 4; ++i) { r.m[i] = c0[i]; r.m[4 + i] = c1[i]; r.m[8 + i] = c2[i]; r.m[12 + i] = c3[i]; }
        return r;
    }
    void store(float* dst) const { std::memcpy(dst, m, sizeof(m)); }
};
inline mat3 upper3(const mat4& a) {
    mat3 r;
    for (int c = 0; c < 3; ++c) for (int row = 0; row < 3; ++row) r.m[c * 3 + row] = a.m[c * 4 + row];
    return r;
}
inline mat4 operator*(const mat4& a, const mat4& b) {
    mat4 r;
    for (int c = 0; c < 4; ++c)
        for (int row = 0; row < 4; ++row) {
            float s = 0.0f;
            for (int k = 0; k < 4; ++k) s += a.m[k * 4 + row] * b.m[c * 4 + k];
            r.m[c * 4 + row] = s;
        }
    return r;
}
inline vec4 operator*(const mat4& a, vec4 v) {
    vec4 r;
    for (int i = 0; i < 4; ++i) r[i] = a.m[i] * v.x + a.m[4 + i] * v.y + a.m[8 + i] * v.z + a.m[12 + i] * v.w;
    return r;
}

// glm::inverse(mat4): adjugate / determinant, evaluated in fp32 like glm does.
inline mat4 inverse(const mat4& a) {
    const float* m = a.m;
    float inv[16];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    float det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    float inv_det = 1.0f / det;
    mat4 r;
    for (int i = 0; i < 16; ++i) r.m[i] = inv[i] * inv_det;
    return r;
}

// glm::perspective (RH, clip z in [-1,1])
inline mat4 perspective(float fovy, float aspect, float zNear, float zFar) {
    float tanHalfFovy = std::tan(fovy / 2.0f);
    mat4 r;
    std::memset(r.m, 0, sizeof(r.m));
    r.at(0, 0) = 1.0f / (aspect * tanHalfFovy);
    r.at(1, 1) = 1.0f / tanHalfFovy;
    r.at(2, 2) = -(zFar + zNear) / (zFar - zNear);
    r.at(2, 3) = -1.0f;
    r.at(3, 2) = -(2.0f * zFar * zNear) / (zFar - zNear);
    return r;
}
// glm::lookAt (RH)
inline mat4 lookAt(vec3 eye, vec3 center, vec3 up) {
    vec3 f = normalize(center - eye);
    vec3 s = normalize(cross(f, up));
    vec3 u = cross(s, f);
    mat4 r;
    r.at(0, 0) = s.x; r.at(1, 0) = s.y; r.at(2, 0) = s.z;
    r.at(0, 1) = u.x; r.at(1, 1) = u.y; r.at(2, 1) = u.z;
    r.at(0, 2) = -f.x; r.at(1, 2) = -f.y; r.at(2, 2) = -f.z;
    r.at(3, 0) = -dot(s, eye); r.at(3, 1) = -dot(u, eye); r.at(3, 2) = dot(f, eye);
    return r;
}
// glm::ortho (RH, clip z in [-1,1])
inline mat4 ortho(float left, float right, float bottom, float top, float zNear, float zFar) {
    mat4 r;
    r.at(0, 0) = 2.0f / (right - left);
    r.at(1, 1) = 2.0f / (top - bottom);
    r.at(2, 2) = -2.0f / (zFar - zNear);
    r.at(3, 0) = -(right + left) / (right - left);
    r.at(3, 1) = -(top + bottom) / (top - bottom);
    r.at(3, 2) = -(zFar + zNear) / (zFar - zNear);
    return r;
}
inline mat4 translate(const mat4& m, vec3 v) {
    mat4 r = m;
    for (int i = 0; i < 4; ++i) r.m[12 + i] = m.m[i] * v.x + m.m[4 + i] * v.y + m.m[8 + i] * v.z + m.m[12 + i];
    return r;
}
inline mat4 rotate(const mat4& m, float angle, vec3 axis_in) {
    float c = std::cos(angle), s = std::sin(angle);
    vec3 axis = normalize(axis_in);
    vec3 temp = axis * (1.0f - c);
    float R[3][3];
    R[0][0] = c + temp.x * axis.x; R[0][1] = temp.x * axis.y + s * axis.z; R[0][2] = temp.x * axis.z - s * axis.y;
    R[1][0] = temp.y * axis.x - s * axis.z; R[1][1] = c + temp.y * axis.y; R[1][2] = temp.y * axis.z + s * axis.x;
    R[2][0] = temp.z * axis.x + s * axis.y; R[2][1] = temp.z * axis.y - s * axis.x; R[2][2] = c + temp.z * axis.z;
    mat4 r = m;
    for (int col = 0; col < 3; ++col)
        for (int i = 0; i < 4; ++i) r.m[col * 4 + i] = m.m[i] * R[col][0] + m.m[4 + i] * R[col][1] + m.m[8 + i] * R[col][2];
    return r;
}

// src/Base/src/Utils.cpp:23-31
inline vec3 FromThetaPhiToDirection(float theta, float phi) {
    float cos_theta = std::cos(theta), sin_theta = std::sin(theta);
    float cos_phi = std::cos(phi), sin_phi = std::sin(phi);
    return {cos_phi * sin_theta, cos_theta, sin_phi * sin_theta};
}

}  // namespace skyhost
