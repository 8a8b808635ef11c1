/*
 * glsl_math.h — the subset of GLSL 4.60 vector/matrix semantics the reference's
 * shaders use, in plain fp32 C++ (TEST INFRASTRUCTURE — part of the CPU oracle).
 *
 * Built with -ffp-contract=off so that every expression is evaluated as written,
 * left to right, one rounding per operation.  Matrices are column-major like GLSL:
 * m.c[i] is column i; M * v = sum_i c[i] * v[i]; v * M = (dot(v, c[0]), dot(v, c[1]), ...).
 */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace glsl
{

struct vec2
{
    float x, y;
};
struct vec3
{
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
struct vec4
{
    float x, y, z, w;
};

inline vec2 V2(float x, float y) { return vec2 { x, y }; }
inline vec3 V3(float x, float y, float z) { return vec3 { x, y, z }; }
inline vec3 V3(float s) { return vec3 { s, s, s }; }
inline vec4 V4(float x, float y, float z, float w) { return vec4 { x, y, z, w }; }
inline vec4 V4(vec3 v, float w) { return vec4 { v.x, v.y, v.z, w }; }
inline vec3 xyz(vec4 v) { return vec3 { v.x, v.y, v.z }; }

inline vec2 operator+(vec2 a, vec2 b) { return { a.x + b.x, a.y + b.y }; }
inline vec2 operator-(vec2 a, vec2 b) { return { a.x - b.x, a.y - b.y }; }
inline vec2 operator*(vec2 a, float s) { return { a.x * s, a.y * s }; }
inline vec2 operator*(float s, vec2 a) { return { s * a.x, s * a.y }; }
inline vec2 operator/(vec2 a, vec2 b) { return { a.x / b.x, a.y / b.y }; }
inline vec2 operator-(vec2 a, float s) { return { a.x - s, a.y - s }; }

inline vec3 operator+(vec3 a, vec3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline vec3 operator-(vec3 a, vec3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline vec3 operator-(vec3 a) { return { -a.x, -a.y, -a.z }; }
inline vec3 operator*(vec3 a, vec3 b) { return { a.x * b.x, a.y * b.y, a.z * b.z }; }
inline vec3 operator/(vec3 a, vec3 b) { return { a.x / b.x, a.y / b.y, a.z / b.z }; }
inline vec3 operator*(vec3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline vec3 operator*(float s, vec3 a) { return { s * a.x, s * a.y, s * a.z }; }
inline vec3 operator/(vec3 a, float s) { return { a.x / s, a.y / s, a.z / s }; }
inline vec3 operator+(vec3 a, float s) { return { a.x + s, a.y + s, a.z + s }; }
inline vec3 operator-(vec3 a, float s) { return { a.x - s, a.y - s, a.z - s }; }
inline vec3 &operator+=(vec3 &a, vec3 b)
{
    a = a + b;
    return a;
}
inline vec3 &operator-=(vec3 &a, vec3 b)
{
    a = a - b;
    return a;
}
inline vec3 &operator*=(vec3 &a, vec3 b)
{
    a = a * b;
    return a;
}
inline vec3 &operator*=(vec3 &a, float s)
{
    a = a * s;
    return a;
}
inline vec3 &operator/=(vec3 &a, float s)
{
    a = a / s;
    return a;
}

inline vec4 operator+(vec4 a, vec4 b) { return { a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w }; }
inline vec4 operator-(vec4 a, vec4 b) { return { a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w }; }
inline vec4 operator*(vec4 a, float s) { return { a.x * s, a.y * s, a.z * s, a.w * s }; }
inline vec4 operator*(vec4 a, vec4 b) { return { a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w }; }

/* dot products sum left to right, as a scalarised GLSL compiler emits them */
inline float dot(vec2 a, vec2 b) { return a.x * b.x + a.y * b.y; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float dot(vec4 a, vec4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
inline vec3 cross(vec3 a, vec3 b)
{
    return { a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y };
}
inline float length(vec3 a) { return std::sqrt(dot(a, a)); }
inline float distance(vec3 a, vec3 b) { return length(a - b); }
inline float inversesqrt(float x) { return 1.0f / std::sqrt(x); }
/* glm / GLSL.std.450 Normalize: v * inversesqrt(dot(v, v)); normalize(0) = NaN (0 * inf) */
inline vec3 normalize(vec3 a) { return a * inversesqrt(dot(a, a)); }

/* NVIDIA FMNMX semantics: if one operand is NaN the other is returned */
inline float min(float a, float b) { return std::fmin(a, b); }
inline float max(float a, float b) { return std::fmax(a, b); }
inline float clamp(float x, float lo, float hi) { return min(max(x, lo), hi); }
inline vec3 max(vec3 a, float b) { return { max(a.x, b), max(a.y, b), max(a.z, b) }; }
inline float mix(float x, float y, float a) { return x * (1.0f - a) + y * a; }
inline vec3 mix(vec3 x, vec3 y, float a) { return x * (1.0f - a) + y * a; }
inline vec4 mix(vec4 x, vec4 y, float a) { return x * (1.0f - a) + y * a; }
inline float abs(float x) { return std::fabs(x); }
inline bool isnan(float x) { return std::isnan(x); }
inline bool isinf(float x) { return std::isinf(x); }

inline vec3 reflect(vec3 I, vec3 N) { return I - 2.0f * dot(N, I) * N; }
/* GLSL refract: returns the zero vector on total internal reflection (Q12) */
inline vec3 refract(vec3 I, vec3 N, float eta)
{
    const float d = dot(N, I);
    const float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f)
        return V3(0.0f);
    return eta * I - (eta * d + std::sqrt(k)) * N;
}

inline uint32_t floatBitsToUint(float f)
{
    uint32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline int32_t floatBitsToInt(float f)
{
    int32_t u;
    std::memcpy(&u, &f, 4);
    return u;
}
inline float uintBitsToFloat(uint32_t u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}
inline float intBitsToFloat(int32_t u)
{
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

struct mat3
{
    vec3 c[3];
};
struct mat4
{
    vec4 c[4];
};
/* GLSL mat3x4: 3 columns of vec4 */
struct mat3x4
{
    vec4 c[3];
};

inline mat3 M3(vec3 a, vec3 b, vec3 c) { return mat3 { { a, b, c } }; }
inline vec3 operator*(const mat3 &m, vec3 v) { return m.c[0] * v.x + m.c[1] * v.y + m.c[2] * v.z; }
/* glm sums the four products pairwise: (c0 x + c1 y) + (c2 z + c3 w) (vendor/glm/glm/detail/type_mat4x4.inl) */
inline vec4 operator*(const mat4 &m, vec4 v) { return (m.c[0] * v.x + m.c[1] * v.y) + (m.c[2] * v.z + m.c[3] * v.w); }
/* row vector times matrix */
inline vec3 operator*(vec4 v, const mat3x4 &m) { return { dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2]) }; }
inline vec4 operator*(vec4 v, const mat4 &m)
{
    return { dot(v, m.c[0]), dot(v, m.c[1]), dot(v, m.c[2]), dot(v, m.c[3]) };
}
/* mat4 * mat3x4 -> mat3x4 (column j = M * B.c[j]); glm sums these left to right (type_mat4x4.inl:469-484) */
inline mat3x4 operator*(const mat4 &m, const mat3x4 &b)
{
    mat3x4 r;
    for (int j = 0; j < 3; j++)
        r.c[j] = m.c[0] * b.c[j].x + m.c[1] * b.c[j].y + m.c[2] * b.c[j].z + m.c[3] * b.c[j].w;
    return r;
}
/* mat4(mat3x4): missing column comes from the identity */
inline mat4 M4(const mat3x4 &m) { return mat4 { { m.c[0], m.c[1], m.c[2], V4(0.0f, 0.0f, 0.0f, 1.0f) } }; }

inline mat4 transpose(const mat4 &m)
{
    mat4 r;
    r.c[0] = V4(m.c[0].x, m.c[1].x, m.c[2].x, m.c[3].x);
    r.c[1] = V4(m.c[0].y, m.c[1].y, m.c[2].y, m.c[3].y);
    r.c[2] = V4(m.c[0].z, m.c[1].z, m.c[2].z, m.c[3].z);
    r.c[3] = V4(m.c[0].w, m.c[1].w, m.c[2].w, m.c[3].w);
    return r;
}

/* inverse(mat3) by the adjugate, the classic closed form */
inline mat3 inverse(const mat3 &m)
{
    const float a00 = m.c[0].x, a01 = m.c[0].y, a02 = m.c[0].z;
    const float a10 = m.c[1].x, a11 = m.c[1].y, a12 = m.c[1].z;
    const float a20 = m.c[2].x, a21 = m.c[2].y, a22 = m.c[2].z;
    const float det = a00 * (a11 * a22 - a21 * a12) - a10 * (a01 * a22 - a21 * a02) + a20 * (a01 * a12 - a11 * a02);
    const float inv = 1.0f / det;
    mat3 r;
    r.c[0] = V3((a11 * a22 - a21 * a12) * inv, -(a01 * a22 - a21 * a02) * inv, (a01 * a12 - a11 * a02) * inv);
    r.c[1] = V3(-(a10 * a22 - a20 * a12) * inv, (a00 * a22 - a20 * a02) * inv, -(a00 * a12 - a10 * a02) * inv);
    r.c[2] = V3((a10 * a21 - a20 * a11) * inv, -(a00 * a21 - a20 * a01) * inv, (a00 * a11 - a10 * a01) * inv);
    return r;
}

/* inverse(mat4) exactly as the reference's vendored glm evaluates it (vendor/glm/glm/detail/func_matrix.inl,
 * compute_inverse<4, 4>): 18 2x2 sub-determinants, four cofactor columns `a*b - c*d + e*f`, the determinant
 * as the pairwise sum (d.x + d.y) + (d.z + d.w) of column 0 times the first cofactor row, then one multiply
 * by 1/det.  GLSL leaves the algorithm to the implementation; the oracle follows glm so that it can be
 * compared bit for bit with the reference's shaders compiled against glm (oracle/_ref/libglsl_ref.so). */
inline mat4 inverse(const mat4 &mm)
{
    const float *a = &mm.c[0].x;
#define M(c, r) a[(c) * 4 + (r)]
    const float Coef00 = M(2, 2) * M(3, 3) - M(3, 2) * M(2, 3);
    const float Coef02 = M(1, 2) * M(3, 3) - M(3, 2) * M(1, 3);
    const float Coef03 = M(1, 2) * M(2, 3) - M(2, 2) * M(1, 3);
    const float Coef04 = M(2, 1) * M(3, 3) - M(3, 1) * M(2, 3);
    const float Coef06 = M(1, 1) * M(3, 3) - M(3, 1) * M(1, 3);
    const float Coef07 = M(1, 1) * M(2, 3) - M(2, 1) * M(1, 3);
    const float Coef08 = M(2, 1) * M(3, 2) - M(3, 1) * M(2, 2);
    const float Coef10 = M(1, 1) * M(3, 2) - M(3, 1) * M(1, 2);
    const float Coef11 = M(1, 1) * M(2, 2) - M(2, 1) * M(1, 2);
    const float Coef12 = M(2, 0) * M(3, 3) - M(3, 0) * M(2, 3);
    const float Coef14 = M(1, 0) * M(3, 3) - M(3, 0) * M(1, 3);
    const float Coef15 = M(1, 0) * M(2, 3) - M(2, 0) * M(1, 3);
    const float Coef16 = M(2, 0) * M(3, 2) - M(3, 0) * M(2, 2);
    const float Coef18 = M(1, 0) * M(3, 2) - M(3, 0) * M(1, 2);
    const float Coef19 = M(1, 0) * M(2, 2) - M(2, 0) * M(1, 2);
    const float Coef20 = M(2, 0) * M(3, 1) - M(3, 0) * M(2, 1);
    const float Coef22 = M(1, 0) * M(3, 1) - M(3, 0) * M(1, 1);
    const float Coef23 = M(1, 0) * M(2, 1) - M(2, 0) * M(1, 1);
    const vec4 Fac0 = V4(Coef00, Coef00, Coef02, Coef03), Fac1 = V4(Coef04, Coef04, Coef06, Coef07);
    const vec4 Fac2 = V4(Coef08, Coef08, Coef10, Coef11), Fac3 = V4(Coef12, Coef12, Coef14, Coef15);
    const vec4 Fac4 = V4(Coef16, Coef16, Coef18, Coef19), Fac5 = V4(Coef20, Coef20, Coef22, Coef23);
    const vec4 Vec0 = V4(M(1, 0), M(0, 0), M(0, 0), M(0, 0)), Vec1 = V4(M(1, 1), M(0, 1), M(0, 1), M(0, 1));
    const vec4 Vec2 = V4(M(1, 2), M(0, 2), M(0, 2), M(0, 2)), Vec3 = V4(M(1, 3), M(0, 3), M(0, 3), M(0, 3));
#undef M
    const vec4 Inv0 = Vec1 * Fac0 - Vec2 * Fac1 + Vec3 * Fac2;
    const vec4 Inv1 = Vec0 * Fac0 - Vec2 * Fac3 + Vec3 * Fac4;
    const vec4 Inv2 = Vec0 * Fac1 - Vec1 * Fac3 + Vec3 * Fac5;
    const vec4 Inv3 = Vec0 * Fac2 - Vec1 * Fac4 + Vec2 * Fac5;
    const vec4 SignA = V4(+1.0f, -1.0f, +1.0f, -1.0f), SignB = V4(-1.0f, +1.0f, -1.0f, +1.0f);
    mat4 Inverse = { { Inv0 * SignA, Inv1 * SignB, Inv2 * SignA, Inv3 * SignB } };
    const vec4 Row0 = V4(Inverse.c[0].x, Inverse.c[1].x, Inverse.c[2].x, Inverse.c[3].x);
    const vec4 Dot0 = mm.c[0] * Row0;
    const float Dot1 = (Dot0.x + Dot0.y) + (Dot0.z + Dot0.w);
    const float OneOverDeterminant = 1.0f / Dot1;
    for (int c = 0; c < 4; c++)
        Inverse.c[c] = Inverse.c[c] * OneOverDeterminant;
    return Inverse;
}

} // namespace glsl
