// b200_math.cuh -- device-side fp32 vector / quaternion / mat3 arithmetic for the sm_100a rigid-body kernels.
//
// The whole library is compiled with -fmad=false and IEEE div/sqrt (nvcc defaults, no --use_fast_math), so every
// expression rounds once per operation in the order written. The operation ORDER follows the reference's scalar path
// (src/core/math.h / math.cpp, cited below) because the broad-phase parity gate is bit-exact on AABBs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <float.h>
#include <math.h>
#include <string.h>

namespace b200 {

#define B200_HD __host__ __device__ __forceinline__
#define B200_D __device__ __forceinline__

struct f3 { float x, y, z; };
struct q4 { float x, y, z, w; };
struct m3 { float m[9]; };      // column-major like the reference (math.h:389-403): m[c*3+r]
struct m2 { float m00, m10, m01, m11; };

#define B200_PI 3.14159265359f   /* math.h:13 */
#define B200_EPS 1e-6f           /* math.h:22 */

B200_HD f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
B200_HD f3 mk3(float s) { return mk3(s, s, s); }
B200_HD float comp(const f3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
B200_HD void setComp(f3& a, int i, float v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v; }

// math.h:532-547
B200_HD f3 operator+(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
B200_HD f3 operator-(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
B200_HD f3 operator*(f3 a, f3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
B200_HD f3 operator*(f3 a, float b) { return mk3(a.x * b, a.y * b, a.z * b); }
B200_HD f3 operator*(float a, f3 b) { return b * a; }
B200_HD f3 operator/(f3 a, float b) { return mk3(a.x / b, a.y / b, a.z / b); }
B200_HD f3 operator-(f3 a) { return mk3(-a.x, -a.y, -a.z); }
B200_HD f3& operator+=(f3& a, f3 b) { a = a + b; return a; }
B200_HD f3& operator-=(f3& a, f3 b) { a = a - b; return a; }
B200_HD f3& operator*=(f3& a, float b) { a = a * b; return a; }
B200_HD f3& operator/=(f3& a, float b) { a = a / b; return a; }

B200_HD float dot(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }                                   // math.h:580
B200_HD f3 cross(f3 a, f3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); } // math.h:584
B200_HD float sqLen(f3 a) { return dot(a, a); }
B200_HD float len(f3 a) { return sqrtf(sqLen(a)); }
B200_HD f3 noz(f3 a) { float sl = sqLen(a); return (sl < 1e-8f) ? mk3(0.f) : (a * (1.f / sqrtf(sl))); }         // math.h:595
B200_HD f3 normalize(f3 a) { float l = len(a); return a * (1.f / l); }                                          // math.h:599
B200_HD f3 vabs(f3 a) { return mk3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
// pch.h:60-70 : min = (a < b) ? a : b, max = (a < b) ? b : a  (NOT fminf/fmaxf: ties and NaNs behave differently)
B200_HD float fmin2(float a, float b) { return a < b ? a : b; }
B200_HD float fmax2(float a, float b) { return a < b ? b : a; }
B200_HD f3 vmin(f3 a, f3 b) { return mk3(fmin2(a.x, b.x), fmin2(a.y, b.y), fmin2(a.z, b.z)); }
B200_HD f3 vmax(f3 a, f3 b) { return mk3(fmax2(a.x, b.x), fmax2(a.y, b.y), fmax2(a.z, b.z)); }
B200_HD float clampf(float v, float l, float u) { float r = fmax2(l, v); r = fmin2(u, r); return r; }          // math.h:30
B200_HD float clamp01(float v) { return clampf(v, 0.f, 1.f); }
B200_HD float lerpf(float l, float u, float t) { return l + t * (u - l); }                                      // math.h:27
B200_HD f3 lerp(f3 l, f3 u, float t) { return l + t * (u - l); }

B200_HD q4 mkq(float x, float y, float z, float w) { q4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }
B200_HD q4 qIdentity() { return mkq(0.f, 0.f, 0.f, 1.f); }
B200_HD f3 qv(q4 q) { return mk3(q.x, q.y, q.z); }
B200_HD q4 conjugate(q4 a) { return mkq(-a.x, -a.y, -a.z, a.w); }
// math.h:627-633
B200_HD q4 operator*(q4 a, q4 b)
{
    f3 av = qv(a), bv = qv(b);
    float w = a.w * b.w - dot(av, bv);
    f3 v = av * b.w + bv * a.w + cross(av, bv);
    return mkq(v.x, v.y, v.z, w);
}
// math.h:642-646 : q * (v,0) * conj(q)
B200_HD f3 operator*(q4 q, f3 v) { q4 p = mkq(v.x, v.y, v.z, 0.f); return qv(q * p * conjugate(q)); }
B200_HD bool qIsIdentity(q4 a) { return a.x == 0.f && a.y == 0.f && a.z == 0.f && a.w == 1.f; }
B200_HD q4 quatAxisAngle(f3 axis, float angle) { float w = cosf(angle * 0.5f); f3 v = axis * sinf(angle * 0.5f); return mkq(v.x, v.y, v.z, w); }
// vec4 dot = hadd(hadd(a*b)) = (x+y)+(z+w)  (math.h:581, simd.h:349); normalise scales by 1/length (math.h:600,622)
B200_HD float dot4(q4 a, q4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
B200_HD q4 normalize(q4 a) { float l = sqrtf(dot4(a, a)); float s = 1.f / l; return mkq(a.x * s, a.y * s, a.z * s, a.w * s); }

B200_HD float mat(const m3& a, int r, int c) { return a.m[c * 3 + r]; }
B200_HD float& matRef(m3& a, int r, int c) { return a.m[c * 3 + r]; }
B200_HD m3 m3Zero() { m3 r; for (int i = 0; i < 9; ++i) r.m[i] = 0.f; return r; }
B200_HD m3 m3Identity() { m3 r = m3Zero(); r.m[0] = r.m[4] = r.m[8] = 1.f; return r; }
B200_HD f3 m3Row(const m3& a, int r) { return mk3(a.m[r], a.m[r + 3], a.m[r + 6]); }
B200_HD f3 m3Col(const m3& a, int c) { return mk3(a.m[c * 3], a.m[c * 3 + 1], a.m[c * 3 + 2]); }
B200_HD f3 operator*(const m3& a, f3 b) { return mk3(dot(m3Row(a, 0), b), dot(m3Row(a, 1), b), dot(m3Row(a, 2), b)); } // math.h:660
B200_HD m3 operator*(const m3& a, const m3& b)                                                                        // math.cpp:103-118
{
    m3 r;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) matRef(r, i, j) = dot(m3Row(a, i), m3Col(b, j));
    return r;
}
B200_HD m3 operator+(const m3& a, const m3& b) { m3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] + b.m[i]; return r; }
B200_HD m3 operator-(const m3& a, const m3& b) { m3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] - b.m[i]; return r; }
B200_HD m3 operator*(const m3& a, float b) { m3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] * b; return r; }
B200_HD m3 transpose(const m3& a) { m3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) matRef(r, i, j) = mat(a, j, i); return r; }
// math.cpp:644-676 (identity shortcut when q.w == 1)
B200_HD m3 quaternionToMat3(q4 q)
{
    if (q.w == 1.f) return m3Identity();
    float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z;
    float qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
    float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    m3 r;
    matRef(r, 0, 0) = 1.f - 2.f * (qyy + qzz); matRef(r, 1, 0) = 2.f * (qxy + qwz);       matRef(r, 2, 0) = 2.f * (qxz - qwy);
    matRef(r, 0, 1) = 2.f * (qxy - qwz);       matRef(r, 1, 1) = 1.f - 2.f * (qxx + qzz); matRef(r, 2, 1) = 2.f * (qyz + qwx);
    matRef(r, 0, 2) = 2.f * (qxz + qwy);       matRef(r, 1, 2) = 2.f * (qyz - qwx);       matRef(r, 2, 2) = 1.f - 2.f * (qxx + qyy);
    return r;
}
// math.cpp:443-448
B200_HD float determinant(const m3& m)
{
    return mat(m, 0, 0) * (mat(m, 1, 1) * mat(m, 2, 2) - mat(m, 2, 1) * mat(m, 1, 2))
         - mat(m, 0, 1) * (mat(m, 1, 0) * mat(m, 2, 2) - mat(m, 2, 0) * mat(m, 1, 2))
         + mat(m, 0, 2) * (mat(m, 1, 0) * mat(m, 2, 1) - mat(m, 2, 0) * mat(m, 1, 1));
}
// math.cpp:276-306
B200_HD m3 invert(const m3& m)
{
    m3 inv;
    matRef(inv, 0, 0) = mat(m, 1, 1) * mat(m, 2, 2) - mat(m, 2, 1) * mat(m, 1, 2);
    matRef(inv, 0, 1) = mat(m, 0, 2) * mat(m, 2, 1) - mat(m, 2, 2) * mat(m, 0, 1);
    matRef(inv, 0, 2) = mat(m, 0, 1) * mat(m, 1, 2) - mat(m, 1, 1) * mat(m, 0, 2);
    matRef(inv, 1, 0) = mat(m, 1, 2) * mat(m, 2, 0) - mat(m, 2, 2) * mat(m, 1, 0);
    matRef(inv, 1, 1) = mat(m, 0, 0) * mat(m, 2, 2) - mat(m, 2, 0) * mat(m, 0, 2);
    matRef(inv, 1, 2) = mat(m, 0, 2) * mat(m, 1, 0) - mat(m, 1, 2) * mat(m, 0, 0);
    matRef(inv, 2, 0) = mat(m, 1, 0) * mat(m, 2, 1) - mat(m, 2, 0) * mat(m, 1, 1);
    matRef(inv, 2, 1) = mat(m, 0, 1) * mat(m, 2, 0) - mat(m, 2, 1) * mat(m, 0, 0);
    matRef(inv, 2, 2) = mat(m, 0, 0) * mat(m, 1, 1) - mat(m, 1, 0) * mat(m, 0, 1);
    float det = determinant(m);
    if (det == 0.f) return m3Zero();
    det = 1.f / det;
    return inv * det;
}
B200_HD float trace(const m3& a) { return mat(a, 0, 0) + mat(a, 1, 1) + mat(a, 2, 2); }
// math.cpp:778-791
B200_HD m3 outerProduct(f3 a, f3 b)
{
    f3 c0 = a * b.x, c1 = a * b.y, c2 = a * b.z; m3 r;
    matRef(r, 0, 0) = c0.x; matRef(r, 1, 0) = c0.y; matRef(r, 2, 0) = c0.z;
    matRef(r, 0, 1) = c1.x; matRef(r, 1, 1) = c1.y; matRef(r, 2, 1) = c1.z;
    matRef(r, 0, 2) = c2.x; matRef(r, 1, 2) = c2.y; matRef(r, 2, 2) = c2.z;
    return r;
}
// math.cpp:793-810
B200_HD m3 skew(f3 r)
{
    m3 s;
    matRef(s, 0, 0) = 0.f;  matRef(s, 0, 1) = -r.z; matRef(s, 0, 2) = r.y;
    matRef(s, 1, 0) = r.z;  matRef(s, 1, 1) = 0.f;  matRef(s, 1, 2) = -r.x;
    matRef(s, 2, 0) = -r.y; matRef(s, 2, 1) = r.x;  matRef(s, 2, 2) = 0.f;
    return s;
}
// math.cpp:1342-1371 (Cramer; zero determinant stays 0)
B200_HD void solve2(const m2& A, float bx, float by, float& ox, float& oy)
{
    float a11 = A.m00, a12 = A.m01, a21 = A.m10, a22 = A.m11;
    float det = a11 * a22 - a12 * a21;
    if (det != 0.f) det = 1.f / det;
    ox = det * (a22 * bx - a12 * by);
    oy = det * (a11 * by - a21 * bx);
}
B200_HD f3 solve3(const m3& A, f3 b)
{
    f3 ex = m3Col(A, 0), ey = m3Col(A, 1), ez = m3Col(A, 2);
    float det = dot(ex, cross(ey, ez));
    if (det != 0.f) det = 1.f / det;
    f3 x;
    x.x = det * dot(b, cross(ey, ez));
    x.y = det * dot(ex, cross(b, ez));
    x.z = det * dot(ex, cross(ey, b));
    return x;
}
// math.cpp:1388-1405
B200_HD f3 barycentric(f3 a, f3 b, f3 c, f3 p)
{
    f3 v0 = b - a, v1 = c - a, v2 = p - a;
    float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    float denom = d00 * d11 - d01 * d01;
    denom = (fabsf(denom) < B200_EPS) ? 1.f : denom;
    float v = (d11 * d20 - d01 * d21) / denom;
    float w = (d00 * d21 - d01 * d20) / denom;
    float u = 1.0f - v - w;
    return mk3(u, v, w);
}
// math.cpp:1414-1426
B200_HD f3 getTangent(f3 n)
{
    f3 t = (fabsf(n.x) >= 0.57735f) ? mk3(n.y, -n.x, 0.f) : mk3(0.f, n.z, -n.y);
    return normalize(t);
}
B200_HD void getTangents(f3 n, f3& t, f3& b) { t = getTangent(n); b = cross(n, t); }
// math.cpp:538-574
B200_HD q4 rotateFromTo(f3 from_, f3 to_)
{
    f3 from = normalize(from_), to = normalize(to_);
    float d = dot(from, to);
    if (d >= 1.f) return qIdentity();
    q4 q;
    if (d < (1e-6f - 1.f))
    {
        f3 axis = cross(mk3(1.f, 0.f, 0.f), from);
        if (sqLen(axis) == 0.f) axis = cross(mk3(0.f, 1.f, 0.f), from);
        axis = normalize(axis);
        q = normalize(quatAxisAngle(axis, B200_PI));
    }
    else
    {
        float s = sqrtf((1.f + d) * 2.f);
        float invs = 1.f / s;
        f3 c = cross(from, to);
        q = normalize(mkq(c.x * invs, c.y * invs, c.z * invs, s * 0.5f));
    }
    return q;
}
// math.cpp:576-592
B200_HD void getAxisRotation(q4 q, f3& axis, float& angle)
{
    float sq = sqLen(qv(q));
    if (sq > 0.f) { angle = 2.f * acosf(q.w); float inv = 1.f / sqrtf(sq); axis = qv(q) * inv; }
    else { angle = 0.f; axis = mk3(1.f, 0.f, 0.f); }
}

// 128-bit helpers
B200_HD float4 mk4(f3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
B200_HD float4 mk4(q4 q) { return make_float4(q.x, q.y, q.z, q.w); }
B200_HD f3 xyz(float4 v) { return mk3(v.x, v.y, v.z); }
B200_HD q4 toq(float4 v) { return mkq(v.x, v.y, v.z, v.w); }

} // namespace b200
