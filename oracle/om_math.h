// oracle/om_math.h -- TEST INFRASTRUCTURE ONLY (parity oracle). Never included by the product path.
//
// CPU restatement of the arithmetic the reference's scalar physics path uses (src/core/math.h, src/core/math.cpp).
// Built with  g++ -O2 -ffp-contract=off -fno-fast-math  so that every expression rounds once per operation in the
// order written here; this file DEFINES the bits the CUDA path is compared with.
#pragma once
#include <cmath>
#include <cstdint>
#include <cfloat>
#include <cstring>

namespace orc {

// math.h:13,22
static const float kPi = 3.14159265359f;
static const float kEps = 1e-6f;

struct V3 { float x, y, z; };
struct Q4 { float x, y, z, w; };
// mat3 is column-major in the reference (math.h:4,389-403): c[j] is column j, element (row i, col j) = c[j][i]
struct M3 { float m[9]; float& at(int r, int c) { return m[c * 3 + r]; } float at(int r, int c) const { return m[c * 3 + r]; } };
struct M2 { float m00, m10, m01, m11; }; // column-major storage order like math.h:348-353

static inline V3 v3(float x, float y, float z) { V3 r = { x, y, z }; return r; }
static inline V3 v3(float s) { V3 r = { s, s, s }; return r; }
static inline float comp(const V3& a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }
static inline void setComp(V3& a, int i, float v) { if (i == 0) a.x = v; else if (i == 1) a.y = v; else a.z = v; }

// math.h:532-547
static inline V3 operator+(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline V3 operator-(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline V3 operator*(V3 a, V3 b) { return v3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline V3 operator/(V3 a, V3 b) { return v3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline V3 operator*(V3 a, float b) { return v3(a.x * b, a.y * b, a.z * b); }
static inline V3 operator*(float a, V3 b) { return b * a; }
static inline V3 operator/(V3 a, float b) { return v3(a.x / b, a.y / b, a.z / b); }
static inline V3 operator-(V3 a) { return v3(-a.x, -a.y, -a.z); }
static inline V3& operator+=(V3& a, V3 b) { a = a + b; return a; }
static inline V3& operator-=(V3& a, V3 b) { a = a - b; return a; }
static inline V3& operator*=(V3& a, float b) { a = a * b; return a; }
static inline V3& operator/=(V3& a, float b) { a = a / b; return a; }

// math.h:580,584
static inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline V3 cross(V3 a, V3 b) { return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float sqLen(V3 a) { return dot(a, a); }
static inline float len(V3 a) { return sqrtf(sqLen(a)); }
// math.h:595 : zero below 1e-8 squared length
static inline V3 noz(V3 a) { float sl = sqLen(a); return (sl < 1e-8f) ? v3(0.f) : (a * (1.f / sqrtf(sl))); }
// math.h:599 : multiply by reciprocal of the length
static inline V3 normalize(V3 a) { float l = len(a); return a * (1.f / l); }
static inline V3 vabs(V3 a) { return v3(fabsf(a.x), fabsf(a.y), fabsf(a.z)); }
static inline float fmin2(float a, float b) { return a < b ? a : b; }   // pch.h min/max macros: (a < b) ? a : b
static inline float fmax2(float a, float b) { return a < b ? b : a; }
static inline V3 vmin(V3 a, V3 b) { return v3(fmin2(a.x, b.x), fmin2(a.y, b.y), fmin2(a.z, b.z)); }
static inline V3 vmax(V3 a, V3 b) { return v3(fmax2(a.x, b.x), fmax2(a.y, b.y), fmax2(a.z, b.z)); }
// math.h:30 : clamp = min(u, max(l, v))
static inline float clampf(float v, float l, float u) { float r = fmax2(l, v); r = fmin2(u, r); return r; }
static inline float clamp01(float v) { return clampf(v, 0.f, 1.f); }
// math.h:27,671
static inline float lerpf(float l, float u, float t) { return l + t * (u - l); }
static inline V3 lerp(V3 l, V3 u, float t) { return l + t * (u - l); }

static inline Q4 q4(float x, float y, float z, float w) { Q4 r = { x, y, z, w }; return r; }
static inline Q4 qIdentity() { return q4(0.f, 0.f, 0.f, 1.f); }
static inline V3 qv(Q4 q) { return v3(q.x, q.y, q.z); }
static inline Q4 conjugate(Q4 a) { return q4(-a.x, -a.y, -a.z, a.w); } // math.h:623
// math.h:627-633
static inline Q4 operator*(Q4 a, Q4 b)
{
    V3 av = qv(a), bv = qv(b);
    float w = a.w * b.w - dot(av, bv);
    V3 v = av * b.w + bv * a.w + cross(av, bv);
    return q4(v.x, v.y, v.z, w);
}
// math.h:642-646 : rotate through two quaternion products
static inline V3 operator*(Q4 q, V3 v) { Q4 p = q4(v.x, v.y, v.z, 0.f); return qv(q * p * conjugate(q)); }
static inline bool operator==(Q4 a, Q4 b) { return a.x == b.x && a.y == b.y && a.z == b.z && a.w == b.w; }
// math.h:932-936
static inline Q4 quatAxisAngle(V3 axis, float angle) { float w = cosf(angle * 0.5f); V3 v = axis * sinf(angle * 0.5f); return q4(v.x, v.y, v.z, w); }
// math.h:581 + simd.h:349 : vec4 dot = hadd(hadd(a*b)) = (x+y)+(z+w); math.h:600,622 : scale by 1/length
static inline float dot4(Q4 a, Q4 b) { return (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w); }
static inline Q4 normalize(Q4 a) { float l = sqrtf(dot4(a, a)); float s = 1.f / l; return q4(a.x * s, a.y * s, a.z * s, a.w * s); }

static inline M3 m3Zero() { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = 0.f; return r; }
static inline M3 m3Identity() { M3 r = m3Zero(); r.m[0] = r.m[4] = r.m[8] = 1.f; return r; }
static inline V3 m3Row(const M3& a, int r) { return v3(a.m[r], a.m[r + 3], a.m[r + 6]); }   // math.h:486
static inline V3 m3Col(const M3& a, int c) { return v3(a.m[c * 3], a.m[c * 3 + 1], a.m[c * 3 + 2]); }
static inline V3 operator*(const M3& a, V3 b) { return v3(dot(m3Row(a, 0), b), dot(m3Row(a, 1), b), dot(m3Row(a, 2), b)); } // math.h:660
// math.cpp:103-118
static inline M3 operator*(const M3& a, const M3& b)
{
    M3 r;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.at(i, j) = dot(m3Row(a, i), m3Col(b, j));
    return r;
}
static inline M3 operator+(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] + b.m[i]; return r; }
static inline M3 operator-(const M3& a, const M3& b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] - b.m[i]; return r; }
static inline M3 operator*(const M3& a, float b) { M3 r; for (int i = 0; i < 9; ++i) r.m[i] = a.m[i] * b; return r; }
static inline M3 transpose(const M3& a) { M3 r; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.at(i, j) = a.at(j, i); return r; }
static inline float trace(const M3& a) { return a.at(0, 0) + a.at(1, 1) + a.at(2, 2); }
// math.cpp:443-448
static inline float determinant(const M3& m)
{
    return m.at(0, 0) * (m.at(1, 1) * m.at(2, 2) - m.at(2, 1) * m.at(1, 2))
         - m.at(0, 1) * (m.at(1, 0) * m.at(2, 2) - m.at(2, 0) * m.at(1, 2))
         + m.at(0, 2) * (m.at(1, 0) * m.at(2, 1) - m.at(2, 0) * m.at(1, 1));
}
// math.cpp:276-306 (returns an uninitialised matrix on det == 0 in the reference; zero here)
static inline M3 invert(const M3& m)
{
    M3 inv;
    inv.at(0, 0) = m.at(1, 1) * m.at(2, 2) - m.at(2, 1) * m.at(1, 2);
    inv.at(0, 1) = m.at(0, 2) * m.at(2, 1) - m.at(2, 2) * m.at(0, 1);
    inv.at(0, 2) = m.at(0, 1) * m.at(1, 2) - m.at(1, 1) * m.at(0, 2);
    inv.at(1, 0) = m.at(1, 2) * m.at(2, 0) - m.at(2, 2) * m.at(1, 0);
    inv.at(1, 1) = m.at(0, 0) * m.at(2, 2) - m.at(2, 0) * m.at(0, 2);
    inv.at(1, 2) = m.at(0, 2) * m.at(1, 0) - m.at(1, 2) * m.at(0, 0);
    inv.at(2, 0) = m.at(1, 0) * m.at(2, 1) - m.at(2, 0) * m.at(1, 1);
    inv.at(2, 1) = m.at(0, 1) * m.at(2, 0) - m.at(2, 1) * m.at(0, 0);
    inv.at(2, 2) = m.at(0, 0) * m.at(1, 1) - m.at(1, 0) * m.at(0, 1);
    float det = determinant(m);
    if (det == 0.f) { return m3Zero(); }
    det = 1.f / det;
    return inv * det;
}
// math.cpp:644-676 (identity shortcut when q.w == 1)
static inline M3 quaternionToMat3(Q4 q)
{
    if (q.w == 1.f) { return m3Identity(); }
    float qxx = q.x * q.x, qyy = q.y * q.y, qzz = q.z * q.z;
    float qxz = q.x * q.z, qxy = q.x * q.y, qyz = q.y * q.z;
    float qwx = q.w * q.x, qwy = q.w * q.y, qwz = q.w * q.z;
    M3 r;
    r.at(0, 0) = 1.f - 2.f * (qyy + qzz); r.at(1, 0) = 2.f * (qxy + qwz);       r.at(2, 0) = 2.f * (qxz - qwy);
    r.at(0, 1) = 2.f * (qxy - qwz);       r.at(1, 1) = 1.f - 2.f * (qxx + qzz); r.at(2, 1) = 2.f * (qyz + qwx);
    r.at(0, 2) = 2.f * (qxz + qwy);       r.at(1, 2) = 2.f * (qyz - qwx);       r.at(2, 2) = 1.f - 2.f * (qxx + qyy);
    return r;
}
// math.cpp:778-810
static inline M3 outerProduct(V3 a, V3 b)
{
    V3 c0 = a * b.x, c1 = a * b.y, c2 = a * b.z; M3 r;
    r.at(0, 0) = c0.x; r.at(1, 0) = c0.y; r.at(2, 0) = c0.z;
    r.at(0, 1) = c1.x; r.at(1, 1) = c1.y; r.at(2, 1) = c1.z;
    r.at(0, 2) = c2.x; r.at(1, 2) = c2.y; r.at(2, 2) = c2.z;
    return r;
}
static inline M3 skew(V3 r)
{
    M3 s;
    s.at(0, 0) = 0.f;  s.at(0, 1) = -r.z; s.at(0, 2) = r.y;
    s.at(1, 0) = r.z;  s.at(1, 1) = 0.f;  s.at(1, 2) = -r.x;
    s.at(2, 0) = -r.y; s.at(2, 1) = r.x;  s.at(2, 2) = 0.f;
    return s;
}
// math.cpp:1342-1371 : Cramer; a zero determinant is left as 0 (result 0)
static inline void solve2(const M2& A, float bx, float by, float& ox, float& oy)
{
    float a11 = A.m00, a12 = A.m01, a21 = A.m10, a22 = A.m11;
    float det = a11 * a22 - a12 * a21;
    if (det != 0.f) { det = 1.f / det; }
    ox = det * (a22 * bx - a12 * by);
    oy = det * (a11 * by - a21 * bx);
}
static inline V3 solve3(const M3& A, V3 b)
{
    V3 ex = m3Col(A, 0), ey = m3Col(A, 1), ez = m3Col(A, 2);
    float det = dot(ex, cross(ey, ez));
    if (det != 0.f) { det = 1.f / det; }
    V3 x;
    x.x = det * dot(b, cross(ey, ez));
    x.y = det * dot(ex, cross(b, ez));
    x.z = det * dot(ex, cross(ey, b));
    return x;
}
// math.cpp:1388-1405 (3D variant, |denom| < EPSILON -> 1)
static inline V3 barycentric(V3 a, V3 b, V3 c, V3 p)
{
    V3 v0 = b - a, v1 = c - a, v2 = p - a;
    float d00 = dot(v0, v0), d01 = dot(v0, v1), d11 = dot(v1, v1), d20 = dot(v2, v0), d21 = dot(v2, v1);
    float denom = d00 * d11 - d01 * d01;
    denom = (fabsf(denom) < kEps) ? 1.f : denom;
    float v = (d11 * d20 - d01 * d21) / denom;
    float w = (d00 * d21 - d01 * d20) / denom;
    float u = 1.0f - v - w;
    return v3(u, v, w);
}
// math.cpp:1414-1426
static inline V3 getTangent(V3 n)
{
    V3 t = (fabsf(n.x) >= 0.57735f) ? v3(n.y, -n.x, 0.f) : v3(0.f, n.z, -n.y);
    return normalize(t);
}
static inline void getTangents(V3 n, V3& t, V3& b) { t = getTangent(n); b = cross(n, t); }
// math.cpp:538-574
static inline Q4 rotateFromTo(V3 from_, V3 to_)
{
    V3 from = normalize(from_), to = normalize(to_);
    float d = dot(from, to);
    if (d >= 1.f) { return qIdentity(); }
    Q4 q;
    if (d < (1e-6f - 1.f))
    {
        V3 axis = cross(v3(1.f, 0.f, 0.f), from);
        if (sqLen(axis) == 0.f) { axis = cross(v3(0.f, 1.f, 0.f), from); }
        axis = normalize(axis);
        q = normalize(quatAxisAngle(axis, kPi));
    }
    else
    {
        float s = sqrtf((1.f + d) * 2.f);
        float invs = 1.f / s;
        V3 c = cross(from, to);
        q = normalize(q4(c.x * invs, c.y * invs, c.z * invs, s * 0.5f));
    }
    return q;
}
// math.cpp:576-592
static inline void getAxisRotation(Q4 q, V3& axis, float& angle)
{
    float sq = sqLen(qv(q));
    if (sq > 0.f) { angle = 2.f * acosf(q.w); float inv = 1.f / sqrtf(sq); axis = qv(q) * inv; }
    else { angle = 0.f; axis = v3(1.f, 0.f, 0.f); }
}
// math.cpp:518-536 with unit scale
static inline V3 transformPosition(Q4 r, V3 p, V3 x) { return r * x + p; }
static inline V3 inverseTransformPosition(Q4 r, V3 p, V3 x) { return conjugate(r) * (x - p); }

// core/random.h:5-50 : xorshift64 with shifts 13/7/17
struct Rng
{
    uint64_t state;
    uint64_t u64() { uint64_t x = state; x ^= x << 13; x ^= x >> 7; x ^= x << 17; state = x; return x; }
    uint32_t u32() { return (uint32_t)u64(); }
    float f01() { return u32() / (float)UINT32_MAX; }
    // remap(v,0,1,lo,hi) = lerp(lo,hi,(v-0)/(1-0))  (math.h:27-29)
    float between(float lo, float hi) { float t = (f01() - 0.f) / (1.f - 0.f); return lerpf(lo, hi, t); }
};

} // namespace orc
