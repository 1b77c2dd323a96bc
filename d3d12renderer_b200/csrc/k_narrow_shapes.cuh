// k_narrow_shapes.cuh -- per-pair contact generation on the device: the 21 shape-pair routines of the reference's
// scalar narrow phase (src/physics/collision_narrow.cpp:374-1584) with their helpers (:56-369), the GJK boolean test
// (collision_gjk.h:184-238, collision_gjk.cpp:6-212) and EPA (collision_epa.h:96-168, collision_epa.cpp:5-239).
// One thread owns one pair; simplex / polytope state lives in registers and thread-local memory. Arithmetic order follows
// the reference so that contacts agree with the CPU oracle to the last bits (the library is built with -fmad=false).
#pragma once
#include "b200_util.cuh"
#include <type_traits>
#include "../../include/b200phys.h"

namespace b200 {

struct Man { f3 p[4]; float d[4]; f3 n; uint32_t cnt; };

struct SphereS { f3 c; float r; };
struct CapS { f3 a, b; float r; };          // capsule and cylinder
struct BoxS { f3 mn, mx; };
struct ObbS { q4 q; f3 c, r; };
struct HullS { q4 q; f3 p; const float4* verts; uint32_t n; };

struct NarrowCounters { uint32_t epaOverflow, gjkCap, asserts; };
struct SimplexD;
// GJK and EPA run in separate kernels (k_narrow.cu): PHASE 1 = GJK only, a hit parks its simplex in a work list and the routine
// reports "no manifold yet"; PHASE 2 = reload the simplex and continue from EPA. The arithmetic per pair is that of the fused
// routine, only the lanes of the EPA warps are all busy and the GJK kernel carries no polytope.
template <int PHASE> struct NarrowCtx : NarrowCounters
{
    static constexpr int phase = PHASE;
    float4* sxPlanes;      // 9 planes of float4, plane stride sxCap
    uint32_t sxCap;
    uint32_t* sxCounter;   // PHASE 1: work-list size
    uint32_t sxEntry;      // PHASE 1: entry written (valid when deferred); PHASE 2 / 3: entry to read
    bool deferred;
    struct EpaWarpShared* epaShared;   // PHASE 3: this warp's polytope in shared memory (warp-cooperative EPA, every lane runs the routine uniformly)
};

// ------------------------------------------------------------------------------------------------------- helpers
B200_D f3 closestPointSegment(f3 q, f3 a, f3 b)
{   // bounding_volumes.h:365-371
    f3 ab = b - a;
    float t = dot(q - a, ab) / sqLen(ab);
    t = clampf(t, 0.f, 1.f);
    return a + t * ab;
}
B200_D f3 closestPointBox(f3 q, const BoxS& b)
{   // bounding_volumes.h:373-384
    f3 r;
    float v = q.x; if (v < b.mn.x) v = b.mn.x; if (v > b.mx.x) v = b.mx.x; r.x = v;
    v = q.y; if (v < b.mn.y) v = b.mn.y; if (v > b.mx.y) v = b.mx.y; r.y = v;
    v = q.z; if (v < b.mn.z) v = b.mn.z; if (v > b.mx.z) v = b.mx.z; r.z = v;
    return r;
}
B200_D float closestSegmentSegment(f3 p1, f3 q1, f3 p2, f3 q2, f3& c1, f3& c2)
{   // bounding_volumes.cpp:1251-1318
    float s, t;
    f3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
    float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
    if (a <= B200_EPS && e <= B200_EPS) { c1 = p1; c2 = p2; return dot(c1 - c2, c1 - c2); }
    if (a <= B200_EPS) { s = 0.f; t = f / e; t = clampf(t, 0.f, 1.f); }
    else
    {
        float c = dot(d1, r);
        if (e <= B200_EPS) { t = 0.f; s = clampf(-c / a, 0.f, 1.f); }
        else
        {
            float b = dot(d1, d2);
            float denom = a * e - b * b;
            if (denom != 0.f) s = clampf((b * f - c * e) / denom, 0.f, 1.f); else s = 0.f;
            t = (b * s + f) / e;
            if (t < 0.f) { t = 0.f; s = clampf(-c / a, 0.f, 1.f); }
            else if (t > 1.f) { t = 1.f; s = clampf((b - c) / a, 0.f, 1.f); }
        }
    }
    c1 = p1 + d1 * s; c2 = p2 + d2 * t;
    return sqLen(c1 - c2);
}
struct PlaneS { f3 n; float d; };
B200_D PlaneS createPlane(f3 point, f3 normal) { PlaneS p; p.n = normal; p.d = -dot(normal, point); return p; }
// dot(vec4(p,1), plane) with the reference's horizontal-add order (x+y)+(z+w)
B200_D float signedDistance(f3 p, const PlaneS& pl) { return (p.x * pl.n.x + p.y * pl.n.y) + (p.z * pl.n.z + 1.f * pl.d); }
B200_D f3 boxCenter(const BoxS& b) { return (b.mn + b.mx) * 0.5f; }
B200_D f3 boxRadius(const BoxS& b) { return (b.mx - b.mn) * 0.5f; }

// ------------------------------------------------------------------------------------------------ analytic pairs
// collision_narrow.cpp:374-400
B200_D bool sphereSphere(const SphereS& s1, const SphereS& s2, Man& m, NarrowCounters& nc)
{
    f3 n = s2.c - s1.c;
    float rs = s2.r + s1.r;
    float sq = sqLen(n);
    if (sq <= rs * rs)
    {
        float dist;
        if (sq == 0.f) { dist = 0.f; m.n = mk3(0.f, 1.f, 0.f); }
        else { dist = sqrtf(sq); m.n = n / dist; }
        m.cnt = 1;
        m.d[0] = rs - dist;
        if (!(m.d[0] >= 0.f)) ++nc.asserts;
        m.p[0] = 0.5f * (s1.c + s1.r * m.n + s2.c - s2.r * m.n);
        return true;
    }
    return false;
}
// collision_narrow.cpp:402-406
B200_D bool sphereCapsule(const SphereS& s, const CapS& c, Man& m, NarrowCounters& nc)
{
    SphereS o; o.c = closestPointSegment(s.c, c.a, c.b); o.r = c.r;
    return sphereSphere(s, o, m, nc);
}
// collision_narrow.cpp:408-449
B200_D bool sphereCylinder(const SphereS& s, const CapS& c, Man& m, NarrowCounters& nc)
{
    f3 ab = c.b - c.a;
    float t = dot(s.c - c.a, ab) / sqLen(ab);
    if (t >= 0.f && t <= 1.f) { SphereS o; o.c = lerp(c.a, c.b, t); o.r = c.r; return sphereSphere(s, o, m, nc); }
    f3 p = (t <= 0.f) ? c.a : c.b;
    f3 up = (t <= 0.f) ? -ab : ab;
    f3 proj = normalize(cross(cross(up, s.c - p), up));
    f3 endA = p + proj * c.r, endB = p - proj * c.r;
    f3 closest = closestPointSegment(s.c, endA, endB);
    f3 normal = closest - s.c;
    float sq = sqLen(normal);
    if (sq <= s.r * s.r)
    {
        float dist;
        if (sq == 0.f) { dist = 0.f; m.n = -normalize(up); }
        else { dist = sqrtf(sq); m.n = normal / dist; }
        m.cnt = 1;
        m.d[0] = s.r - dist;
        if (!(m.d[0] >= 0.f)) ++nc.asserts;
        m.p[0] = closest + 0.5f * m.d[0] * normal;
        return true;
    }
    return false;
}
// collision_narrow.cpp:451-478
B200_D bool sphereBox(const SphereS& s, const BoxS& a, Man& m)
{
    f3 p = closestPointBox(s.c, a);
    f3 n = p - s.c;
    float sq = sqLen(n);
    if (sq <= s.r * s.r)
    {
        float dist = 0.f;
        if (sq > 0.f) { dist = sqrtf(sq); n /= dist; } else { n = mk3(0.f, 1.f, 0.f); }
        m.cnt = 1; m.n = n;
        m.d[0] = s.r - dist;
        m.p[0] = 0.5f * (p + s.c + n * s.r);
        return true;
    }
    return false;
}
// collision_narrow.cpp:480-494
B200_D bool sphereObb(const SphereS& s, const ObbS& o, Man& m)
{
    BoxS bb; bb.mn = o.c - o.r; bb.mx = o.c + o.r;
    SphereS s_; s_.c = conjugate(o.q) * (s.c - o.c) + o.c; s_.r = s.r;
    if (sphereBox(s_, bb, m))
    {
        m.n = o.q * m.n;
        m.p[0] = o.q * (m.p[0] - o.c) + o.c;
        return true;
    }
    return false;
}
// capsule-capsule (:523-612) and capsule-cylinder (:614-703)
B200_D bool capsuleVsSegmentShape(const CapS& a, const CapS& b, bool bIsCylinder, Man& m, NarrowCounters& nc)
{
    f3 aDir = a.b - a.a;
    f3 bDir = normalize(b.b - b.a);
    float aLen = len(aDir);
    aDir *= 1.f / aLen;
    float parallel = dot(aDir, bDir);
    if (fabsf(parallel) > 0.99f)
    {
        f3 pAa = a.a, pAb = a.b, pBa = b.a, pBb = b.b;
        if (parallel < 0.f) { f3 t = pBa; pBa = pBb; pBb = t; }
        f3 ref = a.a;
        float a0 = 0.f, a1 = aLen;
        float b0 = dot(aDir, pBa - ref), b1 = dot(aDir, pBb - ref);
        if (!(b1 > b0)) ++nc.asserts;
        float left = fmax2(a0, b0), right = fmin2(a1, b1);
        if (right < left)
        {
            SphereS sa; sa.r = a.r; sa.c = (a0 > b1) ? pAa : pAb;
            if (bIsCylinder) return sphereCylinder(sa, b, m, nc);
            SphereS sb; sb.r = b.r; sb.c = (a0 > b1) ? pBb : pBa;
            return sphereSphere(sa, sb, m, nc);
        }
        f3 cA0 = ref + left * aDir, cA1 = ref + right * aDir;
        f3 cB0 = closestPointSegment(cA0, pBa, pBb);
        f3 cB1 = cB0 + (right - left) * aDir;
        f3 normal = cB0 - cA0;
        float d = len(normal);
        if (d < B200_EPS) { d = 0.f; normal = mk3(0.f, 1.f, 0.f); } else { normal /= d; }
        float rs = a.r + b.r;
        float pen = rs - d;
        if (pen < 0.f) return false;
        m.n = normal; m.cnt = 2;
        m.d[0] = pen; m.p[0] = (cA0 + cB0) * 0.5f;
        m.d[1] = pen; m.p[1] = (cA1 + cB1) * 0.5f;
        return true;
    }
    f3 c1, c2;
    closestSegmentSegment(a.a, a.b, b.a, b.b, c1, c2);
    SphereS sa; sa.c = c1; sa.r = a.r;
    if (bIsCylinder) return sphereCylinder(sa, b, m, nc);
    SphereS sb; sb.c = c2; sb.r = b.r;
    return sphereSphere(sa, sb, m, nc);
}
// collision_narrow.cpp:1074-1140 (depth keeps the sign of the separation direction)
B200_D bool boxBox(const BoxS& a, const BoxS& b, Man& m)
{
    f3 cA = boxCenter(a), cB = boxCenter(b), rA = boxRadius(a), rB = boxRadius(b);
    f3 d = cB - cA;
    f3 p = (rB + rA) - vabs(d);
    if (p.x < 0.f || p.y < 0.f || p.z < 0.f) return false;
    int me = (p.x < p.y) ? ((p.x < p.z) ? 0 : 2) : ((p.y < p.z) ? 1 : 2);
    float s = comp(d, me) < 0.f ? -1.f : 1.f;
    float pen = comp(p, me) * s;
    f3 normal = mk3(0.f); setComp(normal, me, s);
    m.n = normal; m.cnt = 4;
    int a0 = (me + 1) % 3, a1 = (me + 2) % 3;
    float min0 = fmax2(comp(a.mn, a0), comp(b.mn, a0)), min1 = fmax2(comp(a.mn, a1), comp(b.mn, a1));
    float max0 = fmin2(comp(a.mx, a0), comp(b.mx, a0)), max1 = fmin2(comp(a.mx, a1), comp(b.mx, a1));
    float depth = comp(cA, me) + comp(rA, me) - pen * 0.5f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        m.d[i] = pen; m.p[i] = mk3(0.f);
        setComp(m.p[i], a0, (i < 2) ? min0 : max0); setComp(m.p[i], a1, (i & 1) ? max1 : min1); setComp(m.p[i], me, depth);
    }
    return true;
}

// --------------------------------------------------------------------------------- manifold reduction and clipping
struct VPd { f3 v; float depth; };
struct PolyD { VPd p[16]; uint32_t n; };

// collision_narrow.cpp:56-146
B200_D void stableManifold(const VPd* v, uint32_t n, f3 normal, Man& out)
{
    if (n > 4)
    {
        f3 search = getTangent(normal);
        float best = dot(search, v[0].v); uint32_t ri = 0;
        for (uint32_t i = 1; i < n; ++i) { float d = dot(search, v[i].v); if (d > best) { ri = i; best = d; } }
        out.d[0] = v[ri].depth; out.p[0] = v[ri].v;
        best = 0.f; ri = 0;
        for (uint32_t i = 0; i < n; ++i) { float d = sqLen(v[i].v - out.p[0]); if (d > best) { ri = i; best = d; } }
        out.d[1] = v[ri].depth; out.p[1] = v[ri].v;
        best = 0.f; ri = 0;
        for (uint32_t i = 0; i < n; ++i)
        {
            f3 qa = out.p[0] - v[i].v, qb = out.p[1] - v[i].v;
            float area = 0.5f * dot(cross(qa, qb), normal);
            if (area > best) { ri = i; best = area; }
        }
        out.d[2] = v[ri].depth; out.p[2] = v[ri].v;
        best = 0.f; ri = 0;
        for (uint32_t i = 0; i < n; ++i)
        {
            f3 qa = out.p[0] - v[i].v, qb = out.p[1] - v[i].v, qc = out.p[2] - v[i].v;
            float a1 = 0.5f * dot(cross(qa, qb), normal);
            float a2 = 0.5f * dot(cross(qb, qc), normal);
            float a3 = 0.5f * dot(cross(qc, qa), normal);
            float area = fmax2(fmax2(a1, a2), a3);
            if (area > best) { ri = i; best = area; }
        }
        out.d[3] = v[ri].depth; out.p[3] = v[ri].v;
        out.cnt = 4;
    }
    else
    {
        out.cnt = n;
        for (uint32_t i = 0; i < n; ++i) { out.d[i] = v[i].depth; out.p[i] = v[i].v; }
    }
}
// collision_narrow.cpp:154-163
B200_D VPd clipEdge(VPd a, VPd b, float aDist, float bDist)
{
    aDist = fabsf(aDist); bDist = fabsf(bDist);
    float total = aDist + bDist;
    float t = aDist / total;
    VPd r; r.v = lerp(a.v, b.v, t); r.depth = lerpf(a.depth, b.depth, t);
    return r;
}
// collision_narrow.cpp:166-222
B200_D void clipPolygon(PolyD& input, const PlaneS* planes, uint32_t nPlanes, PolyD& output)
{
    PolyD* in = &input; PolyD* out = &output;
    uint32_t ci = 0;
    for (; ci < nPlanes; ++ci)
    {
        PlaneS pl = planes[ci];
        out->n = 0;
        if (in->n == 0) break;
        VPd start = in->p[in->n - 1];
        for (uint32_t i = 0; i < in->n; ++i)
        {
            VPd end = in->p[i];
            float sd = signedDistance(start.v, pl), ed = signedDistance(end.v, pl);
            bool sIn = sd > 0.f, eIn = ed > 0.f;
            if (sIn && eIn) out->p[out->n++] = end;
            else if (sIn) out->p[out->n++] = clipEdge(start, end, sd, ed);
            else if (!sIn && eIn) { out->p[out->n++] = clipEdge(start, end, sd, ed); out->p[out->n++] = end; }
            start = end;
        }
        PolyD* tmp = in; in = out; out = tmp;
    }
    if (ci % 2 == 0) { for (uint32_t i = 0; i < input.n; ++i) output.p[i] = input.p[i]; output.n = input.n; }
}
B200_D int maxAbsAxis(f3 n) { f3 p = vabs(n); return (p.x > p.y) ? ((p.x > p.z) ? 0 : 2) : ((p.y > p.z) ? 1 : 2); }
// collision_narrow.cpp:225-254
B200_D void boxClipPlanes(f3 radius, f3 normal, f3* pts, f3* nrm)
{
    int me = maxAbsAxis(normal), a0 = (me + 1) % 3, a1 = (me + 2) % 3;
    f3 n;
    n = mk3(0.f); setComp(n, a0, 1.f);  nrm[0] = n; pts[0] = -radius;
    n = mk3(0.f); setComp(n, a1, 1.f);  nrm[1] = n; pts[1] = -radius;
    n = mk3(0.f); setComp(n, a0, -1.f); nrm[2] = n; pts[2] = radius;
    n = mk3(0.f); setComp(n, a1, -1.f); nrm[3] = n; pts[3] = radius;
}
// collision_narrow.cpp:257-289
B200_D void boxIncidentVerts(f3 radius, f3 normal, PolyD& poly)
{
    int me = maxAbsAxis(normal), a0 = (me + 1) % 3, a1 = (me + 2) % 3;
    float s = comp(normal, me) < 0.f ? 1.f : -1.f;
    float d = comp(radius, me) * s;
    float min0 = -comp(radius, a0), min1 = -comp(radius, a1), max0 = comp(radius, a0), max1 = comp(radius, a1);
    poly.n = 4;
    for (int i = 0; i < 4; ++i)
    {
        setComp(poly.p[i].v, me, d);
        setComp(poly.p[i].v, a0, (i == 0 || i == 3) ? min0 : max0);
        setComp(poly.p[i].v, a1, (i < 2) ? min1 : max1);
    }
}
// collision_narrow.cpp:291-299
B200_D PlaneS boxReferencePlane(const BoxS& b, f3 normal)
{
    f3 pt = mk3((normal.x < 0.f) ? b.mn.x : b.mx.x, (normal.y < 0.f) ? b.mn.y : b.mx.y, (normal.z < 0.f) ? b.mn.z : b.mx.z);
    return createPlane(pt, normal);
}
// collision_narrow.cpp:301-336
B200_D void boxIncidentEdge(f3 r, f3 normal, f3& oa, f3& ob)
{
    f3 p = vabs(normal);
    oa = r;
    if (p.x > p.y) { if (p.y > p.z) ob = mk3(r.x, r.y, -r.z); else ob = mk3(r.x, -r.y, r.z); }
    else { if (p.x > p.z) ob = mk3(r.x, r.y, -r.z); else ob = mk3(-r.x, r.y, r.z); }
    f3 s = mk3(normal.x < 0.f ? -1.f : 1.f, normal.y < 0.f ? -1.f : 1.f, normal.z < 0.f ? -1.f : 1.f);
    oa = oa * s; ob = ob * s;
}
// collision_narrow.cpp:339-369
B200_D bool clipAndBuild(PolyD& poly, const PlaneS* planes, uint32_t nPlanes, const PlaneS& ref, Man& out)
{
    PolyD cl; cl.n = 0;
    clipPolygon(poly, planes, nPlanes, cl);
    if (cl.n > 0)
    {
        for (uint32_t i = 0; i < cl.n; ++i)
        {
            if (cl.p[i].depth < 0.f) { cl.p[i] = cl.p[cl.n - 1]; --cl.n; --i; }
            else { cl.p[i].v += ref.n * cl.p[i].depth; }
        }
        if (cl.n > 0) { stableManifold(cl.p, cl.n, out.n, out); return true; }
    }
    return false;
}

// ----------------------------------------------------------------------------------------------- support mappings
struct SupportD
{
    int type;
    f3 a, b; float r;     // sphere: a = centre; capsule/cylinder: a, b; aabb: a = min, b = max
    q4 q; f3 c, rad;      // obb: q, c, rad; hull: q, c = position
    const float4* verts; uint32_t nv;

    B200_D f3 eval(f3 dir) const
    {
        switch (type)
        {
        case B200PHYS_SPHERE: return normalize(dir) * r + a;                    // collision_gjk.h:6-14
        case B200PHYS_CAPSULE:                                                  // collision_gjk.h:16-27
        {
            float dA = dot(dir, a), dB = dot(dir, b);
            f3 far = dA > dB ? a : b;
            return normalize(dir) * r + far;
        }
        case B200PHYS_CYLINDER:                                                 // collision_gjk.h:29-44
        {
            float dA = dot(dir, a), dB = dot(dir, b);
            f3 far = dA > dB ? a : b;
            f3 n = a - b;
            f3 proj = noz(cross(cross(n, dir), n));
            return far + proj * r;
        }
        case B200PHYS_AABB:                                                     // collision_gjk.h:46-58
            return mk3((dir.x < 0.f) ? a.x : b.x, (dir.y < 0.f) ? a.y : b.y, (dir.z < 0.f) ? a.z : b.z);
        case B200PHYS_OBB:                                                      // collision_gjk.h:60-75
        {
            f3 d = conjugate(q) * dir;
            f3 rr = mk3(d.x < 0.f ? -rad.x : rad.x, d.y < 0.f ? -rad.y : rad.y, d.z < 0.f ? -rad.z : rad.z);
            return c + q * rr;
        }
        default:                                                                // hull, collision_gjk.h:77-100
        {
            f3 d = conjugate(q) * dir;
            f3 best = mk3(0.f);
            float maxDist = -FLT_MAX;
            for (uint32_t i = 0; i < nv; ++i)
            {
                float4 v4 = verts[i];
                f3 v = mk3(v4.x, v4.y, v4.z);
                float t = dot(d, v);
                if (t > maxDist) { maxDist = t; best = v; }
            }
            return c + q * best;
        }
        }
    }
    // the same mapping evaluated by a whole warp on uniform arguments: hull vertices are split over the lanes (first maximum wins, like the
    // sequential scan), the closed-form shapes are evaluated redundantly. Every lane returns the same point.
    B200_D f3 evalWarp(f3 dir) const
    {
        if (type != B200PHYS_HULL) return eval(dir);
        const uint32_t lane = threadIdx.x & 31u;
        f3 d = conjugate(q) * dir;
        float maxDist = -FLT_MAX; uint32_t bestI = 0xFFFFFFFFu;
        for (uint32_t i = lane; i < nv; i += 32u)
        {
            const float4 v4 = verts[i];
            const float t = dot(d, mk3(v4.x, v4.y, v4.z));
            if (t > maxDist) { maxDist = t; bestI = i; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            const float ot = __shfl_xor_sync(0xffffffffu, maxDist, o); const uint32_t oi = __shfl_xor_sync(0xffffffffu, bestI, o);
            if (oi != 0xFFFFFFFFu && (bestI == 0xFFFFFFFFu || ot > maxDist || (ot == maxDist && oi < bestI))) { maxDist = ot; bestI = oi; }
        }
        f3 best = mk3(0.f);
        if (bestI != 0xFFFFFFFFu) { const float4 v4 = verts[bestI]; best = mk3(v4.x, v4.y, v4.z); }
        return c + q * best;
    }
};
B200_D SupportD supSphere(const SphereS& s) { SupportD r; r.type = B200PHYS_SPHERE; r.a = s.c; r.r = s.r; return r; }
B200_D SupportD supCapsule(const CapS& s) { SupportD r; r.type = B200PHYS_CAPSULE; r.a = s.a; r.b = s.b; r.r = s.r; return r; }
B200_D SupportD supCylinder(const CapS& s) { SupportD r; r.type = B200PHYS_CYLINDER; r.a = s.a; r.b = s.b; r.r = s.r; return r; }
B200_D SupportD supBox(const BoxS& s) { SupportD r; r.type = B200PHYS_AABB; r.a = s.mn; r.b = s.mx; return r; }
B200_D SupportD supObb(const ObbS& s) { SupportD r; r.type = B200PHYS_OBB; r.q = s.q; r.c = s.c; r.rad = s.r; return r; }
B200_D SupportD supHull(const HullS& s) { SupportD r; r.type = B200PHYS_HULL; r.q = s.q; r.c = s.p; r.verts = s.verts; r.nv = s.n; return r; }

// ------------------------------------------------------------------------------------------------------------ GJK
struct SupPt { f3 a, b, m; };
struct SimplexD { SupPt a, b, c, d; uint32_t n; };
B200_D SupPt minkowski(const SupportD& A, const SupportD& B, f3 dir) { SupPt p; p.a = A.eval(dir); p.b = B.eval(-dir); p.m = p.a - p.b; return p; }
B200_D SupPt minkowskiWarp(const SupportD& A, const SupportD& B, f3 dir) { SupPt p; p.a = A.evalWarp(dir); p.b = B.evalWarp(-dir); p.m = p.a - p.b; return p; }
B200_D f3 crossABA(f3 a, f3 b) { return cross(cross(a, b), a); }

#define B200_GJK_STOP 0
#define B200_GJK_CONTINUE 1
#define B200_GJK_ERROR 2
#define B200_GJK_MAX_ROUNDS 64

// collision_gjk.cpp:6-212 (the reference's goto web between the three face regions, written as a small state machine)
B200_D int updateSimplex(SimplexD& s, const SupPt& a, f3& dir)
{
    if (s.n == 2)
    {
        f3 ao = -a.m, ab = s.b.m - a.m, ac = s.c.m - a.m;
        f3 abc = cross(ab, ac);
        f3 abp = cross(ab, abc);
        if (dot(ao, abp) > 0.f) { s.c = a; dir = crossABA(ab, ao); return B200_GJK_CONTINUE; }
        f3 acp = cross(abc, ac);
        if (dot(ao, acp) > 0.f) { s.b = a; dir = crossABA(ac, ao); return B200_GJK_CONTINUE; }
        if (dot(ao, abc) >= 0.f) { s.d = s.b; s.b = a; s.n = 3; dir = abc; return B200_GJK_CONTINUE; }
        if (dot(ao, -abc) >= 0.f) { s.d = s.c; s.c = s.b; s.b = a; s.n = 3; dir = -abc; return B200_GJK_CONTINUE; }
        return B200_GJK_ERROR;
    }
    if (s.n != 3) return B200_GJK_ERROR;

    f3 ao = -a.m, ab = s.b.m - a.m, ac = s.c.m - a.m, ad = s.d.m - a.m;
    f3 bcd = cross(s.c.m - s.b.m, s.d.m - s.b.m);
    if (dot(bcd, dir) > 0.00001f || dot(bcd, s.b.m) < -0.00001f) return B200_GJK_ERROR;
    f3 abc = cross(ac, ab), abd = cross(ab, ad), adc = cross(ad, ac);
    int flags = 0;
    flags |= (dot(abc, ao) > 0.f) ? 1 : 0;
    flags |= (dot(abd, ao) > 0.f) ? 2 : 0;
    flags |= (dot(adc, ao) > 0.f) ? 4 : 0;
    if (flags == 7) return B200_GJK_ERROR;
    if (flags == 0) return B200_GJK_STOP;

    // states: 1 = overABC1, 2 = overABC2, 3 = overABD1, 4 = overABD2, 5 = overADC1, 6 = overADC2
    int state;
    if (flags == 1) state = 1;
    else if (flags == 2) state = 3;
    else if (flags == 4) state = 5;
    else if (flags == 3) state = (dot(cross(abc, ab), ao) > 0.f) ? 3 : 2;
    else if (flags == 6) state = (dot(cross(abd, ad), ao) > 0.f) ? 5 : 4;
    else state = (dot(cross(adc, ac), ao) > 0.f) ? 1 : 6;   // flags == 5

    for (;;)
    {
        switch (state)
        {
        case 1: if (dot(cross(abc, ab), ao) > 0.f) { s.c = a; s.n = 2; dir = crossABA(ab, ao); return B200_GJK_CONTINUE; } state = 2; break;
        case 2: if (dot(cross(ac, abc), ao) > 0.f) { s.b = a; s.n = 2; dir = crossABA(ac, ao); return B200_GJK_CONTINUE; }
                s.d = a; dir = abc; return B200_GJK_CONTINUE;
        case 3: if (dot(cross(abd, ad), ao) > 0.f) { s.b = s.d; s.c = a; s.n = 2; dir = crossABA(ad, ao); return B200_GJK_CONTINUE; } state = 4; break;
        case 4: if (dot(cross(ab, abd), ao) > 0.f) { s.c = a; s.n = 2; dir = crossABA(ab, ao); return B200_GJK_CONTINUE; }
                s.c = a; dir = abd; return B200_GJK_CONTINUE;
        case 5: if (dot(cross(adc, ac), ao) > 0.f) { s.b = a; s.n = 2; dir = crossABA(ac, ao); return B200_GJK_CONTINUE; } state = 6; break;
        default: if (dot(cross(ad, adc), ao) > 0.f) { s.b = a; s.c = s.d; s.n = 2; dir = crossABA(ad, ao); return B200_GJK_CONTINUE; }
                s.b = a; dir = adc; return B200_GJK_CONTINUE;
        }
    }
}

// collision_gjk.h:184-238 (+ a 64-round cap the reference lacks; hits are counted)
B200_D bool gjk(const SupportD& A, const SupportD& B, SimplexD& s, NarrowCounters& nc)
{
    f3 dir = mk3(1.f, 0.1f, -0.2f);
    s.c = minkowski(A, B, dir);
    if (dot(s.c.m, dir) < 0.f) return false;
    dir = -s.c.m;
    s.b = minkowski(A, B, dir);
    if (dot(s.b.m, dir) < 0.f) return false;
    dir = crossABA(s.c.m - s.b.m, -s.b.m);
    s.n = 2;
    for (int round = 0; ; ++round)
    {
        if (round >= B200_GJK_MAX_ROUNDS) { ++nc.gjkCap; return false; }
        if (sqLen(dir) < 0.0001f) return false;
        SupPt a = minkowski(A, B, dir);
        if (dot(a.m, dir) < 0.f) return false;
        int r = updateSimplex(s, a, dir);
        if (r == B200_GJK_STOP) { s.a = a; s.n = 4; break; }
        if (r == B200_GJK_ERROR) return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------------------ EPA
// Append-only polytope like the reference (inactive triangles are never reclaimed, collision_epa.cpp:21-33) with
// device-sized capacities: <= 20 iterations add <= 20 points; triangle / edge arrays that run full behave like the
// reference's "out of memory" exit (best result so far) and are counted in stats.epaOverflows.
#define B200_EPA_MAX_POINTS 24
#define B200_EPA_MAX_TRIS 160
#define B200_EPA_MAX_EDGES 240
#define B200_EPA_MAX_BORDER 64

struct EpaTri { uint16_t a, b, c, eA, eB, eC; float nx, ny, nz, dist; };
struct EpaEdge { uint16_t a, b, t0, t1; };

struct EpaPoly
{
    SupPt pts[B200_EPA_MAX_POINTS];
    EpaTri tris[B200_EPA_MAX_TRIS];
    EpaEdge edges[B200_EPA_MAX_EDGES];
    uint32_t active[(B200_EPA_MAX_TRIS + 31) / 32];
    uint8_t refs[B200_EPA_MAX_EDGES];
    uint16_t newEdgePerPoint[B200_EPA_MAX_POINTS];
    uint32_t nT, nP, nE;

    // an invalid index (0xFFFF left behind by a degenerate horizon; the reference reads out of bounds there) counts as inactive
    B200_D bool isActive(uint32_t i) const { return i < nT && ((active[i >> 5] >> (i & 31)) & 1u); }
    B200_D void setActive(uint32_t i, bool on) { if (on) active[i >> 5] |= (1u << (i & 31)); else active[i >> 5] &= ~(1u << (i & 31)); }
    B200_D static void triInfo(const SupPt& a, const SupPt& b, const SupPt& c, f3& n, float& d)
    {   // collision_epa.cpp:5-11
        n = normalize(cross(b.m - a.m, c.m - a.m));
        d = dot(n, a.m);
    }
    B200_D int pushPoint(const SupPt& p) { if (nP >= B200_EPA_MAX_POINTS) return -1; pts[nP] = p; return (int)nP++; }
    B200_D int pushTri(uint16_t a, uint16_t b, uint16_t c, uint16_t eA, uint16_t eB, uint16_t eC, f3 n, float d)
    {
        if (nT >= B200_EPA_MAX_TRIS) return -1;
        EpaTri& t = tris[nT]; t.a = a; t.b = b; t.c = c; t.eA = eA; t.eB = eB; t.eC = eC; t.nx = n.x; t.ny = n.y; t.nz = n.z; t.dist = d;
        setActive(nT, true); return (int)nT++;
    }
    B200_D int pushEdge(uint16_t a, uint16_t b, uint16_t t0, uint16_t t1)
    {
        if (nE >= B200_EPA_MAX_EDGES) return -1;
        EpaEdge& e = edges[nE]; e.a = a; e.b = b; e.t0 = t0; e.t1 = t1; return (int)nE++;
    }
    B200_D uint32_t closest() const
    {   // collision_epa.cpp:94-112
        uint32_t best = 0xFFFFFFFFu; float mind = FLT_MAX;
        for (uint32_t i = 0; i < nT; ++i) if (isActive(i) && tris[i].dist < mind) { mind = tris[i].dist; best = i; }
        return best;
    }
    // collision_epa.cpp:114-239
    B200_D bool addPoint(const SupPt& np)
    {
        for (uint32_t i = 0; i < nE; ++i) refs[i] = 0;
        for (uint32_t i = 0; i < nT; ++i)
        {
            if (!isActive(i)) continue;
            const EpaTri& t = tris[i];
            float d = dot(mk3(t.nx, t.ny, t.nz), np.m - pts[t.a].m);
            if (d > 0.f) { if (t.eA < nE) ++refs[t.eA]; if (t.eB < nE) ++refs[t.eB]; if (t.eC < nE) ++refs[t.eC]; setActive(i, false); }
        }
        uint16_t border[B200_EPA_MAX_BORDER]; uint32_t nBorder = 0;
        for (uint32_t i = 0; i < nE; ++i)
            if (refs[i] == 1) { if (nBorder >= B200_EPA_MAX_BORDER) return false; border[nBorder++] = (uint16_t)i; }
        for (uint32_t i = 0; i <= nP && i < B200_EPA_MAX_POINTS; ++i) newEdgePerPoint[i] = 0xFFFF;
        int npi = pushPoint(np);
        if (npi < 0) return false;
        const uint16_t triOffset = (uint16_t)nT;
        for (uint32_t i = 0; i < nBorder; ++i)
        {
            const uint16_t ei = border[i];
            EpaEdge e = edges[ei];
            const bool aAct = isActive(e.t0), bAct = isActive(e.t1);
            const uint16_t connect = bAct ? e.a : e.b;
            const uint16_t triIndex = (uint16_t)nT;
            int ne = pushEdge(connect, (uint16_t)npi, 0xFFFF, (uint16_t)nT);
            if (ne < 0) return false;
            newEdgePerPoint[connect] = (uint16_t)ne;
            const uint16_t bI = connect, cI = bAct ? e.b : e.a;
            f3 n; float d; triInfo(np, pts[bI], pts[cI], n, d);
            int ti = pushTri((uint16_t)npi, bI, cI, ei, 0xFFFF, (uint16_t)ne, n, d);
            if (ti < 0) return false;
            if (aAct) edges[ei].t1 = triIndex; else edges[ei].t0 = triIndex;
        }
        for (uint32_t i = 0; i < nBorder; ++i)
        {
            const EpaEdge e = edges[border[i]];
            const bool bNew = e.t1 >= triOffset;
            const uint16_t connect = bNew ? e.a : e.b;
            const uint16_t other = newEdgePerPoint[connect];
            const uint16_t triIndex = (uint16_t)(i + triOffset);
            if (other == 0xFFFF) continue;
            tris[triIndex].eB = other;
            edges[other].t0 = triIndex;
        }
        return true;
    }
};

// collision_epa.h:96-168 ; result used even when not converged (collision_narrow.cpp:508-512)
B200_D void epa(const SimplexD& g, const SupportD& A, const SupportD& B, f3& outPoint, f3& outNormal, float& outDepth, NarrowCounters& nc)
{
    EpaPoly P;
    P.nT = P.nP = P.nE = 0;
    for (uint32_t i = 0; i < (B200_EPA_MAX_TRIS + 31) / 32; ++i) P.active[i] = 0;
    P.pushPoint(g.a); P.pushPoint(g.b); P.pushPoint(g.c); P.pushPoint(g.d);
    f3 n; float d;
    EpaPoly::triInfo(g.a, g.b, g.d, n, d); P.pushTri(0, 1, 3, 4, 3, 0, n, d);
    EpaPoly::triInfo(g.b, g.c, g.d, n, d); P.pushTri(1, 2, 3, 5, 4, 1, n, d);
    EpaPoly::triInfo(g.c, g.a, g.d, n, d); P.pushTri(2, 0, 3, 3, 5, 2, n, d);
    EpaPoly::triInfo(g.a, g.c, g.b, n, d); P.pushTri(0, 2, 1, 1, 0, 2, n, d);
    P.pushEdge(0, 1, 0, 3); P.pushEdge(1, 2, 1, 3); P.pushEdge(2, 0, 2, 3);
    P.pushEdge(0, 3, 2, 0); P.pushEdge(1, 3, 0, 1); P.pushEdge(2, 3, 1, 2);

    uint32_t ci = 0;
    for (uint32_t it = 0; it < 20; ++it)
    {
        ci = P.closest();
        if (ci == 0xFFFFFFFFu) { ++nc.asserts; ci = 0; break; }
        const EpaTri& t = P.tris[ci];
        const f3 tn = mk3(t.nx, t.ny, t.nz);
        SupPt a = minkowski(A, B, tn);
        float dd = dot(a.m, tn);
        if (dd - t.dist < 0.01f) break;
        if (!P.addPoint(a)) { ++nc.epaOverflow; break; }
    }
    const EpaTri& t = P.tris[ci];
    const f3 tn = mk3(t.nx, t.ny, t.nz);
    const SupPt& a = P.pts[t.a]; const SupPt& b = P.pts[t.b]; const SupPt& c = P.pts[t.c];
    f3 bc = barycentric(a.m, b.m, c.m, tn * t.dist);
    f3 pA = bc.x * a.a + bc.y * b.a + bc.z * c.a;
    f3 pB = bc.x * a.b + bc.y * b.b + bc.z * c.b;
    outPoint = 0.5f * (pA + pB);
    outNormal = tn;
    outDepth = t.dist;
}

// ---- small polytope in shared memory (PHASE 4). Still one thread per pair, but the polytope lives in a word-interleaved shared
// memory column (word k of lane l at [k * 32 + l]: whatever triangle a lane indexes, it stays in its own bank) instead of 7.7 KB of
// thread-local memory that 1024 resident threads per SM turn into DRAM traffic. Measured on the settled mixed pile (oracle
// histogram, tests/epa_histogram.py): 98.6 % of all EPA runs end with <= 10 points and <= 30 triangles (edges = triangles + 2), so the
// column holds exactly that much (242 words = 968 B per pair, 7 warps per SM); a run that needs more is abandoned and re-run from
// its simplex by the large-capacity kernel, which gives the same result because the expansion is deterministic.
// Indices are bytes (0xFF stands for the 0xFFFF a degenerate horizon leaves behind), the activity flags one register, the per-edge
// reference counts of collision_epa.cpp:114-139 two register bit masks (seen once / seen more often).
// Two sizes are instantiated (k_narrow.cu): <10, 30, 32> for everybody (242 words, 7 warps per SM, masks in 32-bit registers) and
// <16, 62, 64> for the runs that outgrow it (532 words, 3 warps per SM, 64-bit masks); what outgrows that too goes to the warp-cooperative kernel.
template <uint32_t MAXP, uint32_t MAXT, uint32_t MAXE>
struct EpaSm
{
    static_assert(MAXT <= 64 && MAXE <= 64 && MAXP <= 16 && MAXE < 127, "masks are one machine word, indices bytes");
    typedef typename std::conditional<(MAXT > 32 || MAXE > 32), unsigned long long, uint32_t>::type Mask;
    // a triangle is normal.xyz, distance and its three point + three edge indices: with <= 16 points and <= 32 edges the six indices fit ONE word
    // (4 bits per point, 6 bits per edge, 63 = none), which makes the small polytope 242 words per pair instead of 272: 7 warps per SM instead of 6
    static constexpr bool kPacked = MAXP <= 16 && MAXE <= 32;
    static constexpr uint32_t kTriW = kPacked ? 5u : 6u;
    static constexpr uint32_t kWords = MAXP * 6 + MAXT * kTriW + MAXE;
    uint32_t* w;             // this lane's column
    Mask active; uint32_t nT, nP, nE;
    static constexpr uint32_t kTri = MAXP * 6, kEdge = kTri + MAXT * kTriW;
    B200_D static uint32_t triAt(uint32_t i) { return kTri + kTriW * i; }
    B200_D static uint32_t lowest(Mask m) { return sizeof(Mask) == 8 ? (uint32_t)__ffsll((long long)m) - 1u : (uint32_t)__ffs((int)(uint32_t)m) - 1u; }
    B200_D static Mask bit(uint32_t i) { return (Mask)1 << i; }
    B200_D float ldf(uint32_t k) const { return __uint_as_float(w[k << 5]); }
    B200_D uint32_t ldu(uint32_t k) const { return w[k << 5]; }
    B200_D void stf(uint32_t k, float v) { w[k << 5] = __float_as_uint(v); }
    B200_D void stu(uint32_t k, uint32_t v) { w[k << 5] = v; }
    B200_D static uint32_t wide(uint32_t b) { return b == 0xFFu ? 0xFFFFu : b; }           // byte -> the reference's uint16 index
    B200_D bool isActive(uint32_t i) const { return i < nT && ((active >> i) & (Mask)1); }
    // point indices as bytes a | b << 8 | c << 16; edge indices as bytes eA | eB << 8 | eC << 16 (anything >= nE: none)
    B200_D uint32_t triIdx(uint32_t i) const
    {
        const uint32_t v = ldu(triAt(i) + 4);
        return kPacked ? ((v & 0xFu) | ((v & 0xF0u) << 4) | ((v & 0xF00u) << 8)) : v;
    }
    B200_D uint32_t triEdges(uint32_t i) const
    {
        if (!kPacked) return ldu(triAt(i) + 5);
        const uint32_t v = ldu(triAt(i) + 4) >> 12;
        return (v & 63u) | ((v & (63u << 6)) << 2) | ((v & (63u << 12)) << 4);
    }
    B200_D void setTriEdgeB(uint32_t i, uint32_t e)
    {
        if (kPacked) { const uint32_t k = triAt(i) + 4; stu(k, (ldu(k) & ~(63u << 18)) | ((e & 63u) << 18)); }
        else { const uint32_t k = triAt(i) + 5; stu(k, (ldu(k) & 0xFFFF00FFu) | (e << 8)); }
    }
    B200_D SupPt pt(uint32_t i) const
    {
        SupPt p; const uint32_t k = 6 * i;
        p.a = mk3(ldf(k), ldf(k + 1), ldf(k + 2)); p.b = mk3(ldf(k + 3), ldf(k + 4), ldf(k + 5)); p.m = p.a - p.b;      // m = a - b as minkowski() computed it
        return p;
    }
    B200_D f3 ptM(uint32_t i) const { const uint32_t k = 6 * i; return mk3(ldf(k), ldf(k + 1), ldf(k + 2)) - mk3(ldf(k + 3), ldf(k + 4), ldf(k + 5)); }
    B200_D bool pushPoint(const SupPt& p)
    {
        if (nP >= MAXP) return false;
        const uint32_t k = 6 * nP; stf(k, p.a.x); stf(k + 1, p.a.y); stf(k + 2, p.a.z); stf(k + 3, p.b.x); stf(k + 4, p.b.y); stf(k + 5, p.b.z);
        ++nP; return true;
    }
    B200_D bool pushTri(uint32_t a, uint32_t b, uint32_t c, uint32_t eA, uint32_t eB, uint32_t eC, f3 n, float d)
    {
        if (nT >= MAXT) return false;
        const uint32_t k = triAt(nT);
        stf(k, n.x); stf(k + 1, n.y); stf(k + 2, n.z); stf(k + 3, d);
        if (kPacked) stu(k + 4, (a & 0xFu) | ((b & 0xFu) << 4) | ((c & 0xFu) << 8) | ((eA & 63u) << 12) | ((eB & 63u) << 18) | ((eC & 63u) << 24));
        else { stu(k + 4, a | (b << 8) | (c << 16)); stu(k + 5, (eA & 0xFFu) | ((eB & 0xFFu) << 8) | ((eC & 0xFFu) << 16)); }
        active |= bit(nT); ++nT; return true;
    }
    B200_D bool pushEdge(uint32_t a, uint32_t b, uint32_t t0, uint32_t t1)
    {
        if (nE >= MAXE) return false;
        stu(kEdge + nE, a | (b << 8) | ((t0 & 0xFFu) << 16) | ((t1 & 0xFFu) << 24));
        ++nE; return true;
    }
    B200_D uint32_t closest() const
    {   // collision_epa.cpp:94-112
        uint32_t best = 0xFFFFFFFFu; float mind = FLT_MAX;
        for (Mask rest = active; rest; rest &= rest - 1)
        {
            const uint32_t i = lowest(rest);
            const float d = ldf(triAt(i) + 3);
            if (d < mind) { mind = d; best = i; }
        }
        return best;
    }
    // the tetrahedron GJK ended with (collision_epa.h:98-116)
    B200_D void init(const SimplexD& g)
    {
        active = 0; nT = nP = nE = 0;
        pushPoint(g.a); pushPoint(g.b); pushPoint(g.c); pushPoint(g.d);
        f3 n; float d;
        EpaPoly::triInfo(g.a, g.b, g.d, n, d); pushTri(0, 1, 3, 4, 3, 0, n, d);
        EpaPoly::triInfo(g.b, g.c, g.d, n, d); pushTri(1, 2, 3, 5, 4, 1, n, d);
        EpaPoly::triInfo(g.c, g.a, g.d, n, d); pushTri(2, 0, 3, 3, 5, 2, n, d);
        EpaPoly::triInfo(g.a, g.c, g.b, n, d); pushTri(0, 2, 1, 1, 0, 2, n, d);
        pushEdge(0, 1, 0, 3); pushEdge(1, 2, 1, 3); pushEdge(2, 0, 2, 3);
        pushEdge(0, 3, 2, 0); pushEdge(1, 3, 0, 1); pushEdge(2, 3, 1, 2);
    }
    // collision_epa.h:148-167: contact from the closest triangle
    B200_D void result(uint32_t ci, f3& outPoint, f3& outNormal, float& outDepth) const
    {
        const uint32_t k = triAt(ci);
        const f3 tn = mk3(ldf(k), ldf(k + 1), ldf(k + 2));
        const float tdist = ldf(k + 3);
        const uint32_t idx = triIdx(ci);
        const SupPt a = pt(idx & 0xFFu), b = pt((idx >> 8) & 0xFFu), c = pt((idx >> 16) & 0xFFu);
        f3 bc = barycentric(a.m, b.m, c.m, tn * tdist);
        f3 pA = bc.x * a.a + bc.y * b.a + bc.z * c.a;
        f3 pB = bc.x * a.b + bc.y * b.b + bc.z * c.b;
        outPoint = 0.5f * (pA + pB);
        outNormal = tn;
        outDepth = tdist;
    }
    // collision_epa.cpp:114-239 ; false = a capacity of the SMALL polytope was hit (the caller hands the pair to the large kernel)
    B200_D bool addPoint(const SupPt& np)
    {
        Mask once = 0, more = 0;
        for (Mask rest = active; rest; rest &= rest - 1)
        {
            const uint32_t i = lowest(rest), k = triAt(i);
            const uint32_t idx = triIdx(i);
            const float d = dot(mk3(ldf(k), ldf(k + 1), ldf(k + 2)), np.m - ptM(idx & 0xFFu));
            if (d > 0.f)
            {
                const uint32_t e = triEdges(i);
#pragma unroll
                for (int j = 0; j < 3; ++j)
                {
                    const uint32_t ei = (e >> (8 * j)) & 0xFFu;
                    if (ei < nE) { const Mask b = bit(ei); more |= once & b; once |= b; }
                }
                active &= ~bit(i);
            }
        }
        const Mask border = once & ~more;                    // refs == 1, ascending edge index
        unsigned long long nepp0 = ~0ull, nepp1 = ~0ull;         // newEdgePerPoint: 7 bits per point (8 points per word), 127 = none
        const uint32_t npi = nP;
        if (!pushPoint(np)) return false;
        const uint32_t triOffset = nT;
        for (Mask rest = border; rest; rest &= rest - 1)
        {
            const uint32_t ei = lowest(rest);
            const uint32_t e = ldu(kEdge + ei);
            const uint32_t ea = e & 0xFFu, eb = (e >> 8) & 0xFFu;
            const bool aAct = isActive((e >> 16) & 0xFFu), bAct = isActive(e >> 24);
            const uint32_t connect = bAct ? ea : eb;
            const uint32_t triIndex = nT, ne = nE;
            if (!pushEdge(connect, npi, 0xFFu, triIndex)) return false;
            { const uint32_t sh = 7u * (connect & 7u); unsigned long long& word = connect < 8u ? nepp0 : nepp1; word = (word & ~(127ull << sh)) | ((unsigned long long)ne << sh); }
            const uint32_t bI = connect, cI = bAct ? eb : ea;
            f3 n; float d; EpaPoly::triInfo(np, pt(bI), pt(cI), n, d);
            if (!pushTri(npi, bI, cI, ei, 0xFFu, ne, n, d)) return false;
            const uint32_t e2 = ldu(kEdge + ei);
            stu(kEdge + ei, aAct ? ((e2 & 0x00FFFFFFu) | (triIndex << 24)) : ((e2 & 0xFF00FFFFu) | (triIndex << 16)));
        }
        uint32_t i = 0;
        for (Mask rest = border; rest; rest &= rest - 1, ++i)
        {
            const uint32_t ei = lowest(rest);
            const uint32_t e = ldu(kEdge + ei);
            const bool bNew = wide(e >> 24) >= triOffset;
            const uint32_t connect = bNew ? (e & 0xFFu) : ((e >> 8) & 0xFFu);
            const uint32_t other = (uint32_t)((connect < 8u ? nepp0 : nepp1) >> (7u * (connect & 7u))) & 127u;
            const uint32_t triIndex = i + triOffset;
            if (other == 127u) continue;
            setTriEdgeB(triIndex, other);
            stu(kEdge + other, (ldu(kEdge + other) & 0xFF00FFFFu) | (triIndex << 16));
        }
        return true;
    }
};

typedef EpaSm<10, 30, 32> EpaSmall;
typedef EpaSm<16, 62, 64> EpaMedium;
#define B200_EPAS_WORDS (EpaSmall::kWords)

// ---- warp-cooperative EPA (PHASE 3). One warp expands ONE polytope that lives in shared memory; all 32 lanes execute the calling
// routine uniformly (same pair, same data) and split the per-triangle / per-edge loops of collision_epa.cpp:94-239 among themselves.
// Every step reproduces the sequential result exactly: the closest triangle is the first minimum (ties -> lowest index), horizon edges
// are compacted in ascending edge order (ballot prefix), the i-th horizon edge creates edge nE + i and triangle nT + i like the
// sequential pushes do, and the two places where a later loop iteration overwrites an earlier one (newEdgePerPoint, edges[other].t0)
// become max-reductions over the iteration index. Capacity failures are detected before anything is appended: the sequential code
// would bail out half way, but its result only depends on the closest triangle found BEFORE the failed expansion.
// Capacities follow the reference where it has a real bound (<= 128 border edges, collision_epa.cpp:141; 4 + 20 points) and are sized
// for the observed maxima elsewhere (the reference's arrays hold 1024).
#define B200_EPAW_MAX_POINTS 24
#define B200_EPAW_MAX_TRIS 320
#define B200_EPAW_MAX_EDGES 480
#define B200_EPAW_MAX_BORDER 128
struct EpaWarpShared
{
    float4 triN[B200_EPAW_MAX_TRIS];                    // normal.xyz, distance
    float pts[(B200_EPAW_MAX_POINTS + 1) * 9];          // a.xyz, b.xyz, m.xyz per support point
    int32_t refs[B200_EPAW_MAX_EDGES];                  // per-round reference counts, then the link targets of the second pass
    uint16_t triA[B200_EPAW_MAX_TRIS], triB[B200_EPAW_MAX_TRIS], triC[B200_EPAW_MAX_TRIS], triEA[B200_EPAW_MAX_TRIS], triEB[B200_EPAW_MAX_TRIS], triEC[B200_EPAW_MAX_TRIS];
    uint16_t edgeA[B200_EPAW_MAX_EDGES], edgeB[B200_EPAW_MAX_EDGES], edgeT0[B200_EPAW_MAX_EDGES], edgeT1[B200_EPAW_MAX_EDGES];
    uint16_t border[B200_EPAW_MAX_BORDER];
    int32_t newEdgePerPoint[B200_EPAW_MAX_POINTS + 1];
    uint32_t active[B200_EPAW_MAX_TRIS / 32];
};

struct EpaWarp
{
    EpaWarpShared& S; uint32_t lane; uint32_t nT, nP, nE;
    B200_D EpaWarp(EpaWarpShared& s) : S(s), lane(threadIdx.x & 31u), nT(0), nP(0), nE(0) {}
    B200_D bool isActive(uint32_t i) const { return i < nT && ((S.active[i >> 5] >> (i & 31)) & 1u); }
    B200_D f3 ptM(uint32_t i) const { return mk3(S.pts[9 * i + 6], S.pts[9 * i + 7], S.pts[9 * i + 8]); }
    B200_D SupPt pt(uint32_t i) const { SupPt p; const float* f = S.pts + 9 * i; p.a = mk3(f[0], f[1], f[2]); p.b = mk3(f[3], f[4], f[5]); p.m = mk3(f[6], f[7], f[8]); return p; }
    B200_D void setPoint(uint32_t i, const SupPt& p)
    {   // lanes 0..8 store one float each (all lanes hold the same p)
        const float v[9] = { p.a.x, p.a.y, p.a.z, p.b.x, p.b.y, p.b.z, p.m.x, p.m.y, p.m.z };
        if (lane < 9) S.pts[9 * i + lane] = v[lane];
    }
    B200_D void setTri(uint32_t i, uint16_t a, uint16_t b, uint16_t c, uint16_t eA, uint16_t eB, uint16_t eC, f3 n, float d)
    { S.triA[i] = a; S.triB[i] = b; S.triC[i] = c; S.triEA[i] = eA; S.triEB[i] = eB; S.triEC[i] = eC; S.triN[i] = make_float4(n.x, n.y, n.z, d); }
    B200_D void setEdge(uint32_t i, uint16_t a, uint16_t b, uint16_t t0, uint16_t t1) { S.edgeA[i] = a; S.edgeB[i] = b; S.edgeT0[i] = t0; S.edgeT1[i] = t1; }

    B200_D void init(const SimplexD& g)
    {
        if (lane < B200_EPAW_MAX_TRIS / 32) S.active[lane] = 0;
        setPoint(0, g.a); setPoint(1, g.b); setPoint(2, g.c); setPoint(3, g.d);
        if (lane == 0)
        {
            f3 n; float d;
            EpaPoly::triInfo(g.a, g.b, g.d, n, d); setTri(0, 0, 1, 3, 4, 3, 0, n, d);
            EpaPoly::triInfo(g.b, g.c, g.d, n, d); setTri(1, 1, 2, 3, 5, 4, 1, n, d);
            EpaPoly::triInfo(g.c, g.a, g.d, n, d); setTri(2, 2, 0, 3, 3, 5, 2, n, d);
            EpaPoly::triInfo(g.a, g.c, g.b, n, d); setTri(3, 0, 2, 1, 1, 0, 2, n, d);
            setEdge(0, 0, 1, 0, 3); setEdge(1, 1, 2, 1, 3); setEdge(2, 2, 0, 2, 3);
            setEdge(3, 0, 3, 2, 0); setEdge(4, 1, 3, 0, 1); setEdge(5, 2, 3, 1, 2);
            S.active[0] = 0xFu;
        }
        nP = 4; nT = 4; nE = 6;
        __syncwarp();
    }
    // collision_epa.cpp:94-112: first minimum of dist over the active triangles
    B200_D uint32_t closest() const
    {
        uint32_t best = 0xFFFFFFFFu; float mind = FLT_MAX;
        for (uint32_t i = lane; i < nT; i += 32u) { const float d = S.triN[i].w; if (isActive(i) && d < mind) { mind = d; best = i; } }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            const float od = __shfl_xor_sync(0xffffffffu, mind, o); const uint32_t ob = __shfl_xor_sync(0xffffffffu, best, o);
            if (ob != 0xFFFFFFFFu && (best == 0xFFFFFFFFu || od < mind || (od == mind && ob < best))) { mind = od; best = ob; }
        }
        return best;
    }
    // collision_epa.cpp:114-239
    B200_D bool addPoint(const SupPt& np)
    {
        const uint32_t FULL = 0xffffffffu;
        for (uint32_t i = lane; i < nE; i += 32u) S.refs[i] = 0;
        for (uint32_t i = lane; i <= nP && i < B200_EPAW_MAX_POINTS + 1; i += 32u) S.newEdgePerPoint[i] = -1;
        __syncwarp();
        // triangles that see the new point are removed; their edges are counted
        for (uint32_t base = 0; base < nT; base += 32u)
        {
            const uint32_t i = base + lane;
            bool removed = false;
            if (i < nT && isActive(i))
            {
                const float4 t = S.triN[i];
                const float d = dot(mk3(t.x, t.y, t.z), np.m - ptM(S.triA[i]));
                if (d > 0.f)
                {
                    removed = true;
                    const uint32_t eA = S.triEA[i], eB = S.triEB[i], eC = S.triEC[i];
                    if (eA < nE) atomicAdd(&S.refs[eA], 1); if (eB < nE) atomicAdd(&S.refs[eB], 1); if (eC < nE) atomicAdd(&S.refs[eC], 1);
                }
            }
            const uint32_t rem = __ballot_sync(FULL, removed);
            if (lane == 0 && rem) S.active[base >> 5] &= ~rem;
        }
        __syncwarp();
        // horizon = edges referenced exactly once, in ascending edge order
        uint32_t nBorder = 0;
        for (uint32_t base = 0; base < nE; base += 32u)
        {
            const uint32_t e = base + lane;
            const bool flag = e < nE && S.refs[e] == 1;
            const uint32_t b = __ballot_sync(FULL, flag);
            const uint32_t pos = nBorder + (uint32_t)__popc(b & ((1u << lane) - 1u));
            if (flag && pos < B200_EPAW_MAX_BORDER) S.border[pos] = (uint16_t)e;
            nBorder += (uint32_t)__popc(b);
        }
        if (nBorder > B200_EPAW_MAX_BORDER) return false;
        if (nP >= B200_EPAW_MAX_POINTS) return false;
        if (nE + nBorder > B200_EPAW_MAX_EDGES || nT + nBorder > B200_EPAW_MAX_TRIS) return false;
        __syncwarp();
        const uint32_t npi = nP, triOffset = nT, edgeOffset = nE;
        setPoint(npi, np);
        // one new edge and one new triangle per horizon edge
        for (uint32_t base = 0; base < nBorder; base += 32u)
        {
            const uint32_t i = base + lane;
            if (i < nBorder)
            {
                const uint32_t ei = S.border[i];
                const uint16_t ea = S.edgeA[ei], eb = S.edgeB[ei];
                const bool aAct = isActive(S.edgeT0[ei]), bAct = isActive(S.edgeT1[ei]);
                const uint16_t connect = bAct ? ea : eb;
                const uint32_t triIndex = triOffset + i, ne = edgeOffset + i;
                setEdge(ne, connect, (uint16_t)npi, 0xFFFF, (uint16_t)triIndex);
                atomicMax(&S.newEdgePerPoint[connect], (int32_t)ne);            // the sequential loop keeps the LAST edge written per point
                const uint16_t bI = connect, cI = bAct ? eb : ea;
                f3 n; float d; EpaPoly::triInfo(np, pt(bI), pt(cI), n, d);
                setTri(triIndex, (uint16_t)npi, bI, cI, (uint16_t)ei, 0xFFFF, (uint16_t)ne, n, d);
                if (aAct) S.edgeT1[ei] = (uint16_t)triIndex; else S.edgeT0[ei] = (uint16_t)triIndex;
            }
        }
        nP = npi + 1; nT = triOffset + nBorder; nE = edgeOffset + nBorder;
        __syncwarp();
        // activity bits of the new triangles; link targets start empty
        for (uint32_t i = triOffset + lane; i < nT; i += 32u) atomicOr(&S.active[i >> 5], 1u << (i & 31u));
        for (uint32_t i = edgeOffset + lane; i < nE; i += 32u) S.refs[i] = -1;
        __syncwarp();
        // second pass: connect neighbouring new triangles through the new edges
        for (uint32_t base = 0; base < nBorder; base += 32u)
        {
            const uint32_t i = base + lane;
            if (i < nBorder)
            {
                const uint32_t ei = S.border[i];
                const bool bNew = S.edgeT1[ei] >= triOffset;
                const uint16_t connect = bNew ? S.edgeA[ei] : S.edgeB[ei];
                const int32_t other = S.newEdgePerPoint[connect];
                const uint32_t triIndex = triOffset + i;
                if (other >= 0) { S.triEB[triIndex] = (uint16_t)other; atomicMax(&S.refs[other], (int32_t)triIndex); }   // later iterations overwrite earlier ones
            }
        }
        __syncwarp();
        for (uint32_t i = edgeOffset + lane; i < nE; i += 32u) if (S.refs[i] >= 0) S.edgeT0[i] = (uint16_t)S.refs[i];
        __syncwarp();
        return true;
    }
};

// collision_epa.h:96-168 with the polytope shared by the warp; every lane returns the same result
B200_D void epaWarp(EpaWarpShared& sh, const SimplexD& g, const SupportD& A, const SupportD& B, f3& outPoint, f3& outNormal, float& outDepth, NarrowCounters& nc)
{
    EpaWarp P(sh);
    __syncwarp();
    P.init(g);
    uint32_t ci = 0;
    for (uint32_t it = 0; it < 20; ++it)
    {
        ci = P.closest();
        if (ci == 0xFFFFFFFFu) { if (P.lane == 0) ++nc.asserts; ci = 0; break; }
        const float4 t = sh.triN[ci];
        const f3 tn = mk3(t.x, t.y, t.z);
        SupPt a = minkowskiWarp(A, B, tn);
        float dd = dot(a.m, tn);
        if (dd - t.w < 0.01f) break;
        if (!P.addPoint(a)) { if (P.lane == 0) ++nc.epaOverflow; break; }
    }
    const float4 t = sh.triN[ci];
    const f3 tn = mk3(t.x, t.y, t.z);
    const SupPt a = P.pt(sh.triA[ci]), b = P.pt(sh.triB[ci]), c = P.pt(sh.triC[ci]);
    f3 bc = barycentric(a.m, b.m, c.m, tn * t.w);
    f3 pA = bc.x * a.a + bc.y * b.a + bc.z * c.a;
    f3 pB = bc.x * a.b + bc.y * b.b + bc.z * c.b;
    outPoint = 0.5f * (pA + pB);
    outNormal = tn;
    outDepth = t.w;
    __syncwarp();
}

B200_D void storeSimplex(float4* __restrict__ pl, uint32_t cap, uint32_t e, const SimplexD& s)
{
    const SupPt* pt[4] = { &s.a, &s.b, &s.c, &s.d };
    float v[36];
#pragma unroll
    for (int i = 0; i < 4; ++i) { v[9*i] = pt[i]->a.x; v[9*i+1] = pt[i]->a.y; v[9*i+2] = pt[i]->a.z; v[9*i+3] = pt[i]->b.x; v[9*i+4] = pt[i]->b.y; v[9*i+5] = pt[i]->b.z; v[9*i+6] = pt[i]->m.x; v[9*i+7] = pt[i]->m.y; v[9*i+8] = pt[i]->m.z; }
#pragma unroll
    for (int k = 0; k < 9; ++k) pl[(size_t)k * cap + e] = make_float4(v[4*k], v[4*k+1], v[4*k+2], v[4*k+3]);
}
B200_D void loadSimplex(const float4* __restrict__ pl, uint32_t cap, uint32_t e, SimplexD& s)
{
    float v[36];
#pragma unroll
    for (int k = 0; k < 9; ++k) { const float4 t = pl[(size_t)k * cap + e]; v[4*k] = t.x; v[4*k+1] = t.y; v[4*k+2] = t.z; v[4*k+3] = t.w; }
    SupPt* pt[4] = { &s.a, &s.b, &s.c, &s.d };
#pragma unroll
    for (int i = 0; i < 4; ++i) { pt[i]->a = mk3(v[9*i], v[9*i+1], v[9*i+2]); pt[i]->b = mk3(v[9*i+3], v[9*i+4], v[9*i+5]); pt[i]->m = mk3(v[9*i+6], v[9*i+7], v[9*i+8]); }
    s.n = 4;
}

template <class NC>
B200_D bool gjkEpa(const SupportD& A, const SupportD& B, Man& m, NC& nc)
{
    SimplexD s;
    if (NC::phase == 1)
    {
        if (!gjk(A, B, s, nc)) return false;
        // warp-aggregated append to the EPA work list
        const uint32_t mask = __activemask();
        const uint32_t lane = threadIdx.x & 31u, leader = (uint32_t)__ffs((int)mask) - 1u;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(nc.sxCounter, (uint32_t)__popc(mask));
        base = __shfl_sync(mask, base, (int)leader);
        const uint32_t e = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
        nc.sxEntry = e; nc.deferred = true;
        if (e < nc.sxCap) storeSimplex(nc.sxPlanes, nc.sxCap, e, s);
        return false;
    }
    loadSimplex(nc.sxPlanes, nc.sxCap, nc.sxEntry, s);
    f3 p, n; float d;
    if (NC::phase == 3) epaWarp(*nc.epaShared, s, A, B, p, n, d, nc);
    else epa(s, A, B, p, n, d, nc);
    m.n = n; m.cnt = 1; m.d[0] = d; m.p[0] = p;
    return true;
}

// ------------------------------------------------------------------------------------- GJK/EPA-based shape pairs
// Every GJK based routine of the reference is "set up two support mappings -> gjkIntersectionTest -> epaCollisionInfo -> routine
// specific tail". The three parts are separate functions here so that a kernel has ONE copy of GJK / EPA (the per-routine copies
// made the kernels 200-300 KB of code and instruction fetch their top stall) and can run them as a loop over lanes that are at
// different stages of different pairs (k_narrow.cu).
struct GjkJob
{
    SupportD A, B;
    uint32_t tail;      // 0 none; 1 segment shape vs box: 2-point clip against the box face; 2 the same, then back out of the OBB frame
    q4 oq; f3 oc;       // tail 2: the OBB frame
};

// capsule-aabb (:705-769) / cylinder-aabb (:953-1022): before GJK
B200_D void segmentShapeBoxPre(const CapS& c, bool isCylinder, const BoxS& a, GjkJob& job)
{
    job.A = isCylinder ? supCylinder(c) : supCapsule(c);
    job.B = supBox(a);
    job.tail = 1;
}
// capsule-obb (:771-790) / cylinder-obb (:1024-1043): the segment shape moves into the OBB's frame
B200_D void segmentShapeObbPre(const CapS& c, bool isCylinder, const ObbS& o, GjkJob& job)
{
    BoxS bb; bb.mn = o.c - o.r; bb.mx = o.c + o.r;
    CapS c_; c_.a = conjugate(o.q) * (c.a - o.c) + o.c; c_.b = conjugate(o.q) * (c.b - o.c) + o.c; c_.r = c.r;
    segmentShapeBoxPre(c_, isCylinder, bb, job);
    job.tail = 2; job.oq = o.q; job.oc = o.c;
}
// collision_narrow.cpp:821-951 before GJK: 0 = no contact, 1 = manifold complete (parallel axes: closed form), 2 = GJK job ready
B200_D int cylinderCylinderPre(const CapS& a, const CapS& b, Man& m, NarrowCounters& nc, GjkJob& job)
{
    f3 aDir = a.b - a.a;
    f3 bDir = normalize(b.b - b.a);
    float aLen = len(aDir);
    aDir *= 1.f / aLen;
    float parallel = dot(aDir, bDir);
    if (fabsf(parallel) > 0.99f)
    {
        f3 pBa = b.a, pBb = b.b;
        if (parallel < 0.f) { f3 t = pBa; pBa = pBb; pBb = t; }
        f3 ref = a.a;
        float a0 = 0.f, a1 = aLen;
        float b0 = dot(aDir, pBa - ref), b1 = dot(aDir, pBb - ref);
        if (!(b1 > b0)) ++nc.asserts;
        float left = fmax2(a0, b0), right = fmin2(a1, b1);
        if (right < left) return 0;
        f3 cA0 = ref + left * aDir, cA1 = ref + right * aDir;
        f3 cB0 = closestPointSegment(cA0, pBa, pBb);
        f3 cB1 = cB0 + (right - left) * aDir;
        f3 normal = cB0 - cA0;
        float d = len(normal);
        float rs = a.r + b.r;
        float pen = rs - d;
        if (pen < 0.f) return 0;
        float capPen = right - left;
        if (!(capPen > 0.f)) ++nc.asserts;
        if (capPen < pen)
        {
            m.cnt = 1; m.d[0] = capPen;
            if (b0 > a0) { m.n = aDir; m.p[0] = a.b - mk3(capPen * 0.5f); }      // scalar broadcast, SURVEY quirk B2
            else { m.n = -aDir; m.p[0] = a.a + mk3(capPen * 0.5f); }
        }
        else
        {
            if (d < B200_EPS) { d = 0.f; normal = mk3(0.f, 1.f, 0.f); } else { normal /= d; }
            m.n = normal; m.cnt = 2;
            m.d[0] = pen; m.p[0] = (cA0 + cB0) * 0.5f;
            m.d[1] = pen; m.p[1] = (cA1 + cB1) * 0.5f;
        }
        return 1;
    }
    job.A = supCylinder(a); job.B = supCylinder(b); job.tail = 0;
    return 2;
}
// after EPA (m holds its one contact): the routine's own tail
B200_D void gjkTail(const GjkJob& job, Man& m)
{
    if (job.tail == 0) return;
    {   // :728-766 / :976-1019: an axis aligned normal and the segment parallel to the face -> two clipped points instead of one
        CapS c; c.a = job.A.a; c.b = job.A.b; c.r = job.A.r;
        BoxS a; a.mn = job.B.a; a.mx = job.B.b;
        f3 normal = m.n;
        if (fabsf(normal.x) > 0.99f || fabsf(normal.y) > 0.99f || fabsf(normal.z) > 0.99f)
        {
            f3 axis = normalize(c.b - c.a);
            if (fabsf(dot(normal, axis)) < 0.01f)
            {
                f3 cpp[4], cpn[4]; PlaneS planes[4];
                f3 boxNormal = -normal;
                PlaneS ref = boxReferencePlane(a, boxNormal);
                PolyD poly; poly.n = 2;
                f3 pa = c.a + normal * c.r, pb = c.b + normal * c.r;
                poly.p[0].v = pa; poly.p[0].depth = -signedDistance(pa, ref);
                poly.p[1].v = pb; poly.p[1].depth = -signedDistance(pb, ref);
                f3 ac = boxCenter(a);
                boxClipPlanes(boxRadius(a), boxNormal, cpp, cpn);
                for (int i = 0; i < 4; ++i) { cpp[i] = cpp[i] + ac; planes[i] = createPlane(cpp[i], cpn[i]); }
                clipAndBuild(poly, planes, 4, ref, m);
            }
        }
    }
    if (job.tail == 2)
    {
        m.n = job.oq * m.n;
        for (uint32_t i = 0; i < m.cnt; ++i) m.p[i] = job.oq * (m.p[i] - job.oc) + job.oc;
    }
}

// ------------------------------------------------------------------------------------------------------ OBB - OBB
// collision_narrow.cpp:1179-1527 : 15-axis SAT in A's frame, then face clipping or edge-edge closest points
B200_D bool obbObb(const ObbS& a, const ObbS& b, Man& m)
{
    f3 axA[3] = { a.q * mk3(1.f, 0.f, 0.f), a.q * mk3(0.f, 1.f, 0.f), a.q * mk3(0.f, 0.f, 1.f) };
    f3 axB[3] = { b.q * mk3(1.f, 0.f, 0.f), b.q * mk3(0.f, 1.f, 0.f), b.q * mk3(0.f, 0.f, 1.f) };
    m3 r;
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
        for (int i = 0; i < 3; ++i) matRef(r, i, j) = dot(axA[i], axB[j]);
    f3 tw = b.c - a.c;
    f3 t = conjugate(a.q) * tw;
    bool parallel = false;
    m3 absR;
#pragma unroll
    for (int i = 0; i < 9; ++i) { absR.m[i] = fabsf(r.m[i]) + B200_EPS; if (absR.m[i] >= 0.99f) parallel = true; }

    float ra, rb, minPen = FLT_MAX;
    f3 normal = mk3(0.f);
    bool bFace = false;
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        ra = comp(a.r, i); rb = dot(m3Row(absR, i), b.r);
        float pen = ra + rb - fabsf(comp(t, i));
        if (pen < 0.f) return false;
        if (pen < minPen) { minPen = pen; normal = mk3(0.f); setComp(normal, i, 1.f); }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {
        ra = dot(m3Col(absR, i), a.r); rb = comp(b.r, i);
        float d = dot(m3Col(r, i), t);
        float pen = ra + rb - fabsf(d);
        if (pen < 0.f) return false;
        if (pen < minPen) { minPen = pen; normal = mk3(0.f); setComp(normal, i, 1.f); bFace = true; }
    }
    bool edgeCol = false; f3 edgeN = mk3(0.f);
    if (!parallel)
    {
        // edge axis a_i x b_j, i-major like the reference (:1275-1408); i1 = (i+1)%3, i2 = (i+2)%3 and likewise for j
#pragma unroll
        for (int i = 0; i < 3; ++i)
        {
            const int i1 = (i + 1) % 3, i2 = (i + 2) % 3;
#pragma unroll
            for (int j = 0; j < 3; ++j)
            {
                const int j1 = (j + 1) % 3, j2 = (j + 2) % 3;
                // ra = a.r[i1] * |R[i2][j]| + a.r[i2] * |R[i1][j]| ; rb = b.r[jlo] * |R[i][jhi]| + b.r[jhi] * |R[i][jlo]| with (jlo < jhi) = the two axes != j
                const int jlo = (j1 < j2) ? j1 : j2, jhi = (j1 < j2) ? j2 : j1;
                const int ilo = (i1 < i2) ? i1 : i2, ihi = (i1 < i2) ? i2 : i1;
                ra = comp(a.r, ilo) * mat(absR, ihi, j) + comp(a.r, ihi) * mat(absR, ilo, j);
                rb = comp(b.r, jlo) * mat(absR, i, jhi) + comp(b.r, jhi) * mat(absR, i, jlo);
                // |t[i2] * R[i1][j] - t[i1] * R[i2][j]|
                float dist = fabsf(comp(t, i2) * mat(r, i1, j) - comp(t, i1) * mat(r, i2, j));
                float pen = ra + rb - dist;
                if (pen < 0.f) return false;
                f3 n = mk3(0.f); setComp(n, i1, -mat(r, i2, j)); setComp(n, i2, mat(r, i1, j));
                float l = 1.f / len(n);
                pen *= l;
                if (pen < minPen) { minPen = pen; edgeN = n * l; edgeCol = true; }
            }
        }
    }
    const bool faceCol = !edgeCol;
    if (faceCol) { if (bFace) normal = r * normal; } else { normal = edgeN; }
    normal = a.q * normal;
    if (dot(normal, tw) < 0.f) normal = -normal;
    m.n = normal;

    if (faceCol)
    {
        f3 cpp[4], cpn[4]; PolyD poly; PlaneS plane;
        if (!bFace)
        {
            boxClipPlanes(a.r, conjugate(a.q) * normal, cpp, cpn);
            boxIncidentVerts(b.r, conjugate(b.q) * normal, poly);
            for (int i = 0; i < 4; ++i) { cpp[i] = a.q * cpp[i] + a.c; cpn[i] = a.q * cpn[i]; poly.p[i].v = b.q * poly.p[i].v + b.c; }
            f3 refPt = supObb(a).eval(normal);
            plane = createPlane(refPt, normal);
        }
        else
        {
            boxClipPlanes(b.r, conjugate(b.q) * -normal, cpp, cpn);
            boxIncidentVerts(a.r, conjugate(a.q) * -normal, poly);
            for (int i = 0; i < 4; ++i) { cpp[i] = b.q * cpp[i] + b.c; cpn[i] = b.q * cpn[i]; poly.p[i].v = a.q * poly.p[i].v + a.c; }
            f3 refPt = supObb(b).eval(-normal);
            plane = createPlane(refPt, -normal);
        }
        PlaneS planes[4];
        for (int i = 0; i < 4; ++i) { planes[i] = createPlane(cpp[i], cpn[i]); poly.p[i].depth = -signedDistance(poly.p[i].v, plane); }
        if (!clipAndBuild(poly, planes, 4, plane, m)) return false;
    }
    else
    {
        f3 a0, a1, b0, b1;
        boxIncidentEdge(a.r, conjugate(a.q) * normal, a0, a1);
        boxIncidentEdge(b.r, conjugate(b.q) * -normal, b0, b1);
        a0 = a.q * a0 + a.c; a1 = a.q * a1 + a.c; b0 = b.q * b0 + b.c; b1 = b.q * b1 + b.c;
        f3 pa, pb;
        float sq = closestSegmentSegment(a0, a1, b0, b1, pa, pb);
        m.cnt = 1; m.d[0] = sqrtf(sq); m.p[0] = (pa + pb) * 0.5f;
    }
    return true;
}

// --------------------------------------------------------------------------- boolean overlap tests (event layer)
// bounding_volumes.h:301-363, bounding_volumes.cpp:704-733,742-835,1079-1208,1237-1244: what overlapCheck
// (collision_narrow.cpp:1593-1689) dispatches to for rigid body vs trigger / force field pairs.
B200_D bool sphereVsSphereB(const SphereS& a, const SphereS& b) { const f3 d = a.c - b.c; const float rs = a.r + b.r; return dot(d, d) <= rs * rs; }
B200_D bool sphereVsCapsuleB(const SphereS& s, const CapS& c) { SphereS o; o.c = closestPointSegment(s.c, c.a, c.b); o.r = c.r; return sphereVsSphereB(s, o); }
B200_D bool sphereVsCylinderB(const SphereS& s, const CapS& c)
{
    const f3 ab = c.b - c.a;
    const float t = dot(s.c - c.a, ab) / sqLen(ab);
    if (t >= 0.f && t <= 1.f) { SphereS o; o.c = lerp(c.a, c.b, t); o.r = c.r; return sphereVsSphereB(s, o); }
    const f3 p = (t <= 0.f) ? c.a : c.b;
    const f3 up = (t <= 0.f) ? -ab : ab;
    const f3 proj = normalize(cross(cross(up, s.c - p), up));
    const f3 endA = p + proj * c.r, endB = p - proj * c.r;
    const f3 closest = closestPointSegment(s.c, endA, endB);
    return sqLen(closest - s.c) <= s.r;          // sic: radius, not radius squared (bounding_volumes.cpp:723)
}
B200_D bool sphereVsBoxB(const SphereS& s, const BoxS& a) { const f3 n = closestPointBox(s.c, a) - s.c; return sqLen(n) <= s.r * s.r; }
B200_D BoxS boxOfObb(const ObbS& o) { BoxS b; b.mn = o.c - o.r; b.mx = o.c + o.r; return b; }
B200_D bool sphereVsObbB(const SphereS& s, const ObbS& o) { SphereS s_; s_.c = conjugate(o.q) * (s.c - o.c) + o.c; s_.r = s.r; return sphereVsBoxB(s_, boxOfObb(o)); }
B200_D bool capsuleVsCapsuleB(const CapS& a, const CapS& b)
{
    f3 c1, c2; closestSegmentSegment(a.a, a.b, b.a, b.b, c1, c2);
    SphereS x, y; x.c = c1; x.r = a.r; y.c = c2; y.r = b.r; return sphereVsSphereB(x, y);
}
B200_D bool capsuleVsCylinderB(const CapS& a, const CapS& b)
{
    f3 c1, c2; closestSegmentSegment(a.a, a.b, b.a, b.b, c1, c2);
    SphereS x; x.c = c1; x.r = a.r; return sphereVsCylinderB(x, b);
}
B200_D bool gjkTest(const SupportD& A, const SupportD& B, NarrowCounters& nc) { SimplexD s; return gjk(A, B, s, nc); }   // gjkIntersectionTest
B200_D CapS intoObbFrame(const CapS& c, const ObbS& o) { CapS r; r.a = conjugate(o.q) * (c.a - o.c) + o.c; r.b = conjugate(o.q) * (c.b - o.c) + o.c; r.r = c.r; return r; }
B200_D bool boxVsBoxB(const BoxS& a, const BoxS& b)
{
    if (a.mx.x < b.mn.x || a.mn.x > b.mx.x) return false;
    if (a.mx.y < b.mn.y || a.mn.y > b.mx.y) return false;
    if (a.mx.z < b.mn.z || a.mn.z > b.mx.z) return false;
    return true;
}
// bounding_volumes.cpp:1079-1199 (15-axis SAT, early out; no contact data)
B200_D bool obbVsObbB(const ObbS& a, const ObbS& b)
{
    const f3 ax[3] = { a.q * mk3(1.f, 0.f, 0.f), a.q * mk3(0.f, 1.f, 0.f), a.q * mk3(0.f, 0.f, 1.f) };
    const f3 bx[3] = { b.q * mk3(1.f, 0.f, 0.f), b.q * mk3(0.f, 1.f, 0.f), b.q * mk3(0.f, 0.f, 1.f) };
    float r[3][3], absR[3][3];                     // r[i][j] = m_ij = dot(axesA[i], axesB[j])
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { r[i][j] = dot(ax[i], bx[j]); absR[i][j] = fabsf(r[i][j]) + B200_EPS; }
    const f3 t = conjugate(a.q) * (b.c - a.c);
    const float ta[3] = { t.x, t.y, t.z }, ar[3] = { a.r.x, a.r.y, a.r.z }, br[3] = { b.r.x, b.r.y, b.r.z };
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {   // a's faces: row i of absR
        const float ra = ar[i], rb = dot(mk3(absR[i][0], absR[i][1], absR[i][2]), b.r);
        if (ra + rb - fabsf(ta[i]) < 0.f) return false;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
    {   // b's faces: column i
        const float ra = dot(mk3(absR[0][i], absR[1][i], absR[2][i]), a.r), rb = br[i];
        const float d = dot(mk3(r[0][i], r[1][i], r[2][i]), t);
        if (ra + rb - fabsf(d) < 0.f) return false;
    }
    float ra, rb;
    ra = a.r.y * absR[2][0] + a.r.z * absR[1][0]; rb = b.r.y * absR[0][2] + b.r.z * absR[0][1]; if (ra + rb - fabsf(t.z * r[1][0] - t.y * r[2][0]) < 0.f) return false;
    ra = a.r.y * absR[2][1] + a.r.z * absR[1][1]; rb = b.r.x * absR[0][2] + b.r.z * absR[0][0]; if (ra + rb - fabsf(t.z * r[1][1] - t.y * r[2][1]) < 0.f) return false;
    ra = a.r.y * absR[2][2] + a.r.z * absR[1][2]; rb = b.r.x * absR[0][1] + b.r.y * absR[0][0]; if (ra + rb - fabsf(t.z * r[1][2] - t.y * r[2][2]) < 0.f) return false;
    ra = a.r.x * absR[2][0] + a.r.z * absR[0][0]; rb = b.r.y * absR[1][2] + b.r.z * absR[1][1]; if (ra + rb - fabsf(t.x * r[2][0] - t.z * r[0][0]) < 0.f) return false;
    ra = a.r.x * absR[2][1] + a.r.z * absR[0][1]; rb = b.r.x * absR[1][2] + b.r.z * absR[1][0]; if (ra + rb - fabsf(t.x * r[2][1] - t.z * r[0][1]) < 0.f) return false;
    ra = a.r.x * absR[2][2] + a.r.z * absR[0][2]; rb = b.r.x * absR[1][1] + b.r.y * absR[1][0]; if (ra + rb - fabsf(t.x * r[2][2] - t.z * r[0][2]) < 0.f) return false;
    ra = a.r.x * absR[1][0] + a.r.y * absR[0][0]; rb = b.r.y * absR[2][2] + b.r.z * absR[2][1]; if (ra + rb - fabsf(t.y * r[0][0] - t.x * r[1][0]) < 0.f) return false;
    ra = a.r.x * absR[1][1] + a.r.y * absR[0][1]; rb = b.r.x * absR[2][2] + b.r.z * absR[2][0]; if (ra + rb - fabsf(t.y * r[0][1] - t.x * r[1][1]) < 0.f) return false;
    ra = a.r.x * absR[1][2] + a.r.y * absR[0][2]; rb = b.r.x * absR[2][1] + b.r.y * absR[2][0]; if (ra + rb - fabsf(t.y * r[0][2] - t.x * r[1][2]) < 0.f) return false;
    return true;
}

} // namespace b200
