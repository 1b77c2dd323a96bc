// oracle/om_narrow.cpp -- TEST INFRASTRUCTURE ONLY (parity oracle). Never linked into the product path.
//
// CPU restatement of the reference's SCALAR narrow phase:
//   src/physics/collision_narrow.cpp:56-369 (manifold reduction, clipping, box helpers), :374-1584 (21 shape pairs)
//   src/physics/collision_gjk.h / collision_gjk.cpp (support mappings, GJK boolean test + simplex update)
//   src/physics/collision_epa.h / collision_epa.cpp (expanding polytope)
// Deviations that are deliberate and counted in NarrowStats instead of trapping:
//   * the reference GJK loop has no iteration cap (collision_gjk.h:210-239); the oracle stops after 64 rounds with
//     "no collision" and counts it, the CUDA path does the same
//   * ASSERTs become counters
#include "om_shapes.h"

namespace orc {

// ----------------------------------------------------------------------------------------------- support mappings
struct Support
{
    int type;
    Sphere s; Capsule c; Box b; Obb o; Hull h;

    V3 operator()(V3 dir) const
    {
        switch (type)
        {
        case T_SPHERE:  // collision_gjk.h:6-14
            return normalize(dir) * s.r + s.c;
        case T_CAPSULE: // collision_gjk.h:16-27
        {
            float dA = dot(dir, c.a), dB = dot(dir, c.b);
            V3 far = dA > dB ? c.a : c.b;
            return normalize(dir) * c.r + far;
        }
        case T_CYLINDER: // collision_gjk.h:29-44
        {
            float dA = dot(dir, c.a), dB = dot(dir, c.b);
            V3 far = dA > dB ? c.a : c.b;
            V3 n = c.a - c.b;
            V3 proj = noz(cross(cross(n, dir), n));
            return far + proj * c.r;
        }
        case T_AABB:    // collision_gjk.h:46-58
            return v3((dir.x < 0.f) ? b.mn.x : b.mx.x, (dir.y < 0.f) ? b.mn.y : b.mx.y, (dir.z < 0.f) ? b.mn.z : b.mx.z);
        case T_OBB:     // collision_gjk.h:60-75
        {
            V3 d = conjugate(o.q) * dir;
            V3 r = v3(d.x < 0.f ? -o.r.x : o.r.x, d.y < 0.f ? -o.r.y : o.r.y, d.z < 0.f ? -o.r.z : o.r.z);
            return o.c + o.q * r;
        }
        default:        // hull, collision_gjk.h:77-100 : first maximum in vertex order
        {
            V3 d = conjugate(h.q) * dir;
            V3 best = v3(0.f);
            float maxDist = -FLT_MAX;
            for (const V3& v : h.g->verts)
            {
                float t = dot(d, v);
                if (t > maxDist) { maxDist = t; best = v; }
            }
            return h.p + h.q * best;
        }
        }
    }
};

static Support supportOf(const WCol& w)
{
    Support s; s.type = w.type;
    s.s = w.sphere; s.c = w.cap; s.b = w.box; s.o = w.obb; s.h = w.hull;
    return s;
}
static Support supSphere(const Sphere& x) { Support s; s.type = T_SPHERE; s.s = x; return s; }
static Support supCapsule(const Capsule& x) { Support s; s.type = T_CAPSULE; s.c = x; return s; }
static Support supCylinder(const Capsule& x) { Support s; s.type = T_CYLINDER; s.c = x; return s; }
static Support supBox(const Box& x) { Support s; s.type = T_AABB; s.b = x; return s; }
static Support supObb(const Obb& x) { Support s; s.type = T_OBB; s.o = x; return s; }
static Support supHull(const Hull& x) { Support s; s.type = T_HULL; s.h = x; return s; }

// ------------------------------------------------------------------------------------------------------------ GJK
struct SupPoint { V3 a, b, m; };           // gjk_support_point (collision_gjk.h:137-154)
struct Simplex { SupPoint a, b, c, d; uint32_t n; };

static inline SupPoint minkowski(const Support& A, const Support& B, V3 dir)
{
    SupPoint p; p.a = A(dir); p.b = B(-dir); p.m = p.a - p.b; return p; // collision_gjk.h:163-169
}
static inline V3 crossABA(V3 a, V3 b) { return cross(cross(a, b), a); }

enum { GJK_STOP, GJK_CONTINUE, GJK_ERROR };

// collision_gjk.cpp:6-212. The reference jumps between face cases with goto; here each edge test is a small lambda
// evaluated in the same order.
static int updateSimplex(Simplex& s, const SupPoint& a, V3& dir)
{
    if (s.n == 2)
    {
        V3 ao = -a.m, ab = s.b.m - a.m, ac = s.c.m - a.m;
        V3 abc = cross(ab, ac);
        V3 abp = cross(ab, abc);
        if (dot(ao, abp) > 0.f) { s.c = a; dir = crossABA(ab, ao); return GJK_CONTINUE; }
        V3 acp = cross(abc, ac);
        if (dot(ao, acp) > 0.f) { s.b = a; dir = crossABA(ac, ao); return GJK_CONTINUE; }
        if (dot(ao, abc) >= 0.f) { s.d = s.b; s.b = a; s.n = 3; dir = abc; return GJK_CONTINUE; }
        if (dot(ao, -abc) >= 0.f) { s.d = s.c; s.c = s.b; s.b = a; s.n = 3; dir = -abc; return GJK_CONTINUE; }
        return GJK_ERROR;
    }
    if (s.n == 3)
    {
        V3 ao = -a.m, ab = s.b.m - a.m, ac = s.c.m - a.m, ad = s.d.m - a.m;
        V3 bcd = cross(s.c.m - s.b.m, s.d.m - s.b.m);
        if (dot(bcd, dir) > 0.00001f || dot(bcd, s.b.m) < -0.00001f) { return GJK_ERROR; }

        V3 abc = cross(ac, ab), abd = cross(ab, ad), adc = cross(ad, ac);
        const int F_ABC = 1, F_ABD = 2, F_ADC = 4;
        int flags = 0;
        flags |= (dot(abc, ao) > 0.f) ? F_ABC : 0;
        flags |= (dot(abd, ao) > 0.f) ? F_ABD : 0;
        flags |= (dot(adc, ao) > 0.f) ? F_ADC : 0;
        if (flags == (F_ABC | F_ABD | F_ADC)) { return GJK_ERROR; }
        if (flags == 0) { return GJK_STOP; }

        // the six "new line" exits and three "stay triangle" exits of the reference
        auto lineAB = [&]() -> int { s.c = a; s.n = 2; dir = crossABA(ab, ao); return GJK_CONTINUE; };
        auto lineAC = [&]() -> int { s.b = a; s.n = 2; dir = crossABA(ac, ao); return GJK_CONTINUE; };
        auto lineAD_fromABD = [&]() -> int { s.b = s.d; s.c = a; s.n = 2; dir = crossABA(ad, ao); return GJK_CONTINUE; };
        auto lineAD_fromADC = [&]() -> int { s.b = a; s.c = s.d; s.n = 2; dir = crossABA(ad, ao); return GJK_CONTINUE; };
        // entry points named like the reference labels
        auto overABC2 = [&]() -> int { if (dot(cross(ac, abc), ao) > 0.f) return lineAC(); s.d = a; dir = abc; return (int)GJK_CONTINUE; };
        auto overABC1 = [&]() -> int { if (dot(cross(abc, ab), ao) > 0.f) return lineAB(); return overABC2(); };
        auto overABD2 = [&]() -> int { if (dot(cross(ab, abd), ao) > 0.f) return lineAB(); s.c = a; dir = abd; return (int)GJK_CONTINUE; };
        auto overABD1 = [&]() -> int { if (dot(cross(abd, ad), ao) > 0.f) return lineAD_fromABD(); return overABD2(); };
        auto overADC2 = [&]() -> int { if (dot(cross(ad, adc), ao) > 0.f) return lineAD_fromADC(); s.b = a; dir = adc; return (int)GJK_CONTINUE; };
        auto overADC1 = [&]() -> int { if (dot(cross(adc, ac), ao) > 0.f) return lineAC(); return overADC2(); };

        if (flags == F_ABC) return overABC1();
        if (flags == F_ABD) return overABD1();
        if (flags == F_ADC) return overADC1();
        if (flags == (F_ABC | F_ABD)) { if (dot(cross(abc, ab), ao) > 0.f) return overABD1(); return overABC2(); }
        if (flags == (F_ABD | F_ADC)) { if (dot(cross(abd, ad), ao) > 0.f) return overADC1(); return overABD2(); }
        if (flags == (F_ADC | F_ABC)) { if (dot(cross(adc, ac), ao) > 0.f) return overABC1(); return overADC2(); }
        return GJK_ERROR;
    }
    return GJK_ERROR;
}

static const int kGjkMaxRounds = 64;

// collision_gjk.h:184-238
static bool gjk(const Support& A, const Support& B, Simplex& s, NarrowStats& st)
{
    V3 dir = v3(1.f, 0.1f, -0.2f);
    s.c = minkowski(A, B, dir);
    if (dot(s.c.m, dir) < 0.f) return false;
    dir = -s.c.m;
    s.b = minkowski(A, B, dir);
    if (dot(s.b.m, dir) < 0.f) return false;
    dir = crossABA(s.c.m - s.b.m, -s.b.m);
    s.n = 2;
    for (int round = 0; ; ++round)
    {
        if (round >= kGjkMaxRounds) { ++st.gjkIterCap; return false; }
        if (sqLen(dir) < 0.0001f) return false;
        SupPoint a = minkowski(A, B, dir);
        if (dot(a.m, dir) < 0.f) return false;
        int r = updateSimplex(s, a, dir);
        if (r == GJK_STOP) { s.a = a; s.n = 4; break; }
        if (r == GJK_ERROR) { ++st.gjkErrors; return false; }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------------------ EPA
struct EpaResult { V3 point, normal; float depth; };

struct Polytope   // epa_simplex (collision_epa.h:45-67), same capacities
{
    enum { CAP = 1024, BORDER_CAP = 128 };
    struct Tri { uint16_t a, b, c, eA, eB, eC; V3 n; float dist; };
    struct Edge { uint16_t a, b, tri[2]; };
    SupPoint pts[CAP]; Tri tris[CAP]; Edge edges[CAP];
    bool active[CAP];
    uint32_t nT, nP, nE;

    static void triInfo(const SupPoint& a, const SupPoint& b, const SupPoint& c, V3& n, float& d)
    {   // collision_epa.cpp:5-11
        n = normalize(cross(b.m - a.m, c.m - a.m));
        d = dot(n, a.m);
    }
    int pushPoint(const SupPoint& p) { if (nP >= CAP) return -1; pts[nP] = p; return (int)nP++; }
    int pushTri(uint16_t a, uint16_t b, uint16_t c, uint16_t eA, uint16_t eB, uint16_t eC, V3 n, float d)
    {
        if (nT >= CAP) return -1;
        Tri& t = tris[nT]; t.a = a; t.b = b; t.c = c; t.eA = eA; t.eB = eB; t.eC = eC; t.n = n; t.dist = d;
        active[nT] = true; return (int)nT++;
    }
    int pushEdge(uint16_t a, uint16_t b, uint16_t tA, uint16_t tB)
    {
        if (nE >= CAP) return -1;
        Edge& e = edges[nE]; e.a = a; e.b = b; e.tri[0] = tA; e.tri[1] = tB; return (int)nE++;
    }
    uint32_t closest() const
    {   // collision_epa.cpp:94-112 : first strict minimum among active triangles
        uint32_t best = 0xFFFFFFFFu; float mind = FLT_MAX;
        for (uint32_t i = 0; i < nT; ++i) if (active[i] && tris[i].dist < mind) { mind = tris[i].dist; best = i; }
        return best;
    }
    // collision_epa.cpp:114-239
    bool addPoint(const SupPoint& np)
    {
        // Degenerate horizons (two border edges ending in the same point) leave 0xFFFF indices behind; the reference then
        // reads out of bounds (undefined behaviour). Defined here and on the device: an invalid index is "inactive" / skipped.
        static thread_local uint8_t refs[CAP];
        auto isActive = [&](uint32_t i) { return i < nT && active[i]; };
        auto refEdge = [&](uint32_t e) { if (e < nE) ++refs[e]; };
        for (uint32_t i = 0; i < nE; ++i) refs[i] = 0;
        for (uint32_t i = 0; i < nT; ++i)
        {
            if (!active[i]) continue;
            Tri& t = tris[i];
            float d = dot(t.n, np.m - pts[t.a].m);
            if (d > 0.f) { refEdge(t.eA); refEdge(t.eB); refEdge(t.eC); active[i] = false; }
        }
        uint16_t border[BORDER_CAP]; uint32_t nBorder = 0;
        for (uint32_t i = 0; i < nE; ++i)
        {
            if (refs[i] == 1) { if (nBorder >= BORDER_CAP) return false; border[nBorder++] = (uint16_t)i; }
        }
        static thread_local uint16_t newEdgePerPoint[CAP];
        for (uint32_t i = 0; i <= nP && i < CAP; ++i) newEdgePerPoint[i] = 0xFFFF;
        int npi = pushPoint(np);
        if (npi < 0) return false;
        uint16_t triOffset = (uint16_t)nT;
        for (uint32_t i = 0; i < nBorder; ++i)
        {
            uint16_t ei = border[i];
            Edge& e = edges[ei];
            bool aAct = isActive(e.tri[0]), bAct = isActive(e.tri[1]);
            uint16_t connect = bAct ? e.a : e.b;
            uint16_t triIndex = (uint16_t)nT;
            int ne = pushEdge(connect, (uint16_t)npi, 0xFFFF, (uint16_t)nT);
            if (ne < 0) return false;
            newEdgePerPoint[connect] = (uint16_t)ne;
            uint16_t bI = connect, cI = bAct ? e.b : e.a;
            V3 n; float d; triInfo(np, pts[bI], pts[cI], n, d);
            int ti = pushTri((uint16_t)npi, bI, cI, ei, 0xFFFF, (uint16_t)ne, n, d);
            if (ti < 0) return false;
            Edge& e2 = edges[ei]; // (array is static storage; reference stays valid, re-fetched for clarity)
            if (aAct) e2.tri[1] = triIndex; else e2.tri[0] = triIndex;
        }
        for (uint32_t i = 0; i < nBorder; ++i)
        {
            Edge& e = edges[border[i]];
            bool bNew = e.tri[1] >= triOffset;
            uint16_t connect = bNew ? e.a : e.b;
            uint16_t other = newEdgePerPoint[connect];
            uint16_t triIndex = (uint16_t)(i + triOffset);
            if (other == 0xFFFF) continue;
            tris[triIndex].eB = other;
            edges[other].tri[0] = triIndex;
        }
        return true;
    }
};

// collision_epa.h:96-168. Result is written even when not converged (caller ignores the status, collision_narrow.cpp:508-512).
static void epa(const Simplex& g, const Support& A, const Support& B, EpaResult& out, NarrowStats& st)
{
    static thread_local Polytope P;
    P.nT = P.nP = P.nE = 0;
    P.pushPoint(g.a); P.pushPoint(g.b); P.pushPoint(g.c); P.pushPoint(g.d);
    V3 n; float d;
    Polytope::triInfo(g.a, g.b, g.d, n, d); P.pushTri(0, 1, 3, 4, 3, 0, n, d);
    Polytope::triInfo(g.b, g.c, g.d, n, d); P.pushTri(1, 2, 3, 5, 4, 1, n, d);
    Polytope::triInfo(g.c, g.a, g.d, n, d); P.pushTri(2, 0, 3, 3, 5, 2, n, d);
    Polytope::triInfo(g.a, g.c, g.b, n, d); P.pushTri(0, 2, 1, 1, 0, 2, n, d);
    P.pushEdge(0, 1, 0, 3); P.pushEdge(1, 2, 1, 3); P.pushEdge(2, 0, 2, 3);
    P.pushEdge(0, 3, 2, 0); P.pushEdge(1, 3, 0, 1); P.pushEdge(2, 3, 1, 2);

    uint32_t ci = 0;
    int status = 2; // 0 success, 1 out of memory, 2 max iterations
    for (uint32_t it = 0; it < 20; ++it)
    {
        ci = P.closest();
        if (ci == 0xFFFFFFFFu) { ++st.asserts; ci = 0; break; }
        const Polytope::Tri& t = P.tris[ci];
        SupPoint a = minkowski(A, B, t.n);
        float dd = dot(a.m, t.n);
        if (dd - t.dist < 0.01f) { status = 0; break; }
        if (!P.addPoint(a)) { status = 1; break; }
    }
    if (status == 1) ++st.epaOutOfMemory;
    if (status == 2) ++st.epaNotConverged;
    if (P.nT > st.epaMaxTriangles) st.epaMaxTriangles = P.nT;
    if (P.nE > st.epaMaxEdges) st.epaMaxEdges = P.nE;
    if (P.nP > st.epaMaxPoints) st.epaMaxPoints = P.nP;

    const Polytope::Tri& t = P.tris[ci];
    const SupPoint& a = P.pts[t.a]; const SupPoint& b = P.pts[t.b]; const SupPoint& c = P.pts[t.c];
    V3 bc = barycentric(a.m, b.m, c.m, t.n * t.dist);
    V3 pA = bc.x * a.a + bc.y * b.a + bc.z * c.a;
    V3 pB = bc.x * a.b + bc.y * b.b + bc.z * c.b;
    out.point = 0.5f * (pA + pB);
    out.normal = t.n;
    out.depth = t.dist;
}

static bool gjkEpa(const Support& A, const Support& B, Manifold& m, NarrowStats& st)
{
    Simplex s;
    if (!gjk(A, B, s, st)) return false;
    EpaResult r; epa(s, A, B, r, st);
    m.normal = r.normal; m.n = 1; m.depth[0] = r.depth; m.point[0] = r.point;
    return true;
}

// --------------------------------------------------------------------------------- manifold reduction and clipping
struct VP { V3 v; float depth; };              // vertex_penetration_pair
struct Poly { VP p[16]; uint32_t n; };         // clipping_polygon

// collision_narrow.cpp:56-146
static void stableManifold(VP* v, uint32_t n, V3 normal, Manifold& out)
{
    if (n > 4)
    {
        V3 search = getTangent(normal);
        float best = dot(search, v[0].v); uint32_t ri = 0;
        for (uint32_t i = 1; i < n; ++i) { float d = dot(search, v[i].v); if (d > best) { ri = i; best = d; } }
        out.depth[0] = v[ri].depth; out.point[0] = v[ri].v;

        best = 0.f; ri = 0;
        for (uint32_t i = 0; i < n; ++i) { float d = sqLen(v[i].v - out.point[0]); if (d > best) { ri = i; best = d; } }
        out.depth[1] = v[ri].depth; out.point[1] = v[ri].v;

        best = 0.f; ri = 0;
        for (uint32_t i = 0; i < n; ++i)
        {
            V3 qa = out.point[0] - v[i].v, qb = out.point[1] - v[i].v;
            float area = 0.5f * dot(cross(qa, qb), normal);
            if (area > best) { ri = i; best = area; }
        }
        out.depth[2] = v[ri].depth; out.point[2] = v[ri].v;

        best = 0.f; ri = 0;
        for (uint32_t i = 0; i < n; ++i)
        {
            V3 qa = out.point[0] - v[i].v, qb = out.point[1] - v[i].v, qc = out.point[2] - v[i].v;
            float a1 = 0.5f * dot(cross(qa, qb), normal);
            float a2 = 0.5f * dot(cross(qb, qc), normal);
            float a3 = 0.5f * dot(cross(qc, qa), normal);
            float area = fmax2(fmax2(a1, a2), a3);
            if (area > best) { ri = i; best = area; }
        }
        out.depth[3] = v[ri].depth; out.point[3] = v[ri].v;
        out.n = 4;
    }
    else
    {
        out.n = n;
        for (uint32_t i = 0; i < n; ++i) { out.depth[i] = v[i].depth; out.point[i] = v[i].v; }
    }
}

// collision_narrow.cpp:154-163
static VP clipEdge(VP a, VP b, float aDist, float bDist)
{
    aDist = fabsf(aDist); bDist = fabsf(bDist);
    float total = aDist + bDist;
    float t = aDist / total;
    VP r = { lerp(a.v, b.v, t), lerpf(a.depth, b.depth, t) };
    return r;
}

// collision_narrow.cpp:166-222 (Sutherland-Hodgman, planes point inside). Result always lands in `output`.
static void clipPolygon(Poly& input, const Plane* planes, uint32_t nPlanes, Poly& output)
{
    Poly* in = &input; Poly* out = &output;
    uint32_t ci = 0;
    for (; ci < nPlanes; ++ci)
    {
        Plane pl = planes[ci];
        out->n = 0;
        if (in->n == 0) break;
        VP start = in->p[in->n - 1];
        for (uint32_t i = 0; i < in->n; ++i)
        {
            VP end = in->p[i];
            float sd = signedDistance(start.v, pl), ed = signedDistance(end.v, pl);
            bool sIn = sd > 0.f, eIn = ed > 0.f;
            if (sIn && eIn) out->p[out->n++] = end;
            else if (sIn) out->p[out->n++] = clipEdge(start, end, sd, ed);
            else if (!sIn && eIn) { out->p[out->n++] = clipEdge(start, end, sd, ed); out->p[out->n++] = end; }
            start = end;
        }
        Poly* tmp = in; in = out; out = tmp;
    }
    if (ci % 2 == 0) { for (uint32_t i = 0; i < input.n; ++i) output.p[i] = input.p[i]; output.n = input.n; }
}

static inline int maxAbsAxis(V3 n) { V3 p = vabs(n); return (p.x > p.y) ? ((p.x > p.z) ? 0 : 2) : ((p.y > p.z) ? 1 : 2); }

// collision_narrow.cpp:225-254 (box local space)
static void boxClipPlanes(V3 radius, V3 normal, V3* pts, V3* nrm)
{
    int me = maxAbsAxis(normal), a0 = (me + 1) % 3, a1 = (me + 2) % 3;
    V3 n;
    n = v3(0.f); setComp(n, a0, 1.f);  nrm[0] = n; pts[0] = -radius;
    n = v3(0.f); setComp(n, a1, 1.f);  nrm[1] = n; pts[1] = -radius;
    n = v3(0.f); setComp(n, a0, -1.f); nrm[2] = n; pts[2] = radius;
    n = v3(0.f); setComp(n, a1, -1.f); nrm[3] = n; pts[3] = radius;
}
// collision_narrow.cpp:257-289
static void boxIncidentVerts(V3 radius, V3 normal, Poly& poly)
{
    int me = maxAbsAxis(normal), a0 = (me + 1) % 3, a1 = (me + 2) % 3;
    float s = comp(normal, me) < 0.f ? 1.f : -1.f;
    float d = comp(radius, me) * s;
    float min0 = -comp(radius, a0), min1 = -comp(radius, a1), max0 = comp(radius, a0), max1 = comp(radius, a1);
    poly.n = 4;
    float c0[4] = { min0, max0, max0, min0 }, c1[4] = { min1, min1, max1, max1 };
    for (int i = 0; i < 4; ++i) { setComp(poly.p[i].v, me, d); setComp(poly.p[i].v, a0, c0[i]); setComp(poly.p[i].v, a1, c1[i]); }
}
// collision_narrow.cpp:291-299
static Plane boxReferencePlane(const Box& b, V3 normal)
{
    V3 pt = v3((normal.x < 0.f) ? b.mn.x : b.mx.x, (normal.y < 0.f) ? b.mn.y : b.mx.y, (normal.z < 0.f) ? b.mn.z : b.mx.z);
    return createPlane(pt, normal);
}
// collision_narrow.cpp:301-336
static void boxIncidentEdge(V3 r, V3 normal, V3& oa, V3& ob)
{
    V3 p = vabs(normal);
    oa = r;
    if (p.x > p.y) { if (p.y > p.z) ob = v3(r.x, r.y, -r.z); else ob = v3(r.x, -r.y, r.z); }
    else { if (p.x > p.z) ob = v3(r.x, r.y, -r.z); else ob = v3(-r.x, r.y, r.z); }
    V3 s = v3(normal.x < 0.f ? -1.f : 1.f, normal.y < 0.f ? -1.f : 1.f, normal.z < 0.f ? -1.f : 1.f);
    oa = oa * s; ob = ob * s;
}
// collision_narrow.cpp:339-369
static bool clipAndBuild(Poly& poly, const Plane* planes, uint32_t nPlanes, const Plane& ref, Manifold& out)
{
    Poly cl; cl.n = 0;
    clipPolygon(poly, planes, nPlanes, cl);
    if (cl.n > 0)
    {
        for (uint32_t i = 0; i < cl.n; ++i)
        {
            if (cl.p[i].depth < 0.f) { cl.p[i] = cl.p[cl.n - 1]; --cl.n; --i; }
            else { cl.p[i].v += ref.n * cl.p[i].depth; }
        }
        if (cl.n > 0) { stableManifold(cl.p, cl.n, out.normal, out); return true; }
    }
    return false;
}

// ---------------------------------------------------------------------------------------------------- shape pairs
// collision_narrow.cpp:374-400
static bool sphereSphere(const Sphere& s1, const Sphere& s2, Manifold& m, NarrowStats& st)
{
    V3 n = s2.c - s1.c;
    float rs = s2.r + s1.r;
    float sq = sqLen(n);
    if (sq <= rs * rs)
    {
        float dist;
        if (sq == 0.f) { dist = 0.f; m.normal = v3(0.f, 1.f, 0.f); }
        else { dist = sqrtf(sq); m.normal = n / dist; }
        m.n = 1;
        m.depth[0] = rs - dist;
        if (!(m.depth[0] >= 0.f)) ++st.asserts;
        m.point[0] = 0.5f * (s1.c + s1.r * m.normal + s2.c - s2.r * m.normal);
        return true;
    }
    return false;
}
// collision_narrow.cpp:402-406
static bool sphereCapsule(const Sphere& s, const Capsule& c, Manifold& m, NarrowStats& st)
{
    Sphere o = { closestPointSegment(s.c, c.a, c.b), c.r };
    return sphereSphere(s, o, m, st);
}
// collision_narrow.cpp:408-449
static bool sphereCylinder(const Sphere& s, const Capsule& c, Manifold& m, NarrowStats& st)
{
    V3 ab = c.b - c.a;
    float t = dot(s.c - c.a, ab) / sqLen(ab);
    if (t >= 0.f && t <= 1.f) { Sphere o = { lerp(c.a, c.b, t), c.r }; return sphereSphere(s, o, m, st); }
    V3 p = (t <= 0.f) ? c.a : c.b;
    V3 up = (t <= 0.f) ? -ab : ab;
    V3 proj = normalize(cross(cross(up, s.c - p), up));
    V3 endA = p + proj * c.r, endB = p - proj * c.r;
    V3 closest = closestPointSegment(s.c, endA, endB);
    V3 normal = closest - s.c;
    float sq = sqLen(normal);
    if (sq <= s.r * s.r)
    {
        float dist;
        if (sq == 0.f) { dist = 0.f; m.normal = -normalize(up); }
        else { dist = sqrtf(sq); m.normal = normal / dist; }
        m.n = 1;
        m.depth[0] = s.r - dist;
        if (!(m.depth[0] >= 0.f)) ++st.asserts;
        m.point[0] = closest + 0.5f * m.depth[0] * normal;
        return true;
    }
    return false;
}
// collision_narrow.cpp:451-478
static bool sphereBox(const Sphere& s, const Box& a, Manifold& m)
{
    V3 p = closestPointAABB(s.c, a);
    V3 n = p - s.c;
    float sq = sqLen(n);
    if (sq <= s.r * s.r)
    {
        float dist = 0.f;
        if (sq > 0.f) { dist = sqrtf(sq); n /= dist; } else { n = v3(0.f, 1.f, 0.f); }
        m.n = 1; m.normal = n;
        m.depth[0] = s.r - dist;
        m.point[0] = 0.5f * (p + s.c + n * s.r);
        return true;
    }
    return false;
}
// collision_narrow.cpp:480-494
static bool sphereObb(const Sphere& s, const Obb& o, Manifold& m)
{
    Box bb = boxFromCenterRadius(o.c, o.r);
    Sphere s_ = { conjugate(o.q) * (s.c - o.c) + o.c, s.r };
    if (sphereBox(s_, bb, m))
    {
        m.normal = o.q * m.normal;
        m.point[0] = o.q * (m.point[0] - o.c) + o.c;
        return true;
    }
    return false;
}

// shared skeleton of capsule-capsule (:523-612), capsule-cylinder (:614-703): `bIsCylinder` selects the end cases
static bool capsuleVsSegmentShape(const Capsule& a, const Capsule& b, bool bIsCylinder, Manifold& m, NarrowStats& st)
{
    V3 aDir = a.b - a.a;
    V3 bDir = normalize(b.b - b.a);
    float aLen = len(aDir);
    aDir *= 1.f / aLen;
    float parallel = dot(aDir, bDir);
    if (fabsf(parallel) > 0.99f)
    {
        V3 pAa = a.a, pAb = a.b, pBa = b.a, pBb = b.b;
        if (parallel < 0.f) { V3 t = pBa; pBa = pBb; pBb = t; }
        V3 ref = a.a;
        float a0 = 0.f, a1 = aLen;
        float b0 = dot(aDir, pBa - ref), b1 = dot(aDir, pBb - ref);
        if (!(b1 > b0)) ++st.asserts;
        float left = fmax2(a0, b0), right = fmin2(a1, b1);
        if (right < left)
        {
            if (a0 > b1)
            {
                Sphere sa = { pAa, a.r };
                if (bIsCylinder) return sphereCylinder(sa, b, m, st);
                Sphere sb = { pBb, b.r }; return sphereSphere(sa, sb, m, st);
            }
            else
            {
                Sphere sa = { pAb, a.r };
                if (bIsCylinder) return sphereCylinder(sa, b, m, st);
                Sphere sb = { pBa, b.r }; return sphereSphere(sa, sb, m, st);
            }
        }
        V3 cA0 = ref + left * aDir, cA1 = ref + right * aDir;
        V3 cB0 = closestPointSegment(cA0, pBa, pBb);
        V3 cB1 = cB0 + (right - left) * aDir;
        V3 normal = cB0 - cA0;
        float d = len(normal);
        if (d < kEps) { d = 0.f; normal = v3(0.f, 1.f, 0.f); } else { normal /= d; }
        float rs = a.r + b.r;
        float pen = rs - d;
        if (pen < 0.f) return false;
        m.normal = normal; m.n = 2;
        m.depth[0] = pen; m.point[0] = (cA0 + cB0) * 0.5f;
        m.depth[1] = pen; m.point[1] = (cA1 + cB1) * 0.5f;
        return true;
    }
    else
    {
        V3 c1, c2;
        closestPointSegmentSegment(a.a, a.b, b.a, b.b, c1, c2);
        Sphere sa = { c1, a.r };
        if (bIsCylinder) return sphereCylinder(sa, b, m, st);
        Sphere sb = { c2, b.r }; return sphereSphere(sa, sb, m, st);
    }
}

// capsule-aabb (:705-769) and cylinder-aabb (:953-1022) share everything but the support mapping
static bool segmentShapeBox(const Capsule& c, bool isCylinder, const Box& a, Manifold& m, NarrowStats& st)
{
    Support A = isCylinder ? supCylinder(c) : supCapsule(c);
    Support B = supBox(a);
    if (!gjkEpa(A, B, m, st)) return false;
    V3 normal = m.normal;
    if (fabsf(normal.x) > 0.99f || fabsf(normal.y) > 0.99f || fabsf(normal.z) > 0.99f)
    {
        V3 axis = normalize(c.b - c.a);
        if (fabsf(dot(normal, axis)) < 0.01f)
        {
            V3 cpp[4], cpn[4]; Plane planes[4];
            V3 boxNormal = -normal;
            Plane ref = boxReferencePlane(a, boxNormal);
            Poly poly; poly.n = 2;
            V3 pa = c.a + normal * c.r, pb = c.b + normal * c.r;
            poly.p[0].v = pa; poly.p[0].depth = -signedDistance(pa, ref);
            poly.p[1].v = pb; poly.p[1].depth = -signedDistance(pb, ref);
            V3 ac = boxCenter(a);
            boxClipPlanes(boxRadius(a), boxNormal, cpp, cpn);
            for (int i = 0; i < 4; ++i) { cpp[i] = cpp[i] + ac; planes[i] = createPlane(cpp[i], cpn[i]); }
            clipAndBuild(poly, planes, 4, ref, m);
        }
    }
    return true;
}
// capsule-obb (:771-790), cylinder-obb (:1024-1043)
static bool segmentShapeObb(const Capsule& c, bool isCylinder, const Obb& o, Manifold& m, NarrowStats& st)
{
    Box bb = boxFromCenterRadius(o.c, o.r);
    Capsule c_ = { conjugate(o.q) * (c.a - o.c) + o.c, conjugate(o.q) * (c.b - o.c) + o.c, c.r };
    if (segmentShapeBox(c_, isCylinder, bb, m, st))
    {
        m.normal = o.q * m.normal;
        for (uint32_t i = 0; i < m.n; ++i) m.point[i] = o.q * (m.point[i] - o.c) + o.c;
        return true;
    }
    return false;
}

// collision_narrow.cpp:821-951
static bool cylinderCylinder(const Capsule& a, const Capsule& b, Manifold& m, NarrowStats& st)
{
    V3 aDir = a.b - a.a;
    V3 bDir = normalize(b.b - b.a);
    float aLen = len(aDir);
    aDir *= 1.f / aLen;
    float parallel = dot(aDir, bDir);
    if (fabsf(parallel) > 0.99f)
    {
        V3 pBa = b.a, pBb = b.b;
        if (parallel < 0.f) { V3 t = pBa; pBa = pBb; pBb = t; }
        V3 ref = a.a;
        float a0 = 0.f, a1 = aLen;
        float b0 = dot(aDir, pBa - ref), b1 = dot(aDir, pBb - ref);
        if (!(b1 > b0)) ++st.asserts;
        float left = fmax2(a0, b0), right = fmin2(a1, b1);
        if (right < left) return false;
        V3 cA0 = ref + left * aDir, cA1 = ref + right * aDir;
        V3 cB0 = closestPointSegment(cA0, pBa, pBb);
        V3 cB1 = cB0 + (right - left) * aDir;
        V3 normal = cB0 - cA0;
        float d = len(normal);
        float rs = a.r + b.r;
        float pen = rs - d;
        if (pen < 0.f) return false;
        float capPen = right - left;
        if (!(capPen > 0.f)) ++st.asserts;
        if (capPen < pen)
        {
            m.n = 1; m.depth[0] = capPen;
            // quirk B2 (:891,897): scalar subtracted from / added to every component
            if (b0 > a0) { m.normal = aDir; m.point[0] = a.b - v3(capPen * 0.5f); }
            else { m.normal = -aDir; m.point[0] = a.a + v3(capPen * 0.5f); }
        }
        else
        {
            if (d < kEps) { d = 0.f; normal = v3(0.f, 1.f, 0.f); } else { normal /= d; }
            m.normal = normal; m.n = 2;
            m.depth[0] = pen; m.point[0] = (cA0 + cB0) * 0.5f;
            m.depth[1] = pen; m.point[1] = (cA1 + cB1) * 0.5f;
        }
        return true;
    }
    return gjkEpa(supCylinder(a), supCylinder(b), m, st);
}

// collision_narrow.cpp:1074-1140 (depth carries the sign of the separation direction: quirk B13)
static bool boxBox(const Box& a, const Box& b, Manifold& m)
{
    V3 cA = boxCenter(a), cB = boxCenter(b), rA = boxRadius(a), rB = boxRadius(b);
    V3 d = cB - cA;
    V3 p = (rB + rA) - vabs(d);
    if (p.x < 0.f || p.y < 0.f || p.z < 0.f) return false;
    int me = (p.x < p.y) ? ((p.x < p.z) ? 0 : 2) : ((p.y < p.z) ? 1 : 2);
    float s = comp(d, me) < 0.f ? -1.f : 1.f;
    float pen = comp(p, me) * s;
    V3 normal = v3(0.f); setComp(normal, me, s);
    m.normal = normal; m.n = 4;
    int a0 = (me + 1) % 3, a1 = (me + 2) % 3;
    float min0 = fmax2(comp(a.mn, a0), comp(b.mn, a0)), min1 = fmax2(comp(a.mn, a1), comp(b.mn, a1));
    float max0 = fmin2(comp(a.mx, a0), comp(b.mx, a0)), max1 = fmin2(comp(a.mx, a1), comp(b.mx, a1));
    float depth = comp(cA, me) + comp(rA, me) - pen * 0.5f;
    float c0[4] = { min0, min0, max0, max0 }, c1[4] = { min1, max1, min1, max1 };
    for (int i = 0; i < 4; ++i)
    {
        m.depth[i] = pen; m.point[i] = v3(0.f);
        setComp(m.point[i], a0, c0[i]); setComp(m.point[i], a1, c1[i]); setComp(m.point[i], me, depth);
    }
    return true;
}

// collision_narrow.cpp:1179-1527 (15-axis SAT in A's frame, face clip or edge-edge)
static bool obbObb(const Obb& a, const Obb& b, Manifold& m)
{
    V3 axA[3] = { a.q * v3(1.f, 0.f, 0.f), a.q * v3(0.f, 1.f, 0.f), a.q * v3(0.f, 0.f, 1.f) };
    V3 axB[3] = { b.q * v3(1.f, 0.f, 0.f), b.q * v3(0.f, 1.f, 0.f), b.q * v3(0.f, 0.f, 1.f) };
    M3 r;
    for (int j = 0; j < 3; ++j) for (int i = 0; i < 3; ++i) r.at(i, j) = dot(axA[i], axB[j]);
    V3 tw = b.c - a.c;
    V3 t = conjugate(a.q) * tw;
    bool parallel = false;
    M3 absR;
    for (int i = 0; i < 9; ++i) { absR.m[i] = fabsf(r.m[i]) + kEps; if (absR.m[i] >= 0.99f) parallel = true; }

    float ra, rb, minPen = FLT_MAX;
    V3 normal = v3(0.f);
    bool bFace = false;
    for (int i = 0; i < 3; ++i)
    {
        ra = comp(a.r, i); rb = dot(m3Row(absR, i), b.r);
        float pen = ra + rb - fabsf(comp(t, i));
        if (pen < 0.f) return false;
        if (pen < minPen) { minPen = pen; normal = v3(0.f); setComp(normal, i, 1.f); }
    }
    for (int i = 0; i < 3; ++i)
    {
        ra = dot(m3Col(absR, i), a.r); rb = comp(b.r, i);
        float d = dot(m3Col(r, i), t);
        float pen = ra + rb - fabsf(d);
        if (pen < 0.f) return false;
        if (pen < minPen) { minPen = pen; normal = v3(0.f); setComp(normal, i, 1.f); bFace = true; }
    }
    bool edgeCol = false; V3 edgeN = v3(0.f);
    if (!parallel)
    {
        // nine edge axes a_i x b_j in the reference's order; tables spell out each closed form of :1275-1408
        #define R(i, j) r.at(i, j)
        #define AR(i, j) absR.at(i, j)
        struct EdgeAxis { float ra, rb, dist; V3 n; };
        EdgeAxis e[9] = {
            { a.r.y * AR(2,0) + a.r.z * AR(1,0), b.r.y * AR(0,2) + b.r.z * AR(0,1), fabsf(t.z * R(1,0) - t.y * R(2,0)), v3(0.f, -R(2,0), R(1,0)) },
            { a.r.y * AR(2,1) + a.r.z * AR(1,1), b.r.x * AR(0,2) + b.r.z * AR(0,0), fabsf(t.z * R(1,1) - t.y * R(2,1)), v3(0.f, -R(2,1), R(1,1)) },
            { a.r.y * AR(2,2) + a.r.z * AR(1,2), b.r.x * AR(0,1) + b.r.y * AR(0,0), fabsf(t.z * R(1,2) - t.y * R(2,2)), v3(0.f, -R(2,2), R(1,2)) },
            { a.r.x * AR(2,0) + a.r.z * AR(0,0), b.r.y * AR(1,2) + b.r.z * AR(1,1), fabsf(t.x * R(2,0) - t.z * R(0,0)), v3(R(2,0), 0.f, -R(0,0)) },
            { a.r.x * AR(2,1) + a.r.z * AR(0,1), b.r.x * AR(1,2) + b.r.z * AR(1,0), fabsf(t.x * R(2,1) - t.z * R(0,1)), v3(R(2,1), 0.f, -R(0,1)) },
            { a.r.x * AR(2,2) + a.r.z * AR(0,2), b.r.x * AR(1,1) + b.r.y * AR(1,0), fabsf(t.x * R(2,2) - t.z * R(0,2)), v3(R(2,2), 0.f, -R(0,2)) },
            { a.r.x * AR(1,0) + a.r.y * AR(0,0), b.r.y * AR(2,2) + b.r.z * AR(2,1), fabsf(t.y * R(0,0) - t.x * R(1,0)), v3(-R(1,0), R(0,0), 0.f) },
            { a.r.x * AR(1,1) + a.r.y * AR(0,1), b.r.x * AR(2,2) + b.r.z * AR(2,0), fabsf(t.y * R(0,1) - t.x * R(1,1)), v3(-R(1,1), R(0,1), 0.f) },
            { a.r.x * AR(1,2) + a.r.y * AR(0,2), b.r.x * AR(2,1) + b.r.y * AR(2,0), fabsf(t.y * R(0,2) - t.x * R(1,2)), v3(-R(1,2), R(0,2), 0.f) },
        };
        #undef R
        #undef AR
        for (int k = 0; k < 9; ++k)
        {
            float pen = e[k].ra + e[k].rb - e[k].dist;
            if (pen < 0.f) return false;
            float l = 1.f / len(e[k].n);
            pen *= l;
            if (pen < minPen) { minPen = pen; edgeN = e[k].n * l; edgeCol = true; }
        }
    }
    bool faceCol = !edgeCol;
    if (faceCol) { if (bFace) normal = r * normal; } else { normal = edgeN; }
    normal = a.q * normal;
    if (dot(normal, tw) < 0.f) normal = -normal;
    m.normal = normal;

    if (faceCol)
    {
        V3 cpp[4], cpn[4]; Poly poly; Plane plane;
        if (!bFace)
        {
            boxClipPlanes(a.r, conjugate(a.q) * normal, cpp, cpn);
            boxIncidentVerts(b.r, conjugate(b.q) * normal, poly);
            for (int i = 0; i < 4; ++i)
            {
                cpp[i] = a.q * cpp[i] + a.c; cpn[i] = a.q * cpn[i];
                poly.p[i].v = b.q * poly.p[i].v + b.c;
            }
            V3 refPt = supObb(a)(normal);
            plane = createPlane(refPt, normal);
        }
        else
        {
            boxClipPlanes(b.r, conjugate(b.q) * -normal, cpp, cpn);
            boxIncidentVerts(a.r, conjugate(a.q) * -normal, poly);
            for (int i = 0; i < 4; ++i)
            {
                cpp[i] = b.q * cpp[i] + b.c; cpn[i] = b.q * cpn[i];
                poly.p[i].v = a.q * poly.p[i].v + a.c;
            }
            V3 refPt = supObb(b)(-normal);
            plane = createPlane(refPt, -normal);
        }
        Plane planes[4];
        for (int i = 0; i < 4; ++i)
        {
            planes[i] = createPlane(cpp[i], cpn[i]);
            poly.p[i].depth = -signedDistance(poly.p[i].v, plane);
        }
        if (!clipAndBuild(poly, planes, 4, plane, m)) return false;
    }
    else
    {
        V3 a0, a1, b0, b1;
        boxIncidentEdge(a.r, conjugate(a.q) * normal, a0, a1);
        boxIncidentEdge(b.r, conjugate(b.q) * -normal, b0, b1);
        a0 = a.q * a0 + a.c; a1 = a.q * a1 + a.c; b0 = b.q * b0 + b.c; b1 = b.q * b1 + b.c;
        V3 pa, pb;
        float sq = closestPointSegmentSegment(a0, a1, b0, b1, pa, pb);
        m.n = 1; m.depth[0] = sqrtf(sq); m.point[0] = (pa + pb) * 0.5f;
    }
    return true;
}

// dispatch: the 21 collision<A,B>() instantiations of collision_narrow.cpp:2470-2571
bool collide(const WCol& A, const WCol& B, Manifold& m, NarrowStats& st)
{
    m.n = 0;
    switch (A.type * 6 + B.type)
    {
    case T_SPHERE * 6 + T_SPHERE:    return sphereSphere(A.sphere, B.sphere, m, st);
    case T_SPHERE * 6 + T_CAPSULE:   return sphereCapsule(A.sphere, B.cap, m, st);
    case T_SPHERE * 6 + T_CYLINDER:  return sphereCylinder(A.sphere, B.cap, m, st);
    case T_SPHERE * 6 + T_AABB:      return sphereBox(A.sphere, B.box, m);
    case T_SPHERE * 6 + T_OBB:       return sphereObb(A.sphere, B.obb, m);
    case T_SPHERE * 6 + T_HULL:      return gjkEpa(supSphere(A.sphere), supHull(B.hull), m, st);      // :496-520
    case T_CAPSULE * 6 + T_CAPSULE:  return capsuleVsSegmentShape(A.cap, B.cap, false, m, st);
    case T_CAPSULE * 6 + T_CYLINDER: return capsuleVsSegmentShape(A.cap, B.cap, true, m, st);
    case T_CAPSULE * 6 + T_AABB:     return segmentShapeBox(A.cap, false, B.box, m, st);
    case T_CAPSULE * 6 + T_OBB:      return segmentShapeObb(A.cap, false, B.obb, m, st);
    case T_CAPSULE * 6 + T_HULL:     return gjkEpa(supCapsule(A.cap), supHull(B.hull), m, st);         // :792-818
    case T_CYLINDER * 6 + T_CYLINDER:return cylinderCylinder(A.cap, B.cap, m, st);
    case T_CYLINDER * 6 + T_AABB:    return segmentShapeBox(A.cap, true, B.box, m, st);
    case T_CYLINDER * 6 + T_OBB:     return segmentShapeObb(A.cap, true, B.obb, m, st);
    case T_CYLINDER * 6 + T_HULL:    return gjkEpa(supCylinder(A.cap), supHull(B.hull), m, st);        // :1045-1071
    case T_AABB * 6 + T_AABB:        return boxBox(A.box, B.box, m);
    case T_AABB * 6 + T_OBB:         { Obb o = { qIdentity(), boxCenter(A.box), boxRadius(A.box) }; return obbObb(o, B.obb, m); } // :1142-1148
    case T_AABB * 6 + T_HULL:        return gjkEpa(supBox(A.box), supHull(B.hull), m, st);             // :1150-1176
    case T_OBB * 6 + T_OBB:          return obbObb(A.obb, B.obb, m);
    case T_OBB * 6 + T_HULL:         return gjkEpa(supObb(A.obb), supHull(B.hull), m, st);             // :1529-1555
    case T_HULL * 6 + T_HULL:        return gjkEpa(supHull(A.hull), supHull(B.hull), m, st);           // :1558-1584
    default: ++st.asserts; return false;
    }
}

// --------------------------------------------------------------------------- boolean overlap tests (event layer)
// bounding_volumes.h:301-363, bounding_volumes.cpp:704-733,742-835,1079-1208,1237-1244: what overlapCheck
// (collision_narrow.cpp:1593-1689) dispatches to for rigid body vs trigger / force field pairs.
static bool sphereVsSphere(const Sphere& a, const Sphere& b) { V3 d = a.c - b.c; float rs = a.r + b.r; return dot(d, d) <= rs * rs; }
static bool sphereVsCapsule(const Sphere& s, const Capsule& c) { Sphere o = { closestPointSegment(s.c, c.a, c.b), c.r }; return sphereVsSphere(s, o); }
static bool sphereVsCylinder(const Sphere& s, const Capsule& c)
{
    V3 ab = c.b - c.a;
    float t = dot(s.c - c.a, ab) / sqLen(ab);
    if (t >= 0.f && t <= 1.f) { Sphere o = { lerp(c.a, c.b, t), c.r }; return sphereVsSphere(s, o); }
    V3 p = (t <= 0.f) ? c.a : c.b;
    V3 up = (t <= 0.f) ? -ab : ab;
    V3 proj = normalize(cross(cross(up, s.c - p), up));
    V3 endA = p + proj * c.r, endB = p - proj * c.r;
    V3 closest = closestPointSegment(s.c, endA, endB);
    return sqLen(closest - s.c) <= s.r;          // sic: radius, not radius squared (bounding_volumes.cpp:723)
}
static bool sphereVsAABB(const Sphere& s, const Box& a) { V3 n = closestPointAABB(s.c, a) - s.c; return sqLen(n) <= s.r * s.r; }
static bool sphereVsOBB(const Sphere& s, const Obb& o)
{
    Sphere s_ = { conjugate(o.q) * (s.c - o.c) + o.c, s.r };
    return sphereVsAABB(s_, boxFromCenterRadius(o.c, o.r));
}
static bool capsuleVsCapsule(const Capsule& a, const Capsule& b)
{
    V3 c1, c2; closestPointSegmentSegment(a.a, a.b, b.a, b.b, c1, c2);
    return sphereVsSphere(Sphere{ c1, a.r }, Sphere{ c2, b.r });
}
static bool capsuleVsCylinder(const Capsule& a, const Capsule& b)
{
    V3 c1, c2; closestPointSegmentSegment(a.a, a.b, b.a, b.b, c1, c2);
    return sphereVsCylinder(Sphere{ c1, a.r }, b);
}
static bool gjkTest(const Support& A, const Support& B, NarrowStats& st) { Simplex s; return gjk(A, B, s, st); }   // gjkIntersectionTest
static Capsule intoObbFrame(const Capsule& c, const Obb& o) { Capsule r = { conjugate(o.q) * (c.a - o.c) + o.c, conjugate(o.q) * (c.b - o.c) + o.c, c.r }; return r; }
// bounding_volumes.cpp:1079-1199 (15-axis SAT, early out; no contact data)
static bool obbVsOBB(const Obb& a, const Obb& b)
{
    V3 ax[3] = { a.q * v3(1.f, 0.f, 0.f), a.q * v3(0.f, 1.f, 0.f), a.q * v3(0.f, 0.f, 1.f) };
    V3 bx[3] = { b.q * v3(1.f, 0.f, 0.f), b.q * v3(0.f, 1.f, 0.f), b.q * v3(0.f, 0.f, 1.f) };
    float r[3][3], absR[3][3];                     // r[i][j] = m_ij = dot(axesA[i], axesB[j])
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { r[i][j] = dot(ax[i], bx[j]); absR[i][j] = fabsf(r[i][j]) + kEps; }
    V3 t = conjugate(a.q) * (b.c - a.c);
    const float ta[3] = { t.x, t.y, t.z }, ar[3] = { a.r.x, a.r.y, a.r.z }, br[3] = { b.r.x, b.r.y, b.r.z };
    for (int i = 0; i < 3; ++i)
    {   // a's faces: row i of absR
        float ra = ar[i], rb = dot(v3(absR[i][0], absR[i][1], absR[i][2]), b.r);
        if (ra + rb - fabsf(ta[i]) < 0.f) return false;
    }
    for (int i = 0; i < 3; ++i)
    {   // b's faces: column i
        float ra = dot(v3(absR[0][i], absR[1][i], absR[2][i]), a.r), rb = br[i];
        float d = dot(v3(r[0][i], r[1][i], r[2][i]), t);
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

// overlapCheck (collision_narrow.cpp:1593-1689): A.type <= B.type
bool overlapOnly(const WCol& A, const WCol& B, NarrowStats& st)
{
    switch (A.type * 6 + B.type)
    {
    case T_SPHERE * 6 + T_SPHERE:     return sphereVsSphere(A.sphere, B.sphere);
    case T_SPHERE * 6 + T_CAPSULE:    return sphereVsCapsule(A.sphere, B.cap);
    case T_SPHERE * 6 + T_CYLINDER:   return sphereVsCylinder(A.sphere, B.cap);
    case T_SPHERE * 6 + T_AABB:       return sphereVsAABB(A.sphere, B.box);
    case T_SPHERE * 6 + T_OBB:        return sphereVsOBB(A.sphere, B.obb);
    case T_SPHERE * 6 + T_HULL:       return gjkTest(supSphere(A.sphere), supHull(B.hull), st);
    case T_CAPSULE * 6 + T_CAPSULE:   return capsuleVsCapsule(A.cap, B.cap);
    case T_CAPSULE * 6 + T_CYLINDER:  return capsuleVsCylinder(A.cap, B.cap);
    case T_CAPSULE * 6 + T_AABB:      return gjkTest(supCapsule(A.cap), supBox(B.box), st);
    case T_CAPSULE * 6 + T_OBB:       return gjkTest(supCapsule(intoObbFrame(A.cap, B.obb)), supBox(boxFromCenterRadius(B.obb.c, B.obb.r)), st);
    case T_CAPSULE * 6 + T_HULL:      return gjkTest(supCapsule(A.cap), supHull(B.hull), st);
    case T_CYLINDER * 6 + T_CYLINDER: return gjkTest(supCylinder(A.cap), supCylinder(B.cap), st);
    case T_CYLINDER * 6 + T_AABB:     return gjkTest(supCylinder(A.cap), supBox(B.box), st);
    case T_CYLINDER * 6 + T_OBB:      return gjkTest(supCylinder(intoObbFrame(A.cap, B.obb)), supBox(boxFromCenterRadius(B.obb.c, B.obb.r)), st);
    case T_CYLINDER * 6 + T_HULL:     return gjkTest(supCylinder(A.cap), supHull(B.hull), st);
    case T_AABB * 6 + T_AABB:         return aabbVsAABB(A.box, B.box);
    case T_AABB * 6 + T_OBB:          { Obb o; o.q = qIdentity(); o.c = boxCenter(A.box); o.r = boxRadius(A.box); return obbVsOBB(o, B.obb); }
    case T_AABB * 6 + T_HULL:         return gjkTest(supBox(A.box), supHull(B.hull), st);
    case T_OBB * 6 + T_OBB:           return obbVsOBB(A.obb, B.obb);
    case T_OBB * 6 + T_HULL:          return gjkTest(supObb(A.obb), supHull(B.hull), st);
    case T_HULL * 6 + T_HULL:         return gjkTest(supHull(A.hull), supHull(B.hull), st);
    default: ++st.asserts; return false;
    }
}

} // namespace orc
