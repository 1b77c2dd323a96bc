// oracle/om_shapes.h -- TEST INFRASTRUCTURE ONLY (parity oracle). Never included by the product path.
// World-space shape records, contact manifold and the closest-point helpers the narrow phase uses.
// Follows reference src/physics/bounding_volumes.h / bounding_volumes.cpp (cited per function).
#pragma once
#include "om_math.h"
#include <vector>

namespace orc {

enum { T_SPHERE = 0, T_CAPSULE = 1, T_CYLINDER = 2, T_AABB = 3, T_OBB = 4, T_HULL = 5 }; // physics.h:59-70
enum { O_RIGID = 0, O_STATIC = 1, O_FORCE_FIELD = 2, O_TRIGGER = 3 };                    // physics.h:49-57

struct Box { V3 mn, mx; };  // bounding_box (bounding_volumes.h:104-130)

// bounding_hull_geometry (bounding_volumes.h:208-218): only what the step reads (vertices for GJK support, aabb)
struct HullGeom { std::vector<V3> verts; std::vector<uint32_t> tris; Box aabb; }; // tris: 3 indices per face (mass properties only)

struct Sphere { V3 c; float r; };
struct Capsule { V3 a, b; float r; };   // also used for cylinders (same fields, bounding_volumes.h:44-74)
struct Obb { Q4 q; V3 c; V3 r; };
struct Hull { Q4 q; V3 p; const HullGeom* g; };

// one world-space collider = collider_union after getWorldSpaceColliders (physics.cpp:631-756)
struct WCol
{
    int type;
    int objType;
    uint32_t objIndex;
    float restitution, friction;
    Sphere sphere; Capsule cap; Box box; Obb obb; Hull hull; // only the member selected by `type` is valid
};

// contact_manifold (collision_narrow.cpp:40-46)
struct Manifold { V3 point[4]; float depth[4]; V3 normal; uint32_t n; };

// bounding_volumes.cpp:25-35,38-48
static inline void grow(Box& b, V3 o) { b.mn = vmin(b.mn, o); b.mx = vmax(b.mx, o); }
static inline Box boxNegInf() { Box b = { v3(FLT_MAX), v3(-FLT_MAX) }; return b; }
static inline V3 boxCenter(const Box& b) { return (b.mn + b.mx) * 0.5f; }
static inline V3 boxRadius(const Box& b) { return (b.mx - b.mn) * 0.5f; }
static inline Box boxFromCenterRadius(V3 c, V3 r) { Box b = { c - r, c + r }; return b; }

// bounding_volumes.cpp:58-70 : 8 rotated corners, in this exact order
static inline Box boxTransformToAABB(const Box& b, Q4 q, V3 t)
{
    Box r = boxNegInf();
    grow(r, q * b.mn + t);
    grow(r, q * v3(b.mx.x, b.mn.y, b.mn.z) + t);
    grow(r, q * v3(b.mn.x, b.mx.y, b.mn.z) + t);
    grow(r, q * v3(b.mx.x, b.mx.y, b.mn.z) + t);
    grow(r, q * v3(b.mn.x, b.mn.y, b.mx.z) + t);
    grow(r, q * v3(b.mx.x, b.mn.y, b.mx.z) + t);
    grow(r, q * v3(b.mn.x, b.mx.y, b.mx.z) + t);
    grow(r, q * b.mx + t);
    return r;
}

// bounding_volumes.h:352-358 : inclusive on every axis
static inline bool aabbVsAABB(const Box& a, const Box& b)
{
    if (a.mx.x < b.mn.x || a.mn.x > b.mx.x) return false;
    if (a.mx.y < b.mn.y || a.mn.y > b.mx.y) return false;
    if (a.mx.z < b.mn.z || a.mn.z > b.mx.z) return false;
    return true;
}

// bounding_volumes.h:365-371
static inline V3 closestPointSegment(V3 q, V3 a, V3 b)
{
    V3 ab = b - a;
    float t = dot(q - a, ab) / sqLen(ab);
    t = clampf(t, 0.f, 1.f);
    return a + t * ab;
}
// bounding_volumes.h:373-384
static inline V3 closestPointAABB(V3 q, const Box& b)
{
    V3 r;
    for (int i = 0; i < 3; ++i)
    {
        float v = comp(q, i);
        if (v < comp(b.mn, i)) v = comp(b.mn, i);
        if (v > comp(b.mx, i)) v = comp(b.mx, i);
        setComp(r, i, v);
    }
    return r;
}
// bounding_volumes.cpp:1251-1318
static inline float closestPointSegmentSegment(V3 p1, V3 q1, V3 p2, V3 q2, V3& c1, V3& c2)
{
    float s, t;
    V3 d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
    float a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
    if (a <= kEps && e <= kEps) { c1 = p1; c2 = p2; return dot(c1 - c2, c1 - c2); }
    if (a <= kEps) { s = 0.f; t = f / e; t = clampf(t, 0.f, 1.f); }
    else
    {
        float c = dot(d1, r);
        if (e <= kEps) { t = 0.f; s = clampf(-c / a, 0.f, 1.f); }
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
    c1 = p1 + d1 * s;
    c2 = p2 + d2 * t;
    return sqLen(c1 - c2);
}

// plane helpers (bounding_volumes.h:166-170,296-299): signed distance = dot(vec4(p,1), plane) with the vec4 hadd order
struct Plane { V3 n; float d; };
static inline Plane createPlane(V3 point, V3 normal) { Plane p = { normal, -dot(normal, point) }; return p; }
static inline float signedDistance(V3 p, const Plane& pl) { return (p.x * pl.n.x + p.y * pl.n.y) + (p.z * pl.n.z + 1.f * pl.d); }

// narrow phase entry: pair already ordered so that a.type <= b.type (collision_narrow.cpp:2374)
struct NarrowStats { uint32_t gjkIterCap = 0, gjkErrors = 0, epaNotConverged = 0, epaOutOfMemory = 0, asserts = 0;
                     uint32_t epaMaxTriangles = 0, epaMaxEdges = 0, epaMaxPoints = 0; };
bool collide(const WCol& a, const WCol& b, Manifold& out, NarrowStats& st);
// boolean overlap tests for trigger / force-field pairs (collision_narrow.cpp:1593-1689)
bool overlapOnly(const WCol& a, const WCol& b, NarrowStats& st);

} // namespace orc
