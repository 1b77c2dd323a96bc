// ray_interaction.cpp -- testPhysicsInteraction (src/physics/physics.cpp:555-629) for the host mirror: a ray is cast against every
// collider that belongs to a rigid body (in the body's local frame); the nearest hit gets `direction * strength` as a force at the hit
// point (force + torque accumulators, consumed by the next step). The ray / shape tests follow src/physics/bounding_volumes.cpp:
// intersectAABB :215-240, intersectOBB :242-247, intersectTriangle :249-265, intersectSphere :267-293, intersectCylinder :295-361,
// intersectCapsule :363-384, intersectDisk :386-394, intersectHull :677-702.
#include "physics_host.h"
#include <cfloat>
#include <cmath>

namespace b200host {

namespace {

using b200::mk3; using b200::mkq; using b200::dot; using b200::cross; using b200::len; using b200::sqLen; using b200::normalize; using b200::conjugate;

struct lray { vec3 o, d; };

bool intersectPlane(const lray& r, vec3 normal, vec3 point, float& t)
{
    const float ndotd = dot(r.d, normal);
    if (fabsf(ndotd) < 1e-6f) return false;
    t = -(dot(r.o, normal) - dot(normal, point)) / ndotd;
    return true;
}
bool intersectDisk(const lray& r, vec3 pos, vec3 normal, float radius, float& t)
{
    if (!intersectPlane(r, normal, pos, t)) return false;
    return len(r.o + t * r.d - pos) <= radius;
}
bool intersectSphere(const lray& r, vec3 center, float radius, float& t)
{
    const vec3 m = r.o - center;
    const float b = dot(m, r.d), c = dot(m, m) - radius * radius;
    if (c > 0.f && b > 0.f) return false;
    const float discr = b * b - c;
    if (discr < 0.f) return false;
    t = -b - sqrtf(discr);
    if (t < 0.f) t = 0.f;
    return true;
}
bool intersectAABB(const lray& r, vec3 mn, vec3 mx, float& t)
{
    const vec3 inv = mk3(1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z);      // may be inf, still works (bounding_volumes.cpp:217)
    const float tx1 = (mn.x - r.o.x) * inv.x, tx2 = (mx.x - r.o.x) * inv.x;
    t = fminf(tx1, tx2); float tmax = fmaxf(tx1, tx2);
    const float ty1 = (mn.y - r.o.y) * inv.y, ty2 = (mx.y - r.o.y) * inv.y;
    t = fmaxf(t, fminf(ty1, ty2)); tmax = fminf(tmax, fmaxf(ty1, ty2));
    const float tz1 = (mn.z - r.o.z) * inv.z, tz2 = (mx.z - r.o.z) * inv.z;
    t = fmaxf(t, fminf(tz1, tz2)); tmax = fminf(tmax, fmaxf(tz1, tz2));
    return tmax >= t && t > 0.f;
}
bool intersectCylinder(const lray& r, vec3 posA, vec3 posB, float radius, float& t)
{
    const vec3 axis = posB - posA;
    const float height = len(axis);
    const quat q = b200::rotateFromTo(axis, mk3(0.f, 1.f, 0.f));      // math.cpp:538-575
    const vec3 o = q * (r.o - posA), d = q * r.d;                        // base at the origin, second point above it
    const float epsilon = 1e-6f;
    float y = -1.f;
    if (o.x * o.x + o.z * o.z > radius * radius)
    {   // outside the infinite cylinder: only then the side can be hit
        const float a = d.x * d.x + d.z * d.z, b = d.x * o.x + d.z * o.z, c = o.x * o.x + o.z * o.z - radius * radius;
        const float delta = b * b - a * c;
        if (delta < epsilon) return false;
        t = (-b - sqrtf(delta)) / a;
        if (t <= epsilon) return false;
        y = o.y + t * d.y;
    }
    if (y > height + epsilon || y < -epsilon)
    {   // caps
        const lray local = { o, d };
        float dist;
        if (d.y < 0.f && intersectDisk(local, mk3(0.f, height, 0.f), mk3(0.f, 1.f, 0.f), radius, dist)) t = dist;
        if (d.y > 0.f && intersectDisk(local, mk3(0.f, 0.f, 0.f), mk3(0.f, -1.f, 0.f), radius, dist)) t = dist;
        y = o.y + t * d.y;
    }
    return y > -epsilon && y < height + epsilon;
}
bool intersectCapsule(const lray& r, vec3 a, vec3 b, float radius, float& outT)
{
    outT = FLT_MAX; float t; bool result = false;
    if (intersectCylinder(r, a, b, radius, t)) { outT = t; result = true; }
    if (intersectSphere(r, a, radius, t)) { outT = fminf(outT, t); result = true; }
    if (intersectSphere(r, b, radius, t)) { outT = fminf(outT, t); result = true; }
    return result;
}
bool pointInTriangle(vec3 point, vec3 A, vec3 B, vec3 C)
{   // math.cpp:1273-1290: the sign-bit trick spelled out (z < 0 and x >= 0 and y >= 0, with -0 counting as negative like the bit test)
    const vec3 e10 = B - A, e20 = C - A;
    const float a = dot(e10, e10), b = dot(e10, e20), c = dot(e20, e20), ac_bb = a * c - b * b;
    const vec3 vp = point - A;
    const float d = dot(vp, e10), e = dot(vp, e20);
    const float x = d * c - e * b, y = e * a - d * b, z = x + y - ac_bb;
    return std::signbit(z) && !std::signbit(x) && !std::signbit(y);
}
bool intersectTriangle(const lray& r, vec3 a, vec3 b, vec3 c, float& t)
{
    vec3 normal = cross(b - a, c - a); const float l = len(normal); normal = l > 1e-6f ? normal * (1.f / l) : mk3(0.f);     // noz
    const float d = -dot(normal, a), nDotR = dot(r.d, normal);
    if (fabsf(nDotR) <= 1e-6f) return false;
    t = -(dot(r.o, normal) + d) / nDotR;
    return t >= 0.f && pointInTriangle(r.o + t * r.d, a, b, c);
}

} // namespace

void testPhysicsInteraction(game_scene& scene, vec3 origin, vec3 direction, float strength, uint32_t firstBody, uint32_t numBodies)
{
    float minT = FLT_MAX; int32_t minRB = -1; vec3 force = mk3(0.f), torque = mk3(0.f);
    for (const collider_component& collider : scene.colliders)
    {
        const game_scene::entity_data& parent = scene.entities[collider.parentEntity];
        if (parent.rigidBody < 0) continue;
        if ((uint32_t)parent.rigidBody < firstBody || (uint32_t)parent.rigidBody - firstBody >= numBodies) continue;
        const trs& transform = scene.physicsTransform1[parent.rigidBody];
        const lray local = { conjugate(transform.rotation) * (origin - transform.position), conjugate(transform.rotation) * direction };
        float t = 0.f; bool hit = false;
        switch (collider.type)
        {
        case collider_type_sphere: hit = intersectSphere(local, collider.sphere.center, collider.sphere.radius, t); break;
        case collider_type_capsule: hit = intersectCapsule(local, collider.capsule.positionA, collider.capsule.positionB, collider.capsule.radius, t); break;
        case collider_type_cylinder: hit = intersectCylinder(local, collider.cylinder.positionA, collider.cylinder.positionB, collider.cylinder.radius, t); break;
        case collider_type_aabb: hit = intersectAABB(local, collider.aabb.minCorner, collider.aabb.maxCorner, t); break;
        case collider_type_obb:
        {
            const lray lr = { conjugate(collider.obb.rotation) * (local.o - collider.obb.center), conjugate(collider.obb.rotation) * local.d };
            hit = intersectAABB(lr, mk3(0.f) - collider.obb.radius, collider.obb.radius, t);
        } break;
        default:
        {   // hull: nearest triangle in the hull's frame (bounding_volumes.cpp:677-702)
            const bounding_hull_geometry& g = scene.hullGeometries[collider.hull.geometryIndex];
            const lray lr = { conjugate(collider.hull.rotation) * (local.o - collider.hull.position), conjugate(collider.hull.rotation) * local.d };
            float best = FLT_MAX;
            for (const indexed_triangle32& tri : g.faces) { float tt; if (intersectTriangle(lr, g.vertices[tri.a], g.vertices[tri.b], g.vertices[tri.c], tt) && tt < best) { best = tt; hit = true; } }
            t = best;
        } break;
        }
        if (hit && t < minT)
        {
            minT = t; minRB = parent.rigidBody;
            const vec3 globalHit = transform.rotation * (transform.scale * (local.o + t * local.d)) + transform.position;
            const vec3 cog = transform.position + transform.rotation * scene.rigidBodies[parent.rigidBody].localCOGPosition;
            force = direction * strength;
            torque = cross(globalHit - cog, force);
        }
    }
    if (minRB >= 0) { scene.rigidBodies[minRB].torqueAccumulator += torque; scene.rigidBodies[minRB].forceAccumulator += force; }
}

} // namespace b200host
