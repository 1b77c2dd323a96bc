// ray_interaction.cpp -- testPhysicsInteraction (src/physics/physics.cpp:555-629) for the host mirror: a ray is cast against every
// collider that belongs to a rigid body (in the body's local frame); the nearest hit gets `direction * strength` as a force at the hit
// point (force + torque accumulators, consumed by the next step). The ray / shape tests follow src/physics/bounding_volumes.cpp:
// intersectAABB :215-240, intersectOBB :242-247, intersectTriangle :249-265, intersectSphere :267-293, intersectCylinder :295-361,
// intersectCapsule :363-384, intersectDisk :386-394, intersectHull :677-702.
#include "physics_host.h"
#include "../csrc/b200_ray.cuh"
#include <cfloat>
#include <cmath>

namespace b200host {

using b200::mk3; using b200::mkq; using b200::dot; using b200::cross; using b200::len; using b200::sqLen; using b200::normalize; using b200::conjugate;
using b200::lray; using b200::intersectSphere; using b200::intersectCapsule; using b200::intersectCylinder; using b200::intersectAABB; using b200::intersectTriangle;

void testPhysicsInteraction(game_scene& scene, vec3 origin, vec3 direction, float strength, uint32_t firstBody, uint32_t numBodies, uint32_t firstCollider, uint32_t numColliders)
{
    float minT = FLT_MAX; int32_t minRB = -1; vec3 force = mk3(0.f), torque = mk3(0.f);
    const size_t c0 = firstCollider < scene.colliders.size() ? firstCollider : scene.colliders.size();
    const size_t c1 = (numColliders < scene.colliders.size() - c0) ? c0 + numColliders : scene.colliders.size();
    for (size_t ci = c0; ci < c1; ++ci)
    {
        const collider_component& collider = scene.colliders[ci];
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
