// oracle/om_step.h -- TEST INFRASTRUCTURE ONLY (parity oracle). Never included by the product path.
// State and stage functions of the CPU restatement of physicsStepInternal (reference src/physics/physics.cpp:1180-1362).
#pragma once
#include "om_shapes.h"
#include "../include/b200phys.h"   // boundary POD structs only (no product code)
#include <vector>

namespace orc {

// rigid_body_global_state (rigid_body.h:6-16)
struct BodyG { Q4 rotation; V3 localCOG; V3 position; M3 invInertia; float invMass; V3 v; V3 w; };

struct Contact { V3 point; float depth; V3 normal; uint32_t fr; };  // collision_contact (physics.h:347-354)
struct Pair { uint32_t a, b; };

// collision_constraint (constraints.h:615-631)
struct ContactRow
{
    V3 rA, rB, tangent;
    V3 tAngA, tAngB, nAngA, nAngB;
    float impN, impT, massN, massT, bias;
};

enum PairOrder { ORDER_REFERENCE_SAP = 0, ORDER_CANONICAL = 1 };

// heightmap_collider_component + its chunks (terrain/heightmap_collider.h:19-150), om_heightmap.cpp
struct Heightmap
{
    struct MinMax { uint16_t mn, mx; };
    bool present = false;
    uint32_t chunksPerDim = 0; float chunkSize = 0.f, invChunkSize = 0.f, chunkScale = 0.f, heightScale = 0.f, invAmplitudeScale = 1.f;
    V3 minCorner = v3(0.f); float restitution = 0.f, friction = 0.f;
    std::vector<uint16_t> heights;                                  // chunksPerDim^2 chunks x 129 x 129, chunk (x, z) at z * chunksPerDim + x
    std::vector<std::vector<std::vector<MinMax>>> mips;             // per chunk: 8 levels of (min, max), 128^2 ... 1
    void set(uint32_t chunksPerDim, float chunkSize, V3 minCorner, float amplitudeScale, float restitution, float friction, const uint16_t* heights);
    float heightAt(float x, float z) const;                         // -FLT_MAX outside
};

struct SapEndpoint { float value; uint32_t collider; bool start; };

struct Oracle
{
    // scene (host-order arrays exactly as passed through the C ABI)
    std::vector<b200phys_body> bodies;
    std::vector<b200phys_transform> transforms;
    std::vector<b200phys_collider> colliders;
    std::vector<HullGeom> hulls;
    std::vector<uint8_t> joints[6];                 // raw structs per kind
    std::vector<b200phys_body_pair> jointPairs[6];

    // persistent broad-phase state (sap_context, collision_broad.cpp:20-24)
    std::vector<SapEndpoint> endpoints;
    uint32_t sortingAxis = 0;
    uint32_t sapColliderCount = 0;

    // per-step outputs
    std::vector<WCol> wcol;
    std::vector<Box> aabbs;
    std::vector<Pair> pairs;            // broad phase, in the order selected by `pairOrder`
    std::vector<Pair> collisions;       // colliding collider pairs (type-ordered), one per manifold
    std::vector<uint8_t> contactCount;  // per collision
    std::vector<Contact> contacts;
    std::vector<b200phys_body_pair> contactBodies;
    std::vector<BodyG> g;               // numBodies + 1 (dummy)
    std::vector<ContactRow> rows;
    NarrowStats nstats;
    uint32_t numNonCollisionInteractions = 0;

    // collision begin / end events (handleCollisionCallbacks, physics.cpp:1037-1178), keyed by collider pair instead of entity pair
    struct CollisionKey { uint32_t a, b, contactOffset, numContacts; };
    struct BeginEvent { uint32_t a, b; V3 point, normal, relativeVelocity; };
    std::vector<CollisionKey> prevFrameCollisions;      // event_context::prevFrameCollisions (physics.cpp:946-950)
    std::vector<BeginEvent> beginEvents;
    std::vector<Pair> endEvents;
    void collisionEvents();

    // non-collision interactions (collision_narrow.h:9-14): rigid body vs force field / trigger overlaps of this step
    struct Interaction { uint32_t rigidBodyIndex, otherIndex; int otherType; };
    std::vector<Interaction> interactions;
    std::vector<V3> forceFields;                        // force_field_global_state per localized force field (physics.cpp:47-50,759-787)
    V3 globalForceField = v3(0.f);                      // sum of the force fields without a collider
    std::vector<Pair> prevFrameTriggerOverlaps;         // (trigger, rigid body), sorted, unique (physics.cpp:946-950)
    std::vector<Pair> triggerEnter, triggerLeave;
    void nonCollisionInteractions();                    // handleNonCollisionInteractions, physics.cpp:952-1035

    int pairOrder = ORDER_CANONICAL;
    // timing of the last step by reference stage name (ms)
    double msWorld = 0, msBroad = 0, msNarrow = 0, msForces = 0, msInit = 0, msSolve = 0, msIntegrate = 0;

    void worldColliders();                               // physics.cpp:631-756
    void broadphaseSAP();                                // collision_broad.cpp:297-447 + :87-166
    void broadphaseBrute(std::vector<Pair>& out) const;  // all-pairs, same predicate
    void narrowphase();                                  // collision_narrow.cpp:2328-2603
    Heightmap heightmap;
    void heightmapCollisions();                          // heightmap_collision.cpp:509-618 behind the narrow phase (physics.cpp:1238-1248)
    void integrateForces(float dt);                      // rigid_body.cpp:95-124 over physics.cpp:1266-1279
    void initContacts(float dt);                         // constraints.cpp:3307-3379
    void solveContacts(const uint32_t* order, uint32_t n); // constraints.cpp:3381-3449
    void integrateVelocities(float dt);                  // rigid_body.cpp:126-142

    // joints (om_joints.cpp)
    void initJoints(float dt);
    void solveJoints(int kind, const uint32_t* order);
    std::vector<uint8_t> jointRows[6];

    // one physicsStepInternal. contactOrder / jointOrder[k]: optional visiting order inside an iteration (the analogue of
    // the reference's AVX2 re-batching, constraints.cpp:51-184); null = array order.
    void step(float dt, uint32_t iterations, const uint32_t* contactOrder, const uint32_t* const jointOrder[6]);
};

} // namespace orc
