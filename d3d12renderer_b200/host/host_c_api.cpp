// host_c_api.cpp -- C entry points over the host mirror for the Python tests / bench: synthetic scene builders for the
// BASELINE.json configurations, flat-array export, and the physicsStep() driver.
#include "physics_host.h"
#include <fstream>
#include <cstring>
#include <memory>

using namespace b200;
using namespace b200host;

namespace {

// random_number_generator (reference src/core/random.h:5-50)
struct Rng
{
    uint64_t state;
    uint64_t u64() { uint64_t x = state; x ^= x << 13; x ^= x >> 7; x ^= x << 17; state = x; return x; }
    float f01() { return (uint32_t)u64() / (float)UINT32_MAX; }
    float between(float lo, float hi) { float t = (f01() - 0.f) / (1.f - 0.f); return lo + t * (hi - lo); }
};

const physics_material kDynamicMat = { physics_material_type_wood, 0.1f, 0.5f, 1.f };   // application.cpp:215
const physics_material kStaticMat = { physics_material_type_metal, 0.1f, 1.f, 4.f };    // application.cpp:248-251

void addGround(game_scene& s, float halfX, float halfZ)
{
    s.createEntity("Platform").addTransform(mk3(0.f, -4.f, 0.f)).addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(0.f), mk3(halfX, 4.f, halfZ)), kStaticMat));
}
void addWalls(game_scene& s, float ex, float ez, float h)
{
    scene_entity w = s.createEntity("Walls"); w.addTransform(mk3(0.f));
    w.addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(-(ex + 2.f), h * 0.5f, 0.f), mk3(2.f, h * 0.5f + 2.f, ez + 4.f)), kStaticMat));
    w.addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3((ex + 2.f), h * 0.5f, 0.f), mk3(2.f, h * 0.5f + 2.f, ez + 4.f)), kStaticMat));
    w.addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(0.f, h * 0.5f, -(ez + 2.f)), mk3(ex + 4.f, h * 0.5f + 2.f, 2.f)), kStaticMat));
    w.addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(0.f, h * 0.5f, (ez + 2.f)), mk3(ex + 4.f, h * 0.5f + 2.f, 2.f)), kStaticMat));
}

inline float deg2radf(float d) { return d * (B200_PI / 180.f); }

quat randomRotation(Rng& rng)
{   // axis = normalize(uniform[-1,1]^3), angle = uniform[0, pi]
    vec3 axis = mk3(rng.between(-1.f, 1.f), rng.between(-1.f, 1.f), rng.between(-1.f, 1.f));
    if (sqLen(axis) < 1e-6f) axis = mk3(0.f, 1.f, 0.f);
    return quatAxisAngle(normalize(axis), rng.between(0.f, B200_PI));
}

// two shared convex hulls: an 8-vertex frustum and a 12-vertex icosahedron (radius ~0.5)
void addHullGeometries(game_scene& s)
{
    const vec3 fr[8] = { mk3(-0.5f, -0.4f, -0.5f), mk3(0.5f, -0.4f, -0.5f), mk3(0.5f, -0.4f, 0.5f), mk3(-0.5f, -0.4f, 0.5f),
                         mk3(-0.3f, 0.4f, -0.3f), mk3(0.3f, 0.4f, -0.3f), mk3(0.3f, 0.4f, 0.3f), mk3(-0.3f, 0.4f, 0.3f) };
    const indexed_triangle32 ft[12] = { {0,1,2},{0,2,3}, {4,6,5},{4,7,6}, {0,4,5},{0,5,1}, {1,5,6},{1,6,2}, {2,6,7},{2,7,3}, {3,7,4},{3,4,0} };
    s.allocateBoundingHullGeometry(fr, 8, ft, 12);
    const float t = 1.618034f, k = 0.5f / 1.902113f;
    const vec3 ic[12] = { k * mk3(-1, t, 0), k * mk3(1, t, 0), k * mk3(-1, -t, 0), k * mk3(1, -t, 0), k * mk3(0, -1, t), k * mk3(0, 1, t),
                          k * mk3(0, -1, -t), k * mk3(0, 1, -t), k * mk3(t, 0, -1), k * mk3(t, 0, 1), k * mk3(-t, 0, -1), k * mk3(-t, 0, 1) };
    const indexed_triangle32 it[20] = { {0,11,5},{0,5,1},{0,1,7},{0,7,10},{0,10,11},{1,5,9},{5,11,4},{11,10,2},{10,7,6},{7,1,8},
                                        {3,9,4},{3,4,2},{3,2,6},{3,6,8},{3,8,9},{4,9,5},{2,4,11},{6,2,10},{8,6,7},{9,8,1} };
    s.allocateBoundingHullGeometry(ic, 12, it, 20);
}

} // namespace

extern "C" {

void* b200host_scene_create() { return new game_scene(); }
void b200host_scene_destroy(void* s) { delete (game_scene*)s; }

// config 1: nx*nz spheres over the ground (mode 0: y = 0.5 + rand*10; mode 1: y = 0.45 "at rest")           [SURVEY 8d config 1]
// config 2: nx*ny*nz spheres on a jittered lattice inside 5 static boxes (spacing in 1/1000 via `param`)       [config 2]
// config 3: nx*ny*nz mixed primitives (sphere / capsule / OBB / hull round-robin), random orientations        [config 3]
// config 4: nx*nz humanoid ragdolls (14 bodies, 17 colliders, 13 joints each) above the ground                [config 4]
static std::vector<uint32_t> gTriggerLog;   // (trigger entity, other entity, type) triples appended by the config-5 trigger callback
uint32_t b200host_trigger_log(uint32_t* out, uint32_t capacityTriples, int clear)
{
    const uint32_t n = (uint32_t)gTriggerLog.size() / 3;
    if (out) memcpy(out, gTriggerLog.data(), 4 * (size_t)3 * (n < capacityTriples ? n : capacityTriples));
    if (clear) gTriggerLog.clear();
    return n;
}

int b200host_build_config(void* sp, int config, uint32_t nx, uint32_t ny, uint32_t nz, uint64_t seed, int mode, float spacing)
{
    game_scene& s = *(game_scene*)sp;
    Rng rng = { seed ? seed : 1 };
    if (config == 1)
    {
        const float half = (nx - 1) * spacing * 0.5f, halfz = (nz - 1) * spacing * 0.5f;
        for (uint32_t ix = 0; ix < nx; ++ix) for (uint32_t iz = 0; iz < nz; ++iz)
        {
            float y = (mode == 0) ? 0.5f + rng.between(0.f, 10.f) : 0.45f;
            s.createEntity("Sphere").addTransform(mk3(ix * spacing - half, y, iz * spacing - halfz))
                .addCollider(collider_component::asSphere({ mk3(0.f), 0.5f }, kDynamicMat)).addRigidBody(false, 1.f);
        }
        addGround(s, fmaxf(30.f, half + 2.f), fmaxf(30.f, halfz + 2.f));
        return 0;
    }
    if (config == 5)
    {   // event-layer scene: spheres dropping through a trigger slab and a localized wind box, plus a global force field
        const float half = (nx - 1) * spacing * 0.5f, halfz = (nz - 1) * spacing * 0.5f;
        for (uint32_t ix = 0; ix < nx; ++ix) for (uint32_t iz = 0; iz < nz; ++iz)
            s.createEntity("Sphere").addTransform(mk3(ix * spacing - half, 3.f + rng.between(0.f, 2.f), iz * spacing - halfz))
                .addCollider(collider_component::asSphere({ mk3(0.f), 0.5f }, kDynamicMat)).addRigidBody(false, 1.f);
        addGround(s, fmaxf(30.f, half + 2.f), fmaxf(30.f, halfz + 2.f));
        s.createEntity("Trigger").addTransform(mk3(0.f, 1.6f, 0.f)).addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(0.f), mk3(half + 2.f, 0.3f, halfz + 2.f)), kStaticMat))
            .addTrigger([](const trigger_event& e) { gTriggerLog.push_back(e.trigger); gTriggerLog.push_back(e.other); gTriggerLog.push_back((uint32_t)e.type); });
        s.createEntity("Wind").addTransform(mk3(0.f, 2.6f, 0.f)).addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(0.f), mk3(half + 2.f, 0.5f, halfz + 2.f)), kStaticMat))
            .addForceField(mk3(5.f, 0.f, 0.f));
        s.createEntity("GlobalWind").addForceField(mk3(0.f, 0.f, 1.f));
        return 0;
    }
    if (config == 2 || config == 3)
    {
        addHullGeometries(s);
        const float hx = (nx - 1) * spacing * 0.5f, hz = (nz - 1) * spacing * 0.5f;
        uint32_t counter = 0;
        for (uint32_t ix = 0; ix < nx; ++ix) for (uint32_t iy = 0; iy < ny; ++iy) for (uint32_t iz = 0; iz < nz; ++iz, ++counter)
        {
            vec3 p = mk3(ix * spacing - hx + rng.between(-0.005f, 0.005f), iy * spacing + 0.6f + rng.between(-0.005f, 0.005f), iz * spacing - hz + rng.between(-0.005f, 0.005f));
            scene_entity e = s.createEntity("Body");
            const int kind = (config == 2) ? 0 : (int)(counter & 3u);
            if (kind == 0) { e.addTransform(p).addCollider(collider_component::asSphere({ mk3(0.f), 0.5f }, kDynamicMat)); }
            else
            {
                e.addTransform(p, randomRotation(rng));
                if (kind == 1) e.addCollider(collider_component::asCapsule({ mk3(0.f, -0.25f, 0.f), mk3(0.f, 0.25f, 0.f), 0.25f }, kDynamicMat));
                else if (kind == 2) e.addCollider(collider_component::asOBB({ qIdentity(), mk3(0.f), mk3(0.4f, 0.3f, 0.35f) }, kDynamicMat));
                else e.addCollider(collider_component::asHull({ qIdentity(), mk3(0.f), (counter >> 2) & 1u }, kDynamicMat));
            }
            e.addRigidBody(false, 1.f);
        }
        const float ex = hx + 1.1f, ez = hz + 1.1f, h = ny * spacing + 4.f;
        s.createEntity("Floor").addTransform(mk3(0.f, -4.f, 0.f)).addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(0.f), mk3(ex + 8.f, 4.f, ez + 8.f)), kStaticMat));
        addWalls(s, ex, ez, h);
        return 0;
    }
    if (config == 6)
    {   // joint zoo: nx x nz copies of eleven small mechanisms covering distance / ball / fixed / slider joints and every motor and
        // limit of the hinge, cone-twist and slider joints (constraints.cpp:189-264,460-528,736-823,1079-1307,1782-2070,2638-2846).
        // Mechanisms hang from kinematic anchors (invMass 0), are spaced so that different mechanisms never touch, and start out of
        // equilibrium so that limits and motors are active within the first steps. `mode` seeds small per-copy variations.
        const float DX = spacing > 0.f ? spacing : 30.f;
        const physics_material mat = kDynamicMat;
        for (uint32_t cx = 0; cx < nx; ++cx) for (uint32_t cz = 0; cz < nz; ++cz)
        {
            const vec3 O = mk3(cx * DX, 0.f, cz * 40.f);
            const float jig = rng.between(-0.05f, 0.05f);
            auto anchorAt = [&](vec3 p) { scene_entity a = s.createEntity("Anchor"); a.addTransform(O + p).addRigidBody(true); return a; };
            auto sphereAt = [&](vec3 p, float r) { scene_entity e = s.createEntity("Ball"); e.addTransform(O + p).addCollider(collider_component::asSphere({ mk3(0.f), r }, mat)).addRigidBody(false); return e; };
            auto capsuleAt = [&](vec3 p, quat q, float h, float r) { scene_entity e = s.createEntity("Link"); e.addTransform(O + p, q).addCollider(collider_component::asCapsule({ mk3(0.f, -h, 0.f), mk3(0.f, h, 0.f), r }, mat)).addRigidBody(false); return e; };
            auto boxAt = [&](vec3 p, quat q, vec3 r) { scene_entity e = s.createEntity("Box"); e.addTransform(O + p, q).addCollider(collider_component::asOBB({ qIdentity(), mk3(0.f), r }, mat)).addRigidBody(false); return e; };
            const quat horiz = quatAxisAngle(mk3(0.f, 0.f, 1.f), deg2radf(90.f));      // capsule axis along x
            {   // A: distance chain (pendulum of three spheres), starts sideways
                scene_entity prev = anchorAt(mk3(0.f, 8.f, 0.f)); vec3 pp = O + mk3(0.f, 8.f, 0.f);
                for (int i = 0; i < 3; ++i)
                {
                    const vec3 p = mk3(0.8f * (i + 1) + jig, 8.f - 0.3f * (i + 1), 0.f);
                    scene_entity b = sphereAt(p, 0.25f);
                    addDistanceConstraintFromGlobalPoints(prev, b, pp, O + p);
                    prev = b; pp = O + p;
                }
            }
            {   // B: ball-joint chain of three capsules laid out horizontally
                scene_entity prev = anchorAt(mk3(0.f, 8.f, 3.f));
                for (int i = 0; i < 3; ++i)
                {
                    scene_entity b = capsuleAt(mk3(0.5f + 1.0f * i, 8.f, 3.f), horiz, 0.35f, 0.12f);
                    addBallConstraintFromGlobalPoints(prev, b, O + mk3(1.0f * i, 8.f, 3.f));
                    prev = b;
                }
            }
            {   // C: two boxes welded by a fixed joint, tumbling onto the ground
                scene_entity a = boxAt(mk3(0.f, 0.75f + jig, 6.f), quatAxisAngle(normalize(mk3(1.f, 0.3f, 0.2f)), 0.4f), mk3(0.4f, 0.2f, 0.3f));
                scene_entity b = boxAt(mk3(0.9f, 1.05f + jig, 6.f), quatAxisAngle(normalize(mk3(0.f, 1.f, 0.5f)), 0.9f), mk3(0.3f, 0.3f, 0.2f));
                addFixedConstraintFromGlobalPoints(a, b, O + mk3(0.45f, 0.9f, 6.f));
                a.rigidBody().angularVelocity = mk3(1.f, 2.f, -1.5f); b.rigidBody().linearVelocity = mk3(0.f, -1.f, 0.5f);
            }
            {   // D: slider along x with limits and a velocity motor
                scene_entity base = anchorAt(mk3(0.f, 5.f, 9.f)); scene_entity b = boxAt(mk3(0.f, 5.f, 9.f), qIdentity(), mk3(0.3f, 0.2f, 0.2f));
                slider_constraint& c = getConstraint(s, addSliderConstraintFromGlobalPoints(base, b, O + mk3(0.f, 5.f, 9.f), mk3(1.f, 0.f, 0.f), -0.05f, 0.06f));
                c.motorType = constraint_velocity_motor; c.maxMotorForce = 200.f; c.motorVelocityOrTargetDistance = 1.5f;
            }
            {   // E: vertical slider with a position motor holding the box against gravity, no limits
                scene_entity base = anchorAt(mk3(0.f, 5.f, 12.f)); scene_entity b = boxAt(mk3(0.f, 5.f, 12.f), qIdentity(), mk3(0.2f, 0.2f, 0.2f));
                slider_constraint& c = getConstraint(s, addSliderConstraintFromGlobalPoints(base, b, O + mk3(0.f, 5.f, 12.f), normalize(mk3(0.1f, 1.f, 0.f))));
                c.motorType = constraint_position_motor; c.maxMotorForce = 80.f; c.motorVelocityOrTargetDistance = 0.3f;
            }
            {   // F: hinge with a velocity motor (no limits)
                scene_entity base = anchorAt(mk3(0.f, 5.f, 15.f)); scene_entity b = capsuleAt(mk3(0.5f, 5.f, 15.f), horiz, 0.4f, 0.1f);
                hinge_constraint& c = getConstraint(s, addHingeConstraintFromGlobalPoints(base, b, O + mk3(0.f, 5.f, 15.f), mk3(0.f, 0.f, 1.f)));
                c.motorType = constraint_velocity_motor; c.maxMotorTorque = 50.f; c.motorVelocityOrTargetAngle = 2.f;
            }
            {   // G: hinge with limits and a position motor
                scene_entity base = anchorAt(mk3(0.f, 5.f, 18.f)); scene_entity b = capsuleAt(mk3(0.5f, 5.f, 18.f), horiz, 0.4f, 0.1f);
                hinge_constraint& c = getConstraint(s, addHingeConstraintFromGlobalPoints(base, b, O + mk3(0.f, 5.f, 18.f), mk3(0.f, 0.f, 1.f), deg2radf(-45.f), deg2radf(60.f)));
                c.motorType = constraint_position_motor; c.maxMotorTorque = 40.f; c.motorVelocityOrTargetAngle = 0.5f;
            }
            {   // H: free hinge with a narrow limit: falls under gravity into the limit within a few steps
                scene_entity base = anchorAt(mk3(0.f, 5.f, 21.f)); scene_entity b = capsuleAt(mk3(0.5f, 5.f, 21.f), horiz, 0.4f, 0.1f);
                addHingeConstraintFromGlobalPoints(base, b, O + mk3(0.f, 5.f, 21.f), mk3(0.f, 0.f, 1.f), deg2radf(-2.f), deg2radf(2.f));
                b.rigidBody().angularVelocity = mk3(0.f, 0.f, -3.f);
            }
            {   // I: cone-twist, swing velocity motor + twist limit
                scene_entity base = anchorAt(mk3(0.f, 5.f, 24.f)); scene_entity b = capsuleAt(mk3(0.f, 4.5f, 24.f), qIdentity(), 0.4f, 0.1f);
                cone_twist_constraint& c = getConstraint(s, addConeTwistConstraintFromGlobalPoints(base, b, O + mk3(0.f, 5.f, 24.f), mk3(0.f, -1.f, 0.f), deg2radf(40.f), deg2radf(20.f)));
                c.swingMotorType = constraint_velocity_motor; c.maxSwingMotorTorque = 30.f; c.swingMotorVelocityOrTargetAngle = 1.5f; c.swingMotorAxis = 0.7f;
                b.rigidBody().angularVelocity = mk3(0.f, 4.f, 0.f);          // twists into the twist limit
            }
            {   // J: cone-twist, swing position motor + twist position motor
                scene_entity base = anchorAt(mk3(0.f, 5.f, 27.f)); scene_entity b = capsuleAt(mk3(0.f, 4.5f, 27.f), qIdentity(), 0.4f, 0.1f);
                cone_twist_constraint& c = getConstraint(s, addConeTwistConstraintFromGlobalPoints(base, b, O + mk3(0.f, 5.f, 27.f), mk3(0.f, -1.f, 0.f), deg2radf(60.f), deg2radf(45.f)));
                c.swingMotorType = constraint_position_motor; c.maxSwingMotorTorque = 25.f; c.swingMotorVelocityOrTargetAngle = 0.4f; c.swingMotorAxis = 1.9f;
                c.twistMotorType = constraint_position_motor; c.maxTwistMotorTorque = 10.f; c.twistMotorVelocityOrTargetAngle = 0.3f;
            }
            {   // K: cone-twist, twist velocity motor, narrow swing limit hit by a sideways kick
                scene_entity base = anchorAt(mk3(0.f, 5.f, 30.f)); scene_entity b = capsuleAt(mk3(0.f, 4.5f, 30.f), qIdentity(), 0.4f, 0.1f);
                cone_twist_constraint& c = getConstraint(s, addConeTwistConstraintFromGlobalPoints(base, b, O + mk3(0.f, 5.f, 30.f), mk3(0.f, -1.f, 0.f), deg2radf(3.f), deg2radf(-1.f)));
                c.twistMotorType = constraint_velocity_motor; c.maxTwistMotorTorque = 5.f; c.twistMotorVelocityOrTargetAngle = 3.f;
                b.rigidBody().linearVelocity = mk3(2.f, 0.f, 1.f);
            }
        }
        s.createEntity("Ground").addTransform(mk3(0.f, -4.f, 0.f))
            .addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(0.f), mk3(nx * DX + 20.f, 4.f, nz * 40.f + 20.f)), kStaticMat));
        return 0;
    }
    if (config == 7)
    {   // nx x nz vehicles (vehicle.cpp:303-497) dropped on the ground; mode 1: engine and steering motors driven
        const float DX = spacing > 0.f ? spacing : 8.f;
        for (uint32_t ix = 0; ix < nx; ++ix) for (uint32_t iz = 0; iz < nz; ++iz)
        {
            vehicle v = vehicle::create(s, mk3(ix * DX, 1.1f + rng.between(0.f, 0.2f), iz * DX), (mode == 0) ? 0.f : rng.between(0.f, 6.2831853f));
            for (int i = 0; i < vehicle::numParts; ++i)
            {   // what addComponent<transform_component> does on a rigid body (scene.h:96-107): vehicle::initialize itself leaves the physics transforms unplaced
                const int rb = s.entities[v.parts[i].handle].rigidBody;
                s.physicsTransform0[rb] = s.physicsTransform1[rb] = v.parts[i].transform();
            }
            if (mode)
            {
                getConstraint(s, v.motorConstraint).motorVelocityOrTargetAngle = 6.f;
                getConstraint(s, v.steeringWheelConstraint).motorVelocityOrTargetAngle = 0.35f;
            }
        }
        s.createEntity("Ground").addTransform(mk3(0.f, -4.f, 0.f))
            .addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(0.f), mk3(nx * DX + 20.f, 4.f, nz * DX + 20.f)), kStaticMat));
        return 0;
    }
    if (config == 8)
    {   // ONE vehicle exactly as vehicle::create leaves it (physics transforms unplaced, see vehicle::initialize) over a ground box whose top
        // is at y = -1: the scene tests/test_ref_pins_oracle.py builds with the reference's own vehicle::create
        vehicle::create(s, mk3(3.f, 1.f, -2.f), 0.5f);
        s.createEntity("Ground").addTransform(mk3(0.f, -5.f, 0.f)).addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(0.f), mk3(28.f, 4.f, 28.f)), kStaticMat));
        return 0;
    }
    if (config == 4)
    {
        const float half = (nx - 1) * spacing * 0.5f, halfz = (nz - 1) * spacing * 0.5f;
        for (uint32_t ix = 0; ix < nx; ++ix) for (uint32_t iz = 0; iz < nz; ++iz)
        {
            const float y = 1.25f + ((mode == 0) ? rng.between(0.f, 0.5f) : 0.f);
            const float rot = (mode == 0) ? rng.between(0.f, 6.2831853f) : 0.f;
            humanoid_ragdoll::create(s, mk3(ix * spacing - half, y, iz * spacing - halfz), rot);
        }
        s.createEntity("Test ground").addTransform(mk3(0.f, -4.f, 0.f))
            .addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(0.f), mk3(fmaxf(20.f, half + 4.f), 4.f, fmaxf(20.f, halfz + 4.f))), kStaticMat));   // learned_locomotion.cpp:440-444
        return 0;
    }
    return 1;
}

// x-layers [ix0, ix1) of the config-3 mixed pile of nxTotal x ny x nz bodies + the static boxes of the WHOLE pile (replicated on every
// rank): the multi-GPU slab of the bench. The random stream is advanced for every lattice site, so the union over all slabs is exactly
// the scene b200host_build_config(3, nxTotal, ny, nz, ...) builds, whatever the number of ranks.
// addComponent<heightmap_collider_component>(chunksPerDim, chunkSize, material) + update(minCorner, amplitudeScale) + setHeights per chunk
int b200host_set_heightmap(void* sp, uint32_t chunksPerDim, float chunkSize, const float* minCorner3, float amplitudeScale, float restitution, float friction, const uint16_t* heights)
{
    game_scene& s = *(game_scene*)sp;
    s.heightmaps.clear();
    physics_material m; memset(&m, 0, sizeof(m)); m.restitution = restitution; m.friction = friction; m.density = 1.f;
    s.heightmaps.emplace_back(chunksPerDim, chunkSize, m);
    heightmap_collider_component& hm = s.heightmaps.back();
    hm.update(b200::mk3(minCorner3[0], minCorner3[1], minCorner3[2]), amplitudeScale);
    for (uint32_t z = 0; z < chunksPerDim; ++z) for (uint32_t x = 0; x < chunksPerDim; ++x) hm.setHeights(x, z, heights + (size_t)(z * chunksPerDim + x) * 129u * 129u);
    return 0;
}

// addComponent<cloth_component>(width, height, gridSizeX, gridSizeY, totalMass, ...) -> index; setWorldPositionOfFixedVertices; particle positions after the last step
uint32_t b200host_add_cloth(void* sp, float width, float height, uint32_t gx, uint32_t gy, float totalMass, float stiffness, float damping, float gravityFactor)
{ game_scene& s = *(game_scene*)sp; s.cloths.emplace_back(width, height, gx, gy, totalMass, stiffness, damping, gravityFactor); return (uint32_t)s.cloths.size() - 1; }
void b200host_cloth_set_fixed(void* sp, uint32_t cloth, const float* rot4, const float* pos3, int moveRigid)
{
    game_scene& s = *(game_scene*)sp; if (cloth >= s.cloths.size()) return;
    trs t; t.rotation = b200::mkq(rot4[0], rot4[1], rot4[2], rot4[3]); t.position = b200::mk3(pos3[0], pos3[1], pos3[2]); t.scale = b200::mk3(1.f);
    s.cloths[cloth].setWorldPositionOfFixedVertices(t, moveRigid != 0);
}
uint32_t b200host_cloth_positions(void* sp, uint32_t cloth, float* out3)
{
    game_scene& s = *(game_scene*)sp; if (cloth >= s.cloths.size()) return 0;
    const cloth_component& c = s.cloths[cloth];
    if (out3) memcpy(out3, c.positions.data(), sizeof(float) * 3 * c.positions.size());
    return (uint32_t)c.positions.size();
}

// binary entity stream (scene/serialization_binary.h): entity -> bytes; bytes -> a NEW entity of this scene (returns its handle, 0xFFFFFFFF on failure)
uint64_t b200host_serialize_entity(void* sp, uint32_t entity, uint8_t* out, uint64_t cap)
{ game_scene& s = *(game_scene*)sp; if (entity >= s.entities.size()) return 0; scene_entity e = { entity, &s }; return serializeEntityToMemory(e, out, cap); }
uint32_t b200host_deserialize_entity(void* sp, const uint8_t* bytes, uint64_t size)
{ game_scene& s = *(game_scene*)sp; scene_entity e = s.createEntity(""); return deserializeEntityFromMemory(e, bytes, size) ? e.handle : 0xFFFFFFFFu; }
uint32_t b200host_entity_of_body(void* sp, uint32_t body) { game_scene& s = *(game_scene*)sp; return body < s.rigidBodyEntity.size() ? s.rigidBodyEntity[body] : 0xFFFFFFFFu; }
uint32_t b200host_num_entities(void* sp) { return (uint32_t)((game_scene*)sp)->entities.size(); }

int b200host_build_mixed_slab(void* sp, uint32_t nxTotal, uint32_t ny, uint32_t nz, uint32_t ix0, uint32_t ix1, uint64_t seed, float spacing)
{
    game_scene& s = *(game_scene*)sp;
    Rng rng = { seed ? seed : 1 };
    addHullGeometries(s);
    const float hx = (nxTotal - 1) * spacing * 0.5f, hz = (nz - 1) * spacing * 0.5f;
    uint32_t counter = 0;
    for (uint32_t ix = 0; ix < nxTotal; ++ix) for (uint32_t iy = 0; iy < ny; ++iy) for (uint32_t iz = 0; iz < nz; ++iz, ++counter)
    {
        const vec3 p = mk3(ix * spacing - hx + rng.between(-0.005f, 0.005f), iy * spacing + 0.6f + rng.between(-0.005f, 0.005f), iz * spacing - hz + rng.between(-0.005f, 0.005f));
        const int kind = (int)(counter & 3u);
        quat q = qIdentity();
        if (kind != 0) q = randomRotation(rng);
        if (ix < ix0 || ix >= ix1) continue;
        scene_entity e = s.createEntity("Body");
        e.addTransform(p, q);
        if (kind == 0) e.addCollider(collider_component::asSphere({ mk3(0.f), 0.5f }, kDynamicMat));
        else if (kind == 1) e.addCollider(collider_component::asCapsule({ mk3(0.f, -0.25f, 0.f), mk3(0.f, 0.25f, 0.f), 0.25f }, kDynamicMat));
        else if (kind == 2) e.addCollider(collider_component::asOBB({ qIdentity(), mk3(0.f), mk3(0.4f, 0.3f, 0.35f) }, kDynamicMat));
        else e.addCollider(collider_component::asHull({ qIdentity(), mk3(0.f), (counter >> 2) & 1u }, kDynamicMat));
        e.addRigidBody(false, 1.f);
    }
    const float ex = hx + 1.1f, ez = hz + 1.1f, h = ny * spacing + 4.f;
    s.createEntity("Floor").addTransform(mk3(0.f, -4.f, 0.f)).addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(0.f), mk3(ex + 8.f, 4.f, ez + 8.f)), kStaticMat));
    addWalls(s, ex, ez, h);
    return 0;
}

uint32_t b200host_num_bodies(void* s) { return (uint32_t)((game_scene*)s)->rigidBodies.size(); }
uint32_t b200host_num_colliders(void* s) { return (uint32_t)((game_scene*)s)->colliders.size(); }
uint32_t b200host_num_joints(void* s, int kind) { return (uint32_t)((game_scene*)s)->constraintEntities[kind].size(); }
uint32_t b200host_num_hulls(void* s) { return (uint32_t)((game_scene*)s)->hullGeometries.size(); }
uint32_t b200host_num_hull_vertices(void* s) { uint32_t n = 0; for (auto& g : ((game_scene*)s)->hullGeometries) n += (uint32_t)g.vertices.size(); return n; }
uint32_t b200host_num_hull_triangles(void* s, uint32_t h) { return (uint32_t)((game_scene*)s)->hullGeometries[h].faces.size(); }

void b200host_get_bodies(void* s, b200phys_body* out) { std::vector<b200phys_body> b; ((game_scene*)s)->gatherBodies(b); if (!b.empty()) memcpy(out, b.data(), b.size() * sizeof(b200phys_body)); }
// returns the number of transforms written (capacity must be >= number of entities)
uint32_t b200host_get_colliders(void* s, b200phys_collider* outC, b200phys_transform* outT)
{
    std::vector<b200phys_collider> c; std::vector<b200phys_transform> t; ((game_scene*)s)->gatherColliders(c, t);
    if (!c.empty()) memcpy(outC, c.data(), c.size() * sizeof(b200phys_collider));
    if (!t.empty()) memcpy(outT, t.data(), t.size() * sizeof(b200phys_transform));
    return (uint32_t)t.size();
}
void b200host_get_joints(void* s, int kind, void* outStructs, b200phys_body_pair* outPairs)
{
    std::vector<uint8_t> st; std::vector<b200phys_body_pair> p; ((game_scene*)s)->gatherJoints(kind, st, p);
    if (!st.empty()) memcpy(outStructs, st.data(), st.size());
    if (!p.empty()) memcpy(outPairs, p.data(), p.size() * sizeof(b200phys_body_pair));
}
void b200host_get_hulls(void* s, float* verts, uint32_t* offsets)
{
    std::vector<float> v; std::vector<uint32_t> o; ((game_scene*)s)->gatherHulls(v, o);
    if (!v.empty()) memcpy(verts, v.data(), v.size() * 4);
    memcpy(offsets, o.data(), o.size() * 4);
}
void b200host_get_hull_triangles(void* s, uint32_t h, uint32_t* tris)
{
    const auto& f = ((game_scene*)s)->hullGeometries[h].faces;
    for (size_t i = 0; i < f.size(); ++i) { tris[3 * i] = f[i].a; tris[3 * i + 1] = f[i].b; tris[3 * i + 2] = f[i].c; }
}

// mass properties of one collider in the boundary packing (scene must hold the hull geometries): inertia(9), cog(3), mass
void b200host_mass_properties(void* s, const b200phys_collider* c, float* inertia9, float* cog3, float* mass)
{
    collider_component col; const float* f = c->shape;
    col.type = (collider_type)c->type; col.material = { physics_material_type_none, c->restitution, c->friction, c->density };
    switch (c->type)
    {
    case collider_type_sphere: col.sphere = { mk3(f[0], f[1], f[2]), f[3] }; break;
    case collider_type_capsule: col.capsule = { mk3(f[0], f[1], f[2]), mk3(f[3], f[4], f[5]), f[6] }; break;
    case collider_type_cylinder: col.cylinder = { mk3(f[0], f[1], f[2]), mk3(f[3], f[4], f[5]), f[6] }; break;
    case collider_type_aabb: col.aabb = bounding_box::fromMinMax(mk3(f[0], f[1], f[2]), mk3(f[3], f[4], f[5])); break;
    case collider_type_obb: col.obb = { mkq(f[0], f[1], f[2], f[3]), mk3(f[4], f[5], f[6]), mk3(f[7], f[8], f[9]) }; break;
    default: { uint32_t g; memcpy(&g, &f[7], 4); col.hull = { mkq(f[0], f[1], f[2], f[3]), mk3(f[4], f[5], f[6]), g }; } break;
    }
    physics_properties p = col.calculatePhysicsProperties(*(game_scene*)s);
    memcpy(inertia9, p.inertia.m, 36); cog3[0] = p.cog.x; cog3[1] = p.cog.y; cog3[2] = p.cog.z; *mass = p.mass;
}

// testPhysicsInteraction (physics.cpp:555-629): the pushed body's accumulators show up in b200host_get_bodies
void b200host_test_physics_interaction(void* s, const float* origin3, const float* direction3, float strength)
{
    testPhysicsInteraction(*(game_scene*)s, mk3(origin3[0], origin3[1], origin3[2]), mk3(direction3[0], direction3[1], direction3[2]), strength);
}

// scene files: the physics components of the reference's YAML scenes (scene_yaml.cpp); 0 on success
int b200host_save_yaml(void* s, const char* path)
{
    std::ofstream f(path);
    return f && saveSceneYaml(*(game_scene*)s, f, "b200host") ? 0 : 1;
}
int b200host_load_yaml(void* s, const char* path)       // entities created, -1 on failure
{
    std::ifstream f(path);
    return f ? loadSceneYaml(*(game_scene*)s, f) : -1;
}

// physicsStep(scene, arena, timer, settings, dt) through the host mirror (needs the CUDA library and a GPU)
int b200host_physics_step(void* s, float* timer, float dt, uint32_t iterations, int fixedFrameRate, uint32_t frameRate, int exactOrder)
{
    physics_settings st; st.fixedFrameRate = fixedFrameRate != 0; st.frameRate = frameRate; st.numRigidSolverIterations = iterations; st.exactConstraintOrder = exactOrder != 0;
    memory_arena arena;
    return physicsStep(*(game_scene*)s, arena, *timer, st, dt);
}
// same, with collision callbacks installed that log what they receive (tests): begin = entityA, entityB, position, normal,
// relative velocity (2 + 9 values); end = entityA, entityB. The logs are overwritten by every call.
static std::vector<uint32_t> gBeginEntities, gEndEntities; static std::vector<float> gBeginPayload;
int b200host_physics_step_with_events(void* s, float* timer, float dt, uint32_t iterations, int fixedFrameRate, uint32_t frameRate, int exactOrder)
{
    physics_settings st; st.fixedFrameRate = fixedFrameRate != 0; st.frameRate = frameRate; st.numRigidSolverIterations = iterations; st.exactConstraintOrder = exactOrder != 0;
    gBeginEntities.clear(); gEndEntities.clear(); gBeginPayload.clear();
    st.collisionBeginCallback = [](const collision_begin_event& e)
    {
        gBeginEntities.push_back(e.entityA); gBeginEntities.push_back(e.entityB);
        const float v[9] = { e.position.x, e.position.y, e.position.z, e.normal.x, e.normal.y, e.normal.z, e.relativeVelocity.x, e.relativeVelocity.y, e.relativeVelocity.z };
        gBeginPayload.insert(gBeginPayload.end(), v, v + 9);
    };
    st.collisionEndCallback = [](const collision_end_event& e) { gEndEntities.push_back(e.entityA); gEndEntities.push_back(e.entityB); };
    memory_arena arena;
    return physicsStep(*(game_scene*)s, arena, *timer, st, dt);
}
uint32_t b200host_event_log_counts(uint32_t* numEnd) { *numEnd = (uint32_t)gEndEntities.size() / 2; return (uint32_t)gBeginEntities.size() / 2; }
void b200host_event_log_get(uint32_t* beginEntities, float* beginPayload9, uint32_t* endEntities)
{
    if (!gBeginEntities.empty()) { memcpy(beginEntities, gBeginEntities.data(), 4 * gBeginEntities.size()); memcpy(beginPayload9, gBeginPayload.data(), 4 * gBeginPayload.size()); }
    if (!gEndEntities.empty()) memcpy(endEntities, gEndEntities.data(), 4 * gEndEntities.size());
}
// entity (render) transforms of the rigid bodies after physicsStep: rotation(4) + position(3) per body
void b200host_get_render_transforms(void* sp, float* out7)
{
    game_scene& s = *(game_scene*)sp;
    for (size_t i = 0; i < s.rigidBodies.size(); ++i)
    {
        const trs& t = s.entities[s.rigidBodyEntity[i]].transform; float* o = out7 + 7 * i;
        o[0] = t.rotation.x; o[1] = t.rotation.y; o[2] = t.rotation.z; o[3] = t.rotation.w; o[4] = t.position.x; o[5] = t.position.y; o[6] = t.position.z;
    }
}
void b200host_get_stats(void* sp, b200phys_stats* out) { *out = ((game_scene*)sp)->lastStats; }

} // extern "C"
