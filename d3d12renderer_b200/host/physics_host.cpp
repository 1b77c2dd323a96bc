// physics_host.cpp -- implementation of the host-side mirror (see physics_host.h). Scene bookkeeping, mass properties
// (scene-build time), joint builders, the ragdoll generator and the physicsStep() driver. The per-step physics itself runs on
// the GPU through the C ABI; this file contains no CPU implementation of any step stage.
#include "physics_host.h"
#include <algorithm>
#include <cstring>
#include <cstdio>

using namespace b200;

namespace b200host {

// ------------------------------------------------------------------------------------------------- rigid bodies
rigid_body_component::rigid_body_component(bool kinematic, float gravityFactor_, float linearDamping_, float angularDamping_)
{   // rigid_body.cpp:6-27
    if (kinematic) { invMass = 0.f; invInertia = m3Zero(); } else { invMass = 1.f; invInertia = m3Identity(); }
    gravityFactor = gravityFactor_; linearDamping = linearDamping_; angularDamping = angularDamping_;
    localCOGPosition = mk3(0.f); linearVelocity = mk3(0.f); angularVelocity = mk3(0.f); forceAccumulator = mk3(0.f); torqueAccumulator = mk3(0.f);
}

// ------------------------------------------------------------------------------------------------ mass properties
// collider_union::calculatePhysicsProperties (physics.cpp:1416-1588): inertia about the shape's own centre of gravity
physics_properties collider_union::calculatePhysicsProperties(const game_scene& scene) const
{
    physics_properties result; result.inertia = m3Zero(); result.cog = mk3(0.f); result.mass = 0.f;
    switch (type)
    {
    case collider_type_sphere:
    {
        float sq = sphere.radius * sphere.radius, sqPi = B200_PI * sq;
        float volume = 4.f / 3.f * sqPi * sphere.radius;   // bounding_volumes.h:35-41
        result.mass = volume * material.density; result.cog = sphere.center;
        result.inertia = m3Identity() * (2.f / 5.f * result.mass * sphere.radius * sphere.radius);
    } break;
    case collider_type_capsule:
    {
        vec3 axis = capsule.positionA - capsule.positionB;
        if (axis.y < 0.f) axis *= -1.f;
        float height = len(axis);
        axis *= (1.f / height);
        mat3 rot = quaternionToMat3(rotateFromTo(mk3(0.f, 1.f, 0.f), axis));
        float sq = capsule.radius * capsule.radius, sqPi = B200_PI * sq;
        float volume = 4.f / 3.f * sqPi * capsule.radius + sqPi * len(capsule.positionA - capsule.positionB);   // bounding_volumes.h:50-58
        result.mass = volume * material.density; result.cog = (capsule.positionA + capsule.positionB) * 0.5f;
        float cylMass = material.density * sqPi * height;
        float hemiMass = material.density * 2.f / 3.f * sqPi * capsule.radius;
        float sqH = height * height;
        mat3 I = m3Zero();
        matRef(I, 1, 1) = sq * cylMass * 0.5f;
        matRef(I, 0, 0) = matRef(I, 2, 2) = mat(I, 1, 1) * 0.5f + cylMass * sqH / 12.f;
        float t0 = hemiMass * 2.f * sq / 5.f;
        matRef(I, 1, 1) += t0 * 2.f;
        float t1 = height * 0.5f;
        float t2 = t0 + hemiMass * (t1 * t1 + 3.f / 8.f * sqH);
        matRef(I, 0, 0) += t2 * 2.f; matRef(I, 2, 2) += t2 * 2.f;
        result.inertia = transpose(rot) * I * rot;
    } break;
    case collider_type_cylinder:
    {
        vec3 axis = cylinder.positionA - cylinder.positionB;
        if (axis.y < 0.f) axis *= -1.f;
        float height = len(axis);
        axis *= (1.f / height);
        mat3 rot = quaternionToMat3(rotateFromTo(mk3(0.f, 1.f, 0.f), axis));
        float volume = B200_PI * cylinder.radius * cylinder.radius * len(cylinder.positionA - cylinder.positionB);   // bounding_volumes.h:67-73
        result.mass = volume * material.density; result.cog = (cylinder.positionA + cylinder.positionB) * 0.5f;
        float sq = cylinder.radius * cylinder.radius, sqH = height * height;
        mat3 I = m3Zero();
        matRef(I, 1, 1) = sq * result.mass * 0.5f;
        matRef(I, 0, 0) = matRef(I, 2, 2) = 1.f / 12.f * result.mass * (3.f * sq + sqH);
        result.inertia = transpose(rot) * I * rot;
    } break;
    case collider_type_aabb:
    {
        vec3 dm = aabb.maxCorner - aabb.minCorner;
        result.mass = dm.x * dm.y * dm.z * material.density; result.cog = (aabb.minCorner + aabb.maxCorner) * 0.5f;
        vec3 d = ((aabb.maxCorner - aabb.minCorner) * 0.5f) * 2.f;
        matRef(result.inertia, 0, 0) = 1.f / 12.f * result.mass * (d.y * d.y + d.z * d.z);
        matRef(result.inertia, 1, 1) = 1.f / 12.f * result.mass * (d.x * d.x + d.z * d.z);
        matRef(result.inertia, 2, 2) = 1.f / 12.f * result.mass * (d.x * d.x + d.y * d.y);
    } break;
    case collider_type_obb:
    {
        vec3 d = obb.radius * 2.f;
        result.mass = d.x * d.y * d.z * material.density; result.cog = obb.center;
        mat3 I = m3Zero();
        matRef(I, 0, 0) = 1.f / 12.f * result.mass * (d.y * d.y + d.z * d.z);
        matRef(I, 1, 1) = 1.f / 12.f * result.mass * (d.x * d.x + d.z * d.z);
        matRef(I, 2, 2) = 1.f / 12.f * result.mass * (d.x * d.x + d.y * d.y);
        mat3 rot = quaternionToMat3(obb.rotation);
        result.inertia = transpose(rot) * I * rot;
    } break;
    default:
    {   // hull: accumulate the covariance of the tetrahedra (origin, face), physics.cpp:1520-1580
        const bounding_hull_geometry& geom = scene.hullGeometries[hull.geometryIndex];
        const float s60 = 1.f / 60.f, s120 = 1.f / 120.f;
        mat3 Cc; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) matRef(Cc, i, j) = (i == j) ? s60 : s120;
        float totalMass = 0.f; mat3 totalCov = m3Zero(); vec3 totalCog = mk3(0.f);
        for (const indexed_triangle32& f : geom.faces)
        {
            vec3 w1 = hull.position + hull.rotation * geom.vertices[f.a], w2 = hull.position + hull.rotation * geom.vertices[f.b], w3 = hull.position + hull.rotation * geom.vertices[f.c];
            mat3 A;
            matRef(A, 0, 0) = w1.x; matRef(A, 0, 1) = w2.x; matRef(A, 0, 2) = w3.x;
            matRef(A, 1, 0) = w1.y; matRef(A, 1, 1) = w2.y; matRef(A, 1, 2) = w3.y;
            matRef(A, 2, 0) = w1.z; matRef(A, 2, 1) = w2.z; matRef(A, 2, 2) = w3.z;
            float detA = determinant(A);
            mat3 cov = (A * detA) * Cc * transpose(A);
            float volume = 1.f / 6.f * detA;
            totalMass += volume; totalCov = totalCov + cov; totalCog += ((w1 + w2 + w3) * 0.25f) * volume;
        }
        totalCog /= totalMass;
        mat3 Cp = totalCov - outerProduct(totalCog, totalCog) * totalMass;
        result.cog = totalCog; result.mass = totalMass * material.density;
        result.inertia = (m3Identity() * trace(Cp) - Cp) * material.density;
    } break;
    }
    return result;
}

// rigid_body_component::recalculateProperties (rigid_body.cpp:29-81)
void game_scene::recalculateProperties(entity_handle e)
{
    entity_data& ed = entities[e];
    if (ed.rigidBody < 0) return;
    rigid_body_component& rb = rigidBodies[ed.rigidBody];
    if (rb.invMass == 0.f || ed.numColliders == 0) return;
    std::vector<physics_properties> props;
    for (entity_handle c = ed.firstColliderEntity; c != null_entity; c = colliders[c].nextEntity) props.push_back(colliders[c].calculatePhysicsProperties(*this));
    mat3 inertia = m3Zero(); vec3 cog = mk3(0.f); float mass = 0.f;
    for (const physics_properties& p : props) { mass += p.mass; cog += p.cog * p.mass; }
    rb.invMass = 1.f / mass;
    rb.localCOGPosition = cog = cog * rb.invMass;
    for (const physics_properties& p : props)
    {
        vec3 r = p.cog - cog;
        inertia = inertia + (p.inertia + (m3Identity() * dot(r, r) - outerProduct(r, r)) * p.mass);
    }
    rb.invInertia = invert(inertia);
}

// ----------------------------------------------------------------------------------------------------- entities
scene_entity game_scene::createEntity(const char* name)
{
    entity_data d; d.name = name ? name : "";
    entities.push_back(d);
    scene_entity e; e.handle = (entity_handle)(entities.size() - 1); e.scene = this;
    return e;
}
uint32_t game_scene::allocateBoundingHullGeometry(const vec3* vertices, uint32_t numVertices, const indexed_triangle32* triangles, uint32_t numTriangles)
{
    bounding_hull_geometry g; g.vertices.assign(vertices, vertices + numVertices); g.faces.assign(triangles, triangles + numTriangles);
    hullGeometries.push_back(g); hullsDirty = true;
    return (uint32_t)hullGeometries.size() - 1;
}
scene_entity& scene_entity::addTransform(vec3 position, quat rotation, vec3 scale)
{   // scene.h:96-107 : refreshes physics_transform0/1 when the entity already is a rigid body
    game_scene::entity_data& d = scene->entities[handle];
    d.hasTransform = true; d.transform = makeTrs(position, rotation, scale);
    if (d.rigidBody >= 0) { scene->physicsTransform0[d.rigidBody] = d.transform; scene->physicsTransform1[d.rigidBody] = d.transform; }
    scene->collidersDirty = true;
    return *this;
}
scene_entity& scene_entity::addCollider(const collider_component& c)
{   // scene.h:38-63 : child collider entity, prepended to the parent's list, mass properties recomputed
    game_scene::entity_data& d = scene->entities[handle];
    collider_component col = c; col.parentEntity = handle; col.nextEntity = d.firstColliderEntity;
    scene->colliders.push_back(col);
    d.firstColliderEntity = (entity_handle)(scene->colliders.size() - 1);
    ++d.numColliders;
    scene->recalculateProperties(handle);
    scene->collidersDirty = true;
    return *this;
}
scene_entity& scene_entity::addTrigger(std::function<void(const trigger_event&)> callback)
{
    game_scene::entity_data& d = scene->entities[handle];
    if (d.trigger < 0) { scene->triggers.push_back(trigger_component{ callback }); scene->triggerEntity.push_back(handle); d.trigger = (int32_t)scene->triggers.size() - 1; }
    else scene->triggers[d.trigger].callback = callback;
    scene->collidersDirty = true;
    return *this;
}
scene_entity& scene_entity::addForceField(vec3 force)
{
    game_scene::entity_data& d = scene->entities[handle];
    if (d.forceField < 0) { scene->forceFields.push_back(force_field_component{ force }); scene->forceFieldEntity.push_back(handle); d.forceField = (int32_t)scene->forceFields.size() - 1; }
    else scene->forceFields[d.forceField].force = force;
    scene->collidersDirty = true;
    return *this;
}
scene_entity& scene_entity::addRigidBody(bool kinematic, float gravityFactor, float linearDamping, float angularDamping)
{   // scene.h:69-84
    game_scene::entity_data& d = scene->entities[handle];
    if (d.rigidBody < 0)
    {
        scene->rigidBodies.emplace_back(kinematic, gravityFactor, linearDamping, angularDamping);
        scene->rigidBodyEntity.push_back(handle);
        scene->physicsTransform0.push_back(d.transform); scene->physicsTransform1.push_back(d.transform);
        d.rigidBody = (int32_t)scene->rigidBodies.size() - 1;
    }
    else scene->rigidBodies[d.rigidBody] = rigid_body_component(kinematic, gravityFactor, linearDamping, angularDamping);
    if (!d.hasTransform) { d.transform = makeTrs(mk3(0.f)); }
    scene->physicsTransform0[d.rigidBody] = d.transform; scene->physicsTransform1[d.rigidBody] = d.transform;
    scene->recalculateProperties(handle);
    scene->jointPairsDirty = true;      // joints of this entity now resolve to a body index
    scene->collidersDirty = true;
    return *this;
}
rigid_body_component& scene_entity::rigidBody() { return scene->rigidBodies[scene->entities[handle].rigidBody]; }
trs& scene_entity::transform() { return scene->entities[handle].transform; }

game_scene::~game_scene()
{
    if (stageBodies) b200phys_host_free(stageBodies);
    if (stageStates) b200phys_host_free(stageStates);
    if (ctx) { b200phys_destroy(ctx); ctx = nullptr; }
}

// ------------------------------------------------------------------------------------------------------- gather
static void st3(float* o, vec3 v) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }
static void st4(float* o, quat q) { o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w; }

void game_scene::gatherBodies(std::vector<b200phys_body>& out) const { out.resize(rigidBodies.size()); gatherBodies(out.data()); }
void game_scene::gatherBodies(b200phys_body* out) const
{   // rbGlobal[] is filled in rigid-body storage order (physics.cpp:1269-1274)
    parallelFor((uint32_t)rigidBodies.size(), [&](uint32_t i)
    {
        const rigid_body_component& rb = rigidBodies[i]; const trs& t = physicsTransform1[i]; b200phys_body& b = out[i];
        st4(b.rotation, t.rotation); st3(b.position, t.position); st3(b.localCOGPosition, rb.localCOGPosition);
        b.invMass = rb.invMass; memcpy(b.invInertia, rb.invInertia.m, sizeof(b.invInertia));
        b.gravityFactor = rb.gravityFactor; b.linearDamping = rb.linearDamping; b.angularDamping = rb.angularDamping;
        st3(b.linearVelocity, rb.linearVelocity); st3(b.angularVelocity, rb.angularVelocity);
        st3(b.forceAccumulator, rb.forceAccumulator); st3(b.torqueAccumulator, rb.torqueAccumulator);
    });
}

void game_scene::gatherColliders(std::vector<b200phys_collider>& outC, std::vector<b200phys_transform>& outT) const
{   // world collider index = position in EnTT's back-to-front iteration = reverse creation order (physics.cpp:635-641)
    const size_t C = colliders.size();
    outC.resize(C); outT.clear();
    std::vector<int32_t> transformOf(entities.size(), -1);
    for (size_t i = 0; i < C; ++i)
    {
        const collider_component& c = colliders[C - 1 - i];
        b200phys_collider& o = outC[i]; memset(&o, 0, sizeof(o));
        switch (c.type)
        {
        case collider_type_sphere: st3(o.shape, c.sphere.center); o.shape[3] = c.sphere.radius; break;
        case collider_type_capsule: st3(o.shape, c.capsule.positionA); st3(o.shape + 3, c.capsule.positionB); o.shape[6] = c.capsule.radius; break;
        case collider_type_cylinder: st3(o.shape, c.cylinder.positionA); st3(o.shape + 3, c.cylinder.positionB); o.shape[6] = c.cylinder.radius; break;
        case collider_type_aabb: st3(o.shape, c.aabb.minCorner); st3(o.shape + 3, c.aabb.maxCorner); break;
        case collider_type_obb: st4(o.shape, c.obb.rotation); st3(o.shape + 4, c.obb.center); st3(o.shape + 7, c.obb.radius); break;
        default: st4(o.shape, c.hull.rotation); st3(o.shape + 4, c.hull.position); memcpy(&o.shape[7], &c.hull.geometryIndex, 4); break;
        }
        o.restitution = c.material.restitution; o.friction = c.material.friction; o.density = c.material.density;
        o.type = (uint8_t)c.type;
        const entity_data& parent = entities[c.parentEntity];
        if (parent.rigidBody >= 0) { o.objectType = B200PHYS_OBJ_RIGID_BODY; o.objectIndex = (uint32_t)parent.rigidBody; o.parentTransform = B200PHYS_NO_TRANSFORM; }
        else
        {   // force field > trigger > static collider (physics.cpp:657-671); pose = transform_component or identity (:645-647)
            if (parent.forceField >= 0) { o.objectType = B200PHYS_OBJ_FORCE_FIELD; o.objectIndex = (uint32_t)parent.forceField; }
            else if (parent.trigger >= 0) { o.objectType = B200PHYS_OBJ_TRIGGER; o.objectIndex = (uint32_t)parent.trigger; }
            else { o.objectType = B200PHYS_OBJ_STATIC; o.objectIndex = 0; }
            if (!parent.hasTransform) o.parentTransform = B200PHYS_NO_TRANSFORM;
            else
            {
                if (transformOf[c.parentEntity] < 0)
                {
                    b200phys_transform t; st4(t.rotation, parent.transform.rotation); st3(t.position, parent.transform.position); t.reserved = 0.f;
                    outT.push_back(t); transformOf[c.parentEntity] = (int32_t)outT.size() - 1;
                }
                o.parentTransform = (uint32_t)transformOf[c.parentEntity];
            }
        }
    }
}

void game_scene::gatherJoints(int kind, std::vector<uint8_t>& outStructs, std::vector<b200phys_body_pair>& outPairs) const
{   // constraint structs in raw storage order, body indices = rigid body storage indices (physics.cpp:789-805)
    const void* src = nullptr; size_t n = 0, sz = 0;
    switch (kind)
    {
    case 0: src = distanceConstraints.data(); n = distanceConstraints.size(); sz = sizeof(distance_constraint); break;
    case 1: src = ballConstraints.data(); n = ballConstraints.size(); sz = sizeof(ball_constraint); break;
    case 2: src = fixedConstraints.data(); n = fixedConstraints.size(); sz = sizeof(fixed_constraint); break;
    case 3: src = hingeConstraints.data(); n = hingeConstraints.size(); sz = sizeof(hinge_constraint); break;
    case 4: src = coneTwistConstraints.data(); n = coneTwistConstraints.size(); sz = sizeof(cone_twist_constraint); break;
    default: src = sliderConstraints.data(); n = sliderConstraints.size(); sz = sizeof(slider_constraint); break;
    }
    outStructs.assign((const uint8_t*)src, (const uint8_t*)src + n * sz);
    outPairs.resize(n);
    for (size_t i = 0; i < n; ++i)
    {
        outPairs[i].rbA = (uint32_t)entities[constraintEntities[kind][i].rbA].rigidBody;
        outPairs[i].rbB = (uint32_t)entities[constraintEntities[kind][i].rbB].rigidBody;
    }
}

const void* game_scene::jointStructs(int kind, size_t& n) const
{
    switch (kind)
    {
    case 0: n = distanceConstraints.size(); return distanceConstraints.data();
    case 1: n = ballConstraints.size(); return ballConstraints.data();
    case 2: n = fixedConstraints.size(); return fixedConstraints.data();
    case 3: n = hingeConstraints.size(); return hingeConstraints.data();
    case 4: n = coneTwistConstraints.size(); return coneTwistConstraints.data();
    default: n = sliderConstraints.size(); return sliderConstraints.data();
    }
}
void game_scene::refreshJointPairs()
{   // body indices = rigid body storage indices (physics.cpp:789-805); they only change when joints or bodies are added
    for (int kind = 0; kind < 6; ++kind)
    {
        const size_t n = constraintEntities[kind].size();
        jointPairs[kind].resize(n);
        for (size_t i = 0; i < n; ++i)
        {
            jointPairs[kind][i].rbA = (uint32_t)entities[constraintEntities[kind][i].rbA].rigidBody;
            jointPairs[kind][i].rbB = (uint32_t)entities[constraintEntities[kind][i].rbB].rigidBody;
        }
    }
    jointPairsDirty = false;
}

void game_scene::gatherHulls(std::vector<float>& verts, std::vector<uint32_t>& offsets) const
{
    verts.clear(); offsets.assign(1, 0u);
    for (const bounding_hull_geometry& g : hullGeometries)
    {
        for (const vec3& v : g.vertices) { verts.push_back(v.x); verts.push_back(v.y); verts.push_back(v.z); }
        offsets.push_back((uint32_t)(verts.size() / 3));
    }
}

void game_scene::scatterBodyStates(const b200phys_body_state* s, uint32_t count)
{   // integrateVelocity's outputs (rigid_body.cpp:126-142): physics_transform1 + velocities; accumulators cleared
    const uint32_t n = count < (uint32_t)rigidBodies.size() ? count : (uint32_t)rigidBodies.size();
    parallelFor(n, [&](uint32_t i)
    {
        rigid_body_component& rb = rigidBodies[i]; trs& t = physicsTransform1[i];
        t.rotation = mkq(s[i].rotation[0], s[i].rotation[1], s[i].rotation[2], s[i].rotation[3]);
        t.position = mk3(s[i].position[0], s[i].position[1], s[i].position[2]);
        rb.linearVelocity = mk3(s[i].linearVelocity[0], s[i].linearVelocity[1], s[i].linearVelocity[2]);
        rb.angularVelocity = mk3(s[i].angularVelocity[0], s[i].angularVelocity[1], s[i].angularVelocity[2]);
        rb.forceAccumulator = mk3(0.f); rb.torqueAccumulator = mk3(0.f);
    });
}

// ------------------------------------------------------------------------------------------------------- joints
static vec3 inverseTransformPosition(const trs& m, vec3 pos) { vec3 p = conjugate(m.rotation) * (pos - m.position); return mk3(p.x / m.scale.x, p.y / m.scale.y, p.z / m.scale.z); } // math.cpp:528-531
static vec3 inverseTransformDirection(const trs& m, vec3 dir) { return conjugate(m.rotation) * dir; }

static uint32_t pushPair(game_scene& s, int kind, scene_entity& a, scene_entity& b)
{   // returns the stable id of the new joint (the constraint entity of the reference, physics.cpp:87-126)
    s.constraintEntities[kind].push_back({ a.handle, b.handle }); s.jointsDirty = true; s.jointPairsDirty = true;
    const uint32_t id = (uint32_t)s.constraintIndexOfId[kind].size();
    s.constraintIds[kind].push_back(id); s.constraintIndexOfId[kind].push_back((int32_t)s.constraintEntities[kind].size() - 1);
    return id;
}

// addConstraint (physics.cpp:344-378): the struct is stored as given
distance_constraint_handle addConstraint(scene_entity& a, scene_entity& b, const distance_constraint& c) { a.scene->distanceConstraints.push_back(c); return { pushPair(*a.scene, 0, a, b) }; }
ball_constraint_handle addConstraint(scene_entity& a, scene_entity& b, const ball_constraint& c) { a.scene->ballConstraints.push_back(c); return { pushPair(*a.scene, 1, a, b) }; }
fixed_constraint_handle addConstraint(scene_entity& a, scene_entity& b, const fixed_constraint& c) { a.scene->fixedConstraints.push_back(c); return { pushPair(*a.scene, 2, a, b) }; }
hinge_constraint_handle addConstraint(scene_entity& a, scene_entity& b, const hinge_constraint& c) { a.scene->hingeConstraints.push_back(c); return { pushPair(*a.scene, 3, a, b) }; }
cone_twist_constraint_handle addConstraint(scene_entity& a, scene_entity& b, const cone_twist_constraint& c) { a.scene->coneTwistConstraints.push_back(c); return { pushPair(*a.scene, 4, a, b) }; }
slider_constraint_handle addConstraint(scene_entity& a, scene_entity& b, const slider_constraint& c) { a.scene->sliderConstraints.push_back(c); return { pushPair(*a.scene, 5, a, b) }; }

// getConstraint (physics.cpp:380-408). Joint parameters are host-mutable between steps (motor targets): the step re-uploads them.
static size_t indexOf(game_scene& s, int kind, uint32_t id) { return (size_t)s.constraintIndexOfId[kind].at(id); }
distance_constraint& getConstraint(game_scene& s, distance_constraint_handle h) { s.jointsDirty = true; return s.distanceConstraints.at(indexOf(s, 0, h.entity)); }
ball_constraint& getConstraint(game_scene& s, ball_constraint_handle h) { s.jointsDirty = true; return s.ballConstraints.at(indexOf(s, 1, h.entity)); }
fixed_constraint& getConstraint(game_scene& s, fixed_constraint_handle h) { s.jointsDirty = true; return s.fixedConstraints.at(indexOf(s, 2, h.entity)); }
hinge_constraint& getConstraint(game_scene& s, hinge_constraint_handle h) { s.jointsDirty = true; return s.hingeConstraints.at(indexOf(s, 3, h.entity)); }
cone_twist_constraint& getConstraint(game_scene& s, cone_twist_constraint_handle h) { s.jointsDirty = true; return s.coneTwistConstraints.at(indexOf(s, 4, h.entity)); }
slider_constraint& getConstraint(game_scene& s, slider_constraint_handle h) { s.jointsDirty = true; return s.sliderConstraints.at(indexOf(s, 5, h.entity)); }

// deleteConstraint (physics.cpp:410-520): the pools are EnTT storages, destroying the constraint entity swaps the last element into the
// hole; the same happens here so that joint array order (= Gauss-Seidel order of the scalar solver) stays what the reference would have
template <typename T> static void swapAndPop(std::vector<T>& v, size_t i) { v[i] = v.back(); v.pop_back(); }
static void eraseJoint(game_scene& s, int kind, uint32_t id)
{
    if (id >= s.constraintIndexOfId[kind].size() || s.constraintIndexOfId[kind][id] < 0) return;
    const size_t i = (size_t)s.constraintIndexOfId[kind][id], last = s.constraintEntities[kind].size() - 1;
    switch (kind)
    {
    case 0: swapAndPop(s.distanceConstraints, i); break;
    case 1: swapAndPop(s.ballConstraints, i); break;
    case 2: swapAndPop(s.fixedConstraints, i); break;
    case 3: swapAndPop(s.hingeConstraints, i); break;
    case 4: swapAndPop(s.coneTwistConstraints, i); break;
    default: swapAndPop(s.sliderConstraints, i); break;
    }
    const uint32_t movedId = s.constraintIds[kind][last];
    swapAndPop(s.constraintEntities[kind], i); swapAndPop(s.constraintIds[kind], i);
    s.constraintIndexOfId[kind][id] = -1;
    if (i != last) s.constraintIndexOfId[kind][movedId] = (int32_t)i;
    s.jointsDirty = true; s.jointPairsDirty = true;
}
void deleteConstraint(game_scene& s, distance_constraint_handle h) { eraseJoint(s, 0, h.entity); }
void deleteConstraint(game_scene& s, ball_constraint_handle h) { eraseJoint(s, 1, h.entity); }
void deleteConstraint(game_scene& s, fixed_constraint_handle h) { eraseJoint(s, 2, h.entity); }
void deleteConstraint(game_scene& s, hinge_constraint_handle h) { eraseJoint(s, 3, h.entity); }
void deleteConstraint(game_scene& s, cone_twist_constraint_handle h) { eraseJoint(s, 4, h.entity); }
void deleteConstraint(game_scene& s, slider_constraint_handle h) { eraseJoint(s, 5, h.entity); }
void deleteAllConstraints(game_scene& s)
{   // physics.cpp:410-418
    s.distanceConstraints.clear(); s.ballConstraints.clear(); s.fixedConstraints.clear(); s.hingeConstraints.clear(); s.coneTwistConstraints.clear(); s.sliderConstraints.clear();
    for (int k = 0; k < 6; ++k) { s.constraintEntities[k].clear(); s.constraintIds[k].clear(); for (int32_t& x : s.constraintIndexOfId[k]) x = -1; }
    s.jointsDirty = true; s.jointPairsDirty = true;
}
void deleteAllConstraintsFromEntity(scene_entity& e)
{   // physics.cpp:522-552: every joint with the entity at either end
    game_scene& s = *e.scene;
    for (int k = 0; k < 6; ++k)
        for (size_t i = s.constraintEntities[k].size(); i-- > 0;)
            if (s.constraintEntities[k][i].rbA == e.handle || s.constraintEntities[k][i].rbB == e.handle) eraseJoint(s, k, s.constraintIds[k][i]);
}
constraint_entity_iterator::constraint_entity_iterator(scene_entity e)
{
    if (!e) return;
    game_scene& s = *e.scene;
    for (int k = 0; k < 6; ++k)
        for (size_t i = 0; i < s.constraintEntities[k].size(); ++i)
        {
            const b200phys_body_pair& p = s.constraintEntities[k][i];
            if (p.rbA != e.handle && p.rbB != e.handle) continue;
            scene_entity other; other.scene = e.scene; other.handle = (p.rbA == e.handle) ? p.rbB : p.rbA;     // getOtherEntity, physics.h:266-270
            items.push_back({ other, (constraint_type)k, (uint32_t)i });
        }
}

distance_constraint_handle addDistanceConstraintFromLocalPoints(scene_entity& a, scene_entity& b, vec3 localAnchorA, vec3 localAnchorB, float distance)
{   // physics.cpp:128-146
    distance_constraint c; st3(c.localAnchorA, localAnchorA); st3(c.localAnchorB, localAnchorB); c.globalLength = distance;
    a.scene->distanceConstraints.push_back(c);
    return { pushPair(*a.scene, 0, a, b) };
}
distance_constraint_handle addDistanceConstraintFromGlobalPoints(scene_entity& a, scene_entity& b, vec3 globalAnchorA, vec3 globalAnchorB)
{   // physics.cpp:148-157
    return addDistanceConstraintFromLocalPoints(a, b, inverseTransformPosition(a.transform(), globalAnchorA), inverseTransformPosition(b.transform(), globalAnchorB), len(globalAnchorA - globalAnchorB));
}
ball_constraint_handle addBallConstraintFromLocalPoints(scene_entity& a, scene_entity& b, vec3 localAnchorA, vec3 localAnchorB)
{   // physics.cpp:159-176
    ball_constraint c; st3(c.localAnchorA, localAnchorA); st3(c.localAnchorB, localAnchorB);
    a.scene->ballConstraints.push_back(c);
    return { pushPair(*a.scene, 1, a, b) };
}
ball_constraint_handle addBallConstraintFromGlobalPoints(scene_entity& a, scene_entity& b, vec3 globalAnchor)
{   // physics.cpp:178-186
    return addBallConstraintFromLocalPoints(a, b, inverseTransformPosition(a.transform(), globalAnchor), inverseTransformPosition(b.transform(), globalAnchor));
}
fixed_constraint_handle addFixedConstraintFromGlobalPoints(scene_entity& a, scene_entity& b, vec3 globalAnchor)
{   // physics.cpp:207-230
    fixed_constraint c; memset(&c, 0, sizeof(c));
    const trs& tA = a.transform(); const trs& tB = b.transform();
    st3(c.localAnchorA, inverseTransformPosition(tA, globalAnchor)); st3(c.localAnchorB, inverseTransformPosition(tB, globalAnchor));
    st4(c.initialInvRotationDifference, conjugate(tB.rotation) * tA.rotation);
    a.scene->fixedConstraints.push_back(c);
    return { pushPair(*a.scene, 2, a, b) };
}
hinge_constraint_handle addHingeConstraintFromGlobalPoints(scene_entity& a, scene_entity& b, vec3 globalAnchor, vec3 globalHingeAxis, float minLimit, float maxLimit)
{   // physics.cpp:232-268
    hinge_constraint c; memset(&c, 0, sizeof(c));
    const trs& tA = a.transform(); const trs& tB = b.transform();
    st3(c.localAnchorA, inverseTransformPosition(tA, globalAnchor)); st3(c.localAnchorB, inverseTransformPosition(tB, globalAnchor));
    vec3 axisA = inverseTransformDirection(tA, globalHingeAxis), axisB = inverseTransformDirection(tB, globalHingeAxis);
    st3(c.localHingeAxisA, axisA); st3(c.localHingeAxisB, axisB);
    vec3 tanA, bitanA; getTangents(axisA, tanA, bitanA);
    st3(c.localHingeTangentA, tanA); st3(c.localHingeBitangentA, bitanA);
    st3(c.localHingeTangentB, conjugate(tB.rotation) * (tA.rotation * tanA));
    c.minRotationLimit = minLimit; c.maxRotationLimit = maxLimit;
    c.motorType = constraint_velocity_motor; c.motorVelocityOrTargetAngle = 0.f; c.maxMotorTorque = -1.f;
    a.scene->hingeConstraints.push_back(c);
    return { pushPair(*a.scene, 3, a, b) };
}
cone_twist_constraint_handle addConeTwistConstraintFromGlobalPoints(scene_entity& a, scene_entity& b, vec3 globalAnchor, vec3 globalAxis, float swingLimit, float twistLimit)
{   // physics.cpp:270-310
    cone_twist_constraint c; memset(&c, 0, sizeof(c));
    const trs& tA = a.transform(); const trs& tB = b.transform();
    st3(c.localAnchorA, inverseTransformPosition(tA, globalAnchor)); st3(c.localAnchorB, inverseTransformPosition(tB, globalAnchor));
    c.swingLimit = swingLimit; c.twistLimit = twistLimit;
    vec3 axisA = inverseTransformDirection(tA, globalAxis), axisB = inverseTransformDirection(tB, globalAxis);
    st3(c.localLimitAxisA, axisA); st3(c.localLimitAxisB, axisB);
    vec3 tanA, bitanA; getTangents(axisA, tanA, bitanA);
    st3(c.localLimitTangentA, tanA); st3(c.localLimitBitangentA, bitanA);
    st3(c.localLimitTangentB, conjugate(tB.rotation) * (tA.rotation * tanA));
    c.swingMotorType = constraint_velocity_motor; c.swingMotorVelocityOrTargetAngle = 0.f; c.maxSwingMotorTorque = -1.f; c.swingMotorAxis = 0.f;
    c.twistMotorType = constraint_velocity_motor; c.twistMotorVelocityOrTargetAngle = 0.f; c.maxTwistMotorTorque = -1.f;
    a.scene->coneTwistConstraints.push_back(c);
    return { pushPair(*a.scene, 4, a, b) };
}
slider_constraint_handle addSliderConstraintFromGlobalPoints(scene_entity& a, scene_entity& b, vec3 globalAnchor, vec3 globalAxis, float minLimit, float maxLimit)
{   // physics.cpp:312-342
    slider_constraint c; memset(&c, 0, sizeof(c));
    const trs& tA = a.transform(); const trs& tB = b.transform();
    st3(c.localAnchorA, inverseTransformPosition(tA, globalAnchor)); st3(c.localAnchorB, inverseTransformPosition(tB, globalAnchor));
    st3(c.localAxisA, inverseTransformDirection(tA, globalAxis));
    st4(c.initialInvRotationDifference, conjugate(tB.rotation) * tA.rotation);
    c.negDistanceLimit = minLimit; c.posDistanceLimit = maxLimit;
    c.motorType = constraint_velocity_motor; c.motorVelocityOrTargetDistance = 0.f; c.maxMotorForce = -1.f;
    a.scene->sliderConstraints.push_back(c);
    return { pushPair(*a.scene, 5, a, b) };
}

// ------------------------------------------------------------------------------------------------------ ragdoll
static inline float deg2rad(float d) { return d * (B200_PI / 180.f); }   // math.h:15,24
static vec3 transformPosition(const trs& m, vec3 pos) { return m.rotation * (m.scale * pos) + m.position; }   // math.cpp:518-521
static vec3 transformDirection(const trs& m, vec3 dir) { return m.rotation * dir; }

void humanoid_ragdoll::initialize(game_scene& scene, vec3 initialHipPosition, float initialRotation)
{   // ragdoll.cpp:10-135 : same body parts, shapes, joints and order
    const float scale = 0.42f;
    const bool kin = false; const float gf = 1.f;
    physics_material material = { physics_material_type_flesh, 0.2f, 1.f, 985.f };
    const vec3 Z = mk3(0.f, 0.f, 1.f);
    struct Part { const char* name; trs t; };
    Part parts[14] = {
        { "Torso", makeTrs(scale * mk3(0.f, 0.f, 0.f)) },
        { "Head", makeTrs(scale * mk3(0.f, 1.45f, 0.f)) },
        { "Left upper arm", makeTrs(scale * mk3(-0.6f, 0.75f, 0.f), quatAxisAngle(Z, deg2rad(-30.f))) },
        { "Left lower arm", makeTrs(scale * mk3(-0.884f, 0.044f, -0.043f), quatAxisAngle(Z, deg2rad(-20.f))) },
        { "Right upper arm", makeTrs(scale * mk3(0.6f, 0.75f, 0.f), quatAxisAngle(Z, deg2rad(30.f))) },
        { "Right lower arm", makeTrs(scale * mk3(0.884f, 0.044f, -0.043f), quatAxisAngle(Z, deg2rad(20.f))) },
        { "Left upper leg", makeTrs(scale * mk3(-0.371f, -0.812f, 0.f), quatAxisAngle(Z, deg2rad(-10.f))) },
        { "Left lower leg", makeTrs(scale * mk3(-0.452f, -1.955f, 0.f), quatAxisAngle(Z, deg2rad(-3.5f))) },
        { "Left foot", makeTrs(scale * mk3(-0.498f, -2.585f, -0.18f)) },
        { "Left toes", makeTrs(scale * mk3(-0.498f, -2.585f, -0.637f)) },
        { "Right upper leg", makeTrs(scale * mk3(0.371f, -0.812f, 0.f), quatAxisAngle(Z, deg2rad(10.f))) },
        { "Right lower leg", makeTrs(scale * mk3(0.452f, -1.955f, 0.f), quatAxisAngle(Z, deg2rad(3.5f))) },
        { "Right foot", makeTrs(scale * mk3(0.498f, -2.585f, -0.18f)) },
        { "Right toes", makeTrs(scale * mk3(0.498f, -2.585f, -0.637f)) },
    };
    enum { torso, head, lUpperArm, lLowerArm, rUpperArm, rLowerArm, lUpperLeg, lLowerLeg, lFoot, lToes, rUpperLeg, rLowerLeg, rFoot, rToes };
    auto cap = [&](vec3 a, vec3 b, float r) { bounding_capsule c = { scale * a, scale * b, scale * r }; return collider_component::asCapsule(c, material); };
    for (int i = 0; i < 14; ++i) { bodyParts[i] = scene.createEntity(parts[i].name); bodyParts[i].addTransform(parts[i].t.position, parts[i].t.rotation); }
    bodyParts[torso].addCollider(cap(mk3(-0.2f, 0.f, 0.f), mk3(0.2f, 0.f, 0.f), 0.25f)).addCollider(cap(mk3(-0.16f, 0.32f, 0.f), mk3(0.16f, 0.32f, 0.f), 0.2f))
                    .addCollider(cap(mk3(-0.14f, 0.62f, 0.f), mk3(0.14f, 0.62f, 0.f), 0.22f)).addCollider(cap(mk3(-0.14f, 0.92f, 0.f), mk3(0.14f, 0.92f, 0.f), 0.2f)).addRigidBody(kin, gf);
    bodyParts[head].addCollider(cap(mk3(0.f, -0.075f, 0.f), mk3(0.f, 0.075f, 0.f), 0.25f)).addRigidBody(kin, gf);
    bodyParts[lUpperArm].addCollider(cap(mk3(0.f, -0.2f, 0.f), mk3(0.f, 0.2f, 0.f), 0.15f)).addRigidBody(kin, gf);
    bodyParts[lLowerArm].addCollider(cap(mk3(0.f, -0.2f, 0.f), mk3(0.f, 0.2f, 0.f), 0.15f)).addRigidBody(kin, gf);
    bodyParts[rUpperArm].addCollider(cap(mk3(0.f, -0.2f, 0.f), mk3(0.f, 0.2f, 0.f), 0.15f)).addRigidBody(kin, gf);
    bodyParts[rLowerArm].addCollider(cap(mk3(0.f, -0.2f, 0.f), mk3(0.f, 0.2f, 0.f), 0.15f)).addRigidBody(kin, gf);
    bodyParts[lUpperLeg].addCollider(cap(mk3(0.f, -0.3f, 0.f), mk3(0.f, 0.3f, 0.f), 0.25f)).addRigidBody(kin, gf);
    bodyParts[lLowerLeg].addCollider(cap(mk3(0.f, -0.3f, 0.f), mk3(0.f, 0.3f, 0.f), 0.18f)).addRigidBody(kin, gf);
    bodyParts[lFoot].addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(scale * mk3(0.f), scale * mk3(0.1587f, 0.1f, 0.3424f)), material)).addRigidBody(kin, gf);
    bodyParts[lToes].addCollider(cap(mk3(-0.0587f, 0.f, 0.f), mk3(0.0587f, 0.f, 0.f), 0.1f)).addRigidBody(kin, gf);
    bodyParts[rUpperLeg].addCollider(cap(mk3(0.f, -0.3f, 0.f), mk3(0.f, 0.3f, 0.f), 0.25f)).addRigidBody(kin, gf);
    bodyParts[rLowerLeg].addCollider(cap(mk3(0.f, -0.3f, 0.f), mk3(0.f, 0.3f, 0.f), 0.18f)).addRigidBody(kin, gf);
    bodyParts[rFoot].addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(scale * mk3(0.f), scale * mk3(0.1587f, 0.1f, 0.3424f)), material)).addRigidBody(kin, gf);
    bodyParts[rToes].addCollider(cap(mk3(-0.0587f, 0.f, 0.f), mk3(0.0587f, 0.f, 0.f), 0.1f)).addRigidBody(kin, gf);

    const trs* T[14]; for (int i = 0; i < 14; ++i) T[i] = &parts[i].t;
    addConeTwistConstraintFromGlobalPoints(bodyParts[torso], bodyParts[head], transformPosition(*T[torso], scale * mk3(0.f, 1.2f, 0.f)), mk3(0.f, 1.f, 0.f), deg2rad(50.f), deg2rad(90.f));
    addConeTwistConstraintFromGlobalPoints(bodyParts[torso], bodyParts[lUpperArm], transformPosition(*T[torso], scale * mk3(-0.4f, 1.f, 0.f)), mk3(-1.f, 0.f, 0.f), deg2rad(130.f), deg2rad(90.f));
    addHingeConstraintFromGlobalPoints(bodyParts[lUpperArm], bodyParts[lLowerArm], transformPosition(*T[lUpperArm], scale * mk3(0.f, -0.42f, 0.f)), normalize(mk3(1.f, 0.f, 1.f)), deg2rad(-5.f), deg2rad(85.f));
    addConeTwistConstraintFromGlobalPoints(bodyParts[torso], bodyParts[rUpperArm], transformPosition(*T[torso], scale * mk3(0.4f, 1.f, 0.f)), mk3(1.f, 0.f, 0.f), deg2rad(130.f), deg2rad(90.f));
    addHingeConstraintFromGlobalPoints(bodyParts[rUpperArm], bodyParts[rLowerArm], transformPosition(*T[rUpperArm], scale * mk3(0.f, -0.42f, 0.f)), normalize(mk3(1.f, 0.f, -1.f)), deg2rad(-5.f), deg2rad(85.f));
    addConeTwistConstraintFromGlobalPoints(bodyParts[torso], bodyParts[lUpperLeg], transformPosition(*T[torso], scale * mk3(-0.3f, -0.25f, 0.f)), transformDirection(*T[lUpperLeg], mk3(0.f, -1.f, 0.f)), -1.f, deg2rad(30.f));
    addHingeConstraintFromGlobalPoints(bodyParts[lUpperLeg], bodyParts[lLowerLeg], transformPosition(*T[lUpperLeg], scale * mk3(0.f, -0.6f, 0.f)), mk3(1.f, 0.f, 0.f), deg2rad(-90.f), deg2rad(5.f));
    addConeTwistConstraintFromGlobalPoints(bodyParts[lLowerLeg], bodyParts[lFoot], transformPosition(*T[lLowerLeg], scale * mk3(0.f, -0.52f, 0.f)), transformDirection(*T[lLowerLeg], mk3(0.f, -1.f, 0.f)), deg2rad(75.f), deg2rad(20.f));
    addHingeConstraintFromGlobalPoints(bodyParts[lFoot], bodyParts[lToes], transformPosition(*T[lFoot], scale * mk3(0.f, 0.f, -0.36f)), mk3(1.f, 0.f, 0.f), deg2rad(-45.f), deg2rad(45.f));
    addConeTwistConstraintFromGlobalPoints(bodyParts[torso], bodyParts[rUpperLeg], transformPosition(*T[torso], scale * mk3(0.3f, -0.25f, 0.f)), transformDirection(*T[rUpperLeg], mk3(0.f, -1.f, 0.f)), -1.f, deg2rad(30.f));
    addHingeConstraintFromGlobalPoints(bodyParts[rUpperLeg], bodyParts[rLowerLeg], transformPosition(*T[rUpperLeg], scale * mk3(0.f, -0.6f, 0.f)), mk3(1.f, 0.f, 0.f), deg2rad(-90.f), deg2rad(5.f));
    addConeTwistConstraintFromGlobalPoints(bodyParts[rLowerLeg], bodyParts[rFoot], transformPosition(*T[rLowerLeg], scale * mk3(0.f, -0.52f, 0.f)), transformDirection(*T[rLowerLeg], mk3(0.f, -1.f, 0.f)), deg2rad(75.f), deg2rad(20.f));
    addHingeConstraintFromGlobalPoints(bodyParts[rFoot], bodyParts[rToes], transformPosition(*T[rFoot], scale * mk3(0.f, 0.f, -0.36f)), mk3(1.f, 0.f, 0.f), deg2rad(-45.f), deg2rad(45.f));

    // place the whole figure (ragdoll.cpp:105-113)
    quat rotation = quatAxisAngle(mk3(0.f, 1.f, 0.f), initialRotation);
    for (int i = 0; i < 14; ++i)
    {
        trs& t = bodyParts[i].transform();
        t.rotation = rotation * t.rotation;
        t.position = rotation * t.position + initialHipPosition;
        int rb = scene.entities[bodyParts[i].handle].rigidBody;
        scene.physicsTransform0[rb] = scene.physicsTransform1[rb] = t;
    }
}

// ------------------------------------------------------------------------------------------------------ vehicle
// vehicle::initialize (vehicle.cpp:303-497) without the render meshes: the same 16 parts, colliders (gear teeth = capsules, wheels =
// cylinders, motor = box; rods, suspensions and hollow gear bodies carry no collider, vehicle.cpp:270), joint anchors, limits and motors.
namespace {
struct GearDesc { float height, cylinderRadius, cylinderInnerRadius; uint32_t numTeeth; float toothLength, toothWidth, friction, density; };   // vehicle.cpp:5-16
struct WheelDesc { float height, radius, innerRadius, friction, density; };                                                                       // vehicle.cpp:18-26
struct Attachment { int type; float rodLength, rodThickness; GearDesc gear; WheelDesc wheel; };     // 0 none, 1 gear, 2 wheel (vehicle.cpp:28-52)
Attachment attachNone(float len, float th) { Attachment a = {}; a.type = 0; a.rodLength = len; a.rodThickness = th; return a; }
Attachment attachGear(float len, float th, GearDesc g) { Attachment a = {}; a.type = 1; a.rodLength = len; a.rodThickness = th; a.gear = g; return a; }
Attachment attachWheel(float len, float th, WheelDesc w) { Attachment a = {}; a.type = 2; a.rodLength = len; a.rodThickness = th; a.wheel = w; return a; }

void attachColliders(scene_entity axis, const Attachment& at, float sign)
{   // vehicle.cpp:54-140 (collider part)
    const float rodOffset = at.rodLength * sign;
    if (at.type == 1)
    {
        const GearDesc& d = at.gear;
        for (uint32_t i = 0; i < d.numTeeth; ++i)
        {
            const float angle = i * 6.28318530718f / d.numTeeth;                      // M_TAU (math.h:19)
            const quat localRotation = quatAxisAngle(mk3(0.f, 1.f, 0.f), angle);
            const vec3 center = localRotation * mk3(d.cylinderRadius + d.toothLength * 0.5f, 0.f, 0.f);
            bounding_capsule c;
            c.positionA = center + mk3(0.f, rodOffset, 0.f) - localRotation * mk3(d.toothLength * 0.5f, 0.f, 0.f);
            c.positionB = center + mk3(0.f, rodOffset, 0.f) + localRotation * mk3(d.toothLength * 0.5f, 0.f, 0.f);
            c.radius = d.toothWidth * 0.5f;
            axis.addCollider(collider_component::asCapsule(c, { physics_material_type_wood, 0.2f, d.friction, d.density }));
        }
    }
    else if (at.type == 2)
    {
        const WheelDesc& d = at.wheel;
        bounding_cylinder c = { mk3(0.f, rodOffset - d.height * 0.5f, 0.f), mk3(0.f, rodOffset + d.height * 0.5f, 0.f), d.radius };
        axis.addCollider(collider_component::asCylinder(c, { physics_material_type_wood, 0.2f, d.friction, d.density }));
    }
}
scene_entity makeAxis(game_scene& scene, vec3 position, quat rotation, GearDesc desc, const Attachment* first = nullptr, const Attachment* second = nullptr)
{   // createAxis, vehicle.cpp:142-171
    scene_entity axis = scene.createEntity("Axis"); axis.addTransform(position, rotation);
    attachColliders(axis, attachGear(0.f, 0.f, desc), 1.f);
    if (first) attachColliders(axis, *first, 1.f);
    if (second) attachColliders(axis, *second, -1.f);
    axis.addRigidBody(false);
    return axis;
}
scene_entity makeGearAxis(game_scene& scene, vec3 position, quat rotation, float length, uint32_t numTeeth, float toothLength, float toothWidth, float friction, float density)
{   // createGearAxis, vehicle.cpp:173-217
    scene_entity axis = scene.createEntity("Gear Axis"); axis.addTransform(position, rotation);
    const float distance = length - toothWidth, stride = distance / (numTeeth - 1), leftOffset = -0.5f * length + 0.5f * toothWidth;
    for (uint32_t i = 0; i < numTeeth; ++i)
    {
        const vec3 center = mk3(leftOffset + i * stride, toothWidth * 0.5f, 0.f);
        bounding_capsule c = { center + mk3(0.f, toothLength * 0.5f, 0.f), center - mk3(0.f, toothLength * 0.5f, 0.f), toothWidth * 0.5f };
        axis.addCollider(collider_component::asCapsule(c, { physics_material_type_wood, 0.2f, friction, density }));
    }
    axis.addRigidBody(false);
    return axis;
}
scene_entity makeWheel(game_scene& scene, vec3 position, quat rotation, WheelDesc d)
{   // createWheel, vehicle.cpp:219-247
    scene_entity w = scene.createEntity("Wheel"); w.addTransform(position, rotation);
    bounding_cylinder c = { mk3(0.f, -d.height * 0.5f, 0.f), mk3(0.f, d.height * 0.5f, 0.f), d.radius };
    w.addCollider(collider_component::asCylinder(c, { physics_material_type_wood, 0.2f, d.friction, d.density }));
    w.addRigidBody(false);
    return w;
}
scene_entity makeBare(game_scene& scene, const char* name, vec3 position, quat rotation)
{   // createWheelSuspension / createRod (vehicle.cpp:249-301): rigid bodies without colliders, default mass properties
    scene_entity e = scene.createEntity(name); e.addTransform(position, rotation); e.addRigidBody(false);
    return e;
}
scene_entity makeRod(game_scene& scene, vec3 from, vec3 to) { return makeBare(scene, "Rod", (from + to) * 0.5f, rotateFromTo(mk3(0.f, 1.f, 0.f), normalize(to - from))); }
} // namespace

void vehicle::initialize(game_scene& scene, vec3 initialMotorPosition, float initialRotation)
{
    const float density = 2000.f;
    parts[motor] = scene.createEntity("Motor");
    parts[motor].addTransform(mk3(0.f), qIdentity())
        .addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(mk3(0.f), mk3(0.6f, 0.1f, 1.f)), { physics_material_type_wood, 0.2f, 0.f, density })).addRigidBody(false);
    const GearDesc motorGearDesc = { 0.1f, 0.2f, 0.f, 8, 0.07f, 0.1f, 0.f, density };
    const float rodThickness = 0.05f;
    const GearDesc steeringWheelDesc = { 0.1f, 0.4f, 0.2f, 0, 0.07f, 0.1f, 0.f, density };
    WheelDesc wheelDesc; wheelDesc.height = 0.3f; wheelDesc.radius = 0.7f; wheelDesc.innerRadius = wheelDesc.radius * 0.4f; wheelDesc.friction = 1.f; wheelDesc.density = 50.f;
    const float motorGearY = 0.25f, gearOffset = 0.26f;

    parts[motorGear] = makeAxis(scene, mk3(0.f, motorGearY, 0.f), qIdentity(), motorGearDesc);
    motorConstraint = addHingeConstraintFromGlobalPoints(parts[motor], parts[motorGear], mk3(0.f, motorGearY, 0.f), mk3(0.f, 1.f, 0.f));
    { hinge_constraint& c = getConstraint(scene, motorConstraint); c.maxMotorTorque = 500.f; c.motorVelocityOrTargetAngle = 0.f; }

    const float driveAxisLength = 4.5f;
    const Attachment driveAxisAttachment = attachGear(driveAxisLength * 0.57f - 1.1f, rodThickness, motorGearDesc);
    parts[driveAxis] = makeAxis(scene, mk3(0.f, motorGearY + gearOffset, gearOffset), quatAxisAngle(mk3(-1.f, 0.f, 0.f), deg2rad(90.f)), motorGearDesc, nullptr, &driveAxisAttachment);
    addHingeConstraintFromGlobalPoints(parts[motor], parts[driveAxis], mk3(0.f, motorGearY + gearOffset, gearOffset), mk3(0.f, 0.f, 1.f));

    const float axisLength = 1.5f, suspensionLength = 0.4f;
    const float frontAxisOffsetZ = -driveAxisLength * 0.5f + gearOffset * 2.f;
    const vec3 frontAxisPos = mk3(0.f, motorGearY + gearOffset, frontAxisOffsetZ);
    parts[frontAxis] = makeRod(scene, frontAxisPos + mk3(axisLength, 0.f, 0.f), frontAxisPos - mk3(axisLength, 0.f, 0.f));
    addFixedConstraintFromGlobalPoints(parts[motor], parts[frontAxis], frontAxisPos);

    const Attachment steeringWheelAttachment = attachGear(2.f, rodThickness, motorGearDesc);
    const quat steeringWheelRot = quatAxisAngle(mk3(-1.f, 0.f, 0.f), deg2rad(-80.f));
    const vec3 steeringWheelPos = mk3(0.f, 1.12f, 0.81f);
    parts[steeringWheel] = makeAxis(scene, steeringWheelPos, steeringWheelRot, steeringWheelDesc, nullptr, &steeringWheelAttachment);
    steeringWheelConstraint = addHingeConstraintFromGlobalPoints(parts[motor], parts[steeringWheel], steeringWheelPos, steeringWheelRot * mk3(0.f, -1.f, 0.f));
    { hinge_constraint& c = getConstraint(scene, steeringWheelConstraint); c.motorType = constraint_position_motor; c.maxMotorTorque = 1000.f; c.motorVelocityOrTargetAngle = 0.f; }

    const vec3 steeringAxisPos = mk3(0.f, motorGearY + gearOffset + 0.06f, frontAxisOffsetZ + 0.49f);
    const float steeringAxisLength = axisLength * 1.05f;
    parts[steeringAxis] = makeGearAxis(scene, steeringAxisPos, steeringWheelRot, steeringAxisLength, 8, motorGearDesc.toothLength, motorGearDesc.toothWidth, motorGearDesc.friction, motorGearDesc.density);
    addSliderConstraintFromGlobalPoints(parts[motor], parts[steeringAxis], steeringAxisPos, mk3(1.f, 0.f, 0.f), -4.f, 4.f);
    const vec3 leftSteeringAttach = steeringAxisPos - mk3(steeringAxisLength * 0.5f, 0.f, 0.f), rightSteeringAttach = steeringAxisPos + mk3(steeringAxisLength * 0.5f, 0.f, 0.f);

    const vec3 leftSuspPos = frontAxisPos - mk3(axisLength, 0.f, 0.f), leftSuspAttach = leftSuspPos + mk3(0.f, 0.f, suspensionLength);
    parts[leftWheelSuspension] = makeBare(scene, "Wheel suspension", leftSuspPos, qIdentity());
    addHingeConstraintFromGlobalPoints(parts[motor], parts[leftWheelSuspension], leftSuspPos, mk3(0.f, 1.f, 0.f), deg2rad(-45.f), deg2rad(45.f));
    const vec3 rightSuspPos = frontAxisPos + mk3(axisLength, 0.f, 0.f), rightSuspAttach = rightSuspPos + mk3(0.f, 0.f, suspensionLength);
    parts[rightWheelSuspension] = makeBare(scene, "Wheel suspension", rightSuspPos, qIdentity());
    addHingeConstraintFromGlobalPoints(parts[motor], parts[rightWheelSuspension], rightSuspPos, mk3(0.f, 1.f, 0.f), deg2rad(-45.f), deg2rad(45.f));

    const vec3 leftFrontWheelPos = leftSuspPos - mk3(suspensionLength * 0.5f, 0.f, 0.f), rightFrontWheelPos = rightSuspPos + mk3(suspensionLength * 0.5f, 0.f, 0.f);
    parts[leftFrontWheel] = makeWheel(scene, leftFrontWheelPos, quatAxisAngle(mk3(0.f, 0.f, 1.f), deg2rad(90.f)), wheelDesc);
    parts[rightFrontWheel] = makeWheel(scene, rightFrontWheelPos, quatAxisAngle(mk3(0.f, 0.f, 1.f), deg2rad(90.f)), wheelDesc);
    addHingeConstraintFromGlobalPoints(parts[leftFrontWheel], parts[leftWheelSuspension], leftFrontWheelPos, mk3(1.f, 0.f, 0.f));
    addHingeConstraintFromGlobalPoints(parts[rightFrontWheel], parts[rightWheelSuspension], rightFrontWheelPos, mk3(1.f, 0.f, 0.f));

    parts[leftWheelArm] = makeRod(scene, leftSteeringAttach, leftSuspAttach);
    parts[rightWheelArm] = makeRod(scene, rightSteeringAttach, rightSuspAttach);
    addBallConstraintFromGlobalPoints(parts[leftWheelSuspension], parts[leftWheelArm], leftSuspAttach);
    addBallConstraintFromGlobalPoints(parts[steeringAxis], parts[leftWheelArm], leftSteeringAttach);
    addBallConstraintFromGlobalPoints(parts[rightWheelSuspension], parts[rightWheelArm], rightSuspAttach);
    addBallConstraintFromGlobalPoints(parts[steeringAxis], parts[rightWheelArm], rightSteeringAttach);

    GearDesc rearAxisGearDesc = motorGearDesc; rearAxisGearDesc.cylinderRadius = 0.5f; rearAxisGearDesc.cylinderInnerRadius = 0.4f; rearAxisGearDesc.numTeeth = 17;
    const float rearAxisOffsetZ = driveAxisLength * 0.505f, rearAxisOffsetX = -gearOffset;
    parts[differentialSunGear] = makeAxis(scene, mk3(rearAxisOffsetX, motorGearY + gearOffset, rearAxisOffsetZ), quatAxisAngle(mk3(0.f, 0.f, -1.f), deg2rad(90.f)), rearAxisGearDesc);
    addHingeConstraintFromGlobalPoints(parts[motor], parts[differentialSunGear], mk3(rearAxisOffsetX, motorGearY + gearOffset, rearAxisOffsetZ), mk3(1.f, 0.f, 0.f));

    const vec3 spiderPos = mk3(0.11f, motorGearY + gearOffset * 2.f, rearAxisOffsetZ);
    const Attachment spiderAttachment = attachNone(0.2f, 0.02f);
    parts[differentialSpiderGear] = makeAxis(scene, spiderPos, qIdentity(), motorGearDesc, &spiderAttachment);
    addHingeConstraintFromGlobalPoints(parts[differentialSunGear], parts[differentialSpiderGear], spiderPos, mk3(0.f, 1.f, 0.f));

    const vec3 leftRearWheelPos = spiderPos + mk3(-gearOffset, -gearOffset, 0.f), rightRearWheelPos = spiderPos + mk3(gearOffset, -gearOffset, 0.f);
    const Attachment leftWheelAttachment = attachWheel(axisLength + spiderPos.x, rodThickness, wheelDesc), rightWheelAttachment = attachWheel(axisLength - spiderPos.x, rodThickness, wheelDesc);
    parts[leftRearWheel] = makeAxis(scene, leftRearWheelPos, quatAxisAngle(mk3(0.f, 0.f, -1.f), deg2rad(90.f)), motorGearDesc, nullptr, &leftWheelAttachment);
    parts[rightRearWheel] = makeAxis(scene, rightRearWheelPos, quatAxisAngle(mk3(0.f, 0.f, -1.f), deg2rad(90.f)), motorGearDesc, &rightWheelAttachment, nullptr);
    addHingeConstraintFromGlobalPoints(parts[motor], parts[leftRearWheel], leftRearWheelPos, mk3(1.f, 0.f, 0.f));
    addHingeConstraintFromGlobalPoints(parts[motor], parts[rightRearWheel], rightRearWheelPos, mk3(1.f, 0.f, 0.f));

    const quat rotation = quatAxisAngle(mk3(0.f, 1.f, 0.f), initialRotation);
    for (int i = 0; i < numParts; ++i)
    {   // vehicle.cpp:486-496: ONLY the transform_component moves. Unlike humanoid_ragdoll (ragdoll.cpp:111-112) the reference does not
        // refresh physics_transform0 / 1 here, so the simulation itself starts from the unplaced pose (and overwrites the render transform
        // after the first step). Reproduced as is; a caller that wants the placement to stick re-adds the transform (scene.h:96-107).
        trs& t = parts[i].transform();
        t.position = rotation * t.position + initialMotorPosition;
        t.rotation = rotation * t.rotation;
    }
}

// ------------------------------------------------------------------------------------------------------ the step
static inline quat nlerpQuat(quat l, quat u, float t)
{   // math.h:670-673 : lerp of the 4-vectors, then normalise
    return normalize(mkq(l.x + t * (u.x - l.x), l.y + t * (u.y - l.y), l.z + t * (u.z - l.z), l.w + t * (u.w - l.w)));
}

static int ensureContext(game_scene& scene)
{
    const uint32_t nb = (uint32_t)scene.rigidBodies.size(), nc = (uint32_t)scene.colliders.size();
    size_t nj = 1; for (int k = 0; k < 6; ++k) if (scene.constraintEntities[k].size() > nj) nj = scene.constraintEntities[k].size();
    if (scene.ctx) return 0;
    b200phys_config cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.device = 0; cfg.maxBodies = nb + (nb >> 2) + 16; cfg.maxColliders = nc + (nc >> 2) + 16; cfg.maxTransforms = (uint32_t)scene.entities.size() + 1; cfg.maxJointsPerKind = (uint32_t)nj + 16;
    if (!scene.heightmaps.empty()) cfg.maxContacts = 2u * 16u * cfg.maxColliders + 64u * cfg.maxColliders;     // one contact per touching triangle
    scene.collidersDirty = scene.jointsDirty = scene.hullsDirty = true;
    return b200phys_create(&cfg, &scene.ctx);
}

// the gather half of physicsStepInternal: every pool that changed goes to the device (also used once by the device-side RL environment)
int uploadSceneToDevice(game_scene& scene)
{
    if (scene.rigidBodies.empty()) return 0;
    int r = ensureContext(scene); if (r) return r;
    if (scene.hullsDirty)
    {
        std::vector<float> v; std::vector<uint32_t> o; scene.gatherHulls(v, o);
        r = b200phys_set_hulls(scene.ctx, v.data(), o.data(), (uint32_t)o.size() - 1); if (r) return r;
        scene.hullsDirty = false;
    }
    const uint32_t numBodies = (uint32_t)scene.rigidBodies.size();
    if (numBodies > scene.stageCapacity)
    {   // page-locked staging, kept across steps
        if (scene.stageBodies) b200phys_host_free(scene.stageBodies);
        if (scene.stageStates) b200phys_host_free(scene.stageStates);
        scene.stageCapacity = numBodies + (numBodies >> 2) + 16;
        scene.stageBodies = (b200phys_body*)b200phys_host_alloc(sizeof(b200phys_body) * (size_t)scene.stageCapacity);
        scene.stageStates = (b200phys_body_state*)b200phys_host_alloc(sizeof(b200phys_body_state) * (size_t)scene.stageCapacity);
        if (!scene.stageBodies || !scene.stageStates) return B200PHYS_E_INVALID;
    }
    scene.gatherBodies(scene.stageBodies);
    r = b200phys_upload_bodies(scene.ctx, scene.stageBodies, numBodies); if (r) return r;
    if (scene.collidersDirty)
    {
        std::vector<b200phys_collider> cols; std::vector<b200phys_transform> xf; scene.gatherColliders(cols, xf);
        r = b200phys_upload_transforms(scene.ctx, xf.data(), (uint32_t)xf.size()); if (r) return r;
        r = b200phys_upload_colliders(scene.ctx, cols.data(), (uint32_t)cols.size()); if (r) return r;
        scene.collidersDirty = false;
    }
    if (!scene.heightmaps.empty() && scene.heightmaps[0].dirty)
    {   // physics.cpp:1238-1248 loops over every heightmap entity; scenes have one terrain, the device context holds one
        heightmap_collider_component& hm = scene.heightmaps[0];
        b200phys_heightmap d; memset(&d, 0, sizeof(d));
        d.chunksPerDim = hm.chunksPerDim; d.chunkSize = hm.chunkSize; d.minCorner[0] = hm.minCorner.x; d.minCorner[1] = hm.minCorner.y; d.minCorner[2] = hm.minCorner.z;
        d.amplitudeScale = hm.amplitudeScale; d.restitution = hm.material.restitution; d.friction = hm.material.friction;
        r = b200phys_upload_heightmap(scene.ctx, &d, hm.heights.data()); if (r) return r;
        hm.dirty = false;
    }
    {   // joint parameters (motor targets etc.) are host-mutable between steps (learned_locomotion.cpp:73-91): re-upload every step
        for (int k = 0; k < 6; ++k)
        {
            if (scene.jointPairsDirty) scene.refreshJointPairs();
            size_t n = 0; const void* structs = scene.jointStructs(k, n);        // straight from the scene's arrays
            r = b200phys_upload_joints(scene.ctx, k, structs, scene.jointPairs[k].data(), (uint32_t)n); if (r) return r;
        }
        scene.jointsDirty = false;
    }
    if (!scene.forceFields.empty())
    {   // getForceFieldStates (physics.cpp:759-787): fields with a collider act locally, the others on every body
        std::vector<float> local(3 * scene.forceFields.size(), 0.f); vec3 global = b200::mk3(0.f);
        for (size_t i = 0; i < scene.forceFields.size(); ++i)
        {
            const game_scene::entity_data& e = scene.entities[scene.forceFieldEntity[i]];
            vec3 force = scene.forceFields[i].force;
            if (e.hasTransform) force = e.transform.rotation * force;
            if (e.numColliders) { local[3 * i] = force.x; local[3 * i + 1] = force.y; local[3 * i + 2] = force.z; }
            else global += force;
        }
        r = b200phys_upload_force_fields(scene.ctx, local.data(), (uint32_t)scene.forceFields.size()); if (r) return r;
        const float g3[3] = { global.x, global.y, global.z };
        r = b200phys_set_global_force_field(scene.ctx, g3); if (r) return r;
        scene.lastGlobalForceField = global;
    }
    return 0;
}

// physicsStepInternal (physics.cpp:1180-1362): gather pools -> C ABI -> scatter results
static int physicsStepInternal(game_scene& scene, const physics_settings& settings, float dt)
{
    if (scene.rigidBodies.empty()) return 0;   // physics.cpp:1186-1189
    int r = uploadSceneToDevice(scene); if (r) return r;
    const uint32_t numBodies = (uint32_t)scene.rigidBodies.size();
    const bool wantEvents = (bool)settings.collisionBeginCallback || (bool)settings.collisionEndCallback;
    if (wantEvents != scene.eventsEnabled) { r = b200phys_enable_events(scene.ctx, wantEvents ? 1 : 0); if (r) return r; scene.eventsEnabled = wantEvents; }
    r = b200phys_step(scene.ctx, dt, settings.numRigidSolverIterations, settings.exactConstraintOrder ? B200PHYS_STEP_EXACT_ORDER : B200PHYS_STEP_FAST); if (r) return r;
    if (wantEvents)
    {   // handleCollisionCallbacks (physics.cpp:1037-1178): the device did the frame-to-frame diff, the host only dispatches, in the
        // reference's order: both lists sorted by (colliderA, colliderB) and merged, end before begin on ties of position
        uint32_t nb = 0, ne = 0;
        r = b200phys_download_events(scene.ctx, nullptr, 0, &nb, nullptr, 0, &ne); if (r && r != B200PHYS_E_CAPACITY) return r;
        std::vector<b200phys_collision_begin_event> be(nb); std::vector<b200phys_collider_pair> ee(ne);
        r = b200phys_download_events(scene.ctx, be.data(), nb, &nb, ee.data(), ne, &ne); if (r) return r;
        std::sort(be.begin(), be.end(), [](const b200phys_collision_begin_event& x, const b200phys_collision_begin_event& y) { return x.colliderA != y.colliderA ? x.colliderA < y.colliderA : x.colliderB < y.colliderB; });
        std::sort(ee.begin(), ee.end(), [](const b200phys_collider_pair& x, const b200phys_collider_pair& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; });
        const size_t C = scene.colliders.size();
        auto col = [&](uint32_t worldIndex) -> const collider_component& { return scene.colliders[C - 1 - worldIndex]; };   // physics.cpp:1058-1059
        size_t ib = 0, ie = 0;
        while (ib < be.size() || ie < ee.size())
        {
            const bool takeEnd = ie < ee.size() && (ib >= be.size() || ee[ie].a < be[ib].colliderA || (ee[ie].a == be[ib].colliderA && ee[ie].b < be[ib].colliderB));
            if (takeEnd)
            {
                const collider_component& a = col(ee[ie].a); const collider_component& b = col(ee[ie].b);
                if (settings.collisionEndCallback) settings.collisionEndCallback(collision_end_event{ a.parentEntity, b.parentEntity, a, b });
                ++ie;
            }
            else
            {
                const b200phys_collision_begin_event& e = be[ib];
                const collider_component& a = col(e.colliderA); const collider_component& b = col(e.colliderB);
                if (settings.collisionBeginCallback)
                    settings.collisionBeginCallback(collision_begin_event{ a.parentEntity, b.parentEntity, a, b, b200::mk3(e.point[0], e.point[1], e.point[2]),
                                                                           b200::mk3(e.normal[0], e.normal[1], e.normal[2]), b200::mk3(e.relativeVelocity[0], e.relativeVelocity[1], e.relativeVelocity[2]) });
                ++ib;
            }
        }
    }
    if (!scene.triggers.empty())
    {   // handleNonCollisionInteractions (physics.cpp:970-1034): leave / enter callbacks in the order of the reference's merge loop
        uint32_t ne = 0, nl = 0;
        r = b200phys_download_trigger_events(scene.ctx, nullptr, 0, &ne, nullptr, 0, &nl); if (r && r != B200PHYS_E_CAPACITY) return r;
        std::vector<b200phys_trigger_event> en(ne), le(nl);
        r = b200phys_download_trigger_events(scene.ctx, en.data(), ne, &ne, le.data(), nl, &nl); if (r) return r;
        auto key = [](const b200phys_trigger_event& e) { return ((unsigned long long)e.triggerIndex << 32) | e.rigidBodyIndex; };
        size_t ie = 0, il = 0;
        while (ie < en.size() || il < le.size())
        {
            const bool leave = il < le.size() && (ie >= en.size() || key(le[il]) < key(en[ie]));
            const b200phys_trigger_event& e = leave ? le[il++] : en[ie++];
            const trigger_component& t = scene.triggers[e.triggerIndex];
            if (t.callback) t.callback(trigger_event{ scene.triggerEntity[e.triggerIndex], scene.rigidBodyEntity[e.rigidBodyIndex], leave ? trigger_event_leave : trigger_event_enter });
        }
    }
    if (!scene.cloths.empty())
    {   // physics.cpp:1354-1358: wind from the global force field, then simulate; creation / parameter changes / fixed-row moves first
        for (cloth_component& cl : scene.cloths)
        {
            if (cl.deviceId < 0)
            {
                b200phys_cloth_desc d; memset(&d, 0, sizeof(d));
                d.width = cl.width; d.height = cl.height; d.gridSizeX = cl.gridSizeX; d.gridSizeY = cl.gridSizeY; d.totalMass = cl.totalMass; d.stiffness = cl.stiffness; d.damping = cl.damping; d.gravityFactor = cl.gravityFactor;
                uint32_t id = 0; r = b200phys_cloth_create(scene.ctx, &d, &id); if (r) return r;
                cl.deviceId = (int32_t)id; cl.upTotalMass = cl.totalMass; cl.upStiffness = cl.stiffness; cl.upDamping = cl.damping; cl.upGravityFactor = cl.gravityFactor;
                cl.positions.resize((size_t)cl.gridSizeX * cl.gridSizeY);
            }
            for (const cloth_component::fixed_update& f : cl.pendingFixed)
            {
                const float rot[4] = { f.transform.rotation.x, f.transform.rotation.y, f.transform.rotation.z, f.transform.rotation.w };
                const float pos[3] = { f.transform.position.x, f.transform.position.y, f.transform.position.z }, sc[3] = { f.transform.scale.x, f.transform.scale.y, f.transform.scale.z };
                r = b200phys_cloth_set_fixed_vertices(scene.ctx, (uint32_t)cl.deviceId, rot, pos, sc, f.moveRigid ? 1 : 0); if (r) return r;
            }
            cl.pendingFixed.clear();
            if (cl.totalMass != cl.upTotalMass || cl.stiffness != cl.upStiffness || cl.damping != cl.upDamping || cl.gravityFactor != cl.upGravityFactor)
            {
                r = b200phys_cloth_set_properties(scene.ctx, (uint32_t)cl.deviceId, cl.totalMass, cl.stiffness, cl.damping, cl.gravityFactor); if (r) return r;
                cl.upTotalMass = cl.totalMass; cl.upStiffness = cl.stiffness; cl.upDamping = cl.damping; cl.upGravityFactor = cl.gravityFactor;
            }
        }
        const float w3[3] = { scene.lastGlobalForceField.x, scene.lastGlobalForceField.y, scene.lastGlobalForceField.z };
        r = b200phys_cloth_apply_wind(scene.ctx, 0xFFFFFFFFu, w3); if (r) return r;
        r = b200phys_cloth_simulate(scene.ctx, settings.numClothVelocityIterations, settings.numClothPositionIterations, settings.numClothDriftIterations, dt); if (r) return r;
        for (cloth_component& cl : scene.cloths) { r = b200phys_cloth_download(scene.ctx, (uint32_t)cl.deviceId, &cl.positions[0].x, nullptr); if (r) return r; }
    }
    r = b200phys_download_bodies(scene.ctx, scene.stageStates, numBodies); if (r) return r;
    scene.scatterBodyStates(scene.stageStates, numBodies);
    r = b200phys_get_stats(scene.ctx, &scene.lastStats);
    return r;
}

int physicsStep(game_scene& scene, memory_arena&, float& timer, const physics_settings& settings, float dt)
{   // physics.cpp:1364-1413
    int rc = 0;
    if (settings.fixedFrameRate)
    {
        const float h = 1.f / (float)settings.frameRate;
        timer += dt;
        uint32_t its = 0;
        if (timer >= h)
        {
            for (size_t i = 0; i < scene.rigidBodies.size(); ++i) scene.physicsTransform0[i] = scene.physicsTransform1[i];
            while (timer >= h && its++ < settings.maxPhysicsIterationsPerFrame)
            {
                int r = physicsStepInternal(scene, settings, h); if (r && !rc) rc = r;
                timer -= h;
            }
        }
        if (timer >= h) timer = fmodf(timer, h);   // dropping physics frames
        const float t = timer / h;
        for (size_t i = 0; i < scene.rigidBodies.size(); ++i)
        {
            trs& out = scene.entities[scene.rigidBodyEntity[i]].transform; const trs& a = scene.physicsTransform0[i]; const trs& b = scene.physicsTransform1[i];
            out.position = lerp(a.position, b.position, t);
            out.rotation = nlerpQuat(a.rotation, b.rotation, t);
            out.scale = lerp(a.scale, b.scale, t);
        }
    }
    else
    {
        rc = physicsStepInternal(scene, settings, dt);
        for (size_t i = 0; i < scene.rigidBodies.size(); ++i) scene.entities[scene.rigidBodyEntity[i]].transform = scene.physicsTransform1[i];
    }
    return rc;
}

} // namespace b200host
