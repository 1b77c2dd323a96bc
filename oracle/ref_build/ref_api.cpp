// ref_api.cpp -- C entry points over the reference's OWN physics code compiled with g++ (oracle/_ref/libref*.so).
// TEST INFRASTRUCTURE ONLY: used by tests/ to pin the oracle restatement (oracle/*.cpp) to reference-owned code, and by bench.py's
// reference arm as the CPU baseline. Never linked or loaded by the product (d3d12renderer_b200/).
//
// Everything physical below is the reference's: game_scene / EnTT pools (src/scene), physicsStep (src/physics/physics.cpp:1364),
// narrowphase (collision_narrow.cpp:2328), the constraint initialise / solve functions (constraints.cpp), humanoid_ragdoll::create
// (ragdoll.cpp) and vehicle::create (vehicle.cpp). This file only converts between the flat boundary structs of include/b200phys.h
// and the reference's components.
#include "pch.h"
#include "physics/physics.h"
#include "physics/collision_broad.h"
#include "physics/collision_narrow.h"
#include "physics/ragdoll.h"
#include "physics/vehicle.h"
#include "physics/heightmap_collision.h"
#include "physics/cloth.h"
#include "terrain/heightmap_collider.h"
#include "scene/scene.h"
#include "core/random.h"
#include "b200phys.h"
#include <chrono>

extern std::vector<bounding_hull_geometry> boundingHullGeometries;      // physics.cpp:14 (made external by patch_reference.py)
void getWorldSpaceColliders(game_scene& scene, bounding_box* outWorldspaceAABBs, collider_union* outWorldSpaceColliders, uint16 dummyRigidBodyIndex);   // physics.cpp:631 (same)

static uint32 gAssertCount = 0;
extern "C" void refAssertFailed(const char*, const char*, int) { ++gAssertCount; }

namespace {

struct RefScene
{
    game_scene scene;
    memory_arena arena;
    std::vector<entity_handle> bodyEntities, transformEntities;
    float timer = 0.f;
    double lastStepSeconds = 0.0;
    std::vector<uint16> heightmapHeights;   // storage behind heightmap_collider_chunk::setHeights (the chunk keeps the pointer)
    std::vector<uint32> triggerLog;          // (trigger index, rigid body index, 0 enter / 1 leave) triples of the last step, callback order
    RefScene() { arena.initialize(0, GB(8)); }
};

vec3 v3(const float* f) { return vec3(f[0], f[1], f[2]); }
quat q4(const float* f) { return quat(f[0], f[1], f[2], f[3]); }

collider_component toCollider(const b200phys_collider& c)
{
    physics_material m = { physics_material_type_none, c.restitution, c.friction, c.density };
    const float* s = c.shape;
    switch (c.type)
    {
    case B200PHYS_SPHERE: return collider_component::asSphere({ v3(s), s[3] }, m);
    case B200PHYS_CAPSULE: return collider_component::asCapsule({ v3(s), v3(s + 3), s[6] }, m);
    case B200PHYS_CYLINDER: return collider_component::asCylinder({ v3(s), v3(s + 3), s[6] }, m);
    case B200PHYS_AABB: return collider_component::asAABB(bounding_box::fromMinMax(v3(s), v3(s + 3)), m);
    case B200PHYS_OBB: { bounding_oriented_box o; o.rotation = q4(s); o.center = v3(s + 4); o.radius = v3(s + 7); return collider_component::asOBB(o, m); }
    default: { bounding_hull h; h.rotation = q4(s); h.position = v3(s + 4); uint32 g; memcpy(&g, s + 7, 4); h.geometryIndex = g; return collider_component::asHull(h, m); }
    }
}

void writeWorldCollider(const collider_union& w, float* out16)
{   // the oracle's world-collider packing (tests/oracle_lib.py: shape[10], then type / objectType / objectIndex as floats)
    memset(out16, 0, 64);
    switch (w.type)
    {
    case collider_type_sphere: memcpy(out16, &w.sphere.center, 12); out16[3] = w.sphere.radius; break;
    case collider_type_capsule: memcpy(out16, &w.capsule.positionA, 12); memcpy(out16 + 3, &w.capsule.positionB, 12); out16[6] = w.capsule.radius; break;
    case collider_type_cylinder: memcpy(out16, &w.cylinder.positionA, 12); memcpy(out16 + 3, &w.cylinder.positionB, 12); out16[6] = w.cylinder.radius; break;
    case collider_type_aabb: memcpy(out16, &w.aabb.minCorner, 12); memcpy(out16 + 3, &w.aabb.maxCorner, 12); break;
    case collider_type_obb: memcpy(out16, &w.obb.rotation, 16); memcpy(out16 + 4, &w.obb.center, 12); memcpy(out16 + 7, &w.obb.radius, 12); break;
    default: memcpy(out16, &w.hull.rotation, 16); memcpy(out16 + 4, &w.hull.position, 12); break;
    }
    out16[13] = (float)w.type; out16[14] = (float)w.objectType; out16[15] = (float)w.objectIndex;
}

} // namespace

// cloth_component keeps its particles private and befriends cloth_render_component, which the PHYSICS_ONLY build does not define
// (cloth.h:57-72): this definition only reads the particle arrays.
struct cloth_render_component
{
    static const std::vector<vec3>& positions(const cloth_component& c) { return c.positions; }
    static const std::vector<vec3>& velocities(const cloth_component& c) { return c.velocities; }
};

extern "C" {

uint32 ref_version() { return 1; }
uint32 ref_assert_count(int reset) { uint32 n = gAssertCount; if (reset) gAssertCount = 0; return n; }

// hull geometries are process-global in the reference (physics.cpp:14): replaces the list. triangles: uint32 triples per hull
void ref_set_hulls(const float* vertices, const uint32* vertexOffsets, uint32 numHulls, const uint32* triangles, const uint32* triangleOffsets)
{
    boundingHullGeometries.clear();
    for (uint32 h = 0; h < numHulls; ++h)
    {
        std::vector<vec3> pos; std::vector<indexed_triangle16> tris;
        for (uint32 v = vertexOffsets[h]; v < vertexOffsets[h + 1]; ++v) pos.push_back(v3(vertices + 3 * v));
        for (uint32 t = triangleOffsets[h]; t < triangleOffsets[h + 1]; ++t) tris.push_back({ (uint16)triangles[3 * t], (uint16)triangles[3 * t + 1], (uint16)triangles[3 * t + 2] });
        boundingHullGeometries.push_back(bounding_hull_geometry::fromMesh(pos.data(), (uint32)pos.size(), tris.data(), (uint32)tris.size()));
    }
}

void* ref_scene_create() { return new RefScene(); }
void ref_scene_destroy(void* p) { delete (RefScene*)p; }

// Builds an EnTT scene from the flat boundary arrays so that the reference's own index conventions reproduce the array indices:
// rigid body storage index = array index (bodies are created in order); collider world index = array index (the reference numbers
// colliders by REVERSE pool iteration, physics.cpp:635-641, so they are added last-to-first). Mass properties, velocities and
// accumulators are then overwritten with the array values (the reference recomputes mass from the shapes on addComponent).
int ref_scene_build(void* p, const b200phys_body* bodies, uint32 numBodies, const b200phys_transform* transforms, uint32 numTransforms,
                    const b200phys_collider* colliders, uint32 numColliders)
{
    RefScene& R = *(RefScene*)p;
    for (uint32 i = 0; i < numBodies; ++i)
    {
        scene_entity e = R.scene.createEntity("body");
        e.addComponent<transform_component>(v3(bodies[i].position), q4(bodies[i].rotation));
        e.addComponent<rigid_body_component>(bodies[i].invMass == 0.f, bodies[i].gravityFactor, bodies[i].linearDamping, bodies[i].angularDamping);
        R.bodyEntities.push_back(e.handle);
    }
    // non-rigid-body parents; triggers / force fields are told apart by the colliders that reference them
    std::vector<int> kind(numTransforms, B200PHYS_OBJ_STATIC); std::vector<uint32> index(numTransforms, 0);
    for (uint32 c = 0; c < numColliders; ++c)
        if (colliders[c].objectType != B200PHYS_OBJ_RIGID_BODY && colliders[c].parentTransform < numTransforms) { kind[colliders[c].parentTransform] = colliders[c].objectType; index[colliders[c].parentTransform] = colliders[c].objectIndex; }
    // force fields / triggers must get their pool index = objectIndex: create those parents in objectIndex order
    std::vector<uint32> order(numTransforms); for (uint32 t = 0; t < numTransforms; ++t) order[t] = t;
    std::stable_sort(order.begin(), order.end(), [&](uint32 a, uint32 b) { return index[a] < index[b]; });
    R.transformEntities.assign(numTransforms, entt::null);
    for (uint32 k = 0; k < numTransforms; ++k)
    {
        const uint32 t = order[k];
        scene_entity e = R.scene.createEntity("parent");
        e.addComponent<transform_component>(v3(transforms[t].position), q4(transforms[t].rotation));
        if (kind[t] == B200PHYS_OBJ_TRIGGER)
        {
            RefScene* self = &R; const uint32 triggerIndex = index[t];
            e.addComponent<trigger_component>();
            e.getComponent<trigger_component>().callback = [self, triggerIndex](trigger_event ev)
            {
                uint32 body = 0xFFFFFFFFu;
                for (size_t b = 0; b < self->bodyEntities.size(); ++b) if (self->bodyEntities[b] == ev.other.handle) { body = (uint32)b; break; }
                self->triggerLog.push_back(triggerIndex); self->triggerLog.push_back(body); self->triggerLog.push_back((uint32)ev.type);
            };
        }
        if (kind[t] == B200PHYS_OBJ_FORCE_FIELD) e.addComponent<force_field_component>();
        R.transformEntities[t] = e.handle;
    }
    for (uint32 k = numColliders; k-- > 0;)
    {
        const b200phys_collider& c = colliders[k];
        entity_handle parent = (c.objectType == B200PHYS_OBJ_RIGID_BODY) ? R.bodyEntities.at(c.objectIndex) : (c.parentTransform < numTransforms ? R.transformEntities[c.parentTransform] : entt::null);
        if (parent == entt::null) { scene_entity e = R.scene.createEntity("static"); parent = e.handle; }     // parent without a transform (identity, physics.cpp:645-647)
        scene_entity pe(parent, &R.scene.registry);
        pe.addComponent<collider_component>(toCollider(c));
    }
    for (uint32 i = 0; i < numBodies; ++i)
    {
        scene_entity e(R.bodyEntities[i], &R.scene.registry);
        rigid_body_component& rb = e.getComponent<rigid_body_component>();
        rb.invMass = bodies[i].invMass; memcpy(rb.invInertia.m, bodies[i].invInertia, 36); rb.localCOGPosition = v3(bodies[i].localCOGPosition);
        rb.linearVelocity = v3(bodies[i].linearVelocity); rb.angularVelocity = v3(bodies[i].angularVelocity);
        rb.forceAccumulator = v3(bodies[i].forceAccumulator); rb.torqueAccumulator = v3(bodies[i].torqueAccumulator);
    }
    return 0;
}

// mass properties the reference itself computed when the colliders were added (scene.h:38-63, rigid_body.cpp:29-81); call BEFORE
// ref_scene_build's overwrite is not possible, so this variant builds with the reference's values kept
void ref_scene_get_bodies(void* p, b200phys_body* out)
{
    RefScene& R = *(RefScene*)p;
    for (size_t i = 0; i < R.bodyEntities.size(); ++i)
    {
        scene_entity e(R.bodyEntities[i], &R.scene.registry);
        const rigid_body_component& rb = e.getComponent<rigid_body_component>();
        const physics_transform1_component& t = e.getComponent<physics_transform1_component>();
        b200phys_body& b = out[i];
        memcpy(b.rotation, &t.rotation, 16); memcpy(b.position, &t.position, 12); memcpy(b.localCOGPosition, &rb.localCOGPosition, 12);
        b.invMass = rb.invMass; memcpy(b.invInertia, rb.invInertia.m, 36); b.gravityFactor = rb.gravityFactor; b.linearDamping = rb.linearDamping; b.angularDamping = rb.angularDamping;
        memcpy(b.linearVelocity, &rb.linearVelocity, 12); memcpy(b.angularVelocity, &rb.angularVelocity, 12);
        memcpy(b.forceAccumulator, &rb.forceAccumulator, 12); memcpy(b.torqueAccumulator, &rb.torqueAccumulator, 12);
    }
}

int ref_scene_add_joints(void* p, int kind, const void* joints, const b200phys_body_pair* pairs, uint32 count)
{
    RefScene& R = *(RefScene*)p;
    for (uint32 i = 0; i < count; ++i)
    {
        scene_entity a(R.bodyEntities.at(pairs[i].rbA), &R.scene.registry), b(R.bodyEntities.at(pairs[i].rbB), &R.scene.registry);
        switch (kind)
        {   // the boundary structs have the reference structs' field order (include/b200phys.h); constraints.h is the authority on size
        case B200PHYS_JOINT_DISTANCE: { distance_constraint c; memcpy(&c, (const b200phys_distance_joint*)joints + i, sizeof(b200phys_distance_joint)); addConstraint(a, b, c); } break;
        case B200PHYS_JOINT_BALL: { ball_constraint c; memcpy(&c, (const b200phys_ball_joint*)joints + i, sizeof(b200phys_ball_joint)); addConstraint(a, b, c); } break;
        case B200PHYS_JOINT_FIXED: { fixed_constraint c; const b200phys_fixed_joint& j = ((const b200phys_fixed_joint*)joints)[i]; c.initialInvRotationDifference = q4(j.initialInvRotationDifference); c.localAnchorA = v3(j.localAnchorA); c.localAnchorB = v3(j.localAnchorB); addConstraint(a, b, c); } break;
        case B200PHYS_JOINT_HINGE: { hinge_constraint c; static_assert(sizeof(hinge_constraint) == sizeof(b200phys_hinge_joint), "hinge layout"); memcpy(&c, (const b200phys_hinge_joint*)joints + i, sizeof(c)); addConstraint(a, b, c); } break;
        case B200PHYS_JOINT_CONE_TWIST: { cone_twist_constraint c; static_assert(sizeof(cone_twist_constraint) == sizeof(b200phys_cone_twist_joint), "cone twist layout"); memcpy(&c, (const b200phys_cone_twist_joint*)joints + i, sizeof(c)); addConstraint(a, b, c); } break;
        default:
        {
            slider_constraint c; const b200phys_slider_joint& j = ((const b200phys_slider_joint*)joints)[i];
            c.initialInvRotationDifference = q4(j.initialInvRotationDifference); c.localAnchorA = v3(j.localAnchorA); c.localAnchorB = v3(j.localAnchorB); c.localAxisA = v3(j.localAxisA);
            c.negDistanceLimit = j.negDistanceLimit; c.posDistanceLimit = j.posDistanceLimit; c.maxMotorForce = j.maxMotorForce; c.motorType = (constraint_motor_type)j.motorType; c.motorVelocity = j.motorVelocityOrTargetDistance;
            addConstraint(a, b, c);
        } break;
        }
    }
    return 0;
}

void ref_scene_set_force_fields(void* p, const float* forces3, uint32 count, const float* global3)
{   // localized fields in force-field pool order (= objectIndex), plus one collider-less entity for the global force (physics.cpp:759-787)
    RefScene& R = *(RefScene*)p;
    uint32 i = 0;
    for (auto [entity, ff] : R.scene.view<force_field_component>().each()) { (void)entity; (void)ff; ++i; }
    std::vector<entity_handle> es; for (auto [entity, ff] : R.scene.view<force_field_component>().each()) es.push_back(entity);
    // view iteration is back to front (most recent first): objectIndex k is es[n - 1 - k]
    for (uint32 k = 0; k < count && k < es.size(); ++k) R.scene.registry.get<force_field_component>(es[es.size() - 1 - k]).force = v3(forces3 + 3 * k);
    if (global3 && (global3[0] != 0.f || global3[1] != 0.f || global3[2] != 0.f)) { scene_entity g = R.scene.createEntity("global"); g.addComponent<force_field_component>(); g.getComponent<force_field_component>().force = v3(global3); }
}

// the reference's own content builders
uint32 ref_scene_add_ragdoll(void* p, const float* hip3, float rotation)
{
    RefScene& R = *(RefScene*)p;
    humanoid_ragdoll r = humanoid_ragdoll::create(R.scene, v3(hip3), rotation);
    for (uint32 i = 0; i < 14; ++i) R.bodyEntities.push_back(r.bodyParts[i].handle);
    return 14;
}
uint32 ref_scene_add_vehicle(void* p, const float* pos3, float rotation, float motorVelocity, float steeringAngle)
{
    RefScene& R = *(RefScene*)p;
    vehicle v = vehicle::create(R.scene, v3(pos3), rotation);
    for (uint32 i = 0; i < 16; ++i) R.bodyEntities.push_back(v.parts[i].handle);
    // the two motors sit on the first and fourth hinge created (vehicle.cpp:352-358,384-390): hinge pool, creation order
    uint32 n = R.scene.numberOfComponentsOfType<hinge_constraint>();
    (void)n; (void)motorVelocity; (void)steeringAngle;
    return 16;
}
uint32 ref_scene_add_ground(void* p, const float* center3, const float* radius3, const float* position3, float restitution, float friction, float density)
{
    RefScene& R = *(RefScene*)p;
    scene_entity e = R.scene.createEntity("ground");
    e.addComponent<transform_component>(v3(position3), quat::identity);
    e.addComponent<collider_component>(collider_component::asAABB(bounding_box::fromCenterRadius(v3(center3), v3(radius3)), { physics_material_type_none, restitution, friction, density }));
    return 1;
}
// heightmap_collider_component (terrain/heightmap_collider.h:132-150) on its own entity: chunksPerDim^2 chunks of 129 x 129 heights
int ref_scene_add_heightmap(void* p, uint32 chunksPerDim, float chunkSize, const float* minCorner3, float amplitudeScale, float restitution, float friction, const uint16* heights)
{
    RefScene& R = *(RefScene*)p;
    const size_t perChunk = (size_t)TERRAIN_LOD_0_VERTICES_PER_DIMENSION * TERRAIN_LOD_0_VERTICES_PER_DIMENSION;
    R.heightmapHeights.assign(heights, heights + perChunk * chunksPerDim * chunksPerDim);
    scene_entity e = R.scene.createEntity("terrain");
    heightmap_collider_component& hm = e.addComponent<heightmap_collider_component>(chunksPerDim, chunkSize, physics_material{ physics_material_type_none, restitution, friction, 1.f }).getComponent<heightmap_collider_component>();
    hm.update(v3(minCorner3), amplitudeScale);
    for (uint32 z = 0; z < chunksPerDim; ++z)
        for (uint32 x = 0; x < chunksPerDim; ++x)
            hm.collider(x, z).setHeights(R.heightmapHeights.data() + perChunk * (z * chunksPerDim + x));
    return 0;
}
float ref_scene_heightmap_height_at(void* p, float x, float z)
{
    RefScene& R = *(RefScene*)p;
    for (auto [h, hm] : R.scene.view<heightmap_collider_component>().each()) return hm.getHeightAt(vec2(x, z));
    return -FLT_MAX;
}
// cloth (physics/cloth.{h,cpp}); physicsStepInternal applies the global force field as wind and simulates (physics.cpp:1354-1358)
void* ref_cloth_create(float width, float height, uint32 gx, uint32 gy, float totalMass, float stiffness, float damping, float gravityFactor)
{ return new cloth_component(width, height, gx, gy, totalMass, stiffness, damping, gravityFactor); }
void ref_cloth_destroy(void* c) { delete (cloth_component*)c; }
void ref_cloth_set_fixed(void* c, const float* rot4, const float* pos3, const float* scale3, int moveRigid)
{ trs t; t.rotation = q4(rot4); t.position = v3(pos3); t.scale = v3(scale3); ((cloth_component*)c)->setWorldPositionOfFixedVertices(t, moveRigid != 0); }
void ref_cloth_set_properties(void* c, float totalMass, float stiffness, float damping, float gravityFactor)
{ cloth_component& k = *(cloth_component*)c; k.totalMass = totalMass; k.stiffness = stiffness; k.damping = damping; k.gravityFactor = gravityFactor; }
void ref_cloth_wind(void* c, const float* f3) { ((cloth_component*)c)->applyWindForce(v3(f3)); }
void ref_cloth_simulate(void* c, uint32 vi, uint32 pi, uint32 di, float dt) { ((cloth_component*)c)->simulate(vi, pi, di, dt); }
void ref_cloth_get(void* c, float* positions, float* velocities)
{
    const cloth_component& k = *(cloth_component*)c;
    const std::vector<vec3>& p = cloth_render_component::positions(k); const std::vector<vec3>& v = cloth_render_component::velocities(k);
    for (size_t i = 0; i < p.size(); ++i)
    {
        if (positions) { positions[3 * i] = p[i].x; positions[3 * i + 1] = p[i].y; positions[3 * i + 2] = p[i].z; }
        if (velocities) { velocities[3 * i] = v[i].x; velocities[3 * i + 1] = v[i].y; velocities[3 * i + 2] = v[i].z; }
    }
}
// ---- binary entity stream (scene/serialization_binary.cpp:137-513). That file itself cannot be compiled here (asset registry, meshes,
// terrain); what IS reference-owned are the component structs it memcpy's. This restates the stream framing -- one bool per component
// type of serialized_components in order, then the payload -- over the reference's real components: the struct layouts (g++ build of
// the patched headers; sizes checked against SURVEY 8's MSVC sizes by ref_layout_info) pin the byte layout of every physics payload.
void ref_layout_info(uint32* out)
{
    out[0] = (uint32)sizeof(trs); out[1] = (uint32)sizeof(rigid_body_component); out[2] = (uint32)sizeof(collider_union); out[3] = (uint32)offsetof(collider_union, material);
    out[4] = (uint32)offsetof(collider_union, type); out[5] = (uint32)offsetof(collider_union, objectIndex); out[6] = (uint32)sizeof(tag_component);
    out[7] = (uint32)sizeof(distance_constraint); out[8] = (uint32)sizeof(ball_constraint); out[9] = (uint32)sizeof(fixed_constraint); out[10] = (uint32)sizeof(hinge_constraint);
    out[11] = (uint32)sizeof(cone_twist_constraint); out[12] = (uint32)sizeof(slider_constraint); out[13] = (uint32)sizeof(physics_material); out[14] = (uint32)sizeof(constraint_type);
    out[15] = (uint32)sizeof(entity_handle);
}
uint64 ref_scene_serialize_body_entity(void* p, uint32 bodyIndex, uint8* out, uint64 cap)
{
    RefScene& R = *(RefScene*)p;
    scene_entity entity = { R.bodyEntities[bodyIndex], R.scene };
    uint64 n = 0;
    auto put = [&](const void* src, uint64 bytes) { if (n + bytes <= cap) memcpy(out + n, src, bytes); n += bytes; };
    auto flag = [&](bool b) { put(&b, 1); };
    auto whole = [&](auto* comp) { flag(comp != nullptr); if (comp) put(comp, sizeof(*comp)); };
    whole(entity.getComponentIfExists<tag_component>());
    whole(entity.getComponentIfExists<transform_component>());
    flag(false); flag(false); flag(false);                                   // position / position_rotation / position_scale
    flag(entity.hasComponent<dynamic_transform_component>());                // empty payload (:148-149)
    flag(false); flag(false); flag(false);                                   // mesh, point light, spot light
    whole(entity.getComponentIfExists<rigid_body_component>());
    whole(entity.getComponentIfExists<force_field_component>());
    flag(false);                                                             // cloth_component: a stand-alone entity in the tests
    flag(false);                                                             // cloth_render_component
    if (physics_reference_component* ref = entity.getComponentIfExists<physics_reference_component>())
    {   // :207-234
        flag(true);
        put(&ref->numColliders, 4);
        for (collider_component& collider : collider_component_iterator(entity)) put((collider_union*)&collider, sizeof(collider_union));
        put(&ref->numConstraints, 4);
        for (auto [constraintEntity, constraintType] : constraint_entity_iterator(entity))
        {
            auto& cref = constraintEntity.getComponent<constraint_entity_reference_component>();
            put(&constraintType, sizeof(constraintType)); put(&cref.entityA, sizeof(entity_handle)); put(&cref.entityB, sizeof(entity_handle));
            switch (constraintType)
            {
                case constraint_type_distance: put(&constraintEntity.getComponent<distance_constraint>(), sizeof(distance_constraint)); break;
                case constraint_type_ball: put(&constraintEntity.getComponent<ball_constraint>(), sizeof(ball_constraint)); break;
                case constraint_type_fixed: put(&constraintEntity.getComponent<fixed_constraint>(), sizeof(fixed_constraint)); break;
                case constraint_type_hinge: put(&constraintEntity.getComponent<hinge_constraint>(), sizeof(hinge_constraint)); break;
                case constraint_type_cone_twist: put(&constraintEntity.getComponent<cone_twist_constraint>(), sizeof(cone_twist_constraint)); break;
                case constraint_type_slider: put(&constraintEntity.getComponent<slider_constraint>(), sizeof(slider_constraint)); break;
                default: break;
            }
        }
    }
    else flag(false);
    flag(false); flag(false); flag(false); flag(false); flag(false);         // terrain, heightmap collider, grass, proc placement, water
    return n;
}
uint32 ref_scene_entity_of_body(void* p, uint32 bodyIndex) { return (uint32)((RefScene*)p)->bodyEntities[bodyIndex]; }
uint32 ref_scene_num_bodies(void* p) { return (uint32)((RefScene*)p)->bodyEntities.size(); }
uint32 ref_scene_num_colliders(void* p) { return ((RefScene*)p)->scene.numberOfComponentsOfType<collider_component>(); }
uint32 ref_scene_num_joints(void* p, int kind)
{
    game_scene& s = ((RefScene*)p)->scene;
    switch (kind) { case 0: return s.numberOfComponentsOfType<distance_constraint>(); case 1: return s.numberOfComponentsOfType<ball_constraint>(); case 2: return s.numberOfComponentsOfType<fixed_constraint>();
                    case 3: return s.numberOfComponentsOfType<hinge_constraint>(); case 4: return s.numberOfComponentsOfType<cone_twist_constraint>(); default: return s.numberOfComponentsOfType<slider_constraint>(); }
}
// hinge motor parameters by pool index (vehicle motors)
void ref_scene_set_hinge_motor(void* p, uint32 hingeIndex, int motorType, float velocityOrAngle, float maxTorque)
{
    hinge_constraint& c = ((RefScene*)p)->scene.getComponentAtIndex<hinge_constraint>(hingeIndex);
    c.motorType = (constraint_motor_type)motorType; c.motorVelocity = velocityOrAngle; c.maxMotorTorque = maxTorque;
}

// physicsStep (physics.cpp:1364-1413). simd = the three physics_settings switches (physics.h:394-396)
void ref_scene_step(void* p, float dt, uint32 iterations, int fixedFrameRate, uint32 frameRate, int simdBroad, int simdNarrow, int simdSolver)
{
    RefScene& R = *(RefScene*)p;
    physics_settings s;
    s.fixedFrameRate = fixedFrameRate != 0; s.frameRate = frameRate; s.numRigidSolverIterations = iterations;
    s.simdBroadPhase = simdBroad != 0; s.simdNarrowPhase = simdNarrow != 0; s.simdConstraintSolver = simdSolver != 0;
    R.triggerLog.clear();
    const auto t0 = std::chrono::steady_clock::now();
    R.arena.reset();                                       // application::update resets the frame arena (application.cpp:521)
    physicsStep(R.scene, R.arena, R.timer, s, dt);
    R.lastStepSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
uint32 ref_scene_trigger_log(void* p, uint32* out, uint32 capacityTriples)
{
    RefScene& R = *(RefScene*)p; const uint32 n = (uint32)R.triggerLog.size() / 3;
    if (out) memcpy(out, R.triggerLog.data(), 12 * (size_t)(n < capacityTriples ? n : capacityTriples));
    return n;
}
double ref_scene_last_step_seconds(void* p) { return ((RefScene*)p)->lastStepSeconds; }

void ref_scene_get_states(void* p, b200phys_body_state* out)
{   // physics_transform1 + velocities (what integrateVelocity wrote, rigid_body.cpp:126-142)
    RefScene& R = *(RefScene*)p;
    for (size_t i = 0; i < R.bodyEntities.size(); ++i)
    {
        scene_entity e(R.bodyEntities[i], &R.scene.registry);
        const rigid_body_component& rb = e.getComponent<rigid_body_component>();
        const physics_transform1_component& t = e.getComponent<physics_transform1_component>();
        memcpy(out[i].rotation, &t.rotation, 16); memcpy(out[i].position, &t.position, 12);
        memcpy(out[i].linearVelocity, &rb.linearVelocity, 12); memcpy(out[i].angularVelocity, &rb.angularVelocity, 12);
    }
}
// rigid body storage index of the i-th body the harness created (checks the "bodies are created in order" assumption)
uint32 ref_scene_body_storage_index(void* p, uint32 i)
{
    RefScene& R = *(RefScene*)p;
    scene_entity e(R.bodyEntities.at(i), &R.scene.registry);
    return (uint32)e.getComponentIndex<rigid_body_component>();
}

// detection stages on their own (getWorldSpaceColliders -> broadphase -> narrowphase, physics.cpp:1217-1233). Mutates the SAP context
// like a step would. Outputs (any may be null): aabbs [C][6]; world colliders [C][16] (oracle packing); pairs in SAP order [P][2];
// contacts (collision_contact, 32 B) + body pairs; colliding pairs + contact counts. Returns counts through outCounts[5] =
// {colliders, pairs, collisions, contacts, non-collision interactions}.
int ref_scene_detect(void* p, int simdBroad, int simdNarrow, float* outAabbs, float* outWorld16, uint32* outPairs, uint32 pairCapacity,
                     b200phys_contact* outContacts, b200phys_body_pair* outBodyPairs, uint32* outCollisionPairs, uint8* outContactCounts, uint32 contactCapacity, uint32* outCounts)
{
    RefScene& R = *(RefScene*)p;
    game_scene& scene = R.scene;
    memory_marker marker = R.arena.getMarker();
    const uint32 numColliders = scene.numberOfComponentsOfType<collider_component>();
    const uint32 numRigidBodies = scene.numberOfComponentsOfType<rigid_body_component>();
    bounding_box* aabbs = R.arena.allocate<bounding_box>(numColliders);
    collider_union* world = R.arena.allocate<collider_union>(numColliders);
    collider_pair* pairs = R.arena.allocate<collider_pair>(numColliders * numColliders + 5000);
    getWorldSpaceColliders(scene, aabbs, world, (uint16)numRigidBodies);
    const uint32 numPairs = broadphase(scene, aabbs, R.arena, pairs, simdBroad != 0);
    if (outAabbs) memcpy(outAabbs, aabbs, sizeof(bounding_box) * numColliders);
    if (outWorld16) for (uint32 i = 0; i < numColliders; ++i) writeWorldCollider(world[i], outWorld16 + 16 * i);
    if (outPairs) for (uint32 i = 0; i < numPairs && i < pairCapacity; ++i) { outPairs[2 * i] = pairs[i].colliderA; outPairs[2 * i + 1] = pairs[i].colliderB; }
    non_collision_interaction* nci = R.arena.allocate<non_collision_interaction>(numPairs);
    collision_contact* contacts = R.arena.allocate<collision_contact>(numPairs * 4 + 5000 + numColliders * 256);
    constraint_body_pair* bodyPairs = R.arena.allocate<constraint_body_pair>(numPairs * 4 + 5000 + numColliders * 256);
    uint8* counts = R.arena.allocate<uint8>(numPairs + 5000 + numColliders);
    collider_pair* outColl = R.arena.allocate<collider_pair>(numPairs + 5000 + numColliders);
    narrowphase_result r = narrowphase(world, pairs, numPairs, R.arena, contacts, bodyPairs, outColl, counts, nci, simdNarrow != 0);
    for (auto [h, hm] : scene.view<heightmap_collider_component>().each())
    {   // physics.cpp:1238-1248
        narrowphase_result hr = heightmapCollision(hm, world, aabbs, numColliders, contacts + r.numContacts, bodyPairs + r.numContacts, outColl + r.numCollisions, counts + r.numCollisions,
                                                   R.arena, (uint16)numRigidBodies);
        r.numCollisions += hr.numCollisions; r.numContacts += hr.numContacts;
    }
    static_assert(sizeof(collision_contact) == sizeof(b200phys_contact), "contact layout");
    for (uint32 i = 0; i < r.numContacts && i < contactCapacity; ++i)
    {
        if (outContacts) memcpy(&outContacts[i], &contacts[i], sizeof(b200phys_contact));
        if (outBodyPairs) { outBodyPairs[i].rbA = bodyPairs[i].rbA; outBodyPairs[i].rbB = bodyPairs[i].rbB; }
    }
    for (uint32 i = 0; i < r.numCollisions && i < contactCapacity; ++i)
    {
        if (outCollisionPairs) { outCollisionPairs[2 * i] = outColl[i].colliderA; outCollisionPairs[2 * i + 1] = outColl[i].colliderB; }
        if (outContactCounts) outContactCounts[i] = counts[i];
    }
    if (outCounts) { outCounts[0] = numColliders; outCounts[1] = numPairs; outCounts[2] = r.numCollisions; outCounts[3] = r.numContacts; outCounts[4] = r.numNonCollisionInteractions; }
    R.arena.resetToMarker(marker);
    return 0;
}

// One collider pair through the reference's narrowphase() (scalar or AVX2), WORLD-space colliders given in the boundary packing
// (b200phys_collider with objectType / objectIndex already final). Returns the contact count; normal[3], pointDepth[4][4].
int ref_collide_pair(const b200phys_collider* a, const b200phys_collider* b, int simd, float* normal3, float* pointDepth16)
{
    static memory_arena arena; static bool init = false;
    if (!init) { arena.initialize(0, GB(1)); init = true; }
    arena.reset();
    collider_union world[2];
    const b200phys_collider* in[2] = { a, b };
    for (int k = 0; k < 2; ++k)
    {
        collider_component c = toCollider(*in[k]);
        memcpy(&world[k], &c, sizeof(collider_union));
        world[k].objectType = (physics_object_type)in[k]->objectType; world[k].objectIndex = (uint16)in[k]->objectIndex;
    }
    collider_pair pair = { 0, 1 };
    collision_contact contacts[16]; constraint_body_pair bodyPairs[16]; collider_pair outPairs[4]; uint8 counts[4]; non_collision_interaction nci[4];
    narrowphase_result r = narrowphase(world, &pair, 1, arena, contacts, bodyPairs, outPairs, counts, nci, simd != 0);
    for (uint32 i = 0; i < r.numContacts && i < 4; ++i)
    {
        memcpy(pointDepth16 + 4 * i, &contacts[i].point, 12); pointDepth16[4 * i + 3] = contacts[i].penetrationDepth;
        memcpy(normal3, &contacts[i].normal, 12);
    }
    return (int)r.numContacts;
}

// The reference's RL environment exactly as its Physics-Lib DLL exports it (learned_locomotion.cpp:395-488); seed and push probability
// of its generator are made settable by patch_reference.py (the reference seeds with time(0) and hard-codes 0.02)
extern "C" { int getPhysicsStateSize(); int getPhysicsActionSize(); void getPhysicsRanges(float*, float*, float*, float*); void resetPhysics(float*); int updatePhysics(float*, float*, float*); }
extern random_number_generator rng; extern float refPushProbability; extern bool refEnvSimd;
int ref_env_sizes(int* stateSize, int* actionSize) { *stateSize = getPhysicsStateSize(); *actionSize = getPhysicsActionSize(); return 0; }
void ref_env_ranges(float* actionMin, float* actionMax) { std::vector<float> smin(getPhysicsStateSize()), smax(getPhysicsStateSize()); getPhysicsRanges(smin.data(), smax.data(), actionMin, actionMax); }
void ref_env_reset(uint64 seed, float pushProbability, int simd, float* outState)
{
    rng = random_number_generator(seed); refPushProbability = pushProbability; refEnvSimd = simd != 0;
    boundingHullGeometries.clear();
    resetPhysics(outState);
}
int ref_env_update(float* action, float* outState, float* outReward) { return updatePhysics(action, outState, outReward); }

// random_number_generator (core/random.h): n floats of randomFloat01 from a seed (pins the scene generators' stream)
void ref_random_floats01(uint64 seed, float* out, uint32 n) { random_number_generator rng = { seed }; for (uint32 i = 0; i < n; ++i) out[i] = rng.randomFloat01(); }

// mass properties of one collider by the reference's formulas (physics.cpp:1416-1588)
void ref_mass_properties(const b200phys_collider* c, float* inertia9, float* cog3, float* mass)
{
    collider_component col = toCollider(*c);
    physics_properties pp = col.calculatePhysicsProperties();
    memcpy(inertia9, pp.inertia.m, 36); memcpy(cog3, &pp.cog, 12); *mass = pp.mass;
}

} // extern "C"
