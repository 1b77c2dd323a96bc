// scene_binary.cpp -- the reference's binary entity stream (src/scene/serialization_binary.{h,cpp}: serializeEntityToMemory /
// deserializeEntityFromMemory, used by the editor's undo buffer and "play" snapshots) for the physics components of the host mirror.
// Format (serialization_binary.cpp:102-133, 444-513): for every type of `serialized_components`, in that order, one bool and -- if true --
// the payload: a raw copy of the component struct for plain components (:137-149), field lists for cloth (:169-196), and for
// physics_reference_component the entity's colliders as collider_union structs and its constraints as (type, entityA, entityB, struct)
// (:207-266). Payload byte layouts are the reference's struct layouts (trs 48, rigid_body_component 112, collider_union 80 with the material
// at 48 and type / objectType / objectIndex at 64 / 65 / 66, the six constraint structs 28 / 24 / 48 / 104 / 120 / 80 bytes): verified against
// the compiled reference by tests/test_ref_pins_oracle.py (ref_layout_info). Rendering / terrain components are written as
// "absent"; a stream that contains one cannot be read here (their sizes are not part of the stream).
#include "physics_host.h"
#include <cstring>

namespace b200host {

namespace {

enum { kTag, kTransform, kPosition, kPositionRotation, kPositionScale, kDynamicTransform, kMesh, kPointLight, kSpotLight, kRigidBody, kForceField, kCloth, kClothRender,
       kPhysicsReference, kTerrain, kHeightmapCollider, kGrass, kProcPlacement, kWater, kNumKinds };
const uint32_t kJointBytes[6] = { sizeof(distance_constraint), sizeof(ball_constraint), sizeof(fixed_constraint), sizeof(hinge_constraint), sizeof(cone_twist_constraint), sizeof(slider_constraint) };
static_assert(sizeof(distance_constraint) == 28 && sizeof(ball_constraint) == 24 && sizeof(fixed_constraint) == 48 && sizeof(hinge_constraint) == 104 &&
              sizeof(cone_twist_constraint) == 120 && sizeof(slider_constraint) == 80 && sizeof(rigid_body_component) == 112, "payload sizes of the reference's structs");

struct writer
{   // write_stream (:8-45)
    uint8_t* buffer; uint64_t size; uint64_t offset = 0; bool ok = true;
    void bytes(const void* p, uint64_t n) { if (offset + n >= size) { ok = false; size = 0; return; } memcpy(buffer + offset, p, n); offset += n; }
    template <typename T> void put(T v) { bytes(&v, sizeof(T)); }
    void floats(const float* f, int n) { bytes(f, sizeof(float) * n); }
    void zeros(uint64_t n) { const uint8_t z[16] = {}; while (n) { const uint64_t k = n < 16 ? n : 16; bytes(z, k); n -= k; } }
};
struct reader
{   // read_stream (:47-91)
    const uint8_t* buffer; uint64_t size; uint64_t offset = 0; bool ok = true;
    void bytes(void* p, uint64_t n) { if (size < offset + n) { ok = false; size = 0; memset(p, 0, n); return; } memcpy(p, buffer + offset, n); offset += n; }
    template <typename T> T get() { T v; bytes(&v, sizeof(T)); return v; }
    void skip(uint64_t n) { if (size < offset + n) { ok = false; size = 0; return; } offset += n; }
};

void writeTrs(writer& w, const trs& t) { const float f[10] = { t.rotation.x, t.rotation.y, t.rotation.z, t.rotation.w, t.position.x, t.position.y, t.position.z, t.scale.x, t.scale.y, t.scale.z }; w.floats(f, 10); w.zeros(8); }
trs readTrs(reader& r) { float f[10]; r.bytes(f, sizeof(f)); r.skip(8); trs t; t.rotation = b200::mkq(f[0], f[1], f[2], f[3]); t.position = b200::mk3(f[4], f[5], f[6]); t.scale = b200::mk3(f[7], f[8], f[9]); return t; }

// collider_union, 80 bytes: shape union (48), physics_material (16), type, objectType, objectIndex, padding
void writeCollider(writer& w, const collider_component& c)
{
    float s[12] = {};
    switch (c.type)
    {
    case collider_type_sphere: s[0] = c.sphere.center.x; s[1] = c.sphere.center.y; s[2] = c.sphere.center.z; s[3] = c.sphere.radius; break;
    case collider_type_capsule: case collider_type_cylinder:
        s[0] = c.capsule.positionA.x; s[1] = c.capsule.positionA.y; s[2] = c.capsule.positionA.z; s[3] = c.capsule.positionB.x; s[4] = c.capsule.positionB.y; s[5] = c.capsule.positionB.z; s[6] = c.capsule.radius; break;
    case collider_type_aabb: s[0] = c.aabb.minCorner.x; s[1] = c.aabb.minCorner.y; s[2] = c.aabb.minCorner.z; s[3] = c.aabb.maxCorner.x; s[4] = c.aabb.maxCorner.y; s[5] = c.aabb.maxCorner.z; break;
    case collider_type_obb:
        s[0] = c.obb.rotation.x; s[1] = c.obb.rotation.y; s[2] = c.obb.rotation.z; s[3] = c.obb.rotation.w; s[4] = c.obb.center.x; s[5] = c.obb.center.y; s[6] = c.obb.center.z;
        s[7] = c.obb.radius.x; s[8] = c.obb.radius.y; s[9] = c.obb.radius.z; break;
    default:
        s[0] = c.hull.rotation.x; s[1] = c.hull.rotation.y; s[2] = c.hull.rotation.z; s[3] = c.hull.rotation.w; s[4] = c.hull.position.x; s[5] = c.hull.position.y; s[6] = c.hull.position.z;
        memcpy(&s[8], &c.hull.geometryIndex, 4); break;       // the index shares an 8-byte union with a pointer (bounding_volumes.h:226-230): offset 32
    }
    w.floats(s, 12);
    w.put<int32_t>((int32_t)c.material.type); w.put(c.material.restitution); w.put(c.material.friction); w.put(c.material.density);
    w.put<uint8_t>((uint8_t)c.type); w.put<uint8_t>(0); w.put<uint16_t>(0);      // objectType / objectIndex "are only used internally" (physics.h:101-104)
    w.zeros(12);
}
collider_component readCollider(reader& r)
{
    float s[12]; r.bytes(s, sizeof(s));
    physics_material m; m.type = (physics_material_type)r.get<int32_t>(); m.restitution = r.get<float>(); m.friction = r.get<float>(); m.density = r.get<float>();
    const uint8_t type = r.get<uint8_t>(); r.skip(3 + 12);
    collider_component c;
    switch (type)
    {
    case collider_type_sphere: c = collider_component::asSphere({ b200::mk3(s[0], s[1], s[2]), s[3] }, m); break;
    case collider_type_capsule: c = collider_component::asCapsule({ b200::mk3(s[0], s[1], s[2]), b200::mk3(s[3], s[4], s[5]), s[6] }, m); break;
    case collider_type_cylinder: c = collider_component::asCylinder({ b200::mk3(s[0], s[1], s[2]), b200::mk3(s[3], s[4], s[5]), s[6] }, m); break;
    case collider_type_aabb: c = collider_component::asAABB(bounding_box::fromMinMax(b200::mk3(s[0], s[1], s[2]), b200::mk3(s[3], s[4], s[5])), m); break;
    case collider_type_obb: c = collider_component::asOBB({ b200::mkq(s[0], s[1], s[2], s[3]), b200::mk3(s[4], s[5], s[6]), b200::mk3(s[7], s[8], s[9]) }, m); break;
    default: { bounding_hull h; h.rotation = b200::mkq(s[0], s[1], s[2], s[3]); h.position = b200::mk3(s[4], s[5], s[6]); memcpy(&h.geometryIndex, &s[8], 4); c = collider_component::asHull(h, m); } break;
    }
    return c;
}

const void* jointStruct(const game_scene& s, constraint_type type, uint32_t index)
{
    switch (type)
    {
    case constraint_type_distance: return &s.distanceConstraints[index];
    case constraint_type_ball: return &s.ballConstraints[index];
    case constraint_type_fixed: return &s.fixedConstraints[index];
    case constraint_type_hinge: return &s.hingeConstraints[index];
    case constraint_type_cone_twist: return &s.coneTwistConstraints[index];
    default: return &s.sliderConstraints[index];
    }
}

} // namespace

// serializeEntityToMemory (serialization_binary.cpp:498-503); returns the bytes written (0 = the buffer was too small)
uint64_t serializeEntityToMemory(scene_entity entity, void* memory, uint64_t maxSize)
{
    game_scene& s = *entity.scene; const game_scene::entity_data& d = s.entities[entity.handle];
    writer w = { (uint8_t*)memory, maxSize };
    for (int kind = 0; kind < kNumKinds; ++kind)
    {
        switch (kind)
        {
        case kTag: { w.put(true); char name[16] = {}; strncpy(name, d.name.c_str(), sizeof(name)); name[sizeof(name) - 1] = 0; w.bytes(name, 16); } break;
        case kTransform: w.put(d.hasTransform); if (d.hasTransform) writeTrs(w, d.transform); break;
        case kDynamicTransform: w.put(d.rigidBody >= 0); break;                                  // added together with the rigid body (scene.h:66-77); empty payload
        case kRigidBody: w.put(d.rigidBody >= 0); if (d.rigidBody >= 0) w.bytes(&s.rigidBodies[d.rigidBody], sizeof(rigid_body_component)); break;
        case kForceField: w.put(d.forceField >= 0); if (d.forceField >= 0) { const vec3 f = s.forceFields[d.forceField].force; const float v[3] = { f.x, f.y, f.z }; w.floats(v, 3); } break;
        case kPhysicsReference:
        {
            constraint_entity_iterator joints(entity);
            const bool has = d.numColliders > 0 || !joints.items.empty();
            w.put(has);
            if (!has) break;
            w.put<uint32_t>(d.numColliders);
            for (entity_handle c = d.firstColliderEntity; c != null_entity; c = s.colliders[c].nextEntity) writeCollider(w, s.colliders[c]);
            w.put<uint32_t>((uint32_t)joints.items.size());
            for (const constraint_entity_iterator::item& it : joints)
            {
                const b200phys_body_pair pr = s.constraintEntities[it.type][it.index];
                w.put<int32_t>((int32_t)it.type); w.put<uint32_t>(pr.rbA); w.put<uint32_t>(pr.rbB);
                w.bytes(jointStruct(s, it.type, it.index), kJointBytes[it.type]);
            }
        } break;
        default: w.put(false); break;       // position variants, rendering, cloth (lives in game_scene::cloths, not on an entity), terrain
        }
    }
    return w.ok ? w.offset : 0;
}

// deserializeEntityFromMemory (:505-510): adds the components to `entity`; constraint partners are named by the handles in the stream
bool deserializeEntityFromMemory(scene_entity entity, const void* memory, uint64_t size)
{
    game_scene& s = *entity.scene;
    reader r = { (const uint8_t*)memory, size };
    for (int kind = 0; kind < kNumKinds && r.ok; ++kind)
    {
        const bool has = r.get<bool>();
        if (!has) continue;
        switch (kind)
        {
        case kTag: { char name[16]; r.bytes(name, 16); name[15] = 0; s.entities[entity.handle].name = name; } break;
        case kTransform: { const trs t = readTrs(r); entity.addTransform(t.position, t.rotation, t.scale); } break;
        case kDynamicTransform: case kClothRender: break;
        case kRigidBody:
        {
            rigid_body_component rb; r.bytes(&rb, sizeof(rb));
            entity.addRigidBody(rb.invMass == 0.f, rb.gravityFactor, rb.linearDamping, rb.angularDamping);
            entity.rigidBody() = rb;
        } break;
        case kForceField: { float f[3]; r.bytes(f, sizeof(f)); entity.addForceField(b200::mk3(f[0], f[1], f[2])); } break;
        case kPhysicsReference:
        {
            const uint32_t numColliders = r.get<uint32_t>();
            for (uint32_t i = 0; i < numColliders && r.ok; ++i)
            {
                const collider_component c = readCollider(r);
                if (c.type == collider_type_hull && c.hull.geometryIndex >= s.hullGeometries.size()) r.ok = false;      // the geometry must exist in the receiving scene
                if (r.ok) entity.addCollider(c);
            }
            const uint32_t numConstraints = r.get<uint32_t>();
            for (uint32_t i = 0; i < numConstraints && r.ok; ++i)
            {
                const int32_t type = r.get<int32_t>(); const uint32_t ea = r.get<uint32_t>(), eb = r.get<uint32_t>();
                if (type < 0 || type > 5) { r.ok = false; break; }
                uint8_t raw[128]; r.bytes(raw, kJointBytes[type]);
                if (!r.ok || ea >= s.entities.size() || eb >= s.entities.size()) { r.ok = false; break; }
                scene_entity a = { ea, &s }, b = { eb, &s };
                switch (type)
                {
                case constraint_type_distance: addConstraint(a, b, *(const distance_constraint*)raw); break;
                case constraint_type_ball: addConstraint(a, b, *(const ball_constraint*)raw); break;
                case constraint_type_fixed: addConstraint(a, b, *(const fixed_constraint*)raw); break;
                case constraint_type_hinge: addConstraint(a, b, *(const hinge_constraint*)raw); break;
                case constraint_type_cone_twist: addConstraint(a, b, *(const cone_twist_constraint*)raw); break;
                default: addConstraint(a, b, *(const slider_constraint*)raw); break;
                }
            }
            // like in the reference, adding the colliders has recomputed the mass properties of the rigid body read before (scene.h:59-62)
        } break;
        default: r.ok = false; break;       // a component whose size the stream does not carry
        }
    }
    return r.ok && r.offset == size;
}

} // namespace b200host
