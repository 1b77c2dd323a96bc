// physics_host.h -- host-side C++ mirror of the reference's physics API surface for the rigid-body step
// (pkurth/D3D12Renderer src/physics/physics.h, rigid_body.h, constraints.h, bounding_volumes.h, ragdoll.h).
//
// Same names, argument meaning and index conventions as the reference so that scene-building code and the call
// `physicsStep(scene, arena, timer, settings, dt)` read like the original; the step itself is NOT computed here: it is
// gathered into the flat arrays of include/b200phys.h and executed by the CUDA library through the C ABI.
// The reference stores components in EnTT pools; this mirror uses flat vectors with the same iteration-order semantics
// (SURVEY.md 3.5): collider index = reverse creation order, rigid-body index = creation (storage) order, joint arrays in
// storage order. INTEGRATION.md shows the gather/scatter a maintainer would write against real EnTT pools instead.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>
#include <cstring>
#include <thread>
#include <iosfwd>
#include "../csrc/b200_math.cuh"       // the same fp32 helpers the kernels use (host + device)
#include "../../include/b200phys.h"

namespace b200host {

using vec3 = b200::f3;
using quat = b200::q4;
using mat3 = b200::m3;

struct trs { quat rotation; vec3 position; vec3 scale; };   // math.h:494-504
inline trs makeTrs(vec3 position, quat rotation = b200::qIdentity(), vec3 scale = b200::mk3(1.f)) { trs t; t.rotation = rotation; t.position = position; t.scale = scale; return t; }

#define GRAVITY -9.81f   /* physics.h:11 */

// physics.h:20-47
enum physics_material_type { physics_material_type_none = -1, physics_material_type_metal, physics_material_type_wood, physics_material_type_flesh, physics_material_type_count };
struct physics_material { physics_material_type type; float restitution; float friction; float density; };

// physics.h:49-70
enum physics_object_type : uint8_t { physics_object_type_rigid_body, physics_object_type_static_collider, physics_object_type_force_field, physics_object_type_trigger, physics_object_type_count };
enum collider_type : uint8_t { collider_type_sphere, collider_type_capsule, collider_type_cylinder, collider_type_aabb, collider_type_obb, collider_type_hull, collider_type_count };

// bounding_volumes.h:30-231
struct bounding_sphere { vec3 center; float radius; };
struct bounding_capsule { vec3 positionA; vec3 positionB; float radius; };
struct bounding_cylinder { vec3 positionA; vec3 positionB; float radius; };
struct bounding_box
{
    vec3 minCorner, maxCorner;
    static bounding_box fromCenterRadius(vec3 center, vec3 radius) { bounding_box b; b.minCorner = center - radius; b.maxCorner = center + radius; return b; }
    static bounding_box fromMinMax(vec3 mn, vec3 mx) { bounding_box b; b.minCorner = mn; b.maxCorner = mx; return b; }
};
struct bounding_oriented_box { quat rotation; vec3 center; vec3 radius; };
struct bounding_hull { quat rotation; vec3 position; uint32_t geometryIndex; };
struct indexed_triangle32 { uint32_t a, b, c; };

// physics.h:13-18
struct physics_properties { mat3 inertia; vec3 cog; float mass; };

typedef uint32_t entity_handle;
static const entity_handle null_entity = 0xFFFFFFFFu;

// physics.h:84-171
struct collider_union
{
    union { bounding_sphere sphere; bounding_capsule capsule; bounding_cylinder cylinder; bounding_box aabb; bounding_oriented_box obb; bounding_hull hull; };
    physics_material material;
    collider_type type;
    collider_union() {}
    physics_properties calculatePhysicsProperties(const struct game_scene& scene) const;   // physics.cpp:1416-1588
};
struct collider_component : collider_union
{
    static collider_component asSphere(bounding_sphere s, physics_material m) { collider_component r; r.sphere = s; r.type = collider_type_sphere; r.material = m; return r; }
    static collider_component asCapsule(bounding_capsule c, physics_material m) { collider_component r; r.capsule = c; r.type = collider_type_capsule; r.material = m; return r; }
    static collider_component asCylinder(bounding_cylinder c, physics_material m) { collider_component r; r.cylinder = c; r.type = collider_type_cylinder; r.material = m; return r; }
    static collider_component asAABB(bounding_box b, physics_material m) { collider_component r; r.aabb = b; r.type = collider_type_aabb; r.material = m; return r; }
    static collider_component asOBB(bounding_oriented_box b, physics_material m) { collider_component r; r.obb = b; r.type = collider_type_obb; r.material = m; return r; }
    static collider_component asHull(bounding_hull h, physics_material m) { collider_component r; r.hull = h; r.type = collider_type_hull; r.material = m; return r; }
    entity_handle parentEntity = null_entity;
    entity_handle nextEntity = null_entity;     // next collider of the same parent (most recently added first, scene.h:53-56)
};

// rigid_body.h:18-46
struct rigid_body_component
{
    rigid_body_component() : rigid_body_component(true, 1.f) {}
    rigid_body_component(bool kinematic, float gravityFactor = 1.f, float linearDamping = 0.4f, float angularDamping = 0.4f);
    vec3 localCOGPosition; float invMass; mat3 invInertia;
    float gravityFactor, linearDamping, angularDamping;
    vec3 linearVelocity, angularVelocity, forceAccumulator, torqueAccumulator;
};

// constraints.h:41-45
enum constraint_motor_type { constraint_velocity_motor, constraint_position_motor };
// constraints.h:73-80,129-135,175-183,229-257,346-380,497-520 : the boundary structs ARE these structs (same field order)
typedef b200phys_distance_joint distance_constraint;
typedef b200phys_ball_joint ball_constraint;
typedef b200phys_fixed_joint fixed_constraint;
typedef b200phys_hinge_joint hinge_constraint;
typedef b200phys_cone_twist_joint cone_twist_constraint;
typedef b200phys_slider_joint slider_constraint;
// physics.h:209-214: a handle names the constraint ENTITY; here a stable id (constraints move inside their pool when another one
// is deleted: EnTT swap-and-pop, mirrored by deleteConstraint)
struct distance_constraint_handle { uint32_t entity; };
struct ball_constraint_handle { uint32_t entity; };
struct fixed_constraint_handle { uint32_t entity; };
struct hinge_constraint_handle { uint32_t entity; };
struct cone_twist_constraint_handle { uint32_t entity; };
struct slider_constraint_handle { uint32_t entity; };
enum constraint_type { constraint_type_distance, constraint_type_ball, constraint_type_fixed, constraint_type_hinge, constraint_type_cone_twist, constraint_type_slider, constraint_type_collision, constraint_type_count }; // constraints.h:14-28

// physics.h:291-345: triggers and force fields
enum trigger_event_type { trigger_event_enter, trigger_event_leave };
struct trigger_event { entity_handle trigger, other; trigger_event_type type; };
struct trigger_component { std::function<void(const trigger_event&)> callback; };
struct force_field_component { vec3 force; };

// physics.h:356-380: collision events handed to the callbacks of physics_settings
struct game_scene;
struct collision_begin_event
{
    entity_handle entityA, entityB;                  // parents of the two colliders (scene_entity in the reference)
    const collider_component& colliderA; const collider_component& colliderB;
    vec3 position, normal, relativeVelocity;
};
struct collision_end_event
{
    entity_handle entityA, entityB;
    const collider_component& colliderA; const collider_component& colliderB;
};
typedef std::function<void(const collision_begin_event&)> collision_begin_event_func;
typedef std::function<void(const collision_end_event&)> collision_end_event_func;

// terrain/heightmap_collider.h:132-150: the terrain's collision representation, chunksPerDim^2 chunks of 129 x 129 uint16 heights
// (heightmap_collider_chunk::setHeights keeps a pointer per chunk; here the component owns one contiguous copy, chunk (x, z) at
// z * chunksPerDim + x). update() moves / rescales it like the reference (heightmap_collider.cpp:14-19).
struct heightmap_collider_component
{
    heightmap_collider_component(uint32_t chunksPerDim, float chunkSize, physics_material material) : chunksPerDim(chunksPerDim), chunkSize(chunkSize), material(material)
    { heights.assign((size_t)chunksPerDim * chunksPerDim * 129u * 129u, 0); }
    void update(vec3 minCorner, float amplitudeScale) { this->minCorner = minCorner; this->amplitudeScale = amplitudeScale; dirty = true; }
    void setHeights(uint32_t x, uint32_t z, const uint16_t* h) { memcpy(heights.data() + (size_t)(z * chunksPerDim + x) * 129u * 129u, h, sizeof(uint16_t) * 129u * 129u); dirty = true; }
    uint32_t chunksPerDim; float chunkSize; physics_material material;
    vec3 minCorner = b200::mk3(0.f); float amplitudeScale = 1.f;
    std::vector<uint16_t> heights; bool dirty = true;
};

// cloth.h:5-55: the parameters live here, the particles on the device (b200phys_cloth_*); `positions` is refreshed after every simulate
// (what cloth_render_component reads)
struct cloth_component
{
    cloth_component(float width, float height, uint32_t gridSizeX, uint32_t gridSizeY, float totalMass, float stiffness = 0.5f, float damping = 0.3f, float gravityFactor = 1.f)
        : totalMass(totalMass), gravityFactor(gravityFactor), damping(damping), stiffness(stiffness), gridSizeX(gridSizeX), gridSizeY(gridSizeY), width(width), height(height) {}
    void setWorldPositionOfFixedVertices(const trs& transform, bool moveRigid = false) { pendingFixed.push_back({ transform, moveRigid }); }
    float totalMass, gravityFactor, damping, stiffness;
    uint32_t gridSizeX, gridSizeY; float width, height;
    std::vector<vec3> positions;
    struct fixed_update { trs transform; bool moveRigid; };
    std::vector<fixed_update> pendingFixed;
    int32_t deviceId = -1; float upTotalMass = 0.f, upStiffness = 0.f, upDamping = 0.f, upGravityFactor = 0.f;      // what the device has
};

// physics.h:382-400
struct physics_settings
{
    bool fixedFrameRate = true;
    uint32_t frameRate = 120;
    uint32_t maxPhysicsIterationsPerFrame = 4;
    uint32_t numRigidSolverIterations = 30;
    uint32_t numClothVelocityIterations = 0, numClothPositionIterations = 1, numClothDriftIterations = 0;
    bool simdBroadPhase = true, simdNarrowPhase = true, simdConstraintSolver = true;   // accepted, ignored: one CUDA path
    bool exactConstraintOrder = false;   // extension: B200PHYS_STEP_EXACT_ORDER (canonical Gauss-Seidel order)
    collision_begin_event_func collisionBeginCallback;   // physics.h:398-399; called in the reference's order (sorted by collider pair)
    collision_end_event_func collisionEndCallback;
};

// independent per-element loops (gather / scatter / per-environment work) spread over the host cores
template <typename F> inline void parallelFor(uint32_t n, F&& f, uint32_t minParallel = 4096u)
{
    const uint32_t hw = std::thread::hardware_concurrency();
    const uint32_t workers = n < minParallel ? 1u : (hw ? (hw > 32u ? 32u : hw) : 1u);
    if (workers <= 1) { for (uint32_t i = 0; i < n; ++i) f(i); return; }
    std::vector<std::thread> pool;
    const uint32_t chunk = (n + workers - 1) / workers;
    for (uint32_t w = 0; w < workers; ++w)
    {
        const uint32_t b = w * chunk, e = b + chunk < n ? b + chunk : n;
        if (b >= e) break;
        pool.emplace_back([b, e, &f]() { for (uint32_t i = b; i < e; ++i) f(i); });
    }
    for (std::thread& t : pool) t.join();
}

struct memory_arena {};   // the reference's per-step scratch source (core/memory.h:52-112); device scratch lives in the context

struct bounding_hull_geometry { std::vector<vec3> vertices; std::vector<indexed_triangle32> faces; };

struct game_scene;
struct scene_entity
{
    entity_handle handle = null_entity; game_scene* scene = nullptr;
    explicit operator bool() const { return handle != null_entity; }
    scene_entity& addTransform(vec3 position, quat rotation = b200::qIdentity(), vec3 scale = b200::mk3(1.f));                 // addComponent<transform_component>
    scene_entity& addCollider(const collider_component& c);                                                                     // addComponent<collider_component>
    scene_entity& addRigidBody(bool kinematic, float gravityFactor = 1.f, float linearDamping = 0.4f, float angularDamping = 0.4f); // addComponent<rigid_body_component>
    scene_entity& addTrigger(std::function<void(const trigger_event&)> callback);   // addComponent<trigger_component>
    scene_entity& addForceField(vec3 force);                                        // addComponent<force_field_component>
    rigid_body_component& rigidBody();
    trs& transform();
};

struct game_scene
{
    struct entity_data
    {
        std::string name; bool hasTransform = false; trs transform;
        int32_t rigidBody = -1;                 // index into rigidBodies (storage order)
        int32_t trigger = -1, forceField = -1;  // index into triggers / forceFields
        entity_handle firstColliderEntity = null_entity; uint32_t numColliders = 0;
    };
    std::vector<entity_data> entities;
    std::vector<rigid_body_component> rigidBodies; std::vector<entity_handle> rigidBodyEntity;
    std::vector<trs> physicsTransform0, physicsTransform1;      // per rigid body (rigid_body.h:48-58)
    std::vector<collider_component> colliders;                  // creation order; world index = reverse (physics.cpp:635-641)
    bool eventsEnabled = false;                                 // collision events switched on in the device context
    std::vector<trigger_component> triggers; std::vector<entity_handle> triggerEntity;
    std::vector<force_field_component> forceFields; std::vector<entity_handle> forceFieldEntity;
    std::vector<bounding_hull_geometry> hullGeometries;
    std::vector<cloth_component> cloths;                         // scene.view<cloth_component>() (physics.cpp:1354)
    vec3 lastGlobalForceField = b200::mk3(0.f);                  // what applyWindForce receives
    std::vector<heightmap_collider_component> heightmaps;       // scene.view<heightmap_collider_component>() (physics.cpp:1238); the device holds one
    std::vector<distance_constraint> distanceConstraints; std::vector<ball_constraint> ballConstraints; std::vector<fixed_constraint> fixedConstraints;
    std::vector<hinge_constraint> hingeConstraints; std::vector<cone_twist_constraint> coneTwistConstraints; std::vector<slider_constraint> sliderConstraints;
    std::vector<b200phys_body_pair> constraintEntities[6];      // (entityA, entityB) per joint, storage order
    std::vector<uint32_t> constraintIds[6];                     // stable id of the joint at each storage position (the constraint entity)
    std::vector<int32_t> constraintIndexOfId[6];                // id -> storage position, -1 once deleted

    scene_entity createEntity(const char* name);
    uint32_t allocateBoundingHullGeometry(const vec3* vertices, uint32_t numVertices, const indexed_triangle32* triangles, uint32_t numTriangles); // physics.h:207
    void recalculateProperties(entity_handle e);               // rigid_body.cpp:29-81

    // ---- boundary (what the EnTT shim of INTEGRATION.md does with real pools)
    void gatherBodies(std::vector<b200phys_body>& out) const;
    void gatherBodies(b200phys_body* out) const;                 // out holds rigidBodies.size() entries
    const void* jointStructs(int kind, size_t& count) const;    // the constraint structs ARE the boundary structs: no copy
    void refreshJointPairs();
    void gatherColliders(std::vector<b200phys_collider>& outColliders, std::vector<b200phys_transform>& outTransforms) const;
    void gatherJoints(int kind, std::vector<uint8_t>& outStructs, std::vector<b200phys_body_pair>& outPairs) const;
    void gatherHulls(std::vector<float>& verts, std::vector<uint32_t>& offsets) const;
    void scatterBodyStates(const b200phys_body_state* states, uint32_t count);

    b200phys_context* ctx = nullptr;  // lazily created device context (registry context variable in the reference's terms)
    bool collidersDirty = true, jointsDirty = true, hullsDirty = true;
    b200phys_stats lastStats;
    // page-locked staging reused from step to step (gather -> upload, download -> scatter) and the cached joint body pairs
    b200phys_body* stageBodies = nullptr; b200phys_body_state* stageStates = nullptr; uint32_t stageCapacity = 0;
    std::vector<b200phys_body_pair> jointPairs[6]; bool jointPairsDirty = true;
    ~game_scene();
};

// physics.h:217-235 (entities must have transform_component; anchors in world space)
distance_constraint_handle addDistanceConstraintFromLocalPoints(scene_entity& a, scene_entity& b, vec3 localAnchorA, vec3 localAnchorB, float distance);
distance_constraint_handle addDistanceConstraintFromGlobalPoints(scene_entity& a, scene_entity& b, vec3 globalAnchorA, vec3 globalAnchorB);
ball_constraint_handle addBallConstraintFromLocalPoints(scene_entity& a, scene_entity& b, vec3 localAnchorA, vec3 localAnchorB);
ball_constraint_handle addBallConstraintFromGlobalPoints(scene_entity& a, scene_entity& b, vec3 globalAnchor);
fixed_constraint_handle addFixedConstraintFromGlobalPoints(scene_entity& a, scene_entity& b, vec3 globalAnchor);
hinge_constraint_handle addHingeConstraintFromGlobalPoints(scene_entity& a, scene_entity& b, vec3 globalAnchor, vec3 globalHingeAxis, float minLimit = 1.f, float maxLimit = -1.f);
cone_twist_constraint_handle addConeTwistConstraintFromGlobalPoints(scene_entity& a, scene_entity& b, vec3 globalAnchor, vec3 globalAxis, float swingLimit, float twistLimit);
slider_constraint_handle addSliderConstraintFromGlobalPoints(scene_entity& a, scene_entity& b, vec3 globalAnchor, vec3 globalAxis, float minLimit = 1.f, float maxLimit = -1.f);

// physics.h:237-264
distance_constraint_handle addConstraint(scene_entity& a, scene_entity& b, const distance_constraint& c);
ball_constraint_handle addConstraint(scene_entity& a, scene_entity& b, const ball_constraint& c);
fixed_constraint_handle addConstraint(scene_entity& a, scene_entity& b, const fixed_constraint& c);
hinge_constraint_handle addConstraint(scene_entity& a, scene_entity& b, const hinge_constraint& c);
cone_twist_constraint_handle addConstraint(scene_entity& a, scene_entity& b, const cone_twist_constraint& c);
slider_constraint_handle addConstraint(scene_entity& a, scene_entity& b, const slider_constraint& c);
distance_constraint& getConstraint(game_scene& scene, distance_constraint_handle handle);
ball_constraint& getConstraint(game_scene& scene, ball_constraint_handle handle);
fixed_constraint& getConstraint(game_scene& scene, fixed_constraint_handle handle);
hinge_constraint& getConstraint(game_scene& scene, hinge_constraint_handle handle);
cone_twist_constraint& getConstraint(game_scene& scene, cone_twist_constraint_handle handle);
slider_constraint& getConstraint(game_scene& scene, slider_constraint_handle handle);
void deleteAllConstraints(game_scene& scene);
void deleteConstraint(game_scene& scene, distance_constraint_handle handle);
void deleteConstraint(game_scene& scene, ball_constraint_handle handle);
void deleteConstraint(game_scene& scene, fixed_constraint_handle handle);
void deleteConstraint(game_scene& scene, hinge_constraint_handle handle);
void deleteConstraint(game_scene& scene, cone_twist_constraint_handle handle);
void deleteConstraint(game_scene& scene, slider_constraint_handle handle);
void deleteAllConstraintsFromEntity(scene_entity& entity);

// physics.h:273-337: range-for over the colliders of an entity (most recently added first) and over its constraints
struct collider_component_iterator
{
    explicit collider_component_iterator(scene_entity e) : scene(e.scene), first(e ? e.scene->entities[e.handle].firstColliderEntity : null_entity) {}
    struct iterator
    {
        game_scene* scene; entity_handle collider;
        friend bool operator!=(const iterator& a, const iterator& b) { return a.collider != b.collider; }
        iterator& operator++() { collider = scene->colliders[collider].nextEntity; return *this; }
        collider_component& operator*() { return scene->colliders[collider]; }
    };
    iterator begin() { return iterator{ scene, first }; }
    iterator end() { return iterator{ scene, null_entity }; }
    game_scene* scene; entity_handle first;
};
struct constraint_entity_iterator
{   // yields (other entity, constraint_type, storage index) for every joint attached to the entity; order: by kind, then storage order
    // (the reference walks a per-entity edge list, physics.cpp:87-126: most recently added first -- the SET is the same)
    struct item { scene_entity other; constraint_type type; uint32_t index; };
    explicit constraint_entity_iterator(scene_entity e);
    std::vector<item> items;
    std::vector<item>::iterator begin() { return items.begin(); }
    std::vector<item>::iterator end() { return items.end(); }
};

// vehicle.h:5-40 (16 parts; hinge joints with a velocity motor (engine) and a position motor (steering wheel), hinge limits, one fixed,
// one slider with limits, four ball joints): the acceptance scene for the joint kinds the ragdoll does not use
struct vehicle
{
    enum { motor, motorGear, driveAxis, frontAxis, steeringWheel, steeringAxis, leftWheelSuspension, rightWheelSuspension, leftFrontWheel, rightFrontWheel,
           leftWheelArm, rightWheelArm, differentialSunGear, differentialSpiderGear, leftRearWheel, rightRearWheel, numParts };
    scene_entity parts[numParts];
    hinge_constraint_handle motorConstraint, steeringWheelConstraint;
    void initialize(game_scene& scene, vec3 initialMotorPosition, float initialRotation = 0.f);   // vehicle.cpp:303-497
    static vehicle create(game_scene& scene, vec3 initialMotorPosition, float initialRotation = 0.f) { vehicle v; v.initialize(scene, initialMotorPosition, initialRotation); return v; }
};

// ragdoll.h:7-60 (14 bodies, 17 colliders, 7 cone-twist + 6 hinge joints)
struct humanoid_ragdoll
{
    scene_entity bodyParts[14];
    void initialize(game_scene& scene, vec3 initialHipPosition, float initialRotation = 0.f);   // ragdoll.cpp:10-135
    static humanoid_ragdoll create(game_scene& scene, vec3 initialHipPosition, float initialRotation = 0.f) { humanoid_ragdoll r; r.initialize(scene, initialHipPosition, initialRotation); return r; }
};

// physics.h:405 : returns 0 or the C ABI error code of the failing call (the reference returns void)
// scene files (physics components of the reference's YAML scenes, src/scene/serialization_yaml.cpp): scene_yaml.cpp
bool saveSceneYaml(const game_scene& scene, std::ostream& out, const char* sceneName = nullptr);
int loadSceneYaml(game_scene& scene, std::istream& in);      // entities created, -1 if the input has no "Entities" list

// physics.h:404: a ray pushes the first rigid body it hits (force + torque accumulators for the next step); ray_interaction.cpp
void testPhysicsInteraction(game_scene& scene, vec3 rayOrigin, vec3 rayDirection, float strength = 1000.f, uint32_t firstBody = 0, uint32_t numBodies = 0xFFFFFFFFu,
                            uint32_t firstCollider = 0, uint32_t numColliders = 0xFFFFFFFFu);   // body / collider (creation order) ranges restrict the test (batched environments)

int physicsStep(game_scene& scene, memory_arena& arena, float& timer, const physics_settings& settings, float dt);
// scene/serialization_binary.h: the binary entity stream (physics components), scene_binary.cpp
uint64_t serializeEntityToMemory(scene_entity entity, void* memory, uint64_t maxSize);
bool deserializeEntityFromMemory(scene_entity entity, const void* memory, uint64_t size);
int uploadSceneToDevice(game_scene& scene);     // creates the device context if needed and uploads whatever changed, without stepping

} // namespace b200host
