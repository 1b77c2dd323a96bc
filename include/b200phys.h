/*
 * b200phys.h -- thin C ABI of the B200-native rigid-body stepper (libb200phys.so).
 *
 * Drop-in boundary for ONE path of pkurth/D3D12Renderer: physicsStepInternal()
 * (reference src/physics/physics.cpp:1180-1362): world-space colliders -> broad phase ->
 * narrow phase -> constraint init -> sequential-impulse iterations -> velocity integration.
 * The host side (C++ mirror of physicsStep(), d3d12renderer_b200/host/) gathers the
 * reference's component pools into the flat arrays declared here, exactly like the
 * reference's own "Physics-Lib" DLL precedent (src/learning/learned_locomotion.cpp:395-488).
 *
 * Conventions
 *  - plain C structs, caller-owned HOST memory, no torch / C++ types in any signature
 *  - every call returns int: 0 = ok, <0 = CUDA error (-(cudaError_t)), >0 = B200PHYS_E_*
 *  - all indices are 32 bit (the reference uses uint16: collision_broad.h:8-13, constraints.h:53-56)
 *  - mat3 is column-major like the reference (src/core/math.h:389-403): m[0..2] is column 0
 *  - quaternions are (x, y, z, w) (src/core/math.h:292-297)
 *  - there is NO CPU fallback: without a CUDA device b200phys_create fails with <0
 */
#ifndef B200PHYS_H
#define B200PHYS_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200PHYS_VERSION 1

/* error codes (>0) */
#define B200PHYS_E_INVALID        1   /* bad argument */
#define B200PHYS_E_CAPACITY       2   /* a device buffer overflowed during the step; see b200phys_stats.overflow_* */
#define B200PHYS_E_NOT_READY      3   /* download before the first step / before upload */

/* collider_type (reference src/physics/physics.h:59-70; the ORDER is load-bearing for pair bucketing) */
enum { B200PHYS_SPHERE = 0, B200PHYS_CAPSULE = 1, B200PHYS_CYLINDER = 2, B200PHYS_AABB = 3, B200PHYS_OBB = 4, B200PHYS_HULL = 5,
       B200PHYS_COLLIDER_TYPE_COUNT = 6 };

/* physics_object_type (reference src/physics/physics.h:49-57) */
enum { B200PHYS_OBJ_RIGID_BODY = 0, B200PHYS_OBJ_STATIC = 1, B200PHYS_OBJ_FORCE_FIELD = 2, B200PHYS_OBJ_TRIGGER = 3,
       B200PHYS_OBJ_NONE = 255 /* inert padding slot (fixed-size halo messages); overlaps nothing */ };

/* constraint_type (reference src/physics/constraints.h:14-28); also the per-iteration kind order (constraints.cpp:3748-3772) */
enum { B200PHYS_JOINT_DISTANCE = 0, B200PHYS_JOINT_BALL = 1, B200PHYS_JOINT_FIXED = 2, B200PHYS_JOINT_HINGE = 3,
       B200PHYS_JOINT_CONE_TWIST = 4, B200PHYS_JOINT_SLIDER = 5, B200PHYS_JOINT_KIND_COUNT = 6 };

/* constraint_motor_type (constraints.h:41-45) */
enum { B200PHYS_MOTOR_VELOCITY = 0, B200PHYS_MOTOR_POSITION = 1 };

#define B200PHYS_NO_TRANSFORM 0xFFFFFFFFu

/* step flags */
#define B200PHYS_STEP_EXACT_ORDER 1u /* Gauss-Seidel in canonical array order, exactly (parity mode) */
#define B200PHYS_STEP_FAST        0u /* colour-sorted order: short dependency chains (the reference's AVX2 path also re-orders:
                                        constraints.cpp:51-184); the order is exported by b200phys_download_solve_order */
#define B200PHYS_STEP_PERMUTED    2u /* like FAST, but a fixed pseudo-random permutation instead of a colouring: cheaper to set up,
                                        longer dependency chains in the sweep */

/* rigid_body_component + physics_transform1_component (reference src/physics/rigid_body.h:18-58). 35 floats. */
typedef struct b200phys_body {
    float rotation[4];          /* physics_transform1.rotation */
    float position[3];          /* physics_transform1.position */
    float localCOGPosition[3];
    float invMass;              /* 0 = kinematic */
    float invInertia[9];        /* local space, column-major */
    float gravityFactor;
    float linearDamping;
    float angularDamping;
    float linearVelocity[3];
    float angularVelocity[3];
    float forceAccumulator[3];
    float torqueAccumulator[3];
} b200phys_body;

/* what a step changes on a body (rigid_body.cpp:126-142): pose + velocities; accumulators are cleared. 13 floats. */
typedef struct b200phys_body_state {
    float rotation[4];
    float position[3];
    float linearVelocity[3];
    float angularVelocity[3];
} b200phys_body_state;

/* pose of a collider parent that is NOT a rigid body (transform_component of a static / trigger / force-field entity,
 * physics.cpp:645-647) */
typedef struct b200phys_transform {
    float rotation[4];
    float position[3];
    float reserved;
} b200phys_transform;

/* collider_union in the PARENT's local space (reference src/physics/physics.h:84-106), 64 bytes.
 * shape[] by type (bounding_volumes.h:30-231):
 *   sphere   : center.xyz, radius
 *   capsule  : positionA.xyz, positionB.xyz, radius
 *   cylinder : positionA.xyz, positionB.xyz, radius
 *   aabb     : minCorner.xyz, maxCorner.xyz
 *   obb      : rotation.xyzw, center.xyz, radius.xyz
 *   hull     : rotation.xyzw, position.xyz, geometryIndex (uint32 bit pattern in shape[7]) */
typedef struct b200phys_collider {
    float    shape[10];
    float    restitution;
    float    friction;
    float    density;
    uint8_t  type;              /* B200PHYS_SPHERE ... */
    uint8_t  objectType;        /* B200PHYS_OBJ_* */
    uint16_t reserved;
    uint32_t objectIndex;       /* rigid body index if objectType == RIGID_BODY, force-field / trigger index otherwise */
    uint32_t parentTransform;   /* for non-rigid-body parents: index into the uploaded transforms, or B200PHYS_NO_TRANSFORM */
} b200phys_collider;

/* bounding_hull_geometry (bounding_volumes.h:208-218) flattened: vertices of hull g are
 * vertices[3*vertexOffsets[g] .. 3*vertexOffsets[g+1]); its local aabb is computed on upload (bounding_volumes.cpp:1393-1397). */

/* --- joints: field order identical to the reference structs (constraints.h) ------------------------------------ */
typedef struct b200phys_distance_joint {    /* distance_constraint, constraints.h:73-80 */
    float localAnchorA[3], localAnchorB[3], globalLength;
} b200phys_distance_joint;

typedef struct b200phys_ball_joint {        /* ball_constraint, constraints.h:129-135 */
    float localAnchorA[3], localAnchorB[3];
} b200phys_ball_joint;

typedef struct b200phys_fixed_joint {       /* fixed_constraint, constraints.h:175-183 */
    float initialInvRotationDifference[4];
    float localAnchorA[3], localAnchorB[3];
    float reserved[2];
} b200phys_fixed_joint;

typedef struct b200phys_hinge_joint {       /* hinge_constraint, constraints.h:229-257; 26 x 4 bytes */
    float localAnchorA[3], localAnchorB[3], localHingeAxisA[3], localHingeAxisB[3];
    float minRotationLimit, maxRotationLimit;
    float maxMotorTorque;
    int32_t motorType;
    float motorVelocityOrTargetAngle;
    float localHingeTangentA[3], localHingeBitangentA[3], localHingeTangentB[3];
} b200phys_hinge_joint;

typedef struct b200phys_cone_twist_joint {  /* cone_twist_constraint, constraints.h:346-380; 30 x 4 bytes */
    float localAnchorA[3], localAnchorB[3];
    float localLimitAxisA[3], localLimitAxisB[3];
    float localLimitTangentA[3], localLimitBitangentA[3], localLimitTangentB[3];
    float swingLimit, twistLimit;
    int32_t swingMotorType;
    float swingMotorVelocityOrTargetAngle;
    float maxSwingMotorTorque;
    float swingMotorAxis;
    int32_t twistMotorType;
    float twistMotorVelocityOrTargetAngle;
    float maxTwistMotorTorque;
} b200phys_cone_twist_joint;

typedef struct b200phys_slider_joint {      /* slider_constraint, constraints.h:497-520 */
    float initialInvRotationDifference[4];
    float localAnchorA[3], localAnchorB[3], localAxisA[3];
    float negDistanceLimit, posDistanceLimit;
    float maxMotorForce;
    int32_t motorType;
    float motorVelocityOrTargetDistance;
    float reserved[2];
} b200phys_slider_joint;

/* constraint_body_pair with 32-bit indices (constraints.h:53-56) */
typedef struct b200phys_body_pair { uint32_t rbA, rbB; } b200phys_body_pair;

/* collider_pair with 32-bit indices (collision_broad.h:8-13). Broad-phase output is canonical: a < b, sorted by (a, b). */
typedef struct b200phys_collider_pair { uint32_t a, b; } b200phys_collider_pair;

/* collision_contact (physics.h:347-354) */
typedef struct b200phys_contact {
    float    point[3];
    float    penetrationDepth;
    float    normal[3];             /* from collider A to collider B of the (type-ordered) pair */
    uint32_t friction_restitution;  /* (uint16 friction << 16) | uint16 restitution, truncating (collision_narrow.cpp:2237) */
} b200phys_contact;

/* Collision begin event (collision_begin_event, physics.h:363-375): contact point and normal averaged over the manifold, relative
 * velocity (B - A) of the two bodies at that point before the solver ran. */
typedef struct b200phys_collision_begin_event
{
    uint32_t colliderA, colliderB;      /* in the narrow phase's type order */
    float    point[3], normal[3], relativeVelocity[3];
} b200phys_collision_begin_event;

typedef struct b200phys_config {
    int32_t  device;                /* CUDA device ordinal */
    uint32_t maxBodies;
    uint32_t maxColliders;
    uint32_t maxTransforms;
    uint32_t maxPairs;              /* broad-phase pair capacity   (0: 16 * maxColliders) */
    uint32_t maxManifolds;          /* colliding pairs capacity    (0: maxPairs) */
    uint32_t maxJointsPerKind;      /* 0: maxBodies */
    uint32_t maxGridCells;          /* uniform grid cell budget    (0: 1<<24) */
    float    cellSize;              /* 0: automatic (largest small-collider AABB extent) */
    float    bigFactor;             /* colliders whose AABB extent exceeds bigFactor * mean extent skip the grid (0: 3) */
    uint32_t maxContacts;           /* contact capacity            (0: 2 * maxManifolds + 1024) */
    uint32_t reserved[5];           /* reserved[0]: fast mode colours the contact graph afresh every this many steps and lets the
                                       manifolds inherit their colour in between (0: 8; 1: colour every step);
                                       reserved[1]: grid cells per largest small extent (0: automatic);
                                       reserved[2]: 1 = no broad-phase candidate reuse (b200phys_get_broadphase_reuse) */
} b200phys_config;

/* the five counters the reference publishes (physics.cpp:1258-1262) + capacities + per-stage device times */
typedef struct b200phys_stats {
    uint32_t numBodies, numColliders;
    uint32_t numBroadphaseOverlaps, numCollisions, numContacts;
    uint32_t numBigColliders, numGridCells, numColors[7]; /* colours/levels per constraint kind, [6] = contacts */
    uint32_t overflowPairs, overflowManifolds, overflowGrid, epaOverflows, gjkIterationCaps, assertCount;
    float    cellSize;
    float    msWorld, msBroad, msNarrow, msInit, msSolve, msIntegrate, msTotal; /* only when timing is enabled */
    uint32_t kernelLaunches;        /* kernels launched by the last step */
    uint32_t reserved[7];           /* [1], [2]: owned bodies / colliders (the rest are ghosts); [3]: overlapping rigid body vs trigger /
                                       force field interactions of the last step; [4]: EPA runs (GJK hits) of the last step,
                                       [5]: those that outgrew the small shared-memory polytope and were re-run in the large one;
                                       [6] / [0]: steps replayed from / captured into a CUDA graph since the context was created */
} b200phys_stats;

typedef struct b200phys_context b200phys_context;

/* lifetime ------------------------------------------------------------------------------------------------------ */
int b200phys_version(void);
const char* b200phys_error_string(int code);
int b200phys_create(const b200phys_config* config, b200phys_context** out);  /* replaces arena scratch (physics.cpp:1205-1212) */
int b200phys_destroy(b200phys_context* ctx);
int b200phys_set_stream(b200phys_context* ctx, void* cudaStream);            /* run on the caller's stream (0 = own stream) */
int b200phys_enable_timing(b200phys_context* ctx, int enabled);
/* CUDA graph replay of b200phys_step (on by default; also off while timing is enabled or with B200PHYS_GRAPH=0 in the environment).
 * The result of a step does not depend on it. */
int b200phys_enable_graphs(b200phys_context* ctx, int enabled);
/* page-locked host memory: uploads / downloads from such buffers are direct DMA transfers (optional, any host memory works) */
void* b200phys_host_alloc(size_t bytes);
void b200phys_host_free(void* p);

/* scene upload (host -> device): replaces the EnTT pool walks of physics.cpp:631-756, 789-805, 1266-1276 ------------
 * Buffer lifetime: uploads are enqueued (cudaMemcpyAsync) on the context's stream and return immediately. Pageable host
 * memory is staged by the driver before the call returns and may be reused at once. PAGE-LOCKED memory (b200phys_host_alloc,
 * cudaHostAlloc) is read by DMA later: it must stay untouched until the next call that synchronises the context has returned
 * (b200phys_synchronize, any b200phys_download_*, b200phys_get_stats). */
/* batched locomotion environment on the device (reference: src/learning/learned_locomotion.{h,cpp}, the RL environment its Physics-Lib
 * DLL exports as getPhysicsStateSize / getPhysicsActionSize / getPhysicsRanges / resetPhysics / updatePhysics, :393-488). The scene
 * (E identical articulated figures + static ground) is uploaded through the normal entry points; this table tells the library where
 * each environment's bodies / joints / colliders are and what its training targets are. Per step only E x actionFloats go up and
 * E x (stateFloats + 2) come down; action smoothing + motor targets, the random pushes, state, reward and failure flag are evaluated
 * on the device (csrc/k_env.cu). actionFloats = 3 * coneTwistsPerEnv + hingesPerEnv, stateFloats = 39 + actionFloats. */
typedef struct b200phys_locomotion_env {
    uint32_t numEnvs, partsPerEnv /* <= 16 */, coneTwistsPerEnv, hingesPerEnv, collidersPerEnv;
    int32_t  parentPart[16];            /* ragdoll.cpp:156-169, -1 = root */
    uint32_t stateParts[6];             /* the body parts of learning_state in its order: left toes, right toes, torso, head, left / right lower arm */
    uint32_t torsoPart, headPart;
    const uint32_t* firstBody;          /* [numEnvs] first rigid body index (parts are consecutive) */
    const uint32_t* firstConeTwist;     /* [numEnvs] first cone-twist joint index */
    const uint32_t* firstHinge;         /* [numEnvs] */
    const uint32_t* firstCollider;      /* [numEnvs] first (lowest) collider index of the environment's collidersPerEnv consecutive colliders */
    const float* initialRotations;      /* [numEnvs][parts][4] reset pose */
    const float* initialPositions;      /* [numEnvs][parts][3] */
    const float* localPositions;        /* [parts][6][3] the six face centres of each part's local bounds (training_locomotion::getLocalPositions) */
    const float* targetPositions;       /* [numEnvs][parts][6][3] */
    const float* targetVelocities;      /* [numEnvs][parts][6][3] */
    const float* targetRotations;       /* [numEnvs][parts][4] local target rotation */
    const float* headTargetHeights;     /* [numEnvs] */
    const unsigned long long* seeds;    /* [numEnvs] xorshift64 state of the push generator, NULL = default per-environment seeds */
} b200phys_locomotion_env;
int b200phys_env_create(b200phys_context* ctx, const b200phys_locomotion_env* desc);   /* after bodies / colliders / joints are uploaded */
int b200phys_env_set_seed(b200phys_context* ctx, uint32_t env, unsigned long long seed);
int b200phys_env_reset(b200phys_context* ctx, const uint32_t* envIndices, uint32_t count);                      /* resetPhysics for these environments */
int b200phys_env_step(b200phys_context* ctx, const float* actions, float dt, uint32_t numSolverIterations, uint32_t flags, float pushProbability,
                      float* outStates, float* outRewards, int32_t* outFailed);                                   /* updatePhysics for all of them */

/* cloth (reference: cloth_component, src/physics/cloth.{h,cpp}): a gridSizeX x gridSizeY particle grid (top row fixed) with stretch / shear /
 * bend constraints, solved by Gauss-Seidel sweeps in the reference's constraint order; physicsStepInternal applies the global force field
 * as wind and simulates every cloth after the rigid bodies (physics.cpp:1354-1358): b200phys_cloth_apply_wind(ctx, 0xFFFFFFFF, force) then
 * b200phys_cloth_simulate(ctx, numClothVelocityIterations, numClothPositionIterations, numClothDriftIterations, dt). */
typedef struct b200phys_cloth_desc {
    float    width, height;
    uint32_t gridSizeX, gridSizeY;
    float    totalMass, stiffness /* 0.5 */, damping /* 0.3 */, gravityFactor /* 1 */;
} b200phys_cloth_desc;
int b200phys_cloth_create(b200phys_context* ctx, const b200phys_cloth_desc* desc, uint32_t* outId);
int b200phys_cloth_count(b200phys_context* ctx, uint32_t* count);
int b200phys_cloth_set_fixed_vertices(b200phys_context* ctx, uint32_t id, const float rotation[4], const float position[3], const float scale[3], int moveRigid); /* setWorldPositionOfFixedVertices */
int b200phys_cloth_set_properties(b200phys_context* ctx, uint32_t id, float totalMass, float stiffness, float damping, float gravityFactor);
int b200phys_cloth_apply_wind(b200phys_context* ctx, uint32_t id /* 0xFFFFFFFF = all */, const float force3[3]);                                                  /* applyWindForce */
int b200phys_cloth_simulate(b200phys_context* ctx, uint32_t velocityIterations, uint32_t positionIterations, uint32_t driftIterations, float dt);                 /* simulate, all cloths */
int b200phys_cloth_download(b200phys_context* ctx, uint32_t id, float* positions3, float* velocities3);                                                           /* gridSizeX * gridSizeY x 3 each, NULL = skip */

/* heightmap collision (reference: heightmap_collider_component, terrain/heightmap_collider.h:132-150, collided against every rigid body
 * collider of type sphere / capsule / AABB / OBB by heightmapCollision, physics/heightmap_collision.cpp:509-618, physics.cpp:1238-1248).
 * heights: chunksPerDim^2 chunks of 129 x 129 uint16 (TERRAIN_LOD_0_VERTICES_PER_DIMENSION), chunk (x, z) at z * chunksPerDim + x, row
 * major in z; world height = minCorner[1] + h * amplitudeScale / 65535. desc == NULL or heights == NULL removes the heightmap. The contacts
 * are appended behind the narrow phase's (one collision {collider, 0xFFFFFFFF} per touching collider, second body = the static dummy);
 * a collider produces one contact per touching triangle (not de-duplicated, heightmap_collision.cpp:432; at most 255): size
 * b200phys_config.maxContacts accordingly. */
typedef struct b200phys_heightmap {
    uint32_t chunksPerDim;
    float    chunkSize;          /* metres per chunk edge (128 segments) */
    float    minCorner[3];
    float    amplitudeScale;
    float    restitution, friction;
} b200phys_heightmap;
int b200phys_upload_heightmap(b200phys_context* ctx, const b200phys_heightmap* desc, const uint16_t* heights);
int b200phys_set_hulls(b200phys_context* ctx, const float* vertices, const uint32_t* vertexOffsets, uint32_t numHulls);
int b200phys_upload_bodies(b200phys_context* ctx, const b200phys_body* bodies, uint32_t count);
int b200phys_upload_transforms(b200phys_context* ctx, const b200phys_transform* transforms, uint32_t count);
int b200phys_upload_colliders(b200phys_context* ctx, const b200phys_collider* colliders, uint32_t count);
int b200phys_upload_joints(b200phys_context* ctx, int kind, const void* joints, const b200phys_body_pair* pairs, uint32_t count);
/* pose + velocities only (the per-step mutable state, same struct b200phys_download_bodies writes); mass properties, damping and
 * accumulators keep their uploaded values. count must equal the uploaded body count */
int b200phys_upload_body_states(b200phys_context* ctx, const b200phys_body_state* states, uint32_t count);
/* accumulators only (what gameplay code mutates between steps: physics.cpp:626-627); count must equal the uploaded body count.
 * The transfer runs on the context's copy stream and overlaps the collision detection of the next b200phys_step, which waits for
 * it right before the force integration: with page-locked host memory (b200phys_host_alloc) the buffer must stay untouched until
 * that step (or b200phys_synchronize) has returned; pageable memory is staged by the driver before this call returns. */
int b200phys_upload_forces(b200phys_context* ctx, const float* forceTorque6, uint32_t count);

/* the step: physicsStepInternal (physics.cpp:1180-1362) ------------------------------------------------------------- */
int b200phys_step(b200phys_context* ctx, float dt, uint32_t numSolverIterations, uint32_t flags);
int b200phys_synchronize(b200phys_context* ctx);

/* results (device -> host) --------------------------------------------------------------------------------------- */
int b200phys_download_bodies(b200phys_context* ctx, b200phys_body_state* out, uint32_t count);
/* Force fields and triggers (reference: handleNonCollisionInteractions, physics.cpp:952-1035; getForceFieldStates :759-787).
 * A collider with objectType FORCE_FIELD / TRIGGER carries the index of its force field / trigger in objectIndex. Every step the
 * rigid-body colliders overlapping such a collider (boolean shape tests, collision_narrow.cpp:1593-1689) are collected on the device:
 * a localized force field adds its force to the accumulator of each overlapping body (once per overlapping collider, like the
 * reference) before the forces are integrated; the global force is added to every body. Trigger overlaps are de-duplicated per
 * (trigger, rigid body) and diffed against the previous call: download_trigger_events must be called once after every step. */
typedef struct b200phys_trigger_event { uint32_t triggerIndex, rigidBodyIndex; } b200phys_trigger_event;
int b200phys_upload_force_fields(b200phys_context* ctx, const float* forces3, uint32_t count);   /* world-space force per localized field */
int b200phys_set_global_force_field(b200phys_context* ctx, const float force3[3]);
int b200phys_download_trigger_events(b200phys_context* ctx, b200phys_trigger_event* enterOut, uint32_t enterCapacity, uint32_t* numEnter,
                                     b200phys_trigger_event* leaveOut, uint32_t leaveCapacity, uint32_t* numLeave);
/* Collision begin / end events (reference: handleCollisionCallbacks, physics.cpp:1037-1178). Off by default; enabling allocates the
 * per-manifold key lists. After every step: the collider pairs that collide now but did not in the previous step (begin, with the
 * averaged contact) and those that stopped colliding (end), each list ordered by (type bucket, lower collider, higher collider).
 * The first enabled step reports every collision as a begin. Trigger / force-field overlap events are not produced yet. */
int b200phys_enable_events(b200phys_context* ctx, int enable);
int b200phys_download_events(b200phys_context* ctx, b200phys_collision_begin_event* beginOut, uint32_t beginCapacity, uint32_t* numBegin,
                             b200phys_collider_pair* endOut, uint32_t endCapacity, uint32_t* numEnd);
int b200phys_get_stats(b200phys_context* ctx, b200phys_stats* out);
/* Broad-phase candidate reuse (no counterpart in the reference, whose sweep-and-prune sorts afresh every step, collision_broad.cpp:87-166;
 * the OUTPUT is the same pair list either way). The grid pipeline runs on AABBs grown by a margin and leaves a candidate list; while every
 * collider's AABB stays inside its grown box a step only filters that list by the exact overlap predicate. On by default;
 * b200phys_config.reserved[2] = 1 or B200PHYS_BP_REUSE=0 turns it off (the pipeline then runs in every step).
 * out[0] steps, out[1] steps that ran the grid pipeline, out[2] candidates in the current list, out[3] pairs of the last step,
 * out[4] margin of the current list (float bits; 0: the last build was exact), out[5] age of the list in steps, out[6] largest movement of
 * an AABB face per step since the last build (float bits), out[7] mean AABB extent of the grid colliders (float bits). Synchronises the stream. */
int b200phys_get_broadphase_reuse(b200phys_context* ctx, uint32_t out[8]);
/* Per-kernel-group device times of the LAST step (timing must be enabled): *count intervals in launch order; names are written as one
 * '\n'-separated string (kernel group names such as "bp_query", "np_epa", "solve"), msOut[i] is the CUDA-event time of interval i.
 * The analogue of the reference's CPU_PROFILE_BLOCK tree (physics.cpp:1182, collision_broad.cpp:299, ...), per kernel group. */
int b200phys_get_timeline(b200phys_context* ctx, char* namesOut, uint32_t namesCapacity, float* msOut, uint32_t msCapacity, uint32_t* count);
/* parity / event-layer taps (each returns the element count through *count; out may be NULL to query the count) */
int b200phys_download_aabbs(b200phys_context* ctx, float* outMinMax6, uint32_t* count);
int b200phys_download_pairs(b200phys_context* ctx, b200phys_collider_pair* out, uint32_t* count);
int b200phys_download_contacts(b200phys_context* ctx, b200phys_contact* outContacts, b200phys_body_pair* outBodyPairs, uint32_t* count);
int b200phys_download_collisions(b200phys_context* ctx, b200phys_collider_pair* outPairs, uint8_t* outContactCounts, uint32_t* count);
/* order in which the solver visited contacts / joints of one kind inside an iteration (kind 6 = contacts) */
int b200phys_download_solve_order(b200phys_context* ctx, int kind, uint32_t* outOrder, uint32_t* count);

/* multi-GPU (one context per GPU / rank, bodies sharded by spatial slab; transport is the caller's: NCCL send/recv on these
 * DEVICE buffers, see d3d12renderer_b200/parallel.py) ------------------------------------------------------------------------
 * halo record formats: body = b200phys_body (35 floats) + owner-local index (36 x 4 bytes); collider = b200phys_collider whose
 * objectIndex is the position of its body inside the packed body list (64 bytes). */
uint32_t b200phys_halo_record_bytes(int which /* 0 body, 1 collider */);
/* packs every OWNED body (with all its colliders) that has a collider AABB overlapping [lo, hi) along `axis` into the caller's
 * device buffers; outCounts[2] = {bodies, colliders} (synchronous). Returns B200PHYS_E_CAPACITY if a buffer is too small. */
int b200phys_halo_pack(b200phys_context* ctx, int axis, float lo, float hi, void* dstBodiesDevice, void* dstCollidersDevice,
                       uint32_t capBodies, uint32_t capColliders, uint32_t* outCounts);
/* appends received ghost bodies / colliders behind the owned ones; they take part in the next step, their results are discarded */
int b200phys_halo_unpack(b200phys_context* ctx, const void* srcBodiesDevice, uint32_t numBodies, const void* srcCollidersDevice, uint32_t numColliders);
/* Fixed-size halo messages (no host round trip): [uint32 numBodies, numColliders, 0, 0][capBodies body records][capColliders
 * collider records]. pack_message fills a device buffer asynchronously on the context's stream; unpack_message appends ALL
 * capBodies + capColliders slots behind the current bodies / colliders, the ones beyond the counts in the header as inert
 * padding (B200PHYS_OBJ_NONE), so the host never needs the counts. A header count above the capacity raises the sticky
 * overflow flag (b200phys_get_stats -> B200PHYS_E_CAPACITY, overflowGrid bit 16). */
uint64_t b200phys_halo_message_bytes(uint32_t capBodies, uint32_t capColliders);
int b200phys_halo_pack_message(b200phys_context* ctx, int axis, float lo, float hi, void* dstMessageDev, uint32_t capBodies, uint32_t capColliders);
int b200phys_halo_unpack_message(b200phys_context* ctx, const void* srcMessageDev, uint32_t capBodies, uint32_t capColliders);
/* Migration (multi-GPU): bodies whose position left [lo, hi) on `axis` move, with all their colliders, into two messages of the
 * halo message layout (down = below lo, up = at or above hi) and are removed from this context (the body and collider arrays
 * are compacted, order preserved). Synchronous; requires no ghosts (b200phys_halo_clear first) and no joints.
 * outDest (host, one entry per previously owned body, may be NULL): 0 stays / 1 down / 2 up, so that callers can carry per-body
 * data of their own along. outCounts[4] = {bodies down, colliders down, bodies up, colliders up}.
 * migrate_unpack appends a received message as OWNED bodies / colliders; outCounts[2] = {bodies, colliders}. */
int b200phys_migrate_pack(b200phys_context* ctx, int axis, float lo, float hi, void* dstMessageDownDev, void* dstMessageUpDev,
                          uint32_t capBodies, uint32_t capColliders, uint32_t* outDest, uint32_t* outCounts);
int b200phys_migrate_unpack(b200phys_context* ctx, const void* srcMessageDev, uint32_t capBodies, uint32_t capColliders, uint32_t* outCounts);
int b200phys_halo_clear(b200phys_context* ctx);

/* Multi-GPU with the transport inside the library (what a C++ host uses; SURVEY 8b `_create_distributed`): one process / thread and one
 * context per GPU, bodies sharded by slab along `axis`, this rank owning [slabLo, slabHi); ranks are ordered along the axis (rank r's
 * upper neighbour is r + 1). The ghost halo is exchanged with ncclSend / ncclRecv on the context's stream inside b200phys_step_distributed.
 * Rank 0 obtains the 128-byte NCCL id with b200phys_dist_unique_id and the host hands it to every rank (MPI_Bcast, a file, a socket).
 * NCCL (libnccl.so.2) is loaded at run time; without it these calls return B200PHYS_E_NOT_READY. */
typedef struct b200phys_dist_config {
    int32_t  rank, nranks;
    int32_t  axis;                          /* 0 x, 1 y, 2 z */
    float    slabLo, slabHi;                /* -3e38 / +3e38 on the outermost ranks */
    float    ghostWidth;                    /* >= the largest collider extent (+ the distance a body travels between two migrations) */
    uint32_t maxGhostBodies, maxGhostColliders;   /* per neighbour; maxBodies / maxColliders of the config count the OWNED ones */
} b200phys_dist_config;
int b200phys_dist_unique_id(uint8_t* out128);
int b200phys_create_distributed(const b200phys_config* config, const b200phys_dist_config* dist, const uint8_t* uniqueId128, b200phys_context** out);
int b200phys_step_distributed(b200phys_context* ctx, float dt, uint32_t numSolverIterations, uint32_t flags);
int b200phys_dist_ghost_counts(b200phys_context* ctx, uint32_t* outBodies, uint32_t* outColliders);

#ifdef __cplusplus
}
#endif
#endif /* B200PHYS_H */
