// locomotion_env.cpp -- the reference's RL environment (src/learning/learned_locomotion.{h,cpp}) over the host mirror, batched:
// E humanoid ragdolls stand on one ground box, far enough apart never to touch, and ONE physicsStep (one b200phys_step on the GPU)
// advances all of them; state, reward and failure test are the reference's formulas evaluated per environment on the host from the
// downloaded poses. The single-environment C interface the reference exports to Python (learned_locomotion.cpp:393-488:
// getPhysicsStateSize / getPhysicsActionSize / getPhysicsRanges / resetPhysics / updatePhysics) is the E = 1 case.
// The random pushes (a ray cast at a random body part with probability 0.02 per step, :448-458) use one xorshift64 stream per
// environment (the reference seeds its single stream with the wall clock) and only test the environment's own bodies.
#include "physics_host.h"
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

namespace b200host {

namespace {

const uint32_t kParts = 14, kConeTwists = 7, kHinges = 6;
const uint32_t kActionFloats = kConeTwists * 3 + kHinges;            // learning_action (learned_locomotion.h:19-35)
const uint32_t kStateFloats = 13 * 3 + kActionFloats;                // learning_state (:44-67)
enum { torso, head, lUpperArm, lLowerArm, rUpperArm, rLowerArm, lUpperLeg, lLowerLeg, lFoot, lToes, rUpperLeg, rLowerLeg, rFoot, rToes };
const int kParent[kParts] = { -1, torso, torso, lUpperArm, torso, rUpperArm, torso, lUpperLeg, lLowerLeg, lFoot, torso, rUpperLeg, rLowerLeg, rFoot };   // ragdoll.cpp:156-169
const float kPi = 3.14159265358979323846f;

vec3 transformPositionT(const trs& m, vec3 p) { return m.rotation * (m.scale * p) + m.position; }                       // math.cpp:518-521
vec3 inverseTransformPositionT(const trs& m, vec3 p) { vec3 q = conjugate(m.rotation) * (p - m.position); return b200::mk3(q.x / m.scale.x, q.y / m.scale.y, q.z / m.scale.z); }
vec3 inverseTransformDirectionT(const trs& m, vec3 d) { return conjugate(m.rotation) * d; }

struct Env
{
    humanoid_ragdoll ragdoll;
    float lastSmoothedAction[kActionFloats];
    float headTargetHeight; vec3 torsoVelocityTarget;
    vec3 localPositions[kParts][6];
    vec3 targetPositions[kParts][6], targetVelocities[kParts][6]; quat localTargetRotation[kParts];
    trs initialTransform[kParts];
    uint32_t firstBody, firstConeTwist, firstHinge, firstColliderCreated, numColliders;
    uint64_t rngState;
};

} // namespace

struct locomotion_env_batch
{
    game_scene scene; std::vector<Env> envs; float physicsTimer = 0.f;
    float pushProbability = 0.02f;      // learned_locomotion.cpp:448
    bool exactOrder = false;            // canonical Gauss-Seidel order (B200PHYS_STEP_EXACT_ORDER) instead of the colour-sorted fast order
    bool deviceMode = false, deviceReady = false;   // everything per environment on the GPU (csrc/k_env.cu): only actions go up, states / rewards / flags come down

    static uint64_t nextRandom(uint64_t& state) { uint64_t x = state; x ^= x << 13; x ^= x >> 7; x ^= x << 17; state = x; return x; }          // core/random.h:14-24
    static float random01(uint64_t& state) { return (uint32_t)nextRandom(state) / (float)0xFFFFFFFFu; }
    // updatePhysics :448-458: push a random body part sideways from 5 m away
    void maybePush(Env& e)
    {
        if (!(random01(e.rngState) < pushProbability)) return;
        const uint32_t part = (uint32_t)nextRandom(e.rngState) % (kParts - 1);                 // randomUint32Between(0, NUM_BODY_PARTS - 1)
        const vec3 target = partTransform(e, (int)part).position + b200::mk3(0.f, 0.2f, 0.f);
        // vec3(rng.randomFloatBetween(-1, 1), 0, rng.randomFloatBetween(-1, 1)) (:455): the order of the two draws is unspecified in C++;
        // g++ (which compiles the reference for the parity tests) and MSVC x64 both evaluate constructor arguments right to left, so z is drawn first
        const float dz = -1.f + 2.f * random01(e.rngState), dx = -1.f + 2.f * random01(e.rngState);
        const vec3 direction = b200::normalize(b200::mk3(dx, 0.f, dz));
        testPhysicsInteraction(scene, target - direction * 5.f, direction, 1000.f, e.firstBody, kParts, e.firstColliderCreated, e.numColliders);
    }

    // The reference reads transform_component everywhere in the environment (learned_locomotion.cpp:113-129, :284-311, :452), not
    // physics_transform1: with the fixed 60 Hz step and exactly one sub-step per call the render transform is interpolated with t = 0,
    // i.e. it shows the pose BEFORE the step just taken (physics.cpp:1396-1402), while velocities are current. Reproduced.
    const trs& partTransform(const Env& e, int part) const { return scene.entities[scene.rigidBodyEntity[e.firstBody + part]].transform; }
    const rigid_body_component& partBody(const Env& e, int part) const { return scene.rigidBodies[e.firstBody + part]; }
    vec3 globalCOG(const Env& e, int part) const { const trs& t = partTransform(e, part); return t.position + t.rotation * partBody(e, part).localCOGPosition; }   // rigid_body.cpp:83-86

    // learned_locomotion::getCoordinateSystem (:113-120): torso COG projected to the ground, identity rotation
    trs coordinateSystem(const Env& e) const { vec3 cog = globalCOG(e, torso); cog.y = 0.f; return makeTrs(cog); }

    // training_locomotion::getLocalPositions (:191-250): the six face centres of the part's local collider bounds
    void localPositions(Env& e, int part)
    {
        vec3 mn = b200::mk3(3.4e38f), mx = b200::mk3(-3.4e38f);
        auto grow = [&](vec3 p) { mn = b200::mk3(fminf(mn.x, p.x), fminf(mn.y, p.y), fminf(mn.z, p.z)); mx = b200::mk3(fmaxf(mx.x, p.x), fmaxf(mx.y, p.y), fmaxf(mx.z, p.z)); };
        const game_scene::entity_data& d = scene.entities[e.ragdoll.bodyParts[part].handle];
        for (entity_handle c = d.firstColliderEntity; c != null_entity; c = scene.colliders[c].nextEntity)
        {
            const collider_component& col = scene.colliders[c];
            if (col.type == collider_type_sphere) { grow(col.sphere.center - b200::mk3(col.sphere.radius)); grow(col.sphere.center + b200::mk3(col.sphere.radius)); }
            else if (col.type == collider_type_capsule)
            {
                const vec3 r3 = b200::mk3(col.capsule.radius);
                grow(col.capsule.positionA + r3); grow(col.capsule.positionA - r3); grow(col.capsule.positionB + r3); grow(col.capsule.positionB - r3);
            }
            else if (col.type == collider_type_aabb) { grow(col.aabb.minCorner); grow(col.aabb.maxCorner); }
        }
        const vec3 c = (mn + mx) * 0.5f, r = (mx - mn) * 0.5f;
        e.localPositions[part][0] = c - b200::mk3(r.x, 0.f, 0.f); e.localPositions[part][1] = c - b200::mk3(0.f, r.y, 0.f); e.localPositions[part][2] = c - b200::mk3(0.f, 0.f, r.z);
        e.localPositions[part][3] = c + b200::mk3(r.x, 0.f, 0.f); e.localPositions[part][4] = c + b200::mk3(0.f, r.y, 0.f); e.localPositions[part][5] = c + b200::mk3(0.f, 0.f, r.z);
    }
    vec3 pointVelocity(const Env& e, int part, vec3 localP) const
    {   // rigid_body_component::getGlobalPointVelocity (rigid_body.cpp:88-93)
        const trs& t = partTransform(e, part); const rigid_body_component& rb = partBody(e, part);
        return rb.linearVelocity + cross(rb.angularVelocity, transformPositionT(t, localP) - globalCOG(e, part));
    }
    quat localRotation(const Env& e, int part) const
    {
        const quat parentRotation = kParent[part] >= 0 ? partTransform(e, kParent[part]).rotation : b200::qIdentity();
        return partTransform(e, part).rotation * conjugate(parentRotation);
    }

    // learned_locomotion::updateConstraint (:71-90) for every joint of the environment
    void applyAction(Env& e, const float* action)
    {   // :92-111 : exponential smoothing with beta = 0.1, then position motors with 200 Nm
        for (uint32_t i = 0; i < kActionFloats; ++i) e.lastSmoothedAction[i] += (action[i] - e.lastSmoothedAction[i]) * 0.1f;     // lerp(out, in, beta)
        for (uint32_t i = 0; i < kConeTwists; ++i)
        {
            cone_twist_constraint& c = scene.coneTwistConstraints[e.firstConeTwist + i]; const float* a = e.lastSmoothedAction + 3 * i;
            c.maxSwingMotorTorque = 200.f; c.maxTwistMotorTorque = 200.f;
            c.swingMotorType = constraint_position_motor; c.twistMotorType = constraint_position_motor;
            c.twistMotorVelocityOrTargetAngle = a[0]; c.swingMotorVelocityOrTargetAngle = a[1]; c.swingMotorAxis = a[2];
        }
        for (uint32_t i = 0; i < kHinges; ++i)
        {
            hinge_constraint& c = scene.hingeConstraints[e.firstHinge + i];
            c.maxMotorTorque = 200.f; c.motorType = constraint_position_motor; c.motorVelocityOrTargetAngle = e.lastSmoothedAction[3 * kConeTwists + i];
        }
    }

    // learned_locomotion::getState (:131-146) + hasFallen (:148-151)
    bool getState(const Env& e, float* out) const
    {
        const trs cs = coordinateSystem(e);
        vec3* v = reinterpret_cast<vec3*>(out);
        auto read = [&](int part, vec3& pos, vec3& vel) { pos = inverseTransformPositionT(cs, globalCOG(e, part)); vel = inverseTransformDirectionT(cs, partBody(e, part).linearVelocity); };
        v[0] = inverseTransformDirectionT(cs, partBody(e, torso).linearVelocity);       // cogVelocity
        read(lToes, v[1], v[2]); read(rToes, v[3], v[4]); read(torso, v[5], v[6]); read(head, v[7], v[8]); read(lLowerArm, v[9], v[10]); read(rLowerArm, v[11], v[12]);
        memcpy(out + 39, e.lastSmoothedAction, sizeof(float) * kActionFloats);
        return v[7].y < 1.f;
    }

    // training_locomotion::getReward (:329-355) with readPartDifference (:284-311)
    float getReward(const Env& e) const
    {
        float positionError = 0.f, velocityError = 0.f, rotationError = 0.f;
        for (uint32_t p = 0; p < kParts; ++p)
        {
            const trs& t = partTransform(e, p);
            for (int i = 0; i < 6; ++i)
            {
                positionError += b200::len(transformPositionT(t, e.localPositions[p][i]) - e.targetPositions[p][i]);
                velocityError += b200::len(pointVelocity(e, p, e.localPositions[p][i]) - e.targetVelocities[p][i]);
            }
            const quat diff = e.localTargetRotation[p] * conjugate(localRotation(e, p));
            rotationError += 2.f * acosf(fminf(fmaxf(diff.w, -1.f), 1.f));
        }
        const float vcmError = b200::len(partBody(e, torso).linearVelocity - e.torsoVelocityTarget);
        const float rp = expf(-10.f / kParts * positionError), rv = expf(-1.f / kParts * velocityError), rlocal = expf(-10.f / kParts * rotationError), rvcm = expf(-vcmError);
        const float headHeight = partTransform(e, head).position.y;
        const float fall = fminf(fmaxf(1.3f - 1.4f * (e.headTargetHeight - headHeight), 0.f), 1.f);
        return fall * (rp + rv + rlocal + rvcm);
    }

    // resetPhysics (:436-460) for one environment: initial pose, zero velocities, training targets from that pose, zero action
    void reset(uint32_t index)
    {
        Env& e = envs[index];
        for (uint32_t p = 0; p < kParts; ++p)
        {
            const uint32_t b = e.firstBody + p;
            scene.physicsTransform0[b] = scene.physicsTransform1[b] = e.initialTransform[p];
            scene.entities[scene.rigidBodyEntity[b]].transform = e.initialTransform[p];
            rigid_body_component& rb = scene.rigidBodies[b];
            rb.linearVelocity = rb.angularVelocity = rb.forceAccumulator = rb.torqueAccumulator = b200::mk3(0.f);
        }
        for (uint32_t p = 0; p < kParts; ++p)
        {   // training_locomotion::reset (:313-324): targets = the pose right now
            localPositions(e, p);
            for (int i = 0; i < 6; ++i) { e.targetPositions[p][i] = transformPositionT(partTransform(e, p), e.localPositions[p][i]); e.targetVelocities[p][i] = pointVelocity(e, p, e.localPositions[p][i]); }
            e.localTargetRotation[p] = localRotation(e, p);
        }
        memset(e.lastSmoothedAction, 0, sizeof(e.lastSmoothedAction));          // learned_locomotion::reset (:36-43)
        const float zero[kActionFloats] = {};
        applyAction(e, zero); scene.jointsDirty = true;
        e.headTargetHeight = partTransform(e, head).position.y;
        e.torsoVelocityTarget = b200::mk3(0.f);
    }

    void create(uint32_t numEnvs, float spacing)
    {
        const uint32_t side = (uint32_t)ceilf(sqrtf((float)numEnvs));
        const float half = (side - 1) * spacing * 0.5f;
        physics_material groundMaterial = { physics_material_type_metal, 0.1f, 1.f, 4.f };
        envs.resize(numEnvs);
        for (uint32_t i = 0; i < numEnvs; ++i)
        {
            Env& e = envs[i];
            e.rngState = 0x9E3779B97F4A7C15ull ^ ((uint64_t)(i + 1) * 0xD1B54A32D192ED03ull);
            e.firstBody = (uint32_t)scene.rigidBodies.size(); e.firstConeTwist = (uint32_t)scene.coneTwistConstraints.size(); e.firstHinge = (uint32_t)scene.hingeConstraints.size();
            e.firstColliderCreated = (uint32_t)scene.colliders.size();
            e.ragdoll = humanoid_ragdoll::create(scene, b200::mk3((i % side) * spacing - half, 1.25f, (i / side) * spacing - half));
            e.numColliders = (uint32_t)scene.colliders.size() - e.firstColliderCreated;
            for (uint32_t p = 0; p < kParts; ++p) e.initialTransform[p] = scene.physicsTransform1[e.firstBody + p];
        }
        const float ext = fmaxf(20.f, half + 10.f);
        scene.createEntity("Test ground").addTransform(b200::mk3(0.f, -4.f, 0.f)).addCollider(collider_component::asAABB(bounding_box::fromCenterRadius(b200::mk3(0.f), b200::mk3(ext, 4.f, ext)), groundMaterial));
        for (uint32_t i = 0; i < numEnvs; ++i) reset(i);
    }

    // updatePhysics (:462-488) for all environments: apply the actions, ONE 60 Hz physics step, read state / reward / failure
    // one-time hand-over to the device: the scene as it stands + the per-environment tables of b200phys_env_create
    int ensureDevice()
    {
        if (deviceReady) return 0;
        int r = uploadSceneToDevice(scene); if (r) return r;
        const uint32_t E = (uint32_t)envs.size(), nC = (uint32_t)scene.colliders.size();
        std::vector<uint32_t> fb(E), fc(E), fh(E), fcol(E); std::vector<float> rot((size_t)E * kParts * 4), pos((size_t)E * kParts * 3), lp((size_t)kParts * 18), tp((size_t)E * kParts * 18), tv((size_t)E * kParts * 18), tr((size_t)E * kParts * 4), hh(E);
        std::vector<unsigned long long> seeds(E);
        for (uint32_t i = 0; i < E; ++i)
        {
            const Env& e = envs[i];
            fb[i] = e.firstBody; fc[i] = e.firstConeTwist; fh[i] = e.firstHinge;
            fcol[i] = nC - e.firstColliderCreated - e.numColliders;            // device collider index = reverse creation order (physics.cpp:635-641)
            hh[i] = e.headTargetHeight; seeds[i] = e.rngState;
            for (uint32_t p = 0; p < kParts; ++p)
            {
                const size_t k = (size_t)i * kParts + p; const trs& t = e.initialTransform[p];
                rot[4 * k] = t.rotation.x; rot[4 * k + 1] = t.rotation.y; rot[4 * k + 2] = t.rotation.z; rot[4 * k + 3] = t.rotation.w;
                pos[3 * k] = t.position.x; pos[3 * k + 1] = t.position.y; pos[3 * k + 2] = t.position.z;
                tr[4 * k] = e.localTargetRotation[p].x; tr[4 * k + 1] = e.localTargetRotation[p].y; tr[4 * k + 2] = e.localTargetRotation[p].z; tr[4 * k + 3] = e.localTargetRotation[p].w;
                for (int j = 0; j < 6; ++j)
                {
                    const size_t q = (k * 6 + j) * 3;
                    tp[q] = e.targetPositions[p][j].x; tp[q + 1] = e.targetPositions[p][j].y; tp[q + 2] = e.targetPositions[p][j].z;
                    tv[q] = e.targetVelocities[p][j].x; tv[q + 1] = e.targetVelocities[p][j].y; tv[q + 2] = e.targetVelocities[p][j].z;
                    if (i == 0) { const size_t l = ((size_t)p * 6 + j) * 3; lp[l] = e.localPositions[p][j].x; lp[l + 1] = e.localPositions[p][j].y; lp[l + 2] = e.localPositions[p][j].z; }
                }
            }
        }
        b200phys_locomotion_env d; memset(&d, 0, sizeof(d));
        d.numEnvs = E; d.partsPerEnv = kParts; d.coneTwistsPerEnv = kConeTwists; d.hingesPerEnv = kHinges; d.collidersPerEnv = envs[0].numColliders;
        for (uint32_t p = 0; p < 16; ++p) d.parentPart[p] = p < kParts ? kParent[p] : -1;
        const uint32_t sp[6] = { lToes, rToes, torso, head, lLowerArm, rLowerArm }; memcpy(d.stateParts, sp, sizeof(sp));
        d.torsoPart = torso; d.headPart = head;
        d.firstBody = fb.data(); d.firstConeTwist = fc.data(); d.firstHinge = fh.data(); d.firstCollider = fcol.data();
        d.initialRotations = rot.data(); d.initialPositions = pos.data(); d.localPositions = lp.data(); d.targetPositions = tp.data(); d.targetVelocities = tv.data();
        d.targetRotations = tr.data(); d.headTargetHeights = hh.data(); d.seeds = seeds.data();
        r = b200phys_env_create(scene.ctx, &d); if (r) return r;
        deviceReady = true;
        return 0;
    }

    int step(const float* actions, float* outStates, float* outRewards, int* outFailed)
    {
        if (deviceMode)
        {
            int r = ensureDevice(); if (r) return r;
            return b200phys_env_step(scene.ctx, actions, 1.f / 60.f, physics_settings().numRigidSolverIterations, exactOrder ? B200PHYS_STEP_EXACT_ORDER : B200PHYS_STEP_FAST,
                                     pushProbability, outStates, outRewards, (int32_t*)outFailed);
        }
        parallelFor((uint32_t)envs.size(), [&](uint32_t i) { applyAction(envs[i], actions + (size_t)i * kActionFloats); }, 256u);
        scene.jointsDirty = true;
        if (pushProbability > 0.f) for (Env& e : envs) maybePush(e);
        physics_settings settings; settings.frameRate = 60; settings.exactConstraintOrder = exactOrder;
        memory_arena arena; physicsTimer = 0.f;
        const int rc = physicsStep(scene, arena, physicsTimer, settings, 1.f / (float)settings.frameRate);
        if (rc) return rc;
        parallelFor((uint32_t)envs.size(), [&](uint32_t i)
        {
            const bool failed = getState(envs[i], outStates + (size_t)i * kStateFloats);
            outFailed[i] = failed ? 1 : 0;
            outRewards[i] = failed ? 0.f : getReward(envs[i]);
        }, 256u);
        return 0;
    }
};

} // namespace b200host

using namespace b200host;

extern "C" {

// ---- batched interface
void* b200env_create(uint32_t numEnvs, float spacing) { locomotion_env_batch* b = new locomotion_env_batch(); b->create(numEnvs ? numEnvs : 1, spacing > 0.f ? spacing : 4.f); return b; }
void b200env_destroy(void* env) { delete (locomotion_env_batch*)env; }
uint32_t b200env_state_size(void) { return kStateFloats; }
uint32_t b200env_action_size(void) { return kActionFloats; }
void b200env_set_push_probability(void* env, float p) { ((locomotion_env_batch*)env)->pushProbability = p; }
void b200env_set_exact_order(void* env, int on) { ((locomotion_env_batch*)env)->exactOrder = on != 0; }
// on: everything per environment runs on the GPU (csrc/k_env.cu); choose before the first step
void b200env_set_device_mode(void* env, int on) { locomotion_env_batch& b = *(locomotion_env_batch*)env; if (!b.deviceReady) b.deviceMode = on != 0; }
// seed of one environment's push generator (xorshift64 state, core/random.h: the reference seeds its single generator with the wall clock)
void b200env_set_seed(void* env, uint32_t index, uint64_t seed)
{
    locomotion_env_batch& b = *(locomotion_env_batch*)env;
    if (index >= b.envs.size()) return;
    b.envs[index].rngState = seed;
    if (b.deviceReady) b200phys_env_set_seed(b.scene.ctx, index, seed);
}
uint32_t b200env_num_envs(void* env) { return (uint32_t)((locomotion_env_batch*)env)->envs.size(); }
// index < 0: every environment; writes the state(s) after the reset
void b200env_reset(void* env, int index, float* outStates)
{
    locomotion_env_batch& b = *(locomotion_env_batch*)env;
    for (uint32_t i = 0; i < b.envs.size(); ++i)
        if (index < 0 || (uint32_t)index == i)
        {   // in device mode the host copy of an environment only ever holds its reset state: reset it too and read the state from it
            b.reset(i); if (outStates) b.getState(b.envs[i], outStates + (size_t)(index < 0 ? i : 0) * kStateFloats);
            if (b.deviceReady) b200phys_env_reset(b.scene.ctx, &i, 1);
        }
}
float b200env_reward(void* env, uint32_t index) { locomotion_env_batch& b = *(locomotion_env_batch*)env; return b.getReward(b.envs[index]); }
int b200env_step(void* env, const float* actions, float* outStates, float* outRewards, int* outFailed) { return ((locomotion_env_batch*)env)->step(actions, outStates, outRewards, outFailed); }
// action limits from the joints' own limits (getLimits, learned_locomotion.cpp:357-383); the state is unbounded
void b200env_ranges(void* env, float* actionMin, float* actionMax)
{
    locomotion_env_batch& b = *(locomotion_env_batch*)env; const Env& e = b.envs[0];
    uint32_t k = 0;
    for (uint32_t i = 0; i < kConeTwists; ++i)
    {
        const cone_twist_constraint& c = b.scene.coneTwistConstraints[e.firstConeTwist + i];
        actionMin[k] = c.twistLimit >= 0.f ? -c.twistLimit : -kPi; actionMax[k++] = c.twistLimit >= 0.f ? c.twistLimit : kPi;
        actionMin[k] = c.swingLimit >= 0.f ? -c.swingLimit : -kPi; actionMax[k++] = c.swingLimit >= 0.f ? c.swingLimit : kPi;
        actionMin[k] = -kPi; actionMax[k++] = kPi;
    }
    for (uint32_t i = 0; i < kHinges; ++i)
    {
        const hinge_constraint& c = b.scene.hingeConstraints[e.firstHinge + i];
        actionMin[k] = c.minRotationLimit <= 0.f ? c.minRotationLimit : -kPi; actionMax[k++] = c.maxRotationLimit >= 0.f ? c.maxRotationLimit : kPi;
    }
}

// ---- the reference's five functions (learned_locomotion.cpp:393-488), one environment
static locomotion_env_batch* gSingle = nullptr;
int getPhysicsStateSize(void) { return (int)kStateFloats; }
int getPhysicsActionSize(void) { return (int)kActionFloats; }
void getPhysicsRanges(float* stateMin, float* stateMax, float* actionMin, float* actionMax)
{
    for (uint32_t i = 0; i < kStateFloats; ++i) { stateMin[i] = -3.402823466e38f; stateMax[i] = 3.402823466e38f; }
    locomotion_env_batch tmp; tmp.create(1, 4.f);
    b200env_ranges(&tmp, actionMin, actionMax);
}
void resetPhysics(float* outState)
{
    if (!gSingle) { gSingle = new locomotion_env_batch(); gSingle->create(1, 4.f); }
    gSingle->reset(0);
    if (outState) gSingle->getState(gSingle->envs[0], outState);
}
int updatePhysics(float* action, float* outState, float* outReward)
{
    if (!gSingle) resetPhysics(nullptr);
    int failed = 0;
    const int rc = gSingle->step(action, outState, outReward, &failed);
    return rc ? 1 : failed;
}

} // extern "C"
