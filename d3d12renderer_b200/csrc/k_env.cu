// k_env.cu -- the batched RL environment of the reference (src/learning/learned_locomotion.{h,cpp}: the one in-repo consumer of the
// physics step) with EVERYTHING per environment on the device: action smoothing + motor targets written straight into
// the uploaded joint structs (:71-111), the random pushes (a ray cast at a random body part, :448-458, testPhysicsInteraction
// physics.cpp:555-629), the 66-float state (:131-146), the reward (:284-355) and the failure test (:148-151). Per step the host sends
// E x 27 action floats and receives E x (66 + 1 + 1) values; bodies and joints never leave the GPU (the host-mirror version re-uploads
// ~230 k bodies and ~210 k joint structs and evaluates the formulas on the CPU: 20.5 ms per step for 16 384 environments, 3 ms of it GPU).
// The formulas are the host mirror's (host/locomotion_env.cpp), which is pinned against the reference's own environment
// (tests/test_gpu_parity_scale.py::test_locomotion_env_matches_the_reference_environment).
#include "b200_context.h"
#include "b200_ray.cuh"
#include "../../include/b200phys.h"
#include <cstring>
#include <vector>

namespace b200 {

static const uint32_t kMaxParts = 16, kMaxActions = 32;

struct EnvDevice
{
    uint32_t numEnvs = 0, parts = 0, coneTwists = 0, hinges = 0, colliders = 0, actionFloats = 0, stateFloats = 0;
    int32_t parent[kMaxParts]; uint32_t stateParts[6]; uint32_t torso = 0, head = 0;
    // per environment
    uint32_t* firstBody = nullptr; uint32_t* firstConeTwist = nullptr; uint32_t* firstHinge = nullptr; uint32_t* firstCollider = nullptr;
    float* lastAction = nullptr;            // E x actionFloats, exponentially smoothed
    unsigned long long* rng = nullptr;      // xorshift64 state per environment
    float4* initRot = nullptr; float4* initPos = nullptr;       // E x parts
    float* localPositions = nullptr;        // parts x 6 x 3 (the same ragdoll everywhere)
    float* targetPositions = nullptr;       // E x parts x 6 x 3
    float* targetVelocities = nullptr;      // E x parts x 6 x 3
    float4* targetRotation = nullptr;       // E x parts
    float* headTargetHeight = nullptr;      // E
    float4* prevRot = nullptr; float4* prevPos = nullptr;       // per body: the pose before the last step (what the reference's render transform shows)
    float* dActions = nullptr; float* dStates = nullptr; float* dRewards = nullptr; int32_t* dFailed = nullptr; uint32_t* dResetList = nullptr;
};

struct EnvParams
{
    uint32_t numEnvs, parts, coneTwists, hinges, colliders, actionFloats, stateFloats, torso, head;
    int32_t parent[kMaxParts]; uint32_t stateParts[6];
    const uint32_t* firstBody; const uint32_t* firstConeTwist; const uint32_t* firstHinge; const uint32_t* firstCollider;
    float* lastAction; unsigned long long* rng;
    const float4* initRot; const float4* initPos; const float* localPositions; const float* targetPositions; const float* targetVelocities; const float4* targetRotation; const float* headTargetHeight;
    float4* prevRot; float4* prevPos;
};

B200_D q4 asQ(float4 v) { return mkq(v.x, v.y, v.z, v.w); }
B200_D f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }

// learned_locomotion::applyAction (:92-111) + updateConstraint (:71-90): lerp(last, action, 0.1), position motors with 200 Nm
__global__ void __launch_bounds__(128) envApplyActionsKernel(EnvParams P, const float* __restrict__ actions, b200phys_cone_twist_joint* __restrict__ coneTwists, b200phys_hinge_joint* __restrict__ hinges, int zero)
{
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < P.numEnvs; e += gridDim.x * blockDim.x)
    {
        float* last = P.lastAction + (size_t)e * P.actionFloats;
        for (uint32_t i = 0; i < P.actionFloats; ++i) { const float a = zero ? 0.f : actions[(size_t)e * P.actionFloats + i]; last[i] += (a - last[i]) * 0.1f; }
        for (uint32_t i = 0; i < P.coneTwists; ++i)
        {
            b200phys_cone_twist_joint& c = coneTwists[P.firstConeTwist[e] + i]; const float* a = last + 3 * i;
            c.maxSwingMotorTorque = 200.f; c.maxTwistMotorTorque = 200.f; c.swingMotorType = 1; c.twistMotorType = 1;     // constraint_position_motor
            c.twistMotorVelocityOrTargetAngle = a[0]; c.swingMotorVelocityOrTargetAngle = a[1]; c.swingMotorAxis = a[2];
        }
        for (uint32_t i = 0; i < P.hinges; ++i)
        {
            b200phys_hinge_joint& c = hinges[P.firstHinge[e] + i];
            c.maxMotorTorque = 200.f; c.motorType = 1; c.motorVelocityOrTargetAngle = last[3 * P.coneTwists + i];
        }
    }
}

B200_D unsigned long long envNextRandom(unsigned long long& s) { unsigned long long x = s; x ^= x << 13; x ^= x >> 7; x ^= x << 17; s = x; return x; }   // core/random.h:14-24
B200_D float envRandom01(unsigned long long& s) { return (float)(uint32_t)envNextRandom(s) / (float)0xFFFFFFFFu; }

// updatePhysics :448-458 + testPhysicsInteraction (physics.cpp:555-629) restricted to the environment's own bodies
__global__ void __launch_bounds__(128) envPushKernel(EnvParams P, float pushProbability, const float4* __restrict__ bRot, const float4* __restrict__ bPos, const float4* __restrict__ bCog,
    float4* __restrict__ bForce, float4* __restrict__ bTorque, const float4* __restrict__ cLocal)
{
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < P.numEnvs; e += gridDim.x * blockDim.x)
    {
        unsigned long long s = P.rng[e];
        const bool push = envRandom01(s) < pushProbability;
        if (push)
        {
            const uint32_t part = (uint32_t)envNextRandom(s) % (P.parts - 1u);
            const uint32_t fb = P.firstBody[e];
            const float4 pp = P.prevPos[fb + part];
            const f3 target = mk3(pp.x, pp.y, pp.z) + mk3(0.f, 0.2f, 0.f);
            const float dz = -1.f + 2.f * envRandom01(s), dx = -1.f + 2.f * envRandom01(s);       // argument evaluation order: see host/locomotion_env.cpp
            const f3 direction = normalize(mk3(dx, 0.f, dz));
            const f3 origin = target - direction * 5.f;
            float minT = FLT_MAX; int32_t minRB = -1; f3 force = mk3(0.f), torque = mk3(0.f);
            // the host walks the colliders in creation order; device index = reverse creation order (physics.cpp:635-641)
            for (uint32_t k = 0; k < P.colliders; ++k)
            {
                const uint32_t ci = P.firstCollider[e] + (P.colliders - 1u - k);
                const float4 c0 = cLocal[4 * ci], c1 = cLocal[4 * ci + 1], c2 = cLocal[4 * ci + 2], c3 = cLocal[4 * ci + 3];
                // uploaded b200phys_collider record: shape[10], restitution, friction | density, (type, objectType, reserved), objectIndex, parentTransform
                const uint32_t tw = __float_as_uint(c3.y), type = tw & 0xFFu, objType = (tw >> 8) & 0xFFu;
                if (objType != B200PHYS_OBJ_RIGID_BODY) continue;
                const uint32_t rb = __float_as_uint(c3.z);
                if (rb < fb || rb - fb >= P.parts) continue;
                const q4 rot = asQ(bRot[rb]); const float4 p4 = bPos[rb]; const f3 pos = mk3(p4.x, p4.y, p4.z);
                const lray local = { conjugate(rot) * (origin - pos), conjugate(rot) * direction };
                float t = 0.f; bool hit = false;
                if (type == B200PHYS_SPHERE) hit = intersectSphere(local, mk3(c0.x, c0.y, c0.z), c0.w, t);
                else if (type == B200PHYS_CAPSULE) hit = intersectCapsule(local, mk3(c0.x, c0.y, c0.z), mk3(c0.w, c1.x, c1.y), c1.z, t);
                else if (type == B200PHYS_CYLINDER) hit = intersectCylinder(local, mk3(c0.x, c0.y, c0.z), mk3(c0.w, c1.x, c1.y), c1.z, t);
                else if (type == B200PHYS_AABB) hit = intersectAABB(local, mk3(c0.x, c0.y, c0.z), mk3(c0.w, c1.x, c1.y), t);
                else if (type == B200PHYS_OBB)
                {
                    const q4 oq = mkq(c0.x, c0.y, c0.z, c0.w); const f3 oc = mk3(c1.x, c1.y, c1.z), orad = mk3(c1.w, c2.x, c2.y);
                    const lray lr = { conjugate(oq) * (local.o - oc), conjugate(oq) * local.d };
                    hit = intersectAABB(lr, mk3(0.f) - orad, orad, t);
                }
                // hulls: the environment's ragdolls have none
                if (hit && t < minT)
                {
                    minT = t; minRB = (int32_t)rb;
                    const f3 globalHit = rot * (local.o + t * local.d) + pos;            // scale 1
                    const float4 cg = bCog[rb];
                    const f3 cog = pos + rot * mk3(cg.x, cg.y, cg.z);
                    force = direction * 1000.f;
                    torque = cross(globalHit - cog, force);
                }
            }
            if (minRB >= 0)
            {
                float4 T = bTorque[minRB], F = bForce[minRB];
                T.x += torque.x; T.y += torque.y; T.z += torque.z; F.x += force.x; F.y += force.y; F.z += force.z;
                bTorque[minRB] = T; bForce[minRB] = F;
            }
        }
        P.rng[e] = s;
    }
}

__global__ void __launch_bounds__(256) envSavePoseKernel(uint32_t n, const float4* __restrict__ bRot, const float4* __restrict__ bPos, float4* __restrict__ prevRot, float4* __restrict__ prevPos)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) { prevRot[i] = bRot[i]; prevPos[i] = bPos[i]; }
}

// learned_locomotion::getState (:131-146), hasFallen (:148-151), training_locomotion::getReward (:329-355) with readPartDifference (:284-311).
// Positions come from the render transform = the pose BEFORE the step just taken (interpolation factor 0, physics.cpp:1396-1402), velocities
// are current -- like the reference's environment and the host mirror's.
__global__ void __launch_bounds__(64) envStateRewardKernel(EnvParams P, const float4* __restrict__ bCog, const float4* __restrict__ bVel, const float4* __restrict__ bAng,
    float* __restrict__ states, float* __restrict__ rewards, int32_t* __restrict__ failed)
{
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < P.numEnvs; e += gridDim.x * blockDim.x)
    {
        const uint32_t fb = P.firstBody[e];
        auto rotOf = [&](uint32_t part) { return asQ(P.prevRot[fb + part]); };
        auto posOf = [&](uint32_t part) { const float4 p = P.prevPos[fb + part]; return mk3(p.x, p.y, p.z); };
        auto cogOf = [&](uint32_t part) { const float4 c = bCog[fb + part]; return posOf(part) + rotOf(part) * mk3(c.x, c.y, c.z); };        // rigid_body.cpp:83-86
        auto velOf = [&](uint32_t part) { const float4 v = bVel[fb + part]; return mk3(v.x, v.y, v.z); };
        // getCoordinateSystem (:113-120): torso COG projected to the ground, identity rotation, unit scale
        f3 cs = cogOf(P.torso); cs.y = 0.f;
        const q4 ident = qIdentity();
        auto invPos = [&](f3 p) { const f3 q = conjugate(ident) * (p - cs); return mk3(q.x / 1.f, q.y / 1.f, q.z / 1.f); };
        auto invDir = [&](f3 d) { return conjugate(ident) * d; };
        float* out = states + (size_t)e * P.stateFloats;
        auto st3 = [&](uint32_t slot, f3 v) { out[3 * slot] = v.x; out[3 * slot + 1] = v.y; out[3 * slot + 2] = v.z; };
        st3(0, invDir(velOf(P.torso)));
        f3 headLocal = mk3(0.f);
        for (uint32_t k = 0; k < 6; ++k)
        {
            const uint32_t part = P.stateParts[k];
            const f3 p = invPos(cogOf(part)), v = invDir(velOf(part));
            st3(1 + 2 * k, p); st3(2 + 2 * k, v);
            if (k == 3) headLocal = p;
        }
        const float* last = P.lastAction + (size_t)e * P.actionFloats;
        for (uint32_t i = 0; i < P.actionFloats; ++i) out[39 + i] = last[i];
        const bool fell = headLocal.y < 1.f;
        failed[e] = fell ? 1 : 0;
        float reward = 0.f;
        if (!fell)
        {
            float positionError = 0.f, velocityError = 0.f, rotationError = 0.f;
            for (uint32_t p = 0; p < P.parts; ++p)
            {
                const q4 rot = rotOf(p); const f3 pos = posOf(p);
                const f3 v = velOf(p); const float4 w4 = bAng[fb + p]; const f3 w = mk3(w4.x, w4.y, w4.z);
                const f3 cog = cogOf(p);
                for (uint32_t i = 0; i < 6; ++i)
                {
                    const f3 lp = ld3(P.localPositions + (size_t)(p * 6 + i) * 3);
                    const f3 wp = rot * (mk3(1.f) * lp) + pos;                                                  // transformPosition (math.cpp:518-521)
                    const size_t ti = (((size_t)e * P.parts + p) * 6 + i) * 3;
                    positionError += len(wp - ld3(P.targetPositions + ti));
                    const f3 pv = v + cross(w, wp - cog);                                                       // getGlobalPointVelocity (rigid_body.cpp:88-93)
                    velocityError += len(pv - ld3(P.targetVelocities + ti));
                }
                const q4 parentRot = P.parent[p] >= 0 ? rotOf((uint32_t)P.parent[p]) : qIdentity();
                const q4 localRot = rot * conjugate(parentRot);
                const q4 diff = asQ(P.targetRotation[(size_t)e * P.parts + p]) * conjugate(localRot);
                rotationError += 2.f * acosf(fminf(fmaxf(diff.w, -1.f), 1.f));
            }
            const f3 torsoVelocityTarget = mk3(0.f);
            const float vcmError = len(velOf(P.torso) - torsoVelocityTarget);
            const float np = (float)P.parts;
            const float rp = expf(-10.f / np * positionError), rv = expf(-1.f / np * velocityError), rlocal = expf(-10.f / np * rotationError), rvcm = expf(-vcmError);
            const float headHeight = posOf(P.head).y;
            const float fall = fminf(fmaxf(1.3f - 1.4f * (P.headTargetHeight[e] - headHeight), 0.f), 1.f);
            reward = fall * (rp + rv + rlocal + rvcm);
        }
        rewards[e] = reward;
    }
}

// resetPhysics (:436-460) for a list of environments: initial pose (current and previous), zero velocities / accumulators / smoothed action
__global__ void __launch_bounds__(64) envResetKernel(EnvParams P, const uint32_t* __restrict__ list, uint32_t count, float4* __restrict__ bRot, float4* __restrict__ bPos,
    float4* __restrict__ bVel, float4* __restrict__ bAng, float4* __restrict__ bForce, float4* __restrict__ bTorque)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    {
        const uint32_t e = list[i]; if (e >= P.numEnvs) continue;
        const uint32_t fb = P.firstBody[e];
        for (uint32_t p = 0; p < P.parts; ++p)
        {
            const float4 r = P.initRot[(size_t)e * P.parts + p]; float4 x = P.initPos[(size_t)e * P.parts + p];
            x.w = bPos[fb + p].w;                                                   // invMass rides in .w
            bRot[fb + p] = r; bPos[fb + p] = x; P.prevRot[fb + p] = r; P.prevPos[fb + p] = x;
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            bVel[fb + p] = z; bAng[fb + p] = z; bForce[fb + p] = z; bTorque[fb + p] = z;
        }
        float* last = P.lastAction + (size_t)e * P.actionFloats;
        for (uint32_t k = 0; k < P.actionFloats; ++k) last[k] = 0.f;
    }
}

static EnvParams paramsOf(const EnvDevice& E)
{
    EnvParams P; memset(&P, 0, sizeof(P));
    P.numEnvs = E.numEnvs; P.parts = E.parts; P.coneTwists = E.coneTwists; P.hinges = E.hinges; P.colliders = E.colliders; P.actionFloats = E.actionFloats; P.stateFloats = E.stateFloats;
    P.torso = E.torso; P.head = E.head; memcpy(P.parent, E.parent, sizeof(P.parent)); memcpy(P.stateParts, E.stateParts, sizeof(P.stateParts));
    P.firstBody = E.firstBody; P.firstConeTwist = E.firstConeTwist; P.firstHinge = E.firstHinge; P.firstCollider = E.firstCollider;
    P.lastAction = E.lastAction; P.rng = E.rng; P.initRot = E.initRot; P.initPos = E.initPos; P.localPositions = E.localPositions; P.targetPositions = E.targetPositions;
    P.targetVelocities = E.targetVelocities; P.targetRotation = E.targetRotation; P.headTargetHeight = E.headTargetHeight; P.prevRot = E.prevRot; P.prevPos = E.prevPos;
    return P;
}

void destroyEnv(Context* c)
{
    EnvDevice* E = (EnvDevice*)c->env; if (!E) return;
    void* ptrs[] = { E->firstBody, E->firstConeTwist, E->firstHinge, E->firstCollider, E->lastAction, E->rng, E->initRot, E->initPos, E->localPositions, E->targetPositions, E->targetVelocities,
                     E->targetRotation, E->headTargetHeight, E->prevRot, E->prevPos, E->dActions, E->dStates, E->dRewards, E->dFailed, E->dResetList };
    for (void* p : ptrs) if (p) cudaFree(p);
    delete E; c->env = nullptr;
}

} // namespace b200

using namespace b200;

struct b200phys_context { Context c; };

#define CKE(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return -(int)e_; } while (0)
template <class T> static cudaError_t upload(T*& dst, const T* src, size_t n) { cudaError_t e = cudaMalloc((void**)&dst, sizeof(T) * (n ? n : 1)); if (e != cudaSuccess) return e; return n ? cudaMemcpy(dst, src, sizeof(T) * n, cudaMemcpyHostToDevice) : cudaSuccess; }

extern "C" {

int b200phys_env_create(b200phys_context* ctx, const b200phys_locomotion_env* d)
{
    if (!ctx || !d || !d->numEnvs || d->partsPerEnv > kMaxParts || d->partsPerEnv < 2 || 3 * d->coneTwistsPerEnv + d->hingesPerEnv > kMaxActions) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CKE(cudaSetDevice(c.device));
    CKE(cudaStreamSynchronize(c.stream));
    destroyEnv(&c);
    EnvDevice* E = new EnvDevice(); c.env = E;
    E->numEnvs = d->numEnvs; E->parts = d->partsPerEnv; E->coneTwists = d->coneTwistsPerEnv; E->hinges = d->hingesPerEnv; E->colliders = d->collidersPerEnv;
    E->actionFloats = 3 * d->coneTwistsPerEnv + d->hingesPerEnv; E->stateFloats = 39 + E->actionFloats;
    for (uint32_t p = 0; p < kMaxParts; ++p) E->parent[p] = p < d->partsPerEnv ? d->parentPart[p] : -1;
    for (int k = 0; k < 6; ++k) E->stateParts[k] = d->stateParts[k];
    E->torso = d->torsoPart; E->head = d->headPart;
    const size_t n = d->numEnvs, np = n * d->partsPerEnv;
    CKE(upload(E->firstBody, d->firstBody, n)); CKE(upload(E->firstConeTwist, d->firstConeTwist, n)); CKE(upload(E->firstHinge, d->firstHinge, n)); CKE(upload(E->firstCollider, d->firstCollider, n));
    std::vector<float4> rot(np), pos(np), trot(np);
    for (size_t i = 0; i < np; ++i)
    {
        rot[i] = make_float4(d->initialRotations[4 * i], d->initialRotations[4 * i + 1], d->initialRotations[4 * i + 2], d->initialRotations[4 * i + 3]);
        pos[i] = make_float4(d->initialPositions[3 * i], d->initialPositions[3 * i + 1], d->initialPositions[3 * i + 2], 0.f);
        trot[i] = make_float4(d->targetRotations[4 * i], d->targetRotations[4 * i + 1], d->targetRotations[4 * i + 2], d->targetRotations[4 * i + 3]);
    }
    CKE(upload(E->initRot, rot.data(), np)); CKE(upload(E->initPos, pos.data(), np)); CKE(upload(E->targetRotation, trot.data(), np));
    CKE(upload(E->localPositions, d->localPositions, (size_t)d->partsPerEnv * 18)); CKE(upload(E->targetPositions, d->targetPositions, np * 18)); CKE(upload(E->targetVelocities, d->targetVelocities, np * 18));
    CKE(upload(E->headTargetHeight, d->headTargetHeights, n));
    std::vector<unsigned long long> seeds(n); for (size_t i = 0; i < n; ++i) seeds[i] = d->seeds ? d->seeds[i] : (0x9E3779B97F4A7C15ull ^ ((unsigned long long)(i + 1) * 0xD1B54A32D192ED03ull));
    CKE(upload(E->rng, seeds.data(), n));
    CKE(cudaMalloc((void**)&E->lastAction, sizeof(float) * n * E->actionFloats)); CKE(cudaMemset(E->lastAction, 0, sizeof(float) * n * E->actionFloats));
    const size_t nb = (size_t)c.cfg.maxBodies + 2;
    CKE(cudaMalloc((void**)&E->prevRot, sizeof(float4) * nb)); CKE(cudaMalloc((void**)&E->prevPos, sizeof(float4) * nb));
    CKE(cudaMemcpy(E->prevRot, c.bRot, sizeof(float4) * c.numBodies, cudaMemcpyDeviceToDevice)); CKE(cudaMemcpy(E->prevPos, c.bPos, sizeof(float4) * c.numBodies, cudaMemcpyDeviceToDevice));
    CKE(cudaMalloc((void**)&E->dActions, sizeof(float) * n * E->actionFloats)); CKE(cudaMalloc((void**)&E->dStates, sizeof(float) * n * E->stateFloats));
    CKE(cudaMalloc((void**)&E->dRewards, sizeof(float) * n)); CKE(cudaMalloc((void**)&E->dFailed, sizeof(int32_t) * n)); CKE(cudaMalloc((void**)&E->dResetList, sizeof(uint32_t) * n));
    // zero action -> motor parameters of every joint (learned_locomotion::reset, :36-43)
    const EnvParams P = paramsOf(*E);
    envApplyActionsKernel<<<gridFor(&c, 2), 128, 0, c.stream>>>(P, nullptr, (b200phys_cone_twist_joint*)c.jointIn[B200PHYS_JOINT_CONE_TWIST], (b200phys_hinge_joint*)c.jointIn[B200PHYS_JOINT_HINGE], 1);
    CKE(cudaStreamSynchronize(c.stream));
    return 0;
}

int b200phys_env_set_seed(b200phys_context* ctx, uint32_t env, unsigned long long seed)
{
    if (!ctx || !ctx->c.env) return B200PHYS_E_INVALID;
    Context& c = ctx->c; EnvDevice& E = *(EnvDevice*)c.env; CKE(cudaSetDevice(c.device));
    if (env >= E.numEnvs) return B200PHYS_E_INVALID;
    CKE(cudaStreamSynchronize(c.stream));
    CKE(cudaMemcpy(E.rng + env, &seed, sizeof(seed), cudaMemcpyHostToDevice));
    return 0;
}

int b200phys_env_reset(b200phys_context* ctx, const uint32_t* envIndices, uint32_t count)
{
    if (!ctx || !ctx->c.env || (count && !envIndices)) return B200PHYS_E_INVALID;
    Context& c = ctx->c; EnvDevice& E = *(EnvDevice*)c.env; CKE(cudaSetDevice(c.device));
    if (count > E.numEnvs) return B200PHYS_E_INVALID;
    if (!count) return 0;
    CKE(cudaMemcpyAsync(E.dResetList, envIndices, sizeof(uint32_t) * count, cudaMemcpyHostToDevice, c.stream));
    const EnvParams P = paramsOf(E);
    envResetKernel<<<gridFor(&c, 1), 64, 0, c.stream>>>(P, E.dResetList, count, c.bRot, c.bPos, c.bVel, c.bAng, c.bForce, c.bTorque);
    // the joints of ALL environments get their motor targets rewritten from the smoothed actions at the next step; the reset ones read zero then
    CKE(cudaStreamSynchronize(c.stream));       // envIndices may be pageable / reused by the caller
    return 0;
}

int b200phys_env_step(b200phys_context* ctx, const float* actions, float dt, uint32_t numSolverIterations, uint32_t flags, float pushProbability,
                      float* outStates, float* outRewards, int32_t* outFailed)
{
    if (!ctx || !ctx->c.env || !actions || !outStates || !outRewards || !outFailed) return B200PHYS_E_INVALID;
    Context& c = ctx->c; EnvDevice& E = *(EnvDevice*)c.env; CKE(cudaSetDevice(c.device));
    const size_t n = E.numEnvs;
    CKE(cudaMemcpyAsync(E.dActions, actions, sizeof(float) * n * E.actionFloats, cudaMemcpyHostToDevice, c.stream));
    const EnvParams P = paramsOf(E);
    const int G = gridFor(&c, 2);
    envApplyActionsKernel<<<G, 128, 0, c.stream>>>(P, E.dActions, (b200phys_cone_twist_joint*)c.jointIn[B200PHYS_JOINT_CONE_TWIST], (b200phys_hinge_joint*)c.jointIn[B200PHYS_JOINT_HINGE], 0);
    if (pushProbability > 0.f) envPushKernel<<<G, 128, 0, c.stream>>>(P, pushProbability, c.bRot, c.bPos, c.bCog, c.bForce, c.bTorque, c.cLocal);
    envSavePoseKernel<<<gridFor(&c, 4), 256, 0, c.stream>>>(c.numBodies, c.bRot, c.bPos, E.prevRot, E.prevPos);
    const int r = b200phys_step(ctx, dt, numSolverIterations, flags);
    if (r) return r;
    envStateRewardKernel<<<gridFor(&c, 2), 64, 0, c.stream>>>(P, c.bCog, c.bVel, c.bAng, E.dStates, E.dRewards, E.dFailed);
    CKE(cudaMemcpyAsync(outStates, E.dStates, sizeof(float) * n * E.stateFloats, cudaMemcpyDeviceToHost, c.stream));
    CKE(cudaMemcpyAsync(outRewards, E.dRewards, sizeof(float) * n, cudaMemcpyDeviceToHost, c.stream));
    CKE(cudaMemcpyAsync(outFailed, E.dFailed, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, c.stream));
    CKE(cudaStreamSynchronize(c.stream));
    CKE(cudaGetLastError());
    return 0;
}

} // extern "C"
