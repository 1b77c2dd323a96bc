// k_events.cu -- collision begin / end events (reference: handleCollisionCallbacks, src/physics/physics.cpp:1037-1178).
//
// The reference sorts this frame's colliding collider pairs and merges them against last frame's list on the host; a pair
// only in the new list raises a begin event (contact point / normal averaged over the manifold, relative velocity of the two
// bodies at that point from the post-force, pre-solve state), a pair only in the old list an end event. Here both lists stay
// on the device: the manifold list of a step is already sorted by (type bucket, min collider, max collider), so a 64-bit key
// per manifold + one binary search per element in the other frame's list is the whole diff; flags -> scan -> compaction gives
// dense, deterministically ordered event arrays that the host downloads (the host mirror turns them into callbacks).
#include "b200_context.h"
#include "b200_util.cuh"

namespace b200 {

static const int kThreads = 256;

B200_D uint32_t evBucketOf(uint32_t ta, uint32_t tb) { if (ta > tb) { const uint32_t t = ta; ta = tb; tb = t; } return ta * 6u - (ta * (ta - 1u)) / 2u + (tb - ta); }

__global__ void __launch_bounds__(kThreads) evKeysKernel(const uint2* __restrict__ collisions, const float4* __restrict__ wCol, unsigned long long* __restrict__ keys, uint2* __restrict__ pairsOut,
                                                         const DeviceCounts* __restrict__ cnt)
{
    const uint32_t n = cnt->numManifolds;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
    {
        const uint2 pr = collisions[i];
        const uint32_t ta = __float_as_uint(wCol[4 * pr.x + 3].x) & 0xFFu, tb = __float_as_uint(wCol[4 * pr.y + 3].x) & 0xFFu;
        const uint32_t lo = min(pr.x, pr.y), hi = max(pr.x, pr.y);
        keys[i] = ((unsigned long long)evBucketOf(ta, tb) << 56) | ((unsigned long long)lo << 28) | (unsigned long long)hi;
        pairsOut[i] = pr;
    }
}

B200_D bool evContains(const unsigned long long* __restrict__ keys, uint32_t n, unsigned long long k)
{
    uint32_t lo = 0, hi = n;
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (keys[mid] < k) lo = mid + 1; else hi = mid; }
    return lo < n && keys[lo] == k;
}

__global__ void __launch_bounds__(kThreads) evDiffKernel(const unsigned long long* __restrict__ keysCur, const unsigned long long* __restrict__ keysPrev,
                                                         uint32_t* __restrict__ flagBegin, uint32_t* __restrict__ flagEnd, const DeviceCounts* __restrict__ cnt)
{
    const uint32_t nCur = cnt->numManifolds, nPrev = cnt->numPrevCollisions;
    const uint32_t n = max(nCur, nPrev);
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
    {
        if (i < nCur) flagBegin[i] = evContains(keysPrev, nPrev, keysCur[i]) ? 0u : 1u;
        if (i < nPrev) flagEnd[i] = evContains(keysCur, nCur, keysPrev[i]) ? 0u : 1u;
    }
}

__global__ void __launch_bounds__(kThreads) evEmitKernel(const uint32_t* __restrict__ flagBegin, const uint32_t* __restrict__ slotBegin, const uint32_t* __restrict__ flagEnd, const uint32_t* __restrict__ slotEnd,
    const uint2* __restrict__ collisions, const uint2* __restrict__ pairsPrev, const uint4* __restrict__ manifolds, const float4* __restrict__ contacts,
    const float4* __restrict__ gVel, const float4* __restrict__ gPos, b200phys_collision_begin_event* __restrict__ beginOut, b200phys_collider_pair* __restrict__ endOut,
    const DeviceCounts* __restrict__ cnt)
{
    const uint32_t nCur = cnt->numManifolds, nPrev = cnt->numPrevCollisions;
    const uint32_t n = max(nCur, nPrev);
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
    {
        if (i < nCur && flagBegin[i])
        {
            const uint4 man = manifolds[i];     // bodyA, bodyB, first contact, contact count
            const float norm = 1.f / (float)man.w;
            f3 point = mk3(0.f), normal = mk3(0.f);
            for (uint32_t k = 0; k < man.w; ++k) { point += xyz(contacts[2 * (man.z + k)]); normal += xyz(contacts[2 * (man.z + k) + 1]); }
            point *= norm; normal *= norm;
            const f3 velA = xyz(gVel[2 * man.x]) + cross(xyz(gVel[2 * man.x + 1]), point - xyz(gPos[man.x]));
            const f3 velB = xyz(gVel[2 * man.y]) + cross(xyz(gVel[2 * man.y + 1]), point - xyz(gPos[man.y]));
            const f3 rel = velB - velA;
            b200phys_collision_begin_event e;
            const uint2 pr = collisions[i];
            e.colliderA = pr.x; e.colliderB = pr.y;
            e.point[0] = point.x; e.point[1] = point.y; e.point[2] = point.z;
            e.normal[0] = normal.x; e.normal[1] = normal.y; e.normal[2] = normal.z;
            e.relativeVelocity[0] = rel.x; e.relativeVelocity[1] = rel.y; e.relativeVelocity[2] = rel.z;
            beginOut[slotBegin[i]] = e;
        }
        if (i < nPrev && flagEnd[i]) { const uint2 pr = pairsPrev[i]; b200phys_collider_pair e; e.a = pr.x; e.b = pr.y; endOut[slotEnd[i]] = e; }
    }
}

__global__ void evFinishKernel(const uint32_t* __restrict__ slotBegin, const uint32_t* __restrict__ slotEnd, DeviceCounts* cnt)
{
    if (threadIdx.x == 0)
    {
        cnt->numBeginEvents = slotBegin[cnt->numManifolds];
        cnt->numEndEvents = slotEnd[cnt->numPrevCollisions];
        cnt->numPrevCollisions = cnt->numManifolds;       // this frame's list is next frame's "previous"
    }
}

// after the narrow phase and the force integration, before the solver (physics.cpp:1284)
void launchCollisionEvents(Context* c)
{
    if (!c->eventsEnabled) return;
    DeviceCounts* cnt = c->counts;
    const int G = gridFor(c, 4);
    unsigned long long* keysCur = c->evKeys[c->evParity]; unsigned long long* keysPrev = c->evKeys[c->evParity ^ 1];
    uint2* pairsCur = c->evPairs[c->evParity]; uint2* pairsPrev = c->evPairs[c->evParity ^ 1];
    launchK(c, evKeysKernel, G, kThreads, 0, c->collisions, c->wCol, keysCur, pairsCur, cnt);
    launchK(c, evDiffKernel, G, kThreads, 0, keysCur, keysPrev, c->evFlag[0], c->evFlag[1], cnt);
    c->launches += 2;
    scanExclusiveU32(c, c->evFlag[0], c->evSlot[0], &cnt->numManifolds, c->maxManifolds, nullptr);
    scanExclusiveU32(c, c->evFlag[1], c->evSlot[1], &cnt->numPrevCollisions, c->maxManifolds, nullptr);
    launchK(c, evEmitKernel, G, kThreads, 0, c->evFlag[0], c->evSlot[0], c->evFlag[1], c->evSlot[1], c->collisions, pairsPrev, c->manifolds, c->contacts, c->gVel, c->gPos,
                                                c->evBegin, c->evEnd, cnt);
    launchK(c, evFinishKernel, 1, 32, 0, c->evSlot[0], c->evSlot[1], cnt);
    c->launches += 2;
    c->evParity ^= 1;
}

} // namespace b200
