// k_solver.cu -- sequential-impulse (PGS) constraint solver on the device.
//
// Reference: constraint_solver (src/physics/constraints.cpp:3711-3772), collision constraints :3307-3449.
// The reference visits constraints Gauss-Seidel in array order on one thread; its AVX2 path re-orders them into
// conflict-free 8-wide batches (constraints.cpp:51-184). Here the unit of scheduling is a MANIFOLD (all contacts of one
// collider pair, solved in order by one thread) or a joint; units are scheduled in "colours":
//   * exact mode : colour = wavefront level = 1 + max(level of earlier units sharing a dynamic body). Executing levels in
//                  order reproduces sequential Gauss-Seidel in canonical array order bit for bit.
//   * fast mode  : Jones-Plassmann colouring with a fixed hash priority (deterministic), <= 64 colours.
// Inside one colour no two units touch the same dynamic body, so a colour is one embarrassingly parallel sweep; colours are
// separated by a hand-written grid barrier inside ONE persistent cooperative kernel (grid = SM count), i.e. the
// iterations x colours loop costs no kernel launches. The static dummy body and kinematic bodies never conflict
// (constraints.cpp:82-83).
//
// HBM traffic per contact and iteration: one 128-byte row (8 coalesced 128-bit planes) + 8 B impulses read/written +
// two 32-byte body records gathered and scattered (the only per-body data the iterations touch).
#include "b200_context.h"
#include "b200_util.cuh"
#include "k_joints.cuh"

namespace b200 {

static_assert(sizeof(ConeTwistRowD) <= 512 && sizeof(SliderRowD) <= 512 && sizeof(HingeRowD) <= 512 && sizeof(FixedRowD) <= 512, "joint rows are allocated 512 bytes each");

static const int kThreads = 256;
static const uint32_t kUncolored = 0xFFFFFFFFu;

B200_D bool bodyIsDynamic(const float4* __restrict__ gVel, uint32_t b, uint32_t N) { return b < N && gVel[2 * b].w > 0.f; }
B200_D uint32_t hashPriority(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// ------------------------------------------------------------------------------------------------------ adjacency
__global__ void __launch_bounds__(kThreads) adjCountKernel(const uint2* __restrict__ unitBodies, const uint32_t* __restrict__ numUnits, const float4* __restrict__ gVel, uint32_t N, uint32_t* __restrict__ adjCount)
{
    const uint32_t n = *numUnits;
    for (uint32_t u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads)
    {
        const uint2 b = unitBodies[u];
        if (bodyIsDynamic(gVel, b.x, N)) atomicAdd(&adjCount[b.x], 1u);
        if (b.y != b.x && bodyIsDynamic(gVel, b.y, N)) atomicAdd(&adjCount[b.y], 1u);
    }
}
__global__ void __launch_bounds__(kThreads) adjFillKernel(const uint2* __restrict__ unitBodies, const uint32_t* __restrict__ numUnits, const float4* __restrict__ gVel, uint32_t N,
                                                          const uint32_t* __restrict__ adjStart, uint32_t* __restrict__ adjFill, uint32_t* __restrict__ adjList)
{
    const uint32_t n = *numUnits;
    for (uint32_t u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads)
    {
        const uint2 b = unitBodies[u];
        if (bodyIsDynamic(gVel, b.x, N)) adjList[adjStart[b.x] + atomicAdd(&adjFill[b.x], 1u)] = u;
        if (b.y != b.x && bodyIsDynamic(gVel, b.y, N)) adjList[adjStart[b.y] + atomicAdd(&adjFill[b.y], 1u)] = u;
    }
}

// ------------------------------------------------------------------------------------------- colouring / levelling
// One cooperative kernel: rounds of "colour every unit whose higher-priority neighbours are coloured", then a counting
// sort of the units by colour (colour histogram -> scan by CTA 0 -> scatter). roundDone[r] is written only in round r,
// so every CTA reads the same finished value after the barrier and takes the same exit decision.
__global__ void __launch_bounds__(kThreads) colorKernel(int exact, uint32_t group, const uint2* __restrict__ unitBodies, const uint32_t* __restrict__ numUnitsPtr,
    const float4* __restrict__ gVel, uint32_t N, const uint32_t* __restrict__ adjStart, const uint32_t* __restrict__ adjList,
    uint32_t* __restrict__ color, uint32_t* __restrict__ order, uint32_t* __restrict__ roundDone, uint32_t* __restrict__ colorCount /* kMaxLevels+1 */,
    uint32_t* __restrict__ colorStartOut /* kMaxLevels+1 */, uint32_t* __restrict__ colorFill, DeviceCounts* cnt, uint32_t barrierSlot)
{
    __shared__ uint32_t sRed[kThreads / 32];
    uint32_t epoch = 0;
    uint32_t* bar = &cnt->barrier[barrierSlot];
    const uint32_t n = *numUnitsPtr;
    const uint32_t maxColors = exact ? kMaxLevels : kMaxColors;
    const uint32_t tid = blockIdx.x * kThreads + threadIdx.x, stride = gridDim.x * kThreads;

    for (uint32_t u = tid; u < n; u += stride) color[u] = kUncolored;
    gridBarrier(bar, epoch);

    uint32_t total = 0, localMax = 0;
    for (uint32_t round = 0; round < kMaxLevels && total < n; ++round)
    {
        uint32_t colouredHere = 0;
        for (uint32_t u = tid; u < n; u += stride)
        {
            if (ldCgU32(color + u) != kUncolored) continue;
            const uint2 b = unitBodies[u];
            const uint32_t pu = hashPriority(u);
            unsigned long long mask = 0ull; uint32_t level = 0; bool wait = false;
            for (int side = 0; side < 2 && !wait; ++side)
            {
                const uint32_t body = side ? b.y : b.x;
                if (side && b.y == b.x) break;
                if (!bodyIsDynamic(gVel, body, N)) continue;
                const uint32_t a0 = adjStart[body], a1 = adjStart[body + 1];
                for (uint32_t a = a0; a < a1; ++a)
                {
                    const uint32_t v = adjList[a];
                    if (v == u) continue;
                    if (exact)
                    {
                        if (v > u) continue;                       // only earlier units constrain the level
                        const uint32_t cv = ldCgU32(color + v);
                        if (cv == kUncolored) { wait = true; break; }
                        level = max(level, cv + 1u);
                    }
                    else
                    {
                        const uint32_t cv = ldCgU32(color + v);
                        if (cv == kUncolored)
                        {
                            const uint32_t pv = hashPriority(v);
                            if (pv > pu || (pv == pu && v > u)) { wait = true; break; }
                        }
                        else mask |= (1ull << cv);
                    }
                }
            }
            if (wait) continue;
            uint32_t c = exact ? level : (uint32_t)(__ffsll((long long)~mask) - 1);
            if (c >= maxColors) { c = maxColors - 1; cnt->overflowColors = 1; }
            asm volatile("st.global.cg.u32 [%0], %1;" :: "l"(color + u), "r"(c) : "memory");
            atomicAdd(&colorCount[c], 1u);
            localMax = max(localMax, c + 1u);
            ++colouredHere;
        }
        // block-reduce the progress counter, one atomic per CTA into this round's private slot
        uint32_t v = colouredHere;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (laneId() == 0) sRed[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < kThreads / 32; ++w) t += sRed[w]; if (t) atomicAdd(&roundDone[round], t); }
        gridBarrier(bar, epoch);
        const uint32_t done = ldCgU32(roundDone + round);
        if (done == 0 && total < n) { if (tid == 0) cnt->overflowColors = 2; break; }   // no progress: refuse to spin
        total += done;
    }
    if (localMax) atomicMax(&cnt->numColors[group], localMax);
    gridBarrier(bar, epoch);

    // counting sort by colour: exclusive scan of the histogram (one CTA; <= 65536 entries), then scatter
    const uint32_t numColors = min(ldCgU32(&cnt->numColors[group]), maxColors);
    if (blockIdx.x == 0)
    {
        __shared__ uint32_t carry;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (uint32_t base = 0; base < numColors; base += kThreads)
        {
            const uint32_t i = base + threadIdx.x;
            const uint32_t x = (i < numColors) ? ldCgU32(colorCount + i) : 0u;
            uint32_t incl = x;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (laneId() >= (uint32_t)o) incl += t; }
            if (laneId() == 31) sRed[threadIdx.x >> 5] = incl;
            __syncthreads();
            uint32_t woff = 0; for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) woff += sRed[w];
            const uint32_t c0 = carry;
            if (i < numColors) { colorStartOut[i] = c0 + woff + incl - x; if (!exact && i <= kMaxColors) cnt->colorStart[group][i] = c0 + woff + incl - x; }
            __syncthreads();
            if (threadIdx.x == kThreads - 1) carry = c0 + woff + incl;
            __syncthreads();
        }
        if (threadIdx.x == 0) { colorStartOut[numColors] = carry; if (!exact && numColors <= kMaxColors) cnt->colorStart[group][numColors] = carry; }
    }
    gridBarrier(bar, epoch);
    for (uint32_t u = tid; u < n; u += stride)
    {
        const uint32_t c = min(ldCgU32(color + u), maxColors - 1);
        const uint32_t pos = ldCgU32(colorStartOut + c) + atomicAdd(&colorFill[c], 1u);
        order[pos] = u;
    }
}

// ------------------------------------------------------------------------------------------------ contact rows
// row planes (float4 each, plane stride = maxContacts), indexed by solve slot:
//  0: rA.xyz, massN   1: rB.xyz, massT   2: normal.xyz, bias   3: tangent.xyz, friction
//  4: nAngA.xyz       5: nAngB.xyz       6: tAngA.xyz          7: tAngB.xyz
__global__ void __launch_bounds__(kThreads) slotCountKernel(const uint32_t* __restrict__ order, const uint4* __restrict__ manifolds, uint32_t* __restrict__ slotCount, const DeviceCounts* __restrict__ cnt)
{
    const uint32_t n = cnt->numManifolds;
    for (uint32_t p = blockIdx.x * kThreads + threadIdx.x; p < n; p += gridDim.x * kThreads) slotCount[p] = manifolds[order[p]].w;
}

B200_D m3 loadInvInertia(const float4* __restrict__ gInvI, uint32_t b)
{
    const float4 r0 = gInvI[3 * b], r1 = gInvI[3 * b + 1], r2 = gInvI[3 * b + 2];
    m3 I; // rows -> column-major storage
    I.m[0] = r0.x; I.m[3] = r0.y; I.m[6] = r0.z;
    I.m[1] = r1.x; I.m[4] = r1.y; I.m[7] = r1.z;
    I.m[2] = r2.x; I.m[5] = r2.y; I.m[8] = r2.z;
    return I;
}

// initializeCollisionVelocityConstraints (constraints.cpp:3307-3379)
__global__ void __launch_bounds__(kThreads) initContactRowsKernel(const uint32_t* __restrict__ order, const uint32_t* __restrict__ unitSlotStart, const uint4* __restrict__ manifolds,
    const float4* __restrict__ contacts, const float4* __restrict__ gVel, const float4* __restrict__ gPos, const float4* __restrict__ gInvI,
    float4* __restrict__ rows, float2* __restrict__ rowImp, uint32_t* __restrict__ slotOfContact, uint32_t maxContacts, float dt, const DeviceCounts* __restrict__ cnt)
{
    const uint32_t n = cnt->numManifolds;
    const float invDt = 1.f / dt;
    for (uint32_t p = blockIdx.x * kThreads + threadIdx.x; p < n; p += gridDim.x * kThreads)
    {
        const uint4 man = manifolds[order[p]];
        const uint32_t slot0 = unitSlotStart[p];
        const float4 vA4 = gVel[2 * man.x], wA4 = gVel[2 * man.x + 1], vB4 = gVel[2 * man.y], wB4 = gVel[2 * man.y + 1];
        const f3 posA = xyz(gPos[man.x]), posB = xyz(gPos[man.y]);
        const m3 IA = loadInvInertia(gInvI, man.x), IB = loadInvInertia(gInvI, man.y);
        const float imA = vA4.w, imB = vB4.w;
        for (uint32_t k = 0; k < man.w; ++k)
        {
            const uint32_t ci = man.z + k, slot = slot0 + k;
            const float4 c0 = contacts[2 * ci], c1 = contacts[2 * ci + 1];
            const f3 point = xyz(c0), normal = xyz(c1); const float depth = c0.w; const uint32_t fr = __float_as_uint(c1.w);
            const f3 rA = point - posA, rB = point - posB;
            const f3 avA = xyz(vA4) + cross(xyz(wA4), rA), avB = xyz(vB4) + cross(xyz(wB4), rB);
            const f3 rel = avB - avA;
            const f3 tangent = noz(rel - dot(normal, rel) * normal);
            const f3 crAt = cross(rA, tangent), crBt = cross(rB, tangent);
            const float invT = imA + dot(crAt, IA * crAt) + imB + dot(crBt, IB * crBt);
            const float massT = (invT != 0.f) ? (1.f / invT) : 0.f;
            const f3 tAngA = IA * crAt, tAngB = IB * crBt;
            const f3 crAn = cross(rA, normal), crBn = cross(rB, normal);
            const float invN = imA + dot(crAn, IA * crAn) + imB + dot(crBn, IB * crBn);
            const float massN = (invN != 0.f) ? (1.f / invN) : 0.f;
            float bias = 0.f;
            if (dt > 1e-5f)
            {
                const float vRel = dot(normal, rel);
                const float slop = -0.001f;
                if (-depth < slop && vRel < 0.f)
                {
                    const float restitution = (float)(fr & 0xFFFFu) / (float)0xFFFF;
                    bias = -restitution * vRel - 0.1f * (-depth - slop) * invDt;
                }
            }
            const f3 nAngA = IA * crAn, nAngB = IB * crBn;
            const float friction = (float)(fr >> 16) / (float)0xFFFF;
            rows[slot] = mk4(rA, massN);
            rows[(size_t)maxContacts + slot] = mk4(rB, massT);
            rows[(size_t)2 * maxContacts + slot] = mk4(normal, bias);
            rows[(size_t)3 * maxContacts + slot] = mk4(tangent, friction);
            rows[(size_t)4 * maxContacts + slot] = mk4(nAngA, 0.f);
            rows[(size_t)5 * maxContacts + slot] = mk4(nAngB, 0.f);
            rows[(size_t)6 * maxContacts + slot] = mk4(tAngA, 0.f);
            rows[(size_t)7 * maxContacts + slot] = mk4(tAngB, 0.f);
            rowImp[slot] = make_float2(0.f, 0.f);
            slotOfContact[slot] = ci;
        }
    }
}

// solveCollisionVelocityConstraints (constraints.cpp:3381-3449) for one manifold
B200_D void solveManifold(uint32_t p, const uint32_t* __restrict__ order, const uint32_t* __restrict__ unitSlotStart, const uint4* __restrict__ manifolds,
                          const float4* __restrict__ rows, float2* __restrict__ rowImp, float4* __restrict__ gVel, uint32_t maxContacts)
{
    const uint4 man = manifolds[order[p]];
    const uint32_t slot0 = unitSlotStart[p];
    const float4 a0 = ldCg(gVel + 2 * man.x), a1 = ldCg(gVel + 2 * man.x + 1), b0 = ldCg(gVel + 2 * man.y), b1 = ldCg(gVel + 2 * man.y + 1);
    const float imA = a0.w, imB = b0.w;
    if (imA == 0.f && imB == 0.f) return;                        // :3394
    f3 vA = xyz(a0), wA = xyz(a1), vB = xyz(b0), wB = xyz(b1);
    for (uint32_t k = 0; k < man.w; ++k)
    {
        const uint32_t slot = slot0 + k;
        const float4 r0 = ldStream(rows + slot), r1 = ldStream(rows + (size_t)maxContacts + slot), r2 = ldStream(rows + (size_t)2 * maxContacts + slot),
                     r3 = ldStream(rows + (size_t)3 * maxContacts + slot), r4 = ldStream(rows + (size_t)4 * maxContacts + slot), r5 = ldStream(rows + (size_t)5 * maxContacts + slot),
                     r6 = ldStream(rows + (size_t)6 * maxContacts + slot), r7 = ldStream(rows + (size_t)7 * maxContacts + slot);
        float2 imp = rowImp[slot];
        const f3 rA = xyz(r0), rB = xyz(r1), normal = xyz(r2), tangent = xyz(r3);
        {   // tangent (friction) row first, clamped by the normal impulse of the PREVIOUS iteration (:3404-3424)
            const f3 rel = (vB + cross(wB, rB)) - (vA + cross(wA, rA));
            const float vt = dot(rel, tangent);
            float lambda = -r1.w * vt;
            const float maxF = r3.w * imp.x;
            const float ni = clampf(imp.y + lambda, -maxF, maxF);
            lambda = ni - imp.y; imp.y = ni;
            const f3 P = lambda * tangent;
            vA -= imA * P; wA -= xyz(r6) * lambda;
            vB += imB * P; wB += xyz(r7) * lambda;
        }
        {   // normal row (:3426-3442)
            const f3 rel = (vB + cross(wB, rB)) - (vA + cross(wA, rA));
            const float vn = dot(rel, normal);
            float lambda = -r0.w * (vn - r2.w);
            const float ni = fmax2(imp.x + lambda, 0.f);
            lambda = ni - imp.x; imp.x = ni;
            const f3 P = lambda * normal;
            vA -= imA * P; wA -= xyz(r4) * lambda;
            vB += imB * P; wB += xyz(r5) * lambda;
        }
        rowImp[slot] = imp;
    }
    if (imA != 0.f) { stCg(gVel + 2 * man.x, mk4(vA, imA)); stCg(gVel + 2 * man.x + 1, mk4(wA, 0.f)); }
    if (imB != 0.f) { stCg(gVel + 2 * man.y, mk4(vB, imB)); stCg(gVel + 2 * man.y + 1, mk4(wB, 0.f)); }
}

struct SolveArgs
{
    // contacts
    const uint32_t* order; const uint32_t* unitSlotStart; const uint4* manifolds; const float4* rows; float2* rowImp; uint32_t maxContacts;
    const uint32_t* colorStart;      // contacts: per-colour unit ranges (kMaxLevels + 1)
    JointSolveArgs joints;
    float4* gVel; const float4* gInvI;
    DeviceCounts* cnt; uint32_t iterations; uint32_t barrierSlot;
};

// the persistent PGS kernel: iterations x (joint kinds in reference order, then contacts) x colours, grid barrier between colours
__global__ void __launch_bounds__(kThreads) solveKernel(SolveArgs A)
{
    uint32_t epoch = 0;
    uint32_t* bar = &A.cnt->barrier[A.barrierSlot];
    const uint32_t tid = blockIdx.x * kThreads + threadIdx.x, stride = gridDim.x * kThreads;
    const uint32_t numContactColors = A.cnt->numColors[6];
    for (uint32_t it = 0; it < A.iterations; ++it)
    {
        for (int kind = 0; kind < 6; ++kind)
        {
            const uint32_t nc = A.cnt->numColors[kind];
            for (uint32_t c = 0; c < nc; ++c)
            {
                const uint32_t p0 = A.joints.colorStart[kind][c], p1 = A.joints.colorStart[kind][c + 1];
                for (uint32_t p = p0 + tid; p < p1; p += stride) solveJoint(kind, A.joints, p, A.gVel, A.gInvI);
                gridBarrier(bar, epoch);
            }
        }
        for (uint32_t c = 0; c < numContactColors; ++c)
        {
            const uint32_t p0 = A.colorStart[c], p1 = A.colorStart[c + 1];
            for (uint32_t p = p0 + tid; p < p1; p += stride) solveManifold(p, A.order, A.unitSlotStart, A.manifolds, A.rows, A.rowImp, A.gVel, A.maxContacts);
            gridBarrier(bar, epoch);
        }
    }
}

__global__ void __launch_bounds__(kThreads) manifoldBodiesKernel(const uint4* __restrict__ manifolds, uint2* __restrict__ unitBodies, const DeviceCounts* __restrict__ cnt)
{
    const uint32_t n = cnt->numManifolds;
    for (uint32_t u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads) { const uint4 m = manifolds[u]; unitBodies[u] = make_uint2(m.x, m.y); }
}

struct JointCounts { uint32_t n[6]; };
__global__ void solverResetKernel(DeviceCounts* cnt, JointCounts jc)
{
    if (threadIdx.x < kNumGroups) cnt->numColors[threadIdx.x] = 0;
    if (threadIdx.x < 8) cnt->barrier[threadIdx.x] = 0;
    if (threadIdx.x < 6) cnt->numJoints[threadIdx.x] = jc.n[threadIdx.x];
}

// colour one group of units (joint kind or contacts); leaves order[] and colour ranges behind
static void colorGroup(Context* c, uint32_t group, const uint2* unitBodies, const uint32_t* numUnitsDev, uint32_t maxUnits, bool exact,
                       uint32_t* color, uint32_t* order, uint32_t* colorStartOut, uint32_t barrierSlot)
{
    const uint32_t N = c->numBodies;
    const int G = gridFor(c, 8);
    cudaMemsetAsync(c->adjFill, 0, sizeof(uint32_t) * (N + 2), c->stream);
    cudaMemsetAsync(c->adjStart, 0, sizeof(uint32_t) * (N + 2), c->stream);
    uint32_t* adjCount = c->adjFill;   // counts are built in adjFill, scanned into adjStart, then adjFill is re-zeroed as the fill cursor
    adjCountKernel<<<G, kThreads, 0, c->stream>>>(unitBodies, numUnitsDev, c->gVel, N, adjCount);
    c->launches++;
    scanExclusiveU32(c, adjCount, c->adjStart, nullptr, N + 1, nullptr);
    cudaMemsetAsync(c->adjFill, 0, sizeof(uint32_t) * (N + 2), c->stream);
    adjFillKernel<<<G, kThreads, 0, c->stream>>>(unitBodies, numUnitsDev, c->gVel, N, c->adjStart, c->adjFill, c->adjList);
    c->launches++;

    uint32_t* roundDone = c->levelFill;                       // kMaxLevels + 1
    uint32_t* colorCount = c->levelFill + (kMaxLevels + 1);   // kMaxLevels + 1
    uint32_t* colorFill = c->levelFill + 2 * (kMaxLevels + 1);
    cudaMemsetAsync(c->levelFill, 0, sizeof(uint32_t) * 3 * (kMaxLevels + 1), c->stream);
    int exactI = exact ? 1 : 0;
    DeviceCounts* cnt = c->counts;
    void* args[] = { &exactI, &group, (void*)&unitBodies, (void*)&numUnitsDev, &c->gVel, (void*)&N, &c->adjStart, &c->adjList, &color, &order, &roundDone, &colorCount,
                     &colorStartOut, &colorFill, &cnt, &barrierSlot };
    cudaLaunchCooperativeKernel((void*)colorKernel, dim3(c->numSMs * 2), dim3(kThreads), args, 0, c->stream);
    c->launches++;
    (void)maxUnits;
}

void launchSolver(Context* c, float dt, uint32_t iterations, uint32_t flags)
{
    DeviceCounts* cnt = c->counts;
    const int G = gridFor(c, 8);
    const bool exact = (flags & B200PHYS_STEP_EXACT_ORDER) != 0;
    JointCounts jc; for (int k = 0; k < 6; ++k) jc.n[k] = c->numJoints[k];
    solverResetKernel<<<1, 32, 0, c->stream>>>(cnt, jc);
    c->launches++;

    // ---- joints: colour each kind, initialise rows
    JointSolveArgs ja;
    JointInitArgs ji;
    bool anyJoint = false;
    for (int k = 0; k < 6; ++k)
    {
        ja.order[k] = c->unitOrder[k]; ja.colorStart[k] = c->levelStart + (size_t)k * (kMaxLevels + 1); ja.rows[k] = c->jointRows[k];
        ji.in[k] = c->jointIn[k]; ji.bodies[k] = c->jointBodies[k]; ji.rows[k] = c->jointRows[k]; ji.count[k] = c->numJoints[k];
        if (c->numJoints[k])
        {
            anyJoint = true;
            colorGroup(c, (uint32_t)k, c->jointBodies[k], &cnt->numJoints[k], c->numJoints[k], exact, c->unitColor[k], c->unitOrder[k],
                       c->levelStart + (size_t)k * (kMaxLevels + 1), (uint32_t)k);
        }
    }
    if (anyJoint)
    {
        ji.bRot = c->bRot; ji.bCog = c->bCog; ji.gPos = c->gPos; ji.gVel = c->gVel; ji.gInvI = c->gInvI; ji.N = c->numBodies; ji.dt = dt; ji.cnt = cnt;
        initJointsKernel<<<gridFor(c, 4), 128, 0, c->stream>>>(ji);
        c->launches++;
    }

    // ---- contacts
    uint2* unitBodies = (uint2*)c->pairsTyped;     // scratch reuse: the typed pair list is dead after the bucket sort
    manifoldBodiesKernel<<<G, kThreads, 0, c->stream>>>(c->manifolds, unitBodies, cnt);
    c->launches++;
    uint32_t* contactColorStart = c->levelStart + (size_t)6 * (kMaxLevels + 1);
    colorGroup(c, 6, unitBodies, &cnt->numManifolds, c->maxManifolds, exact, c->unitColor[6], c->unitOrder[6], contactColorStart, 6);
    uint32_t* slotCount = c->pairDst;              // scratch reuse
    slotCountKernel<<<G, kThreads, 0, c->stream>>>(c->unitOrder[6], c->manifolds, slotCount, cnt);
    c->launches++;
    scanExclusiveU32(c, slotCount, c->unitSlotStart, &cnt->numManifolds, c->maxManifolds, nullptr);
    initContactRowsKernel<<<G, kThreads, 0, c->stream>>>(c->unitOrder[6], c->unitSlotStart, c->manifolds, c->contacts, c->gVel, c->gPos, c->gInvI,
                                                         c->rows, c->rowImp, c->slotOfContact, c->maxContacts, dt, cnt);
    c->launches++;

    // ---- iterations
    SolveArgs A;
    A.order = c->unitOrder[6]; A.unitSlotStart = c->unitSlotStart; A.manifolds = c->manifolds; A.rows = c->rows; A.rowImp = c->rowImp; A.maxContacts = c->maxContacts;
    A.colorStart = contactColorStart; A.joints = ja; A.gVel = c->gVel; A.gInvI = c->gInvI; A.cnt = cnt; A.iterations = iterations; A.barrierSlot = 7;
    void* args[] = { &A };
    if (c->timing) cudaEventRecord(c->ev[4], c->stream);
    cudaLaunchCooperativeKernel((void*)solveKernel, dim3(c->numSMs * 2), dim3(kThreads), args, 0, c->stream);
    c->launches++;
}

} // namespace b200
