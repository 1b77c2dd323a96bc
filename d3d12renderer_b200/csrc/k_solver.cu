// k_solver.cu -- sequential-impulse (PGS) constraint solver on the device.
//
// Reference: constraint_solver (src/physics/constraints.cpp:3711-3772), collision constraints :3307-3449.
// The reference visits constraints Gauss-Seidel in array order on one thread; its AVX2 path re-orders them into
// conflict-free 8-wide batches (constraints.cpp:51-184). Here the unit of scheduling is a MANIFOLD (all contacts of one
// collider pair, solved in order by one thread) or a joint. A step fixes a SOLVE ORDER of the units and derives from it,
// for every unit and each of its dynamic bodies, the unit's rank among that body's units; the sweep itself is a dataflow
// over per-body update counters (see solveKernel): no colour phases, no grid barriers.
//   * exact mode    : canonical array order => bit for bit the sequential Gauss-Seidel sweep.
//   * fast mode     : units sorted by a deterministic Jones-Plassmann colouring (fixed hash priority, <= 48 colours), computed
//                     by a queue-driven kernel without rounds or barriers: short dependency chains through the bodies.
//   * permuted mode : a fixed pseudo-random permutation instead of the colouring (nothing to compute, longer chains).
// The order is exported (b200phys_download_solve_order) so that the oracle can replay it. The static dummy body and kinematic
// bodies never order anything (constraints.cpp:82-83).
//
// HBM traffic per contact and iteration: one 128-byte row (8 coalesced 128-bit planes) + 8 B impulses read/written +
// two 32-byte body records gathered and scattered (the only per-body data the iterations touch).
#include "b200_context.h"
#include <cstring>
#include "b200_util.cuh"
#include "k_joints.cuh"
#include <cstdlib>

namespace b200 {

static_assert(sizeof(ConeTwistRowD) <= 512 && sizeof(SliderRowD) <= 512 && sizeof(HingeRowD) <= 512 && sizeof(FixedRowD) <= 512, "joint rows are allocated 512 bytes each");

static const int kThreads = 256;
static const int kSolveThreads = 512;
static const uint32_t kUncolored = 0xFFFFFFFFu;

B200_D bool bodyIsDynamic(const float4* __restrict__ gMass, uint32_t b, uint32_t N) { return b < N && gMass[b].w > 0.f; }
B200_D uint32_t hashPriority(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// ------------------------------------------------------------------------------------------------------ adjacency
__global__ void __launch_bounds__(kThreads) adjCountKernel(const uint2* __restrict__ unitBodies, const uint32_t* __restrict__ numUnits, const float4* __restrict__ gMass, uint32_t N, uint32_t* __restrict__ adjCount)
{
    const uint32_t n = *numUnits;
    for (uint32_t u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads)
    {
        const uint2 b = unitBodies[u];
        if (bodyIsDynamic(gMass, b.x, N)) atomicAdd(&adjCount[b.x], 1u);
        if (b.y != b.x && bodyIsDynamic(gMass, b.y, N)) atomicAdd(&adjCount[b.y], 1u);
    }
}
__global__ void __launch_bounds__(kThreads) adjFillKernel(const uint2* __restrict__ unitBodies, const uint32_t* __restrict__ numUnits, const float4* __restrict__ gMass, uint32_t N,
                                                          const uint32_t* __restrict__ adjStart, uint32_t* __restrict__ adjFill, uint32_t* __restrict__ adjList, uint2* __restrict__ unitSlots)
{
    const uint32_t n = *numUnits;
    for (uint32_t u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads)
    {
        const uint2 b = unitBodies[u];
        // entry = unit << 1 | which of the unit's two bodies this list belongs to
        uint2 slots = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
        if (bodyIsDynamic(gMass, b.x, N)) { slots.x = adjStart[b.x] + atomicAdd(&adjFill[b.x], 1u); adjList[slots.x] = u << 1; }
        if (b.y != b.x && bodyIsDynamic(gMass, b.y, N)) { slots.y = adjStart[b.y] + atomicAdd(&adjFill[b.y], 1u); adjList[slots.y] = (u << 1) | 1u; }
        if (unitSlots) unitSlots[u] = slots;        // where this unit sits in its bodies' lists (inherit mode: its solve position goes there later)
    }
}

// ------------------------------------------------------------------------------------------- colouring / levelling
// One cooperative kernel, "push" formulation of Jones-Plassmann:
//   pending[u]  = number of neighbours that must be coloured before u (higher hash priority; exact mode: smaller index)
//   forbid[u]   = fast: bit mask of colours taken by those neighbours; exact: 1 + max level of those neighbours
// A unit with pending == 0 takes the lowest free colour (or its level), ORs/maxes it into its lower-ranked neighbours and
// only then decrements their pending counters (release: __threadfence between, acquire on the reader's side), so the
// result does not depend on timing: every unit decides from exactly the colours of its higher-ranked neighbours.
// Work is proportional to 3 traversals of each unit's neighbourhood in total, independent of the number of rounds.
// Afterwards the units are counting-sorted by colour (histogram -> scan by CTA 0 -> scatter).
B200_D bool ranksBefore(int exact, uint32_t v, uint32_t pv, uint32_t u, uint32_t pu)
{   // true if v must be coloured before u
    return exact ? (v < u) : (pv > pu || (pv == pu && v > u));
}

static const uint32_t kNbrStride = 16;   // explicit "who waits for me" list per unit (falls back to adjacency traversal above this)

// visits every neighbour v != u of unit u (units sharing a dynamic body; a unit sharing both bodies is visited twice); 4 adjacency entries are loaded per trip so the
// loads overlap (the lists are short: latency, not bandwidth, is what costs here)
template <typename F>
B200_D void forEachNeighbour(uint32_t u, uint2 b, const float4* __restrict__ gMass, uint32_t N, const uint32_t* __restrict__ adjStart, const uint32_t* __restrict__ adjList, F&& f)
{
#pragma unroll
    for (int side = 0; side < 2; ++side)
    {
        const uint32_t body = side ? b.y : b.x;
        if (side && b.y == b.x) break;
        if (!bodyIsDynamic(gMass, body, N)) continue;
        const uint32_t a0 = adjStart[body], a1 = adjStart[body + 1];
        for (uint32_t a = a0; a < a1; a += 4)
        {
            uint32_t v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) v[j] = (a + j < a1) ? adjList[a + j] : (u << 1);
#pragma unroll
            for (int j = 0; j < 4; ++j) if ((v[j] >> 1) != u) f(v[j] >> 1, v[j] & 1u);      // neighbour, and which of ITS bodies is the shared one
        }
    }
}

// ---- queue-driven Jones-Plassmann colouring (fast order). The classic formulation re-scans all units once per round with a
// grid barrier in between (25 rounds x 11 us on the 778 k-manifold pile). Here the thread whose decrement brings a unit's
// pending count to zero hands the unit to a global ready queue, and every warp keeps taking 32 tickets from that queue: a
// unit is coloured as soon as its last higher-priority neighbour is, there is no barrier and no re-scan, and the result is the
// same deterministic colouring (every unit still decides from exactly the colours of its higher-priority neighbours).
// All same-address traffic is warp-aggregated: one atomic per 32 tickets, one per warp pass for the enqueues.
static const uint32_t kQueueEmpty = 0xFFFFFFFFu;
static const uint32_t kColorPollSleepNs = 200;
static const uint32_t kMaxFastColors = 26;     // colour bits per body-side field of the state word
static const int kColorRunThreads = 256;

__global__ void __launch_bounds__(kThreads) colorInitKernel(const uint2* __restrict__ unitBodies, const uint32_t* __restrict__ numUnitsPtr,
    const float4* __restrict__ gMass, uint32_t N, const uint32_t* __restrict__ adjStart, const uint32_t* __restrict__ adjList,
    uint32_t* __restrict__ color, unsigned long long* __restrict__ forbid, uint32_t* __restrict__ nbrCount, uint32_t* __restrict__ nbrList,
    uint32_t* __restrict__ queue, DeviceCounts* cnt)
{
    const uint32_t n = *numUnitsPtr;
    const uint32_t rounds = (n + gridDim.x * kThreads - 1) / (gridDim.x * kThreads);
    for (uint32_t r = 0; r < rounds; ++r)
    {
        const uint32_t u = (r * gridDim.x + blockIdx.x) * kThreads + threadIdx.x;
        bool seed = false;
        if (u < n)
        {
            const uint2 b = unitBodies[u];
            const uint32_t pu = hashPriority(u);
            uint32_t pend = 0, nw = 0;
            forEachNeighbour(u, b, gMass, N, adjStart, adjList, [&](uint32_t v, uint32_t vside)
            {
                if (ranksBefore(0, v, hashPriority(v), u, pu)) ++pend;
                else { if (nw < kNbrStride) nbrList[(size_t)u * kNbrStride + nw] = (v << 1) | vside; ++nw; }
            });
            // ONE 64-bit word per unit: pending << 52 | colours seen through body B << 26 | colours seen through body A. The units
            // around one body are a clique (pairwise distinct colours), so within one field every colour bit is added at most once
            // and a notifier needs a single atomicAdd (+ colour bit, - 1 pending) per waiting unit; the thread whose add brings the
            // count to zero has, by the order of the atomics on that word, every colour bit in what it reads back afterwards.
            if (pend > 0xFFFu) { pend = 0xFFFu; cnt->overflowColors = 1; }
            forbid[u] = (unsigned long long)pend << 52;
            color[u] = kUncolored; nbrCount[u] = nw;
            seed = pend == 0;
        }
        const uint32_t ballot = __ballot_sync(0xffffffffu, seed);
        if (ballot)
        {
            uint32_t base = 0;
            if (laneId() == 0) base = atomicAdd(&cnt->colorQueueTail, (uint32_t)__popc(ballot));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (seed) queue[base + __popc(ballot & laneMaskLt())] = u;
        }
    }
}

__global__ void __launch_bounds__(kColorRunThreads) colorRunKernel(uint32_t group, const uint2* __restrict__ unitBodies, const uint32_t* __restrict__ numUnitsPtr,
    const float4* __restrict__ gMass, uint32_t N, const uint32_t* __restrict__ adjStart, const uint32_t* __restrict__ adjList,
    uint32_t* __restrict__ color, unsigned long long* __restrict__ forbid, const uint32_t* __restrict__ nbrCount, const uint32_t* __restrict__ nbrList,
    uint32_t* __restrict__ queue, uint32_t* __restrict__ colorCount, DeviceCounts* cnt)
{
    __shared__ uint32_t sHist[kMaxColors];
    const uint32_t n = *numUnitsPtr;
    if (threadIdx.x < kMaxColors) sHist[threadIdx.x] = 0;
    __syncthreads();
    uint32_t localMax = 0;
    for (;;)
    {
        uint32_t base = 0;
        if (laneId() == 0) base = atomicAdd(&cnt->colorQueueHead, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t idx = base + laneId();
        bool done = idx >= n;
        uint32_t spins = 0;
        while (__any_sync(0xffffffffu, !done))
        {
            uint32_t ready[kNbrStride]; uint32_t nReady = 0;
            if (!done)
            {
                uint32_t u;
                asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(u) : "l"(queue + idx) : "memory");
                if (u != kQueueEmpty)
                {
                    done = true;
                    unsigned long long f;
                    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(f) : "l"(forbid + u) : "memory");
                    const uint32_t taken = (uint32_t)(f & 0x3FFFFFFull) | (uint32_t)((f >> 26) & 0x3FFFFFFull);
                    uint32_t c = (uint32_t)(__ffs((int)~taken) - 1);
                    if (c >= kMaxFastColors) { c = kMaxFastColors - 1; cnt->overflowColors = 1; }   // 26 colours per field
                    color[u] = c;
                    atomicAdd(&sHist[c], 1u);
                    localMax = max(localMax, c + 1u);
                    const uint32_t nw = nbrCount[u];
                    const unsigned long long dec = 0xFFF0000000000000ull;     // -1 in the pending field
                    if (nw <= kNbrStride)
                    {   // the waiting units are listed explicitly: 4 independent 128-bit loads, then the atomics back to back
                        uint32_t w[kNbrStride];
                        const uint4* L = reinterpret_cast<const uint4*>(nbrList + (size_t)u * kNbrStride);
#pragma unroll
                        for (int j = 0; j < 4; ++j) { const uint4 q = (4u * j < nw) ? L[j] : make_uint4(0, 0, 0, 0); w[4 * j] = q.x; w[4 * j + 1] = q.y; w[4 * j + 2] = q.z; w[4 * j + 3] = q.w; }
                        unsigned long long old[kNbrStride];
#pragma unroll
                        for (int j = 0; j < (int)kNbrStride; ++j) if ((uint32_t)j < nw) old[j] = atomicAdd(&forbid[w[j] >> 1], (1ull << (c + 26u * (w[j] & 1u))) + dec);
#pragma unroll
                        for (int j = 0; j < (int)kNbrStride; ++j) if ((uint32_t)j < nw && (old[j] >> 52) == 1ull) ready[nReady++] = w[j] >> 1;
                    }
                    else
                    {   // rare: more waiting units than the list holds -> walk the adjacency, enqueue one by one
                        const uint2 b = unitBodies[u];
                        const uint32_t pu = hashPriority(u);
                        forEachNeighbour(u, b, gMass, N, adjStart, adjList, [&](uint32_t v, uint32_t vside)
                        {
                            if (!ranksBefore(0, u, pu, v, hashPriority(v))) return;
                            const unsigned long long o = atomicAdd(&forbid[v], (1ull << (c + 26u * vside)) + dec);
                            if ((o >> 52) == 1ull) { const uint32_t t = atomicAdd(&cnt->colorQueueTail, 1u); if (t < n) queue[t] = v; }
                        });
                    }
                }
                else if (++spins > (1u << 22)) { cnt->overflowColors = 2; done = true; }   // refuse to hang (cyclic graph cannot happen)
            }
            // hand the newly ready units to the queue: one reservation per warp and pass (a pass in which nobody ran a unit
            // is a pure poll: skip the bookkeeping and back off a little so that the working warps get the issue slots)
            if (!__any_sync(0xffffffffu, nReady != 0u)) { if (__any_sync(0xffffffffu, !done)) __nanosleep(kColorPollSleepNs); continue; }
            uint32_t incl = nReady;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (laneId() >= (uint32_t)o) incl += t; }
            const uint32_t totalReady = __shfl_sync(0xffffffffu, incl, 31);
            uint32_t qb = 0;
            if (laneId() == 0) qb = atomicAdd(&cnt->colorQueueTail, totalReady);
            qb = __shfl_sync(0xffffffffu, qb, 0) + incl - nReady;
            for (uint32_t j = 0; j < nReady; ++j) if (qb + j < n) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" :: "l"(queue + qb + j), "r"(ready[j]) : "memory");
        }
    }
    localMax = __reduce_max_sync(0xffffffffu, localMax);
    if (laneId() == 0 && localMax) atomicMax(&cnt->numColors[group], localMax);
    __syncthreads();
    if (threadIdx.x < kMaxColors && sHist[threadIdx.x]) atomicAdd(&colorCount[threadIdx.x], sHist[threadIdx.x]);
}

// colour ranges (<= 48 colours: one warp) ...
__global__ void colorScanKernel(uint32_t group, const uint32_t* __restrict__ colorCount, uint32_t* __restrict__ colorStartOut, DeviceCounts* cnt)
{
    if (threadIdx.x == 0)
    {
        const uint32_t numColors = min(cnt->numColors[group], kMaxColors);
        uint32_t run = 0;
        for (uint32_t c = 0; c < numColors; ++c) { colorStartOut[c] = run; cnt->colorStart[group][c] = run; run += colorCount[c]; }
        colorStartOut[numColors] = run; cnt->colorStart[group][numColors] = run;
    }
}
// ... and the counting sort by colour: each CTA reserves one range per colour for its chunk of units, then ranks inside it in
// shared memory (no same-address global atomics per unit). Also collects the colours present at every dynamic body.
__global__ void __launch_bounds__(kThreads) colorScatterKernel(const uint2* __restrict__ unitBodies, const uint32_t* __restrict__ numUnitsPtr, const uint32_t* __restrict__ color,
    const float4* __restrict__ gMass, uint32_t N, const uint32_t* __restrict__ colorStart, uint32_t* __restrict__ colorFill, uint32_t* __restrict__ order, unsigned long long* __restrict__ bodyMask)
{
    __shared__ uint32_t sHist[kMaxColors], sBase[kMaxColors];
    const uint32_t n = *numUnitsPtr;
    const uint32_t per = (n + gridDim.x - 1) / gridDim.x, u0 = blockIdx.x * per, u1 = min(n, u0 + per);
    if (threadIdx.x < kMaxColors) sHist[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t u = u0 + threadIdx.x; u < u1; u += kThreads) atomicAdd(&sHist[min(color[u], kMaxColors - 1)], 1u);
    __syncthreads();
    if (threadIdx.x < kMaxColors) { const uint32_t h = sHist[threadIdx.x]; sBase[threadIdx.x] = h ? (colorStart[threadIdx.x] + atomicAdd(&colorFill[threadIdx.x], h)) : 0u; }
    __syncthreads();
    for (uint32_t u = u0 + threadIdx.x; u < u1; u += kThreads)
    {
        const uint32_t c = min(color[u], kMaxColors - 1);
        order[atomicAdd(&sBase[c], 1u)] = u;
        const uint2 b = unitBodies[u];
        const unsigned long long bit = 1ull << min(c, 63u);
        if (bodyIsDynamic(gMass, b.x, N)) atomicOr(&bodyMask[b.x], bit);
        if (b.y != b.x && bodyIsDynamic(gMass, b.y, N)) atomicOr(&bodyMask[b.y], bit);
    }
}

// ------------------------------------------------------------------------------------------------ contact rows
// row planes (float4 each, plane stride = maxContacts; a warp reads 512 contiguous bytes per plane), indexed by solve slot:
//  0: rA.xyz, massN   1: rB.xyz, massT   2: normal.xyz, bias   3: tangent.xyz, friction
//  4: nAngA.xyz, invMassA   5: nAngB.xyz, invMassB   6: tAngA.xyz   7: tAngB.xyz
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
    float4* __restrict__ rows, float2* __restrict__ rowImp, uint32_t* __restrict__ slotOfContact, uint4* __restrict__ unitInfo, uint32_t maxContacts, float dt, const DeviceCounts* __restrict__ cnt)
{
    const uint32_t n = cnt->numManifolds;
    const float invDt = 1.f / dt;
    for (uint32_t p = blockIdx.x * kThreads + threadIdx.x; p < n; p += gridDim.x * kThreads)
    {
        const uint4 man = manifolds[order[p]];
        const uint32_t slot0 = unitSlotStart[p];
        if (slot0 + man.w > maxContacts) { unitInfo[p] = make_uint4(man.x, man.y, 0u, 0u); continue; }   // cannot happen with consistent counts (npCompactKernel cuts both lists together)
        unitInfo[p] = make_uint4(man.x, man.y, slot0, man.w);   // solve-order record: one 128-bit load per manifold and iteration
        const float4 vA4 = gVel[2 * man.x], wA4 = gVel[2 * man.x + 1], vB4 = gVel[2 * man.y], wB4 = gVel[2 * man.y + 1];
        const float4 pA4 = gPos[man.x], pB4 = gPos[man.y];
        const f3 posA = xyz(pA4), posB = xyz(pB4);
        const m3 IA = loadInvInertia(gInvI, man.x), IB = loadInvInertia(gInvI, man.y);
        const float imA = pA4.w, imB = pB4.w;
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
            rows[(size_t)4 * maxContacts + slot] = mk4(nAngA, imA);
            rows[(size_t)5 * maxContacts + slot] = mk4(nAngB, imB);
            rows[(size_t)6 * maxContacts + slot] = mk4(tAngA, 0.f);
            rows[(size_t)7 * maxContacts + slot] = mk4(tAngB, 0.f);
            rowImp[slot] = make_float2(0.f, 0.f);
            slotOfContact[slot] = ci;
        }
    }
}

// solveCollisionVelocityConstraints (constraints.cpp:3381-3449): one contact row applied to the two body velocity records
struct RowRegs { float4 r[8]; float2 imp; };
B200_D void loadRow(RowRegs& R, const float4* __restrict__ rows, const float2* __restrict__ rowImp, uint32_t slot, uint32_t maxContacts)
{
#pragma unroll
    for (int j = 0; j < 8; ++j) R.r[j] = ldStream(rows + (size_t)j * maxContacts + slot);
    R.imp = rowImp[slot];
}
B200_D void applyRow(RowRegs& R, float imA, float imB, f3& vA, f3& wA, f3& vB, f3& wB)
{
    const f3 rA = xyz(R.r[0]), rB = xyz(R.r[1]), normal = xyz(R.r[2]), tangent = xyz(R.r[3]);
    {   // tangent (friction) row first, clamped by the normal impulse of the PREVIOUS iteration (:3404-3424)
        const f3 rel = (vB + cross(wB, rB)) - (vA + cross(wA, rA));
        const float vt = dot(rel, tangent);
        float lambda = -R.r[1].w * vt;
        const float maxF = R.r[3].w * R.imp.x;
        const float ni = clampf(R.imp.y + lambda, -maxF, maxF);
        lambda = ni - R.imp.y; R.imp.y = ni;
        const f3 P = lambda * tangent;
        vA -= imA * P; wA -= xyz(R.r[6]) * lambda;
        vB += imB * P; wB += xyz(R.r[7]) * lambda;
    }
    {   // normal row (:3426-3442)
        const f3 rel = (vB + cross(wB, rB)) - (vA + cross(wA, rA));
        const float vn = dot(rel, normal);
        float lambda = -R.r[0].w * (vn - R.r[2].w);
        const float ni = fmax2(R.imp.x + lambda, 0.f);
        lambda = ni - R.imp.x; R.imp.x = ni;
        const f3 P = lambda * normal;
        vA -= imA * P; wA -= xyz(R.r[4]) * lambda;
        vB += imB * P; wB += xyz(R.r[5]) * lambda;
    }
}

// ------------------------------------------------------------------------------------------- dataflow bookkeeping
// The PGS sweep is executed as a DATAFLOW, not as colour phases separated by grid barriers: every dynamic body carries an
// update counter (the .w lanes of its two 16-byte velocity halves, both written together with the data, so each half is
// self-validating and no fence is needed on either side). A unit (manifold or joint) knows, for each of its dynamic bodies,
// its rank among the units touching that body within one sweep and the body's degree; in sweep `it` it may run exactly
// when both counters read it * degree + rank, and it publishes rank + 1. The solve order (colour-sorted in fast mode,
// level-sorted canonical order in exact mode) only decides who comes first at each body; a unit never waits for anything
// but its true predecessors, so the sweep overlaps across colours and iterations and the memory system stays busy.
// dep record per solve position: .x / .y = rank | degree << 16 for body a / b, 0xFFFFFFFF = not dynamic (never waited for).
// A cheap bijection of [0, n): invertible mixing on ceil(log2 n) bits (odd multiplications, xor-shifts), cycle-walked back
// into range. The fast solve order is this pseudo-random permutation of the units: in a random order the dependency chains
// through the bodies are short (the same depth a Jones-Plassmann colouring needs rounds), and since the dataflow sweep only
// needs SOME total order per body, no colouring has to be computed at all.
B200_D uint32_t permuteBits(uint32_t x, uint32_t mask, uint32_t half)
{
    x = (x * 0x9E3779B1u) & mask; x ^= x >> half;
    x = (x * 0x85EBCA6Bu) & mask; x ^= x >> half;
    x = (x * 0xC2B2AE35u) & mask; x ^= x >> half;
    return x;
}
B200_D uint32_t permutePos(uint32_t u, uint32_t n)
{
    const uint32_t bits = n > 1u ? 32u - (uint32_t)__clz(n - 1u) : 1u;
    const uint32_t mask = bits >= 32u ? 0xFFFFFFFFu : ((1u << bits) - 1u), half = (bits + 1u) / 2u;
    uint32_t x = permuteBits(u, mask, half);
    while (x >= n) x = permuteBits(x, mask, half);
    return x;
}
// mode 1: canonical order (position = unit index); mode 2: the pseudo-random permutation
__global__ void __launch_bounds__(kThreads) orderKernel(int mode, const uint32_t* __restrict__ numUnits, uint32_t* __restrict__ order)
{
    const uint32_t n = *numUnits;
    for (uint32_t u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads) order[mode == 2 ? permutePos(u, n) : u] = u;
}

__global__ void __launch_bounds__(kThreads) colorQueueResetKernel(const uint32_t* __restrict__ numUnits, uint32_t* __restrict__ queue, DeviceCounts* cnt)
{
    const uint32_t n = *numUnits;
    if (blockIdx.x == 0 && threadIdx.x == 0) { cnt->colorQueueHead = 0; cnt->colorQueueTail = 0; }
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n + 32u; i += gridDim.x * kThreads) queue[i] = 0xFFFFFFFFu;
}

__global__ void __launch_bounds__(kThreads) depRankKernel(int mode /* 0 colours, 1 canonical index order, 2 permuted order, 3 positions from posOf */, const uint32_t* __restrict__ posOf, const uint2* __restrict__ unitBodies, const uint32_t* __restrict__ numUnits, const uint32_t* __restrict__ color,
    const uint32_t* __restrict__ order, const float4* __restrict__ gMass, uint32_t N, const unsigned long long* __restrict__ bodyMask, const uint32_t* __restrict__ bodyBase,
    const uint32_t* __restrict__ adjStart, const uint32_t* __restrict__ adjList, const uint32_t* __restrict__ adjPos, uint2* __restrict__ dep)
{
    const uint32_t n = *numUnits;
    for (uint32_t p = blockIdx.x * kThreads + threadIdx.x; p < n; p += gridDim.x * kThreads)
    {
        const uint32_t u = order[p];
        const uint2 b = unitBodies[u];
        uint32_t enc[2] = { 0xFFFFFFFFu, 0xFFFFFFFFu };
#pragma unroll
        for (int side = 0; side < 2; ++side)
        {
            const uint32_t body = side ? b.y : b.x;
            if (side && b.y == b.x) break;
            if (!bodyIsDynamic(gMass, body, N)) continue;
            uint32_t rank = bodyBase[body], deg = rank;
            if (mode)
            {   // the body's units in solve-position order
                const uint32_t a0 = adjStart[body], a1 = adjStart[body + 1];
                if (mode == 3) { for (uint32_t a = a0; a < a1; ++a) rank += (adjPos[a] < p) ? 1u : 0u; }
                else if (mode == 1) { for (uint32_t a = a0; a < a1; ++a) rank += ((adjList[a] >> 1) < u) ? 1u : 0u; }
                else { for (uint32_t a = a0; a < a1; ++a) rank += (permutePos(adjList[a] >> 1, n) < p) ? 1u : 0u; }
                deg += a1 - a0;
            }
            else
            {   // units sharing a body have distinct colours and the sweep runs colour by colour
                const unsigned long long m = bodyMask[body];
                rank += (uint32_t)__popcll(m & ((1ull << min(color[u], 63u)) - 1ull));
                deg += (uint32_t)__popcll(m);
            }
            enc[side] = (rank & 0xFFFFu) | (deg << 16);
        }
        dep[p] = make_uint2(enc[0], enc[1]);
    }
}

__global__ void __launch_bounds__(kThreads) bodyBaseAddKernel(int exact, uint32_t N, unsigned long long* __restrict__ bodyMask, const uint32_t* __restrict__ adjStart, uint32_t* __restrict__ bodyBase, DeviceCounts* cnt)
{
    for (uint32_t b = blockIdx.x * kThreads + threadIdx.x; b < N; b += gridDim.x * kThreads)
    {
        const uint32_t add = exact ? (adjStart[b + 1] - adjStart[b]) : (uint32_t)__popcll(bodyMask[b]);
        if (!exact) bodyMask[b] = 0ull;
        if (add) { const uint32_t t = bodyBase[b] + add; bodyBase[b] = t; if (t > 0xFFFFu) cnt->overflowColors = 8; }
    }
}

// ---- colour reuse across steps (contacts): the contact graph of a resting or slowly moving scene barely changes from one step to
// the next, and the dataflow sweep does not need a PROPER colouring, only a total order with short dependency chains. Every
// manifold has a key (type bucket, lower collider, higher collider); the manifold list is sorted by it in every step. Between two
// full colourings (every `recolorEvery` steps) a manifold inherits the colour its key had in the previous step (binary search in
// last step's sorted keys); new pairs get a hashed colour. Units are then stable-sorted by colour and ranked by POSITION, so two
// neighbours that ended up with the same colour are simply ordered by index: correct for any input, fast when most colours
// are still conflict-free.
B200_D uint32_t pairBucketOf(uint32_t ta, uint32_t tb) { if (ta > tb) { const uint32_t t = ta; ta = tb; tb = t; } return ta * 6u - (ta * (ta - 1u)) / 2u + (tb - ta); }
__global__ void __launch_bounds__(kThreads) manifoldKeysKernel(const uint2* __restrict__ collisions, const float4* __restrict__ wCol, unsigned long long* __restrict__ keys, const DeviceCounts* __restrict__ cnt)
{
    const uint32_t n = cnt->numManifolds;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
    {
        const uint2 pr = collisions[i];
        if (pr.y == 0xFFFFFFFFu) { keys[i] = (63ull << 56) | ((unsigned long long)pr.x << 28) | 0xFFFFFFFull; continue; }     // collider vs heightmap: no second collider
        const uint32_t ta = __float_as_uint(wCol[4 * pr.x + 3].x) & 0xFFu, tb = __float_as_uint(wCol[4 * pr.y + 3].x) & 0xFFu;
        keys[i] = ((unsigned long long)pairBucketOf(ta, tb) << 56) | ((unsigned long long)min(pr.x, pr.y) << 28) | (unsigned long long)max(pr.x, pr.y);
    }
}
__global__ void __launch_bounds__(kThreads) colorInheritKernel(const unsigned long long* __restrict__ keys, const unsigned long long* __restrict__ refKeys, const uint8_t* __restrict__ refColors,
    uint8_t* __restrict__ colorKey, uint32_t* __restrict__ color, DeviceCounts* cnt)
{
    const uint32_t n = cnt->numManifolds, nRef = cnt->numRefManifolds, numColors = max(cnt->refNumColors, 1u);
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt->numColors[6] = numColors;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
    {
        const unsigned long long k = keys[i];
        uint32_t lo = 0, hi = nRef;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (refKeys[mid] < k) lo = mid + 1; else hi = mid; }
        uint32_t c;
        if (lo < nRef && refKeys[lo] == k) c = refColors[lo];
        else c = hashPriority((uint32_t)k ^ (uint32_t)(k >> 28)) % numColors;       // a pair that did not touch last step
        c = min(c, kMaxColors - 1u);                                                // (a failed colouring can leave 255 behind: the stable sort takes keys < kMaxColors)
        colorKey[i] = (uint8_t)c; color[i] = c;
    }
}
// joints keep their index from step to step: inheriting is a copy
__global__ void __launch_bounds__(kThreads) jointColorCopyKernel(int toRef, const uint32_t* __restrict__ numUnits, uint32_t* __restrict__ color, uint8_t* __restrict__ colorKey, uint8_t* __restrict__ ref)
{
    const uint32_t n = *numUnits;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
    {
        if (toRef) ref[i] = (uint8_t)min(color[i], 255u);
        else { const uint32_t c = min((uint32_t)ref[i], kMaxColors - 1u); colorKey[i] = (uint8_t)c; color[i] = c; }
    }
}
// also copies every unit's solve position next to its entries in its bodies' adjacency lists: depRankKernel then counts the earlier units
// of a body by reading one contiguous list instead of gathering posOf[] per neighbour (164 -> 111 us at 2.8 M manifolds, +32 us for the scattered position writes here)
__global__ void __launch_bounds__(kThreads) orderFromPosKernel(const uint32_t* __restrict__ numUnits, const uint32_t* __restrict__ posOf, uint32_t* __restrict__ order,
                                                               const uint2* __restrict__ unitSlots, uint32_t* __restrict__ adjPos)
{
    const uint32_t n = *numUnits;
    for (uint32_t u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads)
    {
        const uint32_t p = posOf[u];
        order[p] = u;
        const uint2 sl = unitSlots[u];
        if (sl.x != 0xFFFFFFFFu) adjPos[sl.x] = p;
        if (sl.y != 0xFFFFFFFFu) adjPos[sl.y] = p;
    }
}
// remember this step's (key, colour) per manifold for the next step
__global__ void __launch_bounds__(kThreads) colorRememberKernel(const uint32_t* __restrict__ color, uint8_t* __restrict__ refColors, DeviceCounts* cnt)
{
    const uint32_t n = cnt->numManifolds;
    if (blockIdx.x == 0 && threadIdx.x == 0) { cnt->numRefManifolds = n; cnt->refNumColors = cnt->numColors[6]; }
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) refColors[i] = (uint8_t)min(color[i], 255u);
}

// with several groups (joint kinds, then contacts) a body's total degree is only known after the last group
struct DepFinalizeArgs { const uint2* unitBodies[kNumGroups]; const uint32_t* order[kNumGroups]; uint2* dep[kNumGroups]; const uint32_t* numUnits[kNumGroups]; };
__global__ void __launch_bounds__(kThreads) depFinalizeKernel(DepFinalizeArgs F, const uint32_t* __restrict__ bodyBase)
{
    for (uint32_t g = 0; g < kNumGroups; ++g)
    {
        if (!F.numUnits[g]) continue;
        const uint32_t n = *F.numUnits[g];
        for (uint32_t p = blockIdx.x * kThreads + threadIdx.x; p < n; p += gridDim.x * kThreads)
        {
            const uint2 b = F.unitBodies[g][F.order[g][p]];
            uint2 d = F.dep[g][p];
            if (d.x != 0xFFFFFFFFu) d.x = (d.x & 0xFFFFu) | (bodyBase[b.x] << 16);
            if (d.y != 0xFFFFFFFFu) d.y = (d.y & 0xFFFFu) | (bodyBase[b.y] << 16);
            F.dep[g][p] = d;
        }
    }
}

// ------------------------------------------------------------------------------------------------ the PGS sweep
B200_D float4 ldRelaxed4(const float4* p)
{
    float4 v;
    asm volatile("ld.relaxed.gpu.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
B200_D void stRelaxed4(float4* p, float4 v)
{
    asm volatile("st.relaxed.gpu.global.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
B200_D uint32_t ldRelaxedU32(const uint32_t* p) { uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }

struct SolveArgs
{
    const uint4* unitInfo; const uint2* unitDep; const float4* rows; float2* rowImp; uint32_t maxContacts;
    JointSolveArgs joints;
    float4* gVel; const float4* gPos; const float4* gInvI;
    DeviceCounts* cnt; uint32_t iterations;
};

static const uint32_t kSpinLimit = 1u << 22;    // polls before a unit gives up and aborts the sweep (seconds; never reached by a consistent graph)

// Persistent kernel, one CTA per SM (all co-resident: a unit may poll for a predecessor owned by any other warp). Global
// solve position q = joints of kind 0..5 in their solve order, then the manifolds; thread t owns positions t, t + T, ...
// in every sweep. A warp retries its 32 current units together until all have run, so a unit that depends on another lane's
// unit of the same warp (only at colour / level boundaries) cannot block it.
template <bool JOINTS>
__global__ void __launch_bounds__(kSolveThreads, 1) solveKernel(SolveArgs A)
{
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t T = gridDim.x * kSolveThreads;
    const uint32_t warpBase = blockIdx.x * kSolveThreads + (threadIdx.x & ~31u);
    uint32_t segEnd[7];
    {
        uint32_t run = 0;
#pragma unroll
        for (int k = 0; k < 6; ++k) { if (JOINTS) run += A.cnt->numJoints[k]; segEnd[k] = run; }
        run += A.cnt->numManifolds; segEnd[6] = run;
    }
    const uint32_t total = segEnd[6];
    const uint32_t* abortFlag = &A.cnt->solverAbort;
    bool aborted = false;
    for (uint32_t it = 0; it < A.iterations && !aborted; ++it)
    {
        for (uint32_t qb = warpBase; qb < total && !aborted; qb += T)
        {
            const uint32_t q = qb + lane;
            bool done = !(q < total);
            int kind = 6; uint32_t p = q;
            if (JOINTS && !done)
            {
                kind = 0;
#pragma unroll
                for (int k = 0; k < 6; ++k) if (q >= segEnd[k]) kind = k + 1;
                if (kind) p = q - segEnd[kind - 1];
            }
            uint32_t bodyA = 0, bodyB = 0, jointIndex = 0; uint2 dep = make_uint2(0xFFFFFFFFu, 0xFFFFFFFFu);
            uint4 info = make_uint4(0, 0, 0, 0); RowRegs R; float imA = 0.f, imB = 0.f;
            if (!done)
            {
                if (kind == 6)
                {
                    info = A.unitInfo[p]; dep = A.unitDep[p];
                    loadRow(R, A.rows, A.rowImp, info.z, A.maxContacts);
                    bodyA = info.x; bodyB = info.y; imA = R.r[4].w; imB = R.r[5].w;
                }
                else
                {
                    jointIndex = A.joints.order[kind][p]; dep = A.joints.dep[kind][p];
                    const uint2 jb = A.joints.bodies[kind][jointIndex];
                    bodyA = jb.x; bodyB = jb.y; imA = A.gPos[bodyA].w; imB = A.gPos[bodyB].w;
                }
                if (imA == 0.f && imB == 0.f) done = true;                 // constraints.cpp:3394 (and a joint between two fixed bodies moves nothing)
            }
            const bool dynA = dep.x != 0xFFFFFFFFu, dynB = dep.y != 0xFFFFFFFFu;
            const uint32_t expA = it * (dep.x >> 16) + (dep.x & 0xFFFFu), expB = it * (dep.y >> 16) + (dep.y & 0xFFFFu);
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0;
            if (!done)
            {   // bodies nobody writes (static, kinematic stand-in): read once
                if (!dynA) { a0 = A.gVel[2 * bodyA]; a1 = A.gVel[2 * bodyA + 1]; }
                if (!dynB) { b0 = A.gVel[2 * bodyB]; b1 = A.gVel[2 * bodyB + 1]; }
            }
            uint32_t spins = 0;
            while (__any_sync(0xffffffffu, !done))
            {
                if (done) continue;
                if (dynA) { a0 = ldRelaxed4(A.gVel + 2 * bodyA); a1 = ldRelaxed4(A.gVel + 2 * bodyA + 1); }
                if (dynB) { b0 = ldRelaxed4(A.gVel + 2 * bodyB); b1 = ldRelaxed4(A.gVel + 2 * bodyB + 1); }
                const bool ready = (!dynA || (__float_as_uint(a0.w) == expA && __float_as_uint(a1.w) == expA)) &&
                                   (!dynB || (__float_as_uint(b0.w) == expB && __float_as_uint(b1.w) == expB));
                if (!ready)
                {
                    if ((++spins & 255u) == 0u && (spins >= kSpinLimit || ldRelaxedU32(abortFlag)))
                    {   // inconsistent dependency records (colour overflow): stop the whole sweep instead of hanging the device
                        atomicExch(&A.cnt->solverAbort, 1u); A.cnt->overflowColors = 4; done = true; aborted = true;
                    }
                    continue;
                }
                f3 vA = xyz(a0), wA = xyz(a1), vB = xyz(b0), wB = xyz(b1);
                if (kind == 6)
                {
                    applyRow(R, imA, imB, vA, wA, vB, wB);
                    A.rowImp[info.z] = R.imp;
                    for (uint32_t k = 1; k < info.w; ++k)
                    {
                        RowRegs Rk; loadRow(Rk, A.rows, A.rowImp, info.z + k, A.maxContacts);
                        applyRow(Rk, imA, imB, vA, wA, vB, wB);
                        A.rowImp[info.z + k] = Rk.imp;
                    }
                }
                else
                {
                    BodyVel JA, JB; JA.v = vA; JA.w = wA; JA.invMass = imA; JB.v = vB; JB.w = wB; JB.invMass = imB;
                    solveJoint(kind, A.joints, jointIndex, JA, JB, A.gInvI);
                    vA = JA.v; wA = JA.w; vB = JB.v; wB = JB.w;
                }
                if (dynA) { const float ver = __uint_as_float(expA + 1u); stRelaxed4(A.gVel + 2 * bodyA, mk4(vA, ver)); stRelaxed4(A.gVel + 2 * bodyA + 1, mk4(wA, ver)); }
                if (dynB) { const float ver = __uint_as_float(expB + 1u); stRelaxed4(A.gVel + 2 * bodyB, mk4(vB, ver)); stRelaxed4(A.gVel + 2 * bodyB + 1, mk4(wB, ver)); }
                done = true;
            }
            aborted = __any_sync(0xffffffffu, aborted);
        }
    }
}

__global__ void __launch_bounds__(kThreads) manifoldBodiesKernel(const uint4* __restrict__ manifolds, uint2* __restrict__ unitBodies, const DeviceCounts* __restrict__ cnt)
{
    const uint32_t n = cnt->numManifolds;
    for (uint32_t u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads) { const uint4 m = manifolds[u]; unitBodies[u] = make_uint2(m.x, m.y); }
}

struct JointCounts { uint32_t n[6]; };
__global__ void solverResetKernel(DeviceCounts* cnt, JointCounts jc, uint32_t resetGroups)
{
    if (threadIdx.x < kNumGroups && ((resetGroups >> threadIdx.x) & 1u)) cnt->numColors[threadIdx.x] = 0;
    if (threadIdx.x < 6) cnt->numJoints[threadIdx.x] = jc.n[threadIdx.x];
    if (threadIdx.x == 0) cnt->solverAbort = 0;
}

// colour one group of units (joint kind or contacts); leaves order[] and colour ranges behind
static void colorGroup(Context* c, uint32_t group, const uint2* unitBodies, const uint32_t* numUnitsDev, uint32_t maxUnits, bool exact, bool permuted, bool inherit,
                       uint32_t* color, uint32_t* order, uint32_t* colorStartOut, uint32_t barrierSlot)
{
    uint32_t* posOf = c->unitPending;      // scratch, one word per unit
    const uint32_t N = c->numBodies;
    const int G = gridFor(c, 8);
    memsetK(c, c->adjFill, 0, sizeof(uint32_t) * (N + 2));
    memsetK(c, c->adjStart, 0, sizeof(uint32_t) * (N + 2));
    uint32_t* adjCount = c->adjFill;   // counts are built in adjFill, scanned into adjStart, then adjFill is re-zeroed as the fill cursor
    launchK(c, adjCountKernel, G, kThreads, 0, unitBodies, numUnitsDev, c->gPos, N, adjCount);
    c->launches++;
    scanExclusiveU32(c, adjCount, c->adjStart, nullptr, N + 1, nullptr);
    memsetK(c, c->adjFill, 0, sizeof(uint32_t) * (N + 2));
    const int mode = exact ? 1 : (permuted ? 2 : (inherit ? 3 : 0));
    launchK(c, adjFillKernel, G, kThreads, 0, unitBodies, numUnitsDev, c->gPos, N, c->adjStart, c->adjFill, c->adjList, mode == 3 ? c->unitSlots : (uint2*)nullptr);
    c->launches++;

    DeviceCounts* cnt = c->counts;
    if (mode == 3)
    {   // colours inherited from the previous step (colorInheritKernel ran before): stable sort by colour, rank by position
        stableSmallKeySort(c, c->colorKey, numUnitsDev, maxUnits, kMaxColors, posOf, cnt->colorStart[group]);
        launchK(c, orderFromPosKernel, G, kThreads, 0, numUnitsDev, posOf, order, c->unitSlots, c->adjPos);
        c->launches++;
    }
    else if (mode == 0)
    {   // deterministic Jones-Plassmann colouring, queue driven
        uint32_t* colorCount = c->levelFill + (kMaxLevels + 1);   // kMaxColors entries used
        uint32_t* colorFill = c->levelFill + 2 * (kMaxLevels + 1);
        uint32_t* queue = c->colorQueue;
        memsetK(c, colorCount, 0, sizeof(uint32_t) * kMaxColors);
        memsetK(c, colorFill, 0, sizeof(uint32_t) * kMaxColors);
        launchK(c, colorQueueResetKernel, G, kThreads, 0, numUnitsDev, queue, cnt);
        launchK(c, colorInitKernel, G, kThreads, 0, unitBodies, numUnitsDev, c->gPos, N, c->adjStart, c->adjList, color, c->unitForbid, c->nbrCount, c->nbrList, queue, cnt);
        launchK(c, colorRunKernel, c->numSMs * 4, kColorRunThreads, 0, group, unitBodies, numUnitsDev, c->gPos, N, c->adjStart, c->adjList, color, c->unitForbid, c->nbrCount, c->nbrList, queue, colorCount, cnt);
        launchK(c, colorScanKernel, 1, 32, 0, group, colorCount, colorStartOut, cnt);
        launchK(c, colorScatterKernel, G, kThreads, 0, unitBodies, numUnitsDev, color, c->gPos, N, colorStartOut, colorFill, order, c->bodyMask);
        c->launches += 5;
    }
    else
    {
        launchK(c, orderKernel, G, kThreads, 0, mode, numUnitsDev, order);
        c->launches++;
    }
    // dataflow records of this group (rank / degree of every unit at its bodies), then fold the group into the per-body totals
    launchK(c, depRankKernel, G, kThreads, 0, mode, posOf, unitBodies, numUnitsDev, color, order, c->gPos, N, c->bodyMask, c->bodyBase, c->adjStart, c->adjList, c->adjPos, c->unitDep[group]);
    launchK(c, bodyBaseAddKernel, G, kThreads, 0, mode != 0, N, c->bodyMask, c->adjStart, c->bodyBase, cnt);
    c->launches += 2;
    (void)maxUnits;
}

void launchSolver(Context* c, float dt, uint32_t iterations, uint32_t flags)
{
    DeviceCounts* cnt = c->counts;
    const int G = gridFor(c, 8);
    const bool exact = (flags & B200PHYS_STEP_EXACT_ORDER) != 0;
    const bool permuted = !exact && (flags & B200PHYS_STEP_PERMUTED) != 0;
    JointCounts jc; memset(&jc, 0, sizeof(jc)); for (int k = 0; k < 6; ++k) jc.n[k] = c->numJoints[k];
    // the joint groups' solve order / ranks / per-body counts of the previous step are still right if neither the joint topology nor the
    // body set changed and the mode is the same (fast): 11 launches per joint kind saved (22 of 70 on the ragdoll scenes)
    const bool fastMode = !exact && !permuted;
    bool anyJointKind = false; for (int k = 0; k < 6; ++k) anyJointKind |= c->numJoints[k] != 0;
    const bool jointsCached = fastMode && anyJointKind && c->jointCacheBodiesEpoch == c->bodiesEpoch && c->jointCacheJointsEpoch == c->jointsEpoch && c->jointCacheNumBodies == c->numBodies;
    launchK(c, solverResetKernel, 1, 32, 0, cnt, jc, jointsCached ? 0x40u : 0x7Fu);
    c->launches++;
    memsetK(c, c->bodyMask, 0, sizeof(unsigned long long) * (c->numBodies + 2));
    if (jointsCached) memcpyD2DK(c, c->bodyBase, c->bodyBaseJoints, sizeof(uint32_t) * (c->numBodies + 2));
    else memsetK(c, c->bodyBase, 0, sizeof(uint32_t) * (c->numBodies + 2));

    // ---- joints: colour each kind, initialise rows
    JointSolveArgs ja; memset(&ja, 0, sizeof(ja));      // padding bytes are part of the step hash (launchK)
    JointInitArgs ji; memset(&ji, 0, sizeof(ji));
    bool anyJoint = false;
    for (int k = 0; k < 6; ++k)
    {
        ja.order[k] = c->unitOrder[k]; ja.colorStart[k] = c->levelStart + (size_t)k * (kMaxLevels + 1); ja.rows[k] = c->jointRows[k]; ja.bodies[k] = c->jointBodies[k]; ja.dep[k] = c->unitDep[k];
        ji.in[k] = c->jointIn[k]; ji.bodies[k] = c->jointBodies[k]; ji.rows[k] = c->jointRows[k]; ji.count[k] = c->numJoints[k];
        if (c->numJoints[k])
        {
            anyJoint = true;
            if (jointsCached) continue;
            // the joint graph only changes when joints are uploaded with a different count: colour it once, then inherit
            const bool fast = !exact && !permuted;
            const bool inheritJ = fast && c->jointColorCount[k] == c->numJoints[k] && c->jointColorAge[k] > 0 && c->jointColorAge[k] < 64;
            if (inheritJ) { launchK(c, jointColorCopyKernel, G, kThreads, 0, 0, &cnt->numJoints[k], c->unitColor[k], c->colorKey, c->refColorsJoint[k]); c->launches++; }
            colorGroup(c, (uint32_t)k, c->jointBodies[k], &cnt->numJoints[k], c->numJoints[k], exact, permuted, inheritJ, c->unitColor[k], c->unitOrder[k],
                       c->levelStart + (size_t)k * (kMaxLevels + 1), (uint32_t)k);
            if (fast && !inheritJ) { launchK(c, jointColorCopyKernel, G, kThreads, 0, 1, &cnt->numJoints[k], c->unitColor[k], c->colorKey, c->refColorsJoint[k]); c->launches++; c->jointColorCount[k] = c->numJoints[k]; c->jointColorAge[k] = 0; }
            if (fast) ++c->jointColorAge[k]; else c->jointColorAge[k] = 0;
        }
    }
    if (anyJoint && !jointsCached)
    {
        if (fastMode)
        {
            memcpyD2DK(c, c->bodyBaseJoints, c->bodyBase, sizeof(uint32_t) * (c->numBodies + 2));
            c->jointCacheBodiesEpoch = c->bodiesEpoch; c->jointCacheJointsEpoch = c->jointsEpoch; c->jointCacheNumBodies = c->numBodies;
        }
        else c->jointCacheBodiesEpoch = 0;          // exact / permuted order overwrote the joint orders
    }
    if (anyJoint)
    {
        ji.bRot = c->bRot; ji.bCog = c->bCog; ji.gPos = c->gPos; ji.gVel = c->gVel; ji.gInvI = c->gInvI; ji.N = c->numBodies; ji.dt = dt; ji.cnt = cnt;
        launchK(c, initJointsKernel, gridFor(c, 4), 128, 0, ji);
        c->launches++;
    }

    // ---- contacts
    uint2* unitBodies = (uint2*)c->pairsTyped;     // scratch reuse: the typed pair list is dead after the bucket sort
    launchK(c, manifoldBodiesKernel, G, kThreads, 0, c->manifolds, unitBodies, cnt);
    c->launches++;
    uint32_t* contactColorStart = c->levelStart + (size_t)6 * (kMaxLevels + 1);
    // colour reuse: a full colouring every `recolorEvery` steps, inherited colours in between (fast mode only)
    static int recolorEvery = -1;
    if (recolorEvery < 0) { const char* e = getenv("B200PHYS_RECOLOR_EVERY"); recolorEvery = e ? atoi(e) : 0; }     // tuning override
    const int every = recolorEvery > 0 ? recolorEvery : (c->cfg.reserved[0] ? (int)c->cfg.reserved[0] : 8);
    const bool fastColored = !exact && !permuted;
    unsigned long long* keysCur = c->refKeys[c->refParity]; unsigned long long* keysRef = c->refKeys[c->refParity ^ 1];
    bool inherit = false;
    if (fastColored)
    {
        launchK(c, manifoldKeysKernel, G, kThreads, 0, c->collisions, c->wCol, keysCur, cnt);
        c->launches++;
        inherit = c->colorAge > 0 && c->colorAge < every;
        if (inherit) { launchK(c, colorInheritKernel, G, kThreads, 0, keysCur, keysRef, c->refColors, c->colorKey, c->unitColor[6], cnt); c->launches++; }
    }
    colorGroup(c, 6, unitBodies, &cnt->numManifolds, c->maxManifolds, exact, permuted, inherit, c->unitColor[6], c->unitOrder[6], contactColorStart, 6);
    if (fastColored)
    {
        launchK(c, colorRememberKernel, G, kThreads, 0, c->unitColor[6], c->refColors, cnt);
        c->launches++;
        c->refParity ^= 1;
        c->colorAge = (c->colorAge + 1 >= every) ? 0 : c->colorAge + 1;
    }
    else c->colorAge = 0;       // another mode ran: the remembered colours are stale
    timelineMark(c, "solver_order");
    uint32_t* slotCount = c->pairDst;              // scratch reuse
    launchK(c, slotCountKernel, G, kThreads, 0, c->unitOrder[6], c->manifolds, slotCount, cnt);
    c->launches++;
    scanExclusiveU32(c, slotCount, c->unitSlotStart, &cnt->numManifolds, c->maxManifolds, nullptr);
    launchK(c, initContactRowsKernel, G, kThreads, 0, c->unitOrder[6], c->unitSlotStart, c->manifolds, c->contacts, c->gVel, c->gPos, c->gInvI,
                                                         c->rows, c->rowImp, c->slotOfContact, c->unitInfo, c->maxContacts, dt, cnt);
    c->launches++;

    // ---- iterations
    if (anyJoint)
    {   // several groups: the degrees in the records of the earlier groups are partial until now
        DepFinalizeArgs F; memset(&F, 0, sizeof(F));
        for (uint32_t g = 0; g < kNumGroups; ++g)
        {
            const bool used = g == 6 || c->numJoints[g] != 0;
            F.unitBodies[g] = g == 6 ? unitBodies : c->jointBodies[g]; F.order[g] = c->unitOrder[g]; F.dep[g] = c->unitDep[g];
            F.numUnits[g] = !used ? nullptr : (g == 6 ? &cnt->numManifolds : &cnt->numJoints[g]);
        }
        launchK(c, depFinalizeKernel, G, kThreads, 0, F, c->bodyBase);
        c->launches++;
    }
    SolveArgs A; memset(&A, 0, sizeof(A));
    A.unitInfo = c->unitInfo; A.unitDep = c->unitDep[6]; A.rows = c->rows; A.rowImp = c->rowImp; A.maxContacts = c->maxContacts;
    A.joints = ja; A.gVel = c->gVel; A.gPos = c->gPos; A.gInvI = c->gInvI; A.cnt = cnt; A.iterations = iterations;
    void* args[] = { &A };
    if (c->timing) cudaEventRecord(c->ev[4], c->stream);
    timelineMark(c, "solver_rows");
    static int solveBlocksPerSM[2] = { 0, 0 };
    const int variant = anyJoint ? 1 : 0;
    void* fn = anyJoint ? (void*)solveKernel<true> : (void*)solveKernel<false>;
    if (!solveBlocksPerSM[variant]) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&solveBlocksPerSM[variant], fn, kSolveThreads, 0); if (solveBlocksPerSM[variant] < 1) solveBlocksPerSM[variant] = 1; if (solveBlocksPerSM[variant] > 1) solveBlocksPerSM[variant] = 1; }
    const int bps = solveBlocksPerSM[variant];
    if (c->graphMode == 1) { graphHashValue(c, (const void*)fn); graphHashValue(c, c->numSMs * bps); graphHashBytes(c, &A, sizeof(A)); }
    else cudaLaunchCooperativeKernel(fn, dim3(c->numSMs * bps), dim3(kSolveThreads), args, 0, c->stream);
    c->launches++;
    timelineMark(c, "solve");
}

} // namespace b200
