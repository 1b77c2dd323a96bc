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
#include <cstdlib>

namespace b200 {

static_assert(sizeof(ConeTwistRowD) <= 512 && sizeof(SliderRowD) <= 512 && sizeof(HingeRowD) <= 512 && sizeof(FixedRowD) <= 512, "joint rows are allocated 512 bytes each");

static const int kThreads = 256;
static const int kColorThreads = 1024;   // fat CTAs: fewer arrivals per grid barrier
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
                                                          const uint32_t* __restrict__ adjStart, uint32_t* __restrict__ adjFill, uint32_t* __restrict__ adjList)
{
    const uint32_t n = *numUnits;
    for (uint32_t u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads)
    {
        const uint2 b = unitBodies[u];
        if (bodyIsDynamic(gMass, b.x, N)) adjList[adjStart[b.x] + atomicAdd(&adjFill[b.x], 1u)] = u;
        if (b.y != b.x && bodyIsDynamic(gMass, b.y, N)) adjList[adjStart[b.y] + atomicAdd(&adjFill[b.y], 1u)] = u;
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

// visits every neighbour v != u of unit u (units sharing a dynamic body); 4 adjacency entries are loaded per trip so the
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
            for (int j = 0; j < 4; ++j) v[j] = (a + j < a1) ? adjList[a + j] : u;
#pragma unroll
            for (int j = 0; j < 4; ++j) if (v[j] != u) f(v[j]);
        }
    }
}

__global__ void __launch_bounds__(kColorThreads) colorKernel(int exact, uint32_t group, const uint2* __restrict__ unitBodies, const uint32_t* __restrict__ numUnitsPtr,
    const float4* __restrict__ gMass, uint32_t N, const uint32_t* __restrict__ adjStart, const uint32_t* __restrict__ adjList,
    uint32_t* __restrict__ color, uint32_t* __restrict__ order, uint32_t* __restrict__ pending, unsigned long long* __restrict__ forbid,
    uint32_t* __restrict__ nbrCount, uint32_t* __restrict__ nbrList,
    uint32_t* __restrict__ roundDone, uint32_t* __restrict__ colorCount /* kMaxLevels+1 */,
    uint32_t* __restrict__ colorStartOut /* kMaxLevels+1 */, uint32_t* __restrict__ colorFill, DeviceCounts* cnt, uint32_t barrierSlot)
{
    __shared__ uint32_t sRed[kColorThreads / 32];
    __shared__ uint32_t sHist[kMaxColors], sBase[kMaxColors];   // fast mode: per-CTA colour histogram / scatter bases (no same-address global atomics)
    uint32_t epoch = 0;
    uint32_t* bar = cnt->barrier[barrierSlot];
    const uint32_t n = *numUnitsPtr;
    const uint32_t maxColors = exact ? kMaxLevels : kMaxColors;
    const uint32_t tid = blockIdx.x * kColorThreads + threadIdx.x, stride = gridDim.x * kColorThreads;
    if (threadIdx.x < kMaxColors) sHist[threadIdx.x] = 0;
    __syncthreads();

    // ---- phase 0: dependency counts + the list of units that wait for u
    for (uint32_t u = tid; u < n; u += stride)
    {
        const uint2 b = unitBodies[u];
        const uint32_t pu = hashPriority(u);
        uint32_t pend = 0, nw = 0;
        forEachNeighbour(u, b, gMass, N, adjStart, adjList, [&](uint32_t v)
        {
            const uint32_t pv = hashPriority(v);
            if (ranksBefore(exact, v, pv, u, pu)) ++pend;
            else { if (nw < kNbrStride) nbrList[(size_t)u * kNbrStride + nw] = v; ++nw; }
        });
        // fast mode packs both into ONE 64-bit word (pending << 48 | colour mask): the notifier's OR and its decrement go to the
        // same address, which the memory model keeps in program order, so no fence is needed on either side
        if (exact) { pending[u] = pend; forbid[u] = 0ull; } else forbid[u] = (unsigned long long)pend << 48;
        color[u] = kUncolored; nbrCount[u] = nw;
    }
    gridBarrier(bar, epoch);

    uint32_t total = 0, localMax = 0, doneMask = 0;
    for (uint32_t round = 0; round < kMaxLevels && total < n; ++round)
    {
        uint32_t colouredHere = 0;
        for (uint32_t u = tid; u < n; u += stride)
        {
            const uint32_t k = (u - tid) / stride;                 // this thread's k-th unit: done bits live in a register
            if (k < 32 ? ((doneMask >> k) & 1u) : (color[u] != kUncolored)) continue;
            unsigned long long f;
            if (exact)
            {
                if (ldCgU32(pending + u) != 0) continue;
                __threadfence();                                   // acquire: the notifiers' level updates precede their pending decrements
                asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(f) : "l"(forbid + u) : "memory");
            }
            else
            {
                asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(f) : "l"(forbid + u) : "memory");
                if ((f >> 48) != 0ull) continue;
                f |= 0xFFFF000000000000ull;                        // only 48 colours are representable in fast mode
            }
            if (k < 32) doneMask |= (1u << k);
            uint32_t c = exact ? (uint32_t)f : (uint32_t)(__ffsll((long long)~f) - 1);
            if (c >= (exact ? maxColors : 48u)) { c = (exact ? maxColors : 48u) - 1; cnt->overflowColors = 1; }
            color[u] = c;
            if (exact) atomicAdd(&colorCount[c], 1u); else atomicAdd(&sHist[c], 1u);
            localMax = max(localMax, c + 1u);
            ++colouredHere;
            const uint32_t nw = nbrCount[u];
            if (nw <= kNbrStride)
            {   // the waiting units are listed explicitly: 4 independent 128-bit loads, then fire-and-forget atomics
                uint32_t w[kNbrStride];
                const uint4* L = reinterpret_cast<const uint4*>(nbrList + (size_t)u * kNbrStride);
#pragma unroll
                for (int j = 0; j < 4; ++j) { const uint4 q = (4u * j < nw) ? L[j] : make_uint4(0, 0, 0, 0); w[4 * j] = q.x; w[4 * j + 1] = q.y; w[4 * j + 2] = q.z; w[4 * j + 3] = q.w; }
                if (exact)
                {
#pragma unroll
                    for (int j = 0; j < (int)kNbrStride; ++j) if ((uint32_t)j < nw) atomicMax(&forbid[w[j]], (unsigned long long)(c + 1u));
                    __threadfence();
#pragma unroll
                    for (int j = 0; j < (int)kNbrStride; ++j) if ((uint32_t)j < nw) atomicSub(&pending[w[j]], 1u);
                }
                else
                {
#pragma unroll
                    for (int j = 0; j < (int)kNbrStride; ++j) if ((uint32_t)j < nw) { atomicOr(&forbid[w[j]], 1ull << c); atomicAdd(&forbid[w[j]], 0xFFFF000000000000ull); }   // += -(1 << 48)
                }
            }
            else
            {
                const uint2 b = unitBodies[u];
                const uint32_t pu = hashPriority(u);
                forEachNeighbour(u, b, gMass, N, adjStart, adjList, [&](uint32_t v)
                { if (ranksBefore(exact, u, pu, v, hashPriority(v))) { if (exact) atomicMax(&forbid[v], (unsigned long long)(c + 1u)); else { atomicOr(&forbid[v], 1ull << c); atomicAdd(&forbid[v], 0xFFFF000000000000ull); } } });
                if (exact)
                {
                    __threadfence();
                    forEachNeighbour(u, b, gMass, N, adjStart, adjList, [&](uint32_t v)
                    { if (ranksBefore(exact, u, pu, v, hashPriority(v))) atomicSub(&pending[v], 1u); });
                }
            }
        }
        uint32_t v = colouredHere;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (laneId() == 0) sRed[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < kColorThreads / 32; ++w) t += sRed[w]; if (t) atomicAdd(&roundDone[round], t); }
        gridBarrier(bar, epoch);
        const uint32_t done = ldCgU32(roundDone + round);
        if (done == 0 && total < n) { if (tid == 0) cnt->overflowColors = 2; break; }   // no progress: refuse to spin
        total += done;
        if (tid == 0) cnt->remaining = round + 1;   // diagnostic: colouring rounds of the last group
    }
    localMax = __reduce_max_sync(0xffffffffu, localMax);
    if (laneId() == 0 && localMax) atomicMax(&cnt->numColors[group], localMax);
    __syncthreads();
    if (!exact && threadIdx.x < kMaxColors && sHist[threadIdx.x]) atomicAdd(&colorCount[threadIdx.x], sHist[threadIdx.x]);
    gridBarrier(bar, epoch);

    // counting sort by colour: exclusive scan of the histogram (one CTA; <= 65536 entries), then scatter
    const uint32_t numColors = min(ldCgU32(&cnt->numColors[group]), maxColors);
    if (blockIdx.x == 0)
    {
        __shared__ uint32_t carry;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (uint32_t base = 0; base < numColors; base += kColorThreads)
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
            if (threadIdx.x == kColorThreads - 1) carry = c0 + woff + incl;
            __syncthreads();
        }
        if (threadIdx.x == 0) { colorStartOut[numColors] = carry; if (!exact && numColors <= kMaxColors) cnt->colorStart[group][numColors] = carry; }
    }
    gridBarrier(bar, epoch);
    if (exact)
    {
        for (uint32_t u = tid; u < n; u += stride)
        {
            const uint32_t c = min(color[u], maxColors - 1);
            const uint32_t pos = ldCgU32(colorStartOut + c) + atomicAdd(&colorFill[c], 1u);
            order[pos] = u;
        }
    }
    else
    {   // each CTA reserves one range per colour (its histogram from the rounds above), then ranks inside it in shared memory
        if (threadIdx.x < kMaxColors)
        {
            const uint32_t h = sHist[threadIdx.x];
            sBase[threadIdx.x] = h ? (ldCgU32(colorStartOut + threadIdx.x) + atomicAdd(&colorFill[threadIdx.x], h)) : 0u;
        }
        __syncthreads();
        for (uint32_t u = tid; u < n; u += stride)
        {
            const uint32_t c = min(color[u], maxColors - 1);
            order[atomicAdd(&sBase[c], 1u)] = u;
        }
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
__global__ void __launch_bounds__(kThreads) depMaskKernel(const uint2* __restrict__ unitBodies, const uint32_t* __restrict__ numUnits, const uint32_t* __restrict__ color,
                                                          const float4* __restrict__ gMass, uint32_t N, unsigned long long* __restrict__ bodyMask)
{
    const uint32_t n = *numUnits;
    for (uint32_t u = blockIdx.x * kThreads + threadIdx.x; u < n; u += gridDim.x * kThreads)
    {
        const uint2 b = unitBodies[u];
        const unsigned long long bit = 1ull << min(color[u], 63u);
        if (bodyIsDynamic(gMass, b.x, N)) atomicOr(&bodyMask[b.x], bit);
        if (b.y != b.x && bodyIsDynamic(gMass, b.y, N)) atomicOr(&bodyMask[b.y], bit);
    }
}

__global__ void __launch_bounds__(kThreads) depRankKernel(int exact, const uint2* __restrict__ unitBodies, const uint32_t* __restrict__ numUnits, const uint32_t* __restrict__ color,
    const uint32_t* __restrict__ order, const float4* __restrict__ gMass, uint32_t N, const unsigned long long* __restrict__ bodyMask, const uint32_t* __restrict__ bodyBase,
    const uint32_t* __restrict__ adjStart, const uint32_t* __restrict__ adjList, uint2* __restrict__ dep)
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
            if (exact)
            {   // canonical order: by unit index
                const uint32_t a0 = adjStart[body], a1 = adjStart[body + 1];
                for (uint32_t a = a0; a < a1; ++a) rank += (adjList[a] < u) ? 1u : 0u;
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
__global__ void solverResetKernel(DeviceCounts* cnt, JointCounts jc)
{
    if (threadIdx.x < kNumGroups) cnt->numColors[threadIdx.x] = 0;
    for (int b = 0; b < 8; ++b) cnt->barrier[b][threadIdx.x] = 0;
    if (threadIdx.x < 6) cnt->numJoints[threadIdx.x] = jc.n[threadIdx.x];
    if (threadIdx.x == 0) cnt->solverAbort = 0;
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
    adjCountKernel<<<G, kThreads, 0, c->stream>>>(unitBodies, numUnitsDev, c->gPos, N, adjCount);
    c->launches++;
    scanExclusiveU32(c, adjCount, c->adjStart, nullptr, N + 1, nullptr);
    cudaMemsetAsync(c->adjFill, 0, sizeof(uint32_t) * (N + 2), c->stream);
    adjFillKernel<<<G, kThreads, 0, c->stream>>>(unitBodies, numUnitsDev, c->gPos, N, c->adjStart, c->adjFill, c->adjList);
    c->launches++;

    uint32_t* roundDone = c->levelFill;                       // kMaxLevels + 1
    uint32_t* colorCount = c->levelFill + (kMaxLevels + 1);   // kMaxLevels + 1
    uint32_t* colorFill = c->levelFill + 2 * (kMaxLevels + 1);
    cudaMemsetAsync(c->levelFill, 0, sizeof(uint32_t) * 3 * (kMaxLevels + 1), c->stream);
    int exactI = exact ? 1 : 0;
    DeviceCounts* cnt = c->counts;
    void* args[] = { &exactI, &group, (void*)&unitBodies, (void*)&numUnitsDev, &c->gPos /* .w = invMass */, (void*)&N, &c->adjStart, &c->adjList, &color, &order, &c->unitPending, &c->unitForbid, &c->nbrCount, &c->nbrList,
                     &roundDone, &colorCount, &colorStartOut, &colorFill, &cnt, &barrierSlot };
    static int colorBlocksPerSM = 0;
    if (!colorBlocksPerSM) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&colorBlocksPerSM, colorKernel, kColorThreads, 0); if (colorBlocksPerSM < 1) colorBlocksPerSM = 1; if (colorBlocksPerSM > 8) colorBlocksPerSM = 8; }
    cudaLaunchCooperativeKernel((void*)colorKernel, dim3(c->numSMs * colorBlocksPerSM), dim3(kColorThreads), args, 0, c->stream);
    c->launches++;
    // dataflow records of this group (rank / degree of every unit at its bodies), then fold the group into the per-body totals
    if (!exact) { depMaskKernel<<<G, kThreads, 0, c->stream>>>(unitBodies, numUnitsDev, color, c->gPos, N, c->bodyMask); c->launches++; }
    depRankKernel<<<G, kThreads, 0, c->stream>>>(exactI, unitBodies, numUnitsDev, color, order, c->gPos, N, c->bodyMask, c->bodyBase, c->adjStart, c->adjList, c->unitDep[group]);
    bodyBaseAddKernel<<<G, kThreads, 0, c->stream>>>(exactI, N, c->bodyMask, c->adjStart, c->bodyBase, cnt);
    c->launches += 2;
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
    cudaMemsetAsync(c->bodyMask, 0, sizeof(unsigned long long) * (c->numBodies + 2), c->stream);
    cudaMemsetAsync(c->bodyBase, 0, sizeof(uint32_t) * (c->numBodies + 2), c->stream);

    // ---- joints: colour each kind, initialise rows
    JointSolveArgs ja;
    JointInitArgs ji;
    bool anyJoint = false;
    for (int k = 0; k < 6; ++k)
    {
        ja.order[k] = c->unitOrder[k]; ja.colorStart[k] = c->levelStart + (size_t)k * (kMaxLevels + 1); ja.rows[k] = c->jointRows[k]; ja.bodies[k] = c->jointBodies[k]; ja.dep[k] = c->unitDep[k];
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
                                                         c->rows, c->rowImp, c->slotOfContact, c->unitInfo, c->maxContacts, dt, cnt);
    c->launches++;

    // ---- iterations
    if (anyJoint)
    {   // several groups: the degrees in the records of the earlier groups are partial until now
        DepFinalizeArgs F;
        for (uint32_t g = 0; g < kNumGroups; ++g)
        {
            const bool used = g == 6 || c->numJoints[g] != 0;
            F.unitBodies[g] = g == 6 ? unitBodies : c->jointBodies[g]; F.order[g] = c->unitOrder[g]; F.dep[g] = c->unitDep[g];
            F.numUnits[g] = !used ? nullptr : (g == 6 ? &cnt->numManifolds : &cnt->numJoints[g]);
        }
        depFinalizeKernel<<<G, kThreads, 0, c->stream>>>(F, c->bodyBase);
        c->launches++;
    }
    SolveArgs A;
    A.unitInfo = c->unitInfo; A.unitDep = c->unitDep[6]; A.rows = c->rows; A.rowImp = c->rowImp; A.maxContacts = c->maxContacts;
    A.joints = ja; A.gVel = c->gVel; A.gPos = c->gPos; A.gInvI = c->gInvI; A.cnt = cnt; A.iterations = iterations;
    void* args[] = { &A };
    if (c->timing) cudaEventRecord(c->ev[4], c->stream);
    static int solveBlocksPerSM[2] = { 0, 0 };
    const int variant = anyJoint ? 1 : 0;
    void* fn = anyJoint ? (void*)solveKernel<true> : (void*)solveKernel<false>;
    if (!solveBlocksPerSM[variant]) { cudaOccupancyMaxActiveBlocksPerMultiprocessor(&solveBlocksPerSM[variant], fn, kSolveThreads, 0); if (solveBlocksPerSM[variant] < 1) solveBlocksPerSM[variant] = 1; if (solveBlocksPerSM[variant] > 1) solveBlocksPerSM[variant] = 1; }
    const int bps = solveBlocksPerSM[variant];
    cudaLaunchCooperativeKernel(fn, dim3(c->numSMs * bps), dim3(kSolveThreads), args, 0, c->stream);
    c->launches++;
}

} // namespace b200
