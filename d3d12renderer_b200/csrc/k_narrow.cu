// k_narrow.cu -- narrow phase driver on the device (reference src/physics/collision_narrow.cpp:2328-2603):
//   classify   : prune (no rigid body / same body), order each pair by collider type (strict <, :2374), bucket id
//   bucket sort: stable counting sort into the 21 upper-triangular type buckets (:2399-2453) -> warps are type-uniform
//   collide    : one thread per pair, three kernels by cost class (closed form / SAT+clip / GJK+EPA)
//   compact    : exclusive scan of (contacts, collided) -> dense contact + manifold arrays in canonical order
//                (bucket, colliderA, colliderB, slot); material combine and 16+16 bit packing as in :2221-2253
#include "b200_context.h"
#include "k_narrow_shapes.cuh"
#include <cstdlib>
#include <cstring>

namespace b200 {

static const int kThreads = 256;

B200_D uint32_t bucketOf(uint32_t ta, uint32_t tb) { return ta * 6u - (ta * (ta - 1u)) / 2u + (tb - ta); }

// A pair whose bounding spheres (worldCollidersKernel) are clearly apart cannot collide: it is dropped here, before the bucket sort, instead of
// being answered by its routine (the reference runs the routine and gets "no contact", collision_narrow.cpp:2423-2453; a miss leaves no data
// behind, so any proof of separation gives the same result). The slack -- 0.2 % of the radii, 1e-4, and 1e-5 of the coordinates for the rounding
// of the centres -- is orders of magnitude above what fp32 GJK / SAT can misjudge, so no pair the reference's arithmetic would report is lost
// (the contact lists are compared with the oracle's bit for bit). On a mixed pile more than half of the GJK buckets' misses end here.
B200_D bool spheresApart(float4 a, float4 b)
{
    const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
    const float reach = (a.w + b.w) * 1.002f + 1e-4f + 1e-5f * (fabsf(a.x) + fabsf(a.y) + fabsf(a.z) + fabsf(b.x) + fabsf(b.y) + fabsf(b.z));
    return dx * dx + dy * dy + dz * dz > reach * reach;         // (false for NaN: the pair goes to its routine)
}

// Second proof for the pairs whose spheres are not apart: the other shape's sphere against THIS shape's box when it has one (world AABB, OBB,
// or the box of a hull's vertices in the hull's frame): closest point of the box to the sphere centre, in the box frame. Same slack.
B200_D bool sphereBoxApart(float4 s, const float4* __restrict__ wCol, const float4* __restrict__ hullAabb, uint32_t i, uint32_t type, uint32_t geom)
{
    const float4 r0 = wCol[4 * i], r1 = wCol[4 * i + 1];
    f3 p, lo, hi; float ref;
    if (type == B200PHYS_AABB) { p = xyz(s); lo = mk3(r0.x, r0.y, r0.z); hi = mk3(r0.w, r1.x, r1.y); ref = 0.f; }
    else
    {
        const q4 q = toq(r0); const f3 c = mk3(r1.x, r1.y, r1.z);
        p = conjugate(q) * (xyz(s) - c);
        ref = fabsf(c.x) + fabsf(c.y) + fabsf(c.z);
        if (type == B200PHYS_OBB) { const float4 r2 = wCol[4 * i + 2]; hi = mk3(r1.w, r2.x, r2.y); lo = -hi; }
        else { lo = xyz(hullAabb[2 * geom]); hi = xyz(hullAabb[2 * geom + 1]); }
    }
    const float dx = p.x - clampf(p.x, lo.x, hi.x), dy = p.y - clampf(p.y, lo.y, hi.y), dz = p.z - clampf(p.z, lo.z, hi.z);
    const float reach = s.w * 1.002f + 1e-4f + 1e-5f * (fabsf(s.x) + fabsf(s.y) + fabsf(s.z) + ref);
    return dx * dx + dy * dy + dz * dz > reach * reach;
}

__global__ void __launch_bounds__(kThreads) npClassifyKernel(const uint2* __restrict__ pairs, const float4* __restrict__ wCol, const float4* __restrict__ bSphere, const float4* __restrict__ hullAabb,
                                                             uint8_t* __restrict__ pairKey, uint2* __restrict__ typed, const DeviceCounts* __restrict__ cnt)
{
    const uint32_t n = cnt->numPairs;
    for (uint32_t p = blockIdx.x * kThreads + threadIdx.x; p < n; p += gridDim.x * kThreads)
    {
        uint2 pr = pairs[p];
        const float4 ma = wCol[4 * pr.x + 3], mb = wCol[4 * pr.y + 3];
        const uint32_t twa = __float_as_uint(ma.x), twb = __float_as_uint(mb.x);
        uint32_t ta = twa & 0xFFu, tb = twb & 0xFFu; const uint32_t oa = twa >> 8, ob = twb >> 8;
        const uint32_t ba = __float_as_uint(ma.y), bb = __float_as_uint(mb.y);
        uint32_t key;
        if (oa != B200PHYS_OBJ_RIGID_BODY && ob != B200PHYS_OBJ_RIGID_BODY) key = kBucketDropped;                       // :2358-2362
        else if (oa == B200PHYS_OBJ_RIGID_BODY && ob == B200PHYS_OBJ_RIGID_BODY && ba == bb) key = kBucketDropped;      // :2364-2369
        else
        {
            float oa_geom = ma.z, ob_geom = mb.z;                                                                   // hull geometry index bits (worldCollidersKernel)
            if (!(ta < tb)) { uint32_t t = pr.x; pr.x = pr.y; pr.y = t; t = ta; ta = tb; tb = t; const float g = oa_geom; oa_geom = ob_geom; ob_geom = g; }   // equal types swap too (:2374)
            const bool collision = (oa == B200PHYS_OBJ_RIGID_BODY && ob == B200PHYS_OBJ_RIGID_BODY) || oa == B200PHYS_OBJ_STATIC || ob == B200PHYS_OBJ_STATIC; // :2378-2379
            key = collision ? bucketOf(ta, tb) : kBucketOverlap;
            if (collision && bSphere)
            {
                const float4 sa = bSphere[pr.x], sb = bSphere[pr.y];
                bool apart = spheresApart(sa, sb);
                if (!apart && ta >= B200PHYS_AABB && !(ta == B200PHYS_AABB && tb == B200PHYS_AABB))     // (box vs box in world axes: the broad phase already knows)
                    apart = sphereBoxApart(sb, wCol, hullAabb, pr.x, ta, __float_as_uint(oa_geom));
                if (!apart && tb >= B200PHYS_AABB && !(ta == B200PHYS_AABB && tb == B200PHYS_AABB))
                    apart = sphereBoxApart(sa, wCol, hullAabb, pr.y, tb, __float_as_uint(ob_geom));
                if (apart) key = kBucketDropped;
            }
        }
        pairKey[p] = (uint8_t)key;
        typed[p] = pr;
    }
}

__global__ void __launch_bounds__(kThreads) npPermuteKernel(const uint2* __restrict__ typed, const uint32_t* __restrict__ dst, uint2* __restrict__ sorted, const DeviceCounts* __restrict__ cnt)
{
    const uint32_t n = cnt->numPairs;
    for (uint32_t p = blockIdx.x * kThreads + threadIdx.x; p < n; p += gridDim.x * kThreads) sorted[dst[p]] = typed[p];
}

// ------------------------------------------------------------------------------------------------- shape loading
struct ColRec { float4 r0, r1, r2, r3; };
B200_D ColRec loadCol(const float4* __restrict__ wCol, uint32_t i) { ColRec c; c.r0 = wCol[4 * i]; c.r1 = wCol[4 * i + 1]; c.r2 = wCol[4 * i + 2]; c.r3 = wCol[4 * i + 3]; return c; }
B200_D SphereS asSphere(const ColRec& c) { SphereS s; s.c = mk3(c.r0.x, c.r0.y, c.r0.z); s.r = c.r0.w; return s; }
B200_D CapS asCap(const ColRec& c) { CapS s; s.a = mk3(c.r0.x, c.r0.y, c.r0.z); s.b = mk3(c.r0.w, c.r1.x, c.r1.y); s.r = c.r1.z; return s; }
B200_D BoxS asBox(const ColRec& c) { BoxS s; s.mn = mk3(c.r0.x, c.r0.y, c.r0.z); s.mx = mk3(c.r0.w, c.r1.x, c.r1.y); return s; }
B200_D ObbS asObb(const ColRec& c) { ObbS s; s.q = mkq(c.r0.x, c.r0.y, c.r0.z, c.r0.w); s.c = mk3(c.r1.x, c.r1.y, c.r1.z); s.r = mk3(c.r1.w, c.r2.x, c.r2.y); return s; }
B200_D HullS asHull(const ColRec& c, const float4* hullVerts, const uint32_t* hullOffsets)
{
    HullS h; h.q = mkq(c.r0.x, c.r0.y, c.r0.z, c.r0.w); h.p = mk3(c.r1.x, c.r1.y, c.r1.z);
    const uint32_t g = __float_as_uint(c.r3.z);
    const uint32_t o0 = hullOffsets[g], o1 = hullOffsets[g + 1];
    h.verts = hullVerts + o0; h.n = o1 - o0;
    return h;
}
B200_D SupportD supportFor(uint32_t type, const ColRec& c, const float4* hullVerts, const uint32_t* hullOffsets)
{
    switch (type)
    {
    case B200PHYS_SPHERE: return supSphere(asSphere(c));
    case B200PHYS_CAPSULE: return supCapsule(asCap(c));
    case B200PHYS_CYLINDER: return supCylinder(asCap(c));
    case B200PHYS_AABB: return supBox(asBox(c));
    case B200PHYS_OBB: return supObb(asObb(c));
    default: return supHull(asHull(c, hullVerts, hullOffsets));
    }
}

B200_D void storeManifold(float4* __restrict__ manTmp, unsigned long long* __restrict__ manScan, uint32_t q, uint32_t maxPairs, bool hit, const Man& m)
{
    const uint32_t cnt = hit ? m.cnt : 0u;
    manScan[q] = ((unsigned long long)cnt << 32) | (cnt ? 1ull : 0ull);
    if (cnt)
    {
        manTmp[q] = make_float4(m.n.x, m.n.y, m.n.z, 0.f);
        for (uint32_t k = 0; k < cnt; ++k) manTmp[(size_t)(k + 1) * maxPairs + q] = make_float4(m.p[k].x, m.p[k].y, m.p[k].z, m.d[k]);
    }
}

B200_D void flushCounters(const NarrowCounters& nc, DeviceCounts* cnt)
{
    if (nc.epaOverflow) atomicAdd(&cnt->epaOverflows, nc.epaOverflow);
    if (nc.gjkCap) atomicAdd(&cnt->gjkIterationCaps, nc.gjkCap);
    if (nc.asserts) atomicAdd(&cnt->assertCount, nc.asserts);
}

// cost class 0: closed-form pairs (buckets 0,1,2,3,4,6,7,15); class 1: SAT + clipping (16,18); the rest is GJK (+ EPA) based
template <int CLASS>
__global__ void __launch_bounds__(CLASS == 0 ? 256 : 128) npCollideKernel(const uint2* __restrict__ sorted, const float4* __restrict__ wCol, float4* __restrict__ manTmp,
    unsigned long long* __restrict__ manScan, uint32_t maxPairs, DeviceCounts* cnt)
{
    constexpr uint32_t list0[8] = { 0, 1, 2, 3, 4, 6, 7, 15 };
    constexpr uint32_t list1[2] = { 16, 18 };
    constexpr int nList = CLASS == 0 ? 8 : 2;
    NarrowCounters nc; nc.epaOverflow = nc.gjkCap = nc.asserts = 0;
    for (int li = 0; li < nList; ++li)
    {
        const uint32_t bucket = CLASS == 0 ? list0[li] : list1[li];
        const uint32_t q0 = cnt->keyStart[bucket], q1 = cnt->keyStart[bucket + 1];
        for (uint32_t q = q0 + blockIdx.x * blockDim.x + threadIdx.x; q < q1; q += gridDim.x * blockDim.x)
        {
            const uint2 pr = sorted[q];
            Man m; m.cnt = 0;
            bool hit = false;
            if (CLASS == 0 && bucket == 0)
            {   // sphere - sphere, the bulk of a pile: only the first 16 bytes of each 64-byte collider record are needed
                ColRec A, B; A.r0 = wCol[4 * pr.x]; B.r0 = wCol[4 * pr.y];
                hit = sphereSphere(asSphere(A), asSphere(B), m, nc);
                storeManifold(manTmp, manScan, q, maxPairs, hit, m);
                continue;
            }
            const ColRec A = loadCol(wCol, pr.x), B = loadCol(wCol, pr.y);
            if (CLASS == 0)
            {
                switch (bucket)
                {
                case 0: hit = sphereSphere(asSphere(A), asSphere(B), m, nc); break;
                case 1: hit = sphereCapsule(asSphere(A), asCap(B), m, nc); break;
                case 2: hit = sphereCylinder(asSphere(A), asCap(B), m, nc); break;
                case 3: hit = sphereBox(asSphere(A), asBox(B), m); break;
                case 4: hit = sphereObb(asSphere(A), asObb(B), m); break;
                case 6: hit = capsuleVsSegmentShape(asCap(A), asCap(B), false, m, nc); break;
                case 7: hit = capsuleVsSegmentShape(asCap(A), asCap(B), true, m, nc); break;
                default: hit = boxBox(asBox(A), asBox(B), m); break;
                }
            }
            else
            {
                if (bucket == 16) { const BoxS bx = asBox(A); ObbS o; o.q = qIdentity(); o.c = boxCenter(bx); o.r = boxRadius(bx); hit = obbObb(o, asObb(B), m); } // :1142-1148
                else hit = obbObb(asObb(A), asObb(B), m);
            }
            storeManifold(manTmp, manScan, q, maxPairs, hit, m);
        }
    }
    flushCounters(nc, cnt);
}

// GJK/EPA-based buckets (5,8,9,10,11,12,13,14,17,19,20). gjkPrep = the routine up to gjkIntersectionTest: 0 no contact, 1 manifold
// complete (cylinder-cylinder with parallel axes), 2 job ready.
B200_D int gjkPrep(uint32_t bucket, const ColRec& A, const ColRec& B, const float4* __restrict__ hullVerts, const uint32_t* __restrict__ hullOffsets, Man& m, NarrowCounters& nc, GjkJob& job)
{
    switch (bucket)
    {
    case 8:  segmentShapeBoxPre(asCap(A), false, asBox(B), job); return 2;
    case 9:  segmentShapeObbPre(asCap(A), false, asObb(B), job); return 2;
    case 11: return cylinderCylinderPre(asCap(A), asCap(B), m, nc, job);
    case 12: segmentShapeBoxPre(asCap(A), true, asBox(B), job); return 2;
    case 13: segmentShapeObbPre(asCap(A), true, asObb(B), job); return 2;
    default:
    {   // anything vs hull: 5 sphere, 10 capsule, 14 cylinder, 17 aabb, 19 obb, 20 hull (:496-520,792-818,1045-1071,1150-1176,1529-1584)
        const uint32_t ta = __float_as_uint(A.r3.x) & 0xFFu;
        job.A = supportFor(ta, A, hullVerts, hullOffsets); job.B = supHull(asHull(B, hullVerts, hullOffsets)); job.tail = 0;
        return 2;
    }
    }
}
// the whole routine for one pair in the phase NC stands for (one thread, or one warp in phase 3)
template <class NC>
B200_D bool collideGjkBucket(uint32_t bucket, const ColRec& A, const ColRec& B, const float4* __restrict__ hullVerts, const uint32_t* __restrict__ hullOffsets, Man& m, NC& nc)
{
    GjkJob job;
    const int st = gjkPrep(bucket, A, B, hullVerts, hullOffsets, m, nc, job);
    if (st != 2) return st == 1;
    if (!gjkEpa(job.A, job.B, m, nc)) return false;
    gjkTail(job, m);
    return true;
}

// Work distribution of the lane-refill kernels below: a warp owns a private range [next, end) of a global cursor and takes a new
// chunk when it runs dry; idle lanes take consecutive items of the range. Returns this lane's item or 0xFFFFFFFF (none this pass).
struct WarpRange
{
    uint32_t next = 0, end = 0; bool dry = false;       // warp-uniform
    B200_D uint32_t take(bool want, uint32_t* cursor, uint32_t total, uint32_t chunk)
    {
        const uint32_t lane = threadIdx.x & 31u;
        const uint32_t mask = __ballot_sync(0xffffffffu, want);
        if (!mask) return 0xFFFFFFFFu;
        if (next >= end && !dry)
        {
            uint32_t base = 0;
            if (lane == 0u) base = atomicAdd(cursor, chunk);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (base >= total) { dry = true; next = end = total; }
            else { next = base; end = min(base + chunk, total); }
        }
        const uint32_t rank = (uint32_t)__popc(mask & ((1u << lane) - 1u));
        const uint32_t avail = end - next, n = min((uint32_t)__popc(mask), avail);
        const uint32_t item = (want && rank < avail) ? next + rank : 0xFFFFFFFFu;
        next += n;
        return item;
    }
};

// phase 1: GJK (collision_gjk.h:184-238) for every pair of the GJK buckets. Misses (and the closed-form exits of cylinder-cylinder) are
// final; hits are parked with their terminal simplex in the EPA work list. Most pairs leave after one or two support evaluations and
// the rest take 3-10 rounds, so the lanes of a warp are NOT tied to one pair each for a whole test: the loop below performs one
// support evaluation per pass for every lane and a lane whose pair is decided takes the next pair of the bucket at once.
__global__ void __launch_bounds__(128, 4) npGjkKernel(const uint2* __restrict__ sorted, const float4* __restrict__ wCol, const float4* __restrict__ hullVerts,
    const uint32_t* __restrict__ hullOffsets, float4* __restrict__ manTmp, unsigned long long* __restrict__ manScan, uint32_t maxPairs,
    float4* __restrict__ sxPlanes, uint2* __restrict__ epaWork, uint32_t sxCap, DeviceCounts* cnt, uint32_t refillAt)
{
    constexpr uint32_t list2[11] = { 5, 8, 9, 10, 11, 12, 13, 14, 17, 19, 20 };
    NarrowCounters nc; nc.epaOverflow = nc.gjkCap = nc.asserts = 0;
    const uint32_t lane = threadIdx.x & 31u;
    // the 11 GJK buckets as one index space (a warp that drains bucket by bucket idles most of its lanes 11 times): lane l < 11 holds
    // the first pair slot of bucket l and the running total in front of it
    uint32_t myStart = 0, mySize = 0;
    if (lane < 11u) { myStart = cnt->keyStart[list2[lane]]; mySize = cnt->keyStart[list2[lane] + 1] - myStart; }
    uint32_t myPre = mySize;
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, myPre, o); if (lane >= (uint32_t)o) myPre += v; }
    const uint32_t total = __shfl_sync(0xffffffffu, myPre, 10);
    myPre -= mySize;                                         // exclusive
    WarpRange range;
    bool busy = false; uint32_t q = 0, bucket = 0; int stage = 0, round = 0;
    GjkJob job; SimplexD s; f3 dir = mk3(0.f);
    for (;;)
    {
        // idle lanes take new pairs together (the set-up below runs with at least refillAt lanes, or when nothing else is left to do)
        const uint32_t idle = __ballot_sync(0xffffffffu, !busy);
        const bool batch = (uint32_t)__popc(idle) >= refillAt || idle == 0xffffffffu;
        const uint32_t item = range.take(!busy && batch, &cnt->gjkCursor[0], total, 256u);
        uint32_t li = 0;
#pragma unroll
        for (int k = 1; k < 11; ++k) { const uint32_t pk = __shfl_sync(0xffffffffu, myPre, k); if (item != 0xFFFFFFFFu && item >= pk) li = (uint32_t)k; }
        const uint32_t bStart = __shfl_sync(0xffffffffu, myStart, (int)li), bPre = __shfl_sync(0xffffffffu, myPre, (int)li);
        if (item != 0xFFFFFFFFu)
        {
            bucket = list2[li];
            q = bStart + (item - bPre);
            const uint2 pr = sorted[q];
            const ColRec A = loadCol(wCol, pr.x), B = loadCol(wCol, pr.y);
            Man m; m.cnt = 0;
            const int st = gjkPrep(bucket, A, B, hullVerts, hullOffsets, m, nc, job);
            if (st != 2) storeManifold(manTmp, manScan, q, maxPairs, st == 1, m);
            else { busy = true; stage = 0; dir = mk3(1.f, 0.1f, -0.2f); }        // collision_gjk.h:190
        }
        const uint32_t busyMask = __ballot_sync(0xffffffffu, busy);
        if (!busyMask) { if (range.dry) break; else continue; }
        if (busy)
        {
            bool done = false, hit = false;
            if (stage == 2)
            {
                if (round >= B200_GJK_MAX_ROUNDS) { ++nc.gjkCap; done = true; }
                else if (sqLen(dir) < 0.0001f) done = true;
            }
            // lanes arrive here over different paths: bring them back together for the expensive part (without this the support
            // evaluation below was issued ~twice per pass for two halves of the warp, ncu: 15 of 32 lanes active)
            const uint32_t evalMask = __ballot_sync(busyMask, !done);
            if (!done)
            {
                __syncwarp(evalMask);
                const SupPt a = minkowski(job.A, job.B, dir);
                if (dot(a.m, dir) < 0.f) done = true;
                else if (stage == 0) { s.c = a; dir = -a.m; stage = 1; }
                else if (stage == 1) { s.b = a; dir = crossABA(s.c.m - s.b.m, -s.b.m); s.n = 2; stage = 2; round = 0; }
                else
                {
                    const int r = updateSimplex(s, a, dir);
                    if (r == B200_GJK_STOP) { s.a = a; s.n = 4; done = hit = true; }
                    else if (r == B200_GJK_ERROR) done = true;
                    else ++round;
                }
            }
            if (done)
            {
                if (hit)
                {   // warp-aggregated append to the EPA work list
                    const uint32_t mask = __activemask();
                    const uint32_t leader = (uint32_t)__ffs((int)mask) - 1u;
                    uint32_t base = 0;
                    if (lane == leader) base = atomicAdd(&cnt->numEpaWork, (uint32_t)__popc(mask));
                    base = __shfl_sync(mask, base, (int)leader);
                    const uint32_t e = base + (uint32_t)__popc(mask & ((1u << lane) - 1u));
                    if (e < sxCap) { storeSimplex(sxPlanes, sxCap, e, s); epaWork[e] = make_uint2(q, bucket); }
                    else cnt->overflowManifolds = 1;
                }
                manScan[q] = 0ull;            // no manifold (yet)
                busy = false;
            }
        }
    }
    flushCounters(nc, cnt);
}

// phase 2: EPA (and the routine's tail: clipping, transform back) for the parked pairs in the large thread-local polytope.
// redo == nullptr: the whole work list; else only the entries the small-polytope kernel handed back
__global__ void __launch_bounds__(128) npEpaKernel(const uint2* __restrict__ sorted, const float4* __restrict__ wCol, const float4* __restrict__ hullVerts,
    const uint32_t* __restrict__ hullOffsets, float4* __restrict__ manTmp, unsigned long long* __restrict__ manScan, uint32_t maxPairs,
    float4* __restrict__ sxPlanes, const uint2* __restrict__ epaWork, uint32_t sxCap, DeviceCounts* cnt, const uint32_t* __restrict__ redo)
{
    NarrowCtx<2> nc; nc.epaOverflow = nc.gjkCap = nc.asserts = 0; nc.sxPlanes = sxPlanes; nc.sxCap = sxCap; nc.sxCounter = nullptr; nc.deferred = false;
    const uint32_t n = redo ? min(cnt->numEpaRedo, sxCap) : min(cnt->numEpaWork, sxCap);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const uint32_t e = redo ? redo[i] : i;
        const uint2 w = epaWork[e];
        const uint2 pr = sorted[w.x];
        const ColRec A = loadCol(wCol, pr.x), B = loadCol(wCol, pr.y);
        Man m; m.cnt = 0;
        nc.sxEntry = e;
        const bool hit = collideGjkBucket(w.y, A, B, hullVerts, hullOffsets, m, nc);
        storeManifold(manTmp, manScan, w.x, maxPairs, hit, m);
    }
    flushCounters(nc, cnt);
}

// phase 2, small polytope in shared memory (k_narrow_shapes.cuh, EpaSmall): thread per pair, 7 warps per SM. Runs take 0..20
// expansions (most 1-4), so like the GJK kernel the loop performs ONE expansion per pass for every lane and a finished lane takes the
// next work list entry at once. The few pairs that outgrow the small polytope go to the redo list (large polytope kernels).
// POLY = EpaSmall (7 warps per SM, every parked pair) or EpaMedium (3 warps, the entries of `inList` = what the small one handed back);
// entries that outgrow POLY are appended to outList.
static const int kEpaSmallWarps = 7, kEpaMediumWarps = 3;      // 7 x 242 words x 128 B = 212 KB of the 227 KB a CTA may have
template <class POLY, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, 1) npEpaSmKernel(const uint2* __restrict__ sorted, const float4* __restrict__ wCol, const float4* __restrict__ hullVerts,
    const uint32_t* __restrict__ hullOffsets, float4* __restrict__ manTmp, unsigned long long* __restrict__ manScan, uint32_t maxPairs,
    float4* __restrict__ sxPlanes, const uint2* __restrict__ epaWork, uint32_t sxCap, DeviceCounts* cnt,
    const uint32_t* __restrict__ inList, const uint32_t* __restrict__ inCount, uint32_t* cursor, uint32_t* __restrict__ outList, uint32_t* outCount, uint32_t refillAt)
{
    extern __shared__ uint32_t epaColumns[];
    const uint32_t lane = threadIdx.x & 31u;
    NarrowCounters nc; nc.epaOverflow = nc.gjkCap = nc.asserts = 0;
    POLY P; P.w = epaColumns + (size_t)(threadIdx.x >> 5) * (POLY::kWords * 32) + lane; P.active = 0; P.nT = P.nP = P.nE = 0;
    const uint32_t total = min(*inCount, sxCap);
    WarpRange range;
    // lane states: 0 idle, 1 expanding, 2 converged (contact still to be written). Finishing and starting a run cost about as much
    // as an expansion and a run is ~2.5 expansions, so both are done by many lanes together: when refillAt lanes are not expanding.
    int state = 0; uint32_t e = 0, q = 0, it = 0, ci = 0; GjkJob job;
    for (;;)
    {
        const uint32_t waiting = __ballot_sync(0xffffffffu, state != 1);
        const bool batch = (uint32_t)__popc(waiting) >= refillAt || waiting == 0xffffffffu;
        if (batch && state == 2)
        {
            Man m; f3 p;
            P.result(ci, p, m.n, m.d[0]); m.p[0] = p; m.cnt = 1;
            gjkTail(job, m);
            storeManifold(manTmp, manScan, q, maxPairs, true, m);
            state = 0;
        }
        const uint32_t item = range.take(batch && state == 0, cursor, total, 64u);
        if (item != 0xFFFFFFFFu)
        {
            e = inList ? inList[item] : item;
            const uint2 w = epaWork[e];
            q = w.x;
            const uint2 pr = sorted[q];
            const ColRec A = loadCol(wCol, pr.x), B = loadCol(wCol, pr.y);
            Man m; m.cnt = 0;
            if (gjkPrep(w.y, A, B, hullVerts, hullOffsets, m, nc, job) == 2)      // always: the pair was parked by the GJK kernel
            {
                SimplexD g; loadSimplex(sxPlanes, sxCap, e, g);
                P.init(g);
                state = 1; it = 0;
            }
        }
        if (!__any_sync(0xffffffffu, state != 0)) { if (range.dry) break; else continue; }
        const uint32_t expandMask = __ballot_sync(0xffffffffu, state == 1);
        if (state == 1)
        {   // one pass of the loop of collision_epa.h:118-146
            ci = P.closest();
            const uint32_t evalMask = __ballot_sync(expandMask, ci != 0xFFFFFFFFu);      // lanes leave closest() at different times: together again
            if (ci == 0xFFFFFFFFu) { ++nc.asserts; ci = 0; state = 2; }
            else
            {
                __syncwarp(evalMask);
                const uint32_t k = POLY::triAt(ci);
                const f3 tn = mk3(P.ldf(k), P.ldf(k + 1), P.ldf(k + 2));
                const SupPt a = minkowski(job.A, job.B, tn);
                if (dot(a.m, tn) - P.ldf(k + 3) < 0.01f) state = 2;
                else if (!P.addPoint(a)) { outList[atomicAdd(outCount, 1u)] = e; state = 0; }
                else if (++it == 20u) state = 2;
            }
        }
    }
    flushCounters(nc, cnt);
}

// phase 2, warp-cooperative: one warp per parked pair, polytope in shared memory (k_narrow_shapes.cuh, EpaWarp). All lanes run the
// routine's scalar parts uniformly (same loads: broadcast), the EPA loops are split across the lanes; lane 0 writes the manifold.
static const int kEpaWarpsPerCta = 4;
__global__ void __launch_bounds__(kEpaWarpsPerCta * 32) npEpaWarpKernel(const uint2* __restrict__ sorted, const float4* __restrict__ wCol, const float4* __restrict__ hullVerts,
    const uint32_t* __restrict__ hullOffsets, float4* __restrict__ manTmp, unsigned long long* __restrict__ manScan, uint32_t maxPairs,
    float4* __restrict__ sxPlanes, const uint2* __restrict__ epaWork, uint32_t sxCap, DeviceCounts* cnt, const uint32_t* __restrict__ redo, const uint32_t* __restrict__ redoCount, uint32_t* redoHead)
{
    extern __shared__ __align__(16) unsigned char epaSmem[];
    EpaWarpShared* sh = reinterpret_cast<EpaWarpShared*>(epaSmem) + (threadIdx.x >> 5);
    NarrowCtx<3> nc; nc.epaOverflow = nc.gjkCap = nc.asserts = 0; nc.sxPlanes = sxPlanes; nc.sxCap = sxCap; nc.sxCounter = nullptr; nc.deferred = false; nc.epaShared = sh;
    // redo == nullptr: the whole work list; else only the entries the small-polytope kernel handed back
    const uint32_t n = redo ? min(*redoCount, sxCap) : min(cnt->numEpaWork, sxCap);
    for (;;)
    {   // pairs take 1..20 expansions: hand the work list out dynamically, one entry per warp at a time
        uint32_t e = 0;
        if ((threadIdx.x & 31u) == 0u) e = atomicAdd(redo ? redoHead : &cnt->epaWorkHead, 1u);
        e = __shfl_sync(0xffffffffu, e, 0);
        if (e >= n) break;
        if (redo) e = redo[e];
        const uint2 w = epaWork[e];
        const uint2 pr = sorted[w.x];
        const ColRec A = loadCol(wCol, pr.x), B = loadCol(wCol, pr.y);
        Man m; m.cnt = 0;
        nc.sxEntry = e;
        const bool hit = collideGjkBucket(w.y, A, B, hullVerts, hullOffsets, m, nc);
        if ((threadIdx.x & 31u) == 0u) storeManifold(manTmp, manScan, w.x, maxPairs, hit, m);
    }
    if ((threadIdx.x & 31u) == 0u) flushCounters(nc, cnt);
}

// ------------------------------------------------------------------------------ non-collision interactions (event layer)
// overlapCheck (collision_narrow.cpp:1593-1689): boolean shape test for rigid body vs trigger / force field pairs
B200_D bool overlapCheckD(const ColRec& A, const ColRec& B, const float4* __restrict__ hullVerts, const uint32_t* __restrict__ hullOffsets, NarrowCounters& nc)
{
    const uint32_t ta = __float_as_uint(A.r3.x) & 0xFFu, tb = __float_as_uint(B.r3.x) & 0xFFu;
    switch (ta * 6u + tb)
    {
    case B200PHYS_SPHERE * 6 + B200PHYS_SPHERE:     return sphereVsSphereB(asSphere(A), asSphere(B));
    case B200PHYS_SPHERE * 6 + B200PHYS_CAPSULE:    return sphereVsCapsuleB(asSphere(A), asCap(B));
    case B200PHYS_SPHERE * 6 + B200PHYS_CYLINDER:   return sphereVsCylinderB(asSphere(A), asCap(B));
    case B200PHYS_SPHERE * 6 + B200PHYS_AABB:       return sphereVsBoxB(asSphere(A), asBox(B));
    case B200PHYS_SPHERE * 6 + B200PHYS_OBB:        return sphereVsObbB(asSphere(A), asObb(B));
    case B200PHYS_CAPSULE * 6 + B200PHYS_CAPSULE:   return capsuleVsCapsuleB(asCap(A), asCap(B));
    case B200PHYS_CAPSULE * 6 + B200PHYS_CYLINDER:  return capsuleVsCylinderB(asCap(A), asCap(B));
    case B200PHYS_CAPSULE * 6 + B200PHYS_AABB:      return gjkTest(supCapsule(asCap(A)), supBox(asBox(B)), nc);
    case B200PHYS_CAPSULE * 6 + B200PHYS_OBB:       { const ObbS o = asObb(B); return gjkTest(supCapsule(intoObbFrame(asCap(A), o)), supBox(boxOfObb(o)), nc); }
    case B200PHYS_CYLINDER * 6 + B200PHYS_CYLINDER: return gjkTest(supCylinder(asCap(A)), supCylinder(asCap(B)), nc);
    case B200PHYS_CYLINDER * 6 + B200PHYS_AABB:     return gjkTest(supCylinder(asCap(A)), supBox(asBox(B)), nc);
    case B200PHYS_CYLINDER * 6 + B200PHYS_OBB:      { const ObbS o = asObb(B); return gjkTest(supCylinder(intoObbFrame(asCap(A), o)), supBox(boxOfObb(o)), nc); }
    case B200PHYS_AABB * 6 + B200PHYS_AABB:         return boxVsBoxB(asBox(A), asBox(B));
    case B200PHYS_AABB * 6 + B200PHYS_OBB:          { const BoxS bx = asBox(A); ObbS o; o.q = qIdentity(); o.c = boxCenter(bx); o.r = boxRadius(bx); return obbVsObbB(o, asObb(B)); }
    case B200PHYS_OBB * 6 + B200PHYS_OBB:           return obbVsObbB(asObb(A), asObb(B));
    default:                                        // anything vs hull (types are ordered: B is the hull)
        if (tb == B200PHYS_HULL && ta <= B200PHYS_HULL) return gjkTest(supportFor(ta, A, hullVerts, hullOffsets), supHull(asHull(B, hullVerts, hullOffsets)), nc);
        ++nc.asserts; return false;
    }
}

__global__ void __launch_bounds__(128) npOverlapKernel(const uint2* __restrict__ sorted, const float4* __restrict__ wCol, const float4* __restrict__ hullVerts,
    const uint32_t* __restrict__ hullOffsets, uint32_t* __restrict__ flags, DeviceCounts* cnt)
{
    const uint32_t q0 = cnt->keyStart[kBucketOverlap], q1 = cnt->keyStart[kBucketOverlap + 1];
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt->numOverlapPairs = q1 - q0;
    NarrowCounters nc; nc.epaOverflow = nc.gjkCap = nc.asserts = 0;
    for (uint32_t q = q0 + blockIdx.x * blockDim.x + threadIdx.x; q < q1; q += gridDim.x * blockDim.x)
    {
        const uint2 pr = sorted[q];
        flags[q - q0] = overlapCheckD(loadCol(wCol, pr.x), loadCol(wCol, pr.y), hullVerts, hullOffsets, nc) ? 1u : 0u;
    }
    flushCounters(nc, cnt);
}

// compacts the overlapping pairs into non_collision_interaction records (collision_narrow.h:9-14) and applies the localized force
// fields (physics.cpp:960-967). Two force fields on one body commute; three or more are summed in an unspecified order.
__global__ void __launch_bounds__(kThreads) npInteractionKernel(const uint2* __restrict__ sorted, const float4* __restrict__ wCol, const uint32_t* __restrict__ flags,
    const uint32_t* __restrict__ slots, uint4* __restrict__ interactions, uint32_t maxInteractions, const float4* __restrict__ forceFields, uint32_t numForceFields,
    float4* __restrict__ bForce, uint32_t numBodies, DeviceCounts* cnt)
{
    const uint32_t q0 = cnt->keyStart[kBucketOverlap], n = cnt->numOverlapPairs;
    if (blockIdx.x == 0 && threadIdx.x == 0) { const uint32_t t = slots[n]; cnt->numInteractions = min(t, maxInteractions); if (t > maxInteractions) cnt->overflowPairs = 1; }
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads)
    {
        if (!flags[i]) continue;
        const uint2 pr = sorted[q0 + i];
        const float4 a3 = wCol[4 * pr.x + 3], b3 = wCol[4 * pr.y + 3];
        const uint32_t oa = (__float_as_uint(a3.x) >> 8) & 0xFFu, ob = (__float_as_uint(b3.x) >> 8) & 0xFFu;
        uint32_t rb, other, otherType;
        if (oa == B200PHYS_OBJ_RIGID_BODY) { rb = __float_as_uint(a3.y); other = __float_as_uint(b3.y); otherType = ob; }
        else { rb = __float_as_uint(b3.y); other = __float_as_uint(a3.y); otherType = oa; }
        const uint32_t s = slots[i];
        if (s < maxInteractions) interactions[s] = make_uint4(rb, other, otherType, 0u);
        if (otherType == B200PHYS_OBJ_FORCE_FIELD && other < numForceFields && rb < numBodies)
        {
            const float4 f = forceFields[other];
            atomicAdd(&bForce[rb].x, f.x); atomicAdd(&bForce[rb].y, f.y); atomicAdd(&bForce[rb].z, f.z);
        }
    }
}

__global__ void npTotalsKernel(const unsigned long long* __restrict__ scanOut, uint32_t maxManifolds, uint32_t maxContacts, DeviceCounts* cnt)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
    {
        const unsigned long long t = scanOut[cnt->keyStart[kBucketDropped]];
        const uint32_t nm = (uint32_t)(t & 0xFFFFFFFFull), nk = (uint32_t)(t >> 32);
        cnt->numManifoldsWanted = nm;
        cnt->numManifolds = min(nm, maxManifolds);
        cnt->numContacts = min(nk, maxContacts);
        if (nm > maxManifolds || nk > maxContacts) cnt->overflowManifolds = 1;
    }
}

__global__ void __launch_bounds__(kThreads) npCompactKernel(const uint2* __restrict__ sorted, const float4* __restrict__ wCol, const float4* __restrict__ manTmp,
    const unsigned long long* __restrict__ manScan, const unsigned long long* __restrict__ scanOut, uint32_t maxPairs, uint32_t maxManifolds, uint32_t maxContacts,
    uint2* __restrict__ collisions, uint4* __restrict__ manifolds, float4* __restrict__ contacts, uint2* __restrict__ contactBodies, DeviceCounts* __restrict__ cnt)
{
    const uint32_t n = cnt->keyStart[kBucketDropped];   // collision buckets come first
    for (uint32_t q = blockIdx.x * kThreads + threadIdx.x; q < n; q += gridDim.x * kThreads)
    {
        const unsigned long long self = manScan[q];
        const uint32_t count = (uint32_t)(self >> 32);
        if (!count) continue;
        const unsigned long long ex = scanOut[q];
        const uint32_t mi = (uint32_t)(ex & 0xFFFFFFFFull), cs = (uint32_t)(ex >> 32);
        if (mi >= maxManifolds || cs + count > maxContacts)
        {   // out of capacity: manifold and contact prefix sums are monotone, so what fits is a PREFIX of the canonical list; the first
            // manifold that does not fit cuts both counts (npTotalsKernel clamped them to the capacities before this kernel)
            atomicMin(&cnt->numManifolds, mi); atomicMin(&cnt->numContacts, cs);
            continue;
        }
        const uint2 pr = sorted[q];
        const float4 a2 = wCol[4 * pr.x + 2], a3 = wCol[4 * pr.x + 3], b2 = wCol[4 * pr.y + 2], b3 = wCol[4 * pr.y + 3];
        // writeScalarContact (collision_narrow.cpp:2228-2237): truncating 16+16 bit packing
        const float friction = clamp01(sqrtf(a2.w * b2.w));
        const float restitution = clamp01(fmax2(a2.z, b2.z));
        const uint32_t fr = ((uint32_t)(friction * 0xFFFF) << 16) | (uint32_t)(restitution * 0xFFFF);
        const uint32_t rbA = __float_as_uint(a3.y), rbB = __float_as_uint(b3.y);
        collisions[mi] = pr;
        manifolds[mi] = make_uint4(rbA, rbB, cs, count);
        const float4 nrm = manTmp[q];
        for (uint32_t k = 0; k < count; ++k)
        {
            contacts[2 * (cs + k)] = manTmp[(size_t)(k + 1) * maxPairs + q];
            contacts[2 * (cs + k) + 1] = make_float4(nrm.x, nrm.y, nrm.z, __uint_as_float(fr));
            contactBodies[cs + k] = make_uint2(rbA, rbB);
        }
    }
}

void launchNarrowphase(Context* c)
{
    DeviceCounts* cnt = c->counts;
    const int G = gridFor(c, 8);
    static int sphereCull = -1;
    if (sphereCull < 0) { const char* e = getenv("B200PHYS_SPHERE_CULL"); sphereCull = (e && e[0] == '0') ? 0 : 1; }       // tuning / A-B override
    launchK(c, npClassifyKernel, G, kThreads, 0, c->pairs, c->wCol, sphereCull ? c->bSphere : (const float4*)nullptr, c->hullAabb, c->pairKey, c->pairsTyped, cnt);
    c->launches += 1;
    stableSmallKeySort(c, c->pairKey, &cnt->numPairs, c->maxPairs, kNumPairKeys, c->pairDst, cnt->keyStart);
    launchK(c, npPermuteKernel, G, kThreads, 0, c->pairsTyped, c->pairDst, c->pairsSorted, cnt);
    timelineMark(c, "np_classify_sort");
    launchK(c, npCollideKernel<0>, G, 256, 0, c->pairsSorted, c->wCol, c->manTmp, c->manScan, c->maxPairs, cnt);
    launchK(c, npCollideKernel<1>, gridFor(c, 4), 128, 0, c->pairsSorted, c->wCol, c->manTmp, c->manScan, c->maxPairs, cnt);
    timelineMark(c, "np_closed_form_sat");
    static int gjkRefill = -1, epaRefill = -1;     // lanes that must be idle before a warp sets up new pairs (tuning: B200PHYS_GJK_REFILL / B200PHYS_EPA_REFILL)
    if (gjkRefill < 0) { const char* e = getenv("B200PHYS_GJK_REFILL"); gjkRefill = e ? atoi(e) : 16; e = getenv("B200PHYS_EPA_REFILL"); epaRefill = e ? atoi(e) : 16; }
    launchK(c, npGjkKernel, gridFor(c, 8), 128, 0, c->pairsSorted, c->wCol, c->hullVerts, c->hullOffsets, c->manTmp, c->manScan, c->maxPairs, c->epaSimplex, c->epaWork, c->maxEpaWork, cnt, (uint32_t)gjkRefill);
    timelineMark(c, "np_gjk");
    static int epaMode = -1, epaBps = 0;
    {
        // default: small shared-memory polytope + redo of the few that outgrow it. B200PHYS_EPA=thread: every pair in the thread-local
        // kernel (the round-1 path); =warp: the warp-cooperative kernel (measured slower). B200PHYS_EPA_BPS = CTAs per SM (tuning)
        if (epaMode < 0) { const char* e = getenv("B200PHYS_EPA"); epaMode = !e ? 2 : e[0] == 'w' ? 1 : e[0] == 't' ? 0 : 2; const char* b = getenv("B200PHYS_EPA_BPS"); epaBps = b ? atoi(b) : 0; }
        if (epaMode == 2)
        {
            const size_t smem = (size_t)kEpaSmallWarps * EpaSmall::kWords * 32 * sizeof(uint32_t), msmem = (size_t)kEpaMediumWarps * EpaMedium::kWords * 32 * sizeof(uint32_t);
            static bool attr = false;
            if (!attr)
            {
                cudaFuncSetAttribute(npEpaSmKernel<EpaSmall, kEpaSmallWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                cudaFuncSetAttribute(npEpaSmKernel<EpaMedium, kEpaMediumWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)msmem);
                attr = true;
            }
            launchK(c, npEpaSmKernel<EpaSmall, kEpaSmallWarps>, c->numSMs, kEpaSmallWarps * 32, smem, c->pairsSorted, c->wCol, c->hullVerts, c->hullOffsets, c->manTmp, c->manScan, c->maxPairs, c->epaSimplex,
                    c->epaWork, c->maxEpaWork, cnt, nullptr, &cnt->numEpaWork, &cnt->epaWorkHead, c->epaRedo, &cnt->numEpaRedo, (uint32_t)epaRefill);
            timelineMark(c, "np_epa");
            // the ones that outgrew it are the long runs (up to 20 expansions, ~100 triangles): a warp per pair (exact, see EpaWarp).
            // B200PHYS_EPA_REDO=medium puts a 16-point / 62-triangle shared-memory tier (3 warps per SM) in between -- measured: 1.28 -> 1.07 ms
            // while the pile is still falling (74 k redo entries) but 0.12 -> 0.32 ms once it has settled (7 k entries: too few threads, each
            // with a long run), so it is off by default; =thread uses the thread-local kernel for the redo entries (0.43 ms settled).
            static int redoMode = -1;
            if (redoMode < 0) { const char* e = getenv("B200PHYS_EPA_REDO"); redoMode = !e ? 2 : e[0] == 't' ? 1 : e[0] == 'm' ? 0 : 2; }
            const uint32_t* lastList = c->epaRedo; uint32_t* lastCount = &cnt->numEpaRedo; uint32_t* lastHead = &cnt->epaRedoHead;
            if (redoMode == 0)
            {
                launchK(c, npEpaSmKernel<EpaMedium, kEpaMediumWarps>, c->numSMs, kEpaMediumWarps * 32, msmem, c->pairsSorted, c->wCol, c->hullVerts, c->hullOffsets, c->manTmp, c->manScan, c->maxPairs, c->epaSimplex,
                        c->epaWork, c->maxEpaWork, cnt, c->epaRedo, &cnt->numEpaRedo, &cnt->epaRedoHead, c->epaRedo2, &cnt->numEpaRedo2, (uint32_t)epaRefill);
                c->launches += 1;
                lastList = c->epaRedo2; lastCount = &cnt->numEpaRedo2; lastHead = &cnt->epaRedo2Head;
            }
            if (redoMode == 1)
                launchK(c, npEpaKernel, gridFor(c, 2), 64, 0, c->pairsSorted, c->wCol, c->hullVerts, c->hullOffsets, c->manTmp, c->manScan, c->maxPairs, c->epaSimplex, c->epaWork, c->maxEpaWork, cnt, c->epaRedo);
            else
            {
                const size_t wsmem = sizeof(EpaWarpShared) * kEpaWarpsPerCta;
                static bool wattr = false;
                if (!wattr) { cudaFuncSetAttribute(npEpaWarpKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)wsmem); wattr = true; }
                launchK(c, npEpaWarpKernel, gridFor(c, 6), kEpaWarpsPerCta * 32, wsmem, c->pairsSorted, c->wCol, c->hullVerts, c->hullOffsets, c->manTmp, c->manScan, c->maxPairs, c->epaSimplex, c->epaWork, c->maxEpaWork, cnt,
                        lastList, lastCount, lastHead);
            }
            c->launches += 1;
        }
        else if (epaMode == 0)
            launchK(c, npEpaKernel, gridFor(c, epaBps > 0 ? epaBps : 8), 128, 0, c->pairsSorted, c->wCol, c->hullVerts, c->hullOffsets, c->manTmp, c->manScan, c->maxPairs, c->epaSimplex, c->epaWork, c->maxEpaWork, cnt, nullptr);
        else
        {
            const size_t smem = sizeof(EpaWarpShared) * kEpaWarpsPerCta;
            static bool attr = false;
            if (!attr) { cudaFuncSetAttribute(npEpaWarpKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); attr = true; }
            launchK(c, npEpaWarpKernel, gridFor(c, epaBps > 0 ? epaBps : 6), kEpaWarpsPerCta * 32, smem, c->pairsSorted, c->wCol, c->hullVerts, c->hullOffsets, c->manTmp, c->manScan, c->maxPairs, c->epaSimplex, c->epaWork, c->maxEpaWork, cnt, nullptr, nullptr, nullptr);
        }
    }
    c->launches += 1;
    timelineMark(c, epaMode == 2 ? "np_epa_redo" : "np_epa");
    c->launches += 4;
    // manifold / contact offsets: only the collision buckets can hold manifolds (the pruned and overlap-only pairs are sorted behind them)
    scanExclusiveU64(c, c->manScan, c->manScanOut, &cnt->keyStart[kBucketDropped], c->maxPairs, nullptr);
    // rigid body vs trigger / force field: boolean overlap tests, compaction, force-field push (physics.cpp:1255, before the force integration)
    if (c->numNonCollisionColliders)
    {
        uint32_t* ovFlags = c->pairDst; uint32_t* ovSlots = (uint32_t*)c->pairsTyped;   // both free between the bucket sort and the solver
        launchK(c, npOverlapKernel, gridFor(c, 4), 128, 0, c->pairsSorted, c->wCol, c->hullVerts, c->hullOffsets, ovFlags, cnt);
        c->launches++;
        scanExclusiveU32(c, ovFlags, ovSlots, &cnt->numOverlapPairs, c->maxPairs, nullptr);
        launchK(c, npInteractionKernel, G, kThreads, 0, c->pairsSorted, c->wCol, ovFlags, ovSlots, c->interactions, c->maxInteractions, c->forceFields, c->numForceFields,
                                                           c->bForce, c->numBodies, cnt);
        c->launches++;
    }
    launchK(c, npTotalsKernel, 1, 32, 0, c->manScanOut, c->maxManifolds, c->maxContacts, cnt);
    launchK(c, npCompactKernel, G, kThreads, 0, c->pairsSorted, c->wCol, c->manTmp, c->manScan, c->manScanOut, c->maxPairs, c->maxManifolds, c->maxContacts,
                                                   c->collisions, c->manifolds, c->contacts, c->contactBodies, cnt);
    c->launches += 2;
    timelineMark(c, "np_compact");
}

} // namespace b200

#include "k_heightmap.cuh"
