// k_broad.cu -- uniform-grid broad phase. Replaces the reference's single-axis sweep-and-prune
// (src/physics/collision_broad.cpp:87-166,297-447) but emits the SAME SET: every collider pair whose world AABBs overlap
// under the inclusive predicate of bounding_volumes.h:352-358 (same-body and static-static pairs included; the narrow
// phase prunes them, collision_narrow.cpp:2358-2369). Output is canonical: a < b, sorted by (a, b), built in order by a
// count -> scan -> fill scheme (no atomics on the output, no global sort), so it is bit-reproducible run to run.
//
// Build:  extent statistics -> classify (grid vs "big" list) -> grid set-up on device -> bin by AABB-centre cell
//         (count, scan, scatter a cell-ordered copy of the AABBs: 2 x 128-bit per collider).
// Query:  one thread per cell-ordered collider walks the cell rows its expanded AABB can reach (each row is ONE contiguous
//         range of the cell-ordered copy), tests 128-bit AABBs, keeps partners with a larger index; big colliders are
//         tested from a shared-memory list, and each big collider sweeps all colliders with one CTA.
// Correctness does not depend on the cell size: a partner's centre lies inside the query AABB grown by half the largest
// small extent, and cell coordinates are a monotone function of position, so no rounding can hide a pair.
#include "b200_context.h"
#include "b200_util.cuh"

namespace b200 {

static const int kThreads = 256;
static const uint32_t kLocalPartners = 48;

B200_D bool overlapInclusive(float4 amn, float4 amx, float4 bmn, float4 bmx)
{   // bounding_volumes.h:352-358
    if (amx.x < bmn.x || amn.x > bmx.x) return false;
    if (amx.y < bmn.y || amn.y > bmx.y) return false;
    if (amx.z < bmn.z || amn.z > bmx.z) return false;
    return true;
}

B200_D float maxExtent(float4 mn, float4 mx) { return fmaxf(mx.x - mn.x, fmaxf(mx.y - mn.y, mx.z - mn.z)); }

// ------------------------------------------------------------------------------------------------------ statistics
__global__ void __launch_bounds__(kThreads) bpResetKernel(DeviceCounts* cnt)
{
    if (threadIdx.x == 0)
    {
        cnt->extentSumFixed = 0ull; cnt->extentCount = 0; cnt->maxSmallExtentBits = 0; cnt->numBig = 0;
        for (int a = 0; a < 3; ++a) { cnt->centreMinBits[a] = 0xFFFFFFFFu; cnt->centreMaxBits[a] = 0u; }
        cnt->numPairs = 0; cnt->numPairsWanted = 0; cnt->numEpaWork = 0; cnt->epaWorkHead = 0;
    }
}

__global__ void __launch_bounds__(kThreads) bpExtentKernel(uint32_t C, const float4* __restrict__ mn, const float4* __restrict__ mx, DeviceCounts* cnt)
{
    unsigned long long sum = 0; uint32_t n = 0;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < C; i += gridDim.x * kThreads)
    {
        float e = maxExtent(ldStream(mn + i), ldStream(mx + i));
        if (e >= 0.f && e < 1e12f) { sum += (unsigned long long)((double)e * 65536.0); ++n; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_down_sync(0xffffffffu, sum, o); n += __shfl_down_sync(0xffffffffu, n, o); }
    // one pair of atomics per CTA (same-address atomics serialise at one L2 slice)
    __shared__ unsigned long long sSum[kThreads / 32]; __shared__ uint32_t sN[kThreads / 32];
    if (laneId() == 0) { sSum[threadIdx.x >> 5] = sum; sN[threadIdx.x >> 5] = n; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        unsigned long long ts = 0; uint32_t tn = 0;
        for (int w = 0; w < kThreads / 32; ++w) { ts += sSum[w]; tn += sN[w]; }
        if (tn) { atomicAdd(&cnt->extentSumFixed, ts); atomicAdd(&cnt->extentCount, tn); }
    }
}

__global__ void __launch_bounds__(kThreads) bpClassifyKernel(uint32_t C, float bigFactor, float cfgCell, const float4* __restrict__ mn, const float4* __restrict__ mx,
                                                             uint32_t* __restrict__ cellOf, uint32_t* __restrict__ bigList, DeviceCounts* cnt)
{
    const double mean = cnt->extentCount ? ((double)cnt->extentSumFixed / 65536.0 / (double)cnt->extentCount) : 1.0;
    float threshold = (float)(mean * (double)bigFactor);
    if (cfgCell > 0.f) threshold = cfgCell;       // explicit cell size: anything larger goes to the big list
    if (blockIdx.x == 0 && threadIdx.x == 0) cnt->grid.bigThreshold = threshold;
    uint32_t extBits = 0, lo[3] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[3] = { 0u, 0u, 0u };
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < C; i += gridDim.x * kThreads)
    {
        const float4 a = ldStream(mn + i), b = ldStream(mx + i);
        const float e = maxExtent(a, b);
        if (a.x > b.x) { cellOf[i] = 0xFFFFFFFFu; continue; }   // inverted box = halo padding slot: neither binned nor big
        if (!(e <= threshold))
        {   // big (or non-finite): brute force
            const uint32_t k = atomicAdd(&cnt->numBig, 1u);
            if (k < kMaxBig) bigList[k] = i; else cnt->overflowBig = 1;
            cellOf[i] = 0xFFFFFFFFu;
        }
        else
        {
            cellOf[i] = 0;
            extBits = max(extBits, __float_as_uint(fmaxf(e, 0.f)));
            const float cx = (a.x + b.x) * 0.5f, cy = (a.y + b.y) * 0.5f, cz = (a.z + b.z) * 0.5f;
            const uint32_t sx = floatToSortable(cx), sy = floatToSortable(cy), sz = floatToSortable(cz);
            lo[0] = min(lo[0], sx); lo[1] = min(lo[1], sy); lo[2] = min(lo[2], sz);
            hi[0] = max(hi[0], sx); hi[1] = max(hi[1], sy); hi[2] = max(hi[2], sz);
        }
    }
    extBits = __reduce_max_sync(0xffffffffu, extBits);
#pragma unroll
    for (int a = 0; a < 3; ++a) { lo[a] = __reduce_min_sync(0xffffffffu, lo[a]); hi[a] = __reduce_max_sync(0xffffffffu, hi[a]); }
    __shared__ uint32_t sRed[kThreads / 32][7];
    if (laneId() == 0) { sRed[threadIdx.x >> 5][0] = extBits; for (int a = 0; a < 3; ++a) { sRed[threadIdx.x >> 5][1 + a] = lo[a]; sRed[threadIdx.x >> 5][4 + a] = hi[a]; } }
    __syncthreads();
    if (threadIdx.x == 0)
    {   // one set of atomics per CTA
        for (int w = 1; w < kThreads / 32; ++w)
        {
            extBits = max(extBits, sRed[w][0]);
            for (int a = 0; a < 3; ++a) { lo[a] = min(lo[a], sRed[w][1 + a]); hi[a] = max(hi[a], sRed[w][4 + a]); }
        }
        if (extBits) atomicMax(&cnt->maxSmallExtentBits, extBits);
        for (int a = 0; a < 3; ++a) { if (lo[a] != 0xFFFFFFFFu) atomicMin(&cnt->centreMinBits[a], lo[a]); if (hi[a]) atomicMax(&cnt->centreMaxBits[a], hi[a]); }
    }
}

// single thread: derive the grid from the statistics (all on device, no host round trip)
__global__ void bpSetupKernel(uint32_t maxCells, DeviceCounts* cnt)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    GridParams g;
    float E = __uint_as_float(cnt->maxSmallExtentBits);
    float s = E * 1.001f + 1e-6f;
    if (!(s > 1e-6f)) s = 1.f;
    float lo[3], hi[3];
    bool any = cnt->centreMinBits[0] != 0xFFFFFFFFu;
    for (int a = 0; a < 3; ++a) { lo[a] = any ? sortableToFloat(cnt->centreMinBits[a]) : 0.f; hi[a] = any ? sortableToFloat(cnt->centreMaxBits[a]) : 0.f; }
    // shrink the resolution until the dense grid fits the cell budget (a coarser grid is slower, never wrong)
    for (int it = 0; it < 64; ++it)
    {
        double total = 1.0; bool ok = true;
        for (int a = 0; a < 3; ++a)
        {
            double d = floor(((double)hi[a] - (double)lo[a]) / (double)s) + 2.0;
            if (d > 1048576.0) ok = false;
            g.dims[a] = (int32_t)fmin(d, 1048576.0);
            total *= d;
        }
        if (ok && total <= (double)maxCells) break;
        s *= 1.26f;   // ~2x fewer cells per round
    }
    g.cellSize = s; g.invCellSize = 1.f / s;
    for (int a = 0; a < 3; ++a) g.origin[a] = lo[a];
    g.numCells = (uint32_t)g.dims[0] * (uint32_t)g.dims[1] * (uint32_t)g.dims[2];
    if (g.numCells > maxCells) { cnt->overflowGrid = 1; g.dims[0] = g.dims[1] = g.dims[2] = 1; g.numCells = 1; }
    g.bigThreshold = cnt->grid.bigThreshold;
    cnt->grid = g;
}

B200_D int cellCoord(float x, float origin, float inv, int dim)
{
    // monotone non-decreasing in x (fp subtract, multiply by a positive constant, floor, clamp are all monotone)
    float f = floorf((x - origin) * inv);
    f = fminf(fmaxf(f, 0.f), (float)(dim - 1));
    return (int)f;
}

__global__ void __launch_bounds__(kThreads) bpZeroCellsKernel(uint32_t* __restrict__ cellCount, const DeviceCounts* __restrict__ cnt)
{
    const uint32_t n = cnt->grid.numCells + 1;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < n; i += gridDim.x * kThreads) cellCount[i] = 0;
}

__global__ void __launch_bounds__(kThreads) bpBinKernel(uint32_t C, const float4* __restrict__ mn, const float4* __restrict__ mx,
                                                        uint32_t* __restrict__ cellOf, uint32_t* __restrict__ rankInCell, uint32_t* __restrict__ cellCount, const DeviceCounts* __restrict__ cnt)
{
    const GridParams g = cnt->grid;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < C; i += gridDim.x * kThreads)
    {
        if (cellOf[i] == 0xFFFFFFFFu) continue;
        const float4 a = ldStream(mn + i), b = ldStream(mx + i);
        const int cx = cellCoord((a.x + b.x) * 0.5f, g.origin[0], g.invCellSize, g.dims[0]);
        const int cy = cellCoord((a.y + b.y) * 0.5f, g.origin[1], g.invCellSize, g.dims[1]);
        const int cz = cellCoord((a.z + b.z) * 0.5f, g.origin[2], g.invCellSize, g.dims[2]);
        const uint32_t cell = ((uint32_t)cz * (uint32_t)g.dims[1] + (uint32_t)cy) * (uint32_t)g.dims[0] + (uint32_t)cx;
        cellOf[i] = cell;
        rankInCell[i] = atomicAdd(&cellCount[cell], 1u);
    }
}

__global__ void __launch_bounds__(kThreads) bpScatterKernel(uint32_t C, const float4* __restrict__ mn, const float4* __restrict__ mx, const uint32_t* __restrict__ cellOf,
                                                            const uint32_t* __restrict__ rankInCell, const uint32_t* __restrict__ cellStart, float4* __restrict__ sortedMin, float4* __restrict__ sortedMax)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < C; i += gridDim.x * kThreads)
    {
        const uint32_t cell = cellOf[i];
        if (cell == 0xFFFFFFFFu) continue;
        const uint32_t pos = cellStart[cell] + rankInCell[i];
        float4 a = ldStream(mn + i), b = ldStream(mx + i);
        a.w = __uint_as_float(i);
        sortedMin[pos] = a; sortedMax[pos] = b;
    }
}

// ----------------------------------------------------------------------------------------------------------- query
// FILL == false: pairCount[id] = number of partners with a larger index.  FILL == true: write them, ascending.
template <bool FILL>
__global__ void __launch_bounds__(kThreads) bpQueryKernel(const float4* __restrict__ sortedMin, const float4* __restrict__ sortedMax, const uint32_t* __restrict__ cellStart,
    const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax, const uint32_t* __restrict__ bigList,
    uint32_t* __restrict__ pairCount, const uint32_t* __restrict__ pairStart, uint2* __restrict__ pairs, uint32_t maxPairs, DeviceCounts* cnt)
{
    __shared__ float4 sBigMin[kMaxBig / 4], sBigMax[kMaxBig / 4];   // staged in chunks of 256 big colliders
    __shared__ uint32_t sBigId[kMaxBig / 4];
    const GridParams g = cnt->grid;
    const uint32_t numBinned = cellStart[g.numCells];
    const uint32_t numBig = min(cnt->numBig, kMaxBig);
    const float halfE = 0.5f * __uint_as_float(cnt->maxSmallExtentBits);
    const uint32_t rounds = (numBinned + gridDim.x * kThreads - 1) / (gridDim.x * kThreads);
    for (uint32_t r = 0; r < rounds; ++r)
    {
        const uint32_t e = (r * gridDim.x + blockIdx.x) * kThreads + threadIdx.x;
        const bool active = e < numBinned;
        float4 amn = make_float4(0, 0, 0, 0), amx = amn; uint32_t id = 0, count = 0, base = 0, expected = 0;
        uint32_t part[kLocalPartners];   // FILL: partners are collected and sorted in thread-local memory, then written once
        if (active)
        {
            amn = sortedMin[e]; amx = sortedMax[e]; id = __float_as_uint(amn.w);
            if (FILL) { base = pairStart[id]; expected = pairStart[id + 1] - base; }
            // cells that can hold the centre of an overlapping small collider
            const float slackX = halfE + halfE * 1e-3f + fabsf(amn.x) * 1e-6f + fabsf(amx.x) * 1e-6f;
            const float slackY = halfE + halfE * 1e-3f + fabsf(amn.y) * 1e-6f + fabsf(amx.y) * 1e-6f;
            const float slackZ = halfE + halfE * 1e-3f + fabsf(amn.z) * 1e-6f + fabsf(amx.z) * 1e-6f;
            const int x0 = cellCoord(amn.x - slackX, g.origin[0], g.invCellSize, g.dims[0]), x1 = cellCoord(amx.x + slackX, g.origin[0], g.invCellSize, g.dims[0]);
            const int y0 = cellCoord(amn.y - slackY, g.origin[1], g.invCellSize, g.dims[1]), y1 = cellCoord(amx.y + slackY, g.origin[1], g.invCellSize, g.dims[1]);
            const int z0 = cellCoord(amn.z - slackZ, g.origin[2], g.invCellSize, g.dims[2]), z1 = cellCoord(amx.z + slackZ, g.origin[2], g.invCellSize, g.dims[2]);
            for (int z = z0; z <= z1; ++z)
                for (int y = y0; y <= y1; ++y)
                {
                    const uint32_t row = ((uint32_t)z * (uint32_t)g.dims[1] + (uint32_t)y) * (uint32_t)g.dims[0];
                    const uint32_t f0 = cellStart[row + x0], f1 = cellStart[row + x1 + 1];   // one contiguous candidate range per row
                    for (uint32_t f = f0; f < f1; ++f)
                    {
                        const float4 bmn = sortedMin[f];
                        const uint32_t jd = __float_as_uint(bmn.w);
                        if (jd <= id) continue;
                        if (!overlapInclusive(amn, amx, bmn, sortedMax[f])) continue;
                        if (FILL) { if (expected <= kLocalPartners) part[count] = jd; else if (base + count < maxPairs) pairs[base + count] = make_uint2(id, jd); else cnt->overflowPairs = 1; }
                        ++count;
                    }
                }
        }
        // big colliders, staged through shared memory 256 at a time
        for (uint32_t b0 = 0; b0 < numBig; b0 += kMaxBig / 4)
        {
            __syncthreads();
            const uint32_t nb = min(kMaxBig / 4, numBig - b0);
            if (threadIdx.x < nb) { const uint32_t b = bigList[b0 + threadIdx.x]; sBigId[threadIdx.x] = b; sBigMin[threadIdx.x] = aabbMin[b]; sBigMax[threadIdx.x] = aabbMax[b]; }
            __syncthreads();
            if (active)
                for (uint32_t k = 0; k < nb; ++k)
                {
                    const uint32_t jd = sBigId[k];
                    if (jd <= id) continue;
                    if (!overlapInclusive(amn, amx, sBigMin[k], sBigMax[k])) continue;
                    if (FILL) { if (expected <= kLocalPartners) part[count] = jd; else if (base + count < maxPairs) pairs[base + count] = make_uint2(id, jd); else cnt->overflowPairs = 1; }
                    ++count;
                }
        }
        if (active)
        {
            if (!FILL) pairCount[id] = count;
            else if (expected <= kLocalPartners)
            {   // ascending partner order: insertion sort in local memory, then one sequential write of the segment
                for (uint32_t a = 1; a < count; ++a)
                {
                    const uint32_t key = part[a];
                    uint32_t b = a;
                    while (b > 0 && part[b - 1] > key) { part[b] = part[b - 1]; --b; }
                    part[b] = key;
                }
                for (uint32_t a = 0; a < count; ++a) { if (base + a < maxPairs) pairs[base + a] = make_uint2(id, part[a]); else cnt->overflowPairs = 1; }
            }
            else
            {   // rare: more partners than the local buffer -> insertion sort in place in global memory
                const uint32_t n = (base + count <= maxPairs) ? count : (base < maxPairs ? maxPairs - base : 0);
                for (uint32_t a = 1; a < n; ++a)
                {
                    const uint2 key = pairs[base + a];
                    uint32_t b = a;
                    while (b > 0 && pairs[base + b - 1].y > key.y) { pairs[base + b] = pairs[base + b - 1]; --b; }
                    pairs[base + b] = key;
                }
            }
        }
    }
}

// one CTA per big collider: sweep every collider with a larger index, ordered compaction
template <bool FILL>
__global__ void __launch_bounds__(kThreads) bpBigKernel(uint32_t C, uint32_t firstSlot, const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax, const uint32_t* __restrict__ bigList,
    uint32_t* __restrict__ pairCount, const uint32_t* __restrict__ pairStart, uint2* __restrict__ pairs, uint32_t maxPairs, DeviceCounts* cnt)
{
    __shared__ uint32_t warpCnt[kThreads / 32];
    __shared__ uint32_t running;
    const uint32_t numBig = min(cnt->numBig, kMaxBig);
    for (uint32_t k = firstSlot + blockIdx.x; k < numBig; k += gridDim.x)
    {
        const uint32_t id = bigList[k];
        const float4 amn = aabbMin[id], amx = aabbMax[id];
        const uint32_t base = FILL ? pairStart[id] : 0;
        if (threadIdx.x == 0) running = 0;
        __syncthreads();
        for (uint32_t j0 = id + 1; j0 < C; j0 += kThreads)
        {
            const uint32_t j = j0 + threadIdx.x;
            const bool hit = (j < C) && overlapInclusive(amn, amx, ldStream(aabbMin + j), ldStream(aabbMax + j));
            const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
            if (laneId() == 0) warpCnt[threadIdx.x >> 5] = __popc(ballot);
            __syncthreads();
            uint32_t off = running;
            for (uint32_t w = 0; w < (threadIdx.x >> 5); ++w) off += warpCnt[w];
            off += __popc(ballot & laneMaskLt());
            if (FILL && hit) { if (base + off < maxPairs) pairs[base + off] = make_uint2(id, j); else cnt->overflowPairs = 1; }
            __syncthreads();
            if (threadIdx.x == 0) { uint32_t t = 0; for (int w = 0; w < kThreads / 32; ++w) t += warpCnt[w]; running += t; }
            __syncthreads();
        }
        if (!FILL && threadIdx.x == 0) pairCount[id] = running;
        __syncthreads();
    }
}

// Big colliders, parallel path (numBig * C <= scratch): flag matrix [slot][j] = (j > id_slot && overlap), ONE exclusive scan over
// the flattened matrix gives every pair its ordered position inside the big collider's segment.
__global__ void __launch_bounds__(kThreads) bpBigFlagKernel(uint32_t C, const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax, const uint32_t* __restrict__ bigList,
                                                            uint32_t* __restrict__ flags, const DeviceCounts* __restrict__ cnt)
{
    __shared__ float4 sMin[16], sMax[16]; __shared__ uint32_t sId[16];
    const uint32_t nb = min(cnt->numBig, 16u);
    if (threadIdx.x < nb) { const uint32_t b = bigList[threadIdx.x]; sId[threadIdx.x] = b; sMin[threadIdx.x] = aabbMin[b]; sMax[threadIdx.x] = aabbMax[b]; }
    __syncthreads();
    for (uint32_t j = blockIdx.x * kThreads + threadIdx.x; j < C; j += gridDim.x * kThreads)
    {
        const float4 mn = ldStream(aabbMin + j), mx = ldStream(aabbMax + j);
        for (uint32_t k = 0; k < nb; ++k) flags[(size_t)k * C + j] = (j > sId[k] && overlapInclusive(sMin[k], sMax[k], mn, mx)) ? 1u : 0u;
    }
}
__global__ void bpBigCountKernel(uint32_t C, const uint32_t* __restrict__ bigList, const uint32_t* __restrict__ offs, uint32_t* __restrict__ pairCount, DeviceCounts* cnt)
{
    const uint32_t nb = min(cnt->numBig, 16u);
    if (threadIdx.x < nb) pairCount[bigList[threadIdx.x]] = offs[(size_t)(threadIdx.x + 1) * C] - offs[(size_t)threadIdx.x * C];
}
__global__ void __launch_bounds__(kThreads) bpBigFillKernel(uint32_t C, const uint32_t* __restrict__ bigList, const uint32_t* __restrict__ flags, const uint32_t* __restrict__ offs,
                                                            const uint32_t* __restrict__ pairStart, uint2* __restrict__ pairs, uint32_t maxPairs, DeviceCounts* cnt)
{
    const uint32_t nb = min(cnt->numBig, 16u);
    const size_t total = (size_t)nb * C;
    for (size_t idx = (size_t)blockIdx.x * kThreads + threadIdx.x; idx < total; idx += (size_t)gridDim.x * kThreads)
    {
        if (!flags[idx]) continue;
        const uint32_t k = (uint32_t)(idx / C), j = (uint32_t)(idx - (size_t)k * C), id = bigList[k];
        const uint32_t pos = pairStart[id] + (offs[idx] - offs[(size_t)k * C]);
        if (pos < maxPairs) pairs[pos] = make_uint2(id, j); else cnt->overflowPairs = 1;
    }
}
__global__ void bpBigCountDeviceKernel(uint32_t C, DeviceCounts* cnt) { if (threadIdx.x == 0) cnt->bigMatrixCount = min(cnt->numBig, 16u) * C; }

__global__ void bpFinishKernel(uint32_t C, const uint32_t* __restrict__ pairStart, uint32_t maxPairs, DeviceCounts* cnt)
{
    if (threadIdx.x == 0 && blockIdx.x == 0)
    {
        const uint32_t total = pairStart[C];
        cnt->numPairsWanted = total;
        cnt->numPairs = min(total, maxPairs);
        if (total > maxPairs) cnt->overflowPairs = 1;
    }
}

void launchBroadphase(Context* c)
{
    const uint32_t C = c->numColliders;
    DeviceCounts* cnt = c->counts;
    const float bigFactor = c->cfg.bigFactor > 0.f ? c->cfg.bigFactor : 3.f;
    const int G = gridFor(c, 8);
    bpResetKernel<<<1, 32, 0, c->stream>>>(cnt);
    if (C == 0) { c->launches += 1; return; }
    bpExtentKernel<<<G, kThreads, 0, c->stream>>>(C, c->aabbMin, c->aabbMax, cnt);
    bpClassifyKernel<<<G, kThreads, 0, c->stream>>>(C, bigFactor, c->cfg.cellSize, c->aabbMin, c->aabbMax, c->cellOf, c->bigList, cnt);
    bpSetupKernel<<<1, 32, 0, c->stream>>>(c->maxCells, cnt);
    bpZeroCellsKernel<<<G, kThreads, 0, c->stream>>>(c->cellCount, cnt);
    bpBinKernel<<<G, kThreads, 0, c->stream>>>(C, c->aabbMin, c->aabbMax, c->cellOf, c->rankInCell, c->cellCount, cnt);
    c->launches += 6;
    scanExclusiveU32(c, c->cellCount, c->cellStart, &cnt->grid.numCells, c->maxCells, nullptr);
    bpScatterKernel<<<G, kThreads, 0, c->stream>>>(C, c->aabbMin, c->aabbMax, c->cellOf, c->rankInCell, c->cellStart, c->sortedMin, c->sortedMax);
    timelineMark(c, "bp_build");
    cudaMemsetAsync(c->pairCount, 0, sizeof(uint32_t) * (C + 1), c->stream);
    bpQueryKernel<false><<<G, kThreads, 0, c->stream>>>(c->sortedMin, c->sortedMax, c->cellStart, c->aabbMin, c->aabbMax, c->bigList, c->pairCount, c->pairStart, c->pairs, c->maxPairs, cnt);
    // big colliders: parallel flag-matrix path for up to 16 of them (scratch: pair-sized buffers that are free until the narrow
    // phase); beyond that one CTA per big collider sweeps all colliders (bpBigKernel, skips the first 16)
    uint32_t* bigFlags = c->pairDst; uint32_t* bigOffs = (uint32_t*)c->manScan;
    const bool matrixFits = (size_t)16 * C <= (size_t)c->maxPairs;   // both scratch buffers hold maxPairs + 2 words
    if (matrixFits)
    {
        bpBigCountDeviceKernel<<<1, 32, 0, c->stream>>>(C, cnt);
        bpBigFlagKernel<<<G, kThreads, 0, c->stream>>>(C, c->aabbMin, c->aabbMax, c->bigList, bigFlags, cnt);
        c->launches += 2;
        scanExclusiveU32(c, bigFlags, bigOffs, &cnt->bigMatrixCount, 16u * C, nullptr);
        bpBigCountKernel<<<1, 32, 0, c->stream>>>(C, c->bigList, bigOffs, c->pairCount, cnt);
        c->launches += 1;
    }
    bpBigKernel<false><<<c->numSMs, kThreads, 0, c->stream>>>(C, matrixFits ? 16u : 0u, c->aabbMin, c->aabbMax, c->bigList, c->pairCount, c->pairStart, c->pairs, c->maxPairs, cnt);
    c->launches += 3;
    scanExclusiveU32(c, c->pairCount, c->pairStart, nullptr, C, nullptr);
    bpQueryKernel<true><<<G, kThreads, 0, c->stream>>>(c->sortedMin, c->sortedMax, c->cellStart, c->aabbMin, c->aabbMax, c->bigList, c->pairCount, c->pairStart, c->pairs, c->maxPairs, cnt);
    if (matrixFits) { bpBigFillKernel<<<G, kThreads, 0, c->stream>>>(C, c->bigList, bigFlags, bigOffs, c->pairStart, c->pairs, c->maxPairs, cnt); c->launches += 1; }
    bpBigKernel<true><<<c->numSMs, kThreads, 0, c->stream>>>(C, matrixFits ? 16u : 0u, c->aabbMin, c->aabbMax, c->bigList, c->pairCount, c->pairStart, c->pairs, c->maxPairs, cnt);
    bpFinishKernel<<<1, 32, 0, c->stream>>>(C, c->pairStart, c->maxPairs, cnt);
    c->launches += 3;
    timelineMark(c, "bp_query");
}

} // namespace b200
