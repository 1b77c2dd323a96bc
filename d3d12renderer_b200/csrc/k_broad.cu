// k_broad.cu -- uniform-grid broad phase. Replaces the reference's single-axis sweep-and-prune
// (src/physics/collision_broad.cpp:87-166,297-447) but emits the SAME SET: every collider pair whose world AABBs overlap
// under the inclusive predicate of bounding_volumes.h:352-358 (same-body and static-static pairs included; the narrow
// phase prunes them, collision_narrow.cpp:2358-2369). Output is canonical: a < b, sorted by (a, b), and bit-reproducible
// run to run (no order-dependent atomics reach the output).
//
// Eight launches per step (round 1: ~20, two grid traversals):
//   classify : one pass over the AABBs: "big" colliders (extent above the threshold the PREVIOUS step derived from the mean extent)
//              go to a brute-force list, the rest contribute to the grid statistics (max small extent, centre bounds, mean)
//   setup    : one thread: grid resolution, reach, tile shape and the threshold for the next step
//   bin      : cell of every small collider's AABB centre, rank inside the cell
//   scan     : cell counts -> cell starts (the scan also clears the counts for the next step)
//   scatter  : cell-ordered copy of the AABBs as 32-byte records (min.xyz, id | max.xyz, -)
//   tile     : ONE traversal. A CTA owns a run of cells of one grid row; the candidate records of the neighbouring rows are contiguous
//              runs of the cell-ordered array and are staged in shared memory by TMA bulk copies (cp.async.bulk + mbarrier), so a
//              record is fetched once per tile instead of once per collider; every thread tests its collider against the staged
//              records, keeps the partners with a larger index in shared memory, sorts them and writes one 128-byte row.
//              Extra CTAs count the pairs of the big colliders chunk by chunk.
//   scan     : partners per collider -> canonical pair offsets
//   emit     : rows -> pair list with coalesced stores; rows that overflowed, colliders whose reach left the staged window and the
//              big colliders' rows are produced here by direct traversal.
// With candidate reuse (default, see "candidate reuse" below) three more launches frame the pipeline: check (are all boxes still inside the grown
// boxes of the last build?), decide (one thread: rebuild or reuse; every kernel above is gated by its flag) and filter (candidates -> pairs).
// Correctness does not depend on the cell size: a partner's centre lies inside the query AABB grown by half the largest
// small extent, and cell coordinates are a monotone function of position, so no rounding can hide a pair.
#include "b200_context.h"
#include "b200_util.cuh"
#include <cstdlib>

namespace b200 {

static const int kThreads = 256;
static const int kTileThreads = 128;
static const uint32_t kRowCap = 32;              // partners kept per collider by the traversal (more: the row is rebuilt by the emit kernel)
static const uint32_t kPartStride = kRowCap + 1; // shared-memory row stride (conflict-free per thread and per row)
static const uint32_t kMaxReach = 1;
static const uint32_t kMaxRows = (2 * kMaxReach + 1) * (2 * kMaxReach + 1);
static const uint32_t kMaxSegX = 32;
static const uint32_t kMaxWin = kMaxSegX + 2 * kMaxReach;
static const uint32_t kBigStage = 32;            // big colliders staged in shared memory per CTA
static const uint32_t kBigChunk = 2048;          // colliders per big-row chunk

B200_D bool overlapInclusive(float4 amn, float4 amx, float4 bmn, float4 bmx)
{   // bounding_volumes.h:352-358
    if (amx.x < bmn.x || amn.x > bmx.x) return false;
    if (amx.y < bmn.y || amn.y > bmx.y) return false;
    if (amx.z < bmn.z || amn.z > bmx.z) return false;
    return true;
}

B200_D float maxExtent(float4 mn, float4 mx) { return fmaxf(mx.x - mn.x, fmaxf(mx.y - mn.y, mx.z - mn.z)); }

B200_D int cellCoord(float x, float origin, float inv, int dim)
{
    // monotone non-decreasing in x (fp subtract, multiply by a positive constant, floor, clamp are all monotone)
    float f = floorf((x - origin) * inv);
    f = fminf(fmaxf(f, 0.f), (float)(dim - 1));
    return (int)f;
}

// cells that can hold the centre of a small collider overlapping the box [amn, amx]
struct CellRange { int x0, x1, y0, y1, z0, z1; };
B200_D CellRange queryRange(const GridParams& g, float4 amn, float4 amx, float halfE)
{
    const float slackX = halfE + halfE * 1e-3f + fabsf(amn.x) * 1e-6f + fabsf(amx.x) * 1e-6f;
    const float slackY = halfE + halfE * 1e-3f + fabsf(amn.y) * 1e-6f + fabsf(amx.y) * 1e-6f;
    const float slackZ = halfE + halfE * 1e-3f + fabsf(amn.z) * 1e-6f + fabsf(amx.z) * 1e-6f;
    CellRange r;
    r.x0 = cellCoord(amn.x - slackX, g.origin[0], g.invCellSize, g.dims[0]); r.x1 = cellCoord(amx.x + slackX, g.origin[0], g.invCellSize, g.dims[0]);
    r.y0 = cellCoord(amn.y - slackY, g.origin[1], g.invCellSize, g.dims[1]); r.y1 = cellCoord(amx.y + slackY, g.origin[1], g.invCellSize, g.dims[1]);
    r.z0 = cellCoord(amn.z - slackZ, g.origin[2], g.invCellSize, g.dims[2]); r.z1 = cellCoord(amx.z + slackZ, g.origin[2], g.invCellSize, g.dims[2]);
    return r;
}

// ------------------------------------------------------------------------------------------------------ statistics
// first step (and after an upload): mean extent -> threshold, the way every later step gets it from its predecessor's setup
__global__ void __launch_bounds__(kThreads) bpPrimeKernel(uint32_t C, const float4* __restrict__ mn, const float4* __restrict__ mx, DeviceCounts* cnt)
{
    unsigned long long sum = 0; uint32_t n = 0;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < C; i += gridDim.x * kThreads)
    {
        float e = maxExtent(ldStream(mn + i), ldStream(mx + i));
        if (e >= 0.f && e < 1e12f) { sum += (unsigned long long)((double)e * 65536.0); ++n; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_down_sync(0xffffffffu, sum, o); n += __shfl_down_sync(0xffffffffu, n, o); }
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
__global__ void bpPrimeFinishKernel(float bigFactor, float cfgCell, DeviceCounts* cnt)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double mean = cnt->extentCount ? ((double)cnt->extentSumFixed / 65536.0 / (double)cnt->extentCount) : 1.0;
    cnt->grid.bigThreshold = cfgCell > 0.f ? cfgCell : (float)(mean * (double)bigFactor);
    cnt->extentSumFixed = 0ull; cnt->extentCount = 0; cnt->smallSumFixed = 0ull; cnt->smallCount = 0; cnt->maxSmallExtentBits = 0; cnt->numBig = 0;
    for (int a = 0; a < 3; ++a) { cnt->centreMinBits[a] = 0xFFFFFFFFu; cnt->centreMaxBits[a] = 0u; }
}

// With candidate reuse (inflMin != nullptr) the pipeline works on boxes grown by the step's margin: they are produced here, stored for the
// later stages (which only ever see the grown boxes) and for the containment checks of the following steps.
__global__ void __launch_bounds__(kThreads) bpClassifyKernel(uint32_t C, const float4* __restrict__ mn, const float4* __restrict__ mx, float4* __restrict__ inflMin, float4* __restrict__ inflMax,
                                                             uint32_t* __restrict__ cellOf, uint32_t* __restrict__ bigList, DeviceCounts* cnt)
{
    if (!cnt->bpRebuild) return;
    const float margin = inflMin ? __uint_as_float(cnt->bpMarginBits) : 0.f;
    const float threshold = cnt->grid.bigThreshold;
    uint32_t extBits = 0, lo[3] = { 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu }, hi[3] = { 0u, 0u, 0u };
    unsigned long long sumAll = 0, sumSmall = 0; uint32_t nAll = 0, nSmall = 0;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < C; i += gridDim.x * kThreads)
    {
        float4 a = ldStream(mn + i), b = ldStream(mx + i);
        if (inflMin)
        {
            if (!(a.x > b.x)) { a.x -= margin; a.y -= margin; a.z -= margin; b.x += margin; b.y += margin; b.z += margin; }
            inflMin[i] = a; inflMax[i] = b;
        }
        const float e = maxExtent(a, b);
        if (a.x > b.x) { cellOf[i] = 0xFFFFFFFFu; continue; }   // inverted box = halo padding slot: neither binned nor big
        if (e >= 0.f && e < 1e12f) { sumAll += (unsigned long long)((double)e * 65536.0); ++nAll; }
        if (!(e <= threshold))
        {   // big (or non-finite): brute force
            const uint32_t k = atomicAdd(&cnt->numBig, 1u);
            if (k < kMaxBig) bigList[k] = i; else cnt->overflowBig = 1;
            cellOf[i] = 0xFFFFFFFEu;
        }
        else
        {
            cellOf[i] = 0;
            sumSmall += (unsigned long long)((double)fmaxf(e, 0.f) * 65536.0); ++nSmall;
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
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        sumAll += __shfl_down_sync(0xffffffffu, sumAll, o); sumSmall += __shfl_down_sync(0xffffffffu, sumSmall, o);
        nAll += __shfl_down_sync(0xffffffffu, nAll, o); nSmall += __shfl_down_sync(0xffffffffu, nSmall, o);
    }
    __shared__ uint32_t sRed[kThreads / 32][9]; __shared__ unsigned long long sSum[kThreads / 32][2];
    if (laneId() == 0)
    {
        uint32_t* r = sRed[threadIdx.x >> 5];
        r[0] = extBits; for (int a = 0; a < 3; ++a) { r[1 + a] = lo[a]; r[4 + a] = hi[a]; } r[7] = nAll; r[8] = nSmall;
        sSum[threadIdx.x >> 5][0] = sumAll; sSum[threadIdx.x >> 5][1] = sumSmall;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {   // one set of atomics per CTA (integer: the totals do not depend on the order)
        for (int w = 1; w < kThreads / 32; ++w)
        {
            extBits = max(extBits, sRed[w][0]);
            for (int a = 0; a < 3; ++a) { lo[a] = min(lo[a], sRed[w][1 + a]); hi[a] = max(hi[a], sRed[w][4 + a]); }
            nAll += sRed[w][7]; nSmall += sRed[w][8]; sumAll += sSum[w][0]; sumSmall += sSum[w][1];
        }
        if (extBits) atomicMax(&cnt->maxSmallExtentBits, extBits);
        for (int a = 0; a < 3; ++a) { if (lo[a] != 0xFFFFFFFFu) atomicMin(&cnt->centreMinBits[a], lo[a]); if (hi[a]) atomicMax(&cnt->centreMaxBits[a], hi[a]); }
        if (nAll) { atomicAdd(&cnt->extentSumFixed, sumAll); atomicAdd(&cnt->extentCount, nAll); }
        if (nSmall) { atomicAdd(&cnt->smallSumFixed, sumSmall); atomicAdd(&cnt->smallCount, nSmall); }
    }
}

// single thread: derive the grid from the statistics (all on device, no host round trip), leave the accumulators ready for the next step
__global__ void bpSetupKernel(uint32_t maxCells, float bigFactor, float cfgCell, uint32_t cfgCellDiv, uint32_t tileCand, uint32_t tilesMode, DeviceCounts* cnt)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (!cnt->bpRebuild) return;        // the candidate list is reused: grid, statistics and threshold stay as the last build left them (bpDecideKernel did the per-step resets)
    GridParams g;
    const float E = __uint_as_float(cnt->maxSmallExtentBits);
    const float meanSmall = cnt->smallCount ? (float)((double)cnt->smallSumFixed / 65536.0 / (double)cnt->smallCount) : E;
    // cells per largest small extent: 1 when the colliders are about the same size (a sphere pile), 2 when the typical collider is
    // clearly smaller than the largest (mixed shapes: fewer candidates per query at the price of more, emptier cells)
    // cells per largest small extent: 1 when the colliders are about the same size (a sphere pile: 30 candidates for 25 partners), 2 when
    // the typical collider is clearly smaller than the largest (mixed shapes: 146 -> 79 candidates per query, measured 842 -> 643 us at 1 M)
    float lo[3], hi[3];
    const bool any = cnt->centreMinBits[0] != 0xFFFFFFFFu;
    for (int a = 0; a < 3; ++a) { lo[a] = any ? sortableToFloat(cnt->centreMinBits[a]) : 0.f; hi[a] = any ? sortableToFloat(cnt->centreMaxBits[a]) : 0.f; }
    // ... and only where the coarse cells are crowded (>= 2 colliders per cell): a sparse scene (ragdolls over a wide floor) would pay for
    // eight times as many, mostly empty cells in the cell scan
    double cells1 = 1.0;
    { const float s1 = E * 1.001f + 1e-6f; if (s1 > 1e-6f) for (int a = 0; a < 3; ++a) cells1 *= floor(((double)hi[a] - (double)lo[a]) / (double)s1) + 2.0; }
    const bool crowded = (double)cnt->smallCount >= 2.0 * cells1;
    uint32_t div = cfgCellDiv ? cfgCellDiv : ((E > 0.f && meanSmall < 0.8f * E && crowded) ? 2u : 1u);
    if (div > 3u) div = 3u;
    float s = 1.f;
    for (;; --div)
    {
        s = (E * 1.001f + 1e-6f) / (float)div;
        if (!(s > 1e-6f)) s = 1.f;
        double total = 1.0; bool ok = true;
        for (int a = 0; a < 3; ++a)
        {
            double d = floor(((double)hi[a] - (double)lo[a]) / (double)s) + 2.0;
            if (d > 1048576.0) ok = false;
            total *= d;
        }
        if ((ok && total <= (double)maxCells) || div == 1) break;
    }
    // still too many cells at one cell per extent: coarsen until the dense grid fits the budget (a coarser grid is slower, never wrong)
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
    // reach of a query in cells: the query box is at most (1 + 2 * 0.5005) E long, i.e. it ends at most E * 1.0005 / s < div + 1 cells from
    // the centre cell (colliders at huge coordinates can exceed it: the tile kernel checks and falls back to a direct traversal)
    const uint32_t reachNeeded = max((uint32_t)ceilf(E * 1.0006f / s), 1u);
    g.reach = min(reachNeeded, kMaxReach);
    // a tile is one warp's work: about 32 colliders, and the (2 reach + 1)^2 neighbouring runs should fit the staging buffer in one phase
    const double perCell = (double)cnt->smallCount / (double)(g.numCells ? g.numCells : 1u);
    const double rowsN = (double)((2u * g.reach + 1u) * (2u * g.reach + 1u));
    double sx = perCell > 0.0 ? 32.0 / perCell : (double)kMaxSegX;       // (candidate runs beyond the staging buffer are staged in phases)
    (void)rowsN; (void)tileCand;
    g.segX = (uint32_t)fmin(fmax(sx, 1.0), (double)kMaxSegX);
    if (g.segX > (uint32_t)g.dims[0]) g.segX = (uint32_t)g.dims[0];
    g.numSegs = ((uint32_t)g.dims[0] + g.segX - 1u) / g.segX;
    g.numTiles = g.numSegs * (uint32_t)g.dims[1] * (uint32_t)g.dims[2];
    g.cellDiv = div;
    // staging pays when a cell holds several colliders (long candidate runs); around one collider per cell the direct traversal wins
    g.useTiles = tilesMode == 1u ? 1u : (tilesMode == 2u ? 0u : (perCell >= 2.0 ? 1u : 0u));
    if (reachNeeded > kMaxReach) g.useTiles = 0u;       // the staged window covers kMaxReach cells each way
    // threshold for the next step from this step's mean extent; accumulators cleared
    const double mean = cnt->extentCount ? ((double)cnt->extentSumFixed / 65536.0 / (double)cnt->extentCount) : 1.0;
    g.bigThreshold = cfgCell > 0.f ? cfgCell : (float)(mean * (double)bigFactor);
    cnt->grid = g;
    cnt->meanSmallExtent = meanSmall; cnt->bpScanCells = g.numCells;
    cnt->halfSmallExtent = 0.5f * E;
    cnt->numBigNow = min(cnt->numBig, kMaxBig);
    cnt->extentSumFixed = 0ull; cnt->extentCount = 0; cnt->smallSumFixed = 0ull; cnt->smallCount = 0; cnt->maxSmallExtentBits = 0; cnt->numBig = 0;
    for (int a = 0; a < 3; ++a) { cnt->centreMinBits[a] = 0xFFFFFFFFu; cnt->centreMaxBits[a] = 0u; }
    cnt->numPairs = 0; cnt->numPairsWanted = 0; cnt->numEpaWork = 0; cnt->epaWorkHead = 0; cnt->numEpaRedo = 0; cnt->epaRedoHead = 0; cnt->numEpaRedo2 = 0; cnt->epaRedo2Head = 0;
    for (int i = 0; i < 11; ++i) cnt->gjkCursor[i] = 0; cnt->numOverflowRows = 0;
}

__global__ void __launch_bounds__(kThreads) bpBinKernel(uint32_t C, const float4* __restrict__ mn, const float4* __restrict__ mx,
                                                        uint32_t* __restrict__ cellOf, uint32_t* __restrict__ rankInCell, uint32_t* __restrict__ cellCount, const DeviceCounts* __restrict__ cnt)
{
    if (!cnt->bpRebuild) return;
    const GridParams g = cnt->grid;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < C; i += gridDim.x * kThreads)
    {
        if (cellOf[i] >= 0xFFFFFFFEu) continue;
        const float4 a = ldStream(mn + i), b = ldStream(mx + i);
        const int cx = cellCoord((a.x + b.x) * 0.5f, g.origin[0], g.invCellSize, g.dims[0]);
        const int cy = cellCoord((a.y + b.y) * 0.5f, g.origin[1], g.invCellSize, g.dims[1]);
        const int cz = cellCoord((a.z + b.z) * 0.5f, g.origin[2], g.invCellSize, g.dims[2]);
        const uint32_t cell = ((uint32_t)cz * (uint32_t)g.dims[1] + (uint32_t)cy) * (uint32_t)g.dims[0] + (uint32_t)cx;
        cellOf[i] = cell;
        rankInCell[i] = atomicAdd(&cellCount[cell], 1u);   // the rank decides the position inside the cell only: pairs are sorted later
    }
}

// cell-ordered records: rec[2 e] = (min.xyz, collider id), rec[2 e + 1] = (max.xyz, 0); slotOf[id] = e. Colliders outside the grid
// (big, padding) get an empty row here; the big ones are counted by the tile kernel's extra CTAs.
__global__ void __launch_bounds__(kThreads) bpScatterKernel(uint32_t C, const float4* __restrict__ mn, const float4* __restrict__ mx, const uint32_t* __restrict__ cellOf,
                                                            uint32_t* __restrict__ rankToSlot, const uint32_t* __restrict__ cellStart, float4* __restrict__ rec, uint32_t* __restrict__ pairCount, const DeviceCounts* __restrict__ cnt)
{
    if (!cnt->bpRebuild) return;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < C; i += gridDim.x * kThreads)
    {
        const uint32_t cell = cellOf[i];
        if (cell >= 0xFFFFFFFEu) { pairCount[i] = 0; rankToSlot[i] = 0xFFFFFFFFu; continue; }
        const uint32_t pos = cellStart[cell] + rankToSlot[i];
        float4 a = ldStream(mn + i), b = ldStream(mx + i);
        a.w = __uint_as_float(i); b.w = 0.f;
        rec[2 * (size_t)pos] = a; rec[2 * (size_t)pos + 1] = b;
        rankToSlot[i] = pos;
    }
}

// ------------------------------------------------------------------------------------------------- TMA / mbarrier
B200_D uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
B200_D void mbarInit(uint64_t* bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smemAddr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
B200_D void mbarExpectTx(uint64_t* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smemAddr(bar)), "r"(bytes) : "memory");
}
B200_D void mbarWait(uint64_t* bar, uint32_t parity)
{
    uint32_t ok;
    do
    {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smemAddr(bar)), "r"(parity) : "memory");
    } while (!ok);
}
// contiguous global -> shared bulk copy (the 1-D TMA path: SASS UBLKCP); 16-byte aligned, size a multiple of 16
B200_D void bulkCopyG2S(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smemAddr(dstSmem)), "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar)) : "memory");
}

// ------------------------------------------------------------------------------------------------ direct traversal
// every small partner with a larger index (cells read straight from global memory) + every big partner; used where the tiled path
// does not apply: a reach beyond the staged window, rows longer than kRowCap. F is called with each partner id.
template <typename F>
B200_D void traverseDirect(const GridParams& g, float halfE, uint32_t id, float4 amn, float4 amx, const float4* __restrict__ rec, const uint32_t* __restrict__ cellStart,
                           const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax, const uint32_t* __restrict__ bigList, uint32_t numBig, F&& f)
{
    const CellRange r = queryRange(g, amn, amx, halfE);
    for (int z = r.z0; z <= r.z1; ++z)
        for (int y = r.y0; y <= r.y1; ++y)
        {
            const uint32_t row = ((uint32_t)z * (uint32_t)g.dims[1] + (uint32_t)y) * (uint32_t)g.dims[0];
            const uint32_t f0 = cellStart[row + r.x0], f1 = cellStart[row + r.x1 + 1];
            for (uint32_t e = f0; e < f1; ++e)
            {
                const float4 bmn = rec[2 * (size_t)e];
                const uint32_t jd = __float_as_uint(bmn.w);
                if (jd <= id) continue;
                if (overlapInclusive(amn, amx, bmn, rec[2 * (size_t)e + 1])) f(jd);
            }
        }
    for (uint32_t k = 0; k < numBig; ++k)
    {
        const uint32_t jd = bigList[k];
        if (jd <= id) continue;
        if (overlapInclusive(amn, amx, aabbMin[jd], aabbMax[jd])) f(jd);
    }
}

// ----------------------------------------------------------------------------------------------------------- tiles
// A tile = a run of consecutive cells of one grid row; ONE WARP works on it from start to finish (no block-level barrier anywhere):
// the lanes own up to 32 colliders of the run, lane j plans and issues the bulk copy of neighbouring row j, everybody waits on the
// warp's own mbarrier, tests, sorts, writes rows. Sixteen or so such warps are resident per SM, each at a different point of its
// tile, so the copy latency of one warp is covered by the tests of the others.
static const int kWarpsPerCta = 4;
static const uint32_t kBulkMin = 2048;           // bytes from which a staged run is worth a TMA bulk copy
static const uint32_t kWarpCand = 256;           // candidate records staged per warp and phase (8 KB)
struct WarpTile
{
    float4 rec[2 * kWarpCand];                      // staged candidate records
    uint32_t part[32 * kPartStride];                // partners per lane
    uint32_t start[kMaxRows * (kMaxWin + 1)];       // cell starts of the staged window, per neighbouring row
    uint32_t pcG0[kMaxRows], pcLen[kMaxRows], pcS0[kMaxRows];     // piece of row j staged in the current phase (len 0: none)
    alignas(8) uint64_t bar;
};
struct TileCtaShared { WarpTile w[kWarpsPerCta]; float4 bigMin[kBigStage], bigMax[kBigStage]; uint32_t bigId[kBigStage]; };

__global__ void __launch_bounds__(kTileThreads) bpTileKernel(uint32_t C, uint32_t tileCtas, uint32_t numChunks, const float4* __restrict__ rec, const uint32_t* __restrict__ cellStart,
    const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax, const uint32_t* __restrict__ bigList,
    uint32_t* __restrict__ pairCount, uint32_t* __restrict__ rows, uint32_t* __restrict__ ovList, uint32_t* __restrict__ bigChunkCount, DeviceCounts* cnt)
{
    extern __shared__ __align__(128) unsigned char tileSmem[];
    TileCtaShared& CS = *reinterpret_cast<TileCtaShared*>(tileSmem);
    if (!cnt->bpRebuild) return;
    const GridParams g = cnt->grid;
    const uint32_t numBig = cnt->numBigNow;
    const uint32_t tid = threadIdx.x;

    if (blockIdx.x >= tileCtas)
    {   // ---- rows of the big colliders, counting pass: chunk c of the collider array against every big collider
        __shared__ uint32_t sWarp[kTileThreads / 32];
        for (uint32_t chunk = blockIdx.x - tileCtas; chunk < numChunks; chunk += gridDim.x - tileCtas)
        {
            const uint32_t j0 = chunk * kBigChunk;
            for (uint32_t k = 0; k < numBig; ++k)
            {
                const uint32_t id = bigList[k];
                const float4 amn = aabbMin[id], amx = aabbMax[id];
                uint32_t n = 0;
                for (uint32_t j = j0 + tid; j < min(C, j0 + kBigChunk); j += kTileThreads)
                    if (j > id && overlapInclusive(amn, amx, ldStream(aabbMin + j), ldStream(aabbMax + j))) ++n;
                n = __reduce_add_sync(0xffffffffu, n);
                if (laneId() == 0) sWarp[tid >> 5] = n;
                __syncthreads();
                if (tid == 0)
                {
                    uint32_t t = 0; for (int w = 0; w < kTileThreads / 32; ++w) t += sWarp[w];
                    bigChunkCount[(size_t)k * numChunks + chunk] = t;
                    if (t) atomicAdd(&pairCount[id], t);
                }
                __syncthreads();
            }
        }
        return;
    }

    if (!g.useTiles) return;
    const float halfE = cnt->halfSmallExtent;
    const uint32_t R = g.reach, side = 2u * R + 1u, nRows = side * side, nx = (uint32_t)g.dims[0], ny = (uint32_t)g.dims[1], nz = (uint32_t)g.dims[2];
    const uint32_t lane = tid & 31u, warpInCta = tid >> 5;
    WarpTile& S = CS.w[warpInCta];
    if (lane == 0) mbarInit(&S.bar, 1);
    // big colliders staged once per CTA
    const uint32_t nBigS = min(numBig, kBigStage);
    if (tid < nBigS) { const uint32_t b = bigList[tid]; CS.bigId[tid] = b; CS.bigMin[tid] = aabbMin[b]; CS.bigMax[tid] = aabbMax[b]; }
    __syncthreads();
    uint32_t parity = 0;
    const uint32_t numWarps = tileCtas * kWarpsPerCta;

    for (uint32_t tile = blockIdx.x * kWarpsPerCta + warpInCta; tile < g.numTiles; tile += numWarps)
    {
        const uint32_t seg = tile % g.numSegs, rowId = tile / g.numSegs;
        const uint32_t y = rowId % ny, z = rowId / ny;
        const uint32_t xa = seg * g.segX, xb = min(xa + g.segX, nx);
        const uint32_t rowBase = rowId * nx;
        const uint32_t o0 = cellStart[rowBase + xa], o1 = cellStart[rowBase + xb];
        if (o0 == o1) continue;                               // empty tile (uniform across the warp)
        const uint32_t wx0 = xa > R ? xa - R : 0u, wx1 = min(xb + R, nx), W = wx1 - wx0;
        // cell starts of the window in every neighbouring row
        {
            uint32_t j = 0;
            for (int dz = -(int)R; dz <= (int)R; ++dz)
                for (int dy = -(int)R; dy <= (int)R; ++dy, ++j)
                {
                    const int yy = (int)y + dy, zz = (int)z + dz;
                    const bool valid = yy >= 0 && yy < (int)ny && zz >= 0 && zz < (int)nz;
                    const uint32_t* src = cellStart + ((uint32_t)(valid ? zz : 0) * ny + (uint32_t)(valid ? yy : 0)) * nx + wx0;
                    for (uint32_t xi = lane; xi <= W; xi += 32u) S.start[j * (kMaxWin + 1u) + xi] = valid ? src[xi] : 0u;
                }
        }
        __syncwarp();
        // lane j < nRows looks after neighbouring row j: its staged range in global record indices
        uint32_t rowG0 = 0, rowLen = 0;
        if (lane < nRows) { rowG0 = S.start[lane * (kMaxWin + 1u)]; rowLen = S.start[lane * (kMaxWin + 1u) + W] - rowG0; }

        for (uint32_t ob = o0; ob < o1; ob += 32u)
        {
            const uint32_t e = ob + lane;
            const bool active = e < o1;
            float4 amn = make_float4(0, 0, 0, 0), amx = amn; uint32_t id = 0, count = 0; bool direct = false;
            CellRange q = { 0, -1, 0, -1, 0, -1 };
            if (active)
            {
                amn = rec[2 * (size_t)e]; amx = rec[2 * (size_t)e + 1]; id = __float_as_uint(amn.w);
                q = queryRange(g, amn, amx, halfE);
                // the staged window covers cells [wx0, wx1) of the rows within R of this row; anything else is traversed directly
                direct = q.x0 < (int)wx0 || q.x1 >= (int)wx1 || q.y0 + (int)R < (int)y || q.y1 > (int)(y + R) || q.z0 + (int)R < (int)z || q.z1 > (int)(z + R);
            }
            uint32_t doneOff = 0;      // lane j: records of its row already staged in earlier phases
            for (;;)
            {
                // ---- plan one phase: rows in order, as many records as the staging buffer holds (prefix sum over the lanes)
                const uint32_t remaining = rowLen - doneOff;
                uint32_t incl = remaining;
#pragma unroll
                for (int o = 1; o < 16; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += t; }
                const uint32_t excl = incl - remaining;
                const uint32_t take = excl >= kWarpCand ? 0u : min(remaining, kWarpCand - excl);
                const uint32_t total = min(__shfl_sync(0xffffffffu, incl, 15), kWarpCand);
                if (total == 0) break;
                if (lane < nRows) { S.pcG0[lane] = rowG0 + doneOff; S.pcLen[lane] = take; S.pcS0[lane] = excl; }
                // Staging. A run of at least kBulkMin bytes goes through the TMA unit (one cp.async.bulk, completion on the warp's
                // mbarrier); shorter runs are copied by the lanes with 16-byte cp.async: measured, a bulk copy costs the SM's single
                // TMA unit ~150 ns whatever its size, so thousands of sub-kilobyte copies per SM serialise there.
                const bool bulk = take * 32u >= kBulkMin;
                const uint32_t bulkBytes = __reduce_add_sync(0xffffffffu, bulk ? take * 32u : 0u);
                if (bulkBytes && lane == 0) mbarExpectTx(&S.bar, bulkBytes);
                __syncwarp();
                if (bulk) bulkCopyG2S(&S.rec[2 * excl], rec + 2 * (size_t)(rowG0 + doneOff), take * 32u, &S.bar);
                for (uint32_t j = 0; j < nRows; ++j)
                {
                    const uint32_t len = S.pcLen[j];
                    if (!len || len * 32u >= kBulkMin) continue;                      // (uniform)
                    const float4* src = rec + 2 * (size_t)S.pcG0[j]; float4* dst = &S.rec[2 * S.pcS0[j]];
                    for (uint32_t u = lane; u < 2u * len; u += 32u)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" :: "r"(smemAddr(dst + u)), "l"(src + u) : "memory");
                }
                asm volatile("cp.async.wait_all;" ::: "memory");
                doneOff += take;
                if (bulkBytes) { mbarWait(&S.bar, parity); parity ^= 1u; }
                __syncwarp();
                if (active && !direct)
                {
                    for (uint32_t j = 0; j < nRows; ++j)
                    {
                        const uint32_t len = S.pcLen[j];
                        if (!len) continue;
                        const int yy = (int)y + (int)(j % side) - (int)R, zz = (int)z + (int)(j / side) - (int)R;
                        if (yy < q.y0 || yy > q.y1 || zz < q.z0 || zz > q.z1) continue;
                        const uint32_t* st = &S.start[j * (kMaxWin + 1u)];
                        const uint32_t g0 = S.pcG0[j];
                        const uint32_t f0 = max(st[(uint32_t)q.x0 - wx0], g0), f1 = min(st[(uint32_t)q.x1 + 1u - wx0], g0 + len);
                        const uint32_t sBase = S.pcS0[j] - g0;
                        // four candidates per trip (eight independent shared-memory loads in flight); the last trip re-reads the final record
                        // for the slots beyond the range and ignores them
                        for (uint32_t f = f0; f < f1; f += 4u)
                        {
                            float4 bmn[4], bmx[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) { const uint32_t fi = min(f + (uint32_t)u, f1 - 1u) + sBase; bmn[u] = S.rec[2 * fi]; bmx[u] = S.rec[2 * fi + 1]; }
#pragma unroll
                            for (int u = 0; u < 4; ++u)
                            {
                                const uint32_t jd = __float_as_uint(bmn[u].w);
                                if (f + (uint32_t)u < f1 && jd > id && overlapInclusive(amn, amx, bmn[u], bmx[u])) { if (count < kRowCap) S.part[lane * kPartStride + count] = jd; ++count; }
                            }
                        }
                    }
                }
                __syncwarp();                                           // every lane is done with the staged records
            }
            if (active)
            {
                if (!direct)
                {   // big colliders (staged; beyond kBigStage from global memory)
                    for (uint32_t k = 0; k < numBig; ++k)
                    {
                        uint32_t jd; float4 bmn, bmx;
                        if (k < nBigS) { jd = CS.bigId[k]; bmn = CS.bigMin[k]; bmx = CS.bigMax[k]; } else { jd = bigList[k]; bmn = aabbMin[jd]; bmx = aabbMax[jd]; }
                        if (jd <= id || !overlapInclusive(amn, amx, bmn, bmx)) continue;
                        if (count < kRowCap) S.part[lane * kPartStride + count] = jd;
                        ++count;
                    }
                }
                else
                {
                    count = 0;
                    traverseDirect(g, halfE, id, amn, amx, rec, cellStart, aabbMin, aabbMax, bigList, numBig,
                                   [&](uint32_t jd) { if (count < kRowCap) S.part[lane * kPartStride + count] = jd; ++count; });
                }
                pairCount[id] = count;
                if (count > kRowCap) { const uint32_t k = atomicAdd(&cnt->numOverflowRows, 1u); if (k < C) ovList[k] = id; }    // rebuilt by the emit kernel
                // ascending partner order (the canonical pair order)
                const uint32_t n = min(count, kRowCap);
                uint32_t* P = &S.part[lane * kPartStride];
                for (uint32_t a = 1; a < n; ++a)
                {
                    const uint32_t key = P[a]; uint32_t b = a;
                    while (b > 0 && P[b - 1] > key) { P[b] = P[b - 1]; --b; }
                    P[b] = key;
                }
            }
            if (active)
            {   // the lane's 128-byte row, 16 bytes at a time
                const uint32_t n = min(count, kRowCap);
                const uint32_t* P = &S.part[lane * kPartStride];
                uint4* dst = reinterpret_cast<uint4*>(rows + (size_t)e * kRowCap);
                for (uint32_t k = 0; k < n; k += 4u) dst[k >> 2] = make_uint4(P[k], P[k + 1], P[k + 2], P[k + 3]);     // (the row buffer is 33 words: k + 3 <= 32)
            }
            __syncwarp();
        }
    }
}

// The same traversal without staging: one thread per cell-ordered collider reads its candidate runs straight from global memory (L1 / L2).
// Wins where a cell holds about one collider (sphere piles, ragdoll fields): too few records per run for staging to pay.
static const int kDirectThreads = 256;
__global__ void __launch_bounds__(kDirectThreads) bpRowsDirectKernel(uint32_t C, const float4* __restrict__ rec, const uint32_t* __restrict__ cellStart,
    const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax, const uint32_t* __restrict__ bigList,
    uint32_t* __restrict__ pairCount, uint32_t* __restrict__ rows, uint32_t* __restrict__ ovList, DeviceCounts* cnt)
{
    __shared__ uint32_t sPart[kDirectThreads * kPartStride];
    __shared__ float4 sBigMin[kBigStage], sBigMax[kBigStage]; __shared__ uint32_t sBigId[kBigStage];
    const GridParams g = cnt->grid;
    if (g.useTiles || !cnt->bpRebuild) return;
    const uint32_t numBig = cnt->numBigNow, nBigS = min(numBig, kBigStage);
    const float halfE = cnt->halfSmallExtent;
    if (threadIdx.x < nBigS) { const uint32_t b = bigList[threadIdx.x]; sBigId[threadIdx.x] = b; sBigMin[threadIdx.x] = aabbMin[b]; sBigMax[threadIdx.x] = aabbMax[b]; }
    __syncthreads();
    const uint32_t numBinned = cellStart[g.numCells];
    uint32_t* P = &sPart[threadIdx.x * kPartStride];
    for (uint32_t e = blockIdx.x * kDirectThreads + threadIdx.x; e < numBinned; e += gridDim.x * kDirectThreads)
    {
        const float4 amn = rec[2 * (size_t)e], amx = rec[2 * (size_t)e + 1];
        const uint32_t id = __float_as_uint(amn.w);
        uint32_t count = 0;
        const CellRange r = queryRange(g, amn, amx, halfE);
        for (int z = r.z0; z <= r.z1; ++z)
            for (int y = r.y0; y <= r.y1; ++y)
            {
                const uint32_t row = ((uint32_t)z * (uint32_t)g.dims[1] + (uint32_t)y) * (uint32_t)g.dims[0];
                const uint32_t f0 = cellStart[row + r.x0], f1 = cellStart[row + r.x1 + 1];
                for (uint32_t f = f0; f < f1; f += 2u)
                {   // two candidates per trip: four independent loads
                    const uint32_t fb = min(f + 1u, f1 - 1u);
                    const float4 mn0 = rec[2 * (size_t)f], mx0 = rec[2 * (size_t)f + 1], mn1 = rec[2 * (size_t)fb], mx1 = rec[2 * (size_t)fb + 1];
                    const uint32_t j0 = __float_as_uint(mn0.w), j1 = __float_as_uint(mn1.w);
                    if (j0 > id && overlapInclusive(amn, amx, mn0, mx0)) { if (count < kRowCap) P[count] = j0; ++count; }
                    if (f + 1u < f1 && j1 > id && overlapInclusive(amn, amx, mn1, mx1)) { if (count < kRowCap) P[count] = j1; ++count; }
                }
            }
        for (uint32_t k = 0; k < numBig; ++k)
        {
            uint32_t jd; float4 bmn, bmx;
            if (k < nBigS) { jd = sBigId[k]; bmn = sBigMin[k]; bmx = sBigMax[k]; } else { jd = bigList[k]; bmn = aabbMin[jd]; bmx = aabbMax[jd]; }
            if (jd <= id || !overlapInclusive(amn, amx, bmn, bmx)) continue;
            if (count < kRowCap) P[count] = jd;
            ++count;
        }
        pairCount[id] = count;
        if (count > kRowCap) { const uint32_t k = atomicAdd(&cnt->numOverflowRows, 1u); if (k < C) ovList[k] = id; }
        const uint32_t n = min(count, kRowCap);
        for (uint32_t a = 1; a < n; ++a)
        {
            const uint32_t key = P[a]; uint32_t b = a;
            while (b > 0 && P[b - 1] > key) { P[b] = P[b - 1]; --b; }
            P[b] = key;
        }
        uint4* dst = reinterpret_cast<uint4*>(rows + (size_t)e * kRowCap);
        for (uint32_t k = 0; k < n; k += 4u) dst[k >> 2] = make_uint4(P[k], P[k + 1], P[k + 2], P[k + 3]);
    }
}

// -------------------------------------------------------------------------------------------------------- emit
// rows -> canonical pair list. One warp per 32 consecutive collider ids: their pairs are one contiguous run of the output, written
// with coalesced 8-byte stores; each output position finds its row by a search over the warp's 32 row starts.
__global__ void __launch_bounds__(kThreads) bpEmitKernel(uint32_t C, uint32_t emitCtas, uint32_t numChunks, const uint32_t* __restrict__ slotOf, const uint32_t* __restrict__ rows,
    const uint32_t* __restrict__ pairCount, const uint32_t* __restrict__ pairStart, const float4* __restrict__ rec, const uint32_t* __restrict__ cellStart,
    const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax, const uint32_t* __restrict__ bigList, const uint32_t* __restrict__ ovList,
    const uint32_t* __restrict__ bigChunkCount, uint2* __restrict__ pairList, uint32_t maxPairList, uint2* __restrict__ candList, uint32_t maxCandList, DeviceCounts* cnt)
{
    if (!cnt->bpRebuild) return;
    const GridParams g = cnt->grid;
    const uint32_t numBig = cnt->numBigNow;
    // with a margin the rows hold CANDIDATES (pairs of grown boxes): they go to the candidate list and bpFilterKernel derives the step's pairs
    const bool toCand = cnt->bpEmitToCand != 0u;
    uint2* __restrict__ pairs = toCand ? candList : pairList;
    const uint32_t maxPairs = toCand ? maxCandList : maxPairList;
    if (blockIdx.x == 0 && threadIdx.x == 0)
    {
        const uint32_t total = pairStart[C];
        if (toCand)
        {
            cnt->numCandWanted = total; cnt->numCand = min(total, maxPairs);
            if (total > maxPairs) { cnt->overflowPairs = 1; cnt->bpMarginOff = 1; }    // candidates lost: reported like a pair overflow; no margins from now on
        }
        else
        {
            cnt->numPairsWanted = total; cnt->numPairs = min(total, maxPairs);
            if (total > maxPairs) cnt->overflowPairs = 1;
        }
    }
    if (blockIdx.x < emitCtas)
    {
        const uint32_t lane = laneId();
        const uint32_t warpsTotal = emitCtas * (kThreads / 32), warp = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
        for (uint32_t id0 = warp * 32u; id0 < C; id0 += warpsTotal * 32u)
        {
            const uint32_t id = id0 + lane;
            const uint32_t start = id < C ? pairStart[id] : 0xFFFFFFFFu;
            uint32_t cntI = id < C ? pairCount[id] : 0u;
            const uint32_t slot = id < C ? slotOf[id] : 0xFFFFFFFFu;
            if (cntI > kRowCap || slot == 0xFFFFFFFFu) cntI = 0;      // written by the paths below (overflow rows, big colliders)
            // the warp's rows as one run of `total` entries (prefix over the lanes' counts): rows written elsewhere (overflow rows, big colliders --
            // a ground box with a low index has a row of 10^5 pairs) contribute nothing and are not walked over
            uint32_t incl = cntI;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += t; }
            const uint32_t excl = incl - cntI, total = __shfl_sync(0xffffffffu, incl, 31);
            for (uint32_t t0 = 0; t0 < total; t0 += 32u)
            {
                const uint32_t t = t0 + lane;                                   // (all lanes take part in the shuffles below)
                // row i = last lane whose first entry is at or before t (empty rows share their successor's offset and lose against it)
                uint32_t i = 0;
#pragma unroll
                for (int step = 16; step > 0; step >>= 1)
                {
                    const uint32_t cand = i + (uint32_t)step;
                    const uint32_t e = __shfl_sync(0xffffffffu, excl, (int)min(cand, 31u));
                    if (cand < 32u && e <= t) i = cand;
                }
                const uint32_t si = __shfl_sync(0xffffffffu, start, (int)i), ei = __shfl_sync(0xffffffffu, excl, (int)i), sl = __shfl_sync(0xffffffffu, slot, (int)i);
                const uint32_t k = t - ei, p = si + k;
                if (t < total && p < maxPairs) pairs[p] = make_uint2(id0 + i, rows[(size_t)sl * kRowCap + k]);
            }
        }
        return;
    }
    const float halfE = cnt->halfSmallExtent;
    const uint32_t extra = blockIdx.x - emitCtas, numExtra = gridDim.x - emitCtas;
    // ---- rows with more than kRowCap partners: direct traversal straight into the pair list, then sorted in place (rare)
    {
        const uint32_t nOv = min(cnt->numOverflowRows, C);
        for (uint32_t o = extra * kThreads + threadIdx.x; o < nOv; o += numExtra * kThreads)
        {
            const uint32_t id = ovList[o];
            const uint32_t base = pairStart[id];
            const float4 amn = aabbMin[id], amx = aabbMax[id];
            uint32_t count = 0;
            traverseDirect(g, halfE, id, amn, amx, rec, cellStart, aabbMin, aabbMax, bigList, numBig,
                           [&](uint32_t jd) { if (base + count < maxPairs) pairs[base + count] = make_uint2(id, jd); ++count; });
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
    // ---- rows of the big colliders: chunk by chunk in collider order, positions from the chunk counts
    {
        __shared__ uint32_t sWarp[kThreads / 32]; __shared__ uint32_t sOffset;
        for (uint32_t chunk = extra; chunk < numChunks; chunk += numExtra)
        {
            const uint32_t j0 = chunk * kBigChunk;
            for (uint32_t k = 0; k < numBig; ++k)
            {
                if (bigChunkCount[(size_t)k * numChunks + chunk] == 0) continue;      // uniform
                const uint32_t id = bigList[k];
                const float4 amn = aabbMin[id], amx = aabbMax[id];
                if (threadIdx.x < 32)
                {   // pairs of this row in earlier chunks
                    uint32_t s = 0;
                    for (uint32_t c2 = threadIdx.x; c2 < chunk; c2 += 32u) s += bigChunkCount[(size_t)k * numChunks + c2];
                    s = __reduce_add_sync(0xffffffffu, s);
                    if (threadIdx.x == 0) sOffset = pairStart[id] + s;
                }
                __syncthreads();
                uint32_t running = sOffset;
                for (uint32_t jb = j0; jb < min(C, j0 + kBigChunk); jb += kThreads)
                {
                    const uint32_t j = jb + threadIdx.x;
                    const bool hit = j < C && j > id && overlapInclusive(amn, amx, ldStream(aabbMin + j), ldStream(aabbMax + j));
                    const uint32_t ballot = __ballot_sync(0xffffffffu, hit);
                    if (laneId() == 0) sWarp[threadIdx.x >> 5] = (uint32_t)__popc(ballot);
                    __syncthreads();
                    uint32_t off = running, tot = 0;
                    for (uint32_t w = 0; w < kThreads / 32; ++w) { if (w < (threadIdx.x >> 5)) off += sWarp[w]; tot += sWarp[w]; }
                    off += (uint32_t)__popc(ballot & laneMaskLt());
                    if (hit) { if (off < maxPairs) pairs[off] = make_uint2(id, j); else cnt->overflowPairs = 1; }
                    running += tot;
                    __syncthreads();
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- candidate reuse
// The pair set is a function of the collider indices and their world AABBs only. The grid pipeline above therefore may run on boxes grown
// by a margin m and leave a CANDIDATE list (canonical order, like the pair list); as long as every collider's true AABB lies inside its
// grown box, every overlapping pair of true boxes is a pair of overlapping grown boxes, i.e. a candidate, and the step's pair list is the
// candidate list filtered -- in order -- by the reference's predicate on the true boxes: the SAME list the pipeline would have produced,
// bit for bit (tests compare it with the oracle's sweep in every step). In a settled pile the list survives hundreds of steps; a falling
// or thrown scene rebuilds every step with margin 0 and pays one extra pass over the AABBs.
//   check  : every collider: is the true box inside the grown one? (+ how far the faces have moved: the policy's input)
//   decide : one thread: rebuild or reuse, the margin of a new list, the sizes of this step's scans, the per-step counter resets
//   filter : candidates -> pairs, single pass (decoupled look-back over 2048-candidate tiles), stable
// margin policy (bpDecideKernel): a new list gets the margin that lets it live about kTargetLife steps at the speed the box faces moved since the
// last build, between kMinMarginFraction and maxFraction of the mean small extent; no margin (exact build, no filter pass) when even the largest
// margin would not survive kMinReuseSteps steps
static const float kMaxMarginFraction = 0.06f, kMinMarginFraction = 0.02f;
static const float kTargetLife = 6.f, kMinReuseSteps = 2.5f;

__global__ void __launch_bounds__(kThreads) bpCheckKernel(uint32_t force, uint32_t C, const float4* __restrict__ mn, const float4* __restrict__ mx,
                                                          const float4* __restrict__ inflMin, const float4* __restrict__ inflMax, DeviceCounts* cnt)
{
    if (force) return;
    uint32_t viol = 0; float exc = -3.0e38f;
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < C; i += gridDim.x * kThreads)
    {
        const float4 a = ldStream(mn + i), b = ldStream(mx + i);
        if (a.x > b.x) continue;                                  // padding slot now: takes part in no pair
        const float4 p = ldStream(inflMin + i), q = ldStream(inflMax + i);
        if (p.x > q.x) { ++viol; continue; }                      // was a padding slot when the list was built
        const bool inside = a.x >= p.x && a.y >= p.y && a.z >= p.z && b.x <= q.x && b.y <= q.y && b.z <= q.z;      // (false for NaN)
        if (!inside) ++viol;
        exc = fmaxf(exc, fmaxf(fmaxf(p.x - a.x, fmaxf(p.y - a.y, p.z - a.z)), fmaxf(b.x - q.x, fmaxf(b.y - q.y, b.z - q.z))));
    }
    viol = __reduce_add_sync(0xffffffffu, viol);
    uint32_t excBits = __reduce_max_sync(0xffffffffu, floatToSortable(exc));
    __shared__ uint32_t sV[kThreads / 32], sE[kThreads / 32];
    if (laneId() == 0) { sV[threadIdx.x >> 5] = viol; sE[threadIdx.x >> 5] = excBits; }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        for (int w = 1; w < kThreads / 32; ++w) { viol += sV[w]; excBits = max(excBits, sE[w]); }
        if (viol) atomicAdd(&cnt->bpViolations, viol);
        atomicMax(&cnt->bpExcessBits, excBits);
    }
}

__global__ void bpDecideKernel(uint32_t force, uint32_t C, uint32_t maxCand, float maxFraction, float targetLife, DeviceCounts* cnt)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const float lastMargin = __uint_as_float(cnt->bpMarginBits);
    if (force) cnt->bpMarginOff = 0;
    // a list built without margin was written straight into the pair list of its step: nothing to reuse
    // ... and a list that did not fit its buffer (reported as a pair overflow when it happened) is not complete
    const bool rebuild = force || cnt->bpViolations != 0u || !(lastMargin > 0.f) || !cnt->bpEmitToCand || cnt->numCandWanted > cnt->numCand;
    // how far did the box faces move per step since the last build? (the check's excess is relative to the grown box)
    const float moved = force ? 0.f : fmaxf(sortableToFloat(cnt->bpExcessBits) + fmaxf(lastMargin, 0.f), 0.f);
    const float perStep = moved / (float)max(cnt->bpAge, 1u);
    if (!force) cnt->bpMovedPerStep = perStep;
    if (rebuild)
    {
        float margin = 0.f;
        if (!force && !cnt->bpMarginOff)
        {
            const float ext = cnt->meanSmallExtent;
            const float mMax = maxFraction * ext, mMin = kMinMarginFraction * ext;
            // room for the extra candidates? (the last step's exact count is the yardstick)
            const bool room = (double)cnt->numPairsWanted * 1.4 < (double)maxCand;
            if (mMax > 0.f && mMax < 1e30f && perStep * kMinReuseSteps <= mMax && room) margin = fminf(fmaxf(perStep * targetLife, mMin), mMax);
        }
        cnt->bpMarginBits = __float_as_uint(margin);
        cnt->bpEmitToCand = margin > 0.f ? 1u : 0u;
        cnt->bpAge = 0; ++cnt->bpRebuildCount;
        cnt->bpScanColliders = C;          // (bpScanCells: bpSetupKernel, once the grid is known)
    }
    else { cnt->bpScanCells = 0; cnt->bpScanColliders = 0; }
    cnt->bpRebuild = rebuild ? 1u : 0u;
    ++cnt->bpAge; ++cnt->bpStepCount;
    cnt->bpViolations = 0; cnt->bpExcessBits = 0;
    // per-step resets (bpSetupKernel does them when it runs; it does not on a reuse step)
    cnt->numPairs = 0; cnt->numPairsWanted = 0; cnt->numEpaWork = 0; cnt->epaWorkHead = 0; cnt->numEpaRedo = 0; cnt->epaRedoHead = 0; cnt->numEpaRedo2 = 0; cnt->epaRedo2Head = 0;
    for (int i = 0; i < 11; ++i) cnt->gjkCursor[i] = 0; cnt->numOverflowRows = 0;
}

// candidates -> pairs. A tile is 2048 consecutive candidates; warp w takes 256 of them in four coalesced 512-byte loads (lane l: candidates
// 2 l and 2 l + 1 of each 64), tests them against the true AABBs (the first collider repeats along the list: broadcast loads; the second is a
// gather into the L2-resident AABB planes), ranks the survivors with two ballots per load and writes them at the tile's offset, which comes
// from the look-back chain. Order is preserved, so the output is the canonical pair list.
static const int kFilterThreads = 256;
static const uint32_t kFilterTile = 2048;
B200_D bool trueBoxesOverlap(uint32_t ia, uint32_t ib, const float4* __restrict__ mn, const float4* __restrict__ mx)
{
    const float4 amn = mn[ia], amx = mx[ia], bmn = mn[ib], bmx = mx[ib];
    if (amn.x > amx.x || bmn.x > bmx.x) return false;            // padding slots are in no pair (bpClassifyKernel never bins them)
    return overlapInclusive(amn, amx, bmn, bmx);
}
__global__ void __launch_bounds__(kFilterThreads) bpFilterKernel(const uint2* __restrict__ cand, const float4* __restrict__ mn, const float4* __restrict__ mx,
    uint2* __restrict__ pairs, uint32_t maxPairs, unsigned long long* __restrict__ status, ScanLbState* __restrict__ state, DeviceCounts* cnt)
{
    if (cnt->bpRebuild && !cnt->bpEmitToCand) return;          // this step's pipeline wrote the pair list itself (margin 0)
    const uint32_t epoch = scanEpochOf(state);
    __shared__ uint32_t warpSums[kFilterThreads / 32];
    __shared__ uint32_t sTile, sPrefix, sTileTotal;
    const uint32_t n = cnt->numCand;
    const uint32_t numTiles = (n + kFilterTile - 1) / kFilterTile;
    const uint32_t lane = laneId(), warp = threadIdx.x >> 5, lt = laneMaskLt();
    for (;;)
    {
        if (threadIdx.x == 0) sTile = atomicAdd(&state->ticket, 1u);
        __syncthreads();
        const uint32_t tile = sTile;
        if (tile >= numTiles) break;
        const uint32_t warpBase = tile * kFilterTile + warp * 256u;
        uint4 c4[4]; uint32_t rank[4], keep[4], warpTotal = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const uint32_t idx = warpBase + (uint32_t)j * 64u + 2u * lane;
            c4[j] = make_uint4(0, 0, 0, 0);
            if (idx + 1u < n) c4[j] = *reinterpret_cast<const uint4*>(cand + idx);
            else if (idx < n) { const uint2 t = cand[idx]; c4[j].x = t.x; c4[j].y = t.y; }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            const uint32_t idx = warpBase + (uint32_t)j * 64u + 2u * lane;
            const bool k0 = idx < n && trueBoxesOverlap(c4[j].x, c4[j].y, mn, mx);
            const bool k1 = idx + 1u < n && trueBoxesOverlap(c4[j].z, c4[j].w, mn, mx);
            const uint32_t b0 = __ballot_sync(0xffffffffu, k0), b1 = __ballot_sync(0xffffffffu, k1);
            rank[j] = warpTotal + (uint32_t)__popc(b0 & lt) + (uint32_t)__popc(b1 & lt);       // of this lane's first candidate inside the warp's 256
            keep[j] = (k0 ? 1u : 0u) | (k1 ? 2u : 0u);
            warpTotal += (uint32_t)__popc(b0) + (uint32_t)__popc(b1);
        }
        if (lane == 0) warpSums[warp] = warpTotal;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            const uint32_t w = (threadIdx.x < kFilterThreads / 32) ? warpSums[threadIdx.x] : 0u;
            uint32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= (uint32_t)o) wi += t; }
            if (threadIdx.x < kFilterThreads / 32) warpSums[threadIdx.x] = wi - w;
            const uint32_t total = __shfl_sync(0xffffffffu, wi, kFilterThreads / 32 - 1);
            const uint32_t prefix = scanLookbackPrefix(status, tile, total, epoch);
            if (threadIdx.x == 0)
            {
                sPrefix = prefix; sTileTotal = total;
                if (tile == numTiles - 1)
                {
                    const uint32_t all = prefix + total;
                    // a truncated candidate list may have lost pairs: what it wanted bounds what the pair list can want
                    const uint32_t wanted = cnt->numCandWanted > cnt->numCand ? max(all, cnt->numCandWanted) : all;
                    cnt->numPairsWanted = wanted; cnt->numPairs = min(all, maxPairs);
                    if (wanted > maxPairs) cnt->overflowPairs = 1;
                }
            }
        }
        __syncthreads();
        const uint32_t base = sPrefix + warpSums[warp];
#pragma unroll
        for (int j = 0; j < 4; ++j)
        {
            uint32_t pos = base + rank[j];
            if (keep[j] & 1u) { if (pos < maxPairs) pairs[pos] = make_uint2(c4[j].x, c4[j].y); ++pos; }
            if (keep[j] & 2u) { if (pos < maxPairs) pairs[pos] = make_uint2(c4[j].z, c4[j].w); }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        if (n == 0 && blockIdx.x == 0) { cnt->numPairsWanted = cnt->numCandWanted > cnt->numCand ? cnt->numCandWanted : 0u; cnt->numPairs = 0; }
        scanLeave(state, epoch);
    }
}

void launchBroadphase(Context* c)
{
    const uint32_t C = c->numColliders;
    DeviceCounts* cnt = c->counts;
    const float bigFactor = c->cfg.bigFactor > 0.f ? c->cfg.bigFactor : 3.f;
    const int G = gridFor(c, 8);
    static int cellDivEnv = -1;
    if (cellDivEnv < 0) { const char* e = getenv("B200PHYS_CELL_DIV"); cellDivEnv = e ? atoi(e) : 0; }     // tuning override
    const uint32_t cellDiv = cellDivEnv > 0 ? (uint32_t)cellDivEnv : c->cfg.reserved[1];
    const uint32_t numChunks = (C + kBigChunk - 1) / kBigChunk;
    static int tilesEnv = -1;
    if (tilesEnv < 0) { const char* e = getenv("B200PHYS_BP_TILES"); tilesEnv = e ? (atoi(e) ? 1 : 2) : 0; }    // tuning override: 1 always staged tiles, 0 never
    const uint32_t tilesMode = (uint32_t)tilesEnv;
    const bool reuse = c->bpReuse;
    static float maxFraction = -1.f, targetLife = -1.f;
    if (maxFraction < 0.f)
    {   // tuning overrides
        const char* e = getenv("B200PHYS_BP_MARGIN"); maxFraction = e ? (float)atof(e) : kMaxMarginFraction;
        const char* l = getenv("B200PHYS_BP_LIFE"); targetLife = l ? (float)atof(l) : kTargetLife;
    }
    if (C == 0)
    {
        if (reuse) { launchK(c, bpDecideKernel, 1, 32, 0, 1u, 0u, c->maxCand, maxFraction, targetLife, cnt); c->launches++; c->bpListColliders = 0; }
        launchK(c, bpPrimeFinishKernel, 1, 32, 0, bigFactor, c->cfg.cellSize, cnt);
        launchK(c, bpSetupKernel, 1, 32, 0, c->maxCells, bigFactor, c->cfg.cellSize, cellDiv, kWarpCand, tilesMode, cnt);
        c->launches += 2; c->bpPrimed = false;
        return;
    }
    // candidate reuse: a list exists for exactly this collider count and no upload happened since -> the device decides (check + decide);
    // otherwise this step builds (force)
    const uint32_t force = (!c->bpPrimed || c->bpListColliders != C) ? 1u : 0u;
    if (!c->bpPrimed)
    {   // no threshold from a previous step yet (first step, or colliders were uploaded)
        launchK(c, bpPrimeFinishKernel, 1, 32, 0, bigFactor, c->cfg.cellSize, cnt);       // clears the accumulators
        launchK(c, bpPrimeKernel, G, kThreads, 0, C, c->aabbMin, c->aabbMax, cnt);
        launchK(c, bpPrimeFinishKernel, 1, 32, 0, bigFactor, c->cfg.cellSize, cnt);
        c->launches += 3; c->bpPrimed = true;
    }
    // the pipeline's view of the boxes: grown copies with reuse, the world AABBs themselves without
    const float4* boxMin = reuse ? c->bpInflMin : c->aabbMin; const float4* boxMax = reuse ? c->bpInflMax : c->aabbMax;
    if (reuse)
    {
        launchK(c, bpCheckKernel, G, kThreads, 0, force, C, c->aabbMin, c->aabbMax, c->bpInflMin, c->bpInflMax, cnt);
        launchK(c, bpDecideKernel, 1, 32, 0, force, C, c->maxCand, maxFraction, targetLife, cnt);
        c->launches += 2; c->bpListColliders = C;
        timelineMark(c, "bp_check");
    }
    launchK(c, bpClassifyKernel, G, kThreads, 0, C, c->aabbMin, c->aabbMax, reuse ? c->bpInflMin : (float4*)nullptr, reuse ? c->bpInflMax : (float4*)nullptr, c->cellOf, c->bigList, cnt);
    launchK(c, bpSetupKernel, 1, 32, 0, c->maxCells, bigFactor, c->cfg.cellSize, cellDiv, kWarpCand, tilesMode, cnt);
    launchK(c, bpBinKernel, G, kThreads, 0, C, boxMin, boxMax, c->cellOf, c->rankInCell, c->cellCount, cnt);
    c->launches += 3;
    scanExclusiveU32(c, c->cellCount, c->cellStart, reuse ? &cnt->bpScanCells : &cnt->grid.numCells, c->maxCells, nullptr, /*clearInput*/ true);
    launchK(c, bpScatterKernel, G, kThreads, 0, C, boxMin, boxMax, c->cellOf, c->rankInCell, c->cellStart, c->bpRec, c->pairCount, cnt);
    timelineMark(c, "bp_build");
    static bool attr = false;
    if (!attr) { cudaFuncSetAttribute(bpTileKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(TileCtaShared)); attr = true; }
    const uint32_t tileCtas = (uint32_t)c->numSMs * 4u, bigCtas = (uint32_t)c->numSMs * 4u;      // (big rows: one 2048-collider chunk per CTA at 1 M colliders; with one CTA per SM the counting pass was a 0.1 ms latency chain)
    launchK(c, bpTileKernel, tileCtas + bigCtas, kTileThreads, sizeof(TileCtaShared), C, tileCtas, numChunks, c->bpRec, c->cellStart, boxMin, boxMax, c->bigList,
                                                                                      c->pairCount, c->bpRows, c->cellOf, c->bpBigChunks, cnt);
    launchK(c, bpRowsDirectKernel, gridFor(c, 6), kDirectThreads, 0, C, c->bpRec, c->cellStart, boxMin, boxMax, c->bigList, c->pairCount, c->bpRows, c->cellOf, cnt);
    c->launches += 3;
    timelineMark(c, "bp_tiles");
    scanExclusiveU32(c, c->pairCount, c->pairStart, reuse ? &cnt->bpScanColliders : (const uint32_t*)nullptr, C, nullptr);
    const uint32_t emitCtas = (uint32_t)c->numSMs * 6u, extraCtas = (uint32_t)c->numSMs * 4u;
    launchK(c, bpEmitKernel, emitCtas + extraCtas, kThreads, 0, C, emitCtas, numChunks, c->rankInCell, c->bpRows, c->pairCount, c->pairStart, c->bpRec, c->cellStart,
                                                                  boxMin, boxMax, c->bigList, c->cellOf, c->bpBigChunks, c->pairs, c->maxPairs, c->bpCand, c->maxCand, cnt);
    c->launches += 1;
    timelineMark(c, "bp_emit");
    if (reuse)
    {
        int blocks = gridFor(c, 4);
        launchK(c, bpFilterKernel, blocks, kFilterThreads, 0, c->bpCand, c->aabbMin, c->aabbMax, c->pairs, c->maxPairs, c->scanStatus, (ScanLbState*)c->scanState, cnt);
        c->launches += 1;
        timelineMark(c, "bp_filter");
    }
}

} // namespace b200
