// b200_util.cu -- device-wide primitives used by the broad phase, narrow phase and solver set-up:
//   * exclusive scan (u32 / u64) over a DEVICE-side element count: reduce -> spine -> apply, contiguous chunk per CTA
//   * stable counting sort for small key ranges (<= 64 keys): contiguous chunk per WARP, __match_any_sync ranking
// Both are hand-written (no CUB/thrust) and sized as multiples of the SM count.
#include "b200_context.h"
#include "b200_util.cuh"

namespace b200 {

static const int kScanThreads = 256;
static const int kScanItems = 4;                 // per thread per tile
static const int kScanTile = kScanThreads * kScanItems;

B200_D uint32_t chunkSize(uint32_t n, uint32_t parts, uint32_t align)
{
    uint32_t c = (n + parts - 1) / parts;
    return ((c + align - 1) / align) * align;
}

template <typename T>
B200_D T blockReduceSum(T v, T* smem)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    if (laneId() == 0) smem[threadIdx.x >> 5] = v;
    __syncthreads();
    T r = 0;
    if (threadIdx.x < 32)
    {
        r = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : (T)0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
    }
    __syncthreads();
    return r; // valid in thread 0
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scanReduceKernel(const T* __restrict__ in, const uint32_t* __restrict__ devCount, uint32_t hostCount, T* __restrict__ blockSums)
{
    __shared__ T red[kScanThreads / 32];
    const uint32_t n = devCount ? *devCount : hostCount;
    const uint32_t chunk = chunkSize(n, gridDim.x, kScanTile);
    const uint32_t begin = min(n, blockIdx.x * chunk), end = min(n, begin + chunk);
    T sum = 0;
    for (uint32_t i = begin + threadIdx.x; i < end; i += kScanThreads) sum += in[i];
    sum = blockReduceSum(sum, red);
    if (threadIdx.x == 0) blockSums[blockIdx.x] = sum;
}

template <typename T>
__global__ void __launch_bounds__(1024) scanSpineKernel(T* __restrict__ blockSums, uint32_t numBlocks, T* __restrict__ totalOut)
{
    // numBlocks <= 2048: two elements per thread, Hillis-Steele in shared memory
    __shared__ T s[2048];
    for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = (i < numBlocks) ? blockSums[i] : (T)0;
    __syncthreads();
    for (uint32_t off = 1; off < 2048; off <<= 1)
    {
        T a0 = 0, a1 = 0; uint32_t i0 = threadIdx.x, i1 = threadIdx.x + 1024;
        if (i0 >= off) a0 = s[i0 - off];
        if (i1 >= off) a1 = s[i1 - off];
        __syncthreads();
        s[i0] += a0; s[i1] += a1;
        __syncthreads();
    }
    for (uint32_t i = threadIdx.x; i < numBlocks; i += blockDim.x) blockSums[i] = (i == 0) ? (T)0 : s[i - 1];
    if (threadIdx.x == 0 && totalOut) *totalOut = s[numBlocks - 1];
}

template <typename T>
__global__ void __launch_bounds__(kScanThreads) scanApplyKernel(T* __restrict__ in, T* __restrict__ out, const uint32_t* __restrict__ devCount, uint32_t hostCount,
                                                                const T* __restrict__ blockSums, const T* __restrict__ total, bool clearInput)
{
    __shared__ T warpSums[kScanThreads / 32];
    __shared__ T carryS;
    const uint32_t n = devCount ? *devCount : hostCount;
    const uint32_t chunk = chunkSize(n, gridDim.x, kScanTile);
    const uint32_t begin = min(n, blockIdx.x * chunk), end = min(n, begin + chunk);
    T carry = blockSums[blockIdx.x];
    for (uint32_t tile = begin; tile < end; tile += kScanTile)
    {
        const uint32_t base = tile + threadIdx.x * kScanItems;
        T v[kScanItems]; T local = 0;
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) { v[k] = (base + k < end) ? in[base + k] : (T)0; local += v[k]; if (clearInput && base + k < end) in[base + k] = (T)0; }
        // warp inclusive scan of the per-thread sums
        T incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xffffffffu, incl, o); if (laneId() >= (uint32_t)o) incl += t; }
        if (laneId() == 31) warpSums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            T w = (threadIdx.x < kScanThreads / 32) ? warpSums[threadIdx.x] : (T)0;
            T wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { T t = __shfl_up_sync(0xffffffffu, wi, o); if (laneId() >= (uint32_t)o) wi += t; }
            if (threadIdx.x < kScanThreads / 32) warpSums[threadIdx.x] = wi - w;   // exclusive warp offsets
            if (threadIdx.x == kScanThreads / 32 - 1) carryS = wi;                 // tile total
        }
        __syncthreads();
        T excl = carry + warpSums[threadIdx.x >> 5] + (incl - local);
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) { if (base + k < end) out[base + k] = excl; excl += v[k]; }
        carry += carryS;
        __syncthreads();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0 && total) out[n] = *total;   // out has n + 1 entries
}

template <typename T>
static void scanExclusive(Context* c, const T* in, T* out, const uint32_t* devCount, uint32_t maxCount, T* tmp, T* totalOut, bool clearInput = false)
{
    if (maxCount == 0) return;
    int blocks = gridFor(c, 8);
    if (blocks > 2048) blocks = 2048;
    uint32_t maxBlocks = (maxCount + kScanTile - 1) / kScanTile;
    if ((uint32_t)blocks > maxBlocks) blocks = (int)maxBlocks;
    if (blocks < 1) blocks = 1;
    T* total = tmp + 2048;
    launchK(c, scanReduceKernel<T>, blocks, kScanThreads, 0, in, devCount, maxCount, tmp);
    launchK(c, scanSpineKernel<T>, 1, 1024, 0, tmp, (uint32_t)blocks, total);
    launchK(c, scanApplyKernel<T>, blocks, kScanThreads, 0, const_cast<T*>(in), out, devCount, maxCount, tmp, total, clearInput);
    if (totalOut) memcpyD2DK(c, totalOut, total, sizeof(T));
    c->launches += 3;
}

// ---- single-pass exclusive scan (decoupled look-back): ONE launch instead of reduce / spine / apply. Tiles are handed out by an
// atomic ticket so that every tile a CTA looks back at is already owned by a running CTA; a tile publishes (epoch, state, value)
// in one 64-bit word, the epoch (a per-context launch counter) makes clearing the status array unnecessary; the last CTA to leave
// resets the ticket counter for the next launch.
static const int kLbThreads = 256, kLbItems = 8, kLbTile = kLbThreads * kLbItems;
// (ScanLbState, scanEpochOf, scanLeave: b200_util.cuh -- the broad phase's candidate filter runs the same look-back protocol)

__global__ void __launch_bounds__(kLbThreads) scanLookbackKernel(uint32_t* __restrict__ in, uint32_t* __restrict__ out, const uint32_t* __restrict__ devCount, uint32_t hostCount,
                                                                 unsigned long long* __restrict__ status, ScanLbState* __restrict__ state, uint32_t* __restrict__ totalOut, bool clearInput)
{
    const uint32_t epoch = scanEpochOf(state);
    __shared__ uint32_t warpSums[kLbThreads / 32];
    __shared__ uint32_t sTile, sPrefix, sTileTotal;
    const uint32_t n = devCount ? *devCount : hostCount;
    const uint32_t numTiles = (n + kLbTile - 1) / kLbTile;
    const unsigned long long tagAgg = ((unsigned long long)epoch << 34) | (1ull << 32), tagPre = ((unsigned long long)epoch << 34) | (2ull << 32);
    for (;;)
    {
        if (threadIdx.x == 0) sTile = atomicAdd(&state->ticket, 1u);
        __syncthreads();
        const uint32_t tile = sTile;
        if (tile >= numTiles) break;
        const uint32_t base = tile * kLbTile + threadIdx.x * kLbItems;
        uint32_t v[kLbItems]; uint32_t local = 0;
        if (base + kLbItems <= n)
        {
            const uint4 a = *reinterpret_cast<const uint4*>(in + base), b = *reinterpret_cast<const uint4*>(in + base + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
            if (clearInput) { const uint4 z = make_uint4(0, 0, 0, 0); *reinterpret_cast<uint4*>(in + base) = z; *reinterpret_cast<uint4*>(in + base + 4) = z; }
        }
        else
        {
#pragma unroll
            for (int k = 0; k < kLbItems; ++k) { v[k] = (base + k < n) ? in[base + k] : 0u; if (clearInput && base + k < n) in[base + k] = 0u; }
        }
#pragma unroll
        for (int k = 0; k < kLbItems; ++k) local += v[k];
        uint32_t incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if (laneId() >= (uint32_t)o) incl += t; }
        if (laneId() == 31) warpSums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            const uint32_t w = (threadIdx.x < kLbThreads / 32) ? warpSums[threadIdx.x] : 0u;
            uint32_t wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, wi, o); if (laneId() >= (uint32_t)o) wi += t; }
            if (threadIdx.x < kLbThreads / 32) warpSums[threadIdx.x] = wi - w;
            if (threadIdx.x == kLbThreads / 32 - 1) sTileTotal = wi;
        }
        __syncthreads();
        if (threadIdx.x < 32)
        {   // warp 0: publish the tile aggregate, then look back 32 predecessors at a time until an inclusive prefix is found
            const uint32_t total = sTileTotal;
            uint32_t prefix = 0;
            if (tile == 0) { if (threadIdx.x == 0) atomicExch(&status[0], tagPre | total); }
            else
            {
                if (threadIdx.x == 0) atomicExch(&status[tile], tagAgg | total);
                int32_t window = (int32_t)tile - 1;                            // nearest predecessor of this window
                for (;;)
                {
                    const int32_t p = window - (int32_t)threadIdx.x;
                    unsigned long long st = tagPre;                            // lanes before tile 0 act as an empty prefix
                    if (p >= 0)
                        do { asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(st) : "l"(status + p) : "memory"); } while ((st >> 34) != epoch || ((st >> 32) & 3ull) == 0ull);
                    const bool isPrefix = ((st >> 32) & 3ull) == 2ull;
                    const uint32_t prefMask = __ballot_sync(0xffffffffu, isPrefix);
                    const uint32_t first = prefMask ? (uint32_t)__ffs((int)prefMask) - 1u : 32u;     // nearest lane holding a prefix
                    uint32_t v = (threadIdx.x <= first) ? (uint32_t)st : 0u;
                    v = __reduce_add_sync(0xffffffffu, v);
                    prefix += v;
                    if (prefMask) break;
                    window -= 32;
                }
                if (threadIdx.x == 0) atomicExch(&status[tile], tagPre | (prefix + total));
            }
            if (threadIdx.x == 0)
            {
                sPrefix = prefix;
                if (tile == numTiles - 1) { out[n] = prefix + total; if (totalOut) *totalOut = prefix + total; }
            }
        }
        __syncthreads();
        uint32_t excl = sPrefix + warpSums[threadIdx.x >> 5] + (incl - local);
        if (base + kLbItems <= n)
        {
            uint4 a, b;
            a.x = excl; excl += v[0]; a.y = excl; excl += v[1]; a.z = excl; excl += v[2]; a.w = excl; excl += v[3];
            b.x = excl; excl += v[4]; b.y = excl; excl += v[5]; b.z = excl; excl += v[6]; b.w = excl;
            *reinterpret_cast<uint4*>(out + base) = a; *reinterpret_cast<uint4*>(out + base + 4) = b;
        }
        else
        {
#pragma unroll
            for (int k = 0; k < kLbItems; ++k) { if (base + k < n) out[base + k] = excl; excl += v[k]; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        if (n == 0 && blockIdx.x == 0) { out[0] = 0; if (totalOut) *totalOut = 0; }
        scanLeave(state, epoch);
    }
}

void scanExclusiveU32(Context* c, const uint32_t* in, uint32_t* out, const uint32_t* devCount, uint32_t maxCount, uint32_t* totalOut, bool clearInput)
{
    if (maxCount == 0) return;
    const uint32_t maxTiles = (maxCount + kLbTile - 1) / kLbTile;
    if (maxTiles > c->scanStatusTiles || (((uintptr_t)in | (uintptr_t)out) & 15u))
    {   // unaligned buffers or more tiles than the status array holds: the three-kernel version
        scanExclusive<uint32_t>(c, in, out, devCount, maxCount, c->scanTmp, totalOut, clearInput);
        return;
    }
    int blocks = gridFor(c, 4);
    if ((uint32_t)blocks > maxTiles) blocks = (int)maxTiles;
    launchK(c, scanLookbackKernel, blocks, kLbThreads, 0, const_cast<uint32_t*>(in), out, devCount, maxCount, c->scanStatus, (ScanLbState*)c->scanState, totalOut, clearInput);
    c->launches += 1;
}
// 64-bit variant (two packed 32-bit sums per element): the tile status is two self-validating words, (tag | low half) and
// (tag | high half); a reader accepts a tile only when both words carry the same tag
static const int kLb64Items = 4, kLb64Tile = kLbThreads * kLb64Items;
__global__ void __launch_bounds__(kLbThreads) scanLookbackKernel64(const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out, const uint32_t* __restrict__ devCount, uint32_t hostCount,
                                                                   unsigned long long* __restrict__ status /* 2 words per tile */, ScanLbState* __restrict__ state)
{
    const uint32_t epoch = scanEpochOf(state);
    __shared__ unsigned long long warpSums[kLbThreads / 32];
    __shared__ unsigned long long sPrefix, sTileTotal; __shared__ uint32_t sTile;
    const uint32_t n = devCount ? *devCount : hostCount;
    const uint32_t numTiles = (n + kLb64Tile - 1) / kLb64Tile;
    const unsigned long long tagAgg = ((unsigned long long)epoch << 34) | (1ull << 32), tagPre = ((unsigned long long)epoch << 34) | (2ull << 32);
    for (;;)
    {
        if (threadIdx.x == 0) sTile = atomicAdd(&state->ticket, 1u);
        __syncthreads();
        const uint32_t tile = sTile;
        if (tile >= numTiles) break;
        const uint32_t base = tile * kLb64Tile + threadIdx.x * kLb64Items;
        unsigned long long v[kLb64Items]; unsigned long long local = 0;
#pragma unroll
        for (int k = 0; k < kLb64Items; ++k) { v[k] = (base + k < n) ? in[base + k] : 0ull; local += v[k]; }
        unsigned long long incl = local;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o); if (laneId() >= (uint32_t)o) incl += t; }
        if (laneId() == 31) warpSums[threadIdx.x >> 5] = incl;
        __syncthreads();
        if (threadIdx.x < 32)
        {
            const unsigned long long w = (threadIdx.x < kLbThreads / 32) ? warpSums[threadIdx.x] : 0ull;
            unsigned long long wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned long long t = __shfl_up_sync(0xffffffffu, wi, o); if (laneId() >= (uint32_t)o) wi += t; }
            if (threadIdx.x < kLbThreads / 32) warpSums[threadIdx.x] = wi - w;
            if (threadIdx.x == kLbThreads / 32 - 1) sTileTotal = wi;
        }
        __syncthreads();
        if (threadIdx.x < 32)
        {
            const unsigned long long total = sTileTotal;
            unsigned long long prefix = 0;
            auto publish = [&](unsigned long long tag, unsigned long long val)
            { atomicExch(&status[2 * tile], tag | (val & 0xFFFFFFFFull)); atomicExch(&status[2 * tile + 1], tag | (val >> 32)); };
            if (tile == 0) { if (threadIdx.x == 0) publish(tagPre, total); }
            else
            {
                if (threadIdx.x == 0) publish(tagAgg, total);
                int32_t window = (int32_t)tile - 1;
                for (;;)
                {
                    const int32_t p = window - (int32_t)threadIdx.x;
                    unsigned long long lo = tagPre, hi = tagPre;
                    if (p >= 0)
                        do
                        {
                            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(lo) : "l"(status + 2 * p) : "memory");
                            asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(hi) : "l"(status + 2 * p + 1) : "memory");
                        } while ((lo >> 34) != epoch || ((lo >> 32) & 3ull) == 0ull || (lo >> 32) != (hi >> 32));
                    const bool isPrefix = ((lo >> 32) & 3ull) == 2ull;
                    const uint32_t prefMask = __ballot_sync(0xffffffffu, isPrefix);
                    const uint32_t first = prefMask ? (uint32_t)__ffs((int)prefMask) - 1u : 32u;
                    unsigned long long val = (threadIdx.x <= first) ? ((lo & 0xFFFFFFFFull) | (hi << 32)) : 0ull;
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1) val += __shfl_down_sync(0xffffffffu, val, o);
                    val = __shfl_sync(0xffffffffu, val, 0);
                    prefix += val;
                    if (prefMask) break;
                    window -= 32;
                }
                if (threadIdx.x == 0) publish(tagPre, prefix + total);
            }
            if (threadIdx.x == 0) { sPrefix = prefix; if (tile == numTiles - 1) out[n] = prefix + total; }
        }
        __syncthreads();
        unsigned long long excl = sPrefix + warpSums[threadIdx.x >> 5] + (incl - local);
#pragma unroll
        for (int k = 0; k < kLb64Items; ++k) { if (base + k < n) out[base + k] = excl; excl += v[k]; }
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        if (n == 0 && blockIdx.x == 0) out[0] = 0ull;
        scanLeave(state, epoch);
    }
}

void scanExclusiveU64(Context* c, const unsigned long long* in, unsigned long long* out, const uint32_t* devCount, uint32_t maxCount, unsigned long long* totalOut)
{
    if (maxCount == 0) return;
    const uint32_t maxTiles = (maxCount + kLb64Tile - 1) / kLb64Tile;
    if (2 * (size_t)maxTiles > c->scanStatusTiles || totalOut)
    {
        scanExclusive<unsigned long long>(c, in, out, devCount, maxCount, c->scanTmp64, totalOut);
        return;
    }
    int blocks = gridFor(c, 4);
    if ((uint32_t)blocks > maxTiles) blocks = (int)maxTiles;
    launchK(c, scanLookbackKernel64, blocks, kLbThreads, 0, in, out, devCount, maxCount, c->scanStatus, (ScanLbState*)c->scanState);
    c->launches += 1;
}

// ------------------------------------------------------------------------------------------ stable small-key sort
static const int kSortThreads = 256;
static const int kSortWarps = kSortThreads / 32;

__global__ void __launch_bounds__(kSortThreads) sksHistKernel(const uint8_t* __restrict__ keys, const uint32_t* __restrict__ devCount, uint32_t numKeys,
                                                              uint32_t* __restrict__ hist /* [numKeys][totalWarps] */)
{
    __shared__ uint32_t h[kSortWarps][64];
    const uint32_t n = *devCount;
    const uint32_t warp = threadIdx.x >> 5, lane = laneId();
    const uint32_t totalWarps = gridDim.x * kSortWarps, gw = blockIdx.x * kSortWarps + warp;
    for (uint32_t k = lane; k < 64; k += 32) h[warp][k] = 0;
    __syncwarp();
    const uint32_t chunk = chunkSize(n, totalWarps, 32);
    const uint32_t begin = min(n, gw * chunk), end = min(n, begin + chunk);
    for (uint32_t g = begin; g < end; g += 32)       // warp-uniform trip count
    {
        const uint32_t i = g + lane;
        const bool valid = i < end;
        const uint32_t key = valid ? keys[i] : 63u;
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? key : 0x100u);
        if (valid && (peers & laneMaskLt()) == 0) h[warp][key] += __popc(peers);
        __syncwarp();
    }
    __syncwarp();
    for (uint32_t k = lane; k < numKeys; k += 32) hist[k * totalWarps + gw] = h[warp][k];
}

__global__ void __launch_bounds__(kSortThreads) sksScatterKernel(const uint8_t* __restrict__ keys, const uint32_t* __restrict__ devCount, uint32_t numKeys,
                                                                 const uint32_t* __restrict__ histScan, uint32_t* __restrict__ dst)
{
    __shared__ uint32_t base[kSortWarps][64];
    const uint32_t n = *devCount;
    const uint32_t warp = threadIdx.x >> 5, lane = laneId();
    const uint32_t totalWarps = gridDim.x * kSortWarps, gw = blockIdx.x * kSortWarps + warp;
    for (uint32_t k = lane; k < numKeys; k += 32) base[warp][k] = histScan[k * totalWarps + gw];
    __syncwarp();
    const uint32_t chunk = chunkSize(n, totalWarps, 32);
    const uint32_t begin = min(n, gw * chunk), end = min(n, begin + chunk);
    for (uint32_t g = begin; g < end; g += 32)       // warp-uniform trip count
    {
        const uint32_t i = g + lane;
        const bool valid = i < end;
        const uint32_t key = valid ? keys[i] : 63u;
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? key : 0x100u);
        uint32_t b = 0;
        if (valid) b = base[warp][key];
        __syncwarp();
        if (valid)
        {
            dst[i] = b + __popc(peers & laneMaskLt());
            if ((peers & laneMaskLt()) == 0) base[warp][key] = b + __popc(peers);
        }
        __syncwarp();
    }
}

__global__ void sksKeyStartKernel(const uint32_t* __restrict__ histScan, uint32_t numKeys, uint32_t totalWarps, const uint32_t* __restrict__ devCount, uint32_t* __restrict__ keyStart)
{
    uint32_t k = threadIdx.x;
    if (k < numKeys) keyStart[k] = histScan[k * totalWarps];
    if (k == numKeys) keyStart[k] = *devCount;
}

void stableSmallKeySort(Context* c, const uint8_t* keys, const uint32_t* devCount, uint32_t maxCount, uint32_t numKeys, uint32_t* dst, uint32_t* keyStart)
{
    if (maxCount == 0) return;
    const int blocks = gridFor(c, 4);
    const uint32_t totalWarps = (uint32_t)blocks * kSortWarps;
    uint32_t* hist = c->sortHist;                       // numKeys * totalWarps (+1)
    uint32_t* histScan = hist + 64 * totalWarps + 64;
    launchK(c, sksHistKernel, blocks, kSortThreads, 0, keys, devCount, numKeys, hist);
    c->launches += 1;
    scanExclusiveU32(c, hist, histScan, nullptr, numKeys * totalWarps, nullptr);
    launchK(c, sksScatterKernel, blocks, kSortThreads, 0, keys, devCount, numKeys, histScan, dst);
    launchK(c, sksKeyStartKernel, 1, 64, 0, histScan, numKeys, totalWarps, devCount, keyStart);
    c->launches += 2;
}

} // namespace b200
