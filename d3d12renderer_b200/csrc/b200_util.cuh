// b200_util.cuh -- device helpers shared by the kernels: cache-hinted 128-bit accesses, warp helpers, grid barrier.
#pragma once
#include "b200_math.cuh"

namespace b200 {

// streaming 128-bit load that does not allocate in L1 (data read once per kernel)
B200_D float4 ldStream(const float4* p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}
// coherent (L2) 128-bit load / store for data other CTAs write inside the same kernel (solver body records)
B200_D float4 ldCg(const float4* p)
{
    float4 r;
    asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
    return r;
}
B200_D void stCg(float4* p, float4 v)
{
    asm volatile("st.global.cg.v4.f32 [%0], {%1,%2,%3,%4};" :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
B200_D uint32_t ldCgU32(const uint32_t* p)
{
    uint32_t r; asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory"); return r;
}
B200_D uint32_t ldAcquireU32(const uint32_t* p)
{
    uint32_t r; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(r) : "l"(p) : "memory"); return r;
}

B200_D uint32_t laneId() { return threadIdx.x & 31u; }
B200_D uint32_t laneMaskLt() { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// float <-> order-preserving uint (for atomicMin/Max on signed floats)
B200_D uint32_t floatToSortable(float f) { uint32_t u = __float_as_uint(f); return (u & 0x80000000u) ? ~u : (u | 0x80000000u); }
B200_D float sortableToFloat(uint32_t u) { u = (u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u; return __uint_as_float(u); }

// ---- single-pass (decoupled look-back) scan protocol shared by scanLookbackKernel (b200_util.cu) and the broad phase's candidate
// filter (k_broad.cu). Tiles are handed out by an atomic ticket; a tile publishes (epoch, state, value) in one 64-bit status word; the
// epoch (a per-context launch counter kept on the device so that a replayed CUDA graph sees a fresh one) makes clearing unnecessary.
struct ScanLbState { unsigned int ticket, done, epoch, pad; };
// the epoch of this launch (1 .. 2^30-1): read before the CTA takes a ticket, advanced by the last CTA to leave (nobody of this launch
// can read it afterwards)
B200_D uint32_t scanEpochOf(const ScanLbState* state) { uint32_t e; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(e) : "l"(&state->epoch) : "memory"); return e + 1u; }
B200_D void scanLeave(ScanLbState* state, uint32_t epoch)
{
    if (atomicAdd(&state->done, 1u) == gridDim.x - 1) { state->ticket = 0; state->done = 0; state->epoch = (epoch >= 0x3FFFFFFEu) ? 0u : epoch; }     // last one out: ready for the next launch
}
// Called by the 32 threads of warp 0 (threadIdx.x < 32) of the CTA that owns `tile`, with the tile's total: publishes the aggregate,
// looks back 32 predecessors at a time until an inclusive prefix is found, publishes the tile's inclusive prefix and returns the
// EXCLUSIVE prefix of the tile (uniform across the warp). 32-bit sums.
B200_D uint32_t scanLookbackPrefix(unsigned long long* __restrict__ status, uint32_t tile, uint32_t total, uint32_t epoch)
{
    const unsigned long long tagAgg = ((unsigned long long)epoch << 34) | (1ull << 32), tagPre = ((unsigned long long)epoch << 34) | (2ull << 32);
    uint32_t prefix = 0;
    if (tile == 0) { if (threadIdx.x == 0) atomicExch(&status[0], tagPre | total); return 0u; }
    if (threadIdx.x == 0) atomicExch(&status[tile], tagAgg | total);
    int32_t window = (int32_t)tile - 1;
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
    return prefix;
}

} // namespace b200
