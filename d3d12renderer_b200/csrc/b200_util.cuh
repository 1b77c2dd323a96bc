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

} // namespace b200
