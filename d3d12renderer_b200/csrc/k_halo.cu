// k_halo.cu -- multi-GPU ghost-body halo: pack / unpack kernels and their C-ABI entry points (DESIGN.md "Multi-GPU").
//
// Bodies are sharded by spatial slab, one context per GPU / rank. Before a step each rank packs the owned bodies (with all
// their colliders) whose world AABB reaches into a neighbour's slab (+ ghost width) into two dense device arrays; the
// transport is the caller's (NCCL send/recv over NVLink on exactly these device buffers, see d3d12renderer_b200/parallel.py);
// the receiver appends them behind its owned bodies / colliders as GHOSTS. Ghosts take part in the broad phase, narrow phase
// and solver like any body, but their post-step state is discarded: the owner's result wins and is re-sent next step.
// There is no collective in the data path beyond this neighbour exchange.
#include "b200_context.h"
#include "b200_util.cuh"

using namespace b200;

struct b200phys_context { Context c; };

namespace b200 {

static const int kThreads = 256;
static const uint32_t kHaloBodyFloats = 36;      // b200phys_body (35 floats) + owner-local body index
static const uint32_t kHaloColliderFloats = 16;  // b200phys_collider (64 bytes), objectIndex = index inside the packed body list

__global__ void __launch_bounds__(kThreads) haloFlagKernel(uint32_t numColliders, uint32_t numBodies, int axis, float lo, float hi,
    const float4* __restrict__ cLocal, const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax, uint32_t* __restrict__ bodyFlag)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < numColliders; i += gridDim.x * kThreads)
    {
        const float4 L3 = cLocal[4 * i + 3];
        const uint32_t tw = __float_as_uint(L3.y), objType = (tw >> 8) & 0xFFu, body = __float_as_uint(L3.z);
        if (objType != B200PHYS_OBJ_RIGID_BODY || body >= numBodies) continue;
        const float4 mn = aabbMin[i], mx = aabbMax[i];
        const float a = axis == 0 ? mn.x : (axis == 1 ? mn.y : mn.z), b = axis == 0 ? mx.x : (axis == 1 ? mx.y : mx.z);
        if (b >= lo && a < hi) bodyFlag[body] = 1u;      // benign race: every writer stores 1
    }
}

__global__ void __launch_bounds__(kThreads) haloColliderFlagKernel(uint32_t numColliders, uint32_t numBodies, const float4* __restrict__ cLocal,
    const uint32_t* __restrict__ bodyFlag, uint32_t* __restrict__ colFlag)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < numColliders; i += gridDim.x * kThreads)
    {
        const float4 L3 = cLocal[4 * i + 3];
        const uint32_t tw = __float_as_uint(L3.y), objType = (tw >> 8) & 0xFFu, body = __float_as_uint(L3.z);
        colFlag[i] = (objType == B200PHYS_OBJ_RIGID_BODY && body < numBodies && bodyFlag[body]) ? 1u : 0u;
    }
}

__global__ void __launch_bounds__(kThreads) haloPackBodiesKernel(uint32_t numBodies, uint32_t planeStride, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ slot, uint32_t cap,
    const float4* __restrict__ bRot, const float4* __restrict__ bPos, const float4* __restrict__ bCog, const float4* __restrict__ bInvI,
    const float4* __restrict__ bVel, const float4* __restrict__ bAng, const float4* __restrict__ bForce, const float4* __restrict__ bTorque, float* __restrict__ out)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < numBodies; i += gridDim.x * kThreads)
    {
        if (!flag[i]) continue;
        const uint32_t s = slot[i];
        if (s >= cap) continue;
        float* o = out + (size_t)s * kHaloBodyFloats;
        const float4 r = bRot[i], p = bPos[i], cg = bCog[i], c0 = bInvI[i], c1 = bInvI[planeStride + i], c2 = bInvI[2 * planeStride + i];
        const float4 v = bVel[i], w = bAng[i], F = bForce[i], T = bTorque[i];
        o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; o[4] = p.x; o[5] = p.y; o[6] = p.z; o[7] = cg.x; o[8] = cg.y; o[9] = cg.z; o[10] = p.w;
        o[11] = c0.x; o[12] = c0.y; o[13] = c0.z; o[14] = c1.x; o[15] = c1.y; o[16] = c1.z; o[17] = c2.x; o[18] = c2.y; o[19] = c2.z;
        o[20] = cg.w; o[21] = c0.w; o[22] = c1.w; o[23] = v.x; o[24] = v.y; o[25] = v.z; o[26] = w.x; o[27] = w.y; o[28] = w.z;
        o[29] = F.x; o[30] = F.y; o[31] = F.z; o[32] = T.x; o[33] = T.y; o[34] = T.z; o[35] = __uint_as_float(i);
    }
}

__global__ void __launch_bounds__(kThreads) haloPackCollidersKernel(uint32_t numColliders, const uint32_t* __restrict__ colFlag, const uint32_t* __restrict__ colSlot, uint32_t cap,
    const uint32_t* __restrict__ bodySlot, uint32_t capBodies, const float4* __restrict__ cLocal, float4* __restrict__ out)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < numColliders; i += gridDim.x * kThreads)
    {
        if (!colFlag[i]) continue;
        const uint32_t s = colSlot[i];
        if (s >= cap) continue;
        float4 L3 = cLocal[4 * i + 3];
        const uint32_t bs = bodySlot[__float_as_uint(L3.z)];        // body index inside the packed list
        if (bs >= capBodies)
        {   // its body did not fit into the message (overflow, flagged by the receiver through the header): send an inert slot instead of a
            // collider that would point past the packed body list
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            out[4 * s] = z; out[4 * s + 1] = z; out[4 * s + 2] = z;
            out[4 * s + 3] = make_float4(0.f, __uint_as_float((uint32_t)B200PHYS_SPHERE | ((uint32_t)B200PHYS_OBJ_NONE << 8)), __uint_as_float(0xFFFFFFFFu), __uint_as_float(B200PHYS_NO_TRANSFORM));
            continue;
        }
        L3.z = __uint_as_float(bs);
        out[4 * s] = cLocal[4 * i]; out[4 * s + 1] = cLocal[4 * i + 1]; out[4 * s + 2] = cLocal[4 * i + 2]; out[4 * s + 3] = L3;
    }
}

__global__ void __launch_bounds__(kThreads) haloUnpackBodiesKernel(uint32_t count, uint32_t base, uint32_t planeStride, const float* __restrict__ in,
    float4* __restrict__ bRot, float4* __restrict__ bPos, float4* __restrict__ bCog, float4* __restrict__ bInvI,
    float4* __restrict__ bVel, float4* __restrict__ bAng, float4* __restrict__ bForce, float4* __restrict__ bTorque)
{
    for (uint32_t k = blockIdx.x * kThreads + threadIdx.x; k < count; k += gridDim.x * kThreads)
    {
        const float* b = in + (size_t)k * kHaloBodyFloats; const uint32_t i = base + k;
        bRot[i] = make_float4(b[0], b[1], b[2], b[3]);
        bPos[i] = make_float4(b[4], b[5], b[6], b[10]);
        bCog[i] = make_float4(b[7], b[8], b[9], b[20]);
        bInvI[i] = make_float4(b[11], b[12], b[13], b[21]);
        bInvI[planeStride + i] = make_float4(b[14], b[15], b[16], b[22]);
        bInvI[2 * planeStride + i] = make_float4(b[17], b[18], b[19], 0.f);
        bVel[i] = make_float4(b[23], b[24], b[25], 0.f); bAng[i] = make_float4(b[26], b[27], b[28], 0.f);
        bForce[i] = make_float4(b[29], b[30], b[31], 0.f); bTorque[i] = make_float4(b[32], b[33], b[34], 0.f);
    }
}

__global__ void __launch_bounds__(kThreads) haloUnpackCollidersKernel(uint32_t count, uint32_t base, uint32_t bodyBase, const float4* __restrict__ in, float4* __restrict__ cLocal)
{
    for (uint32_t k = blockIdx.x * kThreads + threadIdx.x; k < count; k += gridDim.x * kThreads)
    {
        float4 L3 = in[4 * k + 3];
        L3.z = __uint_as_float(__float_as_uint(L3.z) + bodyBase);
        const uint32_t i = base + k;
        cLocal[4 * i] = in[4 * k]; cLocal[4 * i + 1] = in[4 * k + 1]; cLocal[4 * i + 2] = in[4 * k + 2]; cLocal[4 * i + 3] = L3;
    }
}

__global__ void haloHeaderKernel(const uint32_t* __restrict__ nBodies, const uint32_t* __restrict__ nColliders, uint32_t* __restrict__ header)
{
    if (threadIdx.x == 0) { header[0] = *nBodies; header[1] = *nColliders; header[2] = 0; header[3] = 0; }
}

// fixed-size message: slots below the header counts are ghosts, the rest inert padding (zero-mass body, OBJ_NONE collider)
__global__ void __launch_bounds__(kThreads) haloUnpackMessageKernel(const uint32_t* __restrict__ header, uint32_t capB, uint32_t capC, uint32_t bodyBase, uint32_t colBase, uint32_t planeStride,
    const float* __restrict__ inBodies, const float4* __restrict__ inCols,
    float4* __restrict__ bRot, float4* __restrict__ bPos, float4* __restrict__ bCog, float4* __restrict__ bInvI,
    float4* __restrict__ bVel, float4* __restrict__ bAng, float4* __restrict__ bForce, float4* __restrict__ bTorque, float4* __restrict__ cLocal, DeviceCounts* cnt)
{
    const uint32_t nB = min(header[0], capB), nC = min(header[1], capC);
    if (blockIdx.x == 0 && threadIdx.x == 0 && (header[0] > capB || header[1] > capC)) cnt->overflowHalo = 1;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (uint32_t k = blockIdx.x * kThreads + threadIdx.x; k < capB; k += gridDim.x * kThreads)
    {
        const uint32_t i = bodyBase + k;
        if (k < nB)
        {
            const float* b = inBodies + (size_t)k * kHaloBodyFloats;
            bRot[i] = make_float4(b[0], b[1], b[2], b[3]);
            bPos[i] = make_float4(b[4], b[5], b[6], b[10]);
            bCog[i] = make_float4(b[7], b[8], b[9], b[20]);
            bInvI[i] = make_float4(b[11], b[12], b[13], b[21]);
            bInvI[planeStride + i] = make_float4(b[14], b[15], b[16], b[22]);
            bInvI[2 * planeStride + i] = make_float4(b[17], b[18], b[19], 0.f);
            bVel[i] = make_float4(b[23], b[24], b[25], 0.f); bAng[i] = make_float4(b[26], b[27], b[28], 0.f);
            bForce[i] = make_float4(b[29], b[30], b[31], 0.f); bTorque[i] = make_float4(b[32], b[33], b[34], 0.f);
        }
        else
        {
            bRot[i] = make_float4(0.f, 0.f, 0.f, 1.f); bPos[i] = z; bCog[i] = z; bInvI[i] = z; bInvI[planeStride + i] = z; bInvI[2 * planeStride + i] = z;
            bVel[i] = z; bAng[i] = z; bForce[i] = z; bTorque[i] = z;
        }
    }
    for (uint32_t k = blockIdx.x * kThreads + threadIdx.x; k < capC; k += gridDim.x * kThreads)
    {
        const uint32_t i = colBase + k;
        if (k < nC)
        {
            float4 L3 = inCols[4 * k + 3];
            const uint32_t packedBody = __float_as_uint(L3.z);
            if (packedBody < nB)
            {
                L3.z = __uint_as_float(packedBody + bodyBase);
                cLocal[4 * i] = inCols[4 * k]; cLocal[4 * i + 1] = inCols[4 * k + 1]; cLocal[4 * i + 2] = inCols[4 * k + 2]; cLocal[4 * i + 3] = L3;
                continue;
            }
        }
        // padding slot, or a collider whose body is not in the message (sender-side overflow / inert slot)
        {
            cLocal[4 * i] = z; cLocal[4 * i + 1] = z; cLocal[4 * i + 2] = z;
            cLocal[4 * i + 3] = make_float4(0.f, __uint_as_float((uint32_t)B200PHYS_SPHERE | ((uint32_t)B200PHYS_OBJ_NONE << 8)), __uint_as_float(0xFFFFFFFFu), __uint_as_float(B200PHYS_NO_TRANSFORM));
        }
    }
}

// ------------------------------------------------------------------------------------------------------- migration
// dest[b]: 0 = stays, 1 = goes to the lower neighbour (position below lo), 2 = to the upper neighbour (position >= hi)
__global__ void __launch_bounds__(kThreads) migrateDestKernel(uint32_t numBodies, int axis, float lo, float hi, const float4* __restrict__ bPos, uint32_t* __restrict__ dest,
                                                              uint32_t* __restrict__ flagL, uint32_t* __restrict__ flagR, uint32_t* __restrict__ flagKeep)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < numBodies; i += gridDim.x * kThreads)
    {
        const float4 p = bPos[i];
        const float x = axis == 0 ? p.x : (axis == 1 ? p.y : p.z);
        const uint32_t d = x < lo ? 1u : (x >= hi ? 2u : 0u);
        dest[i] = d; flagL[i] = d == 1u; flagR[i] = d == 2u; flagKeep[i] = d == 0u;
    }
}
// colliders follow their body; everything that is not attached to an owned rigid body (static geometry) stays
__global__ void __launch_bounds__(kThreads) migrateColliderFlagKernel(uint32_t numColliders, uint32_t numBodies, const float4* __restrict__ cLocal, const uint32_t* __restrict__ dest,
                                                                      uint32_t* __restrict__ flagL, uint32_t* __restrict__ flagR, uint32_t* __restrict__ flagKeep)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < numColliders; i += gridDim.x * kThreads)
    {
        const float4 L3 = cLocal[4 * i + 3];
        const uint32_t tw = __float_as_uint(L3.y), objType = (tw >> 8) & 0xFFu, body = __float_as_uint(L3.z);
        const uint32_t d = (objType == B200PHYS_OBJ_RIGID_BODY && body < numBodies) ? dest[body] : 0u;
        flagL[i] = d == 1u; flagR[i] = d == 2u; flagKeep[i] = d == 0u;
    }
}
// out-of-place compaction of the ten body planes (plane p of the scratch starts at p * scratchStride)
__global__ void __launch_bounds__(kThreads) migrateGatherBodiesKernel(uint32_t numBodies, uint32_t planeStride, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ slot,
    const float4* __restrict__ bRot, const float4* __restrict__ bPos, const float4* __restrict__ bCog, const float4* __restrict__ bInvI,
    const float4* __restrict__ bVel, const float4* __restrict__ bAng, const float4* __restrict__ bForce, const float4* __restrict__ bTorque, float4* __restrict__ scratch, uint32_t scratchStride)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < numBodies; i += gridDim.x * kThreads)
    {
        if (!keep[i]) continue;
        const uint32_t s = slot[i];
        scratch[s] = bRot[i]; scratch[(size_t)scratchStride + s] = bPos[i]; scratch[(size_t)2 * scratchStride + s] = bCog[i];
        scratch[(size_t)3 * scratchStride + s] = bInvI[i]; scratch[(size_t)4 * scratchStride + s] = bInvI[planeStride + i]; scratch[(size_t)5 * scratchStride + s] = bInvI[2 * planeStride + i];
        scratch[(size_t)6 * scratchStride + s] = bVel[i]; scratch[(size_t)7 * scratchStride + s] = bAng[i];
        scratch[(size_t)8 * scratchStride + s] = bForce[i]; scratch[(size_t)9 * scratchStride + s] = bTorque[i];
    }
}
__global__ void __launch_bounds__(kThreads) migrateGatherCollidersKernel(uint32_t numColliders, uint32_t numBodies, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ slot,
    const uint32_t* __restrict__ bodySlot, const float4* __restrict__ cLocal, float4* __restrict__ scratch)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < numColliders; i += gridDim.x * kThreads)
    {
        if (!keep[i]) continue;
        const uint32_t s = slot[i];
        float4 L3 = cLocal[4 * i + 3];
        const uint32_t tw = __float_as_uint(L3.y), objType = (tw >> 8) & 0xFFu, body = __float_as_uint(L3.z);
        if (objType == B200PHYS_OBJ_RIGID_BODY && body < numBodies) L3.z = __uint_as_float(bodySlot[body]);   // index after compaction
        scratch[4 * s] = cLocal[4 * i]; scratch[4 * s + 1] = cLocal[4 * i + 1]; scratch[4 * s + 2] = cLocal[4 * i + 2]; scratch[4 * s + 3] = L3;
    }
}

} // namespace b200

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return -(int)e_; } while (0)

extern "C" {

uint32_t b200phys_halo_record_bytes(int which) { return which == 0 ? kHaloBodyFloats * 4u : kHaloColliderFloats * 4u; }

// Packs every OWNED body (and all of its colliders) with a collider AABB overlapping [lo, hi) on `axis` into the caller's
// device buffers (e.g. torch tensors); returns the counts through outCounts[2] = {bodies, colliders} (host, synchronous).
static int haloPackDevice(Context& c, int axis, float lo, float hi, void* dstBodies, void* dstColliders, uint32_t capBodies, uint32_t capColliders, uint32_t* header, uint32_t* hostCounts)
{
    waitPendingForceUpload(&c);     // the body records carry the accumulators
    const uint32_t N = c.numOwnedBodies, C = c.numOwnedColliders;
    // current world AABBs of the owned colliders (the step recomputes them, ghosts included)
    const uint32_t savedB = c.numBodies, savedC = c.numColliders;
    c.numBodies = N; c.numColliders = C;
    launchWorldColliders(&c);
    c.numBodies = savedB; c.numColliders = savedC;
    uint32_t* bodyFlag = c.adjFill;      // scratch (N + 2)
    uint32_t* bodySlot = c.adjStart;     // scratch (N + 2)
    uint32_t* colFlag = c.cellOf;        // scratch (C)
    uint32_t* colSlot = c.pairStart;     // scratch (C + 2)
    const int G = gridFor(&c, 4);
    CK(cudaMemsetAsync(bodyFlag, 0, sizeof(uint32_t) * (N + 1), c.stream));
    haloFlagKernel<<<G, kThreads, 0, c.stream>>>(C, N, axis, lo, hi, c.cLocal, c.aabbMin, c.aabbMax, bodyFlag);
    scanExclusiveU32(&c, bodyFlag, bodySlot, nullptr, N, nullptr);
    haloColliderFlagKernel<<<G, kThreads, 0, c.stream>>>(C, N, c.cLocal, bodyFlag, colFlag);
    scanExclusiveU32(&c, colFlag, colSlot, nullptr, C, nullptr);
    if (hostCounts)
    {
        CK(cudaMemcpyAsync(&hostCounts[0], bodySlot + N, 4, cudaMemcpyDeviceToHost, c.stream));
        CK(cudaMemcpyAsync(&hostCounts[1], colSlot + C, 4, cudaMemcpyDeviceToHost, c.stream));
    }
    if (header) haloHeaderKernel<<<1, 32, 0, c.stream>>>(bodySlot + N, colSlot + C, header);
    if (dstBodies && capBodies)
        haloPackBodiesKernel<<<G, kThreads, 0, c.stream>>>(N, c.cfg.maxBodies + 1, bodyFlag, bodySlot, capBodies, c.bRot, c.bPos, c.bCog, c.bInvI, c.bVel, c.bAng, c.bForce, c.bTorque, (float*)dstBodies);
    if (dstColliders && capColliders)
        haloPackCollidersKernel<<<G, kThreads, 0, c.stream>>>(C, colFlag, colSlot, capColliders, bodySlot, capBodies, c.cLocal, (float4*)dstColliders);
    CK(cudaGetLastError());
    return 0;
}

int b200phys_halo_pack(b200phys_context* ctx, int axis, float lo, float hi, void* dstBodies, void* dstColliders, uint32_t capBodies, uint32_t capColliders, uint32_t* outCounts)
{
    if (!ctx || axis < 0 || axis > 2 || !outCounts) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    outCounts[0] = outCounts[1] = 0;
    if (!c.numOwnedBodies || !c.numOwnedColliders) return 0;
    uint32_t counts[2] = { 0, 0 };
    int rc = haloPackDevice(c, axis, lo, hi, dstBodies, dstColliders, capBodies, capColliders, nullptr, counts);
    if (rc) return rc;
    CK(cudaStreamSynchronize(c.stream));
    outCounts[0] = counts[0]; outCounts[1] = counts[1];
    if (counts[0] > capBodies || counts[1] > capColliders) return B200PHYS_E_CAPACITY;
    return 0;
}

uint64_t b200phys_halo_message_bytes(uint32_t capBodies, uint32_t capColliders) { return 16ull + (uint64_t)capBodies * kHaloBodyFloats * 4ull + (uint64_t)capColliders * kHaloColliderFloats * 4ull; }

// Asynchronous on the context's stream; nothing returns to the host.
int b200phys_halo_pack_message(b200phys_context* ctx, int axis, float lo, float hi, void* dstMessageDev, uint32_t capBodies, uint32_t capColliders)
{
    if (!ctx || axis < 0 || axis > 2 || !dstMessageDev) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    char* base = (char*)dstMessageDev;
    if (!c.numOwnedBodies || !c.numOwnedColliders) { CK(cudaMemsetAsync(base, 0, 16, c.stream)); return 0; }
    return haloPackDevice(c, axis, lo, hi, base + 16, base + 16 + (size_t)capBodies * kHaloBodyFloats * 4, capBodies, capColliders, (uint32_t*)base, nullptr);
}

int b200phys_halo_unpack_message(b200phys_context* ctx, const void* srcMessageDev, uint32_t capBodies, uint32_t capColliders)
{
    if (!ctx || !srcMessageDev) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (c.numBodies + capBodies > c.cfg.maxBodies || c.numColliders + capColliders > c.cfg.maxColliders) return B200PHYS_E_CAPACITY;
    const char* base = (const char*)srcMessageDev;
    haloUnpackMessageKernel<<<gridFor(&c, 4), kThreads, 0, c.stream>>>((const uint32_t*)base, capBodies, capColliders, c.numBodies, c.numColliders, c.cfg.maxBodies + 1,
        (const float*)(base + 16), (const float4*)(base + 16 + (size_t)capBodies * kHaloBodyFloats * 4), c.bRot, c.bPos, c.bCog, c.bInvI, c.bVel, c.bAng, c.bForce, c.bTorque, c.cLocal, c.counts);
    c.numBodies += capBodies; c.numColliders += capColliders;
    CK(cudaGetLastError());
    return 0;
}

// Appends ghost bodies / colliders (device buffers in the halo record formats) behind the owned ones for the next step.
int b200phys_halo_unpack(b200phys_context* ctx, const void* srcBodies, uint32_t numBodies, const void* srcColliders, uint32_t numColliders)
{
    if (!ctx) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (c.numBodies + numBodies > c.cfg.maxBodies || c.numColliders + numColliders > c.cfg.maxColliders) return B200PHYS_E_CAPACITY;
    const int G = gridFor(&c, 4);
    if (numBodies) haloUnpackBodiesKernel<<<G, kThreads, 0, c.stream>>>(numBodies, c.numBodies, c.cfg.maxBodies + 1, (const float*)srcBodies, c.bRot, c.bPos, c.bCog, c.bInvI, c.bVel, c.bAng, c.bForce, c.bTorque);
    if (numColliders) haloUnpackCollidersKernel<<<G, kThreads, 0, c.stream>>>(numColliders, c.numColliders, c.numBodies, (const float4*)srcColliders, c.cLocal);
    c.numBodies += numBodies; c.numColliders += numColliders;
    CK(cudaGetLastError());
    return 0;
}

// Drops all ghosts: the context is back to its owned bodies / colliders.
int b200phys_halo_clear(b200phys_context* ctx)
{
    if (!ctx) return B200PHYS_E_INVALID;
    ctx->c.numBodies = ctx->c.numOwnedBodies; ctx->c.numColliders = ctx->c.numOwnedColliders;
    return 0;
}

// ---- migration: bodies whose position left [lo, hi) on `axis` move to the neighbouring rank with all their colliders.
// Synchronous (the owned counts change, the host must know them); call it every few steps, with no ghosts present.
// Messages have the halo message layout. outDest (host, numOwnedBodies entries, may be null) receives 0 / 1 / 2 per body so
// that the caller can carry its own per-body data (global ids) along; outCounts = {bodies down, colliders down, bodies up,
// colliders up}. Joints are not migrated: a context with joints returns B200PHYS_E_INVALID.
int b200phys_migrate_pack(b200phys_context* ctx, int axis, float lo, float hi, void* dstMessageDown, void* dstMessageUp, uint32_t capBodies, uint32_t capColliders,
                          uint32_t* outDest, uint32_t* outCounts)
{
    if (!ctx || axis < 0 || axis > 2 || !dstMessageDown || !dstMessageUp || !outCounts) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    for (int k = 0; k < 6; ++k) if (c.numJoints[k]) return B200PHYS_E_INVALID;
    if (c.numBodies != c.numOwnedBodies || c.numColliders != c.numOwnedColliders) return B200PHYS_E_INVALID;   // drop the ghosts first (b200phys_halo_clear)
    waitPendingForceUpload(&c);
    const uint32_t N = c.numOwnedBodies, C = c.numOwnedColliders;
    outCounts[0] = outCounts[1] = outCounts[2] = outCounts[3] = 0;
    char* msg[2] = { (char*)dstMessageDown, (char*)dstMessageUp };
    if (!N) { CK(cudaMemsetAsync(msg[0], 0, 16, c.stream)); CK(cudaMemsetAsync(msg[1], 0, 16, c.stream)); CK(cudaStreamSynchronize(c.stream)); return 0; }
    // scratch (all dead between steps): per-body words in the pair-sized buffers, plane scratch in the contact rows
    const size_t P = c.maxPairs;
    if ((size_t)4 * (N + 2) > P || (size_t)4 * (C + 2) > P || (size_t)10 * N > (size_t)8 * c.maxContacts || (size_t)4 * C > (size_t)8 * c.maxContacts) return B200PHYS_E_CAPACITY;
    uint32_t* dest = c.pairDst;                       // N
    uint32_t* flag[3] = { c.pairDst + (N + 2), c.pairDst + 2 * (size_t)(N + 2), c.pairDst + 3 * (size_t)(N + 2) };   // down, up, keep
    uint32_t* slot[3] = { (uint32_t*)c.manScan, (uint32_t*)c.manScan + (N + 2), (uint32_t*)c.manScan + 2 * (size_t)(N + 2) };
    uint32_t* cflag[3] = { (uint32_t*)c.manScanOut, (uint32_t*)c.manScanOut + (C + 2), (uint32_t*)c.manScanOut + 2 * (size_t)(C + 2) };
    uint32_t* cslot[3] = { (uint32_t*)c.pairsTyped, (uint32_t*)c.pairsTyped + (C + 2), (uint32_t*)c.pairsTyped + 2 * (size_t)(C + 2) };
    const int G = gridFor(&c, 4);
    const uint32_t planeStride = c.cfg.maxBodies + 1;
    migrateDestKernel<<<G, kThreads, 0, c.stream>>>(N, axis, lo, hi, c.bPos, dest, flag[0], flag[1], flag[2]);
    migrateColliderFlagKernel<<<G, kThreads, 0, c.stream>>>(C, N, c.cLocal, dest, cflag[0], cflag[1], cflag[2]);
    for (int d = 0; d < 3; ++d)
    {
        scanExclusiveU32(&c, flag[d], slot[d], nullptr, N, nullptr);
        if (C) scanExclusiveU32(&c, cflag[d], cslot[d], nullptr, C, nullptr); else CK(cudaMemsetAsync(cslot[d], 0, 4, c.stream));
    }
    uint32_t counts[6] = { 0, 0, 0, 0, 0, 0 };
    for (int d = 0; d < 3; ++d)
    {
        CK(cudaMemcpyAsync(&counts[2 * d], slot[d] + N, 4, cudaMemcpyDeviceToHost, c.stream));
        if (C) CK(cudaMemcpyAsync(&counts[2 * d + 1], cslot[d] + C, 4, cudaMemcpyDeviceToHost, c.stream));
    }
    for (int d = 0; d < 2; ++d)
    {
        float* outB = (float*)(msg[d] + 16); float4* outC = (float4*)(msg[d] + 16 + (size_t)capBodies * kHaloBodyFloats * 4);
        haloPackBodiesKernel<<<G, kThreads, 0, c.stream>>>(N, planeStride, flag[d], slot[d], capBodies, c.bRot, c.bPos, c.bCog, c.bInvI, c.bVel, c.bAng, c.bForce, c.bTorque, outB);
        if (C) haloPackCollidersKernel<<<G, kThreads, 0, c.stream>>>(C, cflag[d], cslot[d], capColliders, slot[d], capBodies, c.cLocal, outC);
        haloHeaderKernel<<<1, 32, 0, c.stream>>>(slot[d] + N, cslot[d] + C, (uint32_t*)msg[d]);
    }
    // compaction of what stays
    float4* scratch = c.rows;
    migrateGatherBodiesKernel<<<G, kThreads, 0, c.stream>>>(N, planeStride, flag[2], slot[2], c.bRot, c.bPos, c.bCog, c.bInvI, c.bVel, c.bAng, c.bForce, c.bTorque, scratch, N);
    CK(cudaStreamSynchronize(c.stream));
    CK(cudaGetLastError());
    const uint32_t keptB = counts[4], keptC = counts[5];
    if (counts[0] > capBodies || counts[2] > capBodies || counts[1] > capColliders || counts[3] > capColliders) return B200PHYS_E_CAPACITY;   // nothing was modified yet
    if (outDest) CK(cudaMemcpyAsync(outDest, dest, sizeof(uint32_t) * N, cudaMemcpyDeviceToHost, c.stream));
    float4* planes[10] = { c.bRot, c.bPos, c.bCog, c.bInvI, c.bInvI + planeStride, c.bInvI + 2 * (size_t)planeStride, c.bVel, c.bAng, c.bForce, c.bTorque };
    for (int p = 0; p < 10; ++p) if (keptB) CK(cudaMemcpyAsync(planes[p], scratch + (size_t)p * N, sizeof(float4) * keptB, cudaMemcpyDeviceToDevice, c.stream));
    if (C)
    {
        CK(cudaStreamSynchronize(c.stream));          // the plane scratch is reused for the colliders
        migrateGatherCollidersKernel<<<G, kThreads, 0, c.stream>>>(C, N, cflag[2], cslot[2], slot[2], c.cLocal, scratch);
        if (keptC) CK(cudaMemcpyAsync(c.cLocal, scratch, sizeof(float4) * 4 * (size_t)keptC, cudaMemcpyDeviceToDevice, c.stream));
    }
    CK(cudaStreamSynchronize(c.stream));
    CK(cudaGetLastError());
    c.numBodies = c.numOwnedBodies = keptB; c.numColliders = c.numOwnedColliders = keptC; ++c.bodiesEpoch;
    outCounts[0] = counts[0]; outCounts[1] = counts[1]; outCounts[2] = counts[2]; outCounts[3] = counts[3];
    return 0;
}

// Appends the bodies / colliders of a migration message as OWNED; returns the counts through outCounts[2] (synchronous).
int b200phys_migrate_unpack(b200phys_context* ctx, const void* srcMessageDev, uint32_t capBodies, uint32_t capColliders, uint32_t* outCounts)
{
    if (!ctx || !srcMessageDev || !outCounts) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (c.numBodies != c.numOwnedBodies || c.numColliders != c.numOwnedColliders) return B200PHYS_E_INVALID;
    uint32_t hdr[4] = { 0, 0, 0, 0 };
    CK(cudaMemcpyAsync(hdr, srcMessageDev, 16, cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    outCounts[0] = hdr[0]; outCounts[1] = hdr[1];
    if (hdr[0] > capBodies || hdr[1] > capColliders) return B200PHYS_E_CAPACITY;
    const char* base = (const char*)srcMessageDev;
    int rc = b200phys_halo_unpack(ctx, base + 16, hdr[0], base + 16 + (size_t)capBodies * kHaloBodyFloats * 4, hdr[1]);
    if (rc) return rc;
    c.numOwnedBodies = c.numBodies; c.numOwnedColliders = c.numColliders;
    return 0;
}

} // extern "C"
