// k_dist.cu -- multi-GPU entry points of the C ABI that own their transport: one context per GPU / rank, bodies sharded by spatial
// slab, ghost-body halo exchanged with the slab neighbours by NCCL point-to-point operations (ncclSend / ncclRecv inside one group,
// on the context's stream, over NVLink / NVSwitch). The C++ host drives a multi-GPU physicsStep() through these without Python:
// rank 0 calls b200phys_dist_unique_id, the host distributes the 128 bytes (MPI_Bcast, a file, a socket), every rank calls
// b200phys_create_distributed, then b200phys_step_distributed per step. There is no collective on the data path: exactly two
// neighbours per rank and step. NCCL is loaded at run time (dlopen "libnccl.so.2"), so single-GPU users need no NCCL at all.
#include "b200_context.h"
#include <nccl.h>
#include <dlfcn.h>
#include <cstring>
#include <cstdio>
#include <new>

using namespace b200;

struct b200phys_context { Context c; };

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return -(int)e_; } while (0)

namespace {

struct NcclApi
{
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    bool ok = false;
};

NcclApi& nccl()
{
    static NcclApi api;
    if (api.lib) return api;
    api.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) api.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!api.lib) return api;
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.lib, "ncclCommDestroy");
    api.GroupStart = (decltype(api.GroupStart))dlsym(api.lib, "ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.lib, "ncclGroupEnd");
    api.Send = (decltype(api.Send))dlsym(api.lib, "ncclSend");
    api.Recv = (decltype(api.Recv))dlsym(api.lib, "ncclRecv");
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.GroupStart && api.GroupEnd && api.Send && api.Recv;
    return api;
}

} // namespace

namespace b200 {

struct DistState
{
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1, axis = 0;
    float lo = -3.0e38f, hi = 3.0e38f, ghostWidth = 1.f;
    uint32_t capB = 0, capC = 0;
    void* sendMsg[2] = { nullptr, nullptr };   // to the lower / upper neighbour
    void* recvMsg[2] = { nullptr, nullptr };
    size_t msgBytes = 0;
};

void destroyDist(Context* c)
{
    DistState* d = (DistState*)c->dist;
    if (!d) return;
    for (int k = 0; k < 2; ++k) { if (d->sendMsg[k]) cudaFree(d->sendMsg[k]); if (d->recvMsg[k]) cudaFree(d->recvMsg[k]); }
    if (d->comm && nccl().ok) nccl().CommDestroy(d->comm);
    delete d; c->dist = nullptr;
}

} // namespace b200

extern "C" {

int b200phys_dist_unique_id(uint8_t* out128)
{
    if (!out128) return B200PHYS_E_INVALID;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    if (!nccl().ok) return B200PHYS_E_NOT_READY;
    ncclUniqueId id;
    if (nccl().GetUniqueId(&id) != ncclSuccess) return B200PHYS_E_INVALID;
    memcpy(out128, &id, 128);
    return 0;
}

int b200phys_create_distributed(const b200phys_config* config, const b200phys_dist_config* dist, const uint8_t* uniqueId128, b200phys_context** out)
{
    if (!config || !dist || !uniqueId128 || !out || dist->nranks < 1 || dist->rank < 0 || dist->rank >= dist->nranks || dist->axis < 0 || dist->axis > 2) return B200PHYS_E_INVALID;
    if (!nccl().ok) return B200PHYS_E_NOT_READY;
    // room for the two ghost layers behind the owned bodies / colliders
    b200phys_config cfg = *config;
    cfg.maxBodies += 2u * dist->maxGhostBodies; cfg.maxColliders += 2u * dist->maxGhostColliders;
    int r = b200phys_create(&cfg, out);
    if (r) return r;
    Context& c = (*out)->c;
    DistState* d = new (std::nothrow) DistState();
    if (!d) { b200phys_destroy(*out); *out = nullptr; return B200PHYS_E_INVALID; }
    c.dist = d;
    d->rank = dist->rank; d->nranks = dist->nranks; d->axis = dist->axis; d->lo = dist->slabLo; d->hi = dist->slabHi; d->ghostWidth = dist->ghostWidth;
    d->capB = dist->maxGhostBodies; d->capC = dist->maxGhostColliders;
    d->msgBytes = (size_t)b200phys_halo_message_bytes(d->capB, d->capC);
    for (int k = 0; k < 2; ++k)
    {
        if (cudaMalloc(&d->sendMsg[k], d->msgBytes) != cudaSuccess || cudaMalloc(&d->recvMsg[k], d->msgBytes) != cudaSuccess) { b200phys_destroy(*out); *out = nullptr; return -(int)cudaErrorMemoryAllocation; }
        cudaMemset(d->sendMsg[k], 0, d->msgBytes); cudaMemset(d->recvMsg[k], 0, d->msgBytes);
    }
    ncclUniqueId id; memcpy(&id, uniqueId128, 128);
    if (nccl().CommInitRank(&d->comm, dist->nranks, id, dist->rank) != ncclSuccess) { b200phys_destroy(*out); *out = nullptr; return B200PHYS_E_INVALID; }
    return 0;
}

// ghost exchange with the slab neighbours, then the step. Everything is enqueued on the context's stream; no host synchronisation.
int b200phys_step_distributed(b200phys_context* ctx, float dt, uint32_t numSolverIterations, uint32_t flags)
{
    if (!ctx || !ctx->c.dist) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    DistState& d = *(DistState*)c.dist;
    int r = b200phys_halo_clear(ctx); if (r) return r;
    const bool hasLower = d.rank > 0, hasUpper = d.rank < d.nranks - 1;
    // owned bodies whose AABB reaches into a neighbour's slab (+ ghost width) -> fixed-size message per neighbour
    if (hasLower) { r = b200phys_halo_pack_message(ctx, d.axis, -3.0e38f, d.lo + d.ghostWidth, d.sendMsg[0], d.capB, d.capC); if (r) return r; }
    if (hasUpper) { r = b200phys_halo_pack_message(ctx, d.axis, d.hi - d.ghostWidth, 3.0e38f, d.sendMsg[1], d.capB, d.capC); if (r) return r; }
    if (hasLower || hasUpper)
    {
        NcclApi& n = nccl();
        if (n.GroupStart() != ncclSuccess) return B200PHYS_E_INVALID;
        if (hasLower) { n.Send(d.sendMsg[0], d.msgBytes, ncclChar, d.rank - 1, d.comm, c.stream); n.Recv(d.recvMsg[0], d.msgBytes, ncclChar, d.rank - 1, d.comm, c.stream); }
        if (hasUpper) { n.Send(d.sendMsg[1], d.msgBytes, ncclChar, d.rank + 1, d.comm, c.stream); n.Recv(d.recvMsg[1], d.msgBytes, ncclChar, d.rank + 1, d.comm, c.stream); }
        if (n.GroupEnd() != ncclSuccess) return B200PHYS_E_INVALID;
    }
    if (hasLower) { r = b200phys_halo_unpack_message(ctx, d.recvMsg[0], d.capB, d.capC); if (r) return r; }
    if (hasUpper) { r = b200phys_halo_unpack_message(ctx, d.recvMsg[1], d.capB, d.capC); if (r) return r; }
    return b200phys_step(ctx, dt, numSolverIterations, flags);
}

int b200phys_dist_ghost_counts(b200phys_context* ctx, uint32_t* outBodies, uint32_t* outColliders)
{   // ghosts received in the last exchange, summed over the neighbours (reads the message headers: synchronous)
    if (!ctx || !ctx->c.dist) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    DistState& d = *(DistState*)c.dist;
    uint32_t tb = 0, tc = 0;
    CK(cudaStreamSynchronize(c.stream));
    for (int k = 0; k < 2; ++k)
    {
        if ((k == 0 && d.rank == 0) || (k == 1 && d.rank == d.nranks - 1)) continue;
        uint32_t hdr[4] = { 0, 0, 0, 0 };
        CK(cudaMemcpy(hdr, d.recvMsg[k], 16, cudaMemcpyDeviceToHost));
        tb += hdr[0]; tc += hdr[1];
    }
    if (outBodies) *outBodies = tb; if (outColliders) *outColliders = tc;
    return 0;
}

} // extern "C"
