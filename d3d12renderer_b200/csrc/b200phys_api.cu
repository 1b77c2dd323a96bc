// b200phys_api.cu -- the C ABI of libb200phys.so (include/b200phys.h): context lifetime, scene upload, step, downloads.
// Host side only orchestrates: allocate HBM once (replaces the reference's per-step arena scratch, physics.cpp:1205-1230),
// enqueue the fixed kernel sequence of one physicsStepInternal (physics.cpp:1180-1362) on the context's stream, copy results.
// There is no CPU implementation of any stage in this library: without a CUDA device b200phys_create fails.
#include <algorithm>
#include <vector>
#include <unordered_map>
#include "b200_context.h"
#include <cstdio>
#include <cstring>
#include <cmath>
#include <new>

using namespace b200;

struct b200phys_context { Context c; };

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return -(int)e_; } while (0)

namespace {

template <typename T> cudaError_t devAlloc(T*& p, size_t n) { p = nullptr; if (n == 0) n = 1; return cudaMalloc((void**)&p, n * sizeof(T)); }

int allocateAll(Context& c)
{
    const b200phys_config& g = c.cfg;
    const size_t N1 = (size_t)g.maxBodies + 1, C = g.maxColliders ? g.maxColliders : 1, T = g.maxTransforms ? g.maxTransforms : 1;
    c.maxPairs = g.maxPairs ? g.maxPairs : 16u * (uint32_t)C;
    c.maxManifolds = g.maxManifolds ? g.maxManifolds : c.maxPairs;
    if (c.maxManifolds > c.maxPairs) c.maxManifolds = c.maxPairs;
    c.maxContacts = g.maxContacts ? g.maxContacts : 2u * c.maxManifolds + 1024u;
    c.maxJoints = g.maxJointsPerKind ? g.maxJointsPerKind : g.maxBodies;
    c.maxCells = g.maxGridCells ? g.maxGridCells : (1u << 24);
    const size_t P = c.maxPairs, M = c.maxManifolds, K = c.maxContacts, J = c.maxJoints;

    CK(devAlloc(c.bRot, N1)); CK(devAlloc(c.bPos, N1)); CK(devAlloc(c.bCog, N1)); CK(devAlloc(c.bInvI, 3 * N1));
    CK(devAlloc(c.bVel, N1)); CK(devAlloc(c.bAng, N1)); CK(devAlloc(c.bForce, N1)); CK(devAlloc(c.bTorque, N1));
    CK(devAlloc(c.gVel, 2 * N1)); CK(devAlloc(c.gPos, N1)); CK(devAlloc(c.gInvI, 3 * N1));
    CK(devAlloc(c.bodyMask, N1 + 1)); CK(devAlloc(c.bodyBase, N1 + 1)); CK(devAlloc(c.bodyBaseJoints, N1 + 1));
    CK(devAlloc(c.colorQueue, (M > J ? M : J) + 32));
    for (int k = 0; k < 6; ++k) CK(devAlloc(c.refColorsJoint[k], J + 2));
    CK(devAlloc(c.refKeys[0], M + 2)); CK(devAlloc(c.refKeys[1], M + 2)); CK(devAlloc(c.refColors, M + 2)); CK(devAlloc(c.colorKey, M + 2));
    c.scanStatusTiles = (uint32_t)((std::max<size_t>(16 * P, (size_t)c.maxCells + 2) + 2047) / 2048 + 64); CK(devAlloc(c.scanStatus, c.scanStatusTiles)); CK(devAlloc(c.scanState, 4));
    CK(cudaMemset(c.scanStatus, 0, sizeof(unsigned long long) * c.scanStatusTiles)); CK(cudaMemset(c.scanState, 0, 16));
    CK(devAlloc(c.cLocal, 4 * C)); CK(devAlloc(c.xforms, 2 * T)); CK(devAlloc(c.wCol, 4 * C)); CK(devAlloc(c.aabbMin, C)); CK(devAlloc(c.aabbMax, C)); CK(devAlloc(c.bSphere, C));
    CK(devAlloc(c.hullVerts, 1)); CK(devAlloc(c.hullOffsets, 2)); CK(devAlloc(c.hullAabb, 2));
    CK(devAlloc(c.cellOf, C)); CK(devAlloc(c.rankInCell, C)); CK(devAlloc(c.cellCount, (size_t)c.maxCells + 2)); CK(devAlloc(c.cellStart, (size_t)c.maxCells + 2));
    CK(devAlloc(c.bpRec, 2 * C)); CK(devAlloc(c.bpRows, 32 * C)); CK(devAlloc(c.bpBigChunks, (size_t)kMaxBig * ((C + 2047) / 2048 + 1))); CK(devAlloc(c.bigList, kMaxBig));
    CK(cudaMemset(c.cellCount, 0, sizeof(uint32_t) * ((size_t)c.maxCells + 2)));      // the cell scan leaves the counts cleared for the next step
    CK(devAlloc(c.pairCount, C + 2)); CK(devAlloc(c.pairStart, C + 2)); CK(devAlloc(c.pairs, P));
    {   // candidate reuse across steps (k_broad.cu): grown boxes + a candidate list half as large again as the pair list
        const char* e = getenv("B200PHYS_BP_REUSE");
        c.bpReuse = !(e && e[0] == '0') && g.reserved[2] != 1u;
        if (c.bpReuse)
        {
            c.maxCand = (uint32_t)std::min<size_t>(P + P / 2, 0xFFFFFFF0u);
            CK(devAlloc(c.bpInflMin, C)); CK(devAlloc(c.bpInflMax, C)); CK(devAlloc(c.bpCand, (size_t)c.maxCand + 2));
        }
    }
    CK(devAlloc(c.pairKey, P)); CK(devAlloc(c.pairsTyped, P)); CK(devAlloc(c.pairDst, P + 2)); CK(devAlloc(c.pairsSorted, P));
    CK(devAlloc(c.manTmp, 5 * P)); CK(devAlloc(c.manScan, P + 2)); CK(devAlloc(c.manScanOut, P + 2));
    c.maxEpaWork = (uint32_t)M; CK(devAlloc(c.epaSimplex, 9 * M)); CK(devAlloc(c.epaWork, M)); CK(devAlloc(c.epaRedo, M)); CK(devAlloc(c.epaRedo2, M));
    CK(devAlloc(c.collisions, M)); CK(devAlloc(c.manifolds, M)); CK(devAlloc(c.contacts, 2 * K)); CK(devAlloc(c.contactBodies, K));
    for (uint32_t gI = 0; gI < kNumGroups; ++gI)
    {
        const size_t units = (gI == 6) ? M : J;
        CK(devAlloc(c.unitColor[gI], units)); CK(devAlloc(c.unitOrder[gI], units)); CK(devAlloc(c.unitDep[gI], units));
    }
    CK(devAlloc(c.unitPending, (M > J ? M : J))); CK(devAlloc(c.nbrCount, (M > J ? M : J))); CK(devAlloc(c.nbrList, 16 * (M > J ? M : J))); CK(devAlloc(c.unitForbid, (M > J ? M : J)));
    CK(devAlloc(c.adjStart, N1 + 2)); CK(devAlloc(c.adjFill, N1 + 2)); CK(devAlloc(c.adjList, 2 * (M > J ? M : J))); CK(devAlloc(c.adjPos, 2 * (M > J ? M : J))); CK(devAlloc(c.unitSlots, (M > J ? M : J)));
    CK(devAlloc(c.levelStart, (size_t)kNumGroups * (kMaxLevels + 1))); CK(devAlloc(c.levelFill, (size_t)3 * (kMaxLevels + 1)));
    CK(devAlloc(c.slotOfContact, K)); CK(devAlloc(c.unitSlotStart, M + 2)); CK(devAlloc(c.unitInfo, M)); CK(devAlloc(c.rows, 8 * K)); CK(devAlloc(c.rowImp, K));
    static const size_t inBytes[6] = { sizeof(b200phys_distance_joint), sizeof(b200phys_ball_joint), sizeof(b200phys_fixed_joint),
                                       sizeof(b200phys_hinge_joint), sizeof(b200phys_cone_twist_joint), sizeof(b200phys_slider_joint) };
    for (int k = 0; k < 6; ++k)
    {
        uint8_t* p = nullptr; CK(devAlloc(p, J * inBytes[k])); c.jointIn[k] = p;
        CK(devAlloc(c.jointBodies[k], J));
        uint8_t* r = nullptr; CK(devAlloc(r, J * 512)); c.jointRows[k] = r;   // every row struct is < 512 bytes (static_assert in k_solver.cu)
    }
    CK(devAlloc(c.scanTmp, 4096)); CK(devAlloc(c.scanTmp64, 4096));
    const size_t totalWarps = (size_t)c.numSMs * 4 * 8;
    CK(devAlloc(c.sortHist, 2 * (64 * totalWarps + 64) + 64));
    c.stagingBytes = (size_t)g.maxBodies * sizeof(b200phys_body);
    if (c.stagingBytes < C * 24) c.stagingBytes = C * 24;
    if (c.stagingBytes < 1024) c.stagingBytes = 1024;
    CK(devAlloc(c.staging, c.stagingBytes));
    { float* fs = nullptr; CK(devAlloc(fs, (size_t)g.maxBodies * 6 + 8)); c.forceStaging = fs; }
    CK(devAlloc(c.counts, 1));
    CK(cudaMemset(c.counts, 0, sizeof(DeviceCounts)));
    { const uint32_t one = 1u; CK(cudaMemcpy(&c.counts->bpRebuild, &one, sizeof(one), cudaMemcpyHostToDevice)); }     // without candidate reuse the grid pipeline runs in every step
    CK(cudaHostAlloc((void**)&c.hostCounts, sizeof(DeviceCounts), cudaHostAllocDefault));
    memset(c.hostCounts, 0, sizeof(DeviceCounts));
    // the dummy body slot and all planes start zeroed
    CK(cudaMemset(c.bRot, 0, N1 * 16)); CK(cudaMemset(c.bPos, 0, N1 * 16)); CK(cudaMemset(c.bCog, 0, N1 * 16)); CK(cudaMemset(c.bInvI, 0, 3 * N1 * 16));
    CK(cudaMemset(c.bVel, 0, N1 * 16)); CK(cudaMemset(c.bAng, 0, N1 * 16)); CK(cudaMemset(c.bForce, 0, N1 * 16)); CK(cudaMemset(c.bTorque, 0, N1 * 16));
    CK(cudaMemset(c.hullOffsets, 0, 8));
    for (int i = 0; i < 8; ++i) CK(cudaEventCreate(&c.ev[i]));
    return 0;
}

void freeAll(Context& c)
{
    void* ptrs[] = { c.bRot, c.bPos, c.bCog, c.bInvI, c.bVel, c.bAng, c.bForce, c.bTorque, c.gVel, c.gPos, c.gInvI, c.bodyMask, c.bodyBase, c.refKeys[0], c.refKeys[1], c.refColors, c.colorKey, c.colorQueue, c.scanStatus, c.scanState, c.cLocal, c.xforms, c.wCol, c.aabbMin, c.aabbMax,
        c.hullVerts, c.hullOffsets, c.hullAabb, c.cellOf, c.rankInCell, c.cellCount, c.cellStart, c.bpRec, c.bpRows, c.bpBigChunks, c.bigList, c.pairCount, c.pairStart, c.pairs,
        c.pairKey, c.pairsTyped, c.pairDst, c.pairsSorted, c.manTmp, c.manScan, c.manScanOut, c.epaSimplex, c.epaWork, c.epaRedo, c.epaRedo2, c.collisions, c.manifolds, c.contacts, c.contactBodies,
        c.unitPending, c.unitForbid, c.nbrCount, c.nbrList, c.adjStart, c.adjFill, c.adjList, c.adjPos, c.unitSlots, c.levelStart, c.levelFill, c.slotOfContact, c.unitSlotStart, c.unitInfo, c.rows, c.rowImp, c.scanTmp, c.scanTmp64, c.sortHist, c.staging, c.counts, c.haloBuf, c.bpInflMin, c.bpInflMax, c.bpCand, c.bSphere };
    for (void* p : ptrs) if (p) cudaFree(p);
    for (int k = 0; k < 6; ++k) if (c.refColorsJoint[k]) cudaFree(c.refColorsJoint[k]);
    for (int k = 0; k < 2; ++k) { if (c.evKeys[k]) cudaFree(c.evKeys[k]); if (c.evPairs[k]) cudaFree(c.evPairs[k]); if (c.evFlag[k]) cudaFree(c.evFlag[k]); if (c.evSlot[k]) cudaFree(c.evSlot[k]); }
    if (c.evBegin) cudaFree(c.evBegin); if (c.evEnd) cudaFree(c.evEnd);
    if (c.hmHeights) cudaFree(c.hmHeights); if (c.hmMips) cudaFree(c.hmMips);
    if (c.bodyBaseJoints) cudaFree(c.bodyBaseJoints); delete[] c.jointPairsHost;
    if (c.interactions) cudaFree(c.interactions); if (c.forceFields) cudaFree(c.forceFields); delete c.prevTriggerOverlaps;
    for (uint32_t g = 0; g < kNumGroups; ++g) { if (c.unitColor[g]) cudaFree(c.unitColor[g]); if (c.unitOrder[g]) cudaFree(c.unitOrder[g]); if (c.unitDep[g]) cudaFree(c.unitDep[g]); }
    for (int k = 0; k < 6; ++k) { if (c.jointIn[k]) cudaFree(c.jointIn[k]); if (c.jointBodies[k]) cudaFree(c.jointBodies[k]); if (c.jointRows[k]) cudaFree(c.jointRows[k]); }
    if (c.hostCounts) cudaFreeHost(c.hostCounts);
    for (int i = 0; i < 8; ++i) if (c.ev[i]) cudaEventDestroy(c.ev[i]);
    for (int i = 0; i < Context::kMaxMarks; ++i) if (c.markEv[i]) cudaEventDestroy(c.markEv[i]);
    if (c.copyStream) cudaStreamDestroy(c.copyStream);
    if (c.evCopyGate) cudaEventDestroy(c.evCopyGate); if (c.evForcesDone) cudaEventDestroy(c.evForcesDone);
    if (c.forceStaging) cudaFree(c.forceStaging);
    if (c.ownStream) cudaStreamDestroy(c.ownStream);
}

int fetchCounts(Context& c)
{
    CK(cudaMemcpyAsync(c.hostCounts, c.counts, sizeof(DeviceCounts), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    return 0;
}

} // namespace

extern "C" {

int b200phys_version(void) { return B200PHYS_VERSION; }

const char* b200phys_error_string(int code)
{
    if (code == 0) return "ok";
    if (code < 0) return cudaGetErrorString((cudaError_t)(-code));
    switch (code)
    {
    case B200PHYS_E_INVALID: return "invalid argument";
    case B200PHYS_E_CAPACITY: return "device buffer capacity exceeded (see b200phys_stats overflow counters)";
    case B200PHYS_E_NOT_READY: return "not ready (nothing uploaded / no step yet)";
    default: return "unknown error";
    }
}

void* b200phys_host_alloc(size_t bytes) { void* p = nullptr; if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr; return p; }
void b200phys_host_free(void* p) { if (p) cudaFreeHost(p); }

int b200phys_create(const b200phys_config* config, b200phys_context** out)
{
    if (!config || !out) return B200PHYS_E_INVALID;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess) return -(int)e;
    if (count == 0) return -(int)cudaErrorNoDevice;
    if (config->device < 0 || config->device >= count) return B200PHYS_E_INVALID;
    CK(cudaSetDevice(config->device));
    b200phys_context* ctx = new (std::nothrow) b200phys_context();
    if (!ctx) return B200PHYS_E_INVALID;
    Context& c = ctx->c;
    c.cfg = *config; c.device = config->device;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, c.device));
    c.numSMs = prop.multiProcessorCount;
    if (!prop.cooperativeLaunch) { delete ctx; return -(int)cudaErrorNotSupported; }
    CK(cudaStreamCreateWithFlags(&c.ownStream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c.copyStream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c.evCopyGate, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&c.evForcesDone, cudaEventDisableTiming));
    c.stream = c.ownStream;
    int r = allocateAll(c);
    if (r != 0) { freeAll(c); delete ctx; return r; }
    *out = ctx;
    return 0;
}

static void dropGraphs(Context& c);
int b200phys_destroy(b200phys_context* ctx)
{
    if (!ctx) return B200PHYS_E_INVALID;
    cudaSetDevice(ctx->c.device);
    cudaDeviceSynchronize();
    dropGraphs(ctx->c);
    destroyEnv(&ctx->c);
    destroyCloths(&ctx->c);
    destroyDist(&ctx->c);
    freeAll(ctx->c);
    delete ctx;
    return 0;
}

int b200phys_set_stream(b200phys_context* ctx, void* cudaStream)
{
    if (!ctx) return B200PHYS_E_INVALID;
    ctx->c.stream = cudaStream ? (cudaStream_t)cudaStream : ctx->c.ownStream;
    return 0;
}

int b200phys_enable_timing(b200phys_context* ctx, int enabled) { if (!ctx) return B200PHYS_E_INVALID; ctx->c.timing = enabled != 0; return 0; }
int b200phys_enable_graphs(b200phys_context* ctx, int enabled) { if (!ctx) return B200PHYS_E_INVALID; ctx->c.graphsOff = enabled == 0; return 0; }

int b200phys_set_hulls(b200phys_context* ctx, const float* vertices, const uint32_t* vertexOffsets, uint32_t numHulls)
{
    if (!ctx || (numHulls && (!vertices || !vertexOffsets))) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    CK(cudaStreamSynchronize(c.stream));
    const uint32_t nv = numHulls ? vertexOffsets[numHulls] : 0;
    if (c.hullVerts) cudaFree(c.hullVerts); if (c.hullOffsets) cudaFree(c.hullOffsets); if (c.hullAabb) cudaFree(c.hullAabb);
    CK(devAlloc(c.hullVerts, (size_t)nv + 1)); CK(devAlloc(c.hullOffsets, (size_t)numHulls + 2)); CK(devAlloc(c.hullAabb, 2 * (size_t)numHulls + 2));
    float4* hv = new float4[nv + 1]; float4* ha = new float4[2 * numHulls + 2];
    for (uint32_t h = 0; h < numHulls; ++h)
    {   // local aabb like hullFromMesh (bounding_volumes.cpp:1385-1397): grow() over the vertices
        float mn[3] = { 3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f }, mx[3] = { -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f };
        for (uint32_t v = vertexOffsets[h]; v < vertexOffsets[h + 1]; ++v)
        {
            const float* p = vertices + 3 * (size_t)v;
            hv[v] = make_float4(p[0], p[1], p[2], 0.f);
            for (int a = 0; a < 3; ++a) { mn[a] = mn[a] < p[a] ? mn[a] : p[a]; mx[a] = mx[a] < p[a] ? p[a] : mx[a]; }
        }
        // .w of the min corner: radius of the sphere around the box centre that holds every vertex (rounded up; the narrow phase's bounding sphere)
        double r2 = 0.0;
        for (uint32_t v = vertexOffsets[h]; v < vertexOffsets[h + 1]; ++v)
        {
            const float* p = vertices + 3 * (size_t)v; double d2 = 0.0;
            for (int a = 0; a < 3; ++a) { const double d = (double)p[a] - 0.5 * ((double)mn[a] + (double)mx[a]); d2 += d * d; }
            if (d2 > r2) r2 = d2;
        }
        const float radius = vertexOffsets[h + 1] > vertexOffsets[h] ? (float)(sqrt(r2) * 1.00001 + 1e-7) : 0.f;
        ha[2 * h] = make_float4(mn[0], mn[1], mn[2], radius); ha[2 * h + 1] = make_float4(mx[0], mx[1], mx[2], 0.f);
    }
    cudaError_t e1 = cudaMemcpy(c.hullVerts, hv, sizeof(float4) * nv, cudaMemcpyHostToDevice);
    cudaError_t e2 = cudaMemcpy(c.hullOffsets, vertexOffsets, sizeof(uint32_t) * ((size_t)numHulls + (numHulls ? 1 : 0)), cudaMemcpyHostToDevice);
    cudaError_t e3 = cudaMemcpy(c.hullAabb, ha, sizeof(float4) * 2 * numHulls, cudaMemcpyHostToDevice);
    delete[] hv; delete[] ha;
    CK(e1); CK(e2); CK(e3);
    c.numHulls = numHulls;
    return 0;
}

int b200phys_upload_bodies(b200phys_context* ctx, const b200phys_body* bodies, uint32_t count)
{
    if (!ctx || (count && !bodies) || count > ctx->c.cfg.maxBodies) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (count)
    {
        waitPendingForceUpload(&c);                     // a full upload overwrites the accumulators too
        CK(cudaMemcpyAsync(c.staging, bodies, sizeof(b200phys_body) * (size_t)count, cudaMemcpyHostToDevice, c.stream));
        launchUnpackBodies(&c, c.staging, count);
    }
    c.numBodies = c.numOwnedBodies = count; ++c.bodiesEpoch;
    CK(cudaGetLastError());
    return 0;
}

int b200phys_upload_body_states(b200phys_context* ctx, const b200phys_body_state* states, uint32_t count)
{
    if (!ctx || (count && !states) || count != ctx->c.numOwnedBodies) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (count)
    {
        CK(cudaMemcpyAsync(c.staging, states, sizeof(b200phys_body_state) * (size_t)count, cudaMemcpyHostToDevice, c.stream));
        launchUnpackBodyStates(&c, c.staging, count);
    }
    CK(cudaGetLastError());
    return 0;
}

int b200phys_upload_forces(b200phys_context* ctx, const float* forceTorque6, uint32_t count)
{
    if (!ctx || (count && !forceTorque6) || count != ctx->c.numOwnedBodies) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (count)
    {
        // On the copy stream, gated behind everything already queued on the context's stream (the previous step, a state
        // upload): the transfer and the unpack then overlap the next step's collision detection, which does not read the
        // accumulators; the step waits for them right before the force integration (waitPendingForceUpload).
        CK(cudaEventRecord(c.evCopyGate, c.stream));
        CK(cudaStreamWaitEvent(c.copyStream, c.evCopyGate, 0));
        CK(cudaMemcpyAsync(c.forceStaging, forceTorque6, 24 * (size_t)count, cudaMemcpyHostToDevice, c.copyStream));
        launchUnpackForces(&c, c.forceStaging, count, c.copyStream);
        CK(cudaEventRecord(c.evForcesDone, c.copyStream));
        c.forcesPending = true;
    }
    CK(cudaGetLastError());
    return 0;
}

int b200phys_upload_transforms(b200phys_context* ctx, const b200phys_transform* transforms, uint32_t count)
{
    if (!ctx || (count && !transforms) || count > (ctx->c.cfg.maxTransforms ? ctx->c.cfg.maxTransforms : 1)) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (count) CK(cudaMemcpyAsync(c.xforms, transforms, sizeof(b200phys_transform) * (size_t)count, cudaMemcpyHostToDevice, c.stream));
    c.numTransforms = count;
    return 0;
}

int b200phys_upload_colliders(b200phys_context* ctx, const b200phys_collider* colliders, uint32_t count)
{
    if (!ctx || (count && !colliders) || count > ctx->c.cfg.maxColliders) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (count) CK(cudaMemcpyAsync(c.cLocal, colliders, sizeof(b200phys_collider) * (size_t)count, cudaMemcpyHostToDevice, c.stream));
    c.numColliders = c.numOwnedColliders = count;
    c.bpPrimed = false;
    c.numNonCollisionColliders = 0;
    for (uint32_t i = 0; i < count; ++i) if (colliders[i].objectType == B200PHYS_OBJ_FORCE_FIELD || colliders[i].objectType == B200PHYS_OBJ_TRIGGER) ++c.numNonCollisionColliders;
    if (c.numNonCollisionColliders && !c.interactions) { c.maxInteractions = 4u * c.cfg.maxColliders + 1024u; CK(devAlloc(c.interactions, c.maxInteractions)); }
    if (c.numNonCollisionColliders && !c.forceFields)
    {   // no force uploaded yet: fields push with zero force
        c.maxForceFields = c.cfg.maxColliders; CK(devAlloc(c.forceFields, c.maxForceFields + 1)); CK(cudaMemset(c.forceFields, 0, sizeof(float4) * (c.maxForceFields + 1)));
    }
    return 0;
}

int b200phys_upload_joints(b200phys_context* ctx, int kind, const void* joints, const b200phys_body_pair* pairs, uint32_t count)
{
    if (!ctx || kind < 0 || kind >= 6 || (count && (!joints || !pairs)) || count > ctx->c.maxJoints) return B200PHYS_E_INVALID;
    static const size_t inBytes[6] = { sizeof(b200phys_distance_joint), sizeof(b200phys_ball_joint), sizeof(b200phys_fixed_joint),
                                       sizeof(b200phys_hinge_joint), sizeof(b200phys_cone_twist_joint), sizeof(b200phys_slider_joint) };
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (count)
    {
        CK(cudaMemcpyAsync(c.jointIn[kind], joints, inBytes[kind] * (size_t)count, cudaMemcpyHostToDevice, c.stream));
        CK(cudaMemcpyAsync(c.jointBodies[kind], pairs, sizeof(b200phys_body_pair) * (size_t)count, cudaMemcpyHostToDevice, c.stream));
    }
    // topology change? (the host mirror re-uploads the structs every step for the motor targets; the pairs rarely change)
    if (!c.jointPairsHost) c.jointPairsHost = new std::vector<b200phys_body_pair>[6];
    std::vector<b200phys_body_pair>& prev = c.jointPairsHost[kind];
    if (prev.size() != count || (count && memcmp(prev.data(), pairs, sizeof(b200phys_body_pair) * (size_t)count) != 0)) { prev.assign(pairs, pairs + count); ++c.jointsEpoch; }
    c.numJoints[kind] = count;
    return 0;
}

// the stream operations of one step; every launch goes through launchK / memsetK (b200_context.h), so this body can be executed for
// real, captured into a CUDA graph, or as a dry run that only hashes what it would enqueue
static void enqueueStep(Context& c, float dt, uint32_t numSolverIterations, uint32_t flags)
{
    c.launches = 0; c.numMarks = 0;
    if (c.timing) cudaEventRecord(c.ev[0], c.stream);
    timelineMark(&c, "begin");
    if (c.numNonCollisionColliders) waitPendingForceUpload(&c);   // force fields add to the accumulators during the narrow phase
    launchWorldColliders(&c);                               // physics.cpp:1217
    timelineMark(&c, "world_colliders");
    if (c.timing) cudaEventRecord(c.ev[1], c.stream);
    launchBroadphase(&c);                                   // physics.cpp:1222
    if (c.timing) cudaEventRecord(c.ev[2], c.stream);
    launchNarrowphase(&c);                                  // physics.cpp:1233
    if (c.timing) cudaEventRecord(c.ev[3], c.stream);
    waitPendingForceUpload(&c);
    launchIntegrateForces(&c, dt);                          // physics.cpp:1266-1279
    launchCollisionEvents(&c);                              // physics.cpp:1284 (heightmap collisions carry no second collider and take no part, :1049)
    launchHeightmap(&c);                                    // physics.cpp:1238-1248: contacts behind the narrow phase's
    timelineMark(&c, "integrate_forces");
    launchSolver(&c, dt, numSolverIterations, flags);       // physics.cpp:1316-1334 (records ev[4] before the iterations)
    if (c.timing) cudaEventRecord(c.ev[5], c.stream);
    launchIntegrateVelocities(&c, dt);                      // physics.cpp:1338-1348
    timelineMark(&c, "integrate_velocities");
    if (c.timing) cudaEventRecord(c.ev[6], c.stream);
    c.stepped = true;
}

// CUDA graph replay. A step is ~40-50 small launches; at 1 024 bodies the host cannot issue them as fast as the GPU retires them.
// The launch sequence depends on host state with a period of a few steps (colour inheritance age, ping-pong buffers, event parity),
// so a step is first run DRY: the same code advances the host state and hashes every kernel / geometry / argument value it would
// enqueue. A hash seen for the second time is captured (the host state is rewound for that), later occurrences are replayed with one
// cudaGraphLaunch. Nothing is assumed about what the step depends on: whatever changes an argument changes the hash.
struct GraphEntry { cudaGraphExec_t exec = nullptr; uint32_t seen = 0; uint32_t launches = 0; };
typedef std::unordered_map<unsigned long long, GraphEntry> GraphCache;
static const size_t kMaxGraphs = 256;

static void dropGraphs(Context& c)
{
    GraphCache* cache = (GraphCache*)c.graphCache;
    if (!cache) return;
    for (auto& kv : *cache) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    delete cache; c.graphCache = nullptr;
}

int b200phys_step(b200phys_context* ctx, float dt, uint32_t numSolverIterations, uint32_t flags)
{
    if (!ctx || !(dt > 0.f)) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    static int graphsWanted = -1;
    if (graphsWanted < 0) { const char* e = getenv("B200PHYS_GRAPH"); graphsWanted = (e && e[0] == '0') ? 0 : 1; }
    if (!graphsWanted || c.graphsOff || c.graphFailed || c.timing)
    {
        enqueueStep(c, dt, numSolverIterations, flags);
        CK(cudaGetLastError());
        return 0;
    }
    // cross-stream dependency of the accumulator upload: taken up front (a graph is captured from one stream)
    if (c.forcesPending) waitPendingForceUpload(&c);
    const Context saved = c;
    c.graphMode = 1; c.graphHash = 1469598103934665603ull;
    graphHashValue(&c, dt); graphHashValue(&c, numSolverIterations); graphHashValue(&c, flags); graphHashValue(&c, (const void*)c.stream);
    enqueueStep(c, dt, numSolverIterations, flags);
    const unsigned long long key = c.graphHash;
    c.graphMode = 0;
    if (!c.graphCache) c.graphCache = new GraphCache();
    if (((GraphCache*)c.graphCache)->size() > kMaxGraphs) { dropGraphs(c); c.graphCache = new GraphCache(); }      // scenes that keep changing: start over
    GraphEntry& entry = (*(GraphCache*)c.graphCache)[key];
    if (entry.exec)
    {   // the host state has already advanced past this step
        CK(cudaGraphLaunch(entry.exec, c.stream));
        ++c.graphReplays;
        return 0;
    }
    // not captured yet: rewind the host state and enqueue for real
    const uint32_t replays = c.graphReplays, captures = c.graphCaptures; void* const cachePtr = c.graphCache;
    c = saved; c.graphReplays = replays; c.graphCaptures = captures; c.graphCache = cachePtr;
    if (++entry.seen < 2)
    {
        enqueueStep(c, dt, numSolverIterations, flags);
        CK(cudaGetLastError());
        return 0;
    }
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamBeginCapture(c.stream, cudaStreamCaptureModeRelaxed);
    if (e == cudaSuccess)
    {
        enqueueStep(c, dt, numSolverIterations, flags);
        e = cudaStreamEndCapture(c.stream, &graph);
    }
    if (e == cudaSuccess && graph) e = cudaGraphInstantiate(&entry.exec, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess || !entry.exec)
    {   // capture is not possible here (e.g. a cooperative launch the driver refuses to capture): plain launches from now on
        cudaGetLastError();
        entry.exec = nullptr; c.graphFailed = true;
        const uint32_t r2 = c.graphReplays, c2 = c.graphCaptures;
        c = saved; c.graphReplays = r2; c.graphCaptures = c2; c.graphCache = cachePtr; c.graphFailed = true;
        enqueueStep(c, dt, numSolverIterations, flags);
        CK(cudaGetLastError());
        return 0;
    }
    ++c.graphCaptures;
    CK(cudaGraphLaunch(entry.exec, c.stream));
    return 0;
}

// heightmap_collider_component (terrain/heightmap_collider.h:132-150, heightmap_collider.cpp:5-19) + heightmap_collider_chunk::setHeights
// (:45-123): the min/max pyramid of every chunk is built here on the host (integer work, once per upload)
int b200phys_upload_heightmap(b200phys_context* ctx, const b200phys_heightmap* desc, const uint16_t* heights)
{
    if (!ctx) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    CK(cudaStreamSynchronize(c.stream));
    if (c.hmHeights) { cudaFree(c.hmHeights); c.hmHeights = nullptr; }
    if (c.hmMips) { cudaFree(c.hmMips); c.hmMips = nullptr; }
    if (!desc || !heights) return 0;                                   // no heightmap any more
    if (!desc->chunksPerDim || desc->chunksPerDim > 256 || !(desc->chunkSize > 0.f) || !(desc->amplitudeScale > 0.f)) return B200PHYS_E_INVALID;
    const uint32_t V = 129, nChunks = desc->chunksPerDim * desc->chunksPerDim, nodesPerChunk = 21845;
    std::vector<ushort2> mips((size_t)nChunks * nodesPerChunk);
    for (uint32_t ch = 0; ch < nChunks; ++ch)
    {
        const uint16_t* hh = heights + (size_t)ch * V * V;
        ushort2* m = mips.data() + (size_t)ch * nodesPerChunk;
        uint32_t seg = 128;
        for (uint32_t z = 0; z < seg; ++z)
            for (uint32_t x = 0; x < seg; ++x)
            {
                const uint16_t a = hh[V * z + x], b = hh[V * (z + 1) + x], cc = hh[V * z + x + 1], d = hh[V * (z + 1) + x + 1];
                m[seg * z + x] = make_ushort2(std::min(a, std::min(b, std::min(cc, d))), std::max(a, std::max(b, std::max(cc, d))));
            }
        ushort2* rd = m;
        for (uint32_t level = 1; level < 8; ++level)
        {
            const uint32_t stride = seg; seg >>= 1;
            ushort2* wr = rd + stride * stride;
            for (uint32_t z = 0; z < seg; ++z)
                for (uint32_t x = 0; x < seg; ++x)
                {
                    const ushort2 a = rd[stride * (2 * z) + 2 * x], b = rd[stride * (2 * z + 1) + 2 * x], cc = rd[stride * (2 * z) + 2 * x + 1], d = rd[stride * (2 * z + 1) + 2 * x + 1];
                    wr[seg * z + x] = make_ushort2(std::min(a.x, std::min(b.x, std::min(cc.x, d.x))), std::max(a.y, std::max(b.y, std::max(cc.y, d.y))));
                }
            rd = wr;
        }
    }
    CK(devAlloc(c.hmHeights, (size_t)nChunks * V * V));
    ushort2* dm = nullptr; CK(devAlloc(dm, mips.size())); c.hmMips = dm;
    CK(cudaMemcpy(c.hmHeights, heights, sizeof(uint16_t) * (size_t)nChunks * V * V, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(c.hmMips, mips.data(), sizeof(ushort2) * mips.size(), cudaMemcpyHostToDevice));
    c.hm.chunksPerDim = desc->chunksPerDim; c.hm.chunkSize = desc->chunkSize; c.hm.invChunkSize = 1.f / desc->chunkSize;
    c.hm.chunkScale = desc->chunkSize / (float)(V - 1);
    c.hm.invAmplitudeScale = 1.f / desc->amplitudeScale; c.hm.heightScale = desc->amplitudeScale / (float)65535;
    for (int k = 0; k < 3; ++k) c.hm.minCorner[k] = desc->minCorner[k];
    c.hm.restitution = desc->restitution; c.hm.friction = desc->friction;
    return 0;
}

int b200phys_upload_force_fields(b200phys_context* ctx, const float* forces3, uint32_t count)
{
    if (!ctx || (count && !forces3)) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (!c.forceFields) { c.maxForceFields = c.cfg.maxColliders; CK(devAlloc(c.forceFields, c.maxForceFields + 1)); CK(cudaMemset(c.forceFields, 0, sizeof(float4) * (c.maxForceFields + 1))); }
    if (count > c.maxForceFields) return B200PHYS_E_CAPACITY;
    std::vector<float4> tmp(count);
    for (uint32_t i = 0; i < count; ++i) tmp[i] = make_float4(forces3[3 * i], forces3[3 * i + 1], forces3[3 * i + 2], 0.f);
    if (count) { CK(cudaMemcpyAsync(c.forceFields, tmp.data(), sizeof(float4) * count, cudaMemcpyHostToDevice, c.stream)); CK(cudaStreamSynchronize(c.stream)); }
    c.numForceFields = count;
    return 0;
}

int b200phys_set_global_force_field(b200phys_context* ctx, const float force3[3])
{
    if (!ctx || !force3) return B200PHYS_E_INVALID;
    for (int k = 0; k < 3; ++k) ctx->c.globalForce[k] = force3[k];
    return 0;
}

// trigger overlaps of the last step -> sorted, unique (trigger, rigid body) keys -> merge against the previous call's list
// (handleNonCollisionInteractions, physics.cpp:970-1034). The overlap tests ran on the device; this part is a few entries.
int b200phys_download_trigger_events(b200phys_context* ctx, b200phys_trigger_event* enterOut, uint32_t enterCapacity, uint32_t* numEnter,
                                     b200phys_trigger_event* leaveOut, uint32_t leaveCapacity, uint32_t* numLeave)
{
    if (!ctx || !numEnter || !numLeave) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    *numEnter = *numLeave = 0;
    if (!c.stepped) return 0;
    if (!c.prevTriggerOverlaps) c.prevTriggerOverlaps = new std::vector<unsigned long long>();
    std::vector<unsigned long long> cur;
    if (c.numNonCollisionColliders && c.interactions)
    {
        uint32_t n = 0;
        CK(cudaMemcpyAsync(&n, &c.counts->numInteractions, 4, cudaMemcpyDeviceToHost, c.stream));
        CK(cudaStreamSynchronize(c.stream));
        std::vector<uint4> it(n);
        if (n) { CK(cudaMemcpyAsync(it.data(), c.interactions, sizeof(uint4) * (size_t)n, cudaMemcpyDeviceToHost, c.stream)); CK(cudaStreamSynchronize(c.stream)); }
        for (const uint4& x : it) if (x.z == B200PHYS_OBJ_TRIGGER) cur.push_back(((unsigned long long)x.y << 32) | x.x);
        std::sort(cur.begin(), cur.end());
        cur.erase(std::unique(cur.begin(), cur.end()), cur.end());
    }
    const std::vector<unsigned long long>& prev = *c.prevTriggerOverlaps;
    uint32_t ne = 0, nl = 0; size_t p = 0, t = 0;
    auto emit = [](b200phys_trigger_event* out, uint32_t cap, uint32_t& n, unsigned long long key) { if (out && n < cap) { out[n].triggerIndex = (uint32_t)(key >> 32); out[n].rigidBodyIndex = (uint32_t)key; } ++n; };
    while (p < prev.size() && t < cur.size())
    {
        if (prev[p] == cur[t]) { ++p; ++t; continue; }
        if (prev[p] < cur[t]) emit(leaveOut, leaveCapacity, nl, prev[p++]); else emit(enterOut, enterCapacity, ne, cur[t++]);
    }
    while (p < prev.size()) emit(leaveOut, leaveCapacity, nl, prev[p++]);
    while (t < cur.size()) emit(enterOut, enterCapacity, ne, cur[t++]);
    *numEnter = ne; *numLeave = nl;
    if (ne > enterCapacity || nl > leaveCapacity) return B200PHYS_E_CAPACITY;     // nothing consumed: call again with enough room
    *c.prevTriggerOverlaps = std::move(cur);
    return 0;
}

int b200phys_enable_events(b200phys_context* ctx, int enable)
{
    if (!ctx) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (enable && !c.evBegin)
    {
        if (c.cfg.maxColliders >= (1u << 28)) return B200PHYS_E_INVALID;      // 28 bits per collider in the event key
        const size_t M = c.maxManifolds;
        for (int k = 0; k < 2; ++k) { CK(devAlloc(c.evKeys[k], M + 2)); CK(devAlloc(c.evPairs[k], M + 2)); CK(devAlloc(c.evFlag[k], M + 2)); CK(devAlloc(c.evSlot[k], M + 2)); }
        CK(devAlloc(c.evBegin, M + 1)); CK(devAlloc(c.evEnd, M + 1));
    }
    if (enable && !c.eventsEnabled)
    {   // start from an empty "previous frame" (event_context is created empty, physics.cpp:946-950)
        CK(cudaStreamSynchronize(c.stream));
        const uint32_t zero[3] = { 0, 0, 0 };
        CK(cudaMemcpy(&c.counts->numPrevCollisions, zero, sizeof(zero), cudaMemcpyHostToDevice));
    }
    c.eventsEnabled = enable != 0;
    return 0;
}

int b200phys_download_events(b200phys_context* ctx, b200phys_collision_begin_event* beginOut, uint32_t beginCapacity, uint32_t* numBegin,
                             b200phys_collider_pair* endOut, uint32_t endCapacity, uint32_t* numEnd)
{
    if (!ctx || !numBegin || !numEnd) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    *numBegin = *numEnd = 0;
    if (!c.eventsEnabled || !c.stepped) return 0;
    uint32_t n[2] = { 0, 0 };
    CK(cudaMemcpyAsync(n, &c.counts->numBeginEvents, sizeof(n), cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    *numBegin = n[0]; *numEnd = n[1];
    const uint32_t nb = n[0] < beginCapacity ? n[0] : beginCapacity, ne = n[1] < endCapacity ? n[1] : endCapacity;
    if (beginOut && nb) CK(cudaMemcpyAsync(beginOut, c.evBegin, sizeof(b200phys_collision_begin_event) * (size_t)nb, cudaMemcpyDeviceToHost, c.stream));
    if (endOut && ne) CK(cudaMemcpyAsync(endOut, c.evEnd, sizeof(b200phys_collider_pair) * (size_t)ne, cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    return (n[0] > beginCapacity || n[1] > endCapacity) ? B200PHYS_E_CAPACITY : 0;
}

int b200phys_synchronize(b200phys_context* ctx)
{
    if (!ctx) return B200PHYS_E_INVALID;
    CK(cudaSetDevice(ctx->c.device));
    if (ctx->c.forcesPending) CK(cudaStreamSynchronize(ctx->c.copyStream));     // an accumulator upload no step has consumed yet
    CK(cudaStreamSynchronize(ctx->c.stream));
    return 0;
}

int b200phys_download_bodies(b200phys_context* ctx, b200phys_body_state* out, uint32_t count)
{
    if (!ctx || !out || count > ctx->c.numBodies) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (!count) return 0;
    launchPackBodyStates(&c, c.staging, count);
    CK(cudaMemcpyAsync(out, c.staging, sizeof(b200phys_body_state) * (size_t)count, cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    return 0;
}

int b200phys_get_broadphase_reuse(b200phys_context* ctx, uint32_t out[8])
{
    if (!ctx || !out) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    int r = fetchCounts(c); if (r) return r;
    const DeviceCounts& d = *c.hostCounts;
    out[0] = d.bpStepCount; out[1] = d.bpRebuildCount; out[2] = d.bpEmitToCand ? d.numCand : 0u; out[3] = d.numPairs; out[4] = c.bpReuse ? d.bpMarginBits : 0u; out[5] = d.bpAge;
    memcpy(&out[6], &d.bpMovedPerStep, 4); memcpy(&out[7], &d.meanSmallExtent, 4);
    return 0;
}

int b200phys_get_stats(b200phys_context* ctx, b200phys_stats* out)
{
    if (!ctx || !out) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    int r = fetchCounts(c); if (r) return r;
    const DeviceCounts& d = *c.hostCounts;
    memset(out, 0, sizeof(*out));
    out->numBodies = c.numBodies; out->numColliders = c.numColliders; out->reserved[1] = c.numOwnedBodies; out->reserved[2] = c.numOwnedColliders; out->reserved[3] = c.numNonCollisionColliders ? d.numInteractions : 0u;
    out->reserved[4] = d.numEpaWork; out->reserved[5] = d.numEpaRedo | (d.numEpaRedo2 << 20); out->reserved[6] = c.graphReplays; out->reserved[0] = c.graphCaptures;
    out->numBroadphaseOverlaps = d.numPairs; out->numCollisions = d.numManifolds; out->numContacts = d.numContacts;
    out->numBigColliders = d.numBigNow; out->numGridCells = d.grid.numCells;
    for (int g = 0; g < 7; ++g) out->numColors[g] = d.numColors[g];
    out->overflowPairs = d.overflowPairs; out->overflowManifolds = d.overflowManifolds; out->overflowGrid = d.overflowGrid | d.overflowBig | (d.overflowColors << 8) | (d.overflowHalo << 16);
    out->epaOverflows = d.epaOverflows; out->gjkIterationCaps = d.gjkIterationCaps; out->assertCount = d.assertCount;
    out->cellSize = d.grid.cellSize;
    out->kernelLaunches = c.launches;
    if (c.timing && c.stepped)
    {
        float ms[6] = {};
        for (int i = 0; i < 6; ++i) cudaEventElapsedTime(&ms[i], c.ev[i], c.ev[i + 1]);
        out->msWorld = ms[0]; out->msBroad = ms[1]; out->msNarrow = ms[2]; out->msInit = ms[3]; out->msSolve = ms[4]; out->msIntegrate = ms[5];
        cudaEventElapsedTime(&out->msTotal, c.ev[0], c.ev[6]);
    }
    if (d.overflowPairs || d.overflowManifolds || d.overflowGrid || d.overflowBig || d.overflowColors || d.overflowHalo) return B200PHYS_E_CAPACITY;
    return 0;
}

int b200phys_get_timeline(b200phys_context* ctx, char* namesOut, uint32_t namesCapacity, float* msOut, uint32_t msCapacity, uint32_t* count)
{
    if (!ctx || !count) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    *count = 0;
    if (!c.timing || !c.stepped || c.numMarks < 2) return 0;
    CK(cudaStreamSynchronize(c.stream));
    size_t used = 0;
    if (namesOut && namesCapacity) namesOut[0] = 0;
    for (int i = 1; i < c.numMarks; ++i)
    {
        float ms = 0.f; cudaEventElapsedTime(&ms, c.markEv[i - 1], c.markEv[i]);
        if (msOut && (uint32_t)(i - 1) < msCapacity) msOut[i - 1] = ms;
        if (namesOut)
        {
            const size_t len = strlen(c.markName[i]);
            if (used + len + 2 <= namesCapacity) { memcpy(namesOut + used, c.markName[i], len); used += len; namesOut[used++] = '\n'; namesOut[used] = 0; }
        }
    }
    *count = (uint32_t)(c.numMarks - 1);
    return 0;
}

int b200phys_download_aabbs(b200phys_context* ctx, float* outMinMax6, uint32_t* count)
{
    if (!ctx || !count) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (!c.stepped) return B200PHYS_E_NOT_READY;
    *count = c.numColliders;
    if (!outMinMax6 || !c.numColliders) return 0;
    launchPackAabbs(&c, (float*)c.staging);
    CK(cudaMemcpyAsync(outMinMax6, c.staging, 24 * (size_t)c.numColliders, cudaMemcpyDeviceToHost, c.stream));
    CK(cudaStreamSynchronize(c.stream));
    return 0;
}

int b200phys_download_pairs(b200phys_context* ctx, b200phys_collider_pair* out, uint32_t* count)
{
    if (!ctx || !count) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (!c.stepped) return B200PHYS_E_NOT_READY;
    int r = fetchCounts(c); if (r) return r;
    *count = c.hostCounts->numPairs;
    if (out && *count) { CK(cudaMemcpyAsync(out, c.pairs, sizeof(uint2) * (size_t)*count, cudaMemcpyDeviceToHost, c.stream)); CK(cudaStreamSynchronize(c.stream)); }
    return 0;
}

int b200phys_download_contacts(b200phys_context* ctx, b200phys_contact* outContacts, b200phys_body_pair* outBodyPairs, uint32_t* count)
{
    if (!ctx || !count) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (!c.stepped) return B200PHYS_E_NOT_READY;
    int r = fetchCounts(c); if (r) return r;
    *count = c.hostCounts->numContacts;
    if (*count)
    {
        if (outContacts) CK(cudaMemcpyAsync(outContacts, c.contacts, sizeof(b200phys_contact) * (size_t)*count, cudaMemcpyDeviceToHost, c.stream));
        if (outBodyPairs) CK(cudaMemcpyAsync(outBodyPairs, c.contactBodies, sizeof(uint2) * (size_t)*count, cudaMemcpyDeviceToHost, c.stream));
        CK(cudaStreamSynchronize(c.stream));
    }
    return 0;
}

int b200phys_download_collisions(b200phys_context* ctx, b200phys_collider_pair* outPairs, uint8_t* outContactCounts, uint32_t* count)
{
    if (!ctx || !count) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (!c.stepped) return B200PHYS_E_NOT_READY;
    int r = fetchCounts(c); if (r) return r;
    const uint32_t n = c.hostCounts->numManifolds;
    *count = n;
    if (n && outPairs) { CK(cudaMemcpyAsync(outPairs, c.collisions, sizeof(uint2) * (size_t)n, cudaMemcpyDeviceToHost, c.stream)); }
    if (n && outContactCounts)
    {
        uint4* tmp = new uint4[n];
        cudaError_t e = cudaMemcpyAsync(tmp, c.manifolds, sizeof(uint4) * (size_t)n, cudaMemcpyDeviceToHost, c.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c.stream);
        if (e == cudaSuccess) for (uint32_t i = 0; i < n; ++i) outContactCounts[i] = (uint8_t)tmp[i].w;
        delete[] tmp;
        CK(e);
    }
    CK(cudaStreamSynchronize(c.stream));
    return 0;
}

int b200phys_download_solve_order(b200phys_context* ctx, int kind, uint32_t* outOrder, uint32_t* count)
{
    if (!ctx || !count || kind < 0 || kind > 6) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CK(cudaSetDevice(c.device));
    if (!c.stepped) return B200PHYS_E_NOT_READY;
    int r = fetchCounts(c); if (r) return r;
    if (kind == 6)
    {
        *count = c.hostCounts->numContacts;
        if (outOrder && *count) { CK(cudaMemcpyAsync(outOrder, c.slotOfContact, 4 * (size_t)*count, cudaMemcpyDeviceToHost, c.stream)); CK(cudaStreamSynchronize(c.stream)); }
    }
    else
    {
        *count = c.numJoints[kind];
        if (outOrder && *count) { CK(cudaMemcpyAsync(outOrder, c.unitOrder[kind], 4 * (size_t)*count, cudaMemcpyDeviceToHost, c.stream)); CK(cudaStreamSynchronize(c.stream)); }
    }
    return 0;
}

/* multi-GPU halo entry points are implemented in k_halo.cu */

} // extern "C"
