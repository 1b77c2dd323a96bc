// b200_context.h -- device-side data layout of one b200phys_context (see DESIGN.md "Data layout in HBM").
//
// Everything a step needs lives in HBM between steps; the host only uploads what gameplay code mutates and downloads
// poses/velocities. All counts that kernels consume (pairs, manifolds, contacts, colours, grid size) stay ON THE DEVICE in
// DeviceCounts: kernels are grid-stride loops bounded by those counters, so a step is a fixed launch sequence with no
// host round trip in the middle.
#pragma once
#include <cuda_runtime.h>
#include <vector>
#include <stdint.h>
#include "../../include/b200phys.h"

namespace b200 {

static const uint32_t kMaxBig = 1024;          // colliders too large for the grid (brute-forced)
static const uint32_t kNumBuckets = 21;        // upper-triangular 6x6 type pairs (collision_narrow.cpp:2411-2421)
static const uint32_t kBucketDropped = 21;     // pruned pair (no rigid body / same body)
static const uint32_t kBucketOverlap = 22;     // trigger / force-field overlap test only
static const uint32_t kNumPairKeys = 23;
static const uint32_t kMaxColors = 64;         // fast mode: 64-bit colour masks
static const uint32_t kMaxLevels = 1u << 16;   // exact mode: wavefront levels
static const uint32_t kNumGroups = 7;          // 6 joint kinds + contacts (constraints.cpp:3748-3772 order)

struct GridParams
{
    float cellSize, invCellSize;
    float origin[3];
    int32_t dims[3];
    uint32_t numCells;
    float bigThreshold;                 // extent above which a collider skips the grid; derived by one step for the next one
    uint32_t cellDiv, reach;            // cells per largest small extent; how many cells a query can reach beyond the centre cell
    uint32_t segX, numSegs, numTiles;   // tile = segX consecutive cells of one grid row
    uint32_t useTiles;                  // 1: staged tiles (dense cells), 0: direct per-collider traversal (about one collider per cell)
};

// device-resident scalars produced and consumed by kernels
struct DeviceCounts
{
    uint32_t numEpaWork;                // GJK hits waiting for EPA this step
    uint32_t epaWorkHead;               // next work-list entry to hand out (EPA kernels with dynamic distribution)
    uint32_t numEpaRedo;                // entries the small-polytope EPA kernel handed to the large-capacity one
    uint32_t hmBaseManifolds, hmBaseContacts, hmFitManifolds, hmFitContacts;   // heightmap pass: list sizes before it, first entries that did not fit
    uint32_t numEpaRedo2, epaRedo2Head;  // entries the medium polytope handed on to the warp-cooperative kernel
    uint32_t epaRedoHead;               // next redo entry to hand out
    uint32_t gjkCursor[11];             // per GJK bucket: next pair to hand out (npGjkKernel)
    uint32_t numPairs;                  // broad-phase overlaps (clamped to capacity)
    uint32_t numPairsWanted;            // what the scan produced (overflow detection)
    uint32_t keyStart[kNumPairKeys + 1];// bucket ranges of the classified pair list
    uint32_t numManifolds, numManifoldsWanted;
    uint32_t numContacts;
    uint32_t numColors[kNumGroups];
    uint32_t numJoints[6];
    uint32_t colorStart[kNumGroups][kMaxColors + 1]; // fast mode unit ranges per colour
    uint32_t numBig;
    uint32_t bigMatrixCount;            // min(numBig, 16) * C: size of the big-collider flag matrix
    uint32_t numRefManifolds, refNumColors;     // colour reuse across steps: size / colour count of the remembered list
    uint32_t colorQueueHead, colorQueueTail;   // ready queue of the colouring
    uint32_t numOverlapPairs, numInteractions;   // rigid body vs trigger / force field pairs tested / overlapping this step
    uint32_t numPrevCollisions, numBeginEvents, numEndEvents;   // collision event layer (k_events.cu)
    uint32_t solverAbort;               // set when the dataflow sweep gave up (inconsistent dependency records)
    uint32_t overflowPairs, overflowManifolds, overflowGrid, overflowColors, overflowBig, overflowHalo;
    uint32_t epaOverflows, gjkIterationCaps, assertCount;
    unsigned long long extentSumFixed;  // sum of AABB extents, 2^-16 fixed point (deterministic mean)
    uint32_t extentCount;
    unsigned long long smallSumFixed; uint32_t smallCount;      // the same over the colliders that go into the grid
    uint32_t numBigNow;                 // big colliders of THIS step (numBig accumulates for the next classify pass)
    float halfSmallExtent;              // half the largest small extent: how far a partner's centre can lie outside a query box
    uint32_t numOverflowRows;           // colliders with more partners than the traversal keeps per row
    uint32_t maxSmallExtentBits;        // float bits (non-negative floats order like uints)
    uint32_t centreMinBits[3], centreMaxBits[3]; // sortable-uint encoded floats
    GridParams grid;
    // ---- candidate reuse across steps (k_broad.cu, "fattened boxes"): the grid pipeline runs on AABBs grown by a margin and leaves a CANDIDATE
    // list; while every collider's true AABB stays inside its grown box the step's pair list is that list filtered by the exact predicate
    uint32_t bpRebuild;                 // this step runs the grid pipeline (always 1 in contexts without reuse)
    uint32_t bpEmitToCand;              // the pipeline of this step / the last build wrote the candidate list (margin > 0) rather than the pair list itself
    uint32_t bpViolations;              // colliders found outside their grown box by this step's check
    uint32_t bpExcessBits;              // sortable-uint float: largest distance by which a true AABB face lies outside (+) / inside (-) its grown box
    uint32_t bpMarginBits;              // float bits: margin of the current list (0: the list is the exact pair list of its step, not reusable)
    uint32_t bpAge;                     // steps since the list was built
    uint32_t numCand, numCandWanted;    // candidate list size (clamped / as scanned)
    uint32_t bpScanCells, bpScanColliders;   // sizes of this step's two scans (0 when the list is reused)
    uint32_t bpRebuildCount, bpStepCount;    // statistics
    uint32_t bpMarginOff;               // a candidate list outgrew its buffer: no margins any more until the colliders are uploaded again
    float meanSmallExtent;              // mean AABB extent of the colliders that went into the grid at the last build
    float bpMovedPerStep;               // how far a box face moved per step since the last build (largest over all colliders), as of the last check
};

template <typename T> struct DevBuf { T* p = nullptr; size_t n = 0; };

struct Context
{
    b200phys_config cfg;
    int device = 0;
    int numSMs = 148;
    cudaStream_t stream = nullptr;      // the stream every kernel of this context is launched on
    cudaStream_t ownStream = nullptr;
    bool timing = false;
    cudaEvent_t ev[8] = {};
    // per-kernel-group timeline (b200phys_get_timeline): mark i closes the interval that started at mark i - 1
    static const int kMaxMarks = 32;
    cudaEvent_t markEv[kMaxMarks] = {}; const char* markName[kMaxMarks] = {}; int numMarks = 0;

    uint32_t numBodies = 0, numColliders = 0, numTransforms = 0, numHulls = 0;   // incl. ghosts
    uint32_t numOwnedBodies = 0, numOwnedColliders = 0;                            // what the host uploaded
    uint32_t numJoints[6] = {};
    bool stepped = false;
    uint32_t launches = 0;

    // ---- bodies (persistent, SoA float4 planes; N+1 entries where the last is the static dummy, physics.cpp:1279)
    float4* bRot = nullptr;     // rotation
    float4* bPos = nullptr;     // position.xyz, invMass
    float4* bCog = nullptr;     // localCOG.xyz, gravityFactor
    float4* bInvI = nullptr;    // 3 planes: local invInertia columns + (linearDamping, angularDamping, 0)
    float4* bVel = nullptr;     // linearVelocity.xyz
    float4* bAng = nullptr;     // angularVelocity.xyz
    float4* bForce = nullptr;   // forceAccumulator.xyz
    float4* bTorque = nullptr;  // torqueAccumulator.xyz
    // ---- per-step body global state (rigid_body_global_state, rigid_body.h:6-16)
    float4* gVel = nullptr;     // AoS 32 B: (v.xyz, invMass), (w.xyz, 0)  -- the only record the solver iterations touch
    float4* gPos = nullptr;     // world COG position
    float4* gInvI = nullptr;    // AoS 48 B: world inverse inertia rows

    // ---- colliders
    float4* cLocal = nullptr;   // uploaded b200phys_collider records (4 x float4)
    float4* xforms = nullptr;   // uploaded b200phys_transform records (2 x float4)
    float4* wCol = nullptr;     // world colliders, AoS 64 B (gather target of the narrow phase)
    float4* aabbMin = nullptr;  // world AABB min.xyz
    float4* aabbMax = nullptr;  // world AABB max.xyz
    float4* bSphere = nullptr;  // world bounding sphere of the shape (centre.xyz, radius; radius < 0: padding slot): the narrow phase's cheap separation proof
    float4* hullVerts = nullptr;   // all hull vertices (xyz, 0)
    uint32_t* hullOffsets = nullptr;
    float4* hullAabb = nullptr;    // 2 per hull

    // ---- broad phase (uniform grid)
    uint32_t* cellOf = nullptr;       // per collider: cell index, 0xFFFFFFFF = big
    uint32_t* rankInCell = nullptr;
    uint32_t* cellCount = nullptr;    // maxGridCells + 1
    uint32_t* cellStart = nullptr;    // maxGridCells + 1
    float4* bpRec = nullptr;          // AABBs in cell order, 32-byte records: (min.xyz, collider id bits), (max.xyz, 0)
    uint32_t* bpRows = nullptr;       // per cell-ordered collider: up to 32 partners with a larger index, ascending
    uint32_t* bpBigChunks = nullptr;  // pairs of big collider k inside collider chunk c: [k * numChunks + c]
    bool bpPrimed = false;            // the big / small threshold of the previous step is valid
    // candidate reuse: grown boxes of the last build, the candidate list, and what the host knows about it
    float4* bpInflMin = nullptr; float4* bpInflMax = nullptr; uint2* bpCand = nullptr; uint32_t maxCand = 0;
    bool bpReuse = false;             // this context keeps a candidate list (B200PHYS_BP_REUSE=0 turns it off)
    uint32_t bpListColliders = 0;     // collider count the list was built for (0: none)
    uint32_t* bigList = nullptr;
    uint32_t* pairCount = nullptr;    // per collider
    uint32_t* pairStart = nullptr;    // C + 1
    uint2* pairs = nullptr;           // canonical: a < b, sorted by (a, b)

    // ---- narrow phase
    uint8_t* pairKey = nullptr;       // bucket per pair
    uint2* pairsTyped = nullptr;      // type-ordered (collision_narrow.cpp:2374), not yet bucketed
    uint32_t* pairDst = nullptr;      // destination slot of the stable bucket sort
    uint2* pairsSorted = nullptr;     // bucketed, stable
    float4* epaSimplex = nullptr;     // 9 planes x maxEpaWork: GJK terminal simplices parked for the EPA kernel
    uint2* epaWork = nullptr;         // (sorted pair slot, bucket)
    uint32_t* epaRedo = nullptr;      // work-list entries that outgrew the small polytope
    uint32_t* epaRedo2 = nullptr;     // ... and the medium one
    uint32_t maxEpaWork = 0;
    float4* manTmp = nullptr;         // 5 planes x maxPairs: (normal, count), 4 x (point, depth)
    unsigned long long* manScan = nullptr; // per pair (contacts << 32 | collided) and its exclusive scan
    unsigned long long* manScanOut = nullptr;
    uint2* collisions = nullptr;      // colliding collider pairs (one per manifold)
    uint4* manifolds = nullptr;       // (bodyA, bodyB, contactStart, contactCount)
    float4* contacts = nullptr;       // 2 x float4 per contact: (point, depth), (normal, friction_restitution bits)
    uint2* contactBodies = nullptr;

    // ---- solver
    uint32_t* unitColor[kNumGroups] = {};   // colour / level per unit
    uint32_t* unitOrder[kNumGroups] = {};   // units sorted by colour (solve order)
    uint2* unitDep[kNumGroups] = {};        // dataflow record per solve position: rank | degree << 16 for each body
    // non-collision interactions: allocated on the first upload that contains a trigger / force-field collider
    uint32_t numNonCollisionColliders = 0; uint4* interactions = nullptr; uint32_t maxInteractions = 0;
    float4* forceFields = nullptr; uint32_t numForceFields = 0, maxForceFields = 0; float globalForce[3] = { 0.f, 0.f, 0.f };
    std::vector<unsigned long long>* prevTriggerOverlaps = nullptr;   // host side of the trigger enter / leave diff
    // collision begin / end events (allocated by b200phys_enable_events)
    bool eventsEnabled = false; int evParity = 0;
    unsigned long long* evKeys[2] = {}; uint2* evPairs[2] = {}; uint32_t* evFlag[2] = {}; uint32_t* evSlot[2] = {};
    b200phys_collision_begin_event* evBegin = nullptr; b200phys_collider_pair* evEnd = nullptr;
    unsigned long long* scanStatus = nullptr; uint32_t* scanState = nullptr; uint32_t scanStatusTiles = 0, scanEpoch = 0;   // single-pass scan
    unsigned long long* refKeys[2] = {}; uint8_t* refColors = nullptr; uint8_t* colorKey = nullptr; int refParity = 0, colorAge = 0;   // colour reuse across steps
    uint8_t* refColorsJoint[6] = {}; uint32_t jointColorCount[6] = {}; int jointColorAge[6] = {};
    // the joint graph is static between uploads: solve order, dataflow ranks and the per-body unit counts of the joint groups are kept
    // (fast mode) as long as neither the joints' body pairs nor the body set changed
    uint32_t* bodyBaseJoints = nullptr; uint32_t bodiesEpoch = 1, jointsEpoch = 1, jointCacheBodiesEpoch = 0, jointCacheJointsEpoch = 0, jointCacheNumBodies = 0;
    std::vector<b200phys_body_pair>* jointPairsHost = nullptr;      // [6] last uploaded pairs (to tell a parameter refresh from a topology change)
    // the per-step accumulator upload runs on its own stream so that it overlaps the collision detection (b200phys_upload_forces)
    cudaStream_t copyStream = nullptr; cudaEvent_t evCopyGate = nullptr, evForcesDone = nullptr; bool forcesPending = false; void* forceStaging = nullptr;
    uint32_t* colorQueue = nullptr;         // ready queue of the colouring (max units of any group)
    unsigned long long* bodyMask = nullptr; // colours taken at each body (current group)
    uint32_t* bodyBase = nullptr;           // units per body in the groups coloured so far
    uint32_t* unitPending = nullptr;        // colouring: unresolved higher-ranked neighbours per unit
    unsigned long long* unitForbid = nullptr; // colouring: colour mask (fast) / level lower bound (exact) per unit
    uint32_t* nbrCount = nullptr;           // colouring: number of units waiting for this unit
    uint32_t* nbrList = nullptr;            // colouring: 16 explicit waiting units per unit
    uint32_t* adjStart = nullptr;           // N + 2
    uint32_t* adjFill = nullptr;
    uint32_t* adjList = nullptr;            // 2 * max units
    uint32_t* adjPos = nullptr;             // 2 * max units: solve position of the unit behind each adjList entry (inherit mode)
    uint2* unitSlots = nullptr;             // max units: the unit's slots in its two bodies' lists
    uint32_t* levelStart = nullptr;         // 7 x (kMaxLevels + 1): colour ranges per group
    uint32_t* levelFill = nullptr;          // 3 x (kMaxLevels + 1): per-round progress, colour histogram, scatter cursors
    uint32_t* slotOfContact = nullptr;      // solve slot -> canonical contact index
    uint32_t* unitSlotStart = nullptr;      // per sorted manifold: first solve slot
    uint4* unitInfo = nullptr;              // per sorted manifold: (bodyA, bodyB, first slot, contact count)
    float4* rows = nullptr;                 // 8 planes x maxContacts
    float2* rowImp = nullptr;               // accumulated (normal, tangent) impulse per slot
    void* jointIn[6] = {};                  // uploaded joint structs
    uint2* jointBodies[6] = {};
    void* jointRows[6] = {};                // solver rows per kind (structs in k_joints.cuh)

    // ---- scratch
    uint32_t* scanTmp = nullptr;            // block sums of the scan utility
    unsigned long long* scanTmp64 = nullptr;
    uint32_t* sortHist = nullptr;           // numKeys x numWarps histograms of the stable small-key sort
    uint8_t* staging = nullptr;             // host<->device staging (device side)
    size_t stagingBytes = 0;
    void* hostStaging = nullptr;            // pinned
    size_t hostStagingBytes = 0;

    DeviceCounts* counts = nullptr;         // device
    DeviceCounts* hostCounts = nullptr;     // pinned mirror (filled by get_stats / downloads)

    // capacities
    uint32_t maxPairs = 0, maxManifolds = 0, maxContacts = 0, maxJoints = 0, maxCells = 0;

    // heightmap (b200phys_upload_heightmap): heights + min/max mip pyramid per chunk on the device, parameters on the host
    uint16_t* hmHeights = nullptr; void* hmMips = nullptr;
    struct { uint32_t chunksPerDim; float chunkSize, invChunkSize, chunkScale, heightScale, invAmplitudeScale, minCorner[3], restitution, friction; } hm = {};

    void* cloths = nullptr;             // cloth instances (k_cloth.cu)
    void* env = nullptr;                // device-side locomotion environment tables (k_env.cu)

    // CUDA graph replay of the step (b200phys_api.cu): 0 = launch, 1 = dry run that only hashes what would be launched
    int graphMode = 0; unsigned long long graphHash = 0; void* graphCache = nullptr; bool graphFailed = false, graphsOff = false; uint32_t graphReplays = 0, graphCaptures = 0;

    // halo (multi-GPU)
    void* dist = nullptr;               // NCCL communicator + slab + message buffers of a context made by b200phys_create_distributed (k_dist.cu)
    uint8_t* haloBuf = nullptr; uint32_t haloCap = 0;
    uint32_t numGhostBodies = 0, numGhostColliders = 0;
};

// ---- every stream operation of the step goes through these wrappers: in a dry run (graphMode 1) nothing is enqueued, the kernel, its
// launch geometry and the VALUES of its arguments are hashed instead; the hash of a whole step identifies a captured CUDA graph.
static inline void graphHashBytes(Context* c, const void* p, size_t n)
{
    unsigned long long h = c->graphHash; const unsigned char* b = (const unsigned char*)p;
    for (size_t i = 0; i < n; ++i) { h ^= b[i]; h *= 1099511628211ull; }
    c->graphHash = h;
}
template <class T> static inline void graphHashValue(Context* c, T v) { graphHashBytes(c, &v, sizeof(T)); }
template <class... KArgs, class... Args>
static inline void launchK(Context* c, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args)
{
    if (c->graphMode == 1)
    {
        graphHashValue(c, (const void*)kernel); graphHashValue(c, grid.x); graphHashValue(c, grid.y); graphHashValue(c, block.x); graphHashValue(c, smem);
        (graphHashValue<KArgs>(c, static_cast<KArgs>(args)), ...);
        return;
    }
    kernel<<<grid, block, smem, c->stream>>>(static_cast<KArgs>(args)...);
}
static inline void memsetK(Context* c, void* p, int value, size_t bytes)
{
    if (c->graphMode == 1) { graphHashValue(c, p); graphHashValue(c, value); graphHashValue(c, bytes); return; }
    cudaMemsetAsync(p, value, bytes, c->stream);
}
static inline void memcpyD2DK(Context* c, void* dst, const void* src, size_t bytes)
{
    if (c->graphMode == 1) { graphHashValue(c, dst); graphHashValue(c, src); graphHashValue(c, bytes); return; }
    cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, c->stream);
}

// records the end of a kernel group on the context's stream when timing is enabled (no-op otherwise)
static inline void timelineMark(Context* c, const char* name)
{
    if (!c->timing || c->numMarks >= Context::kMaxMarks) return;
    if (!c->markEv[c->numMarks]) cudaEventCreate(&c->markEv[c->numMarks]);
    cudaEventRecord(c->markEv[c->numMarks], c->stream); c->markName[c->numMarks] = name; ++c->numMarks;
}

// launch geometry shared by every grid-stride kernel: a multiple of the SM count (148 on B200)
static inline int gridFor(const Context* c, int blocksPerSM = 8) { return c->numSMs * blocksPerSM; }

void destroyDist(Context* c);   // k_dist.cu

// ---- stage launchers (one per .cu file) ------------------------------------------------------------------------
void launchUnpackBodies(Context* c, const void* devStaging, uint32_t count);
void launchUnpackForces(Context* c, const void* devStaging, uint32_t count, cudaStream_t stream);
void waitPendingForceUpload(Context* c);    // orders the context's stream behind an accumulator upload still running on the copy stream
void launchUnpackBodyStates(Context* c, const void* devStaging, uint32_t count);
void launchPackBodyStates(Context* c, void* devStaging, uint32_t count);
void launchIntegrateForces(Context* c, float dt);
void launchWorldColliders(Context* c);
void launchIntegrateVelocities(Context* c, float dt);
void launchBroadphase(Context* c);
void launchNarrowphase(Context* c);
void launchSolver(Context* c, float dt, uint32_t iterations, uint32_t flags);
void launchPackAabbs(Context* c, float* devOut);

// utilities (b200_util.cu)
void launchCollisionEvents(Context* c);
void launchHeightmap(Context* c);
void destroyEnv(Context* c);             // k_env.cu
void destroyCloths(Context* c);          // k_cloth.cu        // k_heightmap.cuh (compiled with k_narrow.cu)
void scanExclusiveU32(Context* c, const uint32_t* in, uint32_t* out, const uint32_t* devCount, uint32_t maxCount, uint32_t* totalOut /*device, may be null*/, bool clearInput = false /* leave the scanned part of `in` zeroed */);
void scanExclusiveU64(Context* c, const unsigned long long* in, unsigned long long* out, const uint32_t* devCount, uint32_t maxCount, unsigned long long* totalOut);
// stable counting sort of n (device count) elements by small keys (< numKeys <= 64): dst[i] = destination slot, keyStart[k] = ranges
void stableSmallKeySort(Context* c, const uint8_t* keys, const uint32_t* devCount, uint32_t maxCount, uint32_t numKeys, uint32_t* dst, uint32_t* keyStart);

} // namespace b200
