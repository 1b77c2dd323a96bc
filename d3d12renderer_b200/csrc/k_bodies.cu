// k_bodies.cu -- per-body and per-collider streaming kernels (HBM-bound, one pass each):
//   unpackBodies / packBodyStates  : boundary AoS (b200phys_body, 140 B) <-> device float4 planes, staged through shared memory
//   integrateForces                : rigid_body_component::applyGravityAndIntegrateForces (reference rigid_body.cpp:95-124)
//   worldColliders                 : getWorldSpaceColliders (reference physics.cpp:631-756): world shape + world AABB
//   integrateVelocities            : rigid_body_component::integrateVelocity (reference rigid_body.cpp:126-142)
// All are grid-stride over a multiple of the SM count with 128-bit coalesced accesses on the SoA planes.
#include "b200_context.h"
#include "b200_util.cuh"

namespace b200 {

static const int kThreads = 256;
static const int kBodyFloats = 35;   // sizeof(b200phys_body) / 4

// ---------------------------------------------------------------------------------------------------------- unpack
__global__ void __launch_bounds__(kThreads) unpackBodiesKernel(const float* __restrict__ src, uint32_t count,
    float4* __restrict__ bRot, float4* __restrict__ bPos, float4* __restrict__ bCog, float4* __restrict__ bInvI,
    float4* __restrict__ bVel, float4* __restrict__ bAng, float4* __restrict__ bForce, float4* __restrict__ bTorque, uint32_t planeStride)
{
    __shared__ float s[kThreads * kBodyFloats];
    for (uint32_t base = blockIdx.x * kThreads; base < count; base += gridDim.x * kThreads)
    {
        const uint32_t nHere = min((uint32_t)kThreads, count - base);
        const float* g = src + (size_t)base * kBodyFloats;
        for (uint32_t k = threadIdx.x; k < nHere * kBodyFloats; k += kThreads) s[k] = g[k];   // coalesced
        __syncthreads();
        if (threadIdx.x < nHere)
        {
            const float* b = s + threadIdx.x * kBodyFloats;   // stride 35 words: conflict-free
            const uint32_t i = base + threadIdx.x;
            bRot[i] = make_float4(b[0], b[1], b[2], b[3]);
            bPos[i] = make_float4(b[4], b[5], b[6], b[10]);                 // position, invMass
            bCog[i] = make_float4(b[7], b[8], b[9], b[20]);                 // localCOG, gravityFactor
            bInvI[i] = make_float4(b[11], b[12], b[13], b[21]);             // column 0, linearDamping
            bInvI[planeStride + i] = make_float4(b[14], b[15], b[16], b[22]); // column 1, angularDamping
            bInvI[2 * planeStride + i] = make_float4(b[17], b[18], b[19], 0.f);
            bVel[i] = make_float4(b[23], b[24], b[25], 0.f);
            bAng[i] = make_float4(b[26], b[27], b[28], 0.f);
            bForce[i] = make_float4(b[29], b[30], b[31], 0.f);
            bTorque[i] = make_float4(b[32], b[33], b[34], 0.f);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads) unpackForcesKernel(const float* __restrict__ src, uint32_t count, float4* __restrict__ bForce, float4* __restrict__ bTorque)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < count; i += gridDim.x * kThreads)
    {
        const float* f = src + (size_t)i * 6;
        bForce[i] = make_float4(f[0], f[1], f[2], 0.f);
        bTorque[i] = make_float4(f[3], f[4], f[5], 0.f);
    }
}

__global__ void __launch_bounds__(kThreads) unpackBodyStatesKernel(const float* __restrict__ src, uint32_t count,
    float4* __restrict__ bRot, float4* __restrict__ bPos, float4* __restrict__ bVel, float4* __restrict__ bAng)
{
    __shared__ float s[kThreads * 13];
    for (uint32_t base = blockIdx.x * kThreads; base < count; base += gridDim.x * kThreads)
    {
        const uint32_t nHere = min((uint32_t)kThreads, count - base);
        const float* g = src + (size_t)base * 13;
        for (uint32_t k = threadIdx.x; k < nHere * 13; k += kThreads) s[k] = g[k];
        __syncthreads();
        if (threadIdx.x < nHere)
        {
            const float* b = s + threadIdx.x * 13;
            const uint32_t i = base + threadIdx.x;
            bRot[i] = make_float4(b[0], b[1], b[2], b[3]);
            const float invMass = bPos[i].w;
            bPos[i] = make_float4(b[4], b[5], b[6], invMass);
            bVel[i] = make_float4(b[7], b[8], b[9], 0.f);
            bAng[i] = make_float4(b[10], b[11], b[12], 0.f);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads) packBodyStatesKernel(float* __restrict__ dst, uint32_t count,
    const float4* __restrict__ bRot, const float4* __restrict__ bPos, const float4* __restrict__ bVel, const float4* __restrict__ bAng)
{
    __shared__ float s[kThreads * 13];
    for (uint32_t base = blockIdx.x * kThreads; base < count; base += gridDim.x * kThreads)
    {
        const uint32_t nHere = min((uint32_t)kThreads, count - base);
        if (threadIdx.x < nHere)
        {
            const uint32_t i = base + threadIdx.x;
            float4 r = bRot[i], p = bPos[i], v = bVel[i], w = bAng[i];
            float* o = s + threadIdx.x * 13;
            o[0] = r.x; o[1] = r.y; o[2] = r.z; o[3] = r.w; o[4] = p.x; o[5] = p.y; o[6] = p.z;
            o[7] = v.x; o[8] = v.y; o[9] = v.z; o[10] = w.x; o[11] = w.y; o[12] = w.z;
        }
        __syncthreads();
        float* g = dst + (size_t)base * 13;
        for (uint32_t k = threadIdx.x; k < nHere * 13; k += kThreads) g[k] = s[k];
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------- integrate forces
// reads 10 float4 planes, writes the 96-byte global state (gVel 32 B AoS, gPos 16 B, gInvI 48 B AoS) per body
__global__ void __launch_bounds__(kThreads) integrateForcesKernel(uint32_t N, float dt, uint32_t planeStride, float gfx, float gfy, float gfz,
    const float4* __restrict__ bRot, const float4* __restrict__ bPos, const float4* __restrict__ bCog, const float4* __restrict__ bInvI,
    const float4* __restrict__ bVel, const float4* __restrict__ bAng, float4* __restrict__ bForce, const float4* __restrict__ bTorque,
    float4* __restrict__ gVel, float4* __restrict__ gPos, float4* __restrict__ gInvI)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i <= N; i += gridDim.x * kThreads)
    {
        if (i == N)
        {   // kinematic stand-in for static colliders (physics.cpp:1279)
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            gVel[2 * i] = z; gVel[2 * i + 1] = z; gPos[i] = z; gInvI[3 * i] = z; gInvI[3 * i + 1] = z; gInvI[3 * i + 2] = z;
            continue;
        }
        const float4 rot4 = ldStream(bRot + i), pos4 = ldStream(bPos + i), cog4 = ldStream(bCog + i);
        const float4 c0 = ldStream(bInvI + i), c1 = ldStream(bInvI + planeStride + i), c2 = ldStream(bInvI + 2 * planeStride + i);
        const float4 v4 = ldStream(bVel + i), w4 = ldStream(bAng + i), F4 = bForce[i], T4 = ldStream(bTorque + i);
        const q4 rot = toq(rot4);
        const float invMass = pos4.w, gravityFactor = cog4.w, linDamp = c0.w, angDamp = c1.w;

        const f3 cogPos = xyz(pos4) + rot * xyz(cog4);
        const m3 R = quaternionToMat3(rot);
        m3 I; I.m[0] = c0.x; I.m[1] = c0.y; I.m[2] = c0.z; I.m[3] = c1.x; I.m[4] = c1.y; I.m[5] = c1.z; I.m[6] = c2.x; I.m[7] = c2.y; I.m[8] = c2.z;
        const m3 Iw = R * I * transpose(R);

        f3 F = xyz(F4) + mk3(gfx, gfy, gfz);      // global force fields (physics.cpp:1273)
        if (invMass > 0.f) F.y += (-9.81f / invMass * gravityFactor);   // GRAVITY, physics.h:11
        const f3 linAcc = F * invMass;
        const f3 angAcc = Iw * xyz(T4);
        f3 v = xyz(v4), w = xyz(w4);
        v += linAcc * dt; w += angAcc * dt;
        v *= 1.f / (1.f + dt * linDamp);
        w *= 1.f / (1.f + dt * angDamp);

        gVel[2 * i] = mk4(v, 0.f);          // .w of both halves: update counter of the dataflow solver (k_solver.cu), 0 = start of step
        gVel[2 * i + 1] = mk4(w, 0.f);
        gPos[i] = mk4(cogPos, invMass);
        gInvI[3 * i] = mk4(m3Row(Iw, 0), 0.f);
        gInvI[3 * i + 1] = mk4(m3Row(Iw, 1), 0.f);
        gInvI[3 * i + 2] = mk4(m3Row(Iw, 2), 0.f);
        bForce[i] = mk4(F, 0.f);   // the reference adds gravity into the accumulator (rigid_body.cpp:107); cleared again in integrateVelocity
    }
}

// ------------------------------------------------------------------------------------------------- world colliders
struct BoxMM { f3 mn, mx; };
B200_D void growBox(BoxMM& b, f3 o) { b.mn = vmin(b.mn, o); b.mx = vmax(b.mx, o); }
// bounding_box::transformToAABB (bounding_volumes.cpp:58-70): same corner order
B200_D BoxMM transformToAABB(f3 mn, f3 mx, q4 q, f3 t)
{
    BoxMM r; r.mn = mk3(FLT_MAX); r.mx = mk3(-FLT_MAX);
    growBox(r, q * mn + t);
    growBox(r, q * mk3(mx.x, mn.y, mn.z) + t);
    growBox(r, q * mk3(mn.x, mx.y, mn.z) + t);
    growBox(r, q * mk3(mx.x, mx.y, mn.z) + t);
    growBox(r, q * mk3(mn.x, mn.y, mx.z) + t);
    growBox(r, q * mk3(mx.x, mn.y, mx.z) + t);
    growBox(r, q * mk3(mn.x, mx.y, mx.z) + t);
    growBox(r, q * mx + t);
    return r;
}

__global__ void __launch_bounds__(kThreads) worldCollidersKernel(uint32_t C, uint32_t N,
    const float4* __restrict__ cLocal, const float4* __restrict__ xforms, const float4* __restrict__ bRot, const float4* __restrict__ bPos,
    const float4* __restrict__ hullAabb, float4* __restrict__ wCol, float4* __restrict__ aabbMin, float4* __restrict__ aabbMax, float4* __restrict__ bSphere)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < C; i += gridDim.x * kThreads)
    {
        const float4 L0 = ldStream(cLocal + 4 * i), L1 = ldStream(cLocal + 4 * i + 1), L2 = ldStream(cLocal + 4 * i + 2), L3 = ldStream(cLocal + 4 * i + 3);
        const float s[10] = { L0.x, L0.y, L0.z, L0.w, L1.x, L1.y, L1.z, L1.w, L2.x, L2.y };
        const float restitution = L2.z, friction = L2.w;
        const uint32_t tw = __float_as_uint(L3.y);
        uint32_t type = tw & 0xFFu; const uint32_t objType = (tw >> 8) & 0xFFu;
        const uint32_t objIndex = __float_as_uint(L3.z), parent = __float_as_uint(L3.w);
        if (objType == B200PHYS_OBJ_NONE)
        {   // padding slot of a fixed-size halo message: an inverted box overlaps nothing and is never binned
            wCol[4 * i] = wCol[4 * i + 1] = wCol[4 * i + 2] = make_float4(0.f, 0.f, 0.f, 0.f); wCol[4 * i + 3] = make_float4(__uint_as_float(tw & 0xFFFFu), __uint_as_float(0xFFFFFFFFu), 0.f, 0.f);
            aabbMin[i] = make_float4(FLT_MAX, FLT_MAX, FLT_MAX, 0.f); aabbMax[i] = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, 0.f);
            bSphere[i] = make_float4(0.f, 0.f, 0.f, -1.f);
            continue;
        }

        // pose source (physics.cpp:645-647)
        q4 rot = qIdentity(); f3 pos = mk3(0.f);
        if (objType == B200PHYS_OBJ_RIGID_BODY) { rot = toq(bRot[objIndex]); pos = xyz(bPos[objIndex]); }
        else if (parent != B200PHYS_NO_TRANSFORM) { rot = toq(xforms[2 * parent]); pos = xyz(xforms[2 * parent + 1]); }

        float o[10] = { 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f };
        uint32_t geom = 0;
        BoxMM bb;
        // a sphere around the shape (centre, radius): npClassifyKernel drops pairs whose spheres are clearly apart before any routine runs
        f3 bsC; float bsR;
        switch (type)
        {
        case B200PHYS_SPHERE:
        {
            const f3 c = pos + rot * mk3(s[0], s[1], s[2]);
            bb.mn = c - mk3(s[3]); bb.mx = c + mk3(s[3]);
            o[0] = c.x; o[1] = c.y; o[2] = c.z; o[3] = s[3];
            bsC = c; bsR = s[3];
        } break;
        case B200PHYS_CAPSULE:
        {
            const f3 a = rot * mk3(s[0], s[1], s[2]) + pos, b = rot * mk3(s[3], s[4], s[5]) + pos;
            const f3 r3 = mk3(s[6]);
            bb.mn = mk3(FLT_MAX); bb.mx = mk3(-FLT_MAX);
            growBox(bb, a + r3); growBox(bb, a - r3); growBox(bb, b + r3); growBox(bb, b - r3);
            o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z; o[6] = s[6];
            bsC = (a + b) * 0.5f; bsR = 0.5f * sqrtf(dot(b - a, b - a)) + s[6];
        } break;
        case B200PHYS_CYLINDER:
        {
            const f3 a = rot * mk3(s[0], s[1], s[2]) + pos, b = rot * mk3(s[3], s[4], s[5]) + pos;
            const float radius = s[6];
            const f3 d = b - a;
            const float aa = dot(d, d);
            float x = 1.f - d.x * d.x / aa, y = 1.f - d.y * d.y / aa, z = 1.f - d.z * d.z / aa;
            x = sqrtf(fmax2(0.f, x)); y = sqrtf(fmax2(0.f, y)); z = sqrtf(fmax2(0.f, z));
            const f3 e = radius * mk3(x, y, z);
            bb.mn = vmin(a - e, b - e); bb.mx = vmax(a + e, b + e);
            o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z; o[6] = radius;
            bsC = (a + b) * 0.5f; bsR = sqrtf(0.25f * aa + radius * radius);
        } break;
        case B200PHYS_AABB:
        {
            const f3 mn = mk3(s[0], s[1], s[2]), mx = mk3(s[3], s[4], s[5]);
            bb = transformToAABB(mn, mx, rot, pos);
            if (qIsIdentity(rot)) { o[0] = bb.mn.x; o[1] = bb.mn.y; o[2] = bb.mn.z; o[3] = bb.mx.x; o[4] = bb.mx.y; o[5] = bb.mx.z; }
            else
            {   // an AABB under rotation becomes an OBB (physics.cpp:729-733, bounding_volumes.cpp:72-79)
                type = B200PHYS_OBB;
                const f3 c = rot * ((mn + mx) * 0.5f) + pos, r = (mx - mn) * 0.5f;
                o[0] = rot.x; o[1] = rot.y; o[2] = rot.z; o[3] = rot.w; o[4] = c.x; o[5] = c.y; o[6] = c.z; o[7] = r.x; o[8] = r.y; o[9] = r.z;
            }
            bsC = (bb.mn + bb.mx) * 0.5f; bsR = 0.5f * sqrtf(dot(bb.mx - bb.mn, bb.mx - bb.mn));      // (around the world box: contains the shape in both cases)
        } break;
        case B200PHYS_OBB:
        {
            const q4 oq = mkq(s[0], s[1], s[2], s[3]); const f3 oc = mk3(s[4], s[5], s[6]), rad = mk3(s[7], s[8], s[9]);
            const q4 q = rot * oq; const f3 c = rot * oc + pos;
            bb = transformToAABB(-rad, rad, q, c);
            o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w; o[4] = c.x; o[5] = c.y; o[6] = c.z; o[7] = rad.x; o[8] = rad.y; o[9] = rad.z;
            bsC = c; bsR = sqrtf(dot(rad, rad));
        } break;
        default:
        {
            geom = __float_as_uint(s[7]);
            const q4 q = rot * mkq(s[0], s[1], s[2], s[3]); const f3 p = rot * mk3(s[4], s[5], s[6]) + pos;
            bb = transformToAABB(xyz(hullAabb[2 * geom]), xyz(hullAabb[2 * geom + 1]), q, p);
            o[0] = q.x; o[1] = q.y; o[2] = q.z; o[3] = q.w; o[4] = p.x; o[5] = p.y; o[6] = p.z;
            // hullAabb[2 g].w = largest distance of a vertex from the centre of the local box (b200phys_set_hulls)
            const float4 hmn = hullAabb[2 * geom], hmx = hullAabb[2 * geom + 1];
            bsC = q * ((xyz(hmn) + xyz(hmx)) * 0.5f) + p; bsR = hmn.w;
        } break;
        }
        const uint32_t bodyIndex = (objType == B200PHYS_OBJ_STATIC) ? N : objIndex;   // dummy for static colliders (physics.cpp:669)
        wCol[4 * i] = make_float4(o[0], o[1], o[2], o[3]);
        wCol[4 * i + 1] = make_float4(o[4], o[5], o[6], o[7]);
        wCol[4 * i + 2] = make_float4(o[8], o[9], restitution, friction);
        wCol[4 * i + 3] = make_float4(__uint_as_float(type | (objType << 8)), __uint_as_float(bodyIndex), __uint_as_float(geom), 0.f);
        aabbMin[i] = mk4(bb.mn, 0.f);
        aabbMax[i] = mk4(bb.mx, 0.f);
        bSphere[i] = mk4(bsC, bsR);
    }
}

// --------------------------------------------------------------------------------------------- integrate velocities
__global__ void __launch_bounds__(kThreads) integrateVelocitiesKernel(uint32_t N, float dt,
    const float4* __restrict__ gVel, const float4* __restrict__ gPos, const float4* __restrict__ bCog,
    float4* __restrict__ bRot, float4* __restrict__ bPos, float4* __restrict__ bVel, float4* __restrict__ bAng, float4* __restrict__ bForce, float4* __restrict__ bTorque)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < N; i += gridDim.x * kThreads)
    {
        const float4 v4 = ldStream(gVel + 2 * i), w4 = ldStream(gVel + 2 * i + 1), p4 = ldStream(gPos + i), cog4 = ldStream(bCog + i);
        const q4 rot0 = toq(bRot[i]);
        const f3 v = xyz(v4), w = xyz(w4);
        q4 dq = mkq(0.5f * w.x, 0.5f * w.y, 0.5f * w.z, 0.f);
        dq = dq * rot0;
        const q4 rot = normalize(mkq(rot0.x + dq.x * dt, rot0.y + dq.y * dt, rot0.z + dq.z * dt, rot0.w + dq.w * dt));
        const f3 pos = xyz(p4) + v * dt;
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        bRot[i] = mk4(rot);
        bPos[i] = mk4(pos - rot * xyz(cog4), p4.w);   // .w keeps invMass
        bVel[i] = mk4(v, 0.f); bAng[i] = mk4(w, 0.f);
        bForce[i] = z; bTorque[i] = z;
    }
}

__global__ void __launch_bounds__(kThreads) packAabbsKernel(uint32_t C, const float4* __restrict__ mn, const float4* __restrict__ mx, float* __restrict__ out)
{
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < C; i += gridDim.x * kThreads)
    {
        float4 a = mn[i], b = mx[i]; float* o = out + (size_t)i * 6;
        o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = b.x; o[4] = b.y; o[5] = b.z;
    }
}

// ------------------------------------------------------------------------------------------------------- launchers
void launchUnpackBodies(Context* c, const void* dev, uint32_t count)
{
    if (!count) return;
    const uint32_t stride = c->cfg.maxBodies + 1;
    launchK(c, unpackBodiesKernel, gridFor(c, 4), kThreads, 0, (const float*)dev, count, c->bRot, c->bPos, c->bCog, c->bInvI, c->bVel, c->bAng, c->bForce, c->bTorque, stride);
    c->launches++;
}
void waitPendingForceUpload(Context* c)
{
    if (c->forcesPending) { cudaStreamWaitEvent(c->stream, c->evForcesDone, 0); c->forcesPending = false; }
}
void launchUnpackForces(Context* c, const void* dev, uint32_t count, cudaStream_t stream)
{
    if (!count) return;
    unpackForcesKernel<<<gridFor(c, 4), kThreads, 0, stream>>>((const float*)dev, count, c->bForce, c->bTorque);
    c->launches++;
}
void launchUnpackBodyStates(Context* c, const void* dev, uint32_t count)
{
    if (!count) return;
    launchK(c, unpackBodyStatesKernel, gridFor(c, 4), kThreads, 0, (const float*)dev, count, c->bRot, c->bPos, c->bVel, c->bAng);
    c->launches++;
}
void launchPackBodyStates(Context* c, void* dev, uint32_t count)
{
    if (!count) return;
    launchK(c, packBodyStatesKernel, gridFor(c, 4), kThreads, 0, (float*)dev, count, c->bRot, c->bPos, c->bVel, c->bAng);
    c->launches++;
}
void launchIntegrateForces(Context* c, float dt)
{
    const uint32_t stride = c->cfg.maxBodies + 1;
    launchK(c, integrateForcesKernel, gridFor(c, 8), kThreads, 0, c->numBodies, dt, stride, c->globalForce[0], c->globalForce[1], c->globalForce[2], c->bRot, c->bPos, c->bCog, c->bInvI, c->bVel, c->bAng, c->bForce, c->bTorque, c->gVel, c->gPos, c->gInvI);
    c->launches++;
}
void launchWorldColliders(Context* c)
{
    if (!c->numColliders) return;
    launchK(c, worldCollidersKernel, gridFor(c, 8), kThreads, 0, c->numColliders, c->numBodies, c->cLocal, c->xforms, c->bRot, c->bPos, c->hullAabb, c->wCol, c->aabbMin, c->aabbMax, c->bSphere);
    c->launches++;
}
void launchIntegrateVelocities(Context* c, float dt)
{
    if (!c->numBodies) return;
    launchK(c, integrateVelocitiesKernel, gridFor(c, 8), kThreads, 0, c->numBodies, dt, c->gVel, c->gPos, c->bCog, c->bRot, c->bPos, c->bVel, c->bAng, c->bForce, c->bTorque);
    c->launches++;
}
void launchPackAabbs(Context* c, float* devOut)
{
    if (!c->numColliders) return;
    launchK(c, packAabbsKernel, gridFor(c, 4), kThreads, 0, c->numColliders, c->aabbMin, c->aabbMax, devOut);
    c->launches++;
}

} // namespace b200
