// k_cloth.cu -- the reference's cloth (src/physics/cloth.{h,cpp}) on the device: grids of particles with stretch / shear /
// bend distance constraints, solved by Gauss-Seidel sweeps IN CONSTRAINT CREATION ORDER (velocity, position, drift passes, cloth.cpp:
// 191-317), simulated at the end of physicsStepInternal after the global force field has acted as wind (physics.cpp:1354-1358).
// One CTA per cloth. The sequential sweep order is kept exactly: the constraint topology is static, so the host computes once, at creation,
// the wavefront level of every constraint (1 + the highest level of an earlier constraint that shares one of its two particles); inside a
// level no two constraints share a particle, so the threads of the CTA take a level's constraints in parallel and a __syncthreads()
// separates levels -- every particle sees its constraints in creation order, the result equals the sequential sweep bit for bit. The wind
// pass is per particle: a particle adds the contributions of its (up to six) triangles in the order the reference's quad loop produces them.
#include "b200_context.h"
#include "b200_util.cuh"
#include "../../include/b200phys.h"
#include <cstring>
#include <vector>
#include <algorithm>

namespace b200 {

struct ClothConstraint { uint32_t a, b; float restDistance, inverseMassSum; };

struct ClothDev
{
    uint32_t gx, gy, n, numConstraints, numLevels;
    float width, height, totalMass, gravityFactor, damping, stiffness, oldTotalMass, oldStiffness;
    f3* positions; f3* prevPositions; f3* velocities; f3* forces; float* invMasses;
    ClothConstraint* constraints;       // sorted by (level, creation index)
    float4* temp;                       // per constraint: gradient.xyz, inverseScaledGradientSquared (velocity pass)
    uint32_t* levelStart;               // numLevels + 1
};

struct ClothSet { std::vector<ClothDev> host; ClothDev* dev = nullptr; uint32_t devCapacity = 0; };

B200_D f3 clothParticlePosition(const ClothDev& c, float relX, float relY)
{   // cloth.cpp:132-138
    f3 p = mk3(relX * c.width, -relY * c.height, 0.f);
    p.x -= c.width * 0.5f;
    const float t = p.y; p.y = p.z; p.z = t;
    return p;
}

// cloth.cpp:300-317 for the constraints [c0, c1) of one level
B200_D void clothSolvePositionsLevel(const ClothDev& c, uint32_t c0, uint32_t c1)
{
    for (uint32_t i = c0 + threadIdx.x; i < c1; i += blockDim.x)
    {
        const ClothConstraint k = c.constraints[i];
        if (k.inverseMassSum > 0.f)
        {
            const f3 delta = c.positions[k.b] - c.positions[k.a];
            const float l = sqLen(delta);
            const float sqRest = k.restDistance * k.restDistance;
            if (sqRest + l > 1e-5f)
            {
                const float kk = ((sqRest - l) / (k.inverseMassSum * (sqRest + l)));
                c.positions[k.a] -= delta * (kk * c.invMasses[k.a]);
                c.positions[k.b] += delta * (kk * c.invMasses[k.b]);
            }
        }
    }
}
B200_D void clothSolvePositions(const ClothDev& c)
{
    for (uint32_t l = 0; l < c.numLevels; ++l) { clothSolvePositionsLevel(c, c.levelStart[l], c.levelStart[l + 1]); __syncthreads(); }
}

// cloth_component::simulate (cloth.cpp:191-286); properties were refreshed on the host (recalculateProperties, :330-346)
__global__ void __launch_bounds__(256) clothSimulateKernel(const ClothDev* __restrict__ cloths, uint32_t velocityIterations, uint32_t positionIterations, uint32_t driftIterations, float dt)
{
    const ClothDev c = cloths[blockIdx.x];
    const float gravityVelocity = -9.81f * dt * c.gravityFactor;
    for (uint32_t i = threadIdx.x; i < c.n; i += blockDim.x)
    {
        f3 v = c.velocities[i]; const float im = c.invMasses[i];
        if (im > 0.f) v.y += gravityVelocity;
        v += c.forces[i] * (im * dt);
        c.velocities[i] = v;
        const f3 p = c.positions[i];
        c.prevPositions[i] = p;
        c.positions[i] = p + v * dt;
        c.forces[i] = mk3(0.f);
    }
    __syncthreads();
    const float invDt = (dt > 1e-5f) ? (1.f / dt) : 1.f;
    if (velocityIterations > 0)
    {
        for (uint32_t i = threadIdx.x; i < c.numConstraints; i += blockDim.x)
        {
            const ClothConstraint k = c.constraints[i];
            const f3 g = c.prevPositions[k.b] - c.prevPositions[k.a];
            const float inv = (k.inverseMassSum == 0.f) ? 0.f : (1.f / (sqLen(g) * k.inverseMassSum));
            c.temp[i] = make_float4(g.x, g.y, g.z, inv);
        }
        __syncthreads();
        for (uint32_t it = 0; it < velocityIterations; ++it)
            for (uint32_t l = 0; l < c.numLevels; ++l)
            {   // cloth.cpp:288-298
                for (uint32_t i = c.levelStart[l] + threadIdx.x; i < c.levelStart[l + 1]; i += blockDim.x)
                {
                    const ClothConstraint k = c.constraints[i]; const float4 t = c.temp[i]; const f3 g = mk3(t.x, t.y, t.z);
                    const float j = -dot(g, c.velocities[k.a] - c.velocities[k.b]) * t.w;
                    c.velocities[k.a] += g * (j * c.invMasses[k.a]);
                    c.velocities[k.b] -= g * (j * c.invMasses[k.b]);
                }
                __syncthreads();
            }
        for (uint32_t i = threadIdx.x; i < c.n; i += blockDim.x) c.positions[i] = c.prevPositions[i] + c.velocities[i] * dt;
        __syncthreads();
    }
    if (positionIterations > 0)
    {
        for (uint32_t it = 0; it < positionIterations; ++it) clothSolvePositions(c);
        for (uint32_t i = threadIdx.x; i < c.n; i += blockDim.x) c.velocities[i] = (c.positions[i] - c.prevPositions[i]) * invDt;
        __syncthreads();
    }
    if (driftIterations > 0)
    {
        for (uint32_t i = threadIdx.x; i < c.n; i += blockDim.x) c.prevPositions[i] = c.positions[i];
        __syncthreads();
        for (uint32_t it = 0; it < driftIterations; ++it) clothSolvePositions(c);
        for (uint32_t i = threadIdx.x; i < c.n; i += blockDim.x) c.velocities[i] += (c.positions[i] - c.prevPositions[i]) * invDt;
        __syncthreads();
    }
    const float dampingFactor = 1.f / (1.f + dt * c.damping);
    for (uint32_t i = threadIdx.x; i < c.n; i += blockDim.x) c.velocities[i] *= dampingFactor;
}

// applyWindForce (cloth.cpp:145-183): the reference walks the quads row by row and adds each triangle's force to its three corners; here
// every particle adds the same terms in the same order (quads (px-1,py-1), (px,py-1), (px-1,py), (px,py); inside a quad triangle 1, then 2)
__global__ void __launch_bounds__(256) clothWindKernel(const ClothDev* __restrict__ cloths, uint32_t first, float fx, float fy, float fz)
{
    const ClothDev c = cloths[first + blockIdx.x];
    const f3 force = mk3(fx, fy, fz);
    for (uint32_t i = threadIdx.x; i < c.n; i += blockDim.x)
    {
        const uint32_t px = i % c.gx, py = i / c.gx;
        f3 acc = c.forces[i];
        for (int dy = -1; dy <= 0; ++dy)
            for (int dx = -1; dx <= 0; ++dx)
            {
                const int qx = (int)px + dx, qy = (int)py + dy;
                if (qx < 0 || qy < 0 || qx >= (int)c.gx - 1 || qy >= (int)c.gy - 1) continue;
                const uint32_t tl = (uint32_t)qy * c.gx + (uint32_t)qx, tr = tl + 1, bl = tl + c.gx, br = bl + 1;
                if (i == tl || i == tr || i == bl)
                {
                    const f3 normal = cross(c.positions[bl] - c.positions[tl], c.positions[tr] - c.positions[tl]);
                    f3 f = normal * dot(normalize(normal), force);
                    f *= 1.f / 3.f;
                    acc += f;
                }
                if (i == br || i == tr || i == bl)
                {
                    const f3 normal = cross(c.positions[tr] - c.positions[br], c.positions[bl] - c.positions[br]);
                    f3 f = normal * dot(normalize(normal), force);
                    f *= 1.f / 3.f;
                    acc += f;
                }
            }
        c.forces[i] = acc;
    }
}

// setWorldPositionOfFixedVertices (cloth.cpp:84-130); one CTA
__global__ void __launch_bounds__(256) clothSetFixedKernel(const ClothDev* __restrict__ cloths, uint32_t id, float4 rot4, float4 pos4, float4 scale4, int moveRigid)
{
    const ClothDev c = cloths[id];
    const q4 rot = mkq(rot4.x, rot4.y, rot4.z, rot4.w); const f3 pos = mk3(pos4.x, pos4.y, pos4.z), scale = mk3(scale4.x, scale4.y, scale4.z);
    auto xf = [&](f3 p) { return rot * (scale * p) + pos; };        // transformPosition (math.cpp:518-521)
    if (moveRigid)
    {
        f3 pivot;
        if (c.gx % 2 == 1) pivot = c.positions[c.gx / 2];
        else pivot = (c.positions[c.gx / 2] + c.positions[c.gx / 2 - 1]) * 0.5f;
        const f3 currentAxis = normalize(c.positions[c.gx - 1] - c.positions[0]);
        const f3 newAxis = normalize(xf(clothParticlePosition(c, 1.f, 0.f)) - xf(clothParticlePosition(c, 0.f, 0.f)));
        const f3 newPivot = xf(clothParticlePosition(c, 0.5f, 0.f));
        const q4 delta = rotateFromTo(currentAxis, newAxis);
        __syncthreads();                                            // everybody has read row 0 before anybody overwrites anything
        for (uint32_t i = c.gx + threadIdx.x; i < c.n; i += blockDim.x) c.positions[i] = delta * (c.positions[i] - pivot) + newPivot;
    }
    __syncthreads();
    for (uint32_t x = threadIdx.x; x < c.gx; x += blockDim.x) c.positions[x] = xf(clothParticlePosition(c, (float)x / (float)(c.gx - 1), 0.f));
}

void destroyCloths(Context* c)
{
    ClothSet* S = (ClothSet*)c->cloths; if (!S) return;
    for (ClothDev& d : S->host)
    {
        void* ptrs[] = { d.positions, d.prevPositions, d.velocities, d.forces, d.invMasses, d.constraints, d.temp, d.levelStart };
        for (void* p : ptrs) if (p) cudaFree(p);
    }
    if (S->dev) cudaFree(S->dev);
    delete S; c->cloths = nullptr;
}

static cudaError_t syncTable(ClothSet& S)
{
    if (S.host.size() > S.devCapacity)
    {
        if (S.dev) cudaFree(S.dev);
        S.devCapacity = (uint32_t)S.host.size() + 8;
        cudaError_t e = cudaMalloc((void**)&S.dev, sizeof(ClothDev) * S.devCapacity); if (e != cudaSuccess) return e;
    }
    return cudaMemcpy(S.dev, S.host.data(), sizeof(ClothDev) * S.host.size(), cudaMemcpyHostToDevice);
}

} // namespace b200

using namespace b200;
struct b200phys_context { Context c; };
#define CKC(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return -(int)e_; } while (0)
template <class T> static cudaError_t up(T*& dst, const std::vector<T>& src) { cudaError_t e = cudaMalloc((void**)&dst, sizeof(T) * (src.size() ? src.size() : 1)); if (e != cudaSuccess) return e; return src.size() ? cudaMemcpy(dst, src.data(), sizeof(T) * src.size(), cudaMemcpyHostToDevice) : cudaSuccess; }

extern "C" {

// cloth_component::cloth_component (cloth.cpp:7-82): particle grid, constraints in creation order, + the wavefront levels of that order
int b200phys_cloth_create(b200phys_context* ctx, const b200phys_cloth_desc* d, uint32_t* outId)
{
    if (!ctx || !d || !outId || d->gridSizeX < 2 || d->gridSizeY < 2 || !(d->totalMass > 0.f) || !(d->stiffness > 0.f)) return B200PHYS_E_INVALID;
    Context& c = ctx->c; CKC(cudaSetDevice(c.device));
    CKC(cudaStreamSynchronize(c.stream));
    if (!c.cloths) c.cloths = new ClothSet();
    ClothSet& S = *(ClothSet*)c.cloths;
    ClothDev k; memset(&k, 0, sizeof(k));
    k.gx = d->gridSizeX; k.gy = d->gridSizeY; k.n = k.gx * k.gy; k.width = d->width; k.height = d->height;
    k.totalMass = k.oldTotalMass = d->totalMass; k.stiffness = k.oldStiffness = d->stiffness; k.damping = d->damping; k.gravityFactor = d->gravityFactor;
    const float invMassPerParticle = k.n / k.totalMass;
    std::vector<f3> pos(k.n); std::vector<float> invMass(k.n);
    for (uint32_t y = 0; y < k.gy; ++y)
        for (uint32_t x = 0; x < k.gx; ++x)
        {
            const float relX = x / (float)(k.gx - 1), relY = y / (float)(k.gy - 1);
            f3 p = mk3(relX * k.width, -relY * k.height, 0.f); p.x -= k.width * 0.5f; std::swap(p.y, p.z);
            pos[y * k.gx + x] = p; invMass[y * k.gx + x] = (y == 0) ? 0.f : invMassPerParticle;
        }
    std::vector<ClothConstraint> cons;
    auto add = [&](uint32_t a, uint32_t b) { cons.push_back({ a, b, len(pos[a] - pos[b]), (invMass[a] + invMass[b]) / k.stiffness }); };
    for (uint32_t y = 0; y < k.gy; ++y)
        for (uint32_t x = 0; x < k.gx; ++x)
        {
            const uint32_t index = y * k.gx + x;
            if (x < k.gx - 1) add(index, index + 1);
            if (y < k.gy - 1) add(index, index + k.gx);
            if (x < k.gx - 1 && y < k.gy - 1) { add(index, index + k.gx + 1); add(index + k.gx, index + 1); }
            if (x < k.gx - 2) add(index, index + 2);
            if (y < k.gy - 2) add(index, index + k.gx * 2);
        }
    // wavefront levels of the sequential sweep
    std::vector<uint32_t> last(k.n, 0), level(cons.size());
    uint32_t numLevels = 0;
    for (size_t i = 0; i < cons.size(); ++i) { const uint32_t l = std::max(last[cons[i].a], last[cons[i].b]); level[i] = l; last[cons[i].a] = last[cons[i].b] = l + 1; numLevels = std::max(numLevels, l + 1); }
    std::vector<uint32_t> order(cons.size()); for (size_t i = 0; i < order.size(); ++i) order[i] = (uint32_t)i;
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return level[a] < level[b]; });
    std::vector<ClothConstraint> sorted(cons.size()); std::vector<uint32_t> levelStart(numLevels + 1, 0);
    for (size_t i = 0; i < order.size(); ++i) { sorted[i] = cons[order[i]]; levelStart[level[order[i]] + 1]++; }
    for (uint32_t l = 0; l < numLevels; ++l) levelStart[l + 1] += levelStart[l];
    k.numConstraints = (uint32_t)cons.size(); k.numLevels = numLevels;
    std::vector<f3> zeros(k.n, mk3(0.f));
    CKC(up(k.positions, pos)); CKC(up(k.prevPositions, pos)); CKC(up(k.velocities, zeros)); CKC(up(k.forces, zeros)); CKC(up(k.invMasses, invMass));
    CKC(up(k.constraints, sorted)); CKC(up(k.levelStart, levelStart));
    CKC(cudaMalloc((void**)&k.temp, sizeof(float4) * (cons.size() ? cons.size() : 1)));
    S.host.push_back(k);
    CKC(syncTable(S));
    *outId = (uint32_t)S.host.size() - 1;
    return 0;
}

int b200phys_cloth_count(b200phys_context* ctx, uint32_t* count) { if (!ctx || !count) return B200PHYS_E_INVALID; ClothSet* S = (ClothSet*)ctx->c.cloths; *count = S ? (uint32_t)S->host.size() : 0u; return 0; }

int b200phys_cloth_set_fixed_vertices(b200phys_context* ctx, uint32_t id, const float rotation[4], const float position[3], const float scale[3], int moveRigid)
{
    if (!ctx || !ctx->c.cloths || !rotation || !position || !scale) return B200PHYS_E_INVALID;
    Context& c = ctx->c; ClothSet& S = *(ClothSet*)c.cloths; CKC(cudaSetDevice(c.device));
    if (id >= S.host.size()) return B200PHYS_E_INVALID;
    clothSetFixedKernel<<<1, 256, 0, c.stream>>>(S.dev, id, make_float4(rotation[0], rotation[1], rotation[2], rotation[3]), make_float4(position[0], position[1], position[2], 0.f),
                                                 make_float4(scale[0], scale[1], scale[2], 0.f), moveRigid);
    CKC(cudaGetLastError());
    return 0;
}

// totalMass / stiffness / damping / gravityFactor are plain members in the reference; a change of mass or stiffness makes the next
// simulate() call recalculateProperties (cloth.cpp:195-201, 330-346) -- done here, on the host copy and the device arrays
int b200phys_cloth_set_properties(b200phys_context* ctx, uint32_t id, float totalMass, float stiffness, float damping, float gravityFactor)
{
    if (!ctx || !ctx->c.cloths || !(totalMass > 0.f)) return B200PHYS_E_INVALID;
    Context& c = ctx->c; ClothSet& S = *(ClothSet*)c.cloths; CKC(cudaSetDevice(c.device));
    if (id >= S.host.size()) return B200PHYS_E_INVALID;
    ClothDev& k = S.host[id];
    k.totalMass = totalMass; k.stiffness = stiffness; k.damping = damping; k.gravityFactor = gravityFactor;
    CKC(cudaStreamSynchronize(c.stream));
    if (k.totalMass != k.oldTotalMass || k.stiffness != k.oldStiffness)
    {
        std::vector<float> invMass(k.n); std::vector<ClothConstraint> cons(k.numConstraints);
        CKC(cudaMemcpy(invMass.data(), k.invMasses, sizeof(float) * k.n, cudaMemcpyDeviceToHost));
        CKC(cudaMemcpy(cons.data(), k.constraints, sizeof(ClothConstraint) * k.numConstraints, cudaMemcpyDeviceToHost));
        const float invMassPerParticle = k.n / k.totalMass;
        for (float& im : invMass) im = (im != 0.f) ? invMassPerParticle : 0.f;
        k.stiffness = clampf(k.stiffness, 0.01f, 1.f);
        const float invStiffness = 1.f / k.stiffness;
        for (ClothConstraint& q : cons) q.inverseMassSum = (invMass[q.a] + invMass[q.b]) * invStiffness;
        CKC(cudaMemcpy(k.invMasses, invMass.data(), sizeof(float) * k.n, cudaMemcpyHostToDevice));
        CKC(cudaMemcpy(k.constraints, cons.data(), sizeof(ClothConstraint) * k.numConstraints, cudaMemcpyHostToDevice));
        k.oldTotalMass = k.totalMass; k.oldStiffness = k.stiffness;
    }
    CKC(syncTable(S));
    return 0;
}

// id == 0xFFFFFFFF: every cloth (what physicsStepInternal does with the global force field, physics.cpp:1356)
int b200phys_cloth_apply_wind(b200phys_context* ctx, uint32_t id, const float force3[3])
{
    if (!ctx || !force3) return B200PHYS_E_INVALID;
    Context& c = ctx->c; ClothSet* S = (ClothSet*)c.cloths; CKC(cudaSetDevice(c.device));
    if (!S || S->host.empty()) return 0;
    if (id != 0xFFFFFFFFu && id >= S->host.size()) return B200PHYS_E_INVALID;
    const uint32_t first = id == 0xFFFFFFFFu ? 0u : id, count = id == 0xFFFFFFFFu ? (uint32_t)S->host.size() : 1u;
    clothWindKernel<<<count, 256, 0, c.stream>>>(S->dev, first, force3[0], force3[1], force3[2]);
    CKC(cudaGetLastError());
    return 0;
}

// cloth_component::simulate for every cloth (physics.cpp:1357)
int b200phys_cloth_simulate(b200phys_context* ctx, uint32_t velocityIterations, uint32_t positionIterations, uint32_t driftIterations, float dt)
{
    if (!ctx) return B200PHYS_E_INVALID;
    Context& c = ctx->c; ClothSet* S = (ClothSet*)c.cloths; CKC(cudaSetDevice(c.device));
    if (!S || S->host.empty()) return 0;
    clothSimulateKernel<<<(uint32_t)S->host.size(), 256, 0, c.stream>>>(S->dev, velocityIterations, positionIterations, driftIterations, dt);
    CKC(cudaGetLastError());
    return 0;
}

int b200phys_cloth_download(b200phys_context* ctx, uint32_t id, float* positions3, float* velocities3)
{
    if (!ctx || !ctx->c.cloths) return B200PHYS_E_INVALID;
    Context& c = ctx->c; ClothSet& S = *(ClothSet*)c.cloths; CKC(cudaSetDevice(c.device));
    if (id >= S.host.size()) return B200PHYS_E_INVALID;
    const ClothDev& k = S.host[id];
    if (positions3) CKC(cudaMemcpyAsync(positions3, k.positions, sizeof(f3) * k.n, cudaMemcpyDeviceToHost, c.stream));
    if (velocities3) CKC(cudaMemcpyAsync(velocities3, k.velocities, sizeof(f3) * k.n, cudaMemcpyDeviceToHost, c.stream));
    CKC(cudaStreamSynchronize(c.stream));
    return 0;
}

} // extern "C"
