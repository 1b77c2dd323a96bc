// k_heightmap.cuh -- rigid bodies against a heightmap, included at the end of k_narrow.cu (uses its collider records).
//   reference: src/terrain/heightmap_collider.h:36-205, heightmap_collider.cpp:21-157 (min/max mip pyramid per 129 x 129 chunk, triangle
//              iteration inside a volume, bilinear height lookup)
//              src/physics/heightmap_collision.cpp:6-618 (sphere / capsule / AABB / OBB vs triangle, "lowest point" contact)
//              src/physics/physics.cpp:1238-1248 (the contacts go behind the narrow phase's, one collision {collider, 0xFFFF} per collider)
// One thread per collider walks the chunk pyramids with the reference's explicit stack (same push order => same triangle order) and
// runs the per-triangle tests; a counting pass, a scan over the colliders and a writing pass keep the reference's contact order
// (collider index, then triangle order, then the lowest-point contact). Up to 255 contacts per collider form ONE solver unit whose
// second body is the static dummy. Cylinders and hulls have no case in the reference's switch (:539-561, `lowestPoint` is then read
// uninitialised): they are skipped here and in the oracle.
#pragma once

namespace b200 {

struct HeightmapParams
{
    uint32_t chunksPerDim; float chunkSize, invChunkSize, chunkScale, heightScale, invAmplitudeScale;
    float minX, minY, minZ, restitution, friction;
    const uint16_t* heights;        // chunksPerDim^2 x 129 x 129
    const ushort2* mips;            // per chunk 21845 (min, max) nodes: level 0 (128^2) first
};
static const uint32_t kHmVerts = 129u, kHmNodesPerChunk = 21845u;
B200_D uint32_t hmMipOffset(uint32_t level) { uint32_t o = 0; for (uint32_t l = 0; l < level; ++l) o += (128u >> l) * (128u >> l); return o; }

// bounding_volumes.cpp:1317-1367
B200_D f3 closestPointTriangle(f3 p, f3 a, f3 b, f3 c)
{
    f3 ab = b - a, ac = c - a, ap = p - a;
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) return a;
    f3 bp = p - b;
    float d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0.f && d4 <= d3) return b;
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { float v = d1 / (d1 - d3); return a + v * ab; }
    f3 cp = p - c;
    float d5 = dot(ab, cp), d6 = dot(ac, cp);
    if (d6 >= 0.f && d5 <= d6) return c;
    float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { float w = d2 / (d2 - d6); return a + w * ac; }
    float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) { float w = (d4 - d3) / ((d4 - d3) + (d5 - d6)); return b + w * (c - b); }
    float denom = 1.f / (va + vb + vc);
    float v = vb * denom, w = vc * denom;
    return a + ab * v + ac * w;
}

struct HmContact { f3 point, normal; float depth; };

// heightmap_collision.cpp:43-78
B200_D bool hmSphereVsTriangle(f3 center, float radius, f3 a, f3 b, f3 c, HmContact& out)
{
    f3 closest = closestPointTriangle(center, a, b, c);
    f3 n = closest - center;
    float sq = sqLen(n);
    if (!(sq <= radius * radius)) return false;
    float distance;
    if (sq == 0.f) { f3 triNormal = cross(b - a, c - a); n = -triNormal; distance = 0.f; }
    else { distance = sqrtf(sq); n *= 1.f / distance; }
    out.point = closest; out.normal = n; out.depth = radius - distance;
    return true;
}

// heightmap_collision.cpp:6-41
B200_D void hmAabbIncidentEdge(f3 r, f3 normal, f3& outA, f3& outB)
{
    f3 p = vabs(normal);
    outA = mk3(r.x, r.y, r.z);
    if (p.x > p.y) { if (p.y > p.z) outB = mk3(r.x, r.y, -r.z); else outB = mk3(r.x, -r.y, r.z); }
    else { if (p.x > p.z) outB = mk3(r.x, r.y, -r.z); else outB = mk3(-r.x, r.y, r.z); }
    const float sx = normal.x < 0.f ? -1.f : 1.f, sy = normal.y < 0.f ? -1.f : 1.f, sz = normal.z < 0.f ? -1.f : 1.f;
    outA = mk3(outA.x * sx, outA.y * sy, outA.z * sz);
    outB = mk3(outB.x * sx, outB.y * sy, outB.z * sz);
}

// heightmap_collision.cpp:80-420: separating axes (9 edge crosses, 6 box faces, the triangle plane), then a contact point by category
B200_D bool hmAabbVsTriangle(f3 center, f3 radius, f3 a, f3 b, f3 c, HmContact& out)
{
    a = a - center; b = b - center; c = c - center;
    const f3 f0 = b - a, f1 = c - b, f2 = a - c;
    float minPen = FLT_MAX; f3 minNormal = mk3(0.f);
    int category = 0;        // 0,1,2 = edge axis of triangle edge 0,1,2; 3 = box face; 4 = triangle plane
#pragma unroll
    for (int axis = 0; axis < 3; ++axis)
#pragma unroll
        for (int e = 0; e < 3; ++e)
        {
            const f3 fe = e == 0 ? f0 : (e == 1 ? f1 : f2);
            const f3 q = (e == 0) ? c : b;          // edge 0 projects a and c, edges 1 and 2 project a and b (:108-162)
            float p0, p1, r; f3 normal;
            if (axis == 0) { p0 = (a.z * fe.y) - (a.y * fe.z); p1 = (q.z * fe.y) - (q.y * fe.z); r = radius.y * fabsf(fe.z) + radius.z * fabsf(fe.y); normal = mk3(0.f, -fe.z, fe.y); }
            else if (axis == 1) { p0 = (a.x * fe.z) - (a.z * fe.x); p1 = (q.x * fe.z) - (q.z * fe.x); r = radius.x * fabsf(fe.z) + radius.z * fabsf(fe.x); normal = mk3(fe.z, 0.f, -fe.x); }
            else { p0 = (a.y * fe.x) - (a.x * fe.y); p1 = (q.y * fe.x) - (q.x * fe.y); r = radius.x * fabsf(fe.y) + radius.y * fabsf(fe.x); normal = mk3(-fe.y, fe.x, 0.f); }
            float pen = r - fmax2(-fmax2(p0, p1), fmin2(p0, p1));
            if (pen < 0.f) return false;
            const float l = len(normal);
            pen *= 1.f / l;
            if (pen < minPen) { minPen = pen; minNormal = normal * (1.f / l); category = e; }
        }
    {
        float pen = fmax2(a.x, fmax2(b.x, c.x)) + radius.x; if (pen < 0.f) return false;
        if (pen < minPen) { minPen = pen; minNormal = mk3(-1.f, 0.f, 0.f); category = 3; }
        pen = radius.x - fmin2(a.x, fmin2(b.x, c.x)); if (pen < 0.f) return false;
        if (pen < minPen) { minPen = pen; minNormal = mk3(1.f, 0.f, 0.f); category = 3; }
        pen = fmax2(a.y, fmax2(b.y, c.y)) + radius.y; if (pen < 0.f) return false;
        if (pen < minPen) { minPen = pen; minNormal = mk3(0.f, -1.f, 0.f); category = 3; }
        pen = radius.y - fmin2(a.y, fmin2(b.y, c.y)); if (pen < 0.f) return false;
        if (pen < minPen) { minPen = pen; minNormal = mk3(0.f, 1.f, 0.f); category = 3; }
        pen = fmax2(a.z, fmax2(b.z, c.z)) + radius.z; if (pen < 0.f) return false;
        if (pen < minPen) { minPen = pen; minNormal = mk3(0.f, 0.f, -1.f); category = 3; }
        pen = radius.z - fmin2(a.z, fmin2(b.z, c.z)); if (pen < 0.f) return false;
        if (pen < minPen) { minPen = pen; minNormal = mk3(0.f, 0.f, 1.f); category = 3; }
    }
    {
        f3 triNormal = normalize(cross(f0, f1));
        float triD = dot(triNormal, a);
        float r = dot(radius, vabs(triNormal));
        float s = fabsf(triD);
        float pen = r - s;
        if (pen < 0.f) return false;
        if (pen < minPen) { minPen = pen; minNormal = triNormal; category = 4; }
    }
    f3 triCenter = (a + b + c) * (1.f / 3.f);
    if (dot(minNormal, triCenter) < 0.f) minNormal = -minNormal;
    f3 point;
    if (category < 3)
    {
        f3 a0, a1; hmAabbIncidentEdge(radius, minNormal, a0, a1);
        f3 triA = (category == 0) ? a : (category == 1) ? b : c;
        f3 triB = (category == 0) ? b : (category == 1) ? c : a;
        f3 pa, pb;
        closestSegmentSegment(a0, a1, triA, triB, pa, pb);
        point = (pa + pb) * 0.5f;
    }
    else if (category == 3)
    {
        float da = dot(minNormal, a), db = dot(minNormal, b), dc = dot(minNormal, c);
        f3 p = (da < db) ? ((da < dc) ? a : c) : ((db < dc) ? b : c);
        point = p + minNormal * (minPen * 0.5f);
    }
    else
    {
        f3 p = mk3((minNormal.x < 0.f) ? -radius.x : radius.x, (minNormal.y < 0.f) ? -radius.y : radius.y, (minNormal.z < 0.f) ? -radius.z : radius.z);
        point = p - minNormal * (minPen * 0.5f);
    }
    point += center;
    out.point = point; out.normal = minNormal; out.depth = minPen;
    return true;
}

// heightmap_collider_component::getHeightAt (heightmap_collider.cpp:21-43, 125-157); -FLT_MAX outside the map
B200_D float hmHeightAt(const HeightmapParams& H, float cx, float cz)
{
    cx -= H.minX; cz -= H.minZ;
    cx *= H.invChunkSize; cz *= H.invChunkSize;
    if (cx < 0.f || cz < 0.f || cx >= (float)H.chunksPerDim || cz >= (float)H.chunksPerDim) return -FLT_MAX;
    const uint32_t chunkX = (uint32_t)cx, chunkZ = (uint32_t)cz;
    const uint16_t* hh = H.heights + (size_t)(chunkZ * H.chunksPerDim + chunkX) * kHmVerts * kHmVerts;
    cx = fmodf(cx, 1.f); cz = fmodf(cz, 1.f);
    cx *= (float)(kHmVerts - 1); cz *= (float)(kHmVerts - 1);
    const uint32_t x = (uint32_t)cx, z = (uint32_t)cz;
    const float relX = cx - (float)x, relZ = cz - (float)z;
    const float a = (float)hh[kHmVerts * z + x] * H.heightScale, b = (float)hh[kHmVerts * (z + 1) + x] * H.heightScale;
    const float c = (float)hh[kHmVerts * z + x + 1] * H.heightScale, d = (float)hh[kHmVerts * (z + 1) + x + 1] * H.heightScale;
    const float ac = a + relX * (c - a), bd = b + relX * (d - b);
    return (ac + relZ * (bd - ac)) + H.minY;
}

// iterateTrianglesInVolume (heightmap_collider.h:152-205 over :36-122): func(a, b, c) per triangle in the reference's order
template <class F>
B200_D void hmIterateTriangles(const HeightmapParams& H, f3 vmn, f3 vmx, F&& func)
{
    const f3 mc = mk3(H.minX, H.minY, H.minZ);
    vmn = vmn - mc; vmx = vmx - mc;
    vmn.x *= H.invChunkSize; vmn.z *= H.invChunkSize; vmx.x *= H.invChunkSize; vmx.z *= H.invChunkSize;
    const int32_t last = (int32_t)H.chunksPerDim - 1;
    const uint32_t minX = (uint32_t)max((int32_t)vmn.x, 0), minZ = (uint32_t)max((int32_t)vmn.z, 0);
    const uint32_t maxX = (uint32_t)min(max((int32_t)vmx.x, 0), last), maxZ = (uint32_t)min(max((int32_t)vmx.z, 0), last);
    vmn.y *= H.invAmplitudeScale; vmx.y *= H.invAmplitudeScale;
    const uint32_t minHeight = (uint32_t)(uint16_t)(clamp01(vmn.y) * 65535.f), maxHeight = (uint32_t)(uint16_t)(clamp01(vmx.y) * 65535.f);
    for (uint32_t z = minZ; z <= maxZ; ++z)
        for (uint32_t x = minX; x <= maxX; ++x)
        {
            const float relMinX = fmax2(vmn.x - (float)x, 0.f), relMinZ = fmax2(vmn.z - (float)z, 0.f);
            const float relMaxX = (vmx.x > (float)(x + 1)) ? 1.f : fmodf(vmx.x, 1.f);
            const float relMaxZ = (vmx.z > (float)(z + 1)) ? 1.f : fmodf(vmx.z, 1.f);
            const uint32_t volMinX = (uint32_t)(relMinX * (float)kHmVerts), volMinZ = (uint32_t)(relMinZ * (float)kHmVerts);
            const uint32_t volMaxX = (uint32_t)(relMaxX * (float)kHmVerts), volMaxZ = (uint32_t)(relMaxZ * (float)kHmVerts);
            const f3 chunkMinCorner = mk3((float)x * H.chunkSize, 0.f, (float)z * H.chunkSize) + mc;
            const uint32_t chunk = z * H.chunksPerDim + x;
            const uint16_t* hh = H.heights + (size_t)chunk * kHmVerts * kHmVerts;
            const ushort2* mips = H.mips + (size_t)chunk * kHmNodesPerChunk;
            uint32_t stack[32]; uint32_t sp = 0;          // (level << 16) | (x << 8) | z ; at most 3 * 7 + 4 entries
            stack[sp++] = 7u << 16;
            while (sp > 0)
            {
                const uint32_t e = stack[--sp];
                const uint32_t level = e >> 16, ex = (e >> 8) & 0xFFu, ez = e & 0xFFu;
                const uint32_t eMinX = ex << level, eMinZ = ez << level, eMaxX = ((ex + 1u) << level) - 1u, eMaxZ = ((ez + 1u) << level) - 1u;
                if (eMaxX < volMinX || eMinX > volMaxX) continue;
                if (eMaxZ < volMinZ || eMinZ > volMaxZ) continue;
                const uint32_t seg = 128u >> level;
                const ushort2 mm = mips[hmMipOffset(level) + ez * seg + ex];
                if ((uint32_t)mm.y < minHeight || (uint32_t)mm.x > maxHeight) continue;
                if (level == 0)
                {
                    const float x0 = (float)ex * H.chunkScale, x1 = (float)(ex + 1u) * H.chunkScale, z0 = (float)ez * H.chunkScale, z1 = (float)(ez + 1u) * H.chunkScale;
                    const float hA = (float)hh[kHmVerts * ez + ex] * H.heightScale, hB = (float)hh[kHmVerts * (ez + 1u) + ex] * H.heightScale;
                    const float hC = (float)hh[kHmVerts * ez + ex + 1u] * H.heightScale, hD = (float)hh[kHmVerts * (ez + 1u) + ex + 1u] * H.heightScale;
                    const f3 posA = mk3(x0, hA, z0) + chunkMinCorner, posB = mk3(x0, hB, z1) + chunkMinCorner;
                    const f3 posC = mk3(x1, hC, z0) + chunkMinCorner, posD = mk3(x1, hD, z1) + chunkMinCorner;
                    func(posA, posB, posC);
                    func(posC, posB, posD);
                }
                else
                {
                    const uint32_t l1 = (level - 1u) << 16;
                    stack[sp++] = l1 | ((2u * ex) << 8) | (2u * ez);
                    stack[sp++] = l1 | ((2u * ex) << 8) | (2u * ez + 1u);
                    stack[sp++] = l1 | ((2u * ex + 1u) << 8) | (2u * ez);
                    stack[sp++] = l1 | ((2u * ex + 1u) << 8) | (2u * ez + 1u);
                }
            }
        }
}

// heightmapCollision for ONE collider (heightmap_collision.cpp:519-615). emit(contact) is called in the reference's order; returns the count
template <class E>
B200_D uint32_t hmCollideOne(const HeightmapParams& H, const ColRec& col, uint32_t type, f3 aabbMin, f3 aabbMax, E&& emit)
{
    uint32_t n = 0;
    aabbMax.y += 10.f;
    f3 lowest = mk3(0.f);
    const f3 down = mk3(0.f, -1.f, 0.f);
    HmContact k;
    if (type == B200PHYS_SPHERE)
    {
        const SphereS s = asSphere(col);
        hmIterateTriangles(H, aabbMin, aabbMax, [&](f3 a, f3 b, f3 c) { if (n < 255u && hmSphereVsTriangle(s.c, s.r, a, b, c, k)) { emit(n, k); ++n; } });
        lowest = normalize(down) * s.r + s.c;
    }
    else if (type == B200PHYS_CAPSULE)
    {
        const CapS cap = asCap(col);
        const f3 origin = cap.a, dir = normalize(cap.b - cap.a);
        hmIterateTriangles(H, aabbMin, aabbMax, [&](f3 a, f3 b, f3 c)
        {   // :444-459
            f3 triNormal = normalize(cross(b - a, c - a));
            float d = -dot(triNormal, a);
            float ndotd = dot(dir, triNormal);
            float t = -(dot(origin, triNormal) + d) / ndotd;
            f3 trace = origin + t * dir;
            f3 closest = closestPointTriangle(trace, a, b, c);
            f3 reference = closestPointSegment(closest, cap.a, cap.b);
            if (n < 255u && hmSphereVsTriangle(reference, cap.r, a, b, c, k)) { emit(n, k); ++n; }
        });
        const float dA = dot(down, cap.a), dB = dot(down, cap.b);
        lowest = normalize(down) * cap.r + (dA > dB ? cap.a : cap.b);
    }
    else if (type == B200PHYS_AABB)
    {
        const BoxS box = asBox(col);
        const f3 center = boxCenter(box), radius = boxRadius(box);
        hmIterateTriangles(H, aabbMin, aabbMax, [&](f3 a, f3 b, f3 c) { if (n < 255u && hmAabbVsTriangle(center, radius, a, b, c, k)) { emit(n, k); ++n; } });
        lowest = mk3((down.x < 0.f) ? box.mn.x : box.mx.x, (down.y < 0.f) ? box.mn.y : box.mx.y, (down.z < 0.f) ? box.mn.z : box.mx.z);
    }
    else if (type == B200PHYS_OBB)
    {
        const ObbS o = asObb(col);
        hmIterateTriangles(H, aabbMin, aabbMax, [&](f3 a, f3 b, f3 c)
        {
            a = conjugate(o.q) * (a - o.c); b = conjugate(o.q) * (b - o.c); c = conjugate(o.q) * (c - o.c);
            if (n < 255u && hmAabbVsTriangle(mk3(0.f), o.r, a, b, c, k))
            {   // the reference rotates the finished list back (:500-504); per contact it is the same arithmetic
                k.normal = o.q * k.normal; k.point = o.q * k.point + o.c;
                emit(n, k); ++n;
            }
        });
        const f3 d = conjugate(o.q) * down;
        const f3 r = mk3(d.x < 0.f ? -o.r.x : o.r.x, d.y < 0.f ? -o.r.y : o.r.y, d.z < 0.f ? -o.r.z : o.r.z);
        lowest = o.c + o.q * r;
    }
    else return 0u;
    const float h = hmHeightAt(H, lowest.x, lowest.z);
    if (lowest.y < h && n < 255u) { k.normal = mk3(0.f, -1.f, 0.f); k.point = lowest; k.depth = h - lowest.y; emit(n, k); ++n; }
    return n;
}

// pass 1: contacts per collider -> (contacts << 32 | has any) for the scan; pass 2: write behind the narrow phase's lists
template <bool FILL>
__global__ void __launch_bounds__(128) hmCollideKernel(HeightmapParams H, uint32_t numColliders, uint32_t numBodies, const float4* __restrict__ wCol,
    const float4* __restrict__ aabbMin, const float4* __restrict__ aabbMax, unsigned long long* __restrict__ counts, const unsigned long long* __restrict__ scanOut,
    uint32_t maxManifolds, uint32_t maxContacts, uint2* __restrict__ collisions, uint4* __restrict__ manifolds, float4* __restrict__ contacts, uint2* __restrict__ contactBodies,
    DeviceCounts* cnt)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < numColliders; i += gridDim.x * blockDim.x)
    {
        const ColRec col = loadCol(wCol, i);
        const uint32_t tw = __float_as_uint(col.r3.x), type = tw & 0xFFu, objType = tw >> 8;
        if (objType != B200PHYS_OBJ_RIGID_BODY) { if (!FILL) counts[i] = 0ull; continue; }       // :523-526
        const float4 mn4 = aabbMin[i], mx4 = aabbMax[i];
        const f3 mn = mk3(mn4.x, mn4.y, mn4.z), mx = mk3(mx4.x, mx4.y, mx4.z);
        if (!FILL)
        {
            const uint32_t n = hmCollideOne(H, col, type, mn, mx, [](uint32_t, const HmContact&) {});
            counts[i] = ((unsigned long long)n << 32) | (n ? 1ull : 0ull);
            continue;
        }
        const unsigned long long self = counts[i];
        const uint32_t n = (uint32_t)(self >> 32);
        if (!n) continue;
        const unsigned long long ex = scanOut[i];
        const uint32_t mi = cnt->hmBaseManifolds + (uint32_t)(ex & 0xFFFFFFFFull), cs = cnt->hmBaseContacts + (uint32_t)(ex >> 32);
        if (mi >= maxManifolds || cs + n > maxContacts) { cnt->overflowManifolds = 1; atomicMin(&cnt->hmFitManifolds, mi); atomicMin(&cnt->hmFitContacts, cs); continue; }
        // material combine like writeScalarContact (:577-580)
        const float friction = clamp01(sqrtf(col.r2.w * H.friction));
        const float restitution = clamp01(fmax2(col.r2.z, H.restitution));
        const uint32_t fr = ((uint32_t)(friction * 0xFFFF) << 16) | (uint32_t)(restitution * 0xFFFF);
        const uint32_t rb = __float_as_uint(col.r3.y);
        hmCollideOne(H, col, type, mn, mx, [&](uint32_t k, const HmContact& c)
        {
            contacts[2 * (cs + k)] = make_float4(c.point.x, c.point.y, c.point.z, c.depth);
            contacts[2 * (cs + k) + 1] = make_float4(c.normal.x, c.normal.y, c.normal.z, __uint_as_float(fr));
            contactBodies[cs + k] = make_uint2(rb, numBodies);
        });
        collisions[mi] = make_uint2(i, 0xFFFFFFFFu);           // collider_pair{ i, UINT16_MAX } (:590)
        manifolds[mi] = make_uint4(rb, numBodies, cs, n);
    }
}

__global__ void hmBaseKernel(DeviceCounts* cnt) { cnt->hmBaseManifolds = cnt->numManifolds; cnt->hmBaseContacts = cnt->numContacts; cnt->hmFitManifolds = 0xFFFFFFFFu; cnt->hmFitContacts = 0xFFFFFFFFu; }
__global__ void hmTotalsKernel(const unsigned long long* __restrict__ scanOut, uint32_t numColliders, uint32_t maxManifolds, uint32_t maxContacts, DeviceCounts* cnt)
{
    const unsigned long long t = scanOut[numColliders];
    uint32_t nm = cnt->hmBaseManifolds + (uint32_t)(t & 0xFFFFFFFFull), nk = cnt->hmBaseContacts + (uint32_t)(t >> 32);
    // out of capacity: what fits is a prefix of the collider-ordered list (the first collider that does not fit cuts both counts)
    nm = min(nm, min(cnt->hmFitManifolds, maxManifolds)); nk = min(nk, min(cnt->hmFitContacts, maxContacts));
    cnt->numManifolds = nm; cnt->numContacts = nk;
}

void launchHeightmap(Context* c)
{
    if (!c->hmHeights) return;
    DeviceCounts* cnt = c->counts;
    HeightmapParams H; memset(&H, 0, sizeof(H));
    H.chunksPerDim = c->hm.chunksPerDim; H.chunkSize = c->hm.chunkSize; H.invChunkSize = c->hm.invChunkSize; H.chunkScale = c->hm.chunkScale; H.heightScale = c->hm.heightScale;
    H.invAmplitudeScale = c->hm.invAmplitudeScale; H.minX = c->hm.minCorner[0]; H.minY = c->hm.minCorner[1]; H.minZ = c->hm.minCorner[2]; H.restitution = c->hm.restitution; H.friction = c->hm.friction;
    H.heights = c->hmHeights; H.mips = (const ushort2*)c->hmMips;
    const int G = gridFor(c, 8);
    unsigned long long* counts = c->manScan; unsigned long long* scanOut = c->manScanOut;      // free again after npCompactKernel
    launchK(c, hmBaseKernel, 1, 1, 0, cnt);
    launchK(c, hmCollideKernel<false>, G, 128, 0, H, c->numColliders, c->numBodies, c->wCol, c->aabbMin, c->aabbMax, counts, scanOut, c->maxManifolds, c->maxContacts,
            c->collisions, c->manifolds, c->contacts, c->contactBodies, cnt);
    scanExclusiveU64(c, counts, scanOut, nullptr, c->numColliders, nullptr);
    launchK(c, hmCollideKernel<true>, G, 128, 0, H, c->numColliders, c->numBodies, c->wCol, c->aabbMin, c->aabbMax, counts, scanOut, c->maxManifolds, c->maxContacts,
            c->collisions, c->manifolds, c->contacts, c->contactBodies, cnt);
    launchK(c, hmTotalsKernel, 1, 1, 0, scanOut, c->numColliders, c->maxManifolds, c->maxContacts, cnt);
    c->launches += 4;
    timelineMark(c, "heightmap");
}

} // namespace b200
