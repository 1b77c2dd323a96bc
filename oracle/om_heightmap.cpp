// oracle/om_heightmap.cpp -- TEST INFRASTRUCTURE ONLY (parity oracle). Never linked into the product path.
//
// CPU restatement of the reference's heightmap collision:
//   src/terrain/heightmap_collider.h:19-205, heightmap_collider.cpp:5-157   (min/max mips, triangle iteration, height lookup)
//   src/physics/heightmap_collision.cpp:6-618                               (sphere / capsule / AABB / OBB vs triangle, lowest point)
//   src/physics/physics.cpp:1238-1248                                       (contacts appended behind the narrow phase's)
// Reference behaviour kept as it is: cylinder and hull colliders have no case in heightmapCollision's switch (:539-561), their
// `lowestPoint` is read uninitialised; oracle and device skip such colliders (no contact). Contacts are not de-duplicated (:432).
#include "om_step.h"
#include <cmath>

namespace orc {

static const uint32_t kVerts = 129u;            // TERRAIN_LOD_0_VERTICES_PER_DIMENSION (heightmap_collider.h:8-11)

void Heightmap::set(uint32_t chunksPerDim_, float chunkSize_, V3 minCorner_, float amplitudeScale, float restitution_, float friction_, const uint16_t* h)
{
    present = true;
    chunksPerDim = chunksPerDim_; chunkSize = chunkSize_; invChunkSize = 1.f / chunkSize_;          // heightmap_collider.cpp:5-12
    const uint32_t numSegmentsPerDim = kVerts - 1;
    chunkScale = chunkSize_ / numSegmentsPerDim;
    minCorner = minCorner_; invAmplitudeScale = 1.f / amplitudeScale; heightScale = amplitudeScale / UINT16_MAX;   // update(), :14-19
    restitution = restitution_; friction = friction_;
    const uint32_t nChunks = chunksPerDim * chunksPerDim;
    heights.assign(h, h + (size_t)nChunks * kVerts * kVerts);
    mips.assign(nChunks, std::vector<std::vector<MinMax>>());
    for (uint32_t c = 0; c < nChunks; ++c)
    {   // heightmap_collider_chunk::setHeights (:45-123)
        const uint16_t* hh = heights.data() + (size_t)c * kVerts * kVerts;
        uint32_t numSegments = kVerts - 1;
        const uint32_t numMips = (uint32_t)(log2((double)numSegments) + 1);
        mips[c].resize(numMips);
        {
            const uint32_t readStride = kVerts;
            std::vector<MinMax>& mip = mips[c][0];
            mip.resize(numSegments * numSegments);
            for (uint32_t z = 0; z < numSegments; ++z)
                for (uint32_t x = 0; x < numSegments; ++x)
                {
                    const uint16_t a = hh[readStride * z + x], b = hh[readStride * (z + 1) + x], cc = hh[readStride * z + x + 1], d = hh[readStride * (z + 1) + x + 1];
                    MinMax v = { std::min(a, std::min(b, std::min(cc, d))), std::max(a, std::max(b, std::max(cc, d))) };
                    mip[numSegments * z + x] = v;
                }
        }
        for (uint32_t i = 1; i < numMips; ++i)
        {
            const uint32_t readStride = numSegments;
            numSegments >>= 1;
            const std::vector<MinMax>& rd = mips[c][i - 1];
            std::vector<MinMax>& wr = mips[c][i];
            wr.resize(numSegments * numSegments);
            for (uint32_t z = 0; z < numSegments; ++z)
                for (uint32_t x = 0; x < numSegments; ++x)
                {
                    const MinMax a = rd[readStride * (z * 2) + x * 2], b = rd[readStride * (z * 2 + 1) + x * 2], cc = rd[readStride * (z * 2) + x * 2 + 1], d = rd[readStride * (z * 2 + 1) + x * 2 + 1];
                    MinMax v = { std::min(a.mn, std::min(b.mn, std::min(cc.mn, d.mn))), std::max(a.mx, std::max(b.mx, std::max(cc.mx, d.mx))) };
                    wr[numSegments * z + x] = v;
                }
        }
    }
}

// heightmap_collider_chunk::getHeightAt (heightmap_collider.cpp:125-157) through heightmap_collider_component::getHeightAt (:21-43)
float Heightmap::heightAt(float cx, float cz) const
{
    cx -= minCorner.x; cz -= minCorner.z;
    cx *= invChunkSize; cz *= invChunkSize;
    if (cx < 0.f || cz < 0.f || cx >= chunksPerDim || cz >= chunksPerDim) return -FLT_MAX;
    const uint32_t chunkX = (uint32_t)cx, chunkZ = (uint32_t)cz;
    const uint16_t* hh = heights.data() + (size_t)(chunkZ * chunksPerDim + chunkX) * kVerts * kVerts;
    cx = fmodf(cx, 1.f); cz = fmodf(cz, 1.f);                                   // frac (math.h:40)
    cx *= (kVerts - 1); cz *= (kVerts - 1);
    const uint32_t x = (uint32_t)cx, z = (uint32_t)cz;
    const float relX = cx - x, relZ = cz - z;
    const uint32_t stride = kVerts;
    const float a = hh[stride * z + x] * heightScale, b = hh[stride * (z + 1) + x] * heightScale;
    const float c = hh[stride * z + x + 1] * heightScale, d = hh[stride * (z + 1) + x + 1] * heightScale;
    return lerpf(lerpf(a, c, relX), lerpf(b, d, relX), relZ) + minCorner.y;
}

// heightmap_collider_component::iterateTrianglesInVolume (heightmap_collider.h:152-205) over
// heightmap_collider_chunk::iterateTrianglesInVolume (:36-122): triangles in the reference's order
template <class F>
static void iterateTriangles(const Heightmap& H, Box volume, const F& func)
{
    volume.mn = volume.mn - H.minCorner; volume.mx = volume.mx - H.minCorner;
    volume.mn.x *= H.invChunkSize; volume.mn.z *= H.invChunkSize; volume.mx.x *= H.invChunkSize; volume.mx.z *= H.invChunkSize;
    const uint32_t minX = (uint32_t)std::max((int32_t)volume.mn.x, 0), minZ = (uint32_t)std::max((int32_t)volume.mn.z, 0);
    auto clampi = [](int32_t v, int32_t l, int32_t u) { int32_t r = std::max(l, v); r = std::min(u, r); return r; };
    const uint32_t maxX = (uint32_t)clampi((int32_t)volume.mx.x, 0, (int32_t)H.chunksPerDim - 1), maxZ = (uint32_t)clampi((int32_t)volume.mx.z, 0, (int32_t)H.chunksPerDim - 1);
    volume.mn.y *= H.invAmplitudeScale; volume.mx.y *= H.invAmplitudeScale;
    const uint16_t minHeight = (uint16_t)(clamp01(volume.mn.y) * UINT16_MAX), maxHeight = (uint16_t)(clamp01(volume.mx.y) * UINT16_MAX);
    for (uint32_t z = minZ; z <= maxZ; ++z)
        for (uint32_t x = minX; x <= maxX; ++x)
        {
            const float relMinX = fmax2(volume.mn.x - x, 0.f), relMinZ = fmax2(volume.mn.z - z, 0.f);
            const float relMaxX = (volume.mx.x > (x + 1)) ? 1.f : fmodf(volume.mx.x, 1.f);
            const float relMaxZ = (volume.mx.z > (z + 1)) ? 1.f : fmodf(volume.mx.z, 1.f);
            const uint32_t volMinX = (uint32_t)(relMinX * kVerts), volMinZ = (uint32_t)(relMinZ * kVerts);
            const uint32_t volMaxX = (uint32_t)(relMaxX * kVerts), volMaxZ = (uint32_t)(relMaxZ * kVerts);
            const V3 chunkMinCorner = v3(x * H.chunkSize, 0.f, z * H.chunkSize) + H.minCorner;
            // chunk traversal (:36-122): explicit stack, children pushed (0,0),(0,1),(1,0),(1,1) -> popped in reverse
            const uint32_t chunk = z * H.chunksPerDim + x;
            const uint16_t* hh = H.heights.data() + (size_t)chunk * kVerts * kVerts;
            const std::vector<std::vector<Heightmap::MinMax>>& mips = H.mips[chunk];
            struct Entry { uint16_t mipLevel, x, z; };
            Entry stack[1024]; uint32_t stackSize = 0;
            const uint32_t numMips = (uint32_t)mips.size();
            stack[stackSize++] = { (uint16_t)(numMips - 1), 0, 0 };
            while (stackSize > 0)
            {
                const Entry e = stack[--stackSize];
                const uint32_t eMinX = (uint32_t)e.x << e.mipLevel, eMinZ = (uint32_t)e.z << e.mipLevel;
                const uint32_t eMaxX = (((uint32_t)e.x + 1) << e.mipLevel) - 1, eMaxZ = (((uint32_t)e.z + 1) << e.mipLevel) - 1;
                if (eMaxX < volMinX || eMinX > volMaxX) continue;
                if (eMaxZ < volMinZ || eMinZ > volMaxZ) continue;
                const uint32_t numSegmentsPerDim = (kVerts - 1) >> e.mipLevel;
                const Heightmap::MinMax mm = mips[e.mipLevel][e.z * numSegmentsPerDim + e.x];
                if (mm.mx < minHeight || mm.mn > maxHeight) continue;
                if (e.mipLevel == 0)
                {
                    const uint32_t stride = kVerts;
                    const float ax = (float)e.x * H.chunkScale, az = (float)e.z * H.chunkScale;
                    const float bx = (float)e.x * H.chunkScale, bz = (float)(e.z + 1) * H.chunkScale;
                    const float cx = (float)(e.x + 1) * H.chunkScale, cz = (float)e.z * H.chunkScale;
                    const float dx = (float)(e.x + 1) * H.chunkScale, dz = (float)(e.z + 1) * H.chunkScale;
                    const float hA = hh[stride * e.z + e.x] * H.heightScale, hB = hh[stride * (e.z + 1) + e.x] * H.heightScale;
                    const float hC = hh[stride * e.z + e.x + 1] * H.heightScale, hD = hh[stride * (e.z + 1) + e.x + 1] * H.heightScale;
                    const V3 posA = v3(ax, hA, az) + chunkMinCorner, posB = v3(bx, hB, bz) + chunkMinCorner;
                    const V3 posC = v3(cx, hC, cz) + chunkMinCorner, posD = v3(dx, hD, dz) + chunkMinCorner;
                    func(posA, posB, posC);
                    func(posC, posB, posD);
                }
                else
                {
                    stack[stackSize++] = { (uint16_t)(e.mipLevel - 1), (uint16_t)(2 * e.x + 0), (uint16_t)(2 * e.z + 0) };
                    stack[stackSize++] = { (uint16_t)(e.mipLevel - 1), (uint16_t)(2 * e.x + 0), (uint16_t)(2 * e.z + 1) };
                    stack[stackSize++] = { (uint16_t)(e.mipLevel - 1), (uint16_t)(2 * e.x + 1), (uint16_t)(2 * e.z + 0) };
                    stack[stackSize++] = { (uint16_t)(e.mipLevel - 1), (uint16_t)(2 * e.x + 1), (uint16_t)(2 * e.z + 1) };
                }
            }
        }
}

// bounding_volumes.cpp:1317-1367
static V3 closestPointTriangle(V3 p, V3 a, V3 b, V3 c)
{
    V3 ab = b - a, ac = c - a, ap = p - a;
    float d1 = dot(ab, ap), d2 = dot(ac, ap);
    if (d1 <= 0.f && d2 <= 0.f) return a;
    V3 bp = p - b;
    float d3 = dot(ab, bp), d4 = dot(ac, bp);
    if (d3 >= 0.f && d4 <= d3) return b;
    float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { float v = d1 / (d1 - d3); return a + v * ab; }
    V3 cp = p - c;
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

// heightmap_collision.cpp:6-41
static void aabbIncidentEdge(V3 r, V3 normal, V3& outA, V3& outB)
{
    V3 p = vabs(normal);
    outA = v3(r.x, r.y, r.z);
    if (p.x > p.y) { if (p.y > p.z) outB = v3(r.x, r.y, -r.z); else outB = v3(r.x, -r.y, r.z); }
    else { if (p.x > p.z) outB = v3(r.x, r.y, -r.z); else outB = v3(-r.x, r.y, r.z); }
    const float sx = normal.x < 0.f ? -1.f : 1.f, sy = normal.y < 0.f ? -1.f : 1.f, sz = normal.z < 0.f ? -1.f : 1.f;
    outA = v3(outA.x * sx, outA.y * sy, outA.z * sz);
    outB = v3(outB.x * sx, outB.y * sy, outB.z * sz);
}

// heightmap_collision.cpp:43-78
static uint32_t sphereVsTriangle(V3 center, float radius, V3 a, V3 b, V3 c, Contact* out)
{
    V3 closest = closestPointTriangle(center, a, b, c);
    V3 n = closest - center;
    float sq = sqLen(n);
    if (sq <= radius * radius)
    {
        float distance;
        if (sq == 0.f) { V3 triNormal = cross(b - a, c - a); n = -triNormal; distance = 0.f; }
        else { distance = sqrtf(sq); n *= 1.f / distance; }
        out->point = closest; out->normal = n; out->depth = radius - distance;
        return 1;
    }
    return 0;
}

// heightmap_collision.cpp:80-420
static uint32_t aabbVsTriangle(V3 center, V3 radius, V3 a, V3 b, V3 c, Contact* out)
{
    a = a - center; b = b - center; c = c - center;
    const V3 f0 = b - a, f1 = c - b, f2 = a - c;
    float minPen = FLT_MAX; V3 minNormal = v3(0.f);
    int category = 0;        // 0,1,2 = category0_edge0..2, 3 = category1, 4 = category2
    const V3 f[3] = { f0, f1, f2 };
    // nine edge-cross axes, x-axis group first, edges 0,1,2 inside a group (:107-278). Edge 0 projects a and c, edges 1, 2 project a and b.
    for (int axis = 0; axis < 3; ++axis)
        for (int e = 0; e < 3; ++e)
        {
            const V3 fe = f[e];
            const V3 q = (e == 0) ? c : b;
            float p0, p1, r; V3 normal;
            if (axis == 0) { p0 = (a.z * fe.y) - (a.y * fe.z); p1 = (q.z * fe.y) - (q.y * fe.z); r = radius.y * fabsf(fe.z) + radius.z * fabsf(fe.y); normal = v3(0.f, -fe.z, fe.y); }
            else if (axis == 1) { p0 = (a.x * fe.z) - (a.z * fe.x); p1 = (q.x * fe.z) - (q.z * fe.x); r = radius.x * fabsf(fe.z) + radius.z * fabsf(fe.x); normal = v3(fe.z, 0.f, -fe.x); }
            else { p0 = (a.y * fe.x) - (a.x * fe.y); p1 = (q.y * fe.x) - (q.x * fe.y); r = radius.x * fabsf(fe.y) + radius.y * fabsf(fe.x); normal = v3(-fe.y, fe.x, 0.f); }
            float pen = r - fmax2(-fmax2(p0, p1), fmin2(p0, p1));
            if (pen < 0.f) return 0;
            const float l = len(normal);
            pen *= 1.f / l;
            if (pen < minPen) { minPen = pen; minNormal = normal * (1.f / l); category = e; }
        }
    // six box face axes (:281-346)
    {
        float pen = fmax2(a.x, fmax2(b.x, c.x)) + radius.x; if (pen < 0.f) return 0;
        if (pen < minPen) { minPen = pen; minNormal = v3(-1.f, 0.f, 0.f); category = 3; }
        pen = radius.x - fmin2(a.x, fmin2(b.x, c.x)); if (pen < 0.f) return 0;
        if (pen < minPen) { minPen = pen; minNormal = v3(1.f, 0.f, 0.f); category = 3; }
        pen = fmax2(a.y, fmax2(b.y, c.y)) + radius.y; if (pen < 0.f) return 0;
        if (pen < minPen) { minPen = pen; minNormal = v3(0.f, -1.f, 0.f); category = 3; }
        pen = radius.y - fmin2(a.y, fmin2(b.y, c.y)); if (pen < 0.f) return 0;
        if (pen < minPen) { minPen = pen; minNormal = v3(0.f, 1.f, 0.f); category = 3; }
        pen = fmax2(a.z, fmax2(b.z, c.z)) + radius.z; if (pen < 0.f) return 0;
        if (pen < minPen) { minPen = pen; minNormal = v3(0.f, 0.f, -1.f); category = 3; }
        pen = radius.z - fmin2(a.z, fmin2(b.z, c.z)); if (pen < 0.f) return 0;
        if (pen < minPen) { minPen = pen; minNormal = v3(0.f, 0.f, 1.f); category = 3; }
    }
    {   // triangle plane (:348-364)
        V3 triNormal = normalize(cross(f0, f1));
        float triD = dot(triNormal, a);
        float r = dot(radius, vabs(triNormal));
        float s = fabsf(triD);
        float pen = r - s;
        if (pen < 0.f) return 0;
        if (pen < minPen) { minPen = pen; minNormal = triNormal; category = 4; }
    }
    V3 triCenter = (a + b + c) * (1.f / 3.f);
    if (dot(minNormal, triCenter) < 0.f) minNormal = -minNormal;
    V3 point;
    if (category < 3)
    {
        V3 a0, a1; aabbIncidentEdge(radius, minNormal, a0, a1);
        V3 triA = (category == 0) ? a : (category == 1) ? b : c;
        V3 triB = (category == 0) ? b : (category == 1) ? c : a;
        V3 pa, pb;
        closestPointSegmentSegment(a0, a1, triA, triB, pa, pb);
        point = (pa + pb) * 0.5f;
    }
    else if (category == 3)
    {
        float da = dot(minNormal, a), db = dot(minNormal, b), dc = dot(minNormal, c);
        V3 p = (da < db) ? ((da < dc) ? a : c) : ((db < dc) ? b : c);
        point = p + minNormal * (minPen * 0.5f);
    }
    else
    {
        V3 p = v3((minNormal.x < 0.f) ? -radius.x : radius.x, (minNormal.y < 0.f) ? -radius.y : radius.y, (minNormal.z < 0.f) ? -radius.z : radius.z);
        point = p - minNormal * (minPen * 0.5f);
    }
    point += center;
    out->point = point; out->normal = minNormal; out->depth = minPen;
    return 1;
}

// heightmapCollision (heightmap_collision.cpp:509-618) appended behind the narrow phase (physics.cpp:1238-1248)
void Oracle::heightmapCollisions()
{
    if (!heightmap.present) return;
    const Heightmap& H = heightmap;
    const uint32_t dummy = (uint32_t)bodies.size();
    std::vector<Contact> buf(256);
    for (uint32_t i = 0; i < (uint32_t)wcol.size(); ++i)
    {
        const WCol& col = wcol[i];
        if (col.objType != O_RIGID) continue;
        Box aabb = aabbs[i];
        aabb.mx.y += 10.f;
        uint32_t n = 0;
        V3 lowest = v3(0.f);
        const V3 down = v3(0.f, -1.f, 0.f);
        bool handled = true;
        switch (col.type)
        {
        case T_SPHERE:
        {
            const Sphere s = col.sphere;
            iterateTriangles(H, aabb, [&](V3 a, V3 b, V3 c) { if (n < 255) n += sphereVsTriangle(s.c, s.r, a, b, c, &buf[n]); });
            lowest = normalize(down) * s.r + s.c;                                    // sphere_support_fn, collision_gjk.h:6-14
        } break;
        case T_CAPSULE:
        {
            const Capsule cap = col.cap;
            const V3 origin = cap.a, dir = normalize(cap.b - cap.a);
            iterateTriangles(H, aabb, [&](V3 a, V3 b, V3 c)
            {   // :444-459
                V3 triNormal = normalize(cross(b - a, c - a));
                float d = -dot(triNormal, a);
                float ndotd = dot(dir, triNormal);
                float t = -(dot(origin, triNormal) + d) / ndotd;
                V3 trace = origin + t * dir;
                V3 closest = closestPointTriangle(trace, a, b, c);
                V3 reference = closestPointSegment(closest, cap.a, cap.b);
                if (n < 255) n += sphereVsTriangle(reference, cap.r, a, b, c, &buf[n]);
            });
            const float dA = dot(down, cap.a), dB = dot(down, cap.b);                // capsule_support_fn, collision_gjk.h:16-27
            lowest = normalize(down) * cap.r + (dA > dB ? cap.a : cap.b);
        } break;
        case T_AABB:
        {
            const Box box = col.box;
            const V3 center = boxCenter(box), radius = boxRadius(box);
            iterateTriangles(H, aabb, [&](V3 a, V3 b, V3 c) { if (n < 255) n += aabbVsTriangle(center, radius, a, b, c, &buf[n]); });
            lowest = v3((down.x < 0.f) ? box.mn.x : box.mx.x, (down.y < 0.f) ? box.mn.y : box.mx.y, (down.z < 0.f) ? box.mn.z : box.mx.z);   // :46-58
        } break;
        case T_OBB:
        {
            const Obb o = col.obb;
            iterateTriangles(H, aabb, [&](V3 a, V3 b, V3 c)
            {
                a = conjugate(o.q) * (a - o.c); b = conjugate(o.q) * (b - o.c); c = conjugate(o.q) * (c - o.c);
                if (n < 255) n += aabbVsTriangle(v3(0.f), o.r, a, b, c, &buf[n]);
            });
            for (uint32_t j = 0; j < n; ++j) { buf[j].normal = o.q * buf[j].normal; buf[j].point = o.q * buf[j].point + o.c; }
            const V3 d = conjugate(o.q) * down;                                      // obb_support_fn, collision_gjk.h:60-75
            const V3 r = v3(d.x < 0.f ? -o.r.x : o.r.x, d.y < 0.f ? -o.r.y : o.r.y, d.z < 0.f ? -o.r.z : o.r.z);
            lowest = o.c + o.q * r;
        } break;
        default: handled = false; break;
        }
        if (!handled) continue;
        const float h = H.heightAt(lowest.x, lowest.z);
        if (lowest.y < h && n < 255) { Contact& k = buf[n++]; k.normal = v3(0.f, -1.f, 0.f); k.point = lowest; k.depth = h - lowest.y; }
        if (n > 0)
        {
            const float friction = clamp01(sqrtf(col.friction * H.friction));
            const float restitution = clamp01(fmax2(col.restitution, H.restitution));
            const uint32_t fr = ((uint32_t)(friction * 0xFFFF) << 16) | (uint32_t)(restitution * 0xFFFF);
            for (uint32_t j = 0; j < n; ++j) { buf[j].fr = fr; contacts.push_back(buf[j]); contactBodies.push_back({ col.objIndex, dummy }); }
            contactCount.push_back((uint8_t)n);
            collisions.push_back(Pair{ i, 0xFFFFFFFFu });                            // collider_pair{ i, UINT16_MAX } (:590)
        }
    }
}

} // namespace orc
