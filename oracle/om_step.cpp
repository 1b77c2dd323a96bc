// oracle/om_step.cpp -- TEST INFRASTRUCTURE ONLY (parity oracle). Never linked into the product path.
//
// CPU restatement of the reference's scalar physicsStepInternal (src/physics/physics.cpp:1180-1362) and the stages it
// calls: getWorldSpaceColliders (physics.cpp:631-756), broadphase (collision_broad.cpp:87-166,297-447), narrowphase
// driver (collision_narrow.cpp:2221-2253,2328-2603), applyGravityAndIntegrateForces / integrateVelocity
// (rigid_body.cpp:95-142), collision constraints (constraints.cpp:3307-3449) and the per-iteration kind order
// (constraints.cpp:3748-3772). Single thread, 32-bit indices, flat arrays instead of EnTT pools.
#include "om_step.h"
#include <algorithm>
#include <chrono>

namespace orc {

static inline V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
static inline Q4 ld4(const float* p) { return q4(p[0], p[1], p[2], p[3]); }
static inline void st3(float* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
static inline double nowMs() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// ------------------------------------------------------------------------------------------------ world colliders
void Oracle::worldColliders()
{
    const uint32_t C = (uint32_t)colliders.size();
    const uint32_t dummy = (uint32_t)bodies.size();
    wcol.resize(C); aabbs.resize(C);
    for (uint32_t i = 0; i < C; ++i)
    {
        const b200phys_collider& c = colliders[i];
        WCol& w = wcol[i]; Box& bb = aabbs[i];

        // pose source: physics_transform1 of the rigid body, else transform_component, else identity (physics.cpp:645-647)
        Q4 rot = qIdentity(); V3 pos = v3(0.f);
        if (c.objectType == O_RIGID) { const b200phys_body& b = bodies[c.objectIndex]; rot = ld4(b.rotation); pos = ld3(b.position); }
        else if (c.parentTransform != B200PHYS_NO_TRANSFORM) { const b200phys_transform& t = transforms[c.parentTransform]; rot = ld4(t.rotation); pos = ld3(t.position); }

        w.type = c.type; w.objType = c.objectType;
        w.restitution = c.restitution; w.friction = c.friction;
        w.objIndex = (c.objectType == O_STATIC) ? dummy : c.objectIndex;   // physics.cpp:652-671
        w.hull.g = nullptr;

        const float* s = c.shape;
        switch (c.type)
        {
        case T_SPHERE:   // physics.cpp:675-680
        {
            V3 center = pos + rot * ld3(s);
            bb = boxFromCenterRadius(center, v3(s[3]));
            w.sphere.c = center; w.sphere.r = s[3];
        } break;
        case T_CAPSULE:  // physics.cpp:682-697
        {
            V3 pA = rot * ld3(s) + pos, pB = rot * ld3(s + 3) + pos;
            V3 r3 = v3(s[6]);
            bb = boxNegInf();
            grow(bb, pA + r3); grow(bb, pA - r3); grow(bb, pB + r3); grow(bb, pB - r3);
            w.cap.a = pA; w.cap.b = pB; w.cap.r = s[6];
        } break;
        case T_CYLINDER: // physics.cpp:699-720
        {
            V3 pA = rot * ld3(s) + pos, pB = rot * ld3(s + 3) + pos;
            float radius = s[6];
            V3 a = pB - pA;
            float aa = dot(a, a);
            float x = 1.f - a.x * a.x / aa, y = 1.f - a.y * a.y / aa, z = 1.f - a.z * a.z / aa;
            x = sqrtf(fmax2(0.f, x)); y = sqrtf(fmax2(0.f, y)); z = sqrtf(fmax2(0.f, z));
            V3 e = radius * v3(x, y, z);
            bb.mn = vmin(pA - e, pB - e); bb.mx = vmax(pA + e, pB + e);
            w.cap.a = pA; w.cap.b = pB; w.cap.r = radius;
        } break;
        case T_AABB:     // physics.cpp:722-734 : becomes an OBB under any non-identity rotation
        {
            Box local = { ld3(s), ld3(s + 3) };
            bb = boxTransformToAABB(local, rot, pos);
            if (rot == qIdentity()) { w.box = bb; }
            else
            {
                w.type = T_OBB;
                w.obb.c = rot * boxCenter(local) + pos;   // bounding_volumes.cpp:72-79
                w.obb.r = boxRadius(local);
                w.obb.q = rot;
            }
        } break;
        case T_OBB:      // physics.cpp:736-740, bounding_volumes.cpp:130-139
        {
            Q4 oq = ld4(s); V3 oc = ld3(s + 4), orad = ld3(s + 7);
            Box local = { -orad, orad };
            Q4 q = rot * oq; V3 cpos = rot * oc + pos;
            bb = boxTransformToAABB(local, q, cpos);
            w.obb.q = q; w.obb.c = cpos; w.obb.r = orad;
        } break;
        default:         // hull, physics.cpp:742-753
        {
            uint32_t gi; memcpy(&gi, &s[7], 4);
            const HullGeom& geo = hulls[gi];
            Q4 q = rot * ld4(s); V3 p = rot * ld3(s + 4) + pos;
            bb = boxTransformToAABB(geo.aabb, q, p);
            w.hull.q = q; w.hull.p = p; w.hull.g = &geo;
        } break;
        }
    }
}

// ---------------------------------------------------------------------------------------------------- broad phase
void Oracle::broadphaseBrute(std::vector<Pair>& out) const
{   // the reference's own "#if 0" cross-check idea (collision_broad.cpp:316-340) with the real predicate
    out.clear();
    const uint32_t C = (uint32_t)aabbs.size();
    for (uint32_t i = 0; i < C; ++i) for (uint32_t j = i + 1; j < C; ++j) if (aabbVsAABB(aabbs[i], aabbs[j])) out.push_back({ i, j });
}

void Oracle::broadphaseSAP()
{
    const uint32_t C = (uint32_t)aabbs.size();
    pairs.clear();
    if (C == 0) return;
    if (sapColliderCount != C)
    {
        // endpoints are appended start,end per collider in CREATION order (collision_broad.cpp:27-40); collider index i
        // is the position in EnTT's back-to-front iteration, i.e. creation order C-1-i
        endpoints.clear();
        for (uint32_t k = 0; k < C; ++k) { uint32_t ci = C - 1 - k; endpoints.push_back({ 0.f, ci, true }); endpoints.push_back({ 0.f, ci, false }); }
        sortingAxis = 0; sapColliderCount = C;
    }
    const uint32_t nE = 2 * C;
    V3 s = v3(0.f), s2 = v3(0.f);
    // update endpoint values on the current axis (collision_broad.cpp:349-380)
    std::vector<uint32_t> startOf(C), endOf(C);
    for (uint32_t e = 0; e < nE; ++e) { if (endpoints[e].start) startOf[endpoints[e].collider] = e; else endOf[endpoints[e].collider] = e; }
    for (uint32_t i = 0; i < C; ++i)
    {
        const Box& bb = aabbs[i];
        endpoints[startOf[i]].value = comp(bb.mn, (int)sortingAxis);
        endpoints[endOf[i]].value = comp(bb.mx, (int)sortingAxis);
        V3 c = boxCenter(bb);
        s += c; s2 += c * c;
    }
    // insertion sort (collision_broad.cpp:385-398)
    for (uint32_t i = 1; i < nE; ++i)
    {
        SapEndpoint key = endpoints[i];
        uint32_t j = i - 1;
        while (j != UINT32_MAX && endpoints[j].value > key.value) { endpoints[j + 1] = endpoints[j]; j = j - 1; }
        endpoints[j + 1] = key;
    }
    // sweep with swap-remove active list (collision_broad.cpp:87-166)
    std::vector<uint32_t> active(C), posInActive(C);
    std::vector<Box> activeBB(C);
    uint32_t nActive = 0;
    for (uint32_t i = 0; i < nE; ++i)
    {
        SapEndpoint ep = endpoints[i];
        if (ep.start)
        {
            const Box& a = aabbs[ep.collider];
            for (uint32_t k = 0; k < nActive; ++k) if (aabbVsAABB(a, activeBB[k])) pairs.push_back({ ep.collider, active[k] });
            posInActive[ep.collider] = nActive;
            activeBB[nActive] = a;
            active[nActive++] = ep.collider;
        }
        else
        {
            uint32_t pos = posInActive[ep.collider];
            --nActive;
            uint32_t last = active[nActive];
            posInActive[last] = pos;
            active[pos] = active[nActive];
            activeBB[pos] = activeBB[nActive];
        }
    }
    // next frame's axis = largest variance of AABB centres (collision_broad.cpp:443-444)
    V3 var = s2 - s * s / (float)C;
    sortingAxis = (var.x > var.y) ? ((var.x > var.z) ? 0 : 2) : ((var.y > var.z) ? 1 : 2);

    if (pairOrder == ORDER_CANONICAL)
    {
        for (Pair& p : pairs) if (p.a > p.b) std::swap(p.a, p.b);
        std::sort(pairs.begin(), pairs.end(), [](const Pair& x, const Pair& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; });
    }
}

// --------------------------------------------------------------------------------------------------- narrow phase
void Oracle::narrowphase()
{
    collisions.clear(); contactCount.clear(); contacts.clear(); contactBodies.clear();
    numNonCollisionInteractions = 0;
    std::vector<Pair> work = pairs;
    const uint32_t numPairs = (uint32_t)work.size();
    uint32_t numCollisionChecks = 0, numIntersectionChecks = 0;

    // prune / order by type / split (collision_narrow.cpp:2346-2397)
    uint32_t count = numPairs;
    std::vector<Pair> intersectionPairs;
    for (uint32_t i = 0; i < count; ++i)
    {
        Pair p = work[i];
        const WCol* A = &wcol[p.a]; const WCol* B = &wcol[p.b];
        if (A->objType != O_RIGID && B->objType != O_RIGID) continue;
        if (A->objType == O_RIGID && B->objType == O_RIGID && A->objIndex == B->objIndex) continue;
        if (!(A->type < B->type)) { std::swap(p.a, p.b); }     // equal types swap too (strict <, :2374)
        A = &wcol[p.a]; B = &wcol[p.b];
        if ((A->objType == O_RIGID && B->objType == O_RIGID) || A->objType == O_STATIC || B->objType == O_STATIC)
        {
            work[numCollisionChecks++] = p;
        }
        else if (pairOrder == ORDER_REFERENCE_SAP)
        {   // the reference swaps the trigger pair to the back of the array, which permutes the unprocessed tail (:2389-2394)
            uint32_t lastIndex = numPairs - 1 - numIntersectionChecks++;
            work[i] = work[lastIndex];
            work[lastIndex] = p;
            --i; --count;
        }
        else { intersectionPairs.push_back(p); ++numIntersectionChecks; }
    }
    if (pairOrder == ORDER_REFERENCE_SAP)
        for (uint32_t k = 0; k < numIntersectionChecks; ++k) intersectionPairs.push_back(work[numPairs - numIntersectionChecks + k]);

    // stable counting sort into the upper-triangular type buckets (collision_narrow.cpp:2399-2453)
    std::vector<Pair> sorted; sorted.reserve(numCollisionChecks);
    for (int ta = 0; ta < 6; ++ta) for (int tb = ta; tb < 6; ++tb)
        for (uint32_t i = 0; i < numCollisionChecks; ++i)
            if (wcol[work[i].a].type == ta && wcol[work[i].b].type == tb) sorted.push_back(work[i]);

    // collisionScalar + writeScalarContact (collision_narrow.cpp:2221-2253, 2284-2302)
    for (const Pair& p : sorted)
    {
        const WCol& A = wcol[p.a]; const WCol& B = wcol[p.b];
        Manifold m;
        if (!collide(A, B, m, nstats)) continue;
        float friction = clamp01(sqrtf(A.friction * B.friction));
        float restitution = clamp01(fmax2(A.restitution, B.restitution));
        uint32_t fr = ((uint32_t)(friction * 0xFFFF) << 16) | (uint32_t)(restitution * 0xFFFF);
        collisions.push_back(p); contactCount.push_back((uint8_t)m.n);
        for (uint32_t k = 0; k < m.n; ++k)
        {
            Contact c; c.normal = m.normal; c.depth = m.depth[k]; c.point = m.point[k]; c.fr = fr;
            contacts.push_back(c);
            contactBodies.push_back({ A.objIndex, B.objIndex });
        }
    }
    // trigger / force-field overlaps (collision_narrow.cpp:2573-2593, overlapCheck :1593-1612). The reference walks them type
    // bucket by type bucket; here (and on the device) they stay in pair order: only the floating-point summation order of three
    // or more force fields acting on one body could tell the difference.
    interactions.clear();
    for (const Pair& p : intersectionPairs)
    {
        const WCol& A = wcol[p.a]; const WCol& B = wcol[p.b];
        if (!overlapOnly(A, B, nstats)) continue;
        Interaction it;
        if (A.objType == O_RIGID) { it.rigidBodyIndex = A.objIndex; it.otherIndex = B.objIndex; it.otherType = B.objType; }
        else { it.rigidBodyIndex = B.objIndex; it.otherIndex = A.objIndex; it.otherType = A.objType; }
        interactions.push_back(it);
    }
    numNonCollisionInteractions = (uint32_t)interactions.size();
}

// handleNonCollisionInteractions (physics.cpp:952-1035): localized force fields push the overlapping bodies, trigger overlaps
// are sorted, de-duplicated (several colliders of one body may report the same trigger) and merged against last frame's list.
void Oracle::nonCollisionInteractions()
{
    std::vector<Pair> overlaps;
    for (const Interaction& it : interactions)
    {
        if (it.otherType == B200PHYS_OBJ_FORCE_FIELD)
        {
            if (it.otherIndex < forceFields.size() && it.rigidBodyIndex < bodies.size())
            {
                float* F = bodies[it.rigidBodyIndex].forceAccumulator; const V3 f = forceFields[it.otherIndex];
                F[0] += f.x; F[1] += f.y; F[2] += f.z;
            }
        }
        else if (it.otherType == B200PHYS_OBJ_TRIGGER) overlaps.push_back(Pair{ it.otherIndex, it.rigidBodyIndex });
    }
    auto less = [](const Pair& x, const Pair& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; };
    auto same = [](const Pair& x, const Pair& y) { return x.a == y.a && x.b == y.b; };
    std::sort(overlaps.begin(), overlaps.end(), less);
    overlaps.erase(std::unique(overlaps.begin(), overlaps.end(), same), overlaps.end());
    triggerEnter.clear(); triggerLeave.clear();
    size_t p = 0, t = 0;
    while (p < prevFrameTriggerOverlaps.size() && t < overlaps.size())
    {
        if (same(prevFrameTriggerOverlaps[p], overlaps[t])) { ++p; ++t; continue; }
        if (less(prevFrameTriggerOverlaps[p], overlaps[t])) triggerLeave.push_back(prevFrameTriggerOverlaps[p++]);
        else triggerEnter.push_back(overlaps[t++]);
    }
    while (p < prevFrameTriggerOverlaps.size()) triggerLeave.push_back(prevFrameTriggerOverlaps[p++]);
    while (t < overlaps.size()) triggerEnter.push_back(overlaps[t++]);
    prevFrameTriggerOverlaps = std::move(overlaps);
}

// ------------------------------------------------------------------------------------------------- integration
void Oracle::integrateForces(float dt)
{
    const uint32_t N = (uint32_t)bodies.size();
    g.resize(N + 1);
    for (uint32_t i = 0; i < N; ++i)
    {   // rigid_body.cpp:95-124
        b200phys_body& b = bodies[i]; BodyG& G = g[i];
        Q4 rot = ld4(b.rotation); V3 pos = ld3(b.position), cog = ld3(b.localCOGPosition);
        G.rotation = rot;
        G.position = pos + rot * cog;
        M3 R = quaternionToMat3(rot);
        M3 I; memcpy(I.m, b.invInertia, sizeof(I.m));
        G.invInertia = R * I * transpose(R);
        G.invMass = b.invMass;
        V3 F = ld3(b.forceAccumulator) + globalForceField, T = ld3(b.torqueAccumulator);   // physics.cpp:1273
        if (b.invMass > 0.f) { F.y += (-9.81f / b.invMass * b.gravityFactor); }
        V3 linAcc = F * b.invMass;
        V3 angAcc = G.invInertia * T;
        V3 v = ld3(b.linearVelocity), w = ld3(b.angularVelocity);
        v += linAcc * dt; w += angAcc * dt;
        v *= 1.f / (1.f + dt * b.linearDamping);
        w *= 1.f / (1.f + dt * b.angularDamping);
        st3(b.forceAccumulator, F); st3(b.linearVelocity, v); st3(b.angularVelocity, w);
        G.v = v; G.w = w; G.localCOG = cog;
    }
    memset(&g[N], 0, sizeof(BodyG));   // kinematic stand-in for static colliders (physics.cpp:1279)
}

void Oracle::integrateVelocities(float dt)
{
    const uint32_t N = (uint32_t)bodies.size();
    for (uint32_t i = 0; i < N; ++i)
    {   // rigid_body.cpp:126-142
        b200phys_body& b = bodies[i]; const BodyG& G = g[i];
        V3 v = G.v, w = G.w;
        Q4 dq = q4(0.5f * w.x, 0.5f * w.y, 0.5f * w.z, 0.f);
        dq = dq * G.rotation;
        Q4 sum = q4(G.rotation.x + dq.x * dt, G.rotation.y + dq.y * dt, G.rotation.z + dq.z * dt, G.rotation.w + dq.w * dt);
        Q4 rot = normalize(sum);
        V3 pos = G.position + v * dt;
        st3(b.linearVelocity, v); st3(b.angularVelocity, w);
        st3(b.forceAccumulator, v3(0.f)); st3(b.torqueAccumulator, v3(0.f));
        b.rotation[0] = rot.x; b.rotation[1] = rot.y; b.rotation[2] = rot.z; b.rotation[3] = rot.w;
        st3(b.position, pos - rot * G.localCOG);
    }
}

// ------------------------------------------------------------------------------------------- contact constraints
void Oracle::initContacts(float dt)
{
    const uint32_t K = (uint32_t)contacts.size();
    rows.resize(K);
    float invDt = 1.f / dt;
    for (uint32_t i = 0; i < K; ++i)
    {   // constraints.cpp:3316-3371
        ContactRow& r = rows[i]; const Contact& c = contacts[i];
        const BodyG& A = g[contactBodies[i].rbA]; const BodyG& B = g[contactBodies[i].rbB];
        r.impN = 0.f; r.impT = 0.f;
        r.rA = c.point - A.position; r.rB = c.point - B.position;
        V3 avA = A.v + cross(A.w, r.rA), avB = B.v + cross(B.w, r.rB);
        V3 rel = avB - avA;
        r.tangent = noz(rel - dot(c.normal, rel) * c.normal);
        {
            V3 crAt = cross(r.rA, r.tangent), crBt = cross(r.rB, r.tangent);
            float inv = A.invMass + dot(crAt, A.invInertia * crAt) + B.invMass + dot(crBt, B.invInertia * crBt);
            r.massT = (inv != 0.f) ? (1.f / inv) : 0.f;
            r.tAngA = A.invInertia * crAt; r.tAngB = B.invInertia * crBt;
        }
        {
            V3 crAn = cross(r.rA, c.normal), crBn = cross(r.rB, c.normal);
            float inv = A.invMass + dot(crAn, A.invInertia * crAn) + B.invMass + dot(crBn, B.invInertia * crBn);
            r.massN = (inv != 0.f) ? (1.f / inv) : 0.f;
            r.bias = 0.f;
            if (dt > 1e-5f)
            {
                float vRel = dot(c.normal, rel);
                const float slop = -0.001f;
                if (-c.depth < slop && vRel < 0.f)
                {
                    float restitution = (float)(c.fr & 0xFFFF) / (float)0xFFFF;
                    r.bias = -restitution * vRel - 0.1f * (-c.depth - slop) * invDt;
                }
            }
            r.nAngA = A.invInertia * crAn; r.nAngB = B.invInertia * crBn;
        }
    }
}

void Oracle::solveContacts(const uint32_t* order, uint32_t n)
{
    for (uint32_t k = 0; k < n; ++k)
    {   // constraints.cpp:3385-3448
        uint32_t i = order ? order[k] : k;
        const Contact& c = contacts[i]; ContactRow& r = rows[i];
        BodyG& A = g[contactBodies[i].rbA]; BodyG& B = g[contactBodies[i].rbB];
        if (A.invMass == 0.f && B.invMass == 0.f) continue;
        V3 vA = A.v, wA = A.w, vB = B.v, wB = B.w;
        {
            V3 rel = (vB + cross(wB, r.rB)) - (vA + cross(wA, r.rA));
            float vt = dot(rel, r.tangent);
            float lambda = -r.massT * vt;
            float friction = (float)(c.fr >> 16) / (float)0xFFFF;
            float maxF = friction * r.impN;
            if (!(maxF >= 0.f)) ++nstats.asserts;
            float ni = clampf(r.impT + lambda, -maxF, maxF);
            lambda = ni - r.impT; r.impT = ni;
            V3 P = lambda * r.tangent;
            vA -= A.invMass * P; wA -= r.tAngA * lambda;
            vB += B.invMass * P; wB += r.tAngB * lambda;
        }
        {
            V3 rel = (vB + cross(wB, r.rB)) - (vA + cross(wA, r.rA));
            float vn = dot(rel, c.normal);
            float lambda = -r.massN * (vn - r.bias);
            float imp = fmax2(r.impN + lambda, 0.f);
            lambda = imp - r.impN; r.impN = imp;
            V3 P = lambda * c.normal;
            vA -= A.invMass * P; wA -= r.nAngA * lambda;
            vB += B.invMass * P; wB += r.nAngB * lambda;
        }
        A.v = vA; A.w = wA; B.v = vB; B.w = wB;
    }
}

// ------------------------------------------------------------------------------------------------------- the step
// handleCollisionCallbacks (physics.cpp:1037-1178): this frame's colliding pairs, sorted, are merged against last frame's;
// pairs only in the new list raise a begin event (contact point / normal averaged over the manifold, relative velocity of the
// two bodies at that point from the post-force, pre-solve global state), pairs only in the old list an end event.
// The reference keys by EnTT entity handles; collider indices (its own dense stand-in for them, physics.cpp:1058-1059) are used here.
void Oracle::collisionEvents()
{
    std::vector<CollisionKey> cur;
    uint32_t contactOffset = 0;
    for (size_t i = 0; i < collisions.size(); ++i)
    {
        if (!(collisions[i].b < wcol.size())) continue;        // heightmap collisions carry no second collider (physics.cpp:1049)
        cur.push_back(CollisionKey{ collisions[i].a, collisions[i].b, contactOffset, contactCount[i] });
        contactOffset += contactCount[i];
    }
    auto less = [](const CollisionKey& x, const CollisionKey& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; };
    auto same = [](const CollisionKey& x, const CollisionKey& y) { return x.a == y.a && x.b == y.b; };
    std::sort(cur.begin(), cur.end(), less);
    beginEvents.clear(); endEvents.clear();
    const uint32_t dummy = (uint32_t)bodies.size();
    auto begin = [&](const CollisionKey& k)
    {
        const float norm = 1.f / (float)k.numContacts;
        V3 point = v3(0.f), normal = v3(0.f);
        for (uint32_t i = 0; i < k.numContacts; ++i) { point += contacts[k.contactOffset + i].point; normal += contacts[k.contactOffset + i].normal; }
        point *= norm; normal *= norm;
        const WCol& ca = wcol[k.a]; const WCol& cb = wcol[k.b];
        const BodyG& A = g[ca.objType == B200PHYS_OBJ_RIGID_BODY ? ca.objIndex : dummy];
        const BodyG& B = g[cb.objType == B200PHYS_OBJ_RIGID_BODY ? cb.objIndex : dummy];
        const V3 velA = A.v + cross(A.w, point - A.position), velB = B.v + cross(B.w, point - B.position);
        beginEvents.push_back(BeginEvent{ k.a, k.b, point, normal, velB - velA });
    };
    size_t p = 0, t = 0;
    while (p < prevFrameCollisions.size() && t < cur.size())
    {
        if (same(prevFrameCollisions[p], cur[t])) { ++p; ++t; continue; }
        if (less(prevFrameCollisions[p], cur[t])) { endEvents.push_back(Pair{ prevFrameCollisions[p].a, prevFrameCollisions[p].b }); ++p; }
        else { begin(cur[t]); ++t; }
    }
    while (p < prevFrameCollisions.size()) { endEvents.push_back(Pair{ prevFrameCollisions[p].a, prevFrameCollisions[p].b }); ++p; }
    while (t < cur.size()) { begin(cur[t]); ++t; }
    prevFrameCollisions = std::move(cur);
}

void Oracle::step(float dt, uint32_t iterations, const uint32_t* contactOrder, const uint32_t* const jointOrder[6])
{
    nstats = NarrowStats();
    double t0 = nowMs();
    worldColliders();                       double t1 = nowMs();
    broadphaseSAP();                        double t2 = nowMs();
    narrowphase(); heightmapCollisions();   double t3 = nowMs();
    nonCollisionInteractions();             // physics.cpp:1255
    integrateForces(dt);                    double t4 = nowMs();
    collisionEvents();                      // physics.cpp:1284 (after the force integration, before the solver)
    initJoints(dt);
    initContacts(dt);                       double t5 = nowMs();
    for (uint32_t it = 0; it < iterations; ++it)
    {   // kind order per iteration: distance, ball, fixed, hinge, cone-twist, slider, collision (constraints.cpp:3748-3772)
        for (int k = 0; k < 6; ++k) solveJoints(k, jointOrder ? jointOrder[k] : nullptr);
        solveContacts(contactOrder, (uint32_t)contacts.size());
    }
    double t6 = nowMs();
    integrateVelocities(dt);                double t7 = nowMs();
    msWorld = t1 - t0; msBroad = t2 - t1; msNarrow = t3 - t2; msForces = t4 - t3; msInit = t5 - t4; msSolve = t6 - t5; msIntegrate = t7 - t6;
}

} // namespace orc
