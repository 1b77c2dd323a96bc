// oracle/oracle_c.cpp -- TEST INFRASTRUCTURE ONLY (parity oracle). C entry points for ctypes (tests/, bench.py cpu_baseline,
// __graft_entry__.smoke). Never linked into or loaded by the product path.
#include "om_step.h"
#include <cstdio>

using namespace orc;

namespace orc {
// mass properties restatement (om_mass.cpp)
void massProperties(const b200phys_collider& c, const std::vector<HullGeom>& hulls, float inertia9[9], float cog3[3], float* mass);
}

extern "C" {

struct oracle_ctx { Oracle o; };

oracle_ctx* oracle_create() { return new oracle_ctx(); }
void oracle_destroy(oracle_ctx* c) { delete c; }

// 0 = reference SAP order (collision_broad.cpp sweep order), 1 = canonical sorted (a<b, lexicographic)
void oracle_set_pair_order(oracle_ctx* c, int order) { c->o.pairOrder = order; }

void oracle_set_hulls(oracle_ctx* c, const float* verts, const uint32_t* offsets, uint32_t numHulls)
{
    c->o.hulls.clear(); c->o.hulls.resize(numHulls);
    for (uint32_t h = 0; h < numHulls; ++h)
    {   // bounding_volumes.cpp:1385-1397
        HullGeom& g = c->o.hulls[h]; g.aabb = boxNegInf();
        for (uint32_t v = offsets[h]; v < offsets[h + 1]; ++v) { V3 p = v3(verts[3 * v], verts[3 * v + 1], verts[3 * v + 2]); g.verts.push_back(p); grow(g.aabb, p); }
    }
}
void oracle_set_bodies(oracle_ctx* c, const b200phys_body* b, uint32_t n) { c->o.bodies.assign(b, b + n); }
void oracle_set_transforms(oracle_ctx* c, const b200phys_transform* t, uint32_t n) { c->o.transforms.assign(t, t + n); }
void oracle_set_colliders(oracle_ctx* c, const b200phys_collider* col, uint32_t n) { c->o.colliders.assign(col, col + n); }
void oracle_set_joints(oracle_ctx* c, int kind, const void* joints, const b200phys_body_pair* pairs, uint32_t n)
{
    static const size_t sz[6] = { sizeof(b200phys_distance_joint), sizeof(b200phys_ball_joint), sizeof(b200phys_fixed_joint),
                                  sizeof(b200phys_hinge_joint), sizeof(b200phys_cone_twist_joint), sizeof(b200phys_slider_joint) };
    c->o.joints[kind].assign((const uint8_t*)joints, (const uint8_t*)joints + sz[kind] * n);
    c->o.jointPairs[kind].assign(pairs, pairs + n);
}

// one physicsStepInternal; orders may be NULL
void oracle_step(oracle_ctx* c, float dt, uint32_t iterations, const uint32_t* contactOrder,
                 const uint32_t* o0, const uint32_t* o1, const uint32_t* o2, const uint32_t* o3, const uint32_t* o4, const uint32_t* o5)
{
    const uint32_t* jo[6] = { o0, o1, o2, o3, o4, o5 };
    c->o.step(dt, iterations, contactOrder, jo);
}

// staged entry points (tests drive stages separately, e.g. replaying the GPU's solve order on identical contacts)
void oracle_stage_detect(oracle_ctx* c) { c->o.nstats = NarrowStats(); c->o.worldColliders(); c->o.broadphaseSAP(); c->o.narrowphase(); c->o.heightmapCollisions(); }
// heightmap_collider_component (terrain/heightmap_collider.h:132-150): chunksPerDim^2 chunks of 129 x 129 uint16 heights, chunk (x, z) at z * chunksPerDim + x
void oracle_set_heightmap(oracle_ctx* c, uint32_t chunksPerDim, float chunkSize, const float* minCorner3, float amplitudeScale, float restitution, float friction, const uint16_t* heights)
{
    if (!heights) { c->o.heightmap = Heightmap(); return; }
    c->o.heightmap.set(chunksPerDim, chunkSize, v3(minCorner3[0], minCorner3[1], minCorner3[2]), amplitudeScale, restitution, friction, heights);
}
float oracle_heightmap_height_at(oracle_ctx* c, float x, float z) { return c->o.heightmap.heightAt(x, z); }
void oracle_stage_world_and_aabbs(oracle_ctx* c) { c->o.worldColliders(); }
void oracle_stage_forces(oracle_ctx* c, float dt) { c->o.integrateForces(dt); }
void oracle_stage_init(oracle_ctx* c, float dt) { c->o.initJoints(dt); c->o.initContacts(dt); }
void oracle_stage_solve(oracle_ctx* c, uint32_t iterations, const uint32_t* contactOrder,
                        const uint32_t* o0, const uint32_t* o1, const uint32_t* o2, const uint32_t* o3, const uint32_t* o4, const uint32_t* o5)
{
    const uint32_t* jo[6] = { o0, o1, o2, o3, o4, o5 };
    for (uint32_t it = 0; it < iterations; ++it)
    {
        for (int k = 0; k < 6; ++k) c->o.solveJoints(k, jo[k]);
        c->o.solveContacts(contactOrder, (uint32_t)c->o.contacts.size());
    }
}
void oracle_stage_integrate(oracle_ctx* c, float dt) { c->o.integrateVelocities(dt); }

// replace the detected contacts (lets a test feed the solver stages with externally produced contacts)
void oracle_set_contacts(oracle_ctx* c, const b200phys_contact* contacts, const b200phys_body_pair* bodies, uint32_t n)
{
    c->o.contacts.resize(n); c->o.contactBodies.assign(bodies, bodies + n);
    for (uint32_t i = 0; i < n; ++i)
    {
        Contact& k = c->o.contacts[i];
        k.point = v3(contacts[i].point[0], contacts[i].point[1], contacts[i].point[2]); k.depth = contacts[i].penetrationDepth;
        k.normal = v3(contacts[i].normal[0], contacts[i].normal[1], contacts[i].normal[2]); k.fr = contacts[i].friction_restitution;
    }
}

uint32_t oracle_num_bodies(oracle_ctx* c) { return (uint32_t)c->o.bodies.size(); }
void oracle_get_bodies(oracle_ctx* c, b200phys_body* out) { memcpy(out, c->o.bodies.data(), c->o.bodies.size() * sizeof(b200phys_body)); }
void oracle_get_body_states(oracle_ctx* c, b200phys_body_state* out)
{
    for (size_t i = 0; i < c->o.bodies.size(); ++i)
    {
        const b200phys_body& b = c->o.bodies[i];
        memcpy(out[i].rotation, b.rotation, 16); memcpy(out[i].position, b.position, 12);
        memcpy(out[i].linearVelocity, b.linearVelocity, 12); memcpy(out[i].angularVelocity, b.angularVelocity, 12);
    }
}
// global state after integrateForces / solve: v, w per body (6 floats each)
void oracle_get_global_velocities(oracle_ctx* c, float* out6)
{
    for (size_t i = 0; i + 1 < c->o.g.size(); ++i) { const BodyG& g = c->o.g[i]; float* o = out6 + 6 * i; o[0] = g.v.x; o[1] = g.v.y; o[2] = g.v.z; o[3] = g.w.x; o[4] = g.w.y; o[5] = g.w.z; }
}
uint32_t oracle_num_colliders(oracle_ctx* c) { return (uint32_t)c->o.aabbs.size(); }
void oracle_get_aabbs(oracle_ctx* c, float* out6)
{
    for (size_t i = 0; i < c->o.aabbs.size(); ++i) { const Box& b = c->o.aabbs[i]; float* o = out6 + 6 * i; o[0] = b.mn.x; o[1] = b.mn.y; o[2] = b.mn.z; o[3] = b.mx.x; o[4] = b.mx.y; o[5] = b.mx.z; }
}
// world collider dump: type, then 10 shape floats in the b200phys_collider packing
void oracle_get_world_colliders(oracle_ctx* c, int32_t* types, float* shape10)
{
    for (size_t i = 0; i < c->o.wcol.size(); ++i)
    {
        const WCol& w = c->o.wcol[i]; float* s = shape10 + 10 * i; for (int k = 0; k < 10; ++k) s[k] = 0.f;
        types[i] = w.type;
        switch (w.type)
        {
        case T_SPHERE: s[0] = w.sphere.c.x; s[1] = w.sphere.c.y; s[2] = w.sphere.c.z; s[3] = w.sphere.r; break;
        case T_CAPSULE: case T_CYLINDER: s[0] = w.cap.a.x; s[1] = w.cap.a.y; s[2] = w.cap.a.z; s[3] = w.cap.b.x; s[4] = w.cap.b.y; s[5] = w.cap.b.z; s[6] = w.cap.r; break;
        case T_AABB: s[0] = w.box.mn.x; s[1] = w.box.mn.y; s[2] = w.box.mn.z; s[3] = w.box.mx.x; s[4] = w.box.mx.y; s[5] = w.box.mx.z; break;
        case T_OBB: s[0] = w.obb.q.x; s[1] = w.obb.q.y; s[2] = w.obb.q.z; s[3] = w.obb.q.w; s[4] = w.obb.c.x; s[5] = w.obb.c.y; s[6] = w.obb.c.z; s[7] = w.obb.r.x; s[8] = w.obb.r.y; s[9] = w.obb.r.z; break;
        default: s[0] = w.hull.q.x; s[1] = w.hull.q.y; s[2] = w.hull.q.z; s[3] = w.hull.q.w; s[4] = w.hull.p.x; s[5] = w.hull.p.y; s[6] = w.hull.p.z; break;
        }
    }
}
uint32_t oracle_num_pairs(oracle_ctx* c) { return (uint32_t)c->o.pairs.size(); }
void oracle_get_pairs(oracle_ctx* c, b200phys_collider_pair* out) { for (size_t i = 0; i < c->o.pairs.size(); ++i) { out[i].a = c->o.pairs[i].a; out[i].b = c->o.pairs[i].b; } }
// brute-force all-pairs enumeration of the same predicate (independent check of both SAP and the GPU grid)
uint32_t oracle_brute_pairs(oracle_ctx* c, b200phys_collider_pair* out, uint32_t cap)
{
    std::vector<Pair> p; c->o.broadphaseBrute(p);
    for (size_t i = 0; i < p.size() && i < cap; ++i) { out[i].a = p[i].a; out[i].b = p[i].b; }
    return (uint32_t)p.size();
}
uint32_t oracle_num_contacts(oracle_ctx* c) { return (uint32_t)c->o.contacts.size(); }
void oracle_get_contacts(oracle_ctx* c, b200phys_contact* out, b200phys_body_pair* bodies)
{
    for (size_t i = 0; i < c->o.contacts.size(); ++i)
    {
        const Contact& k = c->o.contacts[i];
        out[i].point[0] = k.point.x; out[i].point[1] = k.point.y; out[i].point[2] = k.point.z; out[i].penetrationDepth = k.depth;
        out[i].normal[0] = k.normal.x; out[i].normal[1] = k.normal.y; out[i].normal[2] = k.normal.z; out[i].friction_restitution = k.fr;
        if (bodies) bodies[i] = c->o.contactBodies[i];
    }
}
// localized force fields (one world-space force per force-field index) and the global one (physics.cpp:759-787)
void oracle_set_force_fields(oracle_ctx* c, const float* forces3, uint32_t count, const float* global3)
{
    c->o.forceFields.clear();
    for (uint32_t i = 0; i < count; ++i) c->o.forceFields.push_back(v3(forces3[3 * i], forces3[3 * i + 1], forces3[3 * i + 2]));
    c->o.globalForceField = global3 ? v3(global3[0], global3[1], global3[2]) : v3(0.f);
}
uint32_t oracle_num_trigger_events(oracle_ctx* c, uint32_t* numLeave) { *numLeave = (uint32_t)c->o.triggerLeave.size(); return (uint32_t)c->o.triggerEnter.size(); }
void oracle_get_trigger_events(oracle_ctx* c, uint32_t* enter2, uint32_t* leave2)
{
    for (size_t i = 0; i < c->o.triggerEnter.size(); ++i) { enter2[2 * i] = c->o.triggerEnter[i].a; enter2[2 * i + 1] = c->o.triggerEnter[i].b; }
    for (size_t i = 0; i < c->o.triggerLeave.size(); ++i) { leave2[2 * i] = c->o.triggerLeave[i].a; leave2[2 * i + 1] = c->o.triggerLeave[i].b; }
}
uint32_t oracle_num_interactions(oracle_ctx* c) { return (uint32_t)c->o.interactions.size(); }
// collision begin / end events of the last step: begin = 11 floats (a, b as uint32 bits, point, normal, relative velocity)
uint32_t oracle_num_begin_events(oracle_ctx* c) { return (uint32_t)c->o.beginEvents.size(); }
uint32_t oracle_num_end_events(oracle_ctx* c) { return (uint32_t)c->o.endEvents.size(); }
void oracle_get_events(oracle_ctx* c, uint32_t* beginPairs, float* beginPayload9, uint32_t* endPairs)
{
    for (size_t i = 0; i < c->o.beginEvents.size(); ++i)
    {
        const auto& e = c->o.beginEvents[i];
        beginPairs[2 * i] = e.a; beginPairs[2 * i + 1] = e.b;
        const float v[9] = { e.point.x, e.point.y, e.point.z, e.normal.x, e.normal.y, e.normal.z, e.relativeVelocity.x, e.relativeVelocity.y, e.relativeVelocity.z };
        for (int k = 0; k < 9; ++k) beginPayload9[9 * i + k] = v[k];
    }
    for (size_t i = 0; i < c->o.endEvents.size(); ++i) { endPairs[2 * i] = c->o.endEvents[i].a; endPairs[2 * i + 1] = c->o.endEvents[i].b; }
}
uint32_t oracle_num_collisions(oracle_ctx* c) { return (uint32_t)c->o.collisions.size(); }
void oracle_get_collisions(oracle_ctx* c, b200phys_collider_pair* out, uint8_t* counts)
{
    for (size_t i = 0; i < c->o.collisions.size(); ++i) { out[i].a = c->o.collisions[i].a; out[i].b = c->o.collisions[i].b; counts[i] = c->o.contactCount[i]; }
}
// accumulated impulses of contact rows after solving: (normal, tangent) per contact
void oracle_get_contact_impulses(oracle_ctx* c, float* out2)
{
    for (size_t i = 0; i < c->o.rows.size(); ++i) { out2[2 * i] = c->o.rows[i].impN; out2[2 * i + 1] = c->o.rows[i].impT; }
}
// 8 diagnostic counters: gjkIterCap, gjkErrors, epaNotConverged, epaOutOfMemory, asserts, epaMaxTriangles, epaMaxEdges, epaMaxPoints
void oracle_get_narrow_stats(oracle_ctx* c, uint32_t* out8)
{
    const NarrowStats& s = c->o.nstats;
    out8[0] = s.gjkIterCap; out8[1] = s.gjkErrors; out8[2] = s.epaNotConverged; out8[3] = s.epaOutOfMemory; out8[4] = s.asserts;
    out8[5] = s.epaMaxTriangles; out8[6] = s.epaMaxEdges; out8[7] = s.epaMaxPoints;
}
// per-stage wall time of the last oracle_step in ms, reference stage names:
// world colliders, broad phase, narrow phase, integrate forces, initialize constraints, solve constraints, integrate velocities
void oracle_get_stage_ms(oracle_ctx* c, double* out7)
{
    out7[0] = c->o.msWorld; out7[1] = c->o.msBroad; out7[2] = c->o.msNarrow; out7[3] = c->o.msForces; out7[4] = c->o.msInit; out7[5] = c->o.msSolve; out7[6] = c->o.msIntegrate;
}

// single pair narrow phase in world space (KAT tests): colliders are given ALREADY in world space in the b200phys_collider
// packing (type + shape[10]); returns contact count, writes normal(3) and up to 4 x (point xyz, depth)
int oracle_collide_pair(oracle_ctx* c, const b200phys_collider* a, const b200phys_collider* b, float* normal3, float* pointsDepth16)
{
    Oracle tmp; tmp.hulls = c->o.hulls;
    tmp.colliders.push_back(*a); tmp.colliders.push_back(*b);
    tmp.colliders[0].objectType = O_STATIC; tmp.colliders[1].objectType = O_STATIC;
    tmp.colliders[0].parentTransform = B200PHYS_NO_TRANSFORM; tmp.colliders[1].parentTransform = B200PHYS_NO_TRANSFORM;
    tmp.worldColliders();
    WCol A = tmp.wcol[0], B = tmp.wcol[1];
    if (!(A.type < B.type)) std::swap(A, B);
    Manifold m; NarrowStats st;
    if (!collide(A, B, m, st)) return 0;
    normal3[0] = m.normal.x; normal3[1] = m.normal.y; normal3[2] = m.normal.z;
    for (uint32_t k = 0; k < m.n; ++k) { pointsDepth16[4 * k] = m.point[k].x; pointsDepth16[4 * k + 1] = m.point[k].y; pointsDepth16[4 * k + 2] = m.point[k].z; pointsDepth16[4 * k + 3] = m.depth[k]; }
    return (int)m.n;
}

// boolean overlap test of two colliders (overlapCheck, collision_narrow.cpp:1593-1689): 1 / 0
int oracle_overlap_pair(oracle_ctx* c, const b200phys_collider* a, const b200phys_collider* b)
{
    Oracle tmp; tmp.hulls = c->o.hulls;
    tmp.colliders.push_back(*a); tmp.colliders.push_back(*b);
    tmp.colliders[0].objectType = O_STATIC; tmp.colliders[1].objectType = O_STATIC;
    tmp.colliders[0].parentTransform = B200PHYS_NO_TRANSFORM; tmp.colliders[1].parentTransform = B200PHYS_NO_TRANSFORM;
    tmp.worldColliders();
    WCol A = tmp.wcol[0], B = tmp.wcol[1];
    if (!(A.type < B.type)) std::swap(A, B);
    NarrowStats st;
    return overlapOnly(A, B, st) ? 1 : 0;
}

void oracle_mass_properties(oracle_ctx* c, const b200phys_collider* col, float* inertia9, float* cog3, float* mass)
{
    massProperties(*col, c->o.hulls, inertia9, cog3, mass);
}

} // extern "C"

extern "C" void oracle_set_hull_triangles(oracle_ctx* c, uint32_t hull, const uint32_t* tris, uint32_t numTris)
{
    c->o.hulls[hull].tris.assign(tris, tris + 3 * numTris);
}
