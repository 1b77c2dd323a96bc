// oracle/om_mass.cpp -- TEST INFRASTRUCTURE ONLY (parity oracle). Never linked into the product path.
// CPU restatement of collider_union::calculatePhysicsProperties (reference src/physics/physics.cpp:1416-1588) and
// rigid_body_component::recalculateProperties (src/physics/rigid_body.cpp:29-81). Scene-build time, not per step.
#include "om_step.h"

namespace orc {

static inline V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }

void massProperties(const b200phys_collider& c, const std::vector<HullGeom>& hulls, float inertia9[9], float cog3[3], float* massOut)
{
    const float* s = c.shape;
    M3 inertia = m3Zero(); V3 cog = v3(0.f); float mass = 0.f;
    switch (c.type)
    {
    case T_SPHERE:
    {   // bounding_volumes.h:35-41, physics.cpp:1421-1426
        float r = s[3];
        float sq = r * r, sqPi = kPi * sq;
        float volume = 4.f / 3.f * sqPi * r;
        mass = volume * c.density; cog = ld3(s);
        inertia = m3Identity() * (2.f / 5.f * mass * r * r);
    } break;
    case T_CAPSULE:
    {   // physics.cpp:1428-1464, bounding_volumes.h:50-58
        V3 pA = ld3(s), pB = ld3(s + 3); float r = s[6];
        V3 axis = pA - pB;
        if (axis.y < 0.f) axis *= -1.f;
        float height = len(axis);
        axis *= (1.f / height);
        Q4 rotation = rotateFromTo(v3(0.f, 1.f, 0.f), axis);
        M3 rot = quaternionToMat3(rotation);
        float sq = r * r, sqPi = kPi * sq;
        float sphereVolume = 4.f / 3.f * sqPi * r;
        float h2 = len(pA - pB);
        float volume = sphereVolume + sqPi * h2;
        mass = volume * c.density; cog = (pA + pB) * 0.5f;
        float cylMass = c.density * sqPi * height;
        float hemiMass = c.density * 2.f / 3.f * sqPi * r;
        float sqH = height * height;
        M3 I = m3Zero();
        I.at(1, 1) = sq * cylMass * 0.5f;
        I.at(0, 0) = I.at(2, 2) = I.at(1, 1) * 0.5f + cylMass * sqH / 12.f;
        float t0 = hemiMass * 2.f * sq / 5.f;
        I.at(1, 1) += t0 * 2.f;
        float t1 = height * 0.5f;
        float t2 = t0 + hemiMass * (t1 * t1 + 3.f / 8.f * sqH);
        I.at(0, 0) += t2 * 2.f; I.at(2, 2) += t2 * 2.f;
        inertia = transpose(rot) * I * rot;
    } break;
    case T_CYLINDER:
    {   // physics.cpp:1466-1491, bounding_volumes.h:67-73
        V3 pA = ld3(s), pB = ld3(s + 3); float r = s[6];
        V3 axis = pA - pB;
        if (axis.y < 0.f) axis *= -1.f;
        float height = len(axis);
        axis *= (1.f / height);
        Q4 rotation = rotateFromTo(v3(0.f, 1.f, 0.f), axis);
        M3 rot = quaternionToMat3(rotation);
        float sqPi = kPi * r * r;
        float volume = sqPi * len(pA - pB);
        mass = volume * c.density; cog = (pA + pB) * 0.5f;
        float sq = r * r, sqH = height * height;
        M3 I = m3Zero();
        I.at(1, 1) = sq * mass * 0.5f;
        I.at(0, 0) = I.at(2, 2) = 1.f / 12.f * mass * (3.f * sq + sqH);
        inertia = transpose(rot) * I * rot;
    } break;
    case T_AABB:
    {   // physics.cpp:1493-1503
        Box b = { ld3(s), ld3(s + 3) };
        V3 dm = b.mx - b.mn;
        mass = dm.x * dm.y * dm.z * c.density; cog = boxCenter(b);
        V3 d = boxRadius(b) * 2.f;
        inertia.at(0, 0) = 1.f / 12.f * mass * (d.y * d.y + d.z * d.z);
        inertia.at(1, 1) = 1.f / 12.f * mass * (d.x * d.x + d.z * d.z);
        inertia.at(2, 2) = 1.f / 12.f * mass * (d.x * d.x + d.y * d.y);
    } break;
    case T_OBB:
    {   // physics.cpp:1505-1518
        Q4 q = q4(s[0], s[1], s[2], s[3]); V3 radius = ld3(s + 7);
        V3 d = radius * 2.f;
        mass = d.x * d.y * d.z * c.density; cog = ld3(s + 4);
        M3 I = m3Zero();
        I.at(0, 0) = 1.f / 12.f * mass * (d.y * d.y + d.z * d.z);
        I.at(1, 1) = 1.f / 12.f * mass * (d.x * d.x + d.z * d.z);
        I.at(2, 2) = 1.f / 12.f * mass * (d.x * d.x + d.y * d.y);
        M3 rot = quaternionToMat3(q);
        inertia = transpose(rot) * I * rot;
    } break;
    default:
    {   // hull, physics.cpp:1520-1580 (tetrahedron covariance accumulation)
        uint32_t gi; memcpy(&gi, &s[7], 4);
        const HullGeom& geo = hulls[gi];
        Q4 q = q4(s[0], s[1], s[2], s[3]); V3 p = ld3(s + 4);
        const float s60 = 1.f / 60.f, s120 = 1.f / 120.f;
        M3 Cc; Cc.at(0, 0) = s60; Cc.at(0, 1) = s120; Cc.at(0, 2) = s120; Cc.at(1, 0) = s120; Cc.at(1, 1) = s60; Cc.at(1, 2) = s120; Cc.at(2, 0) = s120; Cc.at(2, 1) = s120; Cc.at(2, 2) = s60;
        float totalMass = 0.f; M3 totalCov = m3Zero(); V3 totalCog = v3(0.f);
        for (size_t f = 0; f + 2 < geo.tris.size(); f += 3)
        {
            V3 w1 = p + q * geo.verts[geo.tris[f]], w2 = p + q * geo.verts[geo.tris[f + 1]], w3 = p + q * geo.verts[geo.tris[f + 2]];
            M3 A; A.at(0, 0) = w1.x; A.at(0, 1) = w2.x; A.at(0, 2) = w3.x; A.at(1, 0) = w1.y; A.at(1, 1) = w2.y; A.at(1, 2) = w3.y; A.at(2, 0) = w1.z; A.at(2, 1) = w2.z; A.at(2, 2) = w3.z;
            float detA = determinant(A);
            M3 cov = A * detA * Cc * transpose(A);  // detA * A * Ccanonical * transpose(A): (detA * A) first, then products left to right
            float volume = 1.f / 6.f * detA;
            V3 tc = (w1 + w2 + w3) * 0.25f;
            totalMass += volume; totalCov = totalCov + cov; totalCog += tc * volume;
        }
        totalCog /= totalMass;
        M3 Cp = totalCov - outerProduct(totalCog, totalCog) * totalMass;
        cog = totalCog; mass = totalMass * c.density;
        inertia = (m3Identity() * trace(Cp) - Cp) * c.density;
    } break;
    }
    memcpy(inertia9, inertia.m, 36); cog3[0] = cog.x; cog3[1] = cog.y; cog3[2] = cog.z; *massOut = mass;
}

} // namespace orc

extern "C" {
// rigid_body.cpp:29-81 : combine per-collider properties (given in the order the reference walks the collider list:
// most recently added collider first, scene.h:53-56) into invMass / localCOG / invInertia
void oracle_combine_mass(const float* inertias9, const float* cogs3, const float* masses, uint32_t n, float* invMassOut, float* cogOut3, float* invInertiaOut9)
{
    using namespace orc;
    M3 inertia = m3Zero(); V3 cog = v3(0.f); float mass = 0.f;
    for (uint32_t i = 0; i < n; ++i) { mass += masses[i]; cog += v3(cogs3[3 * i], cogs3[3 * i + 1], cogs3[3 * i + 2]) * masses[i]; }
    float invMass = 1.f / mass;
    cog = cog * invMass;
    for (uint32_t i = 0; i < n; ++i)
    {
        V3 lc = v3(cogs3[3 * i], cogs3[3 * i + 1], cogs3[3 * i + 2]);
        M3 li; memcpy(li.m, inertias9 + 9 * i, 36);
        V3 r = lc - cog;
        inertia = inertia + (li + (m3Identity() * dot(r, r) - outerProduct(r, r)) * masses[i]);
    }
    M3 inv = invert(inertia);
    *invMassOut = invMass; cogOut3[0] = cog.x; cogOut3[1] = cog.y; cogOut3[2] = cog.z; memcpy(invInertiaOut9, inv.m, 36);
}
void oracle_set_hull_triangles(struct oracle_ctx* c, uint32_t hull, const uint32_t* tris, uint32_t numTris);
}
