// oracle/om_cloth.cpp -- TEST INFRASTRUCTURE ONLY (parity oracle). Never linked into the product path.
// CPU restatement of the reference's cloth: src/physics/cloth.{h,cpp} -- a grid of particles with stretch / shear / bend
// distance constraints solved by sequential Gauss-Seidel sweeps in creation order (velocity, position and drift passes), simulated at the
// end of physicsStepInternal after the global force field has been applied as wind (physics.cpp:1354-1358).
#include "om_math.h"
#include <vector>
#include <cstdint>
#include <utility>

namespace orc {

struct Cloth
{
    float totalMass, gravityFactor, damping, stiffness; uint32_t gridSizeX, gridSizeY; float width, height;
    float oldTotalMass, oldStiffness;
    struct Constraint { uint32_t a, b; float restDistance, inverseMassSum; };
    std::vector<V3> positions, prevPositions, velocities, forces; std::vector<float> invMasses; std::vector<Constraint> constraints;

    V3 particlePosition(float relX, float relY) const
    {   // cloth.cpp:132-138
        V3 p = v3(relX * width, -relY * height, 0.f);
        p.x -= width * 0.5f;
        std::swap(p.y, p.z);
        return p;
    }
    void addConstraint(uint32_t a, uint32_t b) { constraints.push_back({ a, b, len(positions[a] - positions[b]), (invMasses[a] + invMasses[b]) / stiffness }); }   // :319-328
    Cloth(float w, float h, uint32_t gx, uint32_t gy, float mass, float stiff, float damp, float gravity)
    {   // cloth.cpp:7-82
        gridSizeX = gx; gridSizeY = gy; width = w; height = h; gravityFactor = gravity; damping = damp; totalMass = mass; stiffness = stiff;
        const uint32_t n = gx * gy;
        const float invMassPerParticle = n / totalMass;
        velocities.assign(n, v3(0.f)); forces.assign(n, v3(0.f));
        for (uint32_t y = 0; y < gy; ++y)
        {
            const float invMass = (y == 0) ? 0.f : invMassPerParticle;
            for (uint32_t x = 0; x < gx; ++x)
            {
                const float relX = x / (float)(gx - 1), relY = y / (float)(gy - 1);
                const V3 p = particlePosition(relX, relY);
                positions.push_back(p); prevPositions.push_back(p); invMasses.push_back(invMass);
            }
        }
        for (uint32_t y = 0; y < gy; ++y)
            for (uint32_t x = 0; x < gx; ++x)
            {
                const uint32_t index = y * gx + x;
                if (x < gx - 1) addConstraint(index, index + 1);
                if (y < gy - 1) addConstraint(index, index + gx);
                if (x < gx - 1 && y < gy - 1) { addConstraint(index, index + gx + 1); addConstraint(index + gx, index + 1); }
                if (x < gx - 2) addConstraint(index, index + 2);
                if (y < gy - 2) addConstraint(index, index + gx * 2);
            }
        oldTotalMass = totalMass; oldStiffness = stiffness;
    }
    static V3 transformPosition(Q4 rot, V3 pos, V3 scale, V3 p) { return rot * v3(scale.x * p.x, scale.y * p.y, scale.z * p.z) + pos; }     // math.cpp:518-521
    void setFixed(Q4 rot, V3 pos, V3 scale, bool moveRigid)
    {   // cloth.cpp:84-130
        if (moveRigid)
        {
            V3 pivot;
            if (gridSizeX % 2 == 1) pivot = positions[gridSizeX / 2];
            else pivot = (positions[gridSizeX / 2] + positions[gridSizeX / 2 - 1]) * 0.5f;
            const V3 currentAxis = normalize(positions[gridSizeX - 1] - positions[0]);
            const V3 newAxis = normalize(transformPosition(rot, pos, scale, particlePosition(1.f, 0.f)) - transformPosition(rot, pos, scale, particlePosition(0.f, 0.f)));
            const V3 newPivot = transformPosition(rot, pos, scale, particlePosition(0.5f, 0.f));
            const Q4 delta = rotateFromTo(currentAxis, newAxis);
            for (uint32_t y = 1; y < gridSizeY; ++y)
                for (uint32_t x = 0; x < gridSizeX; ++x) { V3& p = positions[y * gridSizeX + x]; p = delta * (p - pivot) + newPivot; }
        }
        for (uint32_t x = 0; x < gridSizeX; ++x) positions[x] = transformPosition(rot, pos, scale, particlePosition(x / (float)(gridSizeX - 1), 0.f));
    }
    void applyWind(V3 force)
    {   // cloth.cpp:145-183
        for (uint32_t y = 0; y < gridSizeY - 1; ++y)
            for (uint32_t x = 0; x < gridSizeX - 1; ++x)
            {
                const uint32_t tl = y * gridSizeX + x, tr = tl + 1, bl = tl + gridSizeX, br = bl + 1;
                {
                    const V3 normal = cross(positions[bl] - positions[tl], positions[tr] - positions[tl]);
                    V3 f = normal * dot(normalize(normal), force);
                    f *= 1.f / 3.f;
                    forces[tl] += f; forces[tr] += f; forces[bl] += f;
                }
                {
                    const V3 normal = cross(positions[tr] - positions[br], positions[bl] - positions[br]);
                    V3 f = normal * dot(normalize(normal), force);
                    f *= 1.f / 3.f;
                    forces[br] += f; forces[tr] += f; forces[bl] += f;
                }
            }
    }
    void recalculate()
    {   // cloth.cpp:330-346
        const uint32_t n = gridSizeX * gridSizeY;
        const float invMassPerParticle = n / totalMass;
        for (float& im : invMasses) im = (im != 0.f) ? invMassPerParticle : 0.f;
        stiffness = clampf(stiffness, 0.01f, 1.f);
        const float invStiffness = 1.f / stiffness;
        for (Constraint& c : constraints) c.inverseMassSum = (invMasses[c.a] + invMasses[c.b]) * invStiffness;
    }
    void solvePositions()
    {   // cloth.cpp:300-317
        for (const Constraint& c : constraints)
            if (c.inverseMassSum > 0.f)
            {
                const V3 delta = positions[c.b] - positions[c.a];
                const float l = sqLen(delta);
                const float sqRest = c.restDistance * c.restDistance;
                if (sqRest + l > 1e-5f)
                {
                    const float k = ((sqRest - l) / (c.inverseMassSum * (sqRest + l)));
                    positions[c.a] -= delta * (k * invMasses[c.a]);
                    positions[c.b] += delta * (k * invMasses[c.b]);
                }
            }
    }
    void simulate(uint32_t velocityIterations, uint32_t positionIterations, uint32_t driftIterations, float dt)
    {   // cloth.cpp:191-286
        if (totalMass != oldTotalMass || stiffness != oldStiffness) { recalculate(); oldTotalMass = totalMass; oldStiffness = stiffness; }
        const float gravityVelocity = -9.81f * dt * gravityFactor;
        const uint32_t n = gridSizeX * gridSizeY;
        for (uint32_t i = 0; i < n; ++i)
        {
            if (invMasses[i] > 0.f) velocities[i].y += gravityVelocity;
            velocities[i] += forces[i] * (invMasses[i] * dt);
            prevPositions[i] = positions[i];
            positions[i] += velocities[i] * dt;
            forces[i] = v3(0.f);
        }
        const float invDt = (dt > 1e-5f) ? (1.f / dt) : 1.f;
        if (velocityIterations > 0)
        {
            struct Temp { V3 gradient; float inv; };
            std::vector<Temp> temp; temp.reserve(constraints.size());
            for (const Constraint& c : constraints)
            {
                Temp t; t.gradient = prevPositions[c.b] - prevPositions[c.a];
                t.inv = (c.inverseMassSum == 0.f) ? 0.f : (1.f / (sqLen(t.gradient) * c.inverseMassSum));
                temp.push_back(t);
            }
            for (uint32_t it = 0; it < velocityIterations; ++it)
                for (size_t i = 0; i < constraints.size(); ++i)
                {   // cloth.cpp:288-298
                    const Constraint& c = constraints[i];
                    const float j = -dot(temp[i].gradient, velocities[c.a] - velocities[c.b]) * temp[i].inv;
                    velocities[c.a] += temp[i].gradient * (j * invMasses[c.a]);
                    velocities[c.b] -= temp[i].gradient * (j * invMasses[c.b]);
                }
            for (uint32_t i = 0; i < n; ++i) positions[i] = prevPositions[i] + velocities[i] * dt;
        }
        if (positionIterations > 0)
        {
            for (uint32_t it = 0; it < positionIterations; ++it) solvePositions();
            for (uint32_t i = 0; i < n; ++i) velocities[i] = (positions[i] - prevPositions[i]) * invDt;
        }
        if (driftIterations > 0)
        {
            for (uint32_t i = 0; i < n; ++i) prevPositions[i] = positions[i];
            for (uint32_t it = 0; it < driftIterations; ++it) solvePositions();
            for (uint32_t i = 0; i < n; ++i) velocities[i] += (positions[i] - prevPositions[i]) * invDt;
        }
        const float dampingFactor = 1.f / (1.f + dt * damping);
        for (uint32_t i = 0; i < n; ++i) velocities[i] *= dampingFactor;
    }
};

} // namespace orc

using namespace orc;
extern "C" {
void* oracle_cloth_create(float width, float height, uint32_t gx, uint32_t gy, float totalMass, float stiffness, float damping, float gravityFactor) { return new Cloth(width, height, gx, gy, totalMass, stiffness, damping, gravityFactor); }
void oracle_cloth_destroy(void* c) { delete (Cloth*)c; }
void oracle_cloth_set_fixed(void* c, const float* rot4, const float* pos3, const float* scale3, int moveRigid)
{ ((Cloth*)c)->setFixed(q4(rot4[0], rot4[1], rot4[2], rot4[3]), v3(pos3[0], pos3[1], pos3[2]), v3(scale3[0], scale3[1], scale3[2]), moveRigid != 0); }
void oracle_cloth_set_properties(void* c, float totalMass, float stiffness, float damping, float gravityFactor) { Cloth& k = *(Cloth*)c; k.totalMass = totalMass; k.stiffness = stiffness; k.damping = damping; k.gravityFactor = gravityFactor; }
void oracle_cloth_wind(void* c, const float* f3) { ((Cloth*)c)->applyWind(v3(f3[0], f3[1], f3[2])); }
void oracle_cloth_simulate(void* c, uint32_t vi, uint32_t pi, uint32_t di, float dt) { ((Cloth*)c)->simulate(vi, pi, di, dt); }
void oracle_cloth_get(void* c, float* positions, float* velocities)
{
    const Cloth& k = *(Cloth*)c;
    for (size_t i = 0; i < k.positions.size(); ++i)
    {
        if (positions) { positions[3 * i] = k.positions[i].x; positions[3 * i + 1] = k.positions[i].y; positions[3 * i + 2] = k.positions[i].z; }
        if (velocities) { velocities[3 * i] = k.velocities[i].x; velocities[3 * i + 1] = k.velocities[i].y; velocities[3 * i + 2] = k.velocities[i].z; }
    }
}
}
