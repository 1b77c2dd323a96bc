// b200_ray.cuh -- ray vs collider shape tests shared by the host mirror (host/ray_interaction.cpp: testPhysicsInteraction, physics.cpp:555-629)
// and the device-side RL environment (k_env.cu: the random pushes of learned_locomotion.cpp:448-458). They follow
// src/physics/bounding_volumes.cpp: intersectAABB :215-240, intersectOBB :242-247, intersectTriangle :249-265, intersectSphere :267-293,
// intersectCylinder :295-361, intersectCapsule :363-384, intersectDisk :386-394.
#pragma once
#include "b200_math.cuh"
#include <cfloat>
#include <cmath>
#include <cstring>

namespace b200 {

B200_HD bool raySignBit(float f) { uint32_t u; memcpy(&u, &f, 4); return (u >> 31) != 0u; }

struct lray { f3 o, d; };

B200_HD bool intersectPlane(const lray& r, f3 normal, f3 point, float& t)
{
    const float ndotd = dot(r.d, normal);
    if (fabsf(ndotd) < 1e-6f) return false;
    t = -(dot(r.o, normal) - dot(normal, point)) / ndotd;
    return true;
}
B200_HD bool intersectDisk(const lray& r, f3 pos, f3 normal, float radius, float& t)
{
    if (!intersectPlane(r, normal, pos, t)) return false;
    return len(r.o + t * r.d - pos) <= radius;
}
B200_HD bool intersectSphere(const lray& r, f3 center, float radius, float& t)
{
    const f3 m = r.o - center;
    const float b = dot(m, r.d), c = dot(m, m) - radius * radius;
    if (c > 0.f && b > 0.f) return false;
    const float discr = b * b - c;
    if (discr < 0.f) return false;
    t = -b - sqrtf(discr);
    if (t < 0.f) t = 0.f;
    return true;
}
B200_HD bool intersectAABB(const lray& r, f3 mn, f3 mx, float& t)
{
    const f3 inv = mk3(1.f / r.d.x, 1.f / r.d.y, 1.f / r.d.z);      // may be inf, still works (bounding_volumes.cpp:217)
    const float tx1 = (mn.x - r.o.x) * inv.x, tx2 = (mx.x - r.o.x) * inv.x;
    t = fminf(tx1, tx2); float tmax = fmaxf(tx1, tx2);
    const float ty1 = (mn.y - r.o.y) * inv.y, ty2 = (mx.y - r.o.y) * inv.y;
    t = fmaxf(t, fminf(ty1, ty2)); tmax = fminf(tmax, fmaxf(ty1, ty2));
    const float tz1 = (mn.z - r.o.z) * inv.z, tz2 = (mx.z - r.o.z) * inv.z;
    t = fmaxf(t, fminf(tz1, tz2)); tmax = fminf(tmax, fmaxf(tz1, tz2));
    return tmax >= t && t > 0.f;
}
B200_HD bool intersectCylinder(const lray& r, f3 posA, f3 posB, float radius, float& t)
{
    const f3 axis = posB - posA;
    const float height = len(axis);
    const q4 q = rotateFromTo(axis, mk3(0.f, 1.f, 0.f));      // math.cpp:538-575
    const f3 o = q * (r.o - posA), d = q * r.d;                        // base at the origin, second point above it
    const float epsilon = 1e-6f;
    float y = -1.f;
    if (o.x * o.x + o.z * o.z > radius * radius)
    {   // outside the infinite cylinder: only then the side can be hit
        const float a = d.x * d.x + d.z * d.z, b = d.x * o.x + d.z * o.z, c = o.x * o.x + o.z * o.z - radius * radius;
        const float delta = b * b - a * c;
        if (delta < epsilon) return false;
        t = (-b - sqrtf(delta)) / a;
        if (t <= epsilon) return false;
        y = o.y + t * d.y;
    }
    if (y > height + epsilon || y < -epsilon)
    {   // caps
        const lray local = { o, d };
        float dist;
        if (d.y < 0.f && intersectDisk(local, mk3(0.f, height, 0.f), mk3(0.f, 1.f, 0.f), radius, dist)) t = dist;
        if (d.y > 0.f && intersectDisk(local, mk3(0.f, 0.f, 0.f), mk3(0.f, -1.f, 0.f), radius, dist)) t = dist;
        y = o.y + t * d.y;
    }
    return y > -epsilon && y < height + epsilon;
}
B200_HD bool intersectCapsule(const lray& r, f3 a, f3 b, float radius, float& outT)
{
    outT = FLT_MAX; float t = 0.f; bool result = false;
    if (intersectCylinder(r, a, b, radius, t)) { outT = t; result = true; }
    if (intersectSphere(r, a, radius, t)) { outT = fminf(outT, t); result = true; }
    if (intersectSphere(r, b, radius, t)) { outT = fminf(outT, t); result = true; }
    return result;
}
B200_HD bool pointInTriangle(f3 point, f3 A, f3 B, f3 C)
{   // math.cpp:1273-1290: the sign-bit trick spelled out (z < 0 and x >= 0 and y >= 0, with -0 counting as negative like the bit test)
    const f3 e10 = B - A, e20 = C - A;
    const float a = dot(e10, e10), b = dot(e10, e20), c = dot(e20, e20), ac_bb = a * c - b * b;
    const f3 vp = point - A;
    const float d = dot(vp, e10), e = dot(vp, e20);
    const float x = d * c - e * b, y = e * a - d * b, z = x + y - ac_bb;
    return raySignBit(z) && !raySignBit(x) && !raySignBit(y);
}
B200_HD bool intersectTriangle(const lray& r, f3 a, f3 b, f3 c, float& t)
{
    f3 normal = cross(b - a, c - a); const float l = len(normal); normal = l > 1e-6f ? normal * (1.f / l) : mk3(0.f);     // noz
    const float d = -dot(normal, a), nDotR = dot(r.d, normal);
    if (fabsf(nDotR) <= 1e-6f) return false;
    t = -(dot(r.o, normal) + d) / nDotR;
    return t >= 0.f && pointInTriangle(r.o + t * r.d, a, b, c);
}

} // namespace b200
