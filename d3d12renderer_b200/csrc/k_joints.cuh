// k_joints.cuh -- joint constraints on the device: distance, ball, fixed, hinge, cone-twist, slider.
// Reference (scalar path): src/physics/constraints.cpp :189-264 distance, :460-528 ball, :736-823 fixed, :1079-1307 hinge,
// :1782-2070 cone-twist, :2638-2846 slider. One thread per joint: initialise a row once per step from the two body global
// states, then the persistent solver kernel (k_solver.cu) visits the rows colour by colour.
// Included by k_solver.cu only (same translation unit as the solver kernel, no device linking).
#pragma once
#include "b200_context.h"
#include "b200_util.cuh"

namespace b200 {

#define B200_DISTANCE_BETA 0.1f
#define B200_BALL_BETA 0.1f
#define B200_SLIDER_BETA 0.1f
#define B200_HINGE_ROT_BETA 0.3f
#define B200_HINGE_LIMIT_BETA 0.1f
#define B200_TWIST_LIMIT_BETA 0.1f
#define B200_SLIDER_LIMIT_BETA 0.1f
#define B200_DT_THRESHOLD 1e-5f   /* constraints.cpp:9-17 */

struct DistanceRowD { uint32_t a, b; f3 rA, rB, angA, angB, u; float bias, mass; };
struct BallRowD { uint32_t a, b; f3 rA, rB, bias; m3 invMass; };
struct FixedRowD { uint32_t a, b; f3 rA, rB, tBias; m3 invTMass; f3 rBias; m3 invRMass; };
struct HingeRowD
{
    uint32_t a, b; f3 rA, rB, tBias; m3 invTMass; float rBiasX, rBiasY; m2 invRMass; f3 bxa, cxa;
    f3 axis; float axialMass; uint32_t solveLimit, solveMotor;
    float limitImpulse, limitBias, limitSign, motorImpulse, maxMotorImpulse, motorVelocity; f3 angA, angB;
};
struct ConeTwistRowD
{
    uint32_t a, b; f3 rA, rB, bias; m3 invMass;
    uint32_t solveSwingLimit, solveTwistLimit, solveSwingMotor, solveTwistMotor;
    f3 swingAxis; float swingImpulse, swingLimitMass, swingLimitBias;
    f3 twistAxis; float twistImpulse, twistLimitSign, twistMass, twistLimitBias;
    float swingMotorImpulse, maxSwingMotorImpulse, swingMotorVelocity; f3 swingMotorAxis; float swingMotorMass;
    float twistMotorImpulse, maxTwistMotorImpulse, twistMotorVelocity;
    f3 twistAngA, twistAngB, swingMotorAngA, swingMotorAngB, swingLimitAngA, swingLimitAngB;
};
struct SliderRowD
{
    uint32_t a, b; f3 rAuxt, rAuxb, rBxt, rBxb, tangent, bitangent; m2 invTMass; float tBiasX, tBiasY; m3 invRMass; f3 rBias;
    uint32_t solveLimit, solveMotor; f3 axis; float axialMass, limitBias, limitImpulse, limitSign; f3 rAuxs, rBxs, limAngA, limAngB;
    float motorVelocity, motorImpulse, maxMotorImpulse;
};

static const size_t kJointRowBytes[6] = { sizeof(DistanceRowD), sizeof(BallRowD), sizeof(FixedRowD), sizeof(HingeRowD), sizeof(ConeTwistRowD), sizeof(SliderRowD) };
static const size_t kJointInBytes[6] = { sizeof(b200phys_distance_joint), sizeof(b200phys_ball_joint), sizeof(b200phys_fixed_joint),
                                         sizeof(b200phys_hinge_joint), sizeof(b200phys_cone_twist_joint), sizeof(b200phys_slider_joint) };

struct JointSolveArgs
{
    const uint32_t* order[6];       // joints sorted by colour
    const uint32_t* colorStart[6];  // colour ranges
    void* rows[6];                  // row of joint j = rows[kind][j]
    const uint2* bodies[6];         // body pair of joint j
    const uint2* dep[6];            // dataflow record of solve position p (see k_solver.cu)
};

// body global state as the joint code sees it (rigid_body_global_state, rigid_body.h:6-16)
struct BodyGD { q4 rot; f3 cog, pos; m3 I; float invMass; };
B200_D m3 loadInvInertiaJ(const float4* __restrict__ gInvI, uint32_t b)
{
    const float4 r0 = gInvI[3 * b], r1 = gInvI[3 * b + 1], r2 = gInvI[3 * b + 2];
    m3 I;
    I.m[0] = r0.x; I.m[3] = r0.y; I.m[6] = r0.z;
    I.m[1] = r1.x; I.m[4] = r1.y; I.m[7] = r1.z;
    I.m[2] = r2.x; I.m[5] = r2.y; I.m[8] = r2.z;
    return I;
}
B200_D BodyGD loadBodyG(uint32_t b, uint32_t N, const float4* __restrict__ bRot, const float4* __restrict__ bCog, const float4* __restrict__ gPos, const float4* __restrict__ gVel, const float4* __restrict__ gInvI)
{
    BodyGD g;
    if (b >= N) { g.rot = mkq(0.f, 0.f, 0.f, 0.f); g.cog = mk3(0.f); }   // memset dummy (physics.cpp:1279)
    else { g.rot = toq(bRot[b]); g.cog = xyz(bCog[b]); }
    g.pos = xyz(gPos[b]); g.I = loadInvInertiaJ(gInvI, b); g.invMass = gPos[b].w;
    return g;
}
B200_D f3 ld3p(const float* p) { return mk3(p[0], p[1], p[2]); }
B200_D q4 ld4p(const float* p) { return mkq(p[0], p[1], p[2], p[3]); }
// skew(rA) I_A skew(rA)^T + skew(rB) I_B skew(rB)^T + (mA + mB) Id   (constraints.cpp:487-492)
B200_D m3 pointMassMatrix(const BodyGD& A, const BodyGD& B, f3 rA, f3 rB)
{
    const m3 sA = skew(rA), sB = skew(rB);
    return sA * A.I * transpose(sA) + sB * B.I * transpose(sB) + m3Identity() * (A.invMass + B.invMass);
}

struct JointInitArgs
{
    const void* in[6]; const uint2* bodies[6]; void* rows[6]; uint32_t count[6];
    const float4* bRot; const float4* bCog; const float4* gPos; const float4* gVel; const float4* gInvI; uint32_t N; float dt;
    DeviceCounts* cnt;
};

__global__ void __launch_bounds__(128) initJointsKernel(JointInitArgs P)
{
    const float dt = P.dt, invDt = 1.f / P.dt;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, stride = gridDim.x * blockDim.x;
    uint32_t asserts = 0;
    // ---- distance (constraints.cpp:189-239)
    for (uint32_t i = tid; i < P.count[0]; i += stride)
    {
        const b200phys_distance_joint J = ((const b200phys_distance_joint*)P.in[0])[i];
        DistanceRowD o; o.a = P.bodies[0][i].x; o.b = P.bodies[0][i].y;
        const BodyGD A = loadBodyG(o.a, P.N, P.bRot, P.bCog, P.gPos, P.gVel, P.gInvI), B = loadBodyG(o.b, P.N, P.bRot, P.bCog, P.gPos, P.gVel, P.gInvI);
        o.rA = A.rot * (ld3p(J.localAnchorA) - A.cog); o.rB = B.rot * (ld3p(J.localAnchorB) - B.cog);
        const f3 gA = A.pos + o.rA, gB = B.pos + o.rB;
        o.u = gB - gA; const float l = len(o.u);
        o.u = (l > 0.001f) ? (o.u * (1.f / l)) : mk3(0.f);
        const f3 crAu = cross(o.rA, o.u), crBu = cross(o.rB, o.u);
        const float inv = A.invMass + dot(crAu, A.I * crAu) + B.invMass + dot(crBu, B.I * crBu);
        o.mass = (inv != 0.f) ? (1.f / inv) : 0.f;
        o.bias = 0.f;
        if (dt > B200_DT_THRESHOLD) o.bias = (l - J.globalLength) * (B200_DISTANCE_BETA * invDt);
        o.angA = A.I * cross(o.rA, crAu); o.angB = B.I * cross(o.rB, crBu);   // sic (:232-233)
        ((DistanceRowD*)P.rows[0])[i] = o;
    }
    // ---- ball (constraints.cpp:460-503)
    for (uint32_t i = tid; i < P.count[1]; i += stride)
    {
        const b200phys_ball_joint J = ((const b200phys_ball_joint*)P.in[1])[i];
        BallRowD o; o.a = P.bodies[1][i].x; o.b = P.bodies[1][i].y;
        const BodyGD A = loadBodyG(o.a, P.N, P.bRot, P.bCog, P.gPos, P.gVel, P.gInvI), B = loadBodyG(o.b, P.N, P.bRot, P.bCog, P.gPos, P.gVel, P.gInvI);
        o.rA = A.rot * (ld3p(J.localAnchorA) - A.cog); o.rB = B.rot * (ld3p(J.localAnchorB) - B.cog);
        const f3 gA = A.pos + o.rA, gB = B.pos + o.rB;
        o.invMass = pointMassMatrix(A, B, o.rA, o.rB);
        o.bias = mk3(0.f);
        if (dt > B200_DT_THRESHOLD) o.bias = (gB - gA) * (B200_BALL_BETA * invDt);
        ((BallRowD*)P.rows[1])[i] = o;
    }
    // ---- fixed (constraints.cpp:736-786)
    for (uint32_t i = tid; i < P.count[2]; i += stride)
    {
        const b200phys_fixed_joint J = ((const b200phys_fixed_joint*)P.in[2])[i];
        FixedRowD o; o.a = P.bodies[2][i].x; o.b = P.bodies[2][i].y;
        const BodyGD A = loadBodyG(o.a, P.N, P.bRot, P.bCog, P.gPos, P.gVel, P.gInvI), B = loadBodyG(o.b, P.N, P.bRot, P.bCog, P.gPos, P.gVel, P.gInvI);
        o.rA = A.rot * (ld3p(J.localAnchorA) - A.cog); o.rB = B.rot * (ld3p(J.localAnchorB) - B.cog);
        const f3 gA = A.pos + o.rA, gB = B.pos + o.rB;
        o.invTMass = pointMassMatrix(A, B, o.rA, o.rB);
        o.invRMass = A.I + B.I;
        o.tBias = mk3(0.f); o.rBias = mk3(0.f);
        if (dt > B200_DT_THRESHOLD)
        {
            o.tBias = (gB - gA) * (B200_BALL_BETA * invDt);
            const q4 err = B.rot * ld4p(J.initialInvRotationDifference) * conjugate(A.rot);
            o.rBias = qv(err) * (B200_SLIDER_BETA * invDt * 2.f);
        }
        ((FixedRowD*)P.rows[2])[i] = o;
    }
    // ---- hinge (constraints.cpp:1079-1213)
    for (uint32_t i = tid; i < P.count[3]; i += stride)
    {
        const b200phys_hinge_joint J = ((const b200phys_hinge_joint*)P.in[3])[i];
        HingeRowD o; o.a = P.bodies[3][i].x; o.b = P.bodies[3][i].y;
        const BodyGD A = loadBodyG(o.a, P.N, P.bRot, P.bCog, P.gPos, P.gVel, P.gInvI), B = loadBodyG(o.b, P.N, P.bRot, P.bCog, P.gPos, P.gVel, P.gInvI);
        o.rA = A.rot * (ld3p(J.localAnchorA) - A.cog); o.rB = B.rot * (ld3p(J.localAnchorB) - B.cog);
        const f3 gA = A.pos + o.rA, gB = B.pos + o.rB;
        o.invTMass = pointMassMatrix(A, B, o.rA, o.rB);
        o.tBias = mk3(0.f);
        if (dt > B200_DT_THRESHOLD) o.tBias = (gB - gA) * (B200_BALL_BETA * invDt);
        const f3 axA = A.rot * ld3p(J.localHingeAxisA), axB = B.rot * ld3p(J.localHingeAxisB);
        f3 tB, bB; getTangents(axB, tB, bB);
        const f3 bxa = cross(tB, axA), cxa = cross(bB, axA);
        const f3 iAb = A.I * bxa, iBb = B.I * bxa, iAc = A.I * cxa, iBc = B.I * cxa;
        o.invRMass.m00 = dot(bxa, iAb) + dot(bxa, iBb);
        o.invRMass.m01 = dot(bxa, iAc) + dot(bxa, iBc);
        o.invRMass.m10 = dot(cxa, iAb) + dot(cxa, iBb);
        o.invRMass.m11 = dot(cxa, iAc) + dot(cxa, iBc);
        o.bxa = bxa; o.cxa = cxa;
        o.rBiasX = 0.f; o.rBiasY = 0.f;
        if (dt > B200_DT_THRESHOLD) { const float k = B200_HINGE_ROT_BETA * invDt; o.rBiasX = dot(axA, tB) * k; o.rBiasY = dot(axA, bB) * k; }
        o.solveLimit = 0; o.solveMotor = 0; o.axis = mk3(0.f);
        o.limitImpulse = o.limitBias = o.limitSign = o.motorImpulse = o.maxMotorImpulse = o.motorVelocity = o.axialMass = 0.f;
        o.angA = o.angB = mk3(0.f);
        if (J.minRotationLimit <= 0.f || J.maxRotationLimit >= 0.f || J.maxMotorTorque > 0.f)
        {
            const f3 cmp = conjugate(A.rot) * (B.rot * ld3p(J.localHingeTangentB));
            const float angle = atan2f(dot(cmp, ld3p(J.localHingeBitangentA)), dot(cmp, ld3p(J.localHingeTangentA)));
            const bool minV = J.minRotationLimit <= 0.f && angle <= J.minRotationLimit;
            const bool maxV = J.maxRotationLimit >= 0.f && angle >= J.maxRotationLimit;
            if (minV && maxV) ++asserts;
            o.solveLimit = (minV || maxV) ? 1u : 0u;
            o.solveMotor = (J.maxMotorTorque > 0.f) ? 1u : 0u;
            if (o.solveLimit || o.solveMotor)
            {
                o.axis = axA; o.limitImpulse = 0.f;
                const float inv = dot(axA, A.I * axA) + dot(axA, B.I * axA);
                o.axialMass = (inv != 0.f) ? (1.f / inv) : 0.f;
                o.limitSign = minV ? 1.f : -1.f;
                o.maxMotorImpulse = J.maxMotorTorque * dt; o.motorImpulse = 0.f;
                o.angA = A.I * o.axis; o.angB = B.I * o.axis;
                o.motorVelocity = J.motorVelocityOrTargetAngle;
                if (J.motorType == B200PHYS_MOTOR_POSITION)
                {
                    const float minL = (J.minRotationLimit <= 0.f) ? J.minRotationLimit : -B200_PI;
                    const float maxL = (J.maxRotationLimit >= 0.f) ? J.maxRotationLimit : B200_PI;
                    const float target = clampf(J.motorVelocityOrTargetAngle, minL, maxL);
                    o.motorVelocity = (dt > B200_DT_THRESHOLD) ? ((target - angle) * invDt) : 0.f;
                }
                o.limitBias = 0.f;
                if (dt > B200_DT_THRESHOLD)
                {
                    const float d = minV ? (angle - J.minRotationLimit) : (J.maxRotationLimit - angle);
                    o.limitBias = d * B200_HINGE_LIMIT_BETA * invDt;
                }
            }
        }
        ((HingeRowD*)P.rows[3])[i] = o;
    }
    // ---- cone-twist (constraints.cpp:1782-1953)
    for (uint32_t i = tid; i < P.count[4]; i += stride)
    {
        const b200phys_cone_twist_joint J = ((const b200phys_cone_twist_joint*)P.in[4])[i];
        ConeTwistRowD o; memset(&o, 0, sizeof(o));
        o.a = P.bodies[4][i].x; o.b = P.bodies[4][i].y;
        const BodyGD A = loadBodyG(o.a, P.N, P.bRot, P.bCog, P.gPos, P.gVel, P.gInvI), B = loadBodyG(o.b, P.N, P.bRot, P.bCog, P.gPos, P.gVel, P.gInvI);
        o.rA = A.rot * (ld3p(J.localAnchorA) - A.cog); o.rB = B.rot * (ld3p(J.localAnchorB) - B.cog);
        const f3 gA = A.pos + o.rA, gB = B.pos + o.rB;
        o.invMass = pointMassMatrix(A, B, o.rA, o.rB);
        o.bias = mk3(0.f);
        if (dt > B200_DT_THRESHOLD) o.bias = (gB - gA) * (B200_BALL_BETA * invDt);

        const q4 btoa = conjugate(A.rot) * B.rot;
        const f3 limitAxisA = ld3p(J.localLimitAxisA);
        const f3 limitAxisCmpA = btoa * ld3p(J.localLimitAxisB);
        const q4 swingRot = rotateFromTo(limitAxisA, limitAxisCmpA);
        const f3 twistTanA = swingRot * ld3p(J.localLimitTangentA), twistBitanA = swingRot * ld3p(J.localLimitBitangentA);
        const f3 tanCmpA = btoa * ld3p(J.localLimitTangentB);
        const float twistAngle = atan2f(dot(tanCmpA, twistBitanA), dot(tanCmpA, twistTanA));
        f3 swingAxis; float swingAngle; getAxisRotation(swingRot, swingAxis, swingAngle);
        if (swingAngle < 0.f) { swingAngle *= -1.f; swingAxis *= -1.f; }

        o.solveSwingLimit = (J.swingLimit >= 0.f && swingAngle >= J.swingLimit) ? 1u : 0u;
        if (o.solveSwingLimit)
        {
            o.swingImpulse = 0.f;
            o.swingAxis = A.rot * swingAxis;
            const float inv = dot(o.swingAxis, A.I * o.swingAxis) + dot(o.swingAxis, B.I * o.swingAxis);
            o.swingLimitMass = (inv != 0.f) ? (1.f / inv) : 0.f;
            o.swingLimitBias = 0.f;
            if (dt > B200_DT_THRESHOLD) o.swingLimitBias = (J.swingLimit - swingAngle) * (B200_HINGE_LIMIT_BETA * invDt);
            o.swingLimitAngA = A.I * o.swingAxis; o.swingLimitAngB = B.I * o.swingAxis;
        }
        o.solveSwingMotor = (J.maxSwingMotorTorque > 0.f) ? 1u : 0u;
        if (o.solveSwingMotor)
        {
            o.maxSwingMotorImpulse = J.maxSwingMotorTorque * dt; o.swingMotorImpulse = 0.f;
            const float axX = cosf(J.swingMotorAxis), axY = sinf(J.swingMotorAxis);
            const f3 localMotorAxis = axX * ld3p(J.localLimitTangentA) + axY * ld3p(J.localLimitBitangentA);
            if (J.swingMotorType == B200PHYS_MOTOR_VELOCITY)
            {
                o.swingMotorAxis = A.rot * localMotorAxis;
                o.swingMotorVelocity = J.swingMotorVelocityOrTargetAngle;
            }
            else
            {
                float target = J.swingMotorVelocityOrTargetAngle;
                if (J.swingLimit >= 0.f) target = clampf(target, -J.swingLimit, J.swingLimit);
                const f3 targetDir = quatAxisAngle(localMotorAxis, target) * limitAxisA;
                const f3 motorAxis2 = noz(cross(limitAxisCmpA, targetDir));
                o.swingMotorAxis = A.rot * motorAxis2;
                const float cosAngle = dot(targetDir, limitAxisCmpA);
                const float delta = acosf(clamp01(cosAngle));
                o.swingMotorVelocity = (dt > B200_DT_THRESHOLD) ? (delta * invDt * 0.2f) : 0.f;
            }
            o.swingMotorAngA = A.I * o.swingMotorAxis; o.swingMotorAngB = B.I * o.swingMotorAxis;
            const float inv = dot(o.swingMotorAxis, A.I * o.swingMotorAxis) + dot(o.swingMotorAxis, B.I * o.swingMotorAxis);
            o.swingMotorMass = (inv != 0.f) ? (1.f / inv) : 0.f;
        }
        const bool minTw = J.twistLimit >= 0.f && twistAngle <= -J.twistLimit;
        const bool maxTw = J.twistLimit >= 0.f && twistAngle >= J.twistLimit;
        if (minTw && maxTw) ++asserts;
        o.solveTwistLimit = (minTw || maxTw) ? 1u : 0u;
        o.solveTwistMotor = (J.maxTwistMotorTorque > 0.f) ? 1u : 0u;
        if (o.solveTwistLimit || o.solveTwistMotor)
        {
            o.twistImpulse = 0.f;
            o.twistAxis = A.rot * limitAxisA;
            const float inv = dot(o.twistAxis, A.I * o.twistAxis) + dot(o.twistAxis, B.I * o.twistAxis);
            o.twistMass = (inv != 0.f) ? (1.f / inv) : 0.f;
            o.twistLimitSign = minTw ? 1.f : -1.f;
            o.maxTwistMotorImpulse = J.maxTwistMotorTorque * dt; o.twistMotorImpulse = 0.f;
            o.twistAngA = A.I * o.twistAxis; o.twistAngB = B.I * o.twistAxis;
            o.twistMotorVelocity = J.twistMotorVelocityOrTargetAngle;
            if (J.twistMotorType == B200PHYS_MOTOR_POSITION)
            {
                const float limit = (J.twistLimit >= 0.f) ? J.twistLimit : B200_PI;
                const float target = clampf(J.twistMotorVelocityOrTargetAngle, -limit, limit);
                o.twistMotorVelocity = (dt > B200_DT_THRESHOLD) ? ((target - twistAngle) * invDt) : 0.f;
            }
            o.twistLimitBias = 0.f;
            if (dt > B200_DT_THRESHOLD)
            {
                const float d = minTw ? (J.twistLimit + twistAngle) : (J.twistLimit - twistAngle);
                o.twistLimitBias = d * B200_TWIST_LIMIT_BETA * invDt;
            }
        }
        ((ConeTwistRowD*)P.rows[4])[i] = o;
    }
    // ---- slider (constraints.cpp:2638-2761)
    for (uint32_t i = tid; i < P.count[5]; i += stride)
    {
        const b200phys_slider_joint J = ((const b200phys_slider_joint*)P.in[5])[i];
        SliderRowD o; memset(&o, 0, sizeof(o));
        o.a = P.bodies[5][i].x; o.b = P.bodies[5][i].y;
        const BodyGD A = loadBodyG(o.a, P.N, P.bRot, P.bCog, P.gPos, P.gVel, P.gInvI), B = loadBodyG(o.b, P.N, P.bRot, P.bCog, P.gPos, P.gVel, P.gInvI);
        const f3 rA = A.rot * (ld3p(J.localAnchorA) - A.cog), rB = B.rot * (ld3p(J.localAnchorB) - B.cog);
        const f3 gA = A.pos + rA, gB = B.pos + rB;
        const f3 axis = A.rot * ld3p(J.localAxisA);
        getTangents(axis, o.tangent, o.bitangent);
        const f3 u = gB - gA;
        const f3 rAu = rA + u;
        o.rBxt = cross(rB, o.tangent); o.rBxb = cross(rB, o.bitangent);
        o.rAuxt = cross(rAu, o.tangent); o.rAuxb = cross(rAu, o.bitangent);
        const f3 iAt = A.I * o.rAuxt, iAb = A.I * o.rAuxb, iBt = B.I * o.rBxt, iBb = B.I * o.rBxb;
        const float ms = A.invMass + B.invMass;
        o.invTMass.m00 = dot(o.rAuxt, iAt) + dot(o.rBxt, iBt) + ms;
        o.invTMass.m01 = dot(o.rAuxt, iAb) + dot(o.rBxt, iBb);
        o.invTMass.m10 = dot(o.rAuxb, iAt) + dot(o.rBxb, iBt);
        o.invTMass.m11 = dot(o.rAuxb, iAb) + dot(o.rBxb, iBb) + ms;
        o.invRMass = A.I + B.I;
        o.tBiasX = o.tBiasY = 0.f; o.rBias = mk3(0.f);
        if (dt > B200_DT_THRESHOLD)
        {
            const float k = B200_SLIDER_BETA * invDt;
            o.tBiasX = dot(u, o.tangent) * k; o.tBiasY = dot(u, o.bitangent) * k;
            const q4 err = B.rot * ld4p(J.initialInvRotationDifference) * conjugate(A.rot);
            o.rBias = qv(err) * (B200_SLIDER_BETA * invDt * 2.f);
        }
        o.axis = axis;
        const float dist = dot(u, axis);
        o.solveLimit = 0;
        if (J.negDistanceLimit <= 0.f || J.posDistanceLimit >= 0.f)
        {
            const bool minV = (J.negDistanceLimit <= 0.f) && (dist < J.negDistanceLimit);
            const bool maxV = (J.posDistanceLimit >= 0.f) && (dist > J.posDistanceLimit);
            if (minV && maxV) ++asserts;
            if (minV || maxV)
            {
                o.solveLimit = 1; o.limitImpulse = 0.f;
                o.rAuxs = cross(rAu, axis); o.rBxs = cross(rB, axis);
                const float inv = ms + dot(o.rAuxs, A.I * o.rAuxs) + dot(o.rBxs, B.I * o.rBxs);
                o.axialMass = (inv != 0.f) ? (1.f / inv) : 0.f;
                o.limitSign = minV ? 1.f : -1.f;
                o.limitBias = 0.f;
                if (dt > B200_DT_THRESHOLD)
                {
                    const float e = minV ? (dist - J.negDistanceLimit) : (J.posDistanceLimit - dist);
                    o.limitBias = e * (B200_SLIDER_LIMIT_BETA * invDt);
                }
                o.limAngA = A.I * o.rAuxs; o.limAngB = B.I * o.rBxs;
            }
        }
        o.solveMotor = 0;
        if (J.maxMotorForce > 0.f)
        {
            o.solveMotor = 1; o.maxMotorImpulse = J.maxMotorForce * dt; o.motorImpulse = 0.f;
            o.motorVelocity = J.motorVelocityOrTargetDistance;
            if (J.motorType == B200PHYS_MOTOR_POSITION)
            {
                const float minL = (J.negDistanceLimit <= 0.f) ? J.negDistanceLimit : -INFINITY;
                const float maxL = (J.posDistanceLimit >= 0.f) ? J.posDistanceLimit : INFINITY;
                const float target = clampf(J.motorVelocityOrTargetDistance, minL, maxL);
                o.motorVelocity = (dt > B200_DT_THRESHOLD) ? ((target - dist) * invDt) : 0.f;
            }
        }
        ((SliderRowD*)P.rows[5])[i] = o;
    }
    if (asserts) atomicAdd(&P.cnt->assertCount, asserts);
}

// ------------------------------------------------------------------------------------------------------ solving
struct BodyVel { f3 v, w; float invMass; };

// one joint (index i of its kind) applied to the velocities of its two bodies; the caller (k_solver.cu) loads A and B when the
// dataflow counters say they are current, and stores them afterwards
B200_D void solveJoint(int kind, const JointSolveArgs& J, uint32_t i, BodyVel& A, BodyVel& B, const float4* __restrict__ gInvI)
{
    switch (kind)
    {
    case 0: // constraints.cpp:242-264
    {
        const DistanceRowD c = ((const DistanceRowD*)J.rows[0])[i];
        const f3 avA = A.v + cross(A.w, c.rA), avB = B.v + cross(B.w, c.rB);
        const float Cdot = dot(c.u, avB - avA) + c.bias;
        const float lambda = -c.mass * Cdot;
        const f3 P = lambda * c.u;
        A.v -= A.invMass * P; A.w -= c.angA * lambda; B.v += B.invMass * P; B.w += c.angB * lambda;
    } break;
    case 1: // constraints.cpp:505-528
    {
        const BallRowD c = ((const BallRowD*)J.rows[1])[i];
        const m3 IA = loadInvInertiaJ(gInvI, c.a), IB = loadInvInertiaJ(gInvI, c.b);
        const f3 avA = A.v + cross(A.w, c.rA), avB = B.v + cross(B.w, c.rB);
        const f3 Cdot = avB - avA + c.bias;
        const f3 P = solve3(c.invMass, -Cdot);
        A.v -= A.invMass * P; A.w -= IA * cross(c.rA, P);
        B.v += B.invMass * P; B.w += IB * cross(c.rB, P);
    } break;
    case 2: // constraints.cpp:788-823
    {
        const FixedRowD c = ((const FixedRowD*)J.rows[2])[i];
        const m3 IA = loadInvInertiaJ(gInvI, c.a), IB = loadInvInertiaJ(gInvI, c.b);
        {
            const f3 Cdot = B.w - A.w;
            const f3 l = solve3(c.invRMass, -(Cdot + c.rBias));
            A.w -= IA * l; B.w += IB * l;
        }
        {
            const f3 avA = A.v + cross(A.w, c.rA), avB = B.v + cross(B.w, c.rB);
            const f3 Cdot = avB - avA + c.tBias;
            const f3 P = solve3(c.invTMass, -Cdot);
            A.v -= A.invMass * P; A.w -= IA * cross(c.rA, P);
            B.v += B.invMass * P; B.w += IB * cross(c.rB, P);
        }
    } break;
    case 3: // constraints.cpp:1215-1307 : motor -> limit -> rotation -> position
    {
        HingeRowD* row = ((HingeRowD*)J.rows[3]) + i;
        const HingeRowD c = *row;
        const m3 IA = loadInvInertiaJ(gInvI, c.a), IB = loadInvInertiaJ(gInvI, c.b);
        f3 vA = A.v, wA = A.w, vB = B.v, wB = B.w;
        const f3 axis = c.axis;
        if (c.solveMotor)
        {
            const float rel = dot(axis, wB) - dot(axis, wA);
            const float Cdot = rel - c.motorVelocity;
            float l = -c.axialMass * Cdot;
            const float old = c.motorImpulse;
            const float ni = clampf(old + l, -c.maxMotorImpulse, c.maxMotorImpulse);
            l = ni - old; row->motorImpulse = ni;
            wA -= c.angA * l; wB += c.angB * l;
        }
        if (c.solveLimit)
        {
            const float s = c.limitSign;
            const float aA = dot(axis, wA), aB = dot(axis, wB);
            const float rel = s * (aB - aA);
            const float Cdot = rel + c.limitBias;
            float l = -c.axialMass * Cdot;
            const float imp = fmax2(c.limitImpulse + l, 0.f);
            l = imp - c.limitImpulse; row->limitImpulse = imp;
            l *= s;
            wA -= c.angA * l; wB += c.angB * l;
        }
        {
            const f3 dw = wB - wA;
            const float cx = dot(c.bxa, dw), cy = dot(c.cxa, dw);
            float lx, ly; solve2(c.invRMass, -(cx + c.rBiasX), -(cy + c.rBiasY), lx, ly);
            const f3 P = c.bxa * lx + c.cxa * ly;
            wA -= IA * P; wB += IB * P;
        }
        {
            const f3 avA = vA + cross(wA, c.rA), avB = vB + cross(wB, c.rB);
            const f3 Cdot = avB - avA + c.tBias;
            const f3 P = solve3(c.invTMass, -Cdot);
            vA -= A.invMass * P; wA -= IA * cross(c.rA, P);
            vB += B.invMass * P; wB += IB * cross(c.rB, P);
        }
        A.v = vA; A.w = wA; B.v = vB; B.w = wB;
    } break;
    case 4: // constraints.cpp:1955-2070 : twist motor -> swing motor -> twist limit -> swing limit -> position
    {
        ConeTwistRowD* row = ((ConeTwistRowD*)J.rows[4]) + i;
        const ConeTwistRowD c = *row;
        const m3 IA = loadInvInertiaJ(gInvI, c.a), IB = loadInvInertiaJ(gInvI, c.b);
        f3 vA = A.v, wA = A.w, vB = B.v, wB = B.w;
        const f3 tw = c.twistAxis;
        if (c.solveTwistMotor)
        {
            const float rel = dot(tw, wB) - dot(tw, wA);
            const float Cdot = rel - c.twistMotorVelocity;
            float l = -c.twistMass * Cdot;
            const float old = c.twistMotorImpulse;
            const float ni = clampf(old + l, -c.maxTwistMotorImpulse, c.maxTwistMotorImpulse);
            l = ni - old; row->twistMotorImpulse = ni;
            wA -= c.twistAngA * l; wB += c.twistAngB * l;
        }
        if (c.solveSwingMotor)
        {
            const f3 ax = c.swingMotorAxis;
            const float rel = dot(ax, wB) - dot(ax, wA);
            const float Cdot = rel - c.swingMotorVelocity;
            float l = -c.swingMotorMass * Cdot;
            const float old = c.swingMotorImpulse;
            const float ni = clampf(old + l, -c.maxSwingMotorImpulse, c.maxSwingMotorImpulse);
            l = ni - old; row->swingMotorImpulse = ni;
            wA -= c.swingMotorAngA * l; wB += c.swingMotorAngB * l;
        }
        if (c.solveTwistLimit)
        {
            const float s = c.twistLimitSign;
            const float aA = dot(tw, wA), aB = dot(tw, wB);
            const float rel = s * (aB - aA);
            const float Cdot = rel + c.twistLimitBias;
            float l = -c.twistMass * Cdot;
            const float imp = fmax2(c.twistImpulse + l, 0.f);
            l = imp - c.twistImpulse; row->twistImpulse = imp;
            l *= s;
            wA -= c.twistAngA * l; wB += c.twistAngB * l;
        }
        if (c.solveSwingLimit)
        {
            const float aA = dot(c.swingAxis, wA), aB = dot(c.swingAxis, wB);
            const float Cdot = aA - aB + c.swingLimitBias;
            float l = -c.swingLimitMass * Cdot;
            const float imp = fmax2(c.swingImpulse + l, 0.f);
            l = imp - c.swingImpulse; row->swingImpulse = imp;
            wA += c.swingLimitAngA * l; wB -= c.swingLimitAngB * l;
        }
        {
            const f3 avA = vA + cross(wA, c.rA), avB = vB + cross(wB, c.rB);
            const f3 Cdot = avB - avA + c.bias;
            const f3 P = solve3(c.invMass, -Cdot);
            vA -= A.invMass * P; wA -= IA * cross(c.rA, P);
            vB += B.invMass * P; wB += IB * cross(c.rB, P);
        }
        A.v = vA; A.w = wA; B.v = vB; B.w = wB;
    } break;
    default: // slider, constraints.cpp:2763-2846 : motor -> limit -> rotation -> position
    {
        SliderRowD* row = ((SliderRowD*)J.rows[5]) + i;
        const SliderRowD c = *row;
        const m3 IA = loadInvInertiaJ(gInvI, c.a), IB = loadInvInertiaJ(gInvI, c.b);
        f3 vA = A.v, wA = A.w, vB = B.v, wB = B.w;
        if (c.solveMotor)
        {
            const float Cdot = dot(vB, c.axis) - dot(vA, c.axis) - c.motorVelocity;
            const float mass = 1.f / (A.invMass + B.invMass);
            float l = -mass * Cdot;
            const float old = c.motorImpulse;
            const float ni = clampf(old + l, -c.maxMotorImpulse, c.maxMotorImpulse);
            l = ni - old; row->motorImpulse = ni;
            const f3 P = l * c.axis;
            vA -= A.invMass * P; vB += B.invMass * P;
        }
        if (c.solveLimit)
        {
            const float Cdot = dot(vB, c.axis) + dot(wB, c.rBxs) - dot(vA, c.axis) - dot(wA, c.rAuxs);
            float l = -c.axialMass * (c.limitSign * Cdot + c.limitBias);
            const float imp = fmax2(c.limitImpulse + l, 0.f);
            l = imp - c.limitImpulse; row->limitImpulse = imp;
            l *= c.limitSign;
            const f3 P = l * c.axis;
            vA -= A.invMass * P; wA -= c.limAngA * l; vB += B.invMass * P; wB += c.limAngB * l;
        }
        {
            const f3 Cdot = wB - wA;
            const f3 l = solve3(c.invRMass, -(Cdot + c.rBias));
            wA -= IA * l; wB += IB * l;
        }
        {
            const float cx = dot(c.tangent, vB) + dot(c.rBxt, wB) - dot(c.tangent, vA) - dot(c.rAuxt, wA);
            const float cy = dot(c.bitangent, vB) + dot(c.rBxb, wB) - dot(c.bitangent, vA) - dot(c.rAuxb, wA);
            float lx, ly; solve2(c.invTMass, -(cx + c.tBiasX), -(cy + c.tBiasY), lx, ly);
            const f3 tb = c.tangent * lx + c.bitangent * ly;
            vA -= A.invMass * tb; wA -= IA * (c.rAuxt * lx + c.rAuxb * ly);
            vB += B.invMass * tb; wB += IB * (c.rBxt * lx + c.rBxb * ly);
        }
        A.v = vA; A.w = wA; B.v = vB; B.w = wB;
    } break;
    }
}

} // namespace b200
