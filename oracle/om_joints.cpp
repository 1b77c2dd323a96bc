// oracle/om_joints.cpp -- TEST INFRASTRUCTURE ONLY (parity oracle). Never linked into the product path.
//
// CPU restatement of the reference's SCALAR joint constraints (src/physics/constraints.cpp):
//   distance :189-264   ball :460-528   fixed :736-823   hinge :1079-1307   cone-twist :1782-2070   slider :2638-2846
// Each kind: initialise once per step from the body global states, then solve Gauss-Seidel in array order (or in the
// order handed in, the analogue of the reference's own AVX2 re-batching, constraints.cpp:51-184).
#include "om_step.h"

namespace orc {

static inline V3 ld3(const float* p) { return v3(p[0], p[1], p[2]); }
static inline Q4 ld4(const float* p) { return q4(p[0], p[1], p[2], p[3]); }

static const float DISTANCE_BETA = 0.1f, BALL_BETA = 0.1f, SLIDER_BETA = 0.1f, HINGE_ROT_BETA = 0.3f,
                   HINGE_LIMIT_BETA = 0.1f, TWIST_LIMIT_BETA = 0.1f, SLIDER_LIMIT_BETA = 0.1f, DT_THRESHOLD = 1e-5f; // constraints.cpp:9-17

struct DistanceRow { uint32_t a, b; V3 rA, rB, angA, angB, u; float bias, mass; };
struct BallRow { uint32_t a, b; V3 rA, rB, bias; M3 invMass; };
struct FixedRow { uint32_t a, b; V3 rA, rB, tBias; M3 invTMass; V3 rBias; M3 invRMass; };
struct HingeRow
{
    uint32_t a, b; V3 rA, rB, tBias; M3 invTMass; float rBiasX, rBiasY; M2 invRMass; V3 bxa, cxa;
    V3 axis; float axialMass; bool solveLimit, solveMotor;
    float limitImpulse, limitBias, limitSign, motorImpulse, maxMotorImpulse, motorVelocity; V3 angA, angB;
};
struct ConeTwistRow
{
    uint32_t a, b; V3 rA, rB, bias; M3 invMass;
    bool solveSwingLimit, solveTwistLimit, solveSwingMotor, solveTwistMotor;
    V3 swingAxis; float swingImpulse, swingLimitMass, swingLimitBias;
    V3 twistAxis; float twistImpulse, twistLimitSign, twistMass, twistLimitBias;
    float swingMotorImpulse, maxSwingMotorImpulse, swingMotorVelocity; V3 swingMotorAxis; float swingMotorMass;
    float twistMotorImpulse, maxTwistMotorImpulse, twistMotorVelocity;
    V3 twistAngA, twistAngB, swingMotorAngA, swingMotorAngB, swingLimitAngA, swingLimitAngB;
};
struct SliderRow
{
    uint32_t a, b; V3 rAuxt, rAuxb, rBxt, rBxb, tangent, bitangent; M2 invTMass; float tBiasX, tBiasY; M3 invRMass; V3 rBias;
    bool solveLimit, solveMotor; V3 axis; float axialMass, limitBias, limitImpulse, limitSign; V3 rAuxs, rBxs, limAngA, limAngB;
    float motorVelocity, motorImpulse, maxMotorImpulse;
};

// shared "ball" part: skew(rA) I_A skew(rA)^T + skew(rB) I_B skew(rB)^T + (mA + mB) Id   (constraints.cpp:487-492)
static M3 pointMassMatrix(const BodyG& A, const BodyG& B, V3 rA, V3 rB)
{
    M3 sA = skew(rA), sB = skew(rB);
    return sA * A.invInertia * transpose(sA) + sB * B.invInertia * transpose(sB) + m3Identity() * (A.invMass + B.invMass);
}

template <typename T> static T* rowsOf(std::vector<uint8_t>& v, size_t n) { v.resize(n * sizeof(T)); return (T*)v.data(); }

void Oracle::initJoints(float dt)
{
    float invDt = 1.f / dt;
    // ---- distance (constraints.cpp:189-239)
    {
        size_t n = jointPairs[0].size(); auto* in = (const b200phys_distance_joint*)joints[0].data(); auto* out = rowsOf<DistanceRow>(jointRows[0], n);
        for (size_t i = 0; i < n; ++i)
        {
            DistanceRow& o = out[i]; o.a = jointPairs[0][i].rbA; o.b = jointPairs[0][i].rbB;
            const BodyG& A = g[o.a]; const BodyG& B = g[o.b];
            o.rA = A.rotation * (ld3(in[i].localAnchorA) - A.localCOG); o.rB = B.rotation * (ld3(in[i].localAnchorB) - B.localCOG);
            V3 gA = A.position + o.rA, gB = B.position + o.rB;
            o.u = gB - gA; float l = len(o.u);
            o.u = (l > 0.001f) ? (o.u * (1.f / l)) : v3(0.f);
            V3 crAu = cross(o.rA, o.u), crBu = cross(o.rB, o.u);
            float inv = A.invMass + dot(crAu, A.invInertia * crAu) + B.invMass + dot(crBu, B.invInertia * crBu);
            o.mass = (inv != 0.f) ? (1.f / inv) : 0.f;
            o.bias = 0.f;
            if (dt > DT_THRESHOLD) o.bias = (l - in[i].globalLength) * (DISTANCE_BETA * invDt);
            o.angA = A.invInertia * cross(o.rA, crAu);   // sic: reference crosses the lever arm with r x u (:232-233)
            o.angB = B.invInertia * cross(o.rB, crBu);
        }
    }
    // ---- ball (constraints.cpp:460-503)
    {
        size_t n = jointPairs[1].size(); auto* in = (const b200phys_ball_joint*)joints[1].data(); auto* out = rowsOf<BallRow>(jointRows[1], n);
        for (size_t i = 0; i < n; ++i)
        {
            BallRow& o = out[i]; o.a = jointPairs[1][i].rbA; o.b = jointPairs[1][i].rbB;
            const BodyG& A = g[o.a]; const BodyG& B = g[o.b];
            o.rA = A.rotation * (ld3(in[i].localAnchorA) - A.localCOG); o.rB = B.rotation * (ld3(in[i].localAnchorB) - B.localCOG);
            V3 gA = A.position + o.rA, gB = B.position + o.rB;
            o.invMass = pointMassMatrix(A, B, o.rA, o.rB);
            o.bias = v3(0.f);
            if (dt > DT_THRESHOLD) o.bias = (gB - gA) * (BALL_BETA * invDt);
        }
    }
    // ---- fixed (constraints.cpp:736-786)
    {
        size_t n = jointPairs[2].size(); auto* in = (const b200phys_fixed_joint*)joints[2].data(); auto* out = rowsOf<FixedRow>(jointRows[2], n);
        for (size_t i = 0; i < n; ++i)
        {
            FixedRow& o = out[i]; o.a = jointPairs[2][i].rbA; o.b = jointPairs[2][i].rbB;
            const BodyG& A = g[o.a]; const BodyG& B = g[o.b];
            o.rA = A.rotation * (ld3(in[i].localAnchorA) - A.localCOG); o.rB = B.rotation * (ld3(in[i].localAnchorB) - B.localCOG);
            V3 gA = A.position + o.rA, gB = B.position + o.rB;
            o.invTMass = pointMassMatrix(A, B, o.rA, o.rB);
            o.invRMass = A.invInertia + B.invInertia;
            o.tBias = v3(0.f); o.rBias = v3(0.f);
            if (dt > DT_THRESHOLD)
            {
                o.tBias = (gB - gA) * (BALL_BETA * invDt);
                Q4 err = B.rotation * ld4(in[i].initialInvRotationDifference) * conjugate(A.rotation);
                o.rBias = qv(err) * (SLIDER_BETA * invDt * 2.f);
            }
        }
    }
    // ---- hinge (constraints.cpp:1079-1213)
    {
        size_t n = jointPairs[3].size(); auto* in = (const b200phys_hinge_joint*)joints[3].data(); auto* out = rowsOf<HingeRow>(jointRows[3], n);
        for (size_t i = 0; i < n; ++i)
        {
            const b200phys_hinge_joint& J = in[i];
            HingeRow& o = out[i]; o.a = jointPairs[3][i].rbA; o.b = jointPairs[3][i].rbB;
            const BodyG& A = g[o.a]; const BodyG& B = g[o.b];
            o.rA = A.rotation * (ld3(J.localAnchorA) - A.localCOG); o.rB = B.rotation * (ld3(J.localAnchorB) - B.localCOG);
            V3 gA = A.position + o.rA, gB = B.position + o.rB;
            o.invTMass = pointMassMatrix(A, B, o.rA, o.rB);
            o.tBias = v3(0.f);
            if (dt > DT_THRESHOLD) o.tBias = (gB - gA) * (BALL_BETA * invDt);

            V3 axA = A.rotation * ld3(J.localHingeAxisA), axB = B.rotation * ld3(J.localHingeAxisB);
            V3 tB, bB; getTangents(axB, tB, bB);
            V3 bxa = cross(tB, axA), cxa = cross(bB, axA);
            V3 iAb = A.invInertia * bxa, iBb = B.invInertia * bxa, iAc = A.invInertia * cxa, iBc = B.invInertia * cxa;
            o.invRMass.m00 = dot(bxa, iAb) + dot(bxa, iBb);
            o.invRMass.m01 = dot(bxa, iAc) + dot(bxa, iBc);
            o.invRMass.m10 = dot(cxa, iAb) + dot(cxa, iBb);
            o.invRMass.m11 = dot(cxa, iAc) + dot(cxa, iBc);
            o.bxa = bxa; o.cxa = cxa;
            o.rBiasX = 0.f; o.rBiasY = 0.f;
            if (dt > DT_THRESHOLD) { float k = HINGE_ROT_BETA * invDt; o.rBiasX = dot(axA, tB) * k; o.rBiasY = dot(axA, bB) * k; }

            o.solveLimit = false; o.solveMotor = false; o.axis = v3(0.f);
            o.limitImpulse = o.limitBias = o.limitSign = o.motorImpulse = o.maxMotorImpulse = o.motorVelocity = o.axialMass = 0.f;
            o.angA = o.angB = v3(0.f);
            if (J.minRotationLimit <= 0.f || J.maxRotationLimit >= 0.f || J.maxMotorTorque > 0.f)
            {
                V3 cmp = conjugate(A.rotation) * (B.rotation * ld3(J.localHingeTangentB));
                float angle = atan2f(dot(cmp, ld3(J.localHingeBitangentA)), dot(cmp, ld3(J.localHingeTangentA)));
                bool minV = J.minRotationLimit <= 0.f && angle <= J.minRotationLimit;
                bool maxV = J.maxRotationLimit >= 0.f && angle >= J.maxRotationLimit;
                if (minV && maxV) ++nstats.asserts;
                o.solveLimit = minV || maxV;
                o.solveMotor = J.maxMotorTorque > 0.f;
                if (o.solveLimit || o.solveMotor)
                {
                    o.axis = axA; o.limitImpulse = 0.f;
                    float inv = dot(axA, A.invInertia * axA) + dot(axA, B.invInertia * axA);
                    o.axialMass = (inv != 0.f) ? (1.f / inv) : 0.f;
                    o.limitSign = minV ? 1.f : -1.f;
                    o.maxMotorImpulse = J.maxMotorTorque * dt; o.motorImpulse = 0.f;
                    o.angA = A.invInertia * o.axis; o.angB = B.invInertia * o.axis;
                    o.motorVelocity = J.motorVelocityOrTargetAngle;
                    if (J.motorType == B200PHYS_MOTOR_POSITION)
                    {
                        float minL = (J.minRotationLimit <= 0.f) ? J.minRotationLimit : -kPi;
                        float maxL = (J.maxRotationLimit >= 0.f) ? J.maxRotationLimit : kPi;
                        float target = clampf(J.motorVelocityOrTargetAngle, minL, maxL);
                        o.motorVelocity = (dt > DT_THRESHOLD) ? ((target - angle) * invDt) : 0.f;
                    }
                    o.limitBias = 0.f;
                    if (dt > DT_THRESHOLD)
                    {
                        float d = minV ? (angle - J.minRotationLimit) : (J.maxRotationLimit - angle);
                        o.limitBias = d * HINGE_LIMIT_BETA * invDt;
                    }
                }
            }
        }
    }
    // ---- cone-twist (constraints.cpp:1782-1953)
    {
        size_t n = jointPairs[4].size(); auto* in = (const b200phys_cone_twist_joint*)joints[4].data(); auto* out = rowsOf<ConeTwistRow>(jointRows[4], n);
        for (size_t i = 0; i < n; ++i)
        {
            const b200phys_cone_twist_joint& J = in[i];
            ConeTwistRow& o = out[i]; memset(&o, 0, sizeof(o));
            o.a = jointPairs[4][i].rbA; o.b = jointPairs[4][i].rbB;
            const BodyG& A = g[o.a]; const BodyG& B = g[o.b];
            o.rA = A.rotation * (ld3(J.localAnchorA) - A.localCOG); o.rB = B.rotation * (ld3(J.localAnchorB) - B.localCOG);
            V3 gA = A.position + o.rA, gB = B.position + o.rB;
            o.invMass = pointMassMatrix(A, B, o.rA, o.rB);
            o.bias = v3(0.f);
            if (dt > DT_THRESHOLD) o.bias = (gB - gA) * (BALL_BETA * invDt);

            Q4 btoa = conjugate(A.rotation) * B.rotation;
            V3 limitAxisA = ld3(J.localLimitAxisA);
            V3 limitAxisCmpA = btoa * ld3(J.localLimitAxisB);
            Q4 swingRot = rotateFromTo(limitAxisA, limitAxisCmpA);
            V3 twistTanA = swingRot * ld3(J.localLimitTangentA), twistBitanA = swingRot * ld3(J.localLimitBitangentA);
            V3 tanCmpA = btoa * ld3(J.localLimitTangentB);
            float twistAngle = atan2f(dot(tanCmpA, twistBitanA), dot(tanCmpA, twistTanA));

            V3 swingAxis; float swingAngle; getAxisRotation(swingRot, swingAxis, swingAngle);
            if (swingAngle < 0.f) { swingAngle *= -1.f; swingAxis *= -1.f; }

            o.solveSwingLimit = J.swingLimit >= 0.f && swingAngle >= J.swingLimit;
            if (o.solveSwingLimit)
            {
                o.swingImpulse = 0.f;
                o.swingAxis = A.rotation * swingAxis;
                float inv = dot(o.swingAxis, A.invInertia * o.swingAxis) + dot(o.swingAxis, B.invInertia * o.swingAxis);
                o.swingLimitMass = (inv != 0.f) ? (1.f / inv) : 0.f;
                o.swingLimitBias = 0.f;
                if (dt > DT_THRESHOLD) o.swingLimitBias = (J.swingLimit - swingAngle) * (HINGE_LIMIT_BETA * invDt);
                o.swingLimitAngA = A.invInertia * o.swingAxis; o.swingLimitAngB = B.invInertia * o.swingAxis;
            }
            o.solveSwingMotor = J.maxSwingMotorTorque > 0.f;
            if (o.solveSwingMotor)
            {
                o.maxSwingMotorImpulse = J.maxSwingMotorTorque * dt; o.swingMotorImpulse = 0.f;
                float axX = cosf(J.swingMotorAxis), axY = sinf(J.swingMotorAxis);
                V3 localMotorAxis = axX * ld3(J.localLimitTangentA) + axY * ld3(J.localLimitBitangentA);
                if (J.swingMotorType == B200PHYS_MOTOR_VELOCITY)
                {
                    o.swingMotorAxis = A.rotation * localMotorAxis;
                    o.swingMotorVelocity = J.swingMotorVelocityOrTargetAngle;
                }
                else
                {
                    float target = J.swingMotorVelocityOrTargetAngle;
                    if (J.swingLimit >= 0.f) target = clampf(target, -J.swingLimit, J.swingLimit);
                    V3 targetDir = quatAxisAngle(localMotorAxis, target) * limitAxisA;
                    V3 motorAxis2 = noz(cross(limitAxisCmpA, targetDir));
                    o.swingMotorAxis = A.rotation * motorAxis2;
                    float cosAngle = dot(targetDir, limitAxisCmpA);
                    float delta = acosf(clamp01(cosAngle));
                    o.swingMotorVelocity = (dt > DT_THRESHOLD) ? (delta * invDt * 0.2f) : 0.f;
                }
                o.swingMotorAngA = A.invInertia * o.swingMotorAxis; o.swingMotorAngB = B.invInertia * o.swingMotorAxis;
                float inv = dot(o.swingMotorAxis, A.invInertia * o.swingMotorAxis) + dot(o.swingMotorAxis, B.invInertia * o.swingMotorAxis);
                o.swingMotorMass = (inv != 0.f) ? (1.f / inv) : 0.f;
            }
            bool minTw = J.twistLimit >= 0.f && twistAngle <= -J.twistLimit;
            bool maxTw = J.twistLimit >= 0.f && twistAngle >= J.twistLimit;
            if (minTw && maxTw) ++nstats.asserts;
            o.solveTwistLimit = minTw || maxTw;
            o.solveTwistMotor = J.maxTwistMotorTorque > 0.f;
            if (o.solveTwistLimit || o.solveTwistMotor)
            {
                o.twistImpulse = 0.f;
                o.twistAxis = A.rotation * limitAxisA;
                float inv = dot(o.twistAxis, A.invInertia * o.twistAxis) + dot(o.twistAxis, B.invInertia * o.twistAxis);
                o.twistMass = (inv != 0.f) ? (1.f / inv) : 0.f;
                o.twistLimitSign = minTw ? 1.f : -1.f;
                o.maxTwistMotorImpulse = J.maxTwistMotorTorque * dt; o.twistMotorImpulse = 0.f;
                o.twistAngA = A.invInertia * o.twistAxis; o.twistAngB = B.invInertia * o.twistAxis;
                o.twistMotorVelocity = J.twistMotorVelocityOrTargetAngle;
                if (J.twistMotorType == B200PHYS_MOTOR_POSITION)
                {
                    float limit = (J.twistLimit >= 0.f) ? J.twistLimit : kPi;
                    float target = clampf(J.twistMotorVelocityOrTargetAngle, -limit, limit);
                    o.twistMotorVelocity = (dt > DT_THRESHOLD) ? ((target - twistAngle) * invDt) : 0.f;
                }
                o.twistLimitBias = 0.f;
                if (dt > DT_THRESHOLD)
                {
                    float d = minTw ? (J.twistLimit + twistAngle) : (J.twistLimit - twistAngle);
                    o.twistLimitBias = d * TWIST_LIMIT_BETA * invDt;
                }
            }
        }
    }
    // ---- slider (constraints.cpp:2638-2761)
    {
        size_t n = jointPairs[5].size(); auto* in = (const b200phys_slider_joint*)joints[5].data(); auto* out = rowsOf<SliderRow>(jointRows[5], n);
        for (size_t i = 0; i < n; ++i)
        {
            const b200phys_slider_joint& J = in[i];
            SliderRow& o = out[i]; memset(&o, 0, sizeof(o));
            o.a = jointPairs[5][i].rbA; o.b = jointPairs[5][i].rbB;
            const BodyG& A = g[o.a]; const BodyG& B = g[o.b];
            V3 rA = A.rotation * (ld3(J.localAnchorA) - A.localCOG), rB = B.rotation * (ld3(J.localAnchorB) - B.localCOG);
            V3 gA = A.position + rA, gB = B.position + rB;
            V3 axis = A.rotation * ld3(J.localAxisA);
            getTangents(axis, o.tangent, o.bitangent);
            V3 u = gB - gA;
            V3 rAu = rA + u;
            o.rBxt = cross(rB, o.tangent); o.rBxb = cross(rB, o.bitangent);
            o.rAuxt = cross(rAu, o.tangent); o.rAuxb = cross(rAu, o.bitangent);
            V3 iAt = A.invInertia * o.rAuxt, iAb = A.invInertia * o.rAuxb, iBt = B.invInertia * o.rBxt, iBb = B.invInertia * o.rBxb;
            float ms = A.invMass + B.invMass;
            o.invTMass.m00 = dot(o.rAuxt, iAt) + dot(o.rBxt, iBt) + ms;
            o.invTMass.m01 = dot(o.rAuxt, iAb) + dot(o.rBxt, iBb);
            o.invTMass.m10 = dot(o.rAuxb, iAt) + dot(o.rBxb, iBt);
            o.invTMass.m11 = dot(o.rAuxb, iAb) + dot(o.rBxb, iBb) + ms;
            o.invRMass = A.invInertia + B.invInertia;
            o.tBiasX = o.tBiasY = 0.f; o.rBias = v3(0.f);
            if (dt > DT_THRESHOLD)
            {
                float k = SLIDER_BETA * invDt;
                o.tBiasX = dot(u, o.tangent) * k; o.tBiasY = dot(u, o.bitangent) * k;
                Q4 err = B.rotation * ld4(J.initialInvRotationDifference) * conjugate(A.rotation);
                o.rBias = qv(err) * (SLIDER_BETA * invDt * 2.f);
            }
            o.axis = axis;
            float dist = dot(u, axis);
            o.solveLimit = false;
            if (J.negDistanceLimit <= 0.f || J.posDistanceLimit >= 0.f)
            {
                bool minV = (J.negDistanceLimit <= 0.f) && (dist < J.negDistanceLimit);
                bool maxV = (J.posDistanceLimit >= 0.f) && (dist > J.posDistanceLimit);
                if (minV && maxV) ++nstats.asserts;
                if (minV || maxV)
                {
                    o.solveLimit = true; o.limitImpulse = 0.f;
                    o.rAuxs = cross(rAu, axis); o.rBxs = cross(rB, axis);
                    float inv = ms + dot(o.rAuxs, A.invInertia * o.rAuxs) + dot(o.rBxs, B.invInertia * o.rBxs);
                    o.axialMass = (inv != 0.f) ? (1.f / inv) : 0.f;
                    o.limitSign = minV ? 1.f : -1.f;
                    o.limitBias = 0.f;
                    if (dt > DT_THRESHOLD)
                    {
                        float e = minV ? (dist - J.negDistanceLimit) : (J.posDistanceLimit - dist);
                        o.limitBias = e * (SLIDER_LIMIT_BETA * invDt);
                    }
                    o.limAngA = A.invInertia * o.rAuxs; o.limAngB = B.invInertia * o.rBxs;
                }
            }
            o.solveMotor = false;
            if (J.maxMotorForce > 0.f)
            {
                o.solveMotor = true; o.maxMotorImpulse = J.maxMotorForce * dt; o.motorImpulse = 0.f;
                o.motorVelocity = J.motorVelocityOrTargetDistance;
                if (J.motorType == B200PHYS_MOTOR_POSITION)
                {
                    float minL = (J.negDistanceLimit <= 0.f) ? J.negDistanceLimit : -INFINITY;
                    float maxL = (J.posDistanceLimit >= 0.f) ? J.posDistanceLimit : INFINITY;
                    float target = clampf(J.motorVelocityOrTargetDistance, minL, maxL);
                    o.motorVelocity = (dt > DT_THRESHOLD) ? ((target - dist) * invDt) : 0.f;
                }
            }
        }
    }
}

void Oracle::solveJoints(int kind, const uint32_t* order)
{
    size_t n = jointPairs[kind].size();
    for (size_t k = 0; k < n; ++k)
    {
        size_t i = order ? order[k] : k;
        switch (kind)
        {
        case 0: // constraints.cpp:242-264
        {
            DistanceRow& c = ((DistanceRow*)jointRows[0].data())[i]; BodyG& A = g[c.a]; BodyG& B = g[c.b];
            V3 avA = A.v + cross(A.w, c.rA), avB = B.v + cross(B.w, c.rB);
            float Cdot = dot(c.u, avB - avA) + c.bias;
            float lambda = -c.mass * Cdot;
            V3 P = lambda * c.u;
            A.v -= A.invMass * P; A.w -= c.angA * lambda; B.v += B.invMass * P; B.w += c.angB * lambda;
        } break;
        case 1: // constraints.cpp:505-528
        {
            BallRow& c = ((BallRow*)jointRows[1].data())[i]; BodyG& A = g[c.a]; BodyG& B = g[c.b];
            V3 avA = A.v + cross(A.w, c.rA), avB = B.v + cross(B.w, c.rB);
            V3 Cdot = avB - avA + c.bias;
            V3 P = solve3(c.invMass, -Cdot);
            A.v -= A.invMass * P; A.w -= A.invInertia * cross(c.rA, P);
            B.v += B.invMass * P; B.w += B.invInertia * cross(c.rB, P);
        } break;
        case 2: // constraints.cpp:788-823
        {
            FixedRow& c = ((FixedRow*)jointRows[2].data())[i]; BodyG& A = g[c.a]; BodyG& B = g[c.b];
            {
                V3 Cdot = B.w - A.w;
                V3 l = solve3(c.invRMass, -(Cdot + c.rBias));
                A.w -= A.invInertia * l; B.w += B.invInertia * l;
            }
            {
                V3 avA = A.v + cross(A.w, c.rA), avB = B.v + cross(B.w, c.rB);
                V3 Cdot = avB - avA + c.tBias;
                V3 P = solve3(c.invTMass, -Cdot);
                A.v -= A.invMass * P; A.w -= A.invInertia * cross(c.rA, P);
                B.v += B.invMass * P; B.w += B.invInertia * cross(c.rB, P);
            }
        } break;
        case 3: // constraints.cpp:1215-1307 : motor -> limit -> rotation -> position
        {
            HingeRow& c = ((HingeRow*)jointRows[3].data())[i]; BodyG& A = g[c.a]; BodyG& B = g[c.b];
            V3 vA = A.v, wA = A.w, vB = B.v, wB = B.w;
            V3 axis = c.axis;
            if (c.solveMotor)
            {
                float rel = dot(axis, wB) - dot(axis, wA);
                float Cdot = rel - c.motorVelocity;
                float l = -c.axialMass * Cdot;
                float old = c.motorImpulse;
                c.motorImpulse = clampf(c.motorImpulse + l, -c.maxMotorImpulse, c.maxMotorImpulse);
                l = c.motorImpulse - old;
                wA -= c.angA * l; wB += c.angB * l;
            }
            if (c.solveLimit)
            {
                float s = c.limitSign;
                float aA = dot(axis, wA), aB = dot(axis, wB);
                float rel = s * (aB - aA);
                float Cdot = rel + c.limitBias;
                float l = -c.axialMass * Cdot;
                float imp = fmax2(c.limitImpulse + l, 0.f);
                l = imp - c.limitImpulse; c.limitImpulse = imp;
                l *= s;
                wA -= c.angA * l; wB += c.angB * l;
            }
            {
                V3 dw = wB - wA;
                float cx = dot(c.bxa, dw), cy = dot(c.cxa, dw);
                float lx, ly; solve2(c.invRMass, -(cx + c.rBiasX), -(cy + c.rBiasY), lx, ly);
                V3 P = c.bxa * lx + c.cxa * ly;
                wA -= A.invInertia * P; wB += B.invInertia * P;
            }
            {
                V3 avA = vA + cross(wA, c.rA), avB = vB + cross(wB, c.rB);
                V3 Cdot = avB - avA + c.tBias;
                V3 P = solve3(c.invTMass, -Cdot);
                vA -= A.invMass * P; wA -= A.invInertia * cross(c.rA, P);
                vB += B.invMass * P; wB += B.invInertia * cross(c.rB, P);
            }
            A.v = vA; A.w = wA; B.v = vB; B.w = wB;
        } break;
        case 4: // constraints.cpp:1955-2070 : twist motor -> swing motor -> twist limit -> swing limit -> position
        {
            ConeTwistRow& c = ((ConeTwistRow*)jointRows[4].data())[i]; BodyG& A = g[c.a]; BodyG& B = g[c.b];
            V3 vA = A.v, wA = A.w, vB = B.v, wB = B.w;
            V3 tw = c.twistAxis;
            if (c.solveTwistMotor)
            {
                float rel = dot(tw, wB) - dot(tw, wA);
                float Cdot = rel - c.twistMotorVelocity;
                float l = -c.twistMass * Cdot;
                float old = c.twistMotorImpulse;
                c.twistMotorImpulse = clampf(c.twistMotorImpulse + l, -c.maxTwistMotorImpulse, c.maxTwistMotorImpulse);
                l = c.twistMotorImpulse - old;
                wA -= c.twistAngA * l; wB += c.twistAngB * l;
            }
            if (c.solveSwingMotor)
            {
                V3 ax = c.swingMotorAxis;
                float rel = dot(ax, wB) - dot(ax, wA);
                float Cdot = rel - c.swingMotorVelocity;
                float l = -c.swingMotorMass * Cdot;
                float old = c.swingMotorImpulse;
                c.swingMotorImpulse = clampf(c.swingMotorImpulse + l, -c.maxSwingMotorImpulse, c.maxSwingMotorImpulse);
                l = c.swingMotorImpulse - old;
                wA -= c.swingMotorAngA * l; wB += c.swingMotorAngB * l;
            }
            if (c.solveTwistLimit)
            {
                float s = c.twistLimitSign;
                float aA = dot(tw, wA), aB = dot(tw, wB);
                float rel = s * (aB - aA);
                float Cdot = rel + c.twistLimitBias;
                float l = -c.twistMass * Cdot;
                float imp = fmax2(c.twistImpulse + l, 0.f);
                l = imp - c.twistImpulse; c.twistImpulse = imp;
                l *= s;
                wA -= c.twistAngA * l; wB += c.twistAngB * l;
            }
            if (c.solveSwingLimit)
            {
                float aA = dot(c.swingAxis, wA), aB = dot(c.swingAxis, wB);
                float Cdot = aA - aB + c.swingLimitBias;
                float l = -c.swingLimitMass * Cdot;
                float imp = fmax2(c.swingImpulse + l, 0.f);
                l = imp - c.swingImpulse; c.swingImpulse = imp;
                wA += c.swingLimitAngA * l; wB -= c.swingLimitAngB * l;
            }
            {
                V3 avA = vA + cross(wA, c.rA), avB = vB + cross(wB, c.rB);
                V3 Cdot = avB - avA + c.bias;
                V3 P = solve3(c.invMass, -Cdot);
                vA -= A.invMass * P; wA -= A.invInertia * cross(c.rA, P);
                vB += B.invMass * P; wB += B.invInertia * cross(c.rB, P);
            }
            A.v = vA; A.w = wA; B.v = vB; B.w = wB;
        } break;
        default: // slider, constraints.cpp:2763-2846 : motor -> limit -> rotation -> position
        {
            SliderRow& c = ((SliderRow*)jointRows[5].data())[i]; BodyG& A = g[c.a]; BodyG& B = g[c.b];
            V3 vA = A.v, wA = A.w, vB = B.v, wB = B.w;
            if (c.solveMotor)
            {
                float Cdot = dot(vB, c.axis) - dot(vA, c.axis) - c.motorVelocity;
                float mass = 1.f / (A.invMass + B.invMass);
                float l = -mass * Cdot;
                float old = c.motorImpulse;
                c.motorImpulse = clampf(c.motorImpulse + l, -c.maxMotorImpulse, c.maxMotorImpulse);
                l = c.motorImpulse - old;
                V3 P = l * c.axis;
                vA -= A.invMass * P; vB += B.invMass * P;
            }
            if (c.solveLimit)
            {
                float Cdot = dot(vB, c.axis) + dot(wB, c.rBxs) - dot(vA, c.axis) - dot(wA, c.rAuxs);
                float l = -c.axialMass * (c.limitSign * Cdot + c.limitBias);
                float imp = fmax2(c.limitImpulse + l, 0.f);
                l = imp - c.limitImpulse; c.limitImpulse = imp;
                l *= c.limitSign;
                V3 P = l * c.axis;
                vA -= A.invMass * P; wA -= c.limAngA * l; vB += B.invMass * P; wB += c.limAngB * l;
            }
            {
                V3 Cdot = wB - wA;
                V3 l = solve3(c.invRMass, -(Cdot + c.rBias));
                wA -= A.invInertia * l; wB += B.invInertia * l;
            }
            {
                float cx = dot(c.tangent, vB) + dot(c.rBxt, wB) - dot(c.tangent, vA) - dot(c.rAuxt, wA);
                float cy = dot(c.bitangent, vB) + dot(c.rBxb, wB) - dot(c.bitangent, vA) - dot(c.rAuxb, wA);
                float lx, ly; solve2(c.invTMass, -(cx + c.tBiasX), -(cy + c.tBiasY), lx, ly);
                V3 tb = c.tangent * lx + c.bitangent * ly;
                vA -= A.invMass * tb; wA -= A.invInertia * (c.rAuxt * lx + c.rAuxb * ly);
                vB += B.invMass * tb; wB += B.invInertia * (c.rBxt * lx + c.rBxb * ly);
            }
            A.v = vA; A.w = wA; B.v = vB; B.w = wB;
        } break;
        }
    }
}

} // namespace orc
