// puckswarm_b200 -- lane-per-island solver with a uniform instruction stream (device only).
//
// 32 small islands share a warp, one per lane. To keep the warp converged the solver is written without the
// body-kind / wall / manifold-type branches of the generic routines (ps_physics.cuh): every staged body, the wall
// included, has the same record (pose, velocity, transform, kind) in a lane-interleaved shared-memory slot, inverse
// masses come from selects on the kind, the face-A / face-B manifolds are one computation with the roles swapped, and
// the lanes walk warm start, velocity sweeps, integration and position sweeps in lock step (trip counts are the warp
// maximum). Only three short regions diverge: circle vs face normals, the 2-point block solver, and the robots' sin /
// cos after a correction. The arithmetic -- every operation and its order -- is that of warmStartOne, solveVelocityOne,
// correctPoint and integrateBody, so the results are bit-identical to the sequential solver.
#pragma once
#include "ps_pipeline.cuh"

#if defined(__CUDACC__)
namespace ps {

enum { LK_PUCK = 0, LK_ROBOT = 1, LK_WALL = 2 };

// A warp's lane-interleaved slot: NBL dynamic bodies (+ the wall) and NCL constraints per lane. 39 bytes per body: the
// velocity (used up to the integration) and the transform's px / py / c (used from the integration on) share storage -- the
// slot's size is what bounds how many islands an SM holds.
template <int NBL, int NCL>
struct LaneSlotsT {
    static constexpr int kBodies = NBL, kContacts = NCL, kLBX = NBL + 1;   // staged bodies per lane + the wall
    float cx[kLBX][32], cy[kLBX][32], a[kLBX][32];
    float u0[kLBX][32], u1[kLBX][32], u2[kLBX][32];   // vx, vy, w until lkIntegrateLane; then px, py, c
    float s[kLBX][32];
    float pm[kLBX][32], pi[kLBX][32];   // position-solver inverse masses (mass * invMass, mass * invI; 0 for the wall): selecting them from the kind in the sweeps costs more than the 8 bytes
    uint16_t gidx[kLBX][32];
    uint8_t kind[kLBX][32];
    // what the position sweeps read of each constraint: localNormal, localPoint, the two manifold points, radius, and
    // bodyA | bodyB << 8 | type << 16 | points << 24 (up to 100 sweeps re-read these: shared memory instead of L2 latency)
    float pc[NCL][9][32];
    uint32_t pcHead[NCL][32];
};
typedef LaneSlotsT<kLaneBodies, kLaneContacts> LaneSlotsU;
typedef LaneSlotsT<kLaneXBodies, kLaneXContacts> LaneSlotsX;

template <class Slots>
struct LaneCtxT {
    typedef Slots SlotsType;
    Slots* S;
    int lane;
    int nb;          // staged dynamic bodies; local index nb is the wall
    int lo, nc;      // this island's range in wk.order
    float pmPuck, piPuck, pmRobot, piRobot;
    // body-table accessors (the solver routines below are written against these; SlotCtx is the other implementation)
    __device__ __forceinline__ float& cx(int l) const { return S->cx[l][lane]; }
    __device__ __forceinline__ float& cy(int l) const { return S->cy[l][lane]; }
    __device__ __forceinline__ float& a(int l) const { return S->a[l][lane]; }
    __device__ __forceinline__ float& vx(int l) const { return S->u0[l][lane]; }
    __device__ __forceinline__ float& vy(int l) const { return S->u1[l][lane]; }
    __device__ __forceinline__ float& w(int l) const { return S->u2[l][lane]; }
    __device__ __forceinline__ float& px(int l) const { return S->u0[l][lane]; }
    __device__ __forceinline__ float& py(int l) const { return S->u1[l][lane]; }
    __device__ __forceinline__ float& c(int l) const { return S->u2[l][lane]; }
    __device__ __forceinline__ float& s(int l) const { return S->s[l][lane]; }
    __device__ __forceinline__ float pm(int l) const { return S->pm[l][lane]; }
    __device__ __forceinline__ float pi(int l) const { return S->pi[l][lane]; }
    __device__ __forceinline__ int kind(int l) const { return S->kind[l][lane]; }
    __device__ __forceinline__ int gidx(int l) const { return S->gidx[l][lane]; }
};
typedef LaneCtxT<LaneSlotsU> LaneCtx;

__device__ __forceinline__ float lkInvMass(const HotConsts& h, int kind) { return kind == LK_PUCK ? h.puck.invMass : (kind == LK_ROBOT ? h.robot.invMass : 0.0f); }
__device__ __forceinline__ float lkInvI(const HotConsts& h, int kind) { return kind == LK_PUCK ? h.puck.invI : (kind == LK_ROBOT ? h.robot.invI : 0.0f); }

// Body.synchronizeTransform of staged body l
template <class BT>
__device__ __forceinline__ void lkSyncXf(const PhysEnv& e, const BT& L, int l) {
    if (L.kind(l) == LK_ROBOT) {
        Xf T = makeXf(e.trig, L.cx(l), L.cy(l), L.a(l), e.hot.robot.lcx, e.hot.robot.lcy);
        L.px(l) = T.px;
        L.py(l) = T.py;
        L.c(l) = T.c;
        L.s(l) = T.s;
    } else {
        L.px(l) = L.cx(l);
        L.py(l) = L.cy(l);
    }
}

// stage island k of the arena behind wk into this lane's slot; constraint body references become staged indices.
// Up to the integration the slot holds poses and velocities; the transforms follow in lkIntegrateLane.
template <class LC>
__device__ __forceinline__ void lkStage(const PhysEnv& e, const PhysWork& wk, int k, LC& L) {
    typename LC::SlotsType& S = *L.S;
    int ln = L.lane;
    int blo = wk.islandBodyStart[k];
    int nb = wk.islandBodyStart[k + 1] - blo;
    int P = e.cfg.d.P;
    L.nb = nb;
    L.lo = wk.islandStart[k];
    L.nc = wk.islandStart[k + 1] - L.lo;
    for (int l = 0; l < nb; ++l) {
        int g = wk.ibody[blo + l];
        S.gidx[l][ln] = (uint16_t)g;
        S.cx[l][ln] = wk.cx[g];
        S.cy[l][ln] = wk.cy[g];
        S.a[l][ln] = wk.a[g];
        S.u0[l][ln] = wk.vx[g];
        S.u1[l][ln] = wk.vy[g];
        S.u2[l][ln] = wk.w[g];
        S.kind[l][ln] = g >= P ? LK_ROBOT : LK_PUCK;
        S.pm[l][ln] = g >= P ? L.pmRobot : L.pmPuck;
        S.pi[l][ln] = g >= P ? L.piRobot : L.piPuck;
    }
    // the wall: a body that never moves (zero inverse masses), sweep.c = (0, 0); its transform follows at the integration
    S.kind[nb][ln] = LK_WALL;
    S.pm[nb][ln] = 0.0f;
    S.pi[nb][ln] = 0.0f;
    S.gidx[nb][ln] = 0xFFFF;
    S.cx[nb][ln] = 0.0f;
    S.cy[nb][ln] = 0.0f;
    S.a[nb][ln] = 0.0f;
    S.u0[nb][ln] = 0.0f;
    S.u1[nb][ln] = 0.0f;
    S.u2[nb][ln] = 0.0f;
    for (int i = 0; i < L.nc; ++i) {
        TC& t = wk.tc[wk.order[L.lo + i]];
        int a = t.bodyA, b = t.bodyB;
        int la = nb, lb = nb;
        for (int l = 0; l < nb; ++l) {
            int g = S.gidx[l][ln];
            if (g == a) la = l;
            if (g == b) lb = l;
        }
        t.bodyA = la;
        t.bodyB = lb;
        S.pc[i][0][ln] = t.localNormal.x;
        S.pc[i][1][ln] = t.localNormal.y;
        S.pc[i][2][ln] = t.localPoint.x;
        S.pc[i][3][ln] = t.localPoint.y;
        S.pc[i][4][ln] = t.lp[0].x;
        S.pc[i][5][ln] = t.lp[0].y;
        S.pc[i][6][ln] = t.lp[1].x;
        S.pc[i][7][ln] = t.lp[1].y;
        S.pc[i][8][ln] = t.radius;
        S.pcHead[i][ln] = (uint32_t)la | ((uint32_t)lb << 8) | ((uint32_t)t.type << 16) | ((uint32_t)t.solverPoints << 24);
    }
}

// the velocities went back at the integration (lkIntegrateLane); what is left are the poses and the robots' transforms
template <class LC>
__device__ __forceinline__ void lkFinish(const PhysEnv& e, const PhysWork& wk, const LC& L) {
    typename LC::SlotsType& S = *L.S;
    int ln = L.lane;
    int P = e.cfg.d.P;
    for (int l = 0; l < L.nb; ++l) {
        int g = S.gidx[l][ln];
        wk.cx[g] = S.cx[l][ln];
        wk.cy[g] = S.cy[l][ln];
        wk.a[g] = S.a[l][ln];
        if (g >= P) {
            int r = g - P;
            wk.rpx[r] = S.u0[l][ln];
            wk.rpy[r] = S.u1[l][ln];
            wk.rc[r] = S.u2[l][ln];
            wk.rs[r] = S.s[l][ln];
        }
    }
}

// ContactSolver.warmStart for one constraint (= warmStartOne)
template <class BT>
__device__ __forceinline__ void lkWarmStart(const PhysEnv& e, const BT& L, const TC& c) {
    int A = c.bodyA, B = c.bodyB;
    int kA = L.kind(A), kB = L.kind(B);
    V2 vA = mk(L.vx(A), L.vy(A)), vB = mk(L.vx(B), L.vy(B));
    float wA = L.w(A), wB = L.w(B);
    float invMassA = lkInvMass(e.hot, kA), invIA = lkInvI(e.hot, kA), invMassB = lkInvMass(e.hot, kB), invIB = lkInvI(e.hot, kB);
    V2 normal = c.normal, tangent = crossVS(normal, 1.0f);
    for (int j = 0; j < c.solverPoints; ++j) {
        V2 P = c.nImp[j] * normal + c.tImp[j] * tangent;
        wA -= invIA * cross(c.rA[j], P);
        vA = vA - invMassA * P;
        wB += invIB * cross(c.rB[j], P);
        vB = vB + invMassB * P;
    }
    if (kA != LK_WALL) {
        L.vx(A) = vA.x;
        L.vy(A) = vA.y;
        L.w(A) = wA;
    }
    if (kB != LK_WALL) {
        L.vx(B) = vB.x;
        L.vy(B) = vB.y;
        L.w(B) = wB;
    }
}

// ContactSolver.solveVelocityConstraints for one constraint (= solveVelocityOne)
template <class BT>
__device__ __forceinline__ void lkSolveVelocity(const PhysEnv& e, const BT& L, TC& c) {
    int A = c.bodyA, B = c.bodyB;
    int kA = L.kind(A), kB = L.kind(B);
    V2 vA = mk(L.vx(A), L.vy(A)), vB = mk(L.vx(B), L.vy(B));
    float wA = L.w(A), wB = L.w(B);
    float invMassA = lkInvMass(e.hot, kA), invIA = lkInvI(e.hot, kA), invMassB = lkInvMass(e.hot, kB), invIB = lkInvI(e.hot, kB);
    V2 normal = c.normal, tangent = crossVS(normal, 1.0f);
    float friction = e.hot.friction;
    for (int j = 0; j < c.solverPoints; ++j) {
        V2 dv = vB + crossSV(wB, c.rB[j]) - vA - crossSV(wA, c.rA[j]);
        float vt = dot(dv, tangent);
        float lambda = c.tMass[j] * (-vt);
        float maxFriction = friction * c.nImp[j];
        float newImpulse = clampf(c.tImp[j] + lambda, -maxFriction, maxFriction);
        lambda = newImpulse - c.tImp[j];
        V2 P = lambda * tangent;
        vA = vA - invMassA * P;
        wA -= invIA * cross(c.rA[j], P);
        vB = vB + invMassB * P;
        wB += invIB * cross(c.rB[j], P);
        c.tImp[j] = newImpulse;
    }
    if (c.solverPoints == 1) {
        V2 dv = vB + crossSV(wB, c.rB[0]) - vA - crossSV(wA, c.rA[0]);
        float vn = dot(dv, normal);
        float lambda = -c.nMass[0] * (vn - 0.0f);
        float newImpulse = fmaxf_(c.nImp[0] + lambda, 0.0f);
        lambda = newImpulse - c.nImp[0];
        V2 P = lambda * normal;
        vA = vA - invMassA * P;
        wA -= invIA * cross(c.rA[0], P);
        vB = vB + invMassB * P;
        wB += invIB * cross(c.rB[0], P);
        c.nImp[0] = newImpulse;
    } else {
        V2 a = mk(c.nImp[0], c.nImp[1]);
        V2 dv1 = vB + crossSV(wB, c.rB[0]) - vA - crossSV(wA, c.rA[0]);
        V2 dv2 = vB + crossSV(wB, c.rB[1]) - vA - crossSV(wA, c.rA[1]);
        float vn1 = dot(dv1, normal), vn2 = dot(dv2, normal);
        V2 b = mk(vn1 - 0.0f, vn2 - 0.0f);
        b = b - mk(c.k11 * a.x + c.k12 * a.y, c.k12 * a.x + c.k22 * a.y);
        V2 x;
        bool ok = false;
        V2 mb = mk(c.m11 * b.x + c.m12 * b.y, c.m12 * b.x + c.m22 * b.y);
        x = -mb;
        if (x.x >= 0.0f && x.y >= 0.0f) ok = true;
        if (!ok) {
            x.x = -c.nMass[0] * b.x;
            x.y = 0.0f;
            vn2 = c.k12 * x.x + b.y;
            if (x.x >= 0.0f && vn2 >= 0.0f) ok = true;
        }
        if (!ok) {
            x.x = 0.0f;
            x.y = -c.nMass[1] * b.y;
            vn1 = c.k12 * x.y + b.x;
            if (x.y >= 0.0f && vn1 >= 0.0f) ok = true;
        }
        if (!ok) {
            x.x = 0.0f;
            x.y = 0.0f;
            if (b.x >= 0.0f && b.y >= 0.0f) ok = true;
        }
        if (ok) {
            V2 dd = x - a;
            V2 P1 = dd.x * normal, P2 = dd.y * normal;
            vA = vA - invMassA * (P1 + P2);
            wA -= invIA * (cross(c.rA[0], P1) + cross(c.rA[1], P2));
            vB = vB + invMassB * (P1 + P2);
            wB += invIB * (cross(c.rB[0], P1) + cross(c.rB[1], P2));
            c.nImp[0] = x.x;
            c.nImp[1] = x.y;
        }
    }
    if (kA != LK_WALL) {
        L.vx(A) = vA.x;
        L.vy(A) = vA.y;
        L.w(A) = wA;
    }
    if (kB != LK_WALL) {
        L.vx(B) = vB.x;
        L.vy(B) = vB.y;
        L.w(B) = wB;
    }
}

// One position correction of one manifold point (= positionManifold + correctPoint). Returns the separation.
template <class BT>
__device__ __forceinline__ float lkCorrectPoint(const PhysEnv& e, const BT& L, int A, int B, int type, V2 localNormal, V2 localPoint, V2 pointLocal,
                                                float radius, bool* changed) {
    Xf xfA, xfB;
    xfA.px = L.px(A);
    xfA.py = L.py(A);
    xfA.c = L.c(A);
    xfA.s = L.s(A);
    xfB.px = L.px(B);
    xfB.py = L.py(B);
    xfB.c = L.c(B);
    xfB.s = L.s(B);
    // circles: pointA = xfA * localPoint, pointB = xfB * pointLocal; face A: plane = xfA * localPoint, clip = xfB * pointLocal;
    // face B: plane = xfB * localPoint, clip = xfA * pointLocal  -> one computation with the reference side selected
    bool faceB = type == MF_FACE_B;
    Xf xfRef = faceB ? xfB : xfA, xfInc = faceB ? xfA : xfB;
    V2 p1 = mulX(xfRef, localPoint), p2 = mulX(xfInc, pointLocal);
    V2 normal, point;
    float separation;
    if (type == MF_CIRCLES) {
        if (distSq(p1, p2) > kEpsilon * kEpsilon) {
            normal = p2 - p1;
            normalize(normal);
        } else {
            normal = mk(1.0f, 0.0f);
        }
        point = 0.5f * (p1 + p2);
        separation = dot(p2 - p1, normal) - radius;
    } else {
        normal = mulR(xfRef, localNormal);
        separation = dot(p2 - p1, normal) - radius;
        point = p2;
        if (faceB) normal = -normal;
    }
    V2 cA = mk(L.cx(A), L.cy(A)), cB = mk(L.cx(B), L.cy(B));
    V2 rA = point - cA, rB = point - cB;
    float C = clampf(kContactBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
    if (C == 0.0f) return separation;
    // mass-scaled inverse masses per staged body; the wall's are zero, so it "moves" by -0 * P: never
    float invMassA = L.pm(A), invIA = L.pi(A), invMassB = L.pm(B), invIB = L.pi(B);
    float rnA = cross(rA, normal), rnB = cross(rB, normal);
    float K = invMassA + invMassB + invIA * rnA * rnA + invIB * rnB * rnB;
    float impulse = K > 0.0f ? -C / K : 0.0f;
    V2 P = impulse * normal;
    bool moved = false;
    {
        V2 nc = cA - invMassA * P;
        float oa = L.a(A);
        float na = oa - invIA * cross(rA, P);
        if (nc.x != cA.x || nc.y != cA.y || na != oa) {
            L.cx(A) = nc.x;
            L.cy(A) = nc.y;
            L.a(A) = na;
            lkSyncXf(e, L, A);
            moved = true;
        }
    }
    {
        V2 nc = cB + invMassB * P;
        float oa = L.a(B);
        float na = oa + invIB * cross(rB, P);
        if (nc.x != cB.x || nc.y != cB.y || na != oa) {
            L.cx(B) = nc.x;
            L.cy(B) = nc.y;
            L.a(B) = na;
            lkSyncXf(e, L, B);
            moved = true;
        }
    }
    if (moved) *changed = true;
    return separation;
}

// The same correction by a PAIR of lanes (warp-per-island solver): both lanes evaluate the point identically, then lane
// `side` 0 moves body A and lane `side` 1 moves body B, so the two bodies' updates (for a robot: the transform with its two
// table look-ups) leave each other's dependency chain. x - y == x + (-y) and (-m) * P == -(m * P) exactly, so "own body,
// signed inverse masses" is the reference's arithmetic bit for bit. Every lane of the warp calls this together (inactive
// lanes with on = false): the reads of a point are separated from its writes, and the writes from the next reads, by
// warp barriers.
template <class BT>
__device__ __forceinline__ float lkCorrectPointPair(const PhysEnv& e, const BT& L, bool on, int side, int A, int B, int type, V2 localNormal, V2 localPoint,
                                                    V2 pointLocal, float radius, bool* changed) {
    float separation = 0.0f;
    bool write = false;
    int me = side ? B : A;
    V2 nc = mk(0.0f, 0.0f);
    float na = 0.0f;
    if (on) {
        Xf xfA, xfB;
        xfA.px = L.px(A);
        xfA.py = L.py(A);
        xfA.c = L.c(A);
        xfA.s = L.s(A);
        xfB.px = L.px(B);
        xfB.py = L.py(B);
        xfB.c = L.c(B);
        xfB.s = L.s(B);
        bool faceB = type == MF_FACE_B;
        Xf xfRef = faceB ? xfB : xfA, xfInc = faceB ? xfA : xfB;
        V2 p1 = mulX(xfRef, localPoint), p2 = mulX(xfInc, pointLocal);
        V2 normal, point;
        if (type == MF_CIRCLES) {
            if (distSq(p1, p2) > kEpsilon * kEpsilon) {
                normal = p2 - p1;
                normalize(normal);
            } else {
                normal = mk(1.0f, 0.0f);
            }
            point = 0.5f * (p1 + p2);
            separation = dot(p2 - p1, normal) - radius;
        } else {
            normal = mulR(xfRef, localNormal);
            separation = dot(p2 - p1, normal) - radius;
            point = p2;
            if (faceB) normal = -normal;
        }
        float C = clampf(kContactBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
        if (C != 0.0f) {
            V2 cA = mk(L.cx(A), L.cy(A)), cB = mk(L.cx(B), L.cy(B));
            V2 rA = point - cA, rB = point - cB;
            float invMassA = L.pm(A), invIA = L.pi(A), invMassB = L.pm(B), invIB = L.pi(B);
            float rnA = cross(rA, normal), rnB = cross(rB, normal);
            float K = invMassA + invMassB + invIA * rnA * rnA + invIB * rnB * rnB;
            float impulse = K > 0.0f ? -C / K : 0.0f;
            V2 P = impulse * normal;
            V2 cM = side ? cB : cA, rM = side ? rB : rA;
            float sm = side ? invMassB : -invMassA, si = side ? invIB : -invIA;
            float oa = L.a(me);
            nc = cM + sm * P;
            na = oa + si * cross(rM, P);
            write = nc.x != cM.x || nc.y != cM.y || na != oa;
        }
    }
    __syncwarp();
    if (write) {
        L.cx(me) = nc.x;
        L.cy(me) = nc.y;
        L.a(me) = na;
        lkSyncXf(e, L, me);
        *changed = true;
    }
    __syncwarp();
    return separation;
}

// Island.solve step 6 for staged body l (= integrateBody); the start-of-step pose / transform go straight to HBM
template <class BT>
__device__ __forceinline__ void lkIntegrate(const PhysEnv& e, const PhysWork& g, const BT& L, int l) {
    int gb = L.gidx(l);
    int P = e.cfg.d.P;
    float dt = e.cfg.dt;
    float vx = L.vx(l), vy = L.vy(l), w = L.w(l);
    V2 tr = mk(dt * vx, dt * vy);
    if (dot(tr, tr) > kMaxTranslationSquared) {
        float ratio = kMaxTranslation / len(tr);
        vx *= ratio;
        vy *= ratio;
    }
    float rot = dt * w;
    if (rot * rot > kMaxRotationSquared) {
        float ratio = kMaxRotation / fabsf(rot);
        w *= ratio;
    }
    L.vx(l) = vx;
    L.vy(l) = vy;
    L.w(l) = w;
    g.c0x[gb] = L.cx(l);
    g.c0y[gb] = L.cy(l);
    g.a0[gb] = L.a(l);
    if (gb >= P) {
        int r = gb - P;
        g.r0px[r] = L.px(l);
        g.r0py[r] = L.py(l);
        g.r0c[r] = L.c(l);
        g.r0s[r] = L.s(l);
    }
    L.cx(l) += dt * vx;
    L.cy(l) += dt * vy;
    L.a(l) += dt * w;
    lkSyncXf(e, L, l);
}

// Island.solve step 6 for staged body l of a lane slot (= integrateBody). The clamped velocity goes back to the arena here (it
// does not change afterwards), the start-of-step pose / transform go straight to HBM, and the slot's velocity storage
// becomes the transform of the advanced pose.
template <class LC>
__device__ __forceinline__ void lkIntegrateLane(const PhysEnv& e, const PhysWork& g, const LC& L, int l) {
    int gb = L.gidx(l);
    int P = e.cfg.d.P;
    float dt = e.cfg.dt;
    float vx = L.vx(l), vy = L.vy(l), w = L.w(l);
    V2 tr = mk(dt * vx, dt * vy);
    if (dot(tr, tr) > kMaxTranslationSquared) {
        float ratio = kMaxTranslation / len(tr);
        vx *= ratio;
        vy *= ratio;
    }
    float rot = dt * w;
    if (rot * rot > kMaxRotationSquared) {
        float ratio = kMaxRotation / fabsf(rot);
        w *= ratio;
    }
    g.vx[gb] = vx;
    g.vy[gb] = vy;
    g.w[gb] = w;
    g.c0x[gb] = L.cx(l);
    g.c0y[gb] = L.cy(l);
    g.a0[gb] = L.a(l);
    L.cx(l) += dt * vx;
    L.cy(l) += dt * vy;
    L.a(l) += dt * w;
    if (gb >= P) {
        int r = gb - P;
        g.r0px[r] = g.rpx[r];
        g.r0py[r] = g.rpy[r];
        g.r0c[r] = g.rc[r];
        g.r0s[r] = g.rs[r];
        Xf T = makeXf(e.trig, L.cx(l), L.cy(l), L.a(l), e.hot.robot.lcx, e.hot.robot.lcy);
        L.px(l) = T.px;
        L.py(l) = T.py;
        L.c(l) = T.c;
        L.s(l) = T.s;
    } else {
        L.px(l) = L.cx(l);
        L.py(l) = L.cy(l);
        L.c(l) = 1.0f;
        L.s(l) = 0.0f;
    }
}
template <class LC>
__device__ __forceinline__ void lkWallXf(const PhysEnv& e, const LC& L) {
    L.px(L.nb) = e.hot.wallXf.px;
    L.py(L.nb) = e.hot.wallXf.py;
    L.c(L.nb) = e.hot.wallXf.c;
    L.s(L.nb) = e.hot.wallXf.s;
}

// Island.solve of up to 32 staged islands in lock step. `has`: this lane holds an island. Returns its position iterations.
template <class LC>
__device__ __forceinline__ int lkSolve(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, const LC& L, bool has) {
    const unsigned FULL = 0xFFFFFFFFu;
    int nc = has ? L.nc : 0;
    int maxNc = nc;
    for (int o = 16; o > 0; o >>= 1) maxNc = max(maxNc, __shfl_xor_sync(FULL, maxNc, o));
    for (int i = 0; i < maxNc; ++i) {
        if (i < nc) lkWarmStart(e, L, wk.tc[wk.order[L.lo + i]]);
        __syncwarp();
    }
    for (int it = 0; it < e.cfg.velIters; ++it)
        for (int i = 0; i < maxNc; ++i) {
            if (i < nc) lkSolveVelocity(e, L, wk.tc[wk.order[L.lo + i]]);
            __syncwarp();
        }
    if (has) {
        for (int i = 0; i < nc; ++i) {
            const TC& t = wk.tc[wk.order[L.lo + i]];
            for (int j = 0; j < t.solverPoints; ++j) {
                st.cimp[4 * t.ci + 2 * j] = t.nImp[j];
                st.cimp[4 * t.ci + 2 * j + 1] = t.tImp[j];
            }
        }
        for (int l = 0; l < L.nb; ++l) lkIntegrateLane(e, wk, L, l);
        lkWallXf(e, L);
    }
    __syncwarp();
    int iters = 0;
    bool running = has && e.cfg.posIters > 0;
    while (__any_sync(FULL, running)) {
        float minSep = 0.0f;
        bool changed = false;
        for (int i = 0; i < maxNc; ++i) {
            if (running && i < nc) {
                const typename LC::SlotsType& S = *L.S;
                int ln = L.lane;
                uint32_t hd = S.pcHead[i][ln];
                int bA = (int)(hd & 255u), bB = (int)((hd >> 8) & 255u), type = (int)((hd >> 16) & 255u), np = (int)(hd >> 24);
                V2 lnm = mk(S.pc[i][0][ln], S.pc[i][1][ln]), lpt = mk(S.pc[i][2][ln], S.pc[i][3][ln]);
                float radius = S.pc[i][8][ln];
                for (int j = 0; j < np; ++j) {
                    float sep = lkCorrectPoint(e, L, bA, bB, type, lnm, lpt, mk(S.pc[i][4 + 2 * j][ln], S.pc[i][5 + 2 * j][ln]), radius, &changed);
                    minSep = fminf_(minSep, sep);
                }
            }
            __syncwarp();
        }
        if (running) {
            ++iters;
            if (minSep >= -1.5f * kLinearSlop || iters >= e.cfg.posIters) running = false;
            else if (!changed) {   // fixed point: every later sweep repeats this one
                iters = e.cfg.posIters;
                running = false;
            }
        }
    }
    return iters;
}

// Warp-per-island solver. The warp stages the island into its shared-memory slot: one body table (the wall included, as
// the body after the last dynamic one), what the position sweeps read of each constraint, and -- when the slot is
// built with STAGE_TC -- the constraints themselves (otherwise the four velocity-phase passes read them from HBM / L2).
//
// Rows. A sweep over the constraints is cut into rows: constraint i (in solve order) goes to the row after the last row
// that holds an earlier constraint sharing a dynamic body with it. The constraints of a row touch disjoint bodies, so
// the lanes solve them together; rows run in order, so every body still sees its constraints in the reference's order
// and the sweep is bit-identical to the sequential one.
//
// (Two ways of running corrections of different sweeps side by side -- a modulo schedule over the rows, and a dataflow
// order driven by per-body correction counters -- were built, verified bit-exact, and measured slower: the solve order of
// real islands leaves ~2 independent corrections per step, less than the bookkeeping costs. The plain row schedule stays.)
static constexpr int kWarpBodies = kQueueWarpBodies, kWarpContacts = kQueueWarpContacts;
static constexpr int kBigBodies = 254, kBigContacts = 640;
// A staged body as the position sweeps read it: two 16-byte loads give a correction everything it needs of a body (sweep.c,
// the mass-scaled inverse masses, the transform); the angle and the kind are read by the lane that moves the body only.
struct __align__(16) BodyRec {
    float cx, cy, pm, pi;     // sweep.c; mass * invMass, mass * invI (0 for the wall)
    float px, py, c, s;       // transform
    float a;                  // sweep.a
    uint32_t kind;            // LK_*
    uint32_t gidx;            // body index in the arena (0xFFFF: the wall)
    uint32_t bodyRow;         // row-schedule scratch
};
// What the position sweeps read of a constraint, in ROW-MAJOR solve order (three 16-byte loads, no index indirection)
struct __align__(16) ConRec {
    float lnx, lny, lpx, lpy;   // localNormal, localPoint
    float p0x, p0y, p1x, p1y;   // the two manifold points
    float radius;
    uint32_t offs;              // byte offsets of bodyA | bodyB << 16 in IslandSlot::br
    uint32_t typeNp;            // manifold type | points << 8
    uint32_t pad;
};
template <int NB_, int NC_, bool STAGE_TC>
struct __align__(16) IslandSlot {
    static constexpr int NBX = NB_ + 1;
    static constexpr int kBodies = NB_, kContacts = NC_;
    static constexpr bool kStageTC = STAGE_TC;
    BodyRec br[NBX];
    ConRec cr[NC_];
    float vx[NBX], vy[NBX], w[NBX];
    uint16_t pcHead[NC_];                  // bodyA | bodyB << 8 (staged indices) in solve order: input of the row schedule
    uint16_t row[NC_], sched[NC_], rowStart[NC_ + 2];
    int nRows;
    TC tc[STAGE_TC ? NC_ : 1];
};
// The velocity-phase constraints (TC, 164 B each) stay in HBM / L2 for both slot sizes: staging them made a warp slot 17.7 KB
// instead of 5.6 KB, and the shared memory the island kernels hold is what keeps the other arena groups' wide kernels off the
// SMs while an island tail runs (measured: +8.5 % on the whole step in the steady state, profiles/r01_summary.md).
#ifndef PS_WARP_STAGE_TC
#define PS_WARP_STAGE_TC 0
#endif
typedef IslandSlot<kWarpBodies, kWarpContacts, PS_WARP_STAGE_TC != 0> WarpSlot;
typedef IslandSlot<kBigBodies, kBigContacts, false> BigSlot;

template <class Slot>
struct SlotCtx {
    Slot* S;
    __device__ __forceinline__ float& cx(int l) const { return S->cx[l]; }
    __device__ __forceinline__ float& cy(int l) const { return S->cy[l]; }
    __device__ __forceinline__ float& a(int l) const { return S->a[l]; }
    __device__ __forceinline__ float& vx(int l) const { return S->vx[l]; }
    __device__ __forceinline__ float& vy(int l) const { return S->vy[l]; }
    __device__ __forceinline__ float& w(int l) const { return S->w[l]; }
    __device__ __forceinline__ float& px(int l) const { return S->px[l]; }
    __device__ __forceinline__ float& py(int l) const { return S->py[l]; }
    __device__ __forceinline__ float& c(int l) const { return S->c[l]; }
    __device__ __forceinline__ float& s(int l) const { return S->s[l]; }
    __device__ __forceinline__ float pm(int l) const { return S->pm[l]; }
    __device__ __forceinline__ float pi(int l) const { return S->pi[l]; }
    __device__ __forceinline__ int kind(int l) const { return S->kind[l]; }
    __device__ __forceinline__ int gidx(int l) const { return S->gidx[l]; }
};

template <class Slot>
struct RecCtx {   // the same accessors over IslandSlot's body records
    Slot* S;
    __device__ __forceinline__ float& cx(int l) const { return S->br[l].cx; }
    __device__ __forceinline__ float& cy(int l) const { return S->br[l].cy; }
    __device__ __forceinline__ float& a(int l) const { return S->br[l].a; }
    __device__ __forceinline__ float& vx(int l) const { return S->vx[l]; }
    __device__ __forceinline__ float& vy(int l) const { return S->vy[l]; }
    __device__ __forceinline__ float& w(int l) const { return S->w[l]; }
    __device__ __forceinline__ float& px(int l) const { return S->br[l].px; }
    __device__ __forceinline__ float& py(int l) const { return S->br[l].py; }
    __device__ __forceinline__ float& c(int l) const { return S->br[l].c; }
    __device__ __forceinline__ float& s(int l) const { return S->br[l].s; }
    __device__ __forceinline__ float pm(int l) const { return S->br[l].pm; }
    __device__ __forceinline__ float pi(int l) const { return S->br[l].pi; }
    __device__ __forceinline__ int kind(int l) const { return (int)S->br[l].kind; }
    __device__ __forceinline__ int gidx(int l) const { return (int)S->br[l].gidx; }
};

// shared-memory accesses of the position sweeps, by 32-bit shared address
__device__ __forceinline__ float4 lds128(unsigned a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ float2 lds64(unsigned a) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(a) : "memory");
    return v;
}
__device__ __forceinline__ void sts128(unsigned a, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ void sts64(unsigned a, float x, float y) { asm volatile("st.shared.v2.f32 [%0], {%1, %2};" ::"r"(a), "f"(x), "f"(y) : "memory"); }
__device__ __forceinline__ void sts32(unsigned a, float x) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(x) : "memory"); }

// lkCorrectPointPair over body records: `br` is the shared address of IslandSlot::br, offA / offB the bodies' byte offsets.
// The arithmetic is lkCorrectPointPair's, operation for operation.
__device__ __forceinline__ float lkCorrectPointPairRec(const PhysEnv& e, unsigned br, bool on, int side, unsigned offA, unsigned offB, int type, V2 localNormal,
                                                       V2 localPoint, V2 pointLocal, float radius, bool* changed) {
    float separation = 0.0f;
    bool write = false;
    unsigned me = br + (side ? offB : offA);
    V2 nc = mk(0.0f, 0.0f);
    float na = 0.0f, pmM = 0.0f, piM = 0.0f;
    bool robot = false;
    if (on) {
        float4 A0 = lds128(br + offA), A1 = lds128(br + offA + 16), B0 = lds128(br + offB), B1 = lds128(br + offB + 16);
        Xf xfA, xfB;
        xfA.px = A1.x;
        xfA.py = A1.y;
        xfA.c = A1.z;
        xfA.s = A1.w;
        xfB.px = B1.x;
        xfB.py = B1.y;
        xfB.c = B1.z;
        xfB.s = B1.w;
        bool faceB = type == MF_FACE_B;
        Xf xfRef = faceB ? xfB : xfA, xfInc = faceB ? xfA : xfB;
        V2 p1 = mulX(xfRef, localPoint), p2 = mulX(xfInc, pointLocal);
        V2 normal, point;
        if (type == MF_CIRCLES) {
            if (distSq(p1, p2) > kEpsilon * kEpsilon) {
                normal = p2 - p1;
                normalize(normal);
            } else {
                normal = mk(1.0f, 0.0f);
            }
            point = 0.5f * (p1 + p2);
            separation = dot(p2 - p1, normal) - radius;
        } else {
            normal = mulR(xfRef, localNormal);
            separation = dot(p2 - p1, normal) - radius;
            point = p2;
            if (faceB) normal = -normal;
        }
        float C = clampf(kContactBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
        if (C != 0.0f) {
            float2 ak = lds64(me + 32);   // own angle, kind
            V2 cA = mk(A0.x, A0.y), cB = mk(B0.x, B0.y);
            V2 rA = point - cA, rB = point - cB;
            float invMassA = A0.z, invIA = A0.w, invMassB = B0.z, invIB = B0.w;
            float rnA = cross(rA, normal), rnB = cross(rB, normal);
            float K = invMassA + invMassB + invIA * rnA * rnA + invIB * rnB * rnB;
            float impulse = K > 0.0f ? -C / K : 0.0f;
            V2 P = impulse * normal;
            V2 cM = side ? cB : cA, rM = side ? rB : rA;
            float sm = side ? invMassB : -invMassA, si = side ? invIB : -invIA;
            float oa = ak.x;
            nc = cM + sm * P;
            na = oa + si * cross(rM, P);
            write = nc.x != cM.x || nc.y != cM.y || na != oa;
            pmM = side ? invMassB : invMassA;
            piM = side ? invIB : invIA;
            robot = __float_as_uint(ak.y) == (unsigned)LK_ROBOT;
        }
    }
    __syncwarp();
    if (write) {
        sts128(me, nc.x, nc.y, pmM, piM);
        sts32(me + 32, na);
        if (robot) {   // Body.synchronizeTransform (lkSyncXf)
            Xf T = makeXf(e.trig, nc.x, nc.y, na, e.hot.robot.lcx, e.hot.robot.lcy);
            sts128(me + 16, T.px, T.py, T.c, T.s);
        } else {
            sts64(me + 16, nc.x, nc.y);
        }
        *changed = true;
    }
    __syncwarp();
    return separation;
}

// Returns false if the island does not fit the slot (the caller falls back to the sequential path).
// POS = false: the velocity half only (warm start, velocity sweeps, impulses, integration); the bodies are written back and
// the row schedule is left in wk.tcPair[lo + q] = constraint (index into wk.tc) | row << 16, q in row-major solve order, for
// the position-sweep kernel (k_island_pos) that follows on the same stream.
template <class Slot, bool POS = true>
__device__ __forceinline__ bool islandSolveSlot(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, int k, Slot& s, int* itersOut) {
    const unsigned FULL = 0xFFFFFFFFu;
    int lane = threadIdx.x & 31;
    int lo = wk.islandStart[k], hi = wk.islandStart[k + 1];
    int blo = wk.islandBodyStart[k], bhi = wk.islandBodyStart[k + 1];
    int nb = bhi - blo, nc = hi - lo;
    int P = e.cfg.d.P;
    if (nb > Slot::kBodies || nc > Slot::kContacts) return false;
    // ---- stage the bodies (island order) and the wall
    for (int i = lane; i <= nb; i += 32) {
        BodyRec r;
        if (i < nb) {
            int g = wk.ibody[blo + i];
            r.gidx = (uint32_t)g;
            r.cx = wk.cx[g];
            r.cy = wk.cy[g];
            r.a = wk.a[g];
            s.vx[i] = wk.vx[g];
            s.vy[i] = wk.vy[g];
            s.w[i] = wk.w[g];
            if (g >= P) {
                int rr = g - P;
                r.kind = LK_ROBOT;
                r.pm = e.hot.pmRobot;
                r.pi = e.hot.piRobot;
                r.px = wk.rpx[rr];
                r.py = wk.rpy[rr];
                r.c = wk.rc[rr];
                r.s = wk.rs[rr];
            } else {
                r.kind = LK_PUCK;
                r.pm = e.hot.pmPuck;
                r.pi = e.hot.piPuck;
                r.px = r.cx;
                r.py = r.cy;
                r.c = 1.0f;
                r.s = 0.0f;
            }
        } else {
            r.gidx = 0xFFFFu;
            r.kind = LK_WALL;
            r.cx = r.cy = r.a = 0.0f;
            s.vx[i] = s.vy[i] = s.w[i] = 0.0f;
            r.pm = r.pi = 0.0f;
            r.px = e.hot.wallXf.px;
            r.py = e.hot.wallXf.py;
            r.c = e.hot.wallXf.c;
            r.s = e.hot.wallXf.s;
        }
        r.bodyRow = 0;
        s.br[i] = r;
    }
    for (int i = lane; i < nc + 2; i += 32) s.rowStart[i] = 0;
    __syncwarp();
    // ---- constraints: body references -> staged indices (wall -> nb)
    for (int i = lane; i < nc; i += 32) {
        TC& g = wk.tc[wk.order[lo + i]];
        int gA = g.bodyA, gB = g.bodyB;
        int la = nb, lb = nb;
        for (int j = 0; j < nb; ++j) {
            int gi = (int)s.br[j].gidx;
            if (gi == gA) la = j;
            if (gi == gB) lb = j;
        }
        if (Slot::kStageTC) {
            TC t = g;
            t.bodyA = la;
            t.bodyB = lb;
            s.tc[i] = t;
        } else {
            g.bodyA = la;
            g.bodyB = lb;
        }
        s.pcHead[i] = (uint16_t)(la | (lb << 8));
    }
    __syncwarp();
    // ---- rows of one Gauss-Seidel sweep (CSR: rowStart / sched)
    if (lane == 0) {
        int nRows = 0;
        for (int i = 0; i < nc; ++i) {
            int hd = s.pcHead[i];
            int a = hd & 255, b = hd >> 8;
            int r = 0;
            if (a < nb) r = (int)s.br[a].bodyRow;
            if (b < nb && (int)s.br[b].bodyRow > r) r = (int)s.br[b].bodyRow;
            s.row[i] = (uint16_t)r;
            s.rowStart[r + 2]++;                       // counts, shifted by two for the in-place prefix sum below
            if (a < nb) s.br[a].bodyRow = (uint32_t)(r + 1);
            if (b < nb) s.br[b].bodyRow = (uint32_t)(r + 1);
            if (r + 1 > nRows) nRows = r + 1;
        }
        for (int r = 0; r < nRows; ++r) s.rowStart[r + 2] += s.rowStart[r + 1];             // rowStart[r + 1] = start of row r, for now
        for (int i = 0; i < nc; ++i) s.sched[s.rowStart[s.row[i] + 1]++] = (uint16_t)i;   // ... and becomes its end = start of row r + 1
        s.nRows = nRows;
    }
    __syncwarp();
    RecCtx<Slot> L;
    L.S = &s;
    int nRows = s.nRows;
    auto constraint = [&](int ci) -> TC& { return Slot::kStageTC ? s.tc[ci] : wk.tc[wk.order[lo + ci]]; };
    // ---- what the position sweeps read of each constraint, in row-major solve order
    if (POS)
        for (int q = lane; q < nc; q += 32) {
            const TC& t = constraint(s.sched[q]);
            ConRec c;
            c.lnx = t.localNormal.x;
            c.lny = t.localNormal.y;
            c.lpx = t.localPoint.x;
            c.lpy = t.localPoint.y;
            c.p0x = t.lp[0].x;
            c.p0y = t.lp[0].y;
            c.p1x = t.lp[1].x;
            c.p1y = t.lp[1].y;
            c.radius = t.radius;
            c.offs = (uint32_t)t.bodyA * (uint32_t)sizeof(BodyRec) | (((uint32_t)t.bodyB * (uint32_t)sizeof(BodyRec)) << 16);
            c.typeNp = (uint32_t)t.type | ((uint32_t)t.solverPoints << 8);
            c.pad = 0;
            s.cr[q] = c;
        }
    // ContactSolver.warmStart
    for (int r = 0; r < nRows; ++r) {
        for (int q = s.rowStart[r] + lane; q < s.rowStart[r + 1]; q += 32) lkWarmStart(e, L, constraint(s.sched[q]));
        __syncwarp();
    }
    // velocity sweeps
    for (int it = 0; it < e.cfg.velIters; ++it)
        for (int r = 0; r < nRows; ++r) {
            for (int q = s.rowStart[r] + lane; q < s.rowStart[r + 1]; q += 32) lkSolveVelocity(e, L, constraint(s.sched[q]));
            __syncwarp();
        }
    // ContactSolver.storeImpulses
    for (int i = lane; i < nc; i += 32) {
        const TC& t = constraint(i);
        for (int j2 = 0; j2 < t.solverPoints; ++j2) {
            st.cimp[4 * t.ci + 2 * j2] = t.nImp[j2];
            st.cimp[4 * t.ci + 2 * j2 + 1] = t.tImp[j2];
        }
    }
    // integrate the island's bodies
    for (int b = lane; b < nb; b += 32) lkIntegrate(e, wk, L, b);
    __syncwarp();
    if (!POS)
        for (int q = lane; q < nc; q += 32) {
            int ci = s.sched[q];
            wk.tcPair[lo + q] = (uint32_t)wk.order[lo + ci] | ((uint32_t)s.row[ci] << 16);
        }
    // ---- position sweeps
    int posIters = e.cfg.posIters;
    int iters = 0;
    bool done = !POS || posIters <= 0;
    const unsigned brAddr = (unsigned)__cvta_generic_to_shared(&s.br[0]), crAddr = (unsigned)__cvta_generic_to_shared(&s.cr[0]);
    while (!done) {
        float minSep = 0.0f;
        bool changed = false;
        int rowBeg = 0;
        for (int r = 0; r < nRows; ++r) {
            // a lane pair per constraint of the row (16 constraints per pass; a row rarely holds more than three)
            int rowEnd = s.rowStart[r + 1];
            for (int base = rowBeg; base < rowEnd; base += 16) {
                int q = base + (lane >> 1);
                bool on = q < rowEnd;
                unsigned ca = crAddr + (unsigned)(on ? q : 0) * (unsigned)sizeof(ConRec);
                float4 c0 = lds128(ca), c1 = lds128(ca + 16), c2 = lds128(ca + 32);
                unsigned offs = __float_as_uint(c2.y), tn = __float_as_uint(c2.z);
                unsigned offA = offs & 0xFFFFu, offB = offs >> 16;
                int type = (int)(tn & 255u), np = on ? (int)(tn >> 8) : 0;
                V2 ln = mk(c0.x, c0.y), lp = mk(c0.z, c0.w);
                float sep0 = lkCorrectPointPairRec(e, brAddr, np > 0, lane & 1, offA, offB, type, ln, lp, mk(c1.x, c1.y), c2.x, &changed);
                minSep = fminf_(minSep, sep0);
                if (__any_sync(FULL, np > 1)) {
                    float sep1 = lkCorrectPointPairRec(e, brAddr, np > 1, lane & 1, offA, offB, type, ln, lp, mk(c1.z, c1.w), c2.x, &changed);
                    minSep = fminf_(minSep, sep1);
                }
            }
            rowBeg = rowEnd;
        }
        for (int o = 16; o > 0; o >>= 1) {
            float other = __shfl_xor_sync(FULL, minSep, o);
            minSep = fminf_(minSep, other);
        }
        bool anyChanged = __any_sync(FULL, changed);
        ++iters;
        if (minSep >= -1.5f * kLinearSlop || iters >= posIters) done = true;
        else if (!anyChanged) {   // fixed point: every later sweep repeats this one
            iters = posIters;
            done = true;
        }
    }
    if (lane == 0) *itersOut = iters;
    __syncwarp();
    for (int i = lane; i < nb; i += 32) {
        BodyRec r = s.br[i];
        int g = (int)r.gidx;
        wk.cx[g] = r.cx;
        wk.cy[g] = r.cy;
        wk.a[g] = r.a;
        wk.vx[g] = s.vx[i];
        wk.vy[g] = s.vy[i];
        wk.w[g] = s.w[i];
        if (g >= P) {
            int rr = g - P;
            wk.rpx[rr] = r.px;
            wk.rpy[rr] = r.py;
            wk.rc[rr] = r.c;
            wk.rs[rr] = r.s;
        }
    }
    __syncwarp();
    return true;
}
__device__ __forceinline__ bool islandSolveWarp(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, int k, WarpSlot& s, int* itersOut) {
    return islandSolveSlot(e, wk, st, k, s, itersOut);
}
__device__ __forceinline__ bool islandVelocityWarp(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, int k, WarpSlot& s, int* itersOut) {
    return islandSolveSlot<WarpSlot, false>(e, wk, st, k, s, itersOut);
}

// ---- Position sweeps of the warp-bucket islands, G lanes per island and 32 / G islands per warp (k_island_pos) ----
// The velocity half ran in k_island_vel (one warp per island); what is left is the long part: up to 100 Gauss-Seidel
// sweeps whose corrections depend on each other, so a warp that holds ONE island issues from ~3 lanes. Here a warp holds
// 32 / G islands in G-lane groups: every group walks its own row schedule (a lane pair per constraint, as in
// lkCorrectPointPair), one pass per loop iteration, all groups through the same instructions. A group that finishes is
// written back and refilled from the queue by the whole warp (staging is warp-cooperative and short), so the groups never
// wait for each other's islands.
struct PosSlot {
    static constexpr int NBX = kWarpBodies + 1;
    float cx[NBX], cy[NBX], a[NBX];
    float px[NBX], py[NBX], c[NBX], s[NBX], pm[NBX], pi[NBX];
    uint16_t gidx[NBX];
    uint8_t kind[NBX];
    float pc[kWarpContacts][9];            // row-major solve order: localNormal, localPoint, two manifold points, radius
    uint32_t pcHead[kWarpContacts];        // bodyA | bodyB << 8 | type << 16 | points << 24
    uint8_t rowStart[kWarpContacts + 2];
    int nRows;
};

// Stage island `entry` (arena * NB + island) into T with all 32 lanes. Returns false when the island is not this kernel's
// (it outgrew the slot, so the velocity kernel's fallback has already solved it completely).
__device__ __forceinline__ bool posStage(const PhysEnv& e, const PhysWork& wk, int k, PosSlot& T, int lane) {
    int lo = wk.islandStart[k], nc = wk.islandStart[k + 1] - lo;
    int blo = wk.islandBodyStart[k], nb = wk.islandBodyStart[k + 1] - blo;
    int P = e.cfg.d.P;
    if (nb > kWarpBodies || nc > kWarpContacts) return false;
    for (int i = lane; i <= nb; i += 32) {
        if (i < nb) {
            int g = wk.ibody[blo + i];
            T.gidx[i] = (uint16_t)g;
            T.cx[i] = wk.cx[g];
            T.cy[i] = wk.cy[g];
            T.a[i] = wk.a[g];
            if (g >= P) {
                int r = g - P;
                T.kind[i] = LK_ROBOT;
                T.pm[i] = e.hot.pmRobot;
                T.pi[i] = e.hot.piRobot;
                T.px[i] = wk.rpx[r];
                T.py[i] = wk.rpy[r];
                T.c[i] = wk.rc[r];
                T.s[i] = wk.rs[r];
            } else {
                T.kind[i] = LK_PUCK;
                T.pm[i] = e.hot.pmPuck;
                T.pi[i] = e.hot.piPuck;
                T.px[i] = T.cx[i];
                T.py[i] = T.cy[i];
                T.c[i] = 1.0f;
                T.s[i] = 0.0f;
            }
        } else {
            T.gidx[i] = 0xFFFF;
            T.kind[i] = LK_WALL;
            T.cx[i] = T.cy[i] = T.a[i] = 0.0f;
            T.pm[i] = T.pi[i] = 0.0f;
            T.px[i] = e.hot.wallXf.px;
            T.py[i] = e.hot.wallXf.py;
            T.c[i] = e.hot.wallXf.c;
            T.s[i] = e.hot.wallXf.s;
        }
    }
    for (int q = lane; q < nc; q += 32) {
        uint32_t v = wk.tcPair[lo + q];
        int row = (int)(v >> 16);
        const TC& t = wk.tc[v & 0xFFFFu];
        T.pc[q][0] = t.localNormal.x;
        T.pc[q][1] = t.localNormal.y;
        T.pc[q][2] = t.localPoint.x;
        T.pc[q][3] = t.localPoint.y;
        T.pc[q][4] = t.lp[0].x;
        T.pc[q][5] = t.lp[0].y;
        T.pc[q][6] = t.lp[1].x;
        T.pc[q][7] = t.lp[1].y;
        T.pc[q][8] = t.radius;
        T.pcHead[q] = (uint32_t)t.bodyA | ((uint32_t)t.bodyB << 8) | ((uint32_t)t.type << 16) | ((uint32_t)t.solverPoints << 24);
        // rows are contiguous in q: a constraint whose predecessor sits in another row starts its row
        int prevRow = q > 0 ? (int)(wk.tcPair[lo + q - 1] >> 16) : -1;
        if (row != prevRow) T.rowStart[row] = (uint8_t)q;
        if (q == nc - 1) {
            T.rowStart[row + 1] = (uint8_t)nc;
            T.nRows = row + 1;
        }
    }
    return true;
}
__device__ __forceinline__ void posFinish(const PhysEnv& e, const PhysWork& wk, int k, const PosSlot& T, int lane) {
    int nb = wk.islandBodyStart[k + 1] - wk.islandBodyStart[k];
    int P = e.cfg.d.P;
    for (int i = lane; i < nb; i += 32) {
        int g = T.gidx[i];
        wk.cx[g] = T.cx[i];
        wk.cy[g] = T.cy[i];
        wk.a[g] = T.a[i];
        if (g >= P) {
            int r = g - P;
            wk.rpx[r] = T.px[i];
            wk.rpy[r] = T.py[i];
            wk.rc[r] = T.c[i];
            wk.rs[r] = T.s[i];
        }
    }
}
__device__ __forceinline__ bool islandSolveBig(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, int k, BigSlot& s, int* itersOut) {
    return islandSolveSlot(e, wk, st, k, s, itersOut);
}

}  // namespace ps
#endif
