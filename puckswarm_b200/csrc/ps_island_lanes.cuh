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

static constexpr int kLB = kLaneBodies + 1;   // staged bodies per lane + the wall
enum { LK_PUCK = 0, LK_ROBOT = 1, LK_WALL = 2 };

struct LaneSlotsU {
    float cx[kLB][32], cy[kLB][32], a[kLB][32], vx[kLB][32], vy[kLB][32], w[kLB][32];
    float px[kLB][32], py[kLB][32], c[kLB][32], s[kLB][32];
    uint16_t gidx[kLB][32];
    uint8_t kind[kLB][32];
};

struct LaneCtx {
    LaneSlotsU* S;
    int lane;
    int nb;          // staged dynamic bodies; local index nb is the wall
    int lo, nc;      // this island's range in wk.order
};

__device__ __forceinline__ float lkInvMass(const HotConsts& h, int kind) { return kind == LK_PUCK ? h.puck.invMass : (kind == LK_ROBOT ? h.robot.invMass : 0.0f); }
__device__ __forceinline__ float lkInvI(const HotConsts& h, int kind) { return kind == LK_PUCK ? h.puck.invI : (kind == LK_ROBOT ? h.robot.invI : 0.0f); }
__device__ __forceinline__ float lkPosInvMass(const HotConsts& h, int kind) { return kind == LK_PUCK ? h.pmPuck : (kind == LK_ROBOT ? h.pmRobot : 0.0f); }
__device__ __forceinline__ float lkPosInvI(const HotConsts& h, int kind) { return kind == LK_PUCK ? h.piPuck : (kind == LK_ROBOT ? h.piRobot : 0.0f); }

// Body.synchronizeTransform of staged body l
__device__ __forceinline__ void lkSyncXf(const PhysEnv& e, const LaneCtx& L, int l) {
    LaneSlotsU& S = *L.S;
    int ln = L.lane;
    if (S.kind[l][ln] == LK_ROBOT) {
        Xf T = makeXf(e.trig, S.cx[l][ln], S.cy[l][ln], S.a[l][ln], e.hot.robot.lcx, e.hot.robot.lcy);
        S.px[l][ln] = T.px;
        S.py[l][ln] = T.py;
        S.c[l][ln] = T.c;
        S.s[l][ln] = T.s;
    } else {
        S.px[l][ln] = S.cx[l][ln];
        S.py[l][ln] = S.cy[l][ln];
    }
}

// stage island k of the arena behind wk into this lane's slot; constraint body references become staged indices
__device__ __forceinline__ void lkStage(const PhysEnv& e, const PhysWork& wk, int k, LaneCtx& L) {
    LaneSlotsU& S = *L.S;
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
        S.vx[l][ln] = wk.vx[g];
        S.vy[l][ln] = wk.vy[g];
        S.w[l][ln] = wk.w[g];
        if (g >= P) {
            int r = g - P;
            S.kind[l][ln] = LK_ROBOT;
            S.px[l][ln] = wk.rpx[r];
            S.py[l][ln] = wk.rpy[r];
            S.c[l][ln] = wk.rc[r];
            S.s[l][ln] = wk.rs[r];
        } else {
            S.kind[l][ln] = LK_PUCK;
            S.px[l][ln] = wk.cx[g];
            S.py[l][ln] = wk.cy[g];
            S.c[l][ln] = 1.0f;
            S.s[l][ln] = 0.0f;
        }
    }
    // the wall: a body that never moves (zero inverse masses), sweep.c = (0, 0), its transform from the scene
    S.kind[nb][ln] = LK_WALL;
    S.cx[nb][ln] = 0.0f;
    S.cy[nb][ln] = 0.0f;
    S.a[nb][ln] = 0.0f;
    S.vx[nb][ln] = 0.0f;
    S.vy[nb][ln] = 0.0f;
    S.w[nb][ln] = 0.0f;
    S.px[nb][ln] = e.hot.wallXf.px;
    S.py[nb][ln] = e.hot.wallXf.py;
    S.c[nb][ln] = e.hot.wallXf.c;
    S.s[nb][ln] = e.hot.wallXf.s;
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
    }
}

__device__ __forceinline__ void lkFinish(const PhysEnv& e, const PhysWork& wk, const LaneCtx& L) {
    LaneSlotsU& S = *L.S;
    int ln = L.lane;
    int P = e.cfg.d.P;
    for (int l = 0; l < L.nb; ++l) {
        int g = S.gidx[l][ln];
        wk.cx[g] = S.cx[l][ln];
        wk.cy[g] = S.cy[l][ln];
        wk.a[g] = S.a[l][ln];
        wk.vx[g] = S.vx[l][ln];
        wk.vy[g] = S.vy[l][ln];
        wk.w[g] = S.w[l][ln];
        if (g >= P) {
            int r = g - P;
            wk.rpx[r] = S.px[l][ln];
            wk.rpy[r] = S.py[l][ln];
            wk.rc[r] = S.c[l][ln];
            wk.rs[r] = S.s[l][ln];
        }
    }
}

// ContactSolver.warmStart for one constraint (= warmStartOne)
__device__ __forceinline__ void lkWarmStart(const PhysEnv& e, const LaneCtx& L, const TC& c) {
    LaneSlotsU& S = *L.S;
    int ln = L.lane;
    int A = c.bodyA, B = c.bodyB;
    int kA = S.kind[A][ln], kB = S.kind[B][ln];
    V2 vA = mk(S.vx[A][ln], S.vy[A][ln]), vB = mk(S.vx[B][ln], S.vy[B][ln]);
    float wA = S.w[A][ln], wB = S.w[B][ln];
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
        S.vx[A][ln] = vA.x;
        S.vy[A][ln] = vA.y;
        S.w[A][ln] = wA;
    }
    if (kB != LK_WALL) {
        S.vx[B][ln] = vB.x;
        S.vy[B][ln] = vB.y;
        S.w[B][ln] = wB;
    }
}

// ContactSolver.solveVelocityConstraints for one constraint (= solveVelocityOne)
__device__ __forceinline__ void lkSolveVelocity(const PhysEnv& e, const LaneCtx& L, TC& c) {
    LaneSlotsU& S = *L.S;
    int ln = L.lane;
    int A = c.bodyA, B = c.bodyB;
    int kA = S.kind[A][ln], kB = S.kind[B][ln];
    V2 vA = mk(S.vx[A][ln], S.vy[A][ln]), vB = mk(S.vx[B][ln], S.vy[B][ln]);
    float wA = S.w[A][ln], wB = S.w[B][ln];
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
        S.vx[A][ln] = vA.x;
        S.vy[A][ln] = vA.y;
        S.w[A][ln] = wA;
    }
    if (kB != LK_WALL) {
        S.vx[B][ln] = vB.x;
        S.vy[B][ln] = vB.y;
        S.w[B][ln] = wB;
    }
}

// One position correction of one manifold point (= positionManifold + correctPoint). Returns the separation.
__device__ __forceinline__ float lkCorrectPoint(const PhysEnv& e, const LaneCtx& L, int A, int B, int type, V2 localNormal, V2 localPoint, V2 pointLocal,
                                                float radius, bool* changed) {
    LaneSlotsU& S = *L.S;
    int ln = L.lane;
    Xf xfA, xfB;
    xfA.px = S.px[A][ln];
    xfA.py = S.py[A][ln];
    xfA.c = S.c[A][ln];
    xfA.s = S.s[A][ln];
    xfB.px = S.px[B][ln];
    xfB.py = S.py[B][ln];
    xfB.c = S.c[B][ln];
    xfB.s = S.s[B][ln];
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
    V2 cA = mk(S.cx[A][ln], S.cy[A][ln]), cB = mk(S.cx[B][ln], S.cy[B][ln]);
    V2 rA = point - cA, rB = point - cB;
    float C = clampf(kContactBaumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
    if (C == 0.0f) return separation;
    int kA = S.kind[A][ln], kB = S.kind[B][ln];
    float invMassA = lkPosInvMass(e.hot, kA), invIA = lkPosInvI(e.hot, kA), invMassB = lkPosInvMass(e.hot, kB), invIB = lkPosInvI(e.hot, kB);
    float rnA = cross(rA, normal), rnB = cross(rB, normal);
    float K = invMassA + invMassB + invIA * rnA * rnA + invIB * rnB * rnB;
    float impulse = K > 0.0f ? -C / K : 0.0f;
    V2 P = impulse * normal;
    bool moved = false;
    {
        V2 nc = cA - invMassA * P;
        float oa = S.a[A][ln];
        float na = oa - invIA * cross(rA, P);
        if (kA != LK_WALL && (nc.x != cA.x || nc.y != cA.y || na != oa)) {
            S.cx[A][ln] = nc.x;
            S.cy[A][ln] = nc.y;
            S.a[A][ln] = na;
            lkSyncXf(e, L, A);
            moved = true;
        }
    }
    {
        V2 nc = cB + invMassB * P;
        float oa = S.a[B][ln];
        float na = oa + invIB * cross(rB, P);
        if (kB != LK_WALL && (nc.x != cB.x || nc.y != cB.y || na != oa)) {
            S.cx[B][ln] = nc.x;
            S.cy[B][ln] = nc.y;
            S.a[B][ln] = na;
            lkSyncXf(e, L, B);
            moved = true;
        }
    }
    if (moved) *changed = true;
    return separation;
}

// Island.solve step 6 for staged body l (= integrateBody); the start-of-step pose / transform go straight to HBM
__device__ __forceinline__ void lkIntegrate(const PhysEnv& e, const PhysWork& g, const LaneCtx& L, int l) {
    LaneSlotsU& S = *L.S;
    int ln = L.lane;
    int gb = S.gidx[l][ln];
    int P = e.cfg.d.P;
    float dt = e.cfg.dt;
    float vx = S.vx[l][ln], vy = S.vy[l][ln], w = S.w[l][ln];
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
    S.vx[l][ln] = vx;
    S.vy[l][ln] = vy;
    S.w[l][ln] = w;
    g.c0x[gb] = S.cx[l][ln];
    g.c0y[gb] = S.cy[l][ln];
    g.a0[gb] = S.a[l][ln];
    if (gb >= P) {
        int r = gb - P;
        g.r0px[r] = S.px[l][ln];
        g.r0py[r] = S.py[l][ln];
        g.r0c[r] = S.c[l][ln];
        g.r0s[r] = S.s[l][ln];
    }
    S.cx[l][ln] += dt * vx;
    S.cy[l][ln] += dt * vy;
    S.a[l][ln] += dt * w;
    lkSyncXf(e, L, l);
}

// Island.solve of up to 32 staged islands in lock step. `has`: this lane holds an island. Returns its position iterations.
__device__ __forceinline__ int lkSolve(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, const LaneCtx& L, bool has) {
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
        for (int l = 0; l < L.nb; ++l) lkIntegrate(e, wk, L, l);
    }
    __syncwarp();
    int iters = 0;
    bool running = has && e.cfg.posIters > 0;
    while (__any_sync(FULL, running)) {
        float minSep = 0.0f;
        bool changed = false;
        for (int i = 0; i < maxNc; ++i) {
            if (running && i < nc) {
                const TC& t = wk.tc[wk.order[L.lo + i]];
                for (int j = 0; j < t.solverPoints; ++j) {
                    float sep = lkCorrectPoint(e, L, t.bodyA, t.bodyB, t.type, t.localNormal, t.localPoint, t.lp[j], t.radius, &changed);
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

}  // namespace ps
#endif
