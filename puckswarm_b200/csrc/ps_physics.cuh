// puckswarm_b200 -- one JBox2D World.step(dt, velIters, posIters) of one arena, executed by one CTA.
//
// Replaces, for the PuckSwarm scene, org.jbox2d.dynamics.World.step as called from
// src/arena/RunOffline.java:94: ContactManager.collide, World.solve (islands by
// DFS, Island.solve, ContactSolver), Body.synchronizeFixtures + ContactManager.findNewContacts,
// World.solveTOI, clearForces. Results are meant to be bit-identical to the sequential reference:
//   * the contact cache is an array in creation order (JBox2D's world / per-body contact lists are that
//     order reversed), compacted stably, new pairs appended in sorted proxy-id order;
//   * islands and the Gauss-Seidel order inside them come from the same stack DFS (newest body first,
//     newest contact edge first) followed by Island.solve's static-last swap partition;
//   * islands are independent, so each island is solved by its own thread while the parallel phases
//     (narrow phase, constraint setup, integration, fixture sync, pair search, TOI per body) use the CTA.
#pragma once
#include "ps_ctx.cuh"
#include "ps_toi.cuh"

namespace ps {

struct PhysCfg {
    Dims d;
    int MC;          // contact cache capacity per arena
    int MT;          // touching-contact capacity per arena
    int MPAIR;       // new-pair buffer capacity (power of two)
    int MW;          // capacity of the wall-contact list
    int warpMinContacts;   // islands with at least this many contacts get a warp of their own (level-parallel sweeps)
    int laneXMaxContacts;  // ... unless they have at most this many and fit the wide lane slot (0: the wide lane slot is not used)
    int velIters, posIters, toiEnabled;
    float dt;
    float damp;      // clamp(1 - dt * 10, 0, 1): linear and angular damping factor of pucks and robots
};

// One touching contact of the current step: ContactConstraint + what the position solver needs.
struct TC {
    int bodyA, bodyB;     // dynamic body index, -1 = wall
    int ci;               // index in the contact cache
    int type;             // manifold type
    int solverPoints;     // ContactConstraint.pointCount (may drop to 1 by the condition-number test)
    int manifoldPoints;
    V2 normal, localNormal, localPoint;
    float radius;
    V2 lp[2], rA[2], rB[2];
    float nMass[2], tMass[2], nImp[2], tImp[2];
    float k11, k12, k22, m11, m12, m22;   // K and normalMass = K^-1 (both symmetric)
    uint32_t key;         // (proxyA << 16) | proxyB
};

// Per-arena persistent state in global memory (layout = ps_state_view of one arena).
struct ArenaState {
    float* body;          // [6][NB] c.x c.y a v.x v.y w
    float* senseAabb;     // [4][NB]
    float* fat;           // [4][NP]
    uint32_t* ckey;       // [MC]
    uint32_t* cinfo;      // [MC]
    float* cimp;          // [MC][4]
    int* nContacts;
    float* invDt0;
    int* errorFlags;
};

// Per-CTA working memory. Pointers may refer to shared or global memory.
struct PhysWork {
    float *cx, *cy, *a, *vx, *vy, *w, *c0x, *c0y, *a0;   // [NB]
    float *rpx, *rpy, *rc, *rs;                          // [R] robot transform (pucks: position = c, R unused)
    float *r0px, *r0py, *r0c, *r0s;                      // [R] robot transform at the start of the step (xf1)
    float *fx, *fy, *tq;                                 // [R] force / torque accumulators
    int* adjOff;                                         // [NB + 1]
    int* adjFill;                                        // [NB]
    uint16_t* adj;                                       // [2 * MT] touching contacts per body, newest first
    uint16_t* order;                                     // [MT] Gauss-Seidel order, island after island
    int* islandStart;                                    // [NB + 1]
    int* islandBodyStart;                                // [NB + 1] ranges into ibody, one per recorded island
    uint16_t* ibody;                                     // [NB] bodies of the recorded islands (DFS pop order)
    uint16_t* stack;                                     // [NB]
    uint8_t* bodyFlag;                                   // [NB]
    uint8_t* tcFlag;                                     // [MT]
    uint32_t* tcPair;                                    // [MT] (bodyA + 1) << 16 | (bodyB + 1) of each touching contact (0 = wall)
    uint8_t* moved;                                      // [NP]
    float* bodyFat;                                      // [4][NB] union of the body's fat AABBs
    int* movedList;                                      // [NP]
    float* fatNow;                                       // [4][NP] the fat AABBs of this step for the pair search: a shared-memory copy, or null = st.fat
    uint16_t* movedBody;                                 // [NB] bodies with at least one moved proxy
    uint32_t* bodyPairs;                                 // [4 * NB] (moved body << 16 | body) whose box unions overlap
    float* fatOld;                                       // [4][NP] fat AABB before this step's moveProxy
    uint32_t* pairs;                                     // [MPAIR]
    int* wallList;                                       // [MW] cached contacts with the wall
    int* wallBody;                                       // [MW] the dynamic body of each wall contact
    float* wallT;                                        // [MW] speculative TimeOfImpact result of the current round
    uint8_t* wallState;                                  // [MW] TOI state of the current round
    uint8_t* wallNeed;                                   // [MW] needs (re-)evaluation with the body's current tMax
    float* toiT;                                         // [NB] World.solveTOI(Body): current minimum toi
    int* toiContact;                                     // [NB] cache index of the current toiContact, -1 = none
    int* toiCur;                                         // [NB] scan position in the wall list (descending)
    int* toiCount;                                       // [NB] count / found / iter of the current do-while pass
    uint8_t* toiFound;
    uint8_t* toiIter;
    uint8_t* toiActive;
    TC* tc;                                              // [MT]
    int* counters;                                       // [8]: 0 nTouching, 1 nIslands, 2 nMoved, 3 nPairs, 4 nWall, 5 scratch
    int* stats;                                          // [8] per-step diagnostics: 0 max position iterations, 1 TOI events, 2 islands
};

enum { ERR_CONTACT_CAP = 1, ERR_LM_CAP = 2, ERR_TOUCH_CAP = 4, ERR_PAIR_CAP = 8, ERR_WALL_CAP = 16 };

// The handful of scene constants the solver loops read on every step, carried by value (kernel parameter /
// constant bank) so that the Gauss-Seidel dependency chain does not wait on loads through the scene pointer.
struct HotConsts {
    BodyConst puck, robot;
    Xf wallXf;
    float friction;
    float pmPuck, piPuck, pmRobot, piRobot;   // mass-scaled inverse masses of the position solver: mass * invMass, mass * invI
};
struct PhysEnv {
    const Scene* scene;
    Trig trig;
    PhysCfg cfg;
    HotConsts hot;
};

// Phase clock for the per-step diagnostics (stats[3..7]: cycles of collide, islands, velocity, position, broadphase + TOI)
PS_HD long long phaseClock() {
#if defined(__CUDA_ARCH__) && defined(PS_PHASE_CLOCKS)
    return clock64();
#else
    return 0;
#endif
}

PS_HD bool isRobot(const Dims& d, int b) { return b >= d.P; }

PS_HD Xf bodyXf(const PhysEnv& e, const PhysWork& wk, int b) {
    Xf T;
    if (b < 0) {
        T = e.hot.wallXf;
    } else if (b < e.cfg.d.P) {
        T.px = wk.cx[b];
        T.py = wk.cy[b];
        T.c = 1.0f;
        T.s = 0.0f;
    } else {
        int r = b - e.cfg.d.P;
        T.px = wk.rpx[r];
        T.py = wk.rpy[r];
        T.c = wk.rc[r];
        T.s = wk.rs[r];
    }
    return T;
}
PS_HD Xf bodyXf0(const PhysEnv& e, const PhysWork& wk, int b) {
    Xf T;
    if (b < e.cfg.d.P) {
        T.px = wk.c0x[b];
        T.py = wk.c0y[b];
        T.c = 1.0f;
        T.s = 0.0f;
    } else {
        int r = b - e.cfg.d.P;
        T.px = wk.r0px[r];
        T.py = wk.r0py[r];
        T.c = wk.r0c[r];
        T.s = wk.r0s[r];
    }
    return T;
}
// Body.synchronizeTransform (only robots carry a rotation that anything reads)
PS_HD void syncBodyXf(const PhysEnv& e, const PhysWork& wk, int b) {
    if (b >= e.cfg.d.P) {
        int r = b - e.cfg.d.P;
        Xf T = makeXf(e.trig, wk.cx[b], wk.cy[b], wk.a[b], e.hot.robot.lcx, e.hot.robot.lcy);
        wk.rpx[r] = T.px;
        wk.rpy[r] = T.py;
        wk.rc[r] = T.c;
        wk.rs[r] = T.s;
    }
}
PS_HD const BodyConst& bodyConst(const PhysEnv& e, int b) { return b < e.cfg.d.P ? e.hot.puck : e.hot.robot; }

PS_HD Aabb loadAabb(const float* base, int n, int i) {
    Aabb a;
    a.lx = base[i];
    a.ly = base[n + i];
    a.ux = base[2 * n + i];
    a.uy = base[3 * n + i];
    return a;
}
PS_HD void storeAabb(float* base, int n, int i, const Aabb& a) {
    base[i] = a.lx;
    base[n + i] = a.ly;
    base[2 * n + i] = a.ux;
    base[3 * n + i] = a.uy;
}
PS_HD Aabb proxyFat(const PhysEnv& e, const ArenaState& st, int proxy) {
    if (proxy < kWallFixtures) return e.scene->wallFat[proxy];
    return loadAabb(st.fat, e.cfg.d.NP, proxy - kWallFixtures);
}
PS_HD const float* fatNowOf(const PhysWork& wk, const ArenaState& st) { return wk.fatNow ? wk.fatNow : st.fat; }

// Contact.update: evaluate the manifold at the current transforms and carry the impulses of points whose id survived.
PS_HD void updateContact(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, int ci, Manifold& m, float* imp /*[4]*/) {
    uint32_t key = st.ckey[ci];
    int pa = (int)(key >> 16), pb = (int)(key & 0xFFFFu);
    const Shape& sA = e.scene->shapes[proxyShape(e.cfg.d, pa)];
    const Shape& sB = e.scene->shapes[proxyShape(e.cfg.d, pb)];
    Xf xfA = bodyXf(e, wk, proxyBody(e.cfg.d, pa)), xfB = bodyXf(e, wk, proxyBody(e.cfg.d, pb));
    evaluateManifold(m, sA, xfA, sB, xfB);
    uint32_t oldInfo = st.cinfo[ci];
    int oldCount = (int)(oldInfo & 3u);
    float o[4];
    o[0] = st.cimp[4 * ci + 0];
    o[1] = st.cimp[4 * ci + 1];
    o[2] = st.cimp[4 * ci + 2];
    o[3] = st.cimp[4 * ci + 3];
    uint32_t info = (uint32_t)m.pointCount;
    imp[0] = imp[1] = imp[2] = imp[3] = 0.0f;
    for (int i = 0; i < m.pointCount; ++i) {
        uint32_t id2 = m.id[i];
        info |= id2 << (8 + 8 * i);
        for (int j = 0; j < oldCount; ++j) {
            uint32_t id1 = (oldInfo >> (8 + 8 * j)) & 0xFFu;
            if (id1 == id2) {
                imp[2 * i] = o[2 * j];
                imp[2 * i + 1] = o[2 * j + 1];
                break;
            }
        }
    }
    st.cinfo[ci] = info;
    st.cimp[4 * ci + 0] = imp[0];
    st.cimp[4 * ci + 1] = imp[1];
    st.cimp[4 * ci + 2] = imp[2];
    st.cimp[4 * ci + 3] = imp[3];
}

// ContactSolver constructor for one touching contact
PS_HD void initConstraint(const PhysEnv& e, const PhysWork& wk, TC& c, const Manifold& m, int pa, int pb, const float* imp, float dtRatio) {
    const Dims& d = e.cfg.d;
    int bA = proxyBody(d, pa), bB = proxyBody(d, pb);
    const Shape& sA = e.scene->shapes[proxyShape(d, pa)];
    const Shape& sB = e.scene->shapes[proxyShape(d, pb)];
    Xf xfA = bodyXf(e, wk, bA), xfB = bodyXf(e, wk, bB);
    float invMassA = 0, invIA = 0, invMassB = 0, invIB = 0;
    V2 cA = mk(0, 0), cB = mk(0, 0);
    if (bA >= 0) {
        const BodyConst& k = bodyConst(e, bA);
        invMassA = k.invMass;
        invIA = k.invI;
        cA = mk(wk.cx[bA], wk.cy[bA]);
    }
    if (bB >= 0) {
        const BodyConst& k = bodyConst(e, bB);
        invMassB = k.invMass;
        invIB = k.invI;
        cB = mk(wk.cx[bB], wk.cy[bB]);
    }
    c.bodyA = bA;
    c.bodyB = bB;
    c.key = ((uint32_t)pa << 16) | (uint32_t)pb;
    c.type = m.type;
    c.manifoldPoints = m.pointCount;
    c.solverPoints = m.pointCount;
    c.localNormal = m.localNormal;
    c.localPoint = m.localPoint;
    c.radius = sA.radius + sB.radius;
    V2 pts[2];
    worldManifold(m, xfA, sA.radius, xfB, sB.radius, c.normal, pts);
    for (int j = 0; j < m.pointCount; ++j) {
        c.nImp[j] = dtRatio * imp[2 * j];
        c.tImp[j] = dtRatio * imp[2 * j + 1];
        c.lp[j] = m.p[j];
        c.rA[j] = pts[j] - cA;
        c.rB[j] = pts[j] - cB;
        float rnA = cross(c.rA[j], c.normal), rnB = cross(c.rB[j], c.normal);
        rnA *= rnA;
        rnB *= rnB;
        float kNormal = invMassA + invMassB + invIA * rnA + invIB * rnB;
        c.nMass[j] = 1.0f / kNormal;
        V2 tangent = crossVS(c.normal, 1.0f);
        float rtA = cross(c.rA[j], tangent), rtB = cross(c.rB[j], tangent);
        rtA *= rtA;
        rtB *= rtB;
        float kTangent = invMassA + invMassB + invIA * rtA + invIB * rtB;
        c.tMass[j] = 1.0f / kTangent;
        // velocityBias = -restitution * vRel with restitution 0 everywhere: always (+-)0, never changes vn - bias
    }
    if (m.pointCount == 2) {
        float rn1A = cross(c.rA[0], c.normal), rn1B = cross(c.rB[0], c.normal);
        float rn2A = cross(c.rA[1], c.normal), rn2B = cross(c.rB[1], c.normal);
        float k11 = invMassA + invMassB + invIA * rn1A * rn1A + invIB * rn1B * rn1B;
        float k22 = invMassA + invMassB + invIA * rn2A * rn2A + invIB * rn2B * rn2B;
        float k12 = invMassA + invMassB + invIA * rn1A * rn2A + invIB * rn1B * rn2B;
        if (k11 * k11 < 100.0f * (k11 * k22 - k12 * k12)) {
            c.k11 = k11;
            c.k12 = k12;
            c.k22 = k22;
            // Mat22.getInverse of [[k11 k12][k12 k22]]
            float det = k11 * k22 - k12 * k12;
            if (det != 0.0f) det = 1.0f / det;
            c.m11 = det * k22;
            c.m12 = -det * k12;
            c.m22 = det * k11;
        } else {
            c.solverPoints = 1;
        }
    }
}

struct BodyVel {
    V2 v;
    float w, invMass, invI;
};
PS_HD BodyVel loadVel(const PhysEnv& e, const PhysWork& wk, int b) {
    BodyVel r;
    if (b < 0) {
        r.v = mk(0, 0);
        r.w = 0;
        r.invMass = 0;
        r.invI = 0;
    } else {
        const BodyConst& k = bodyConst(e, b);
        r.v = mk(wk.vx[b], wk.vy[b]);
        r.w = wk.w[b];
        r.invMass = k.invMass;
        r.invI = k.invI;
    }
    return r;
}
PS_HD void storeVel(const PhysWork& wk, int b, const BodyVel& r) {
    if (b >= 0) {
        wk.vx[b] = r.v.x;
        wk.vy[b] = r.v.y;
        wk.w[b] = r.w;
    }
}

// ContactSolver.warmStart for one constraint
PS_HD void warmStartOne(const PhysEnv& e, const PhysWork& wk, const TC& c) {
    BodyVel A = loadVel(e, wk, c.bodyA), B = loadVel(e, wk, c.bodyB);
    V2 normal = c.normal, tangent = crossVS(normal, 1.0f);
    for (int j = 0; j < c.solverPoints; ++j) {
        V2 P = c.nImp[j] * normal + c.tImp[j] * tangent;
        A.w -= A.invI * cross(c.rA[j], P);
        A.v = A.v - A.invMass * P;
        B.w += B.invI * cross(c.rB[j], P);
        B.v = B.v + B.invMass * P;
    }
    storeVel(wk, c.bodyA, A);
    storeVel(wk, c.bodyB, B);
}

// ContactSolver.solveVelocityConstraints for one constraint (friction first, then normal / block solver)
PS_HD void solveVelocityOne(const PhysEnv& e, const PhysWork& wk, TC& c) {
    BodyVel A = loadVel(e, wk, c.bodyA), B = loadVel(e, wk, c.bodyB);
    V2 vA = A.v, vB = B.v;
    float wA = A.w, wB = B.w;
    float invMassA = A.invMass, invIA = A.invI, invMassB = B.invMass, invIB = B.invI;
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
        // b -= K * a
        b = b - mk(c.k11 * a.x + c.k12 * a.y, c.k12 * a.x + c.k22 * a.y);
        V2 x;
        bool ok = false;
        // case 1: both points active, x = -K^-1 b
        V2 mb = mk(c.m11 * b.x + c.m12 * b.y, c.m12 * b.x + c.m22 * b.y);
        x = -mb;
        if (x.x >= 0.0f && x.y >= 0.0f) ok = true;
        if (!ok) {   // case 2: point 1 active
            x.x = -c.nMass[0] * b.x;
            x.y = 0.0f;
            vn2 = c.k12 * x.x + b.y;
            if (x.x >= 0.0f && vn2 >= 0.0f) ok = true;
        }
        if (!ok) {   // case 3: point 2 active
            x.x = 0.0f;
            x.y = -c.nMass[1] * b.y;
            vn1 = c.k12 * x.y + b.x;
            if (x.y >= 0.0f && vn1 >= 0.0f) ok = true;
        }
        if (!ok) {   // case 4: none active
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
    A.v = vA;
    A.w = wA;
    B.v = vB;
    B.w = wB;
    storeVel(wk, c.bodyA, A);
    storeVel(wk, c.bodyB, B);
}

// One position-correction of one manifold point; massA/massB let the TOI solver freeze one side.
// Returns the separation it saw; *changed is set when a body's (c, a) actually moved (bitwise).
PS_HD float correctPoint(const PhysEnv& e, const PhysWork& wk, int bA, int bB, int type, V2 localNormal, V2 localPoint, V2 pointLocal,
                         float radius, float baumgarte, float invMassA, float invIA, float invMassB, float invIB, bool* changed = nullptr) {
    Xf xfA = bodyXf(e, wk, bA), xfB = bodyXf(e, wk, bB);
    V2 normal, point;
    float separation;
    positionManifold(type, localNormal, localPoint, pointLocal, radius, xfA, xfB, normal, point, separation);
    V2 cA = bA >= 0 ? mk(wk.cx[bA], wk.cy[bA]) : mk(0, 0);
    V2 cB = bB >= 0 ? mk(wk.cx[bB], wk.cy[bB]) : mk(0, 0);
    V2 rA = point - cA, rB = point - cB;
    float C = clampf(baumgarte * (separation + kLinearSlop), -kMaxLinearCorrection, 0.0f);
    // C == 0 (no penetration beyond the slop): the impulse is -0 / K = (-)0, P = 0 and neither body moves
    if (C == 0.0f) return separation;
    float rnA = cross(rA, normal), rnB = cross(rB, normal);
    float K = invMassA + invMassB + invIA * rnA * rnA + invIB * rnB * rnB;
    float impulse = K > 0.0f ? -C / K : 0.0f;
    V2 P = impulse * normal;
    bool moved = false;
    if (bA >= 0) {
        V2 nc = cA - invMassA * P;
        float oa = wk.a[bA];
        float na = oa - invIA * cross(rA, P);
        if (nc.x != cA.x || nc.y != cA.y || na != oa) {
            // (an unchanged body keeps its transform: synchronizeTransform is a pure function of (c, a))
            wk.cx[bA] = nc.x;
            wk.cy[bA] = nc.y;
            wk.a[bA] = na;
            syncBodyXf(e, wk, bA);
            moved = true;
        }
    }
    if (bB >= 0) {
        V2 nc = cB + invMassB * P;
        float oa = wk.a[bB];
        float na = oa + invIB * cross(rB, P);
        if (nc.x != cB.x || nc.y != cB.y || na != oa) {
            wk.cx[bB] = nc.x;
            wk.cy[bB] = nc.y;
            wk.a[bB] = na;
            syncBodyXf(e, wk, bB);
            moved = true;
        }
    }
    if (changed && moved) *changed = true;
    return separation;
}

// ---------------------------------------------------------------------------------------------------------------
// Phase A: load state into working memory, apply the robots' commands (Robot.move, src/arena/Robot.java:175-182)
template <class Ctx>
PS_HD void physLoad(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st, const float* cmdFwd, const float* cmdTorque) {
    const Dims& d = e.cfg.d;
    int NB = d.NB;
    PS_FOR(b, NB) {
        wk.cx[b] = st.body[0 * NB + b];
        wk.cy[b] = st.body[1 * NB + b];
        wk.a[b] = st.body[2 * NB + b];
        wk.vx[b] = st.body[3 * NB + b];
        wk.vy[b] = st.body[4 * NB + b];
        wk.w[b] = st.body[5 * NB + b];
        syncBodyXf(e, wk, b);
        if (b >= d.P) {
            int r = b - d.P;
            Xf T = bodyXf(e, wk, b);
            // f = getWorldVector((forwards, 0)); p = getWorldPoint(localCenter); applyForce(f, p); applyTorque(t)
            V2 f = mulR(T, mk(cmdFwd[r], 0.0f));
            V2 p = mulX(T, mk(e.scene->robot.lcx, e.scene->robot.lcy));
            float fx = 0.0f, fy = 0.0f, tq = 0.0f;
            fx += f.x;
            fy += f.y;
            tq += cross(p - mk(wk.cx[b], wk.cy[b]), f);
            tq += cmdTorque[r];
            wk.fx[r] = fx;
            wk.fy[r] = fy;
            wk.tq[r] = tq;
        }
    }
    if (ctx.tid() == 0) {
        for (int i = 0; i < 8; i++) wk.counters[i] = 0;
        for (int i = 0; i < 8; i++) wk.stats[i] = 0;
    }
    ctx.sync();
}

// Phase B: ContactManager.collide -- destroy contacts whose fat AABBs separated, re-evaluate the others, compact the
// cache stably, and set up the solver constraint of every touching contact.
template <class Ctx>
PS_HD void physCollide(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st, float dtRatio) {
    const Dims& d = e.cfg.d;
    int n = *st.nContacts;
    int nKeep = 0, nTouch = 0;
    int nRound = roundUpThreads(ctx, n);
    for (int base = 0; base < nRound; base += ctx.nthreads()) {
        int i = base + ctx.tid();
        bool keep = false, touching = false;
        uint32_t key = 0, info = 0;
        float imp[4] = {0, 0, 0, 0};
        Manifold m;
        m.pointCount = 0;
        if (i < n) {
            key = st.ckey[i];
            int pa = (int)(key >> 16), pb = (int)(key & 0xFFFFu);
            keep = aabbOverlap(proxyFat(e, st, pa), proxyFat(e, st, pb));
            if (keep) {
                updateContact(e, wk, st, i, m, imp);
                info = st.cinfo[i];
                touching = m.pointCount > 0;
            }
        }
        int totalKeep, totalTouch;
        int posKeep = nKeep + ctx.exscanFlag(keep, totalKeep);
        int posTouch = nTouch + ctx.exscanFlag(touching, totalTouch);
        // (the scans contain barriers: every read of this chunk is done before any write below)
        if (keep) {
            st.ckey[posKeep] = key;
            st.cinfo[posKeep] = info;
            st.cimp[4 * posKeep + 0] = imp[0];
            st.cimp[4 * posKeep + 1] = imp[1];
            st.cimp[4 * posKeep + 2] = imp[2];
            st.cimp[4 * posKeep + 3] = imp[3];
            if (touching) {
                if (posTouch < e.cfg.MT) {
                    TC c;
                    c.ci = posKeep;
                    initConstraint(e, wk, c, m, (int)(key >> 16), (int)(key & 0xFFFFu), imp, dtRatio);
                    wk.tc[posTouch] = c;
                } else {
                    *st.errorFlags |= ERR_TOUCH_CAP;
                }
            }
        }
        nKeep += totalKeep;
        nTouch += totalTouch;
        ctx.sync();
    }
    if (nTouch > e.cfg.MT) nTouch = e.cfg.MT;
    if (ctx.tid() == 0) {
        *st.nContacts = nKeep;
        wk.counters[0] = nTouch;
    }
    ctx.sync();
}

// Phase C: World.solve -- contact graph, islands by DFS, integrate velocities.
template <class Ctx>
PS_HD void physIslands(Ctx& ctx, const PhysEnv& e, const PhysWork& wk) {
    const Dims& d = e.cfg.d;
    int NB = d.NB;
    int nT = wk.counters[0];
    PS_FOR(b, NB + 1) {
        wk.adjOff[b] = 0;
        if (b < NB) {
            wk.adjFill[b] = 0;
            wk.bodyFlag[b] = 0;
        }
    }
    PS_FOR(t, nT) {
        wk.tcFlag[t] = 0;
        wk.tcPair[t] = ((uint32_t)(wk.tc[t].bodyA + 1) << 16) | (uint32_t)(wk.tc[t].bodyB + 1);
    }
    ctx.sync();
    // degree of every dynamic body in the touching-contact graph
    PS_FOR(t, nT) {
        int bA = (int)(wk.tcPair[t] >> 16) - 1, bB = (int)(wk.tcPair[t] & 0xFFFFu) - 1;
        if (bA >= 0) ctx.atomicAddI(&wk.adjOff[bA + 1], 1);
        if (bB >= 0) ctx.atomicAddI(&wk.adjOff[bB + 1], 1);
    }
    ctx.sync();
    // exclusive prefix over bodies (adjOff[b + 1] currently holds deg(b))
    {
        int carry = 0;
        int nRound = roundUpThreads(ctx, NB);
        for (int base = 0; base < nRound; base += ctx.nthreads()) {
            int b = base + ctx.tid();
            int deg = b < NB ? wk.adjOff[b + 1] : 0;
            int total;
            int off = carry + ctx.exscanInt(deg, total);
            ctx.sync();
            if (b < NB) wk.adjOff[b + 1] = off + deg;
            carry += total;
            ctx.sync();
        }
    }
    // fill, then order every body's list newest contact first (= JBox2D's per-body contact edge list)
    PS_FOR(t, nT) {
        int bA = (int)(wk.tcPair[t] >> 16) - 1, bB = (int)(wk.tcPair[t] & 0xFFFFu) - 1;
        if (bA >= 0) wk.adj[wk.adjOff[bA] + ctx.atomicAddI(&wk.adjFill[bA], 1)] = (uint16_t)t;
        if (bB >= 0) wk.adj[wk.adjOff[bB] + ctx.atomicAddI(&wk.adjFill[bB], 1)] = (uint16_t)t;
    }
    ctx.sync();
    PS_FOR(b, NB) {
        int lo = wk.adjOff[b], hi = wk.adjOff[b + 1];
        for (int i = lo + 1; i < hi; ++i) {   // insertion sort, descending
            uint16_t v = wk.adj[i];
            int j = i - 1;
            while (j >= lo && wk.adj[j] < v) {
                wk.adj[j + 1] = wk.adj[j];
                --j;
            }
            wk.adj[j + 1] = v;
        }
    }
    // integrate velocities (Island.solve step 1) -- independent of the island structure
    PS_FOR(b, NB) {
        const BodyConst& k = bodyConst(e, b);
        float fx = 0.0f, fy = 0.0f, tq = 0.0f;
        if (b >= d.P) {
            fx = wk.fx[b - d.P];
            fy = wk.fy[b - d.P];
            tq = wk.tq[b - d.P];
        }
        float dt = e.cfg.dt;
        // v += dt * (gravity + invMass * force), gravity = (0, 0)
        float vx = wk.vx[b] + dt * (0.0f + k.invMass * fx);
        float vy = wk.vy[b] + dt * (0.0f + k.invMass * fy);
        float w = wk.w[b] + dt * k.invI * tq;
        wk.vx[b] = vx * e.cfg.damp;
        wk.vy[b] = vy * e.cfg.damp;
        wk.w[b] = w * e.cfg.damp;
    }
    ctx.sync();
    // stack DFS: seeds in body-list order (newest body first), edges newest first; only islands with contacts are recorded
    if (ctx.tid() == 0) {
        int nIsl = 0, ordN = 0, nIb = 0;
        for (int seed = NB - 1; seed >= 0; --seed) {
            if (wk.bodyFlag[seed]) continue;
            wk.bodyFlag[seed] = 1;
            if (wk.adjOff[seed] == wk.adjOff[seed + 1]) continue;
            int start = ordN;
            wk.islandStart[nIsl] = start;
            wk.islandBodyStart[nIsl] = nIb;
            int sp = 0;
            wk.stack[sp++] = (uint16_t)seed;
            while (sp > 0) {
                int b = wk.stack[--sp];
                wk.ibody[nIb++] = (uint16_t)b;
                for (int i = wk.adjOff[b]; i < wk.adjOff[b + 1]; ++i) {
                    int t = wk.adj[i];
                    if (wk.tcFlag[t]) continue;
                    wk.tcFlag[t] = 1;
                    wk.order[ordN++] = (uint16_t)t;
                    int bA = (int)(wk.tcPair[t] >> 16) - 1, bB = (int)(wk.tcPair[t] & 0xFFFFu) - 1;
                    int other = bA == b ? bB : bA;
                    if (other < 0) continue;   // the wall is static: part of the island, never expanded
                    if (wk.bodyFlag[other]) continue;
                    wk.stack[sp++] = (uint16_t)other;
                    wk.bodyFlag[other] = 1;
                }
            }
            // Island.solve: contacts between two non-static bodies first (swap partition)
            int i1 = start - 1;
            for (int i2 = start; i2 < ordN; ++i2) {
                int t = wk.order[i2];
                bool nonStatic = (wk.tcPair[t] >> 16) != 0 && (wk.tcPair[t] & 0xFFFFu) != 0;
                if (nonStatic) {
                    ++i1;
                    uint16_t tmp = wk.order[i1];
                    wk.order[i1] = wk.order[i2];
                    wk.order[i2] = tmp;
                }
            }
            ++nIsl;
        }
        wk.islandStart[nIsl] = ordN;
        wk.islandBodyStart[nIsl] = nIb;
        wk.counters[1] = nIsl;
    }
    ctx.sync();
}

// Phase D: ContactSolver velocity phase per island (warm start, velIters sweeps, store impulses)
template <class Ctx>
PS_HD void physSolveVelocity(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st) {
    int nIsl = wk.counters[1];
    PS_FOR(k, nIsl) {
        int lo = wk.islandStart[k], hi = wk.islandStart[k + 1];
        for (int i = lo; i < hi; ++i) warmStartOne(e, wk, wk.tc[wk.order[i]]);
        for (int it = 0; it < e.cfg.velIters; ++it)
            for (int i = lo; i < hi; ++i) solveVelocityOne(e, wk, wk.tc[wk.order[i]]);
        // ContactSolver.storeImpulses
        for (int i = lo; i < hi; ++i) {
            const TC& c = wk.tc[wk.order[i]];
            for (int j = 0; j < c.solverPoints; ++j) {
                st.cimp[4 * c.ci + 2 * j] = c.nImp[j];
                st.cimp[4 * c.ci + 2 * j + 1] = c.tImp[j];
            }
        }
    }
    ctx.sync();
}

// Island.solve step 6 for one body: clamp the step, remember the start-of-step transform, advance, synchronizeTransform
PS_HD void integrateBody(const PhysEnv& e, const PhysWork& wk, int b) {
    const Dims& d = e.cfg.d;
    float dt = e.cfg.dt;
    float vx = wk.vx[b], vy = wk.vy[b], w = wk.w[b];
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
    wk.vx[b] = vx;
    wk.vy[b] = vy;
    wk.w[b] = w;
    wk.c0x[b] = wk.cx[b];
    wk.c0y[b] = wk.cy[b];
    wk.a0[b] = wk.a[b];
    if (b >= d.P) {
        int r = b - d.P;
        wk.r0px[r] = wk.rpx[r];
        wk.r0py[r] = wk.rpy[r];
        wk.r0c[r] = wk.rc[r];
        wk.r0s[r] = wk.rs[r];
    }
    wk.cx[b] += dt * vx;
    wk.cy[b] += dt * vy;
    wk.a[b] += dt * w;
    syncBodyXf(e, wk, b);
}

// Phase E: integrate positions (Island.solve step 6) of every body (fused pipeline)
template <class Ctx>
PS_HD void physIntegrate(Ctx& ctx, const PhysEnv& e, const PhysWork& wk) {
    PS_FOR(b, e.cfg.d.NB) integrateBody(e, wk, b);
    ctx.sync();
}

// Phase F: ContactSolver.solvePositionConstraints per island until minSeparation >= -1.5 linearSlop
template <class Ctx>
PS_HD void physSolvePosition(Ctx& ctx, const PhysEnv& e, const PhysWork& wk) {
    int nIsl = wk.counters[1];
    const BodyConst &kp = e.scene->puck, &kr = e.scene->robot;
    // mass-scaled inverse masses: invMass = mass * invMass, invI = mass * invI
    float pmP = kp.mass * kp.invMass, piP = kp.mass * kp.invI, pmR = kr.mass * kr.invMass, piR = kr.mass * kr.invI;
    int P = e.cfg.d.P;
    PS_FOR(k, nIsl) {
        int lo = wk.islandStart[k], hi = wk.islandStart[k + 1];
        int iters = 0;
        for (int it = 0; it < e.cfg.posIters; ++it) {
            ++iters;
            float minSeparation = 0.0f;
            for (int i = lo; i < hi; ++i) {
                const TC& c = wk.tc[wk.order[i]];
                int bA = c.bodyA, bB = c.bodyB;
                float imA = bA < 0 ? 0.0f : (bA < P ? pmP : pmR), iiA = bA < 0 ? 0.0f : (bA < P ? piP : piR);
                float imB = bB < 0 ? 0.0f : (bB < P ? pmP : pmR), iiB = bB < 0 ? 0.0f : (bB < P ? piP : piR);
                for (int j = 0; j < c.solverPoints; ++j) {
                    float sep = correctPoint(e, wk, bA, bB, c.type, c.localNormal, c.localPoint, c.lp[j], c.radius, kContactBaumgarte, imA, iiA, imB, iiB);
                    minSeparation = fminf_(minSeparation, sep);
                }
            }
            if (minSeparation >= -1.5f * kLinearSlop) break;
        }
        ctx.atomicMaxI(&wk.stats[0], iters);
    }
    ctx.sync();
    if (ctx.tid() == 0) {
        wk.stats[2] = nIsl;
        // islands without contacts also run the position loop once (it breaks immediately)
        if (wk.stats[0] < 1 && e.cfg.posIters > 0) wk.stats[0] = 1;
    }
    ctx.sync();
}

// Phase G: Body.synchronizeFixtures for every dynamic body: swept AABBs, DynamicTree.moveProxy
template <class Ctx>
PS_HD void physSyncFixtures(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st) {
    const Dims& d = e.cfg.d;
    int NP = d.NP, NB = d.NB;
    int nMoved = 0;
    int nRound = roundUpThreads(ctx, NP);
    for (int base = 0; base < nRound; base += ctx.nthreads()) {
        int i = base + ctx.tid();
        bool mv = false;
        if (i < NP) {
            int proxy = kWallFixtures + i;
            int b = proxyBody(d, proxy);
            const Shape& s = e.scene->shapes[proxyShape(d, proxy)];
            Xf xf1 = bodyXf0(e, wk, b), xf2 = bodyXf(e, wk, b);
            Aabb a1 = shapeAabb(s, xf1), a2 = shapeAabb(s, xf2);
            Aabb u;
            u.lx = fminf_(a1.lx, a2.lx);
            u.ly = fminf_(a1.ly, a2.ly);
            u.ux = fmaxf_(a1.ux, a2.ux);
            u.uy = fmaxf_(a1.uy, a2.uy);
            // Fixture.m_aabb of the fixture PointSensor can see: robot hull (fixture 0), puck inner disc (fixture 1)
            bool sensed = b < d.P ? (proxy - bodyFirstProxy(d, b)) == 1 : (proxy - bodyFirstProxy(d, b)) == 0;
            if (sensed) storeAabb(st.senseAabb, NB, b, u);
            Aabb fat = loadAabb(st.fat, NP, i);
            if (!aabbContains(fat, u)) {
                storeAabb(wk.fatOld, NP, i, fat);
                float dx = xf2.px - xf1.px, dy = xf2.py - xf1.py;
                Aabb nb;
                nb.lx = u.lx - kAabbExtension;
                nb.ly = u.ly - kAabbExtension;
                nb.ux = u.ux + kAabbExtension;
                nb.uy = u.uy + kAabbExtension;
                float ddx = kAabbMultiplier * dx, ddy = kAabbMultiplier * dy;
                if (ddx < 0.0f) nb.lx += ddx; else nb.ux += ddx;
                if (ddy < 0.0f) nb.ly += ddy; else nb.uy += ddy;
                storeAabb(st.fat, NP, i, nb);
                fat = nb;
                mv = true;
            }
            if (wk.fatNow) storeAabb(wk.fatNow, NP, i, fat);
            wk.moved[i] = mv ? 1 : 0;
        }
        int total;
        int pos = nMoved + ctx.exscanFlag(mv, total);
        if (mv) wk.movedList[pos] = i;
        nMoved += total;
    }
    if (ctx.tid() == 0) wk.counters[2] = nMoved;
    ctx.sync();
}

// Union of each body's fat AABBs, and the list of bodies with at least one moved proxy (pair-search culling)
PS_HD int bodyPairCap(const Dims& d) { return 4 * d.NB; }
template <class Ctx>
PS_HD void physBodyFat(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st) {
    const Dims& d = e.cfg.d;
    if (ctx.tid() == 0) {
        wk.counters[6] = 0;
        wk.counters[7] = 0;
    }
    ctx.sync();
    PS_FOR(b, d.NB) {
        int first = bodyFirstProxy(d, b) - kWallFixtures, cnt = bodyProxyCount(d, b);
        const float* fatNow = fatNowOf(wk, st);
        Aabb u = loadAabb(fatNow, d.NP, first);
        bool mv = wk.moved[first] != 0;
        for (int k = 1; k < cnt; ++k) {
            Aabb f = loadAabb(fatNow, d.NP, first + k);
            u.lx = fminf_(u.lx, f.lx);
            u.ly = fminf_(u.ly, f.ly);
            u.ux = fmaxf_(u.ux, f.ux);
            u.uy = fmaxf_(u.uy, f.uy);
            mv = mv || wk.moved[first + k] != 0;
        }
        storeAabb(wk.bodyFat, d.NB, b, u);
        if (mv) wk.movedBody[ctx.atomicAddI(&wk.counters[6], 1)] = (uint16_t)b;
    }
    ctx.sync();
}

// Candidate test of BroadPhase.updatePairs + ContactManager.addPair for (moved proxy q, other proxy p).
// A pair is new iff the fat AABBs overlap now and did not overlap before this step's moveProxy calls: the contact cache
// always equals the set of filtered pairs with overlapping fat AABBs (contacts die in collide() exactly when they stop
// overlapping, and are born here exactly when they start), so "already has a contact" == "overlapped before".
// Callers pass proxies of different bodies only. filtQ / shapeP: the Filter word of q's shape, the shape index of p.
template <class Ctx>
PS_HD void tryPair(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st, int q, uint32_t filtQ, const Aabb& fatQ, int p, int shapeP,
                   bool allNew) {
    const Dims& d = e.cfg.d;
    if (!filtCollide(filtQ, e.scene->filt[shapeP])) return;
    const float* fatNow = fatNowOf(wk, st);
    Aabb fatP = p < kWallFixtures ? e.scene->wallFat[p] : loadAabb(fatNow, d.NP, p - kWallFixtures);
    if (!aabbOverlap(fatP, fatQ)) return;
    if (!allNew) {
        Aabb oq = loadAabb(wk.fatOld, d.NP, q - kWallFixtures);   // q moved: its old box was saved
        Aabb op = (p >= kWallFixtures && wk.moved[p - kWallFixtures]) ? loadAabb(wk.fatOld, d.NP, p - kWallFixtures) : fatP;
        if (aabbOverlap(op, oq)) return;
    }
    int lo = p < q ? p : q, hi = p < q ? q : p;
    int slot = ctx.atomicAddI(&wk.counters[3], 1);
    if (slot < e.cfg.MPAIR) wk.pairs[slot] = ((uint32_t)lo << 16) | (uint32_t)hi;
}

// BroadPhase.updatePairs' tail + ContactManager.addPair over the candidate keys in wk.pairs[0 .. counters[3]): sort,
// drop duplicates (and, with checkCache, pairs that already have a contact), append the rest in sorted proxy order.
template <class Ctx>
PS_HD void physAppendPairs(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st, bool checkCache) {
    const Dims& d = e.cfg.d;
    int nPairs = wk.counters[3];
    if (nPairs > e.cfg.MPAIR) {
        if (ctx.tid() == 0) *st.errorFlags |= ERR_PAIR_CAP;
        nPairs = e.cfg.MPAIR;
    }
    // bitonic sort of the pair keys (padded with 0xFFFFFFFF), then unique -> append in sorted order
    int n2 = 1;
    while (n2 < nPairs) n2 <<= 1;
    PS_FOR(i, n2) if (i >= nPairs) wk.pairs[i] = 0xFFFFFFFFu;
    ctx.sync();
    for (int k = 2; k <= n2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            PS_FOR(i, n2) {
                int l = i ^ j;
                if (l > i) {
                    uint32_t a = wk.pairs[i], b = wk.pairs[l];
                    bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        wk.pairs[i] = b;
                        wk.pairs[l] = a;
                    }
                }
            }
            ctx.sync();
        }
    int nOld = *st.nContacts;
    int nNew = 0;
    uint8_t* cached = (uint8_t*)wk.tc;   // [MPAIR] flags; the constraint table is idle whenever pairs are appended (MT * sizeof(TC) >= MPAIR)
    if (checkCache) {
        // ContactManager.addPair's "does a contact between these two fixtures exist already" walk, for callers whose cache
        // may hold contacts of separated boxes (Body.setTransform outside World.step): every cached key looks itself up
        // in the sorted pair list and flags the first occurrence
        PS_FOR(i, nPairs) cached[i] = 0;
        ctx.sync();
        PS_FOR(ci, nOld) {
            uint32_t k = st.ckey[ci];
            uint32_t a = k >> 16, b = k & 0xFFFFu;
            uint32_t canon = a < b ? k : ((b << 16) | a);
            int lo = 0, hi = nPairs;
            while (lo < hi) {
                int mid = (lo + hi) >> 1;
                if (wk.pairs[mid] < canon) lo = mid + 1;
                else hi = mid;
            }
            if (lo < nPairs && wk.pairs[lo] == canon) cached[lo] = 1;
        }
        ctx.sync();
    }
    int nRound = roundUpThreads(ctx, nPairs);
    for (int base = 0; base < nRound; base += ctx.nthreads()) {
        int i = base + ctx.tid();
        bool uniq = false;
        uint32_t key = 0;
        if (i < nPairs) {
            key = wk.pairs[i];
            uniq = (i == 0 || wk.pairs[i - 1] != key) && !(checkCache && cached[i]);
        }
        int total;
        int pos = nOld + nNew + ctx.exscanFlag(uniq, total);
        if (uniq) {
            if (pos < e.cfg.MC) {
                int pa = (int)(key >> 16), pb = (int)(key & 0xFFFFu);
                // Contact.create: the polygon is fixture A of a polygon-circle contact
                const Shape& sA = e.scene->shapes[proxyShape(d, pa)];
                const Shape& sB = e.scene->shapes[proxyShape(d, pb)];
                if (sA.type == SHAPE_CIRCLE && sB.type == SHAPE_POLYGON) key = ((uint32_t)pb << 16) | (uint32_t)pa;
                st.ckey[pos] = key;
                st.cinfo[pos] = 0;
                st.cimp[4 * pos + 0] = 0.0f;
                st.cimp[4 * pos + 1] = 0.0f;
                st.cimp[4 * pos + 2] = 0.0f;
                st.cimp[4 * pos + 3] = 0.0f;
            } else {
                *st.errorFlags |= ERR_CONTACT_CAP;
            }
        }
        nNew += total;
    }
    ctx.sync();
    if (ctx.tid() == 0) {
        int nc = nOld + nNew;
        *st.nContacts = nc > e.cfg.MC ? e.cfg.MC : nc;
    }
    ctx.sync();
}

// Phase H: ContactManager.findNewContacts. allNew = every proxy counts as moved and there is no previous cache (reset).
template <class Ctx>
PS_HD void physFindNewContacts(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st, bool allNew) {
    const Dims& d = e.cfg.d;
    int NB = d.NB;
    int nMoved = allNew ? d.NP : wk.counters[2];
    // (a) bodies with a moved proxy x bodies: which box unions overlap. The lists only prune: every pair below still goes
    //     through the reference's tests (tryPair), and the pairs are sorted afterwards, so the order of discovery is free.
    int nMB = wk.counters[6];
    int cap = bodyPairCap(d);
    const float* fatNow = fatNowOf(wk, st);
    {
        // flat index it = mi * NB + b, advanced without divisions
        int nT = ctx.nthreads();
        int mi = ctx.tid() / NB, b = ctx.tid() - mi * NB;
        for (int it = ctx.tid(); it < nMB * NB; it += nT) {
            int mb = wk.movedBody[mi];
            if (b != mb && aabbOverlap(loadAabb(wk.bodyFat, NB, b), loadAabb(wk.bodyFat, NB, mb))) {
                int slot = ctx.atomicAddI(&wk.counters[7], 1);
                if (slot < cap) wk.bodyPairs[slot] = ((uint32_t)mb << 16) | (uint32_t)b;
            }
            b += nT;
            while (b >= NB) {
                b -= NB;
                ++mi;
            }
        }
    }
    ctx.sync();
    int nBP = wk.counters[7];
    if (nBP <= cap) {
        // (b) one work item per (overlapping body pair, proxy of the moved body): that proxy against the other body's proxies
        int maxPer = d.R > 0 ? bodyProxyCount(d, d.P) : bodyProxyCount(d, 0);
        if (d.P > 0 && bodyProxyCount(d, 0) > maxPer) maxPer = bodyProxyCount(d, 0);
        int nT = ctx.nthreads();
        int bp = ctx.tid() / maxPer, k = ctx.tid() - bp * maxPer;
        for (int it = ctx.tid(); it < nBP * maxPer; it += nT) {
            uint32_t pr = wk.bodyPairs[bp];
            int mb = (int)(pr >> 16), b = (int)(pr & 0xFFFFu);
            if (k < bodyProxyCount(d, mb)) {
                int q = bodyFirstProxy(d, mb) + k;
                if (wk.moved[q - kWallFixtures]) {
                    Aabb fatQ = loadAabb(fatNow, d.NP, q - kWallFixtures);
                    if (aabbOverlap(loadAabb(wk.bodyFat, NB, b), fatQ)) {
                        uint32_t filtQ = e.scene->filt[bodyFirstShape(d, mb) + k];
                        int first = bodyFirstProxy(d, b), cnt = bodyProxyCount(d, b), shape0 = bodyFirstShape(d, b);
                        for (int kk = 0; kk < cnt; ++kk) tryPair(ctx, e, wk, st, q, filtQ, fatQ, first + kk, shape0 + kk, allNew);
                    }
                }
            }
            k += nT;
            while (k >= maxPer) {
                k -= maxPer;
                ++bp;
            }
        }
    } else {
        // body-pair list overflow (a very crowded arena): every moved proxy against every body
        PS_FOR(mi, nMoved) {
            int q = kWallFixtures + (allNew ? mi : wk.movedList[mi]);
            Aabb fatQ = loadAabb(fatNow, d.NP, q - kWallFixtures);
            int bq = proxyBody(d, q);
            uint32_t filtQ = e.scene->filt[proxyShape(d, q)];
            for (int b = 0; b < NB; ++b) {
                if (b == bq || !aabbOverlap(loadAabb(wk.bodyFat, NB, b), fatQ)) continue;
                int first = bodyFirstProxy(d, b), cnt = bodyProxyCount(d, b), shape0 = bodyFirstShape(d, b);
                for (int k = 0; k < cnt; ++k) tryPair(ctx, e, wk, st, q, filtQ, fatQ, first + k, shape0 + k, allNew);
            }
        }
    }
    // (c) wall: the grid names the wall proxies whose boxes come near the moved proxy's box
    PS_FOR(mi, nMoved) {
        int q = kWallFixtures + (allNew ? mi : wk.movedList[mi]);
        Aabb fatQ = loadAabb(fatNow, d.NP, q - kWallFixtures);
        float inner = e.scene->cornerX - 1.0f;
        if (fatQ.lx > -inner && fatQ.ux < inner && fatQ.ly > -inner && fatQ.uy < inner) continue;
        const Scene& sc = *e.scene;
        int ix0 = wallGridIndex(fatQ.lx, sc.wgX0, sc.wgInvX), ix1 = wallGridIndex(fatQ.ux, sc.wgX0, sc.wgInvX);
        int iy0 = wallGridIndex(fatQ.ly, sc.wgY0, sc.wgInvY), iy1 = wallGridIndex(fatQ.uy, sc.wgY0, sc.wgInvY);
        uint32_t m[3] = {0u, 0u, 0u};
        for (int ix = ix0; ix <= ix1; ++ix)
            for (int iy = iy0; iy <= iy1; ++iy) {
                const uint32_t* c = sc.wallGrid[ix * kWallGrid + iy];
                m[0] |= c[0];
                m[1] |= c[1];
                m[2] |= c[2];
            }
        uint32_t filtQ = sc.filt[proxyShape(d, q)];
        for (int w = 0; w < 3; ++w) {
            uint32_t bits = m[w];
            while (bits) {
                int k = w * 32 + ctz32(bits);
                bits &= bits - 1u;
                tryPair(ctx, e, wk, st, q, filtQ, fatQ, k, k, allNew);
            }
        }
    }
    ctx.sync();
    physAppendPairs(ctx, e, wk, st, false);
}

// Sweep of a body over the current step. Pucks carry no rotation that anything reads (circle fixtures centred on the
// origin, localCenter = 0), so their transform is a pure translation.
PS_HD Sweep bodySweep(const PhysEnv& e, const PhysWork& wk, int body) {
    Sweep sw;
    if (body < e.cfg.d.P) {
        sw.lcx = 0.0f;
        sw.lcy = 0.0f;
    } else {
        sw.lcx = e.scene->robot.lcx;
        sw.lcy = e.scene->robot.lcy;
    }
    sw.c0x = wk.c0x[body];
    sw.c0y = wk.c0y[body];
    sw.cx = wk.cx[body];
    sw.cy = wk.cy[body];
    sw.a0 = wk.a0[body];
    sw.a = wk.a[body];
    return sw;
}

// TimeOfImpact of one cached wall contact of `body` on [0, tMax]
PS_HD void toiEval(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, int body, int ci, float tMax, int& state, float& t) {
    const Dims& d = e.cfg.d;
    uint32_t key = st.ckey[ci];
    int pa = (int)(key >> 16), pb = (int)(key & 0xFFFFu);
    int bA = proxyBody(d, pa);
    DProxy prA, prB;
    proxyFromShape(prA, e.scene->shapes[proxyShape(d, pa)]);
    proxyFromShape(prB, e.scene->shapes[proxyShape(d, pb)]);
    Sweep sw = bodySweep(e, wk, body);
    Sweep wall;
    wall.lcx = wall.lcy = wall.c0x = wall.c0y = wall.cx = wall.cy = wall.a0 = wall.a = 0.0f;
    int kindBody = body < d.P ? SWEEP_TRANSLATE : SWEEP_FULL;
    if (bA == body) timeOfImpact(state, t, e.trig, prA, sw, kindBody, prB, wall, SWEEP_STATIC, e.scene->wallXf, tMax);
    else timeOfImpact(state, t, e.trig, prA, wall, SWEEP_STATIC, prB, sw, kindBody, e.scene->wallXf, tMax);
}

// Second half of World.solveTOI(Body) once the minimum-TOI contact is known: Body.advance(toi), refresh the body's wall
// contacts (contact-edge order = newest first = descending cache index), then the TOISolver (only this body moves).
PS_HD void toiResolve(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, int body, int toiContact, float toi, int nWall) {
    const Dims& d = e.cfg.d;
    const BodyConst& k = bodyConst(e, body);
    Sweep sw = bodySweep(e, wk, body);
    sweepAdvance(sw, toi);
    wk.c0x[body] = sw.c0x;
    wk.c0y[body] = sw.c0y;
    wk.a0[body] = sw.a0;
    wk.cx[body] = sw.c0x;
    wk.cy[body] = sw.c0y;
    wk.a[body] = sw.a0;
    syncBodyXf(e, wk, body);
    int cIdx[kMaxTOIContacts];
    Manifold mans[kMaxTOIContacts];
    int n = 0;
    {
        Manifold m;
        float imp[4];
        updateContact(e, wk, st, toiContact, m, imp);
    }
    for (int wi = nWall - 1; wi >= 0 && n < kMaxTOIContacts; --wi) {
        if (wk.wallBody[wi] != body) continue;
        int ci = wk.wallList[wi];
        uint32_t key = st.ckey[ci];
        int pa = (int)(key >> 16), pb = (int)(key & 0xFFFFu);
        int bA = proxyBody(d, pa), bB = proxyBody(d, pb);
        Manifold m;
        float imp[4];
        if (ci != toiContact) {
            updateContact(e, wk, st, ci, m, imp);
        } else {
            // already updated above; re-evaluating at the same transforms gives the same manifold
            const Shape& sA = e.scene->shapes[proxyShape(d, pa)];
            const Shape& sB = e.scene->shapes[proxyShape(d, pb)];
            evaluateManifold(m, sA, bodyXf(e, wk, bA), sB, bodyXf(e, wk, bB));
        }
        if (m.pointCount == 0) continue;
        cIdx[n] = ci;
        mans[n] = m;
        ++n;
    }
    float pm = k.mass * k.invMass, pi = k.mass * k.invI;
    for (int it = 0; it < 20; ++it) {
        float minSeparation = 0.0f;
        for (int c = 0; c < n; ++c) {
            uint32_t key = st.ckey[cIdx[c]];
            int pa = (int)(key >> 16), pb = (int)(key & 0xFFFFu);
            int bA = proxyBody(d, pa), bB = proxyBody(d, pb);
            float radius = e.scene->shapes[proxyShape(d, pa)].radius + e.scene->shapes[proxyShape(d, pb)].radius;
            float imA = bA == body ? pm : 0.0f, iiA = bA == body ? pi : 0.0f;
            float imB = bB == body ? pm : 0.0f, iiB = bB == body ? pi : 0.0f;
            const Manifold& m = mans[c];
            for (int j = 0; j < m.pointCount; ++j) {
                float sep = correctPoint(e, wk, bA, bB, m.type, m.localNormal, m.localPoint, m.p[j], radius, kToiBaumgarte, imA, iiA, imB, iiB);
                minSeparation = fminf_(minSeparation, sep);
            }
        }
        if (minSeparation >= -1.5f * kLinearSlop) break;
    }
}

// Phase I: World.solveTOI. Every dynamic body with cached wall contacts runs, independently of the others,
//     do { for each wall contact != toiContact (edge order): TOI on [0, toi]; TOUCHING && t < toi -> toiContact, toi = t }
//     while (found && count > 1 && iter < 50)
// The evaluations of one pass are independent except through tMax = toi, which only changes when a hit is found. So
// all wall contacts of the arena are evaluated speculatively in parallel with their body's current toi; each body then
// walks its list in edge order, accepts the first hit, and only the contacts after it are re-evaluated with the new
// toi in the next round. Every evaluation that is consumed used exactly the tMax the sequential loop would have used.
template <class Ctx>
PS_HD void physSolveTOI(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st) {
    const Dims& d = e.cfg.d;
    int n = *st.nContacts;
    PS_FOR(b, d.NB) {
        wk.bodyFlag[b] = 0;
        wk.toiActive[b] = 0;
    }
    ctx.sync();
    int nWall = 0;
    int nRound = roundUpThreads(ctx, n);
    for (int base = 0; base < nRound; base += ctx.nthreads()) {
        int i = base + ctx.tid();
        bool isWall = false;
        int body = -1;
        if (i < n) {
            uint32_t key = st.ckey[i];
            int pa = (int)(key >> 16), pb = (int)(key & 0xFFFFu);
            isWall = pa < kWallFixtures || pb < kWallFixtures;
            if (isWall) body = proxyBody(d, pa < kWallFixtures ? pb : pa);
        }
        int total;
        int pos = nWall + ctx.exscanFlag(isWall, total);
        if (isWall) {
            if (pos < e.cfg.MW) {
                wk.wallList[pos] = i;
                wk.wallBody[pos] = body;
                wk.wallNeed[pos] = 1;
                wk.bodyFlag[body] = 1;
            } else {
                *st.errorFlags |= ERR_WALL_CAP;
            }
        }
        nWall += total;
    }
    if (nWall > e.cfg.MW) nWall = e.cfg.MW;
    ctx.sync();
    PS_FOR(b, d.NB) {
        if (wk.bodyFlag[b]) {
            wk.toiActive[b] = 1;
            wk.toiT[b] = 1.0f;
            wk.toiContact[b] = -1;
            wk.toiCur[b] = nWall - 1;
            wk.toiCount[b] = 0;
            wk.toiFound[b] = 0;
            wk.toiIter[b] = 0;
        }
    }
    if (ctx.tid() == 0) wk.counters[4] = nWall > 0 ? 1 : 0;
    ctx.sync();
    while (wk.counters[4]) {
        ctx.sync();
        if (ctx.tid() == 0) {
            wk.counters[4] = 0;
            wk.stats[3] += 1;   // diagnostic: speculative rounds
        }
        PS_FOR(w, nWall) {
            if (!wk.wallNeed[w]) continue;
            int body = wk.wallBody[w];
            int state;
            float t;
            toiEval(e, wk, st, body, wk.wallList[w], wk.toiT[body], state, t);
            ctx.atomicAddI(&wk.stats[4], 1);   // diagnostic: TimeOfImpact evaluations
            wk.wallState[w] = (uint8_t)state;
            wk.wallT[w] = t;
            wk.wallNeed[w] = 0;
        }
        ctx.sync();
        PS_FOR(b, d.NB) {
            if (!wk.toiActive[b]) continue;
            bool again = false;
            for (;;) {
                // continue the current pass from toiCur
                bool hit = false;
                int wi = wk.toiCur[b];
                for (; wi >= 0; --wi) {
                    if (wk.wallBody[wi] != b) continue;
                    if (wk.wallList[wi] == wk.toiContact[b]) continue;
                    wk.toiCount[b] += 1;
                    if (wk.wallState[wi] == TOI_TOUCHING && wk.wallT[wi] < wk.toiT[b]) {
                        wk.toiContact[b] = wk.wallList[wi];
                        wk.toiT[b] = wk.wallT[wi];
                        wk.toiFound[b] = 1;
                        hit = true;
                        break;
                    }
                }
                if (hit) {
                    // the rest of the pass runs against the smaller toi
                    wk.toiCur[b] = wi - 1;
                    bool any = false;
                    for (int w2 = wi - 1; w2 >= 0; --w2)
                        if (wk.wallBody[w2] == b && wk.wallList[w2] != wk.toiContact[b]) {
                            wk.wallNeed[w2] = 1;
                            any = true;
                        }
                    if (any) {
                        again = true;
                        break;
                    }
                    // nothing left in this pass: fall through to the end-of-pass test
                }
                // end of pass
                wk.toiIter[b] += 1;
                if (wk.toiFound[b] && wk.toiCount[b] > 1 && wk.toiIter[b] < 50) {
                    wk.toiCount[b] = 0;
                    wk.toiFound[b] = 0;
                    wk.toiCur[b] = nWall - 1;
                    bool any = false;
                    for (int w2 = nWall - 1; w2 >= 0; --w2)
                        if (wk.wallBody[w2] == b && wk.wallList[w2] != wk.toiContact[b]) {
                            wk.wallNeed[w2] = 1;
                            any = true;
                        }
                    if (any) {
                        again = true;
                        break;
                    }
                    continue;   // a pass over zero contacts: count stays 0, the loop condition ends it next time round
                }
                wk.toiActive[b] = 0;
                break;
            }
            if (again) wk.counters[4] = 1;
        }
        ctx.sync();
    }
    // bodies that found a TOI contact: advance + TOISolver (rare)
    PS_FOR(b, d.NB) {
        if (wk.bodyFlag[b] && wk.toiContact[b] >= 0) {
            toiResolve(e, wk, st, b, wk.toiContact[b], wk.toiT[b], nWall);
            wk.stats[1] += 1;   // diagnostic only
        }
    }
    ctx.sync();
}

// Phase J: write the bodies back
template <class Ctx>
PS_HD void physStore(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st) {
    int NB = e.cfg.d.NB;
    PS_FOR(b, NB) {
        st.body[0 * NB + b] = wk.cx[b];
        st.body[1 * NB + b] = wk.cy[b];
        st.body[2 * NB + b] = wk.a[b];
        st.body[3 * NB + b] = wk.vx[b];
        st.body[4 * NB + b] = wk.vy[b];
        st.body[5 * NB + b] = wk.w[b];
    }
    if (ctx.tid() == 0) *st.invDt0 = 1.0f / e.cfg.dt;
    ctx.sync();
}


// World.step
template <class Ctx>
PS_HD void worldStep(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st, const float* cmdFwd, const float* cmdTorque) {
    float dtRatio = *st.invDt0 * e.cfg.dt;
    ctx.sync();
    physLoad(ctx, e, wk, st, cmdFwd, cmdTorque);
    long long t0 = phaseClock();
    physCollide(ctx, e, wk, st, dtRatio);
    long long t1 = phaseClock();
    physIslands(ctx, e, wk);
    long long t2 = phaseClock();
    physSolveVelocity(ctx, e, wk, st);
    physIntegrate(ctx, e, wk);
    long long t3 = phaseClock();
    physSolvePosition(ctx, e, wk);
    long long t4 = phaseClock();
    physSyncFixtures(ctx, e, wk, st);
    physBodyFat(ctx, e, wk, st);
    physFindNewContacts(ctx, e, wk, st, false);
    if (e.cfg.toiEnabled) physSolveTOI(ctx, e, wk, st);
    long long t5 = phaseClock();
    physStore(ctx, e, wk, st);
    if (ctx.tid() == 0) {
        wk.stats[3] = (int)((t1 - t0) >> 4);
        wk.stats[4] = (int)((t2 - t1) >> 4);
        wk.stats[5] = (int)((t3 - t2) >> 4);
        wk.stats[6] = (int)((t4 - t3) >> 4);
        wk.stats[7] = (int)((t5 - t4) >> 4);
    }
    ctx.sync();
}

}  // namespace ps
