// puckswarm_b200 -- World.step as a three-stage pipeline over the whole batch.
//
// The fused per-arena kernel (worldStep in ps_physics.cuh) leaves most lanes idle while one thread per island runs
// the order-dependent Gauss-Seidel sweeps. The pipeline splits the step where the parallel axis changes:
//   stage A  physPre   one CTA per arena : Robot.move, ContactManager.collide, constraint set-up, contact graph,
//                                          islands by DFS; contact-free bodies are integrated right away; every
//                                          island with contacts is pushed on a batch-wide queue (largest first)
//   stage B  island workers, one LANE per island across all arenas: warm start, velocity sweeps, store impulses,
//                                          integrate the island's bodies, position sweeps until converged. A lane
//                                          runs a small state machine (one contact per step) and pulls the next
//                                          island when it finishes, so warps stay converged and full
//   stage C  physPost  one CTA per arena : synchronizeFixtures, findNewContacts, solveTOI, write back
// Islands are independent (disjoint dynamic bodies; the wall is immutable), so the order in which lanes pick
// them up does not influence any result.
#pragma once
#include "ps_physics.cuh"

namespace ps {

struct IslandQueue {
    unsigned* entries;     // [kQueueBuckets][cap]: arena * NB + island. Bucket 0: islands that do not fit a lane slot or have
                           // >= PhysCfg::warpMinContacts contacts (one warp each, level-parallel sweeps); buckets 1..4: lane batches of
                           // islands with 1-2, 3-5, 6-12, 13+ contacts (like with like, so that the lanes of a warp finish together)
    int* counts;           // [kQueueBuckets + 1]: entries per bucket, last = consumer head of bucket 0
    int cap;
};
static constexpr int kLaneBodies = 12, kLaneRobots = 4;
static constexpr int kQueueBuckets = 5, kWarpMinContactsDefault = 16;   // capacity of one lane's shared-memory slot (k_island_batch)

template <class Ctx>
PS_HD void physPre(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st, const float* cmdFwd, const float* cmdTorque,
                   const IslandQueue& q, int arena) {
    float dtRatio = *st.invDt0 * e.cfg.dt;
    ctx.sync();
    physLoad(ctx, e, wk, st, cmdFwd, cmdTorque);
    physCollide(ctx, e, wk, st, dtRatio);
    physIslands(ctx, e, wk);
    const Dims& d = e.cfg.d;
    // bodies without touching contacts are islands of their own: Island.solve reduces to the position integration
    PS_FOR(b, d.NB) {
        if (wk.adjOff[b] == wk.adjOff[b + 1]) integrateBody(e, wk, b);
    }
    int nIsl = wk.counters[1];
    PS_FOR(k, nIsl) {
        int nb = wk.islandBodyStart[k + 1] - wk.islandBodyStart[k], nr = 0;
        for (int i = wk.islandBodyStart[k]; i < wk.islandBodyStart[k + 1]; ++i) nr += wk.ibody[i] >= d.P ? 1 : 0;
        int nc = wk.islandStart[k + 1] - wk.islandStart[k];
        int bucket = (nc >= e.cfg.warpMinContacts || nb > kLaneBodies || nr > kLaneRobots) ? 0 : (nc <= 2 ? 1 : (nc <= 5 ? 2 : (nc <= 12 ? 3 : 4)));
        int pos = ctx.atomicAddI(&q.counts[bucket], 1);
        if (pos < q.cap) q.entries[(size_t)bucket * q.cap + pos] = (unsigned)arena * (unsigned)d.NB + (unsigned)k;
    }
    if (ctx.tid() == 0 && e.cfg.posIters > 0) wk.stats[0] = 1;   // contact-free islands run the position loop once
    ctx.sync();
}

// One island's Island.solve as a resumable state machine: each step handles one contact
struct IslandCursor {
    int lo, hi;        // range in wk.order
    int blo, bhi;      // range in wk.ibody
    int phase;         // 0 warm start, 1 velocity sweeps, 2 position sweeps
    int it, i;
    float minSep;
    bool changed;      // some body moved during the current position sweep
};
PS_HD void islandBegin(IslandCursor& c, const PhysWork& wk, int k) {
    c.lo = wk.islandStart[k];
    c.hi = wk.islandStart[k + 1];
    c.blo = wk.islandBodyStart[k];
    c.bhi = wk.islandBodyStart[k + 1];
    c.phase = 0;
    c.it = 0;
    c.i = c.lo;
    c.minSep = 0.0f;
    c.changed = false;
}
struct DefaultIntegrator {   // integrates the i-th entry of the island's body list
    PS_HD void operator()(const PhysEnv& e, const PhysWork& wk, int i) const { integrateBody(e, wk, wk.ibody[i]); }
};
// Returns false when the island is finished; *itersOut then holds its position iterations.
template <class Integrator>
PS_HD bool islandStepT(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, IslandCursor& c, int* itersOut, const Integrator& integrate) {
    if (c.phase == 0) {
        warmStartOne(e, wk, wk.tc[wk.order[c.i]]);
        if (++c.i < c.hi) return true;
        c.i = c.lo;
        c.phase = 1;
        if (e.cfg.velIters > 0) return true;
    } else if (c.phase == 1) {
        solveVelocityOne(e, wk, wk.tc[wk.order[c.i]]);
        if (++c.i < c.hi) return true;
        c.i = c.lo;
        if (++c.it < e.cfg.velIters) return true;
    }
    if (c.phase <= 1) {
        // ContactSolver.storeImpulses, then integrate the island's bodies (Island.solve step 6)
        for (int i = c.lo; i < c.hi; ++i) {
            const TC& t = wk.tc[wk.order[i]];
            for (int j = 0; j < t.solverPoints; ++j) {
                st.cimp[4 * t.ci + 2 * j] = t.nImp[j];
                st.cimp[4 * t.ci + 2 * j + 1] = t.tImp[j];
            }
        }
        for (int i = c.blo; i < c.bhi; ++i) integrate(e, wk, i);
        c.phase = 2;
        c.it = 0;
        c.i = c.lo;
        c.minSep = 0.0f;
        c.changed = false;
        if (e.cfg.posIters > 0) return true;
        *itersOut = 0;
        return false;
    }
    // position sweep: one contact
    {
        int P = e.cfg.d.P;
        const TC& t = wk.tc[wk.order[c.i]];
        int bA = t.bodyA, bB = t.bodyB;
        float imA = bA < 0 ? 0.0f : (bA < P ? e.hot.pmPuck : e.hot.pmRobot);
        float iiA = bA < 0 ? 0.0f : (bA < P ? e.hot.piPuck : e.hot.piRobot);
        float imB = bB < 0 ? 0.0f : (bB < P ? e.hot.pmPuck : e.hot.pmRobot);
        float iiB = bB < 0 ? 0.0f : (bB < P ? e.hot.piPuck : e.hot.piRobot);
        for (int j = 0; j < t.solverPoints; ++j) {
            float sep = correctPoint(e, wk, bA, bB, t.type, t.localNormal, t.localPoint, t.lp[j], t.radius, kContactBaumgarte, imA, iiA, imB, iiB, &c.changed);
            c.minSep = fminf_(c.minSep, sep);
        }
        if (++c.i < c.hi) return true;
        ++c.it;
        if (c.minSep >= -1.5f * kLinearSlop || c.it >= e.cfg.posIters) {
            *itersOut = c.it;
            return false;
        }
        if (!c.changed) {
            // a sweep that moved nothing leaves the island exactly where it started, so every later sweep repeats it bit
            // for bit and the loop can only end at the iteration cap: report that outcome now
            *itersOut = e.cfg.posIters;
            return false;
        }
        c.i = c.lo;
        c.minSep = 0.0f;
        c.changed = false;
        return true;
    }
}

PS_HD bool islandStep(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, IslandCursor& c, int* itersOut) {
    return islandStepT(e, wk, st, c, itersOut, DefaultIntegrator());
}

#if defined(__CUDACC__)
// Lane-per-island solver: every lane of a warp runs a different island. The island's bodies live in a lane-interleaved
// shared-memory slot (element e of lane l at [e][l], conflict-free); the constraints stay in HBM / L2 (they are read,
// not chased: no dependent latency). Body indices inside the staged island are pre-multiplied by 32 so that the
// ordinary solver routines address the interleaved arrays unchanged: wk.cx[b * 32] with wk.cx offset by the lane.
struct LaneSlots {
    float body[6][kLaneBodies][32];    // cx cy a vx vy w
    float rob[4][kLaneRobots][32];     // rpx rpy rc rs
    uint16_t gidx[kLaneBodies][32];    // local -> arena body index
};
struct StagedIntegrator {
    const PhysWork* g;                 // the arena's arrays in HBM (c0 / xf0 are only needed by the later stages)
    const uint16_t* gidx;              // lane's column of LaneSlots::gidx (stride 32)
    int nPuckLocal, P;
    __device__ __forceinline__ void operator()(const PhysEnv& e, const PhysWork& wk, int local) const {
        // staged index = local * 32; same arithmetic as integrateBody
        int bl = local * 32;
        int gb = gidx[local * 32];
        float dt = e.cfg.dt;
        float vx = wk.vx[bl], vy = wk.vy[bl], w = wk.w[bl];
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
        wk.vx[bl] = vx;
        wk.vy[bl] = vy;
        wk.w[bl] = w;
        g->c0x[gb] = wk.cx[bl];
        g->c0y[gb] = wk.cy[bl];
        g->a0[gb] = wk.a[bl];
        if (gb >= P) {
            int r = gb - P, lr = (local - nPuckLocal) * 32;
            g->r0px[r] = wk.rpx[lr];
            g->r0py[r] = wk.rpy[lr];
            g->r0c[r] = wk.rc[lr];
            g->r0s[r] = wk.rs[lr];
        }
        wk.cx[bl] += dt * vx;
        wk.cy[bl] += dt * vy;
        wk.a[bl] += dt * w;
        syncBodyXf(e, wk, bl);
    }
};
// One lane's staged island: the cursor plus the rebased working set
struct LaneIsland {
    IslandCursor c;
    PhysEnv le;
    PhysWork lw;
    StagedIntegrator integ;
    int nb, nPuck;
};
// Stage island k of the arena behind wk into this lane's slot. The caller guarantees the island fits.
__device__ __forceinline__ void islandStageLane(const PhysEnv& e, const PhysWork& wk, int k, LaneSlots& s, int lane, LaneIsland& L) {
    int lo = wk.islandStart[k], hi = wk.islandStart[k + 1];
    int blo = wk.islandBodyStart[k], bhi = wk.islandBodyStart[k + 1];
    int nb = bhi - blo;
    int P = e.cfg.d.P;
    int nPuck = 0;
    for (int i = 0; i < nb; ++i) nPuck += wk.ibody[blo + i] < P ? 1 : 0;
    // local numbering: pucks first (ibody order), then robots
    int ip = 0, ir = nPuck;
    for (int i = 0; i < nb; ++i) {
        int g = wk.ibody[blo + i];
        int l = g < P ? ip++ : ir++;
        s.gidx[l][lane] = (uint16_t)g;
        s.body[0][l][lane] = wk.cx[g];
        s.body[1][l][lane] = wk.cy[g];
        s.body[2][l][lane] = wk.a[g];
        s.body[3][l][lane] = wk.vx[g];
        s.body[4][l][lane] = wk.vy[g];
        s.body[5][l][lane] = wk.w[g];
        if (g >= P) {
            int r = g - P, lr = l - nPuck;
            s.rob[0][lr][lane] = wk.rpx[r];
            s.rob[1][lr][lane] = wk.rpy[r];
            s.rob[2][lr][lane] = wk.rc[r];
            s.rob[3][lr][lane] = wk.rs[r];
        }
    }
    // constraints: body references -> staged indices (the arena-level indices are not needed again this step)
    for (int i = lo; i < hi; ++i) {
        TC& t = wk.tc[wk.order[i]];
        int a = t.bodyA, b = t.bodyB;
        int la = -1, lb = -1;
        for (int j = 0; j < nb; ++j) {
            int g = s.gidx[j][lane];
            if (g == a) la = j * 32;
            if (g == b) lb = j * 32;
        }
        t.bodyA = a < 0 ? -1 : la;
        t.bodyB = b < 0 ? -1 : lb;
    }
    L.le = e;
    L.le.cfg.d.P = nPuck * 32;
    L.lw = wk;
    L.lw.cx = &s.body[0][0][lane];
    L.lw.cy = &s.body[1][0][lane];
    L.lw.a = &s.body[2][0][lane];
    L.lw.vx = &s.body[3][0][lane];
    L.lw.vy = &s.body[4][0][lane];
    L.lw.w = &s.body[5][0][lane];
    L.lw.rpx = &s.rob[0][0][lane];
    L.lw.rpy = &s.rob[1][0][lane];
    L.lw.rc = &s.rob[2][0][lane];
    L.lw.rs = &s.rob[3][0][lane];
    L.integ.gidx = &s.gidx[0][lane];
    L.integ.nPuckLocal = nPuck;
    L.integ.P = P;
    L.c.lo = lo;
    L.c.hi = hi;
    L.c.blo = 0;       // the integrator is handed the local body numbers 0 .. nb-1
    L.c.bhi = nb;
    L.c.phase = 0;
    L.c.it = 0;
    L.c.i = lo;
    L.c.minSep = 0.0f;
    L.c.changed = false;
    L.nb = nb;
    L.nPuck = nPuck;
}
// Phase-synchronous solve of up to 32 staged islands, one per lane: the lanes walk warm start, velocity sweeps, integration
// and position sweeps together (trip counts are the warp maximum, a lane without work at a trip idles), so the warp stays
// converged and every warp instruction advances all its islands. `has` = this lane holds an island. Returns the lane's
// position iterations.
__device__ __forceinline__ int islandSolveLanes(const PhysEnv& e, const ArenaState& st, LaneIsland& L, bool has) {
    const unsigned FULL = 0xFFFFFFFFu;
    int nc = has ? L.c.hi - L.c.lo : 0;
    int lo = L.c.lo;
    int maxNc = nc;
    for (int o = 16; o > 0; o >>= 1) maxNc = max(maxNc, __shfl_xor_sync(FULL, maxNc, o));
    const PhysEnv& le = L.le;
    const PhysWork& lw = L.lw;
    for (int i = 0; i < maxNc; ++i) {
        if (i < nc) warmStartOne(le, lw, lw.tc[lw.order[lo + i]]);
        __syncwarp();
    }
    for (int it = 0; it < e.cfg.velIters; ++it)
        for (int i = 0; i < maxNc; ++i) {
            if (i < nc) solveVelocityOne(le, lw, lw.tc[lw.order[lo + i]]);
            __syncwarp();
        }
    if (has) {
        for (int i = 0; i < nc; ++i) {
            const TC& t = lw.tc[lw.order[lo + i]];
            for (int j = 0; j < t.solverPoints; ++j) {
                st.cimp[4 * t.ci + 2 * j] = t.nImp[j];
                st.cimp[4 * t.ci + 2 * j + 1] = t.tImp[j];
            }
        }
        for (int b = 0; b < L.nb; ++b) L.integ(le, lw, b);
    }
    __syncwarp();
    int iters = 0;
    bool running = has && e.cfg.posIters > 0;
    int P = le.cfg.d.P;
    while (__any_sync(FULL, running)) {
        float minSep = 0.0f;
        bool changed = false;
        for (int i = 0; i < maxNc; ++i) {
            if (running && i < nc) {
                const TC& t = lw.tc[lw.order[lo + i]];
                int bA = t.bodyA, bB = t.bodyB;
                float imA = bA < 0 ? 0.0f : (bA < P ? e.hot.pmPuck : e.hot.pmRobot);
                float iiA = bA < 0 ? 0.0f : (bA < P ? e.hot.piPuck : e.hot.piRobot);
                float imB = bB < 0 ? 0.0f : (bB < P ? e.hot.pmPuck : e.hot.pmRobot);
                float iiB = bB < 0 ? 0.0f : (bB < P ? e.hot.piPuck : e.hot.piRobot);
                for (int j = 0; j < t.solverPoints; ++j) {
                    float sep = correctPoint(le, lw, bA, bB, t.type, t.localNormal, t.localPoint, t.lp[j], t.radius, kContactBaumgarte, imA, iiA, imB, iiB, &changed);
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

__device__ __forceinline__ void islandFinishLane(const PhysEnv& e, const PhysWork& wk, LaneSlots& s, int lane, const LaneIsland& L) {
    int P = e.cfg.d.P;
    for (int l = 0; l < L.nb; ++l) {
        int g = s.gidx[l][lane];
        wk.cx[g] = s.body[0][l][lane];
        wk.cy[g] = s.body[1][l][lane];
        wk.a[g] = s.body[2][l][lane];
        wk.vx[g] = s.body[3][l][lane];
        wk.vy[g] = s.body[4][l][lane];
        wk.w[g] = s.body[5][l][lane];
        if (g >= P) {
            int r = g - P, lr = l - L.nPuck;
            wk.rpx[r] = s.rob[0][lr][lane];
            wk.rpy[r] = s.rob[1][lr][lane];
            wk.rc[r] = s.rob[2][lr][lane];
            wk.rs[r] = s.rob[3][lr][lane];
        }
    }
}

// Warp-per-island solver for islands of up to kWarpBodies bodies / kWarpContacts contacts: the warp stages the island's
// bodies and constraints into its shared-memory slot (local body indices: pucks first, then robots), lane 0 runs the
// same state machine on the staged copy (shared-memory latency instead of L2 on the Gauss-Seidel dependency chain),
// and the warp writes the bodies back. Same arithmetic on the same values: results are identical to the lane path.
static constexpr int kWarpBodies = 48, kWarpRobots = 16, kWarpContacts = 64;
struct WarpSlot {
    float body[9][kWarpBodies];          // cx cy a vx vy w c0x c0y a0
    float rob[8][kWarpRobots];           // rpx rpy rc rs r0px r0py r0c r0s
    TC tc[kWarpContacts];
    uint16_t order[kWarpContacts];
    uint16_t ibody[kWarpBodies];
    uint16_t lmap[kWarpBodies];          // local -> arena body index
    uint8_t level[kWarpContacts], sched[kWarpContacts], bodyLevel[kWarpBodies];
    int levelStart[kWarpContacts + 3];
    int nLevels;
};
// Returns false if the island does not fit the slot (the caller falls back to the lane path).
__device__ __forceinline__ bool islandSolveWarp(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, int k, WarpSlot& s, int* itersOut) {
    const unsigned FULL = 0xFFFFFFFFu;
    int lane = threadIdx.x & 31;
    int lo = wk.islandStart[k], hi = wk.islandStart[k + 1];
    int blo = wk.islandBodyStart[k], bhi = wk.islandBodyStart[k + 1];
    int nb = bhi - blo, nc = hi - lo;
    int P = e.cfg.d.P;
    if (nb > kWarpBodies || nc > kWarpContacts) return false;
    // local numbering: pucks first (ibody order), then robots
    int nPuck = 0, nRob = 0;
    for (int base = 0; base < nb; base += 32) {
        int i = base + lane;
        int g = i < nb ? (int)wk.ibody[blo + i] : -1;
        bool isP = g >= 0 && g < P, isR = g >= P;
        unsigned mp = __ballot_sync(FULL, isP), mr = __ballot_sync(FULL, isR);
        unsigned below = (1u << lane) - 1u;
        if (isP) s.lmap[nPuck + __popc(mp & below)] = (uint16_t)g;
        nPuck += __popc(mp);
        // robots are numbered in a second pass (they go after all pucks)
        nRob += __popc(mr);
    }
    if (nRob > kWarpRobots) return false;
    {
        int r = 0;
        for (int base = 0; base < nb; base += 32) {
            int i = base + lane;
            int g = i < nb ? (int)wk.ibody[blo + i] : -1;
            bool isR = g >= P;
            unsigned mr = __ballot_sync(FULL, isR);
            if (isR) s.lmap[nPuck + r + __popc(mr & ((1u << lane) - 1u))] = (uint16_t)g;
            r += __popc(mr);
        }
    }
    __syncwarp();
    for (int i = lane; i < nb; i += 32) {
        int g = s.lmap[i];
        s.body[0][i] = wk.cx[g];
        s.body[1][i] = wk.cy[g];
        s.body[2][i] = wk.a[g];
        s.body[3][i] = wk.vx[g];
        s.body[4][i] = wk.vy[g];
        s.body[5][i] = wk.w[g];
        s.ibody[i] = (uint16_t)i;
        if (g >= P) {
            int r = g - P, lr = i - nPuck;
            s.rob[0][lr] = wk.rpx[r];
            s.rob[1][lr] = wk.rpy[r];
            s.rob[2][lr] = wk.rc[r];
            s.rob[3][lr] = wk.rs[r];
        }
    }
    for (int i = lane; i < nc; i += 32) {
        TC t = wk.tc[wk.order[lo + i]];
        int la = -1, lb = -1;
        for (int j = 0; j < nb; ++j) {
            int g = s.lmap[j];
            if (g == t.bodyA) la = j;
            if (g == t.bodyB) lb = j;
        }
        t.bodyA = t.bodyA < 0 ? -1 : la;
        t.bodyB = t.bodyB < 0 ? -1 : lb;
        s.tc[i] = t;
        s.order[i] = (uint16_t)i;
    }
    __syncwarp();
    // ---- level schedule of one Gauss-Seidel sweep: contact i (in solve order) goes to level 1 + the highest level of an
    // earlier contact that shares a dynamic body with it. Contacts of one level touch disjoint bodies, so the lanes
    // process them together; levels run in order, so every body still sees its contacts in the reference's order and
    // the sweep is bit-identical to the sequential one.
    if (lane == 0) {
        int nLevels = 0;
        for (int b = 0; b < nb; ++b) s.bodyLevel[b] = 0;
        for (int i = 0; i < nc; ++i) {
            int a = s.tc[i].bodyA, b = s.tc[i].bodyB;
            int lv = 0;
            if (a >= 0) lv = s.bodyLevel[a];
            if (b >= 0 && s.bodyLevel[b] > lv) lv = s.bodyLevel[b];
            lv += 1;
            if (a >= 0) s.bodyLevel[a] = (uint8_t)lv;
            if (b >= 0) s.bodyLevel[b] = (uint8_t)lv;
            s.level[i] = (uint8_t)lv;
            if (lv > nLevels) nLevels = lv;
        }
        // counting sort by level (stable)
        for (int l = 0; l <= nLevels + 1; ++l) s.levelStart[l] = 0;
        for (int i = 0; i < nc; ++i) s.levelStart[s.level[i] + 1]++;
        for (int l = 1; l <= nLevels + 1; ++l) s.levelStart[l] += s.levelStart[l - 1];
        for (int i = 0; i < nc; ++i) s.sched[s.levelStart[s.level[i]]++] = (uint8_t)i;
        // levelStart[l] now holds the end of level l == start of level l + 1; shift back
        for (int l = nLevels + 1; l >= 1; --l) s.levelStart[l] = s.levelStart[l - 1];
        s.levelStart[0] = 0;
        s.nLevels = nLevels;
    }
    __syncwarp();
    {
        PhysEnv le = e;
        le.cfg.d.P = nPuck;
        PhysWork lw = wk;
        lw.cx = s.body[0];
        lw.cy = s.body[1];
        lw.a = s.body[2];
        lw.vx = s.body[3];
        lw.vy = s.body[4];
        lw.w = s.body[5];
        lw.c0x = s.body[6];
        lw.c0y = s.body[7];
        lw.a0 = s.body[8];
        lw.rpx = s.rob[0];
        lw.rpy = s.rob[1];
        lw.rc = s.rob[2];
        lw.rs = s.rob[3];
        lw.r0px = s.rob[4];
        lw.r0py = s.rob[5];
        lw.r0c = s.rob[6];
        lw.r0s = s.rob[7];
        lw.tc = s.tc;
        int nLevels = s.nLevels;
        // levels are 1-based: level l holds sched[levelStart[l] .. levelStart[l + 1])
        // ContactSolver.warmStart
        for (int l = 1; l <= nLevels; ++l) {
            for (int q = s.levelStart[l] + lane; q < s.levelStart[l + 1]; q += 32) warmStartOne(le, lw, s.tc[s.sched[q]]);
            __syncwarp();
        }
        // velocity sweeps
        for (int it = 0; it < e.cfg.velIters; ++it)
            for (int l = 1; l <= nLevels; ++l) {
                for (int q = s.levelStart[l] + lane; q < s.levelStart[l + 1]; q += 32) solveVelocityOne(le, lw, s.tc[s.sched[q]]);
                __syncwarp();
            }
        // ContactSolver.storeImpulses
        for (int i = lane; i < nc; i += 32) {
            const TC& t = s.tc[i];
            for (int j2 = 0; j2 < t.solverPoints; ++j2) {
                st.cimp[4 * t.ci + 2 * j2] = t.nImp[j2];
                st.cimp[4 * t.ci + 2 * j2 + 1] = t.tImp[j2];
            }
        }
        // integrate the island's bodies
        for (int b = lane; b < nb; b += 32) integrateBody(le, lw, b);
        __syncwarp();
        // position sweeps
        int iters = 0;
        int P = le.cfg.d.P;
        for (int it = 0; it < e.cfg.posIters; ++it) {
            float minSep = 0.0f;
            bool changed = false;
            for (int l = 1; l <= nLevels; ++l) {
                for (int q = s.levelStart[l] + lane; q < s.levelStart[l + 1]; q += 32) {
                    const TC& t = s.tc[s.sched[q]];
                    int bA = t.bodyA, bB = t.bodyB;
                    float imA = bA < 0 ? 0.0f : (bA < P ? e.hot.pmPuck : e.hot.pmRobot);
                    float iiA = bA < 0 ? 0.0f : (bA < P ? e.hot.piPuck : e.hot.piRobot);
                    float imB = bB < 0 ? 0.0f : (bB < P ? e.hot.pmPuck : e.hot.pmRobot);
                    float iiB = bB < 0 ? 0.0f : (bB < P ? e.hot.piPuck : e.hot.piRobot);
                    for (int j2 = 0; j2 < t.solverPoints; ++j2) {
                        float sep = correctPoint(le, lw, bA, bB, t.type, t.localNormal, t.localPoint, t.lp[j2], t.radius, kContactBaumgarte, imA, iiA, imB, iiB, &changed);
                        minSep = fminf_(minSep, sep);
                    }
                }
                __syncwarp();
            }
            for (int o = 16; o > 0; o >>= 1) {
                float other = __shfl_xor_sync(FULL, minSep, o);
                minSep = fminf_(minSep, other);
            }
            bool anyChanged = __any_sync(FULL, changed);
            ++iters;
            if (minSep >= -1.5f * kLinearSlop) break;
            if (!anyChanged) {   // fixed point: every later sweep repeats this one
                iters = e.cfg.posIters;
                break;
            }
        }
        if (lane == 0) *itersOut = iters;
    }
    __syncwarp();
    for (int i = lane; i < nb; i += 32) {
        int g = s.lmap[i];
        wk.cx[g] = s.body[0][i];
        wk.cy[g] = s.body[1][i];
        wk.a[g] = s.body[2][i];
        wk.vx[g] = s.body[3][i];
        wk.vy[g] = s.body[4][i];
        wk.w[g] = s.body[5][i];
        wk.c0x[g] = s.body[6][i];
        wk.c0y[g] = s.body[7][i];
        wk.a0[g] = s.body[8][i];
        if (g >= P) {
            int r = g - P, lr = i - nPuck;
            wk.rpx[r] = s.rob[0][lr];
            wk.rpy[r] = s.rob[1][lr];
            wk.rc[r] = s.rob[2][lr];
            wk.rs[r] = s.rob[3][lr];
            wk.r0px[r] = s.rob[4][lr];
            wk.r0py[r] = s.rob[5][lr];
            wk.r0c[r] = s.rob[6][lr];
            wk.r0s[r] = s.rob[7][lr];
        }
    }
    __syncwarp();
    return true;
}
#endif

template <class Ctx>
PS_HD void physPost(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st) {
    ctx.sync();
    if (ctx.tid() == 0) wk.stats[2] = wk.counters[1];
    long long t0 = phaseClock();
    physSyncFixtures(ctx, e, wk, st);
    physBodyFat(ctx, e, wk, st);
    long long t1 = phaseClock();
    physFindNewContacts(ctx, e, wk, st, false);
    long long t2 = phaseClock();
    if (e.cfg.toiEnabled) physSolveTOI(ctx, e, wk, st);
    long long t3 = phaseClock();
    physStore(ctx, e, wk, st);
    if (ctx.tid() == 0) {   // diagnostics: cycles / 16 of fixture sync, pair search, TOI
        wk.stats[5] = (int)((t1 - t0) >> 4);
        wk.stats[6] = (int)((t2 - t1) >> 4);
        wk.stats[7] = (int)((t3 - t2) >> 4);
    }
    ctx.sync();
}

}  // namespace ps
