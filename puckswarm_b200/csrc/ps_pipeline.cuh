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
    unsigned* entries;     // [3][cap]: arena * NB + island, bucketed by contact count (>= 8, 3..7, 1..2)
    int* counts;           // [4]: entries per bucket, [3] = consumer head
    int cap;
};
PS_HD int islandBucket(int nContacts) { return nContacts >= 8 ? 0 : (nContacts >= 3 ? 1 : 2); }

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
        int bucket = islandBucket(wk.islandStart[k + 1] - wk.islandStart[k]);
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
// Returns false when the island is finished; *itersOut then holds its position iterations.
PS_HD bool islandStep(const PhysEnv& e, const PhysWork& wk, const ArenaState& st, IslandCursor& c, int* itersOut) {
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
        for (int i = c.blo; i < c.bhi; ++i) integrateBody(e, wk, wk.ibody[i]);
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
        const BodyConst &kp = e.scene->puck, &kr = e.scene->robot;
        int P = e.cfg.d.P;
        const TC& t = wk.tc[wk.order[c.i]];
        int bA = t.bodyA, bB = t.bodyB;
        float imA = bA < 0 ? 0.0f : (bA < P ? kp.mass * kp.invMass : kr.mass * kr.invMass);
        float iiA = bA < 0 ? 0.0f : (bA < P ? kp.mass * kp.invI : kr.mass * kr.invI);
        float imB = bB < 0 ? 0.0f : (bB < P ? kp.mass * kp.invMass : kr.mass * kr.invMass);
        float iiB = bB < 0 ? 0.0f : (bB < P ? kp.mass * kp.invI : kr.mass * kr.invI);
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

#if defined(__CUDACC__)
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
    if (lane == 0) {
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
        lw.order = s.order;
        lw.ibody = s.ibody;
        IslandCursor c;
        c.lo = 0;
        c.hi = nc;
        c.blo = 0;
        c.bhi = nb;
        c.phase = 0;
        c.it = 0;
        c.i = 0;
        c.minSep = 0.0f;
        c.changed = false;
        int iters = 0;
        while (islandStep(le, lw, st, c, &iters)) {
        }
        *itersOut = iters;
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
