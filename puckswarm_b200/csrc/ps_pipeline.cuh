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
                           // >= PhysCfg::warpMinContacts contacts (one warp each, level-parallel sweeps; those that fit the wide lane slot go to bucket kBucketLaneX); buckets 1..4: lane batches of
                           // islands with 1-2, 3-5, 6-12, 13+ contacts (like with like, so that the lanes of a warp finish together)
    int* counts;           // [kQueueSlots]: entries per bucket, then kQHeadWarp / kQHeadPos (consumer heads of the warp-bucket kernels)
    int cap;
    int* load;             // [2] or null: islands of >= kWarpMinContactsDefault contacts (or too many bodies for a lane) of the previous tick / counted so far
                           // this tick; swapped by stage C. While the previous tick had few of them per arena (right after a reset) the warp kernel has
                           // room, and islands of 13-16 contacts -- the lane batches' longest chains -- take a warp each as well
    int nArenasLoad;       // arenas the load counts (the group's), times adaptPerArenaX100 / 100 = the switch-over
    int adaptX100;
    int* score;            // [nArenas] or null: heaviest island of the arena lately (contacts / manifold points of its warp-bucket and big islands),
                           // written here, read and decayed by the heavy / light arena split (k_classify)
};
static constexpr int kQueueWarpBodies = 48, kQueueWarpContacts = 64;   // = kWarpBodies, kWarpContacts of ps_island_lanes.cuh
static constexpr int kLaneBodies = 12, kLaneContacts = 16;   // what a lane slot of the lane-per-island solver holds
static constexpr int kLaneXBodies = 24, kLaneXContacts = 28;   // ... and the wide lane slot (bucket kBucketLaneX): the smaller half of what used to take a warp each
static constexpr int kQueueBuckets = 10, kLaneBuckets = 4, kBucketBig = 5, kBucketLarge = 6, kBucketXL = 7, kBucketXXL = 8, kBucketLaneX = 9, kWarpMinContactsDefault = 17, kWarpMinContactsLight = 13;   // every island that fits a lane slot takes one (measured: 16 -> 17 shortens the island phase by 2 %, 10 lengthens it by 11 %)
static constexpr int kQHeadWarp = kQueueBuckets, kQHeadPos = kQueueBuckets + 1, kQHeadVel = kQueueBuckets + 2, kQueueSlots = kQueueBuckets + 3;
static constexpr int kXLPoints = 40, kXXLPoints = 56;   // manifold points of an island from which it is handed out before the other large ones: the tick ends when the
                                                       // longest island ends, and an island's sweeps cost its points (robot jams: two per contact) x up to 100
static constexpr int kLargeContacts = 28;   // warp-slot islands with at least this many contacts are handed out first (they are the tail)   // capacity of one lane's shared-memory slot (k_island_batch)

struct NoHook {
    PS_HD void operator()() const {}
};
// afterCollide: called by every thread between ContactManager.collide and the island search (the device kernel points the
// contact-graph arrays at shared memory there when the arena's touching contacts fit)
template <class Ctx, class Hook = NoHook>
PS_HD void physPre(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st, const float* cmdFwd, const float* cmdTorque,
                   const IslandQueue& q, int arena, long long* preCycles = nullptr, const Hook& afterCollide = Hook()) {
    float dtRatio = *st.invDt0 * e.cfg.dt;
    ctx.sync();
    long long t0 = phaseClock();
    physLoad(ctx, e, wk, st, cmdFwd, cmdTorque);
    long long t1 = phaseClock();
    physCollide(ctx, e, wk, st, dtRatio);
    long long t2 = phaseClock();
    afterCollide();
    physIslands(ctx, e, wk);
    long long t3 = phaseClock();
    const Dims& d = e.cfg.d;
    // bodies without touching contacts are islands of their own: Island.solve reduces to the position integration
    PS_FOR(b, d.NB) {
        if (wk.adjOff[b] == wk.adjOff[b + 1]) integrateBody(e, wk, b);
    }
    if (ctx.tid() == 0 && e.cfg.posIters > 0) wk.stats[0] = 1;   // contact-free islands run the position loop once
    ctx.sync();
    int nIsl = wk.counters[1];
    int warpMin = e.cfg.warpMinContacts;
    if (q.load && (long long)q.load[0] * 100 < (long long)q.nArenasLoad * q.adaptX100 && warpMin > kWarpMinContactsLight) warpMin = kWarpMinContactsLight;
    PS_FOR(k, nIsl) {
        int nb = wk.islandBodyStart[k + 1] - wk.islandBodyStart[k];
        int nc = wk.islandStart[k + 1] - wk.islandStart[k];
        if (q.load && (nc >= kWarpMinContactsDefault || nb > kLaneBodies)) ctx.atomicAddI(&q.load[1], 1);
        // 0: a warp each (WarpSlot); kBucketBig: what outgrows that (BigSlot); 1..4: a lane each, by contact count
        int bucket = (nc >= warpMin || nc > kLaneContacts || nb > kLaneBodies) ? ((nb > kQueueWarpBodies || nc > kQueueWarpContacts) ? kBucketBig : (nc >= kLargeContacts ? kBucketLarge : 0))
                                                                                               : (nc <= 2 ? 1 : (nc <= 5 ? 2 : (nc <= 12 ? 3 : 4)));
        // the smaller warp-bucket islands take a lane of a wide lane slot instead (32 islands per instruction stream)
        if (bucket == 0 && nc <= e.cfg.laneXMaxContacts && nc <= kLaneXContacts && nb <= kLaneXBodies) bucket = kBucketLaneX;
        if (bucket == kBucketLarge) {
            // the heaviest warp-slot islands first: estimated by their manifold points
            int lo = wk.islandStart[k], points = 0;
            for (int i = 0; i < nc; ++i) points += wk.tc[wk.order[lo + i]].solverPoints;
            if (points >= kXXLPoints) bucket = kBucketXXL;
            else if (points >= kXLPoints) bucket = kBucketXL;
        }
        if (q.score && (bucket == 0 || bucket >= kBucketBig) && bucket != kBucketLaneX) {
            // a sweep costs the island's manifold points; robots' contacts carry two (estimated for the buckets that did not count them)
            int w = bucket == kBucketBig ? 2 * nc : (bucket == 0 ? nc : (bucket == kBucketLarge ? 32 : (bucket == kBucketXL ? 48 : 64)) + nc / 4);
            ctx.atomicMaxI(&q.score[arena], w);
        }
        int pos = ctx.atomicAddI(&q.counts[bucket], 1);
        if (pos < q.cap) q.entries[(size_t)bucket * q.cap + pos] = (unsigned)arena * (unsigned)d.NB + (unsigned)k;
    }
    ctx.sync();
    if (preCycles && ctx.tid() == 0) {   // diagnostics: cycles of load, collide, graph + islands, integration + queue
        long long t4 = phaseClock();
        preCycles[0] = t1 - t0;
        preCycles[1] = t2 - t1;
        preCycles[2] = t3 - t2;
        preCycles[3] = t4 - t3;
    }
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
