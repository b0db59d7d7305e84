// puckswarm_b200 -- seeded arena construction on the device.
//
// Replaces `new Experiment(seed)` + `new Arena(world, null, false)` for one arena
// (src/experiment/Experiment.java:21-24, src/arena/Arena.java:45-120,143-181,
// src/arena/RoundedRectangleEnclosure.java:55-93): java.util.Random(seed), rejection-sampled puck and robot
// poses, fixture proxies (tight AABB + 0.1) and the first BroadPhase.updatePairs over all proxies.
#pragma once
#include "ps_physics.cuh"
#include "../../include/puckswarm_b200.h"

namespace ps {

// java.util.Random
struct JRandom {
    uint64_t seed;
    int haveNextNextGaussian;
    double nextNextGaussian;
};
PS_HD void jrSetSeed(JRandom& r, int64_t s) {
    r.seed = ((uint64_t)s ^ 0x5DEECE66DULL) & ((1ULL << 48) - 1);
    r.haveNextNextGaussian = 0;
    r.nextNextGaussian = 0.0;
}
PS_HD int32_t jrNext(JRandom& r, int bits) {
    r.seed = (r.seed * 0x5DEECE66DULL + 0xBULL) & ((1ULL << 48) - 1);
    return (int32_t)(r.seed >> (48 - bits));
}
PS_HD float jrNextFloat(JRandom& r) { return jrNext(r, 24) / (float)(1 << 24); }
PS_HD int32_t jrNextInt(JRandom& r, int32_t n) {
    if ((n & -n) == n) return (int32_t)(((int64_t)n * (int64_t)jrNext(r, 31)) >> 31);
    int32_t bits, val;
    do {
        bits = jrNext(r, 31);
        val = bits % n;
    } while ((int32_t)((uint32_t)bits - (uint32_t)val + (uint32_t)(n - 1)) < 0);
    return val;
}
PS_HD double jrNextDouble(JRandom& r) {
    int64_t hi = (int64_t)jrNext(r, 26) << 27;
    return (double)(hi + jrNext(r, 27)) * (1.0 / (double)(1LL << 53));
}
PS_HD double jrNextGaussian(JRandom& r) {
    if (r.haveNextNextGaussian) {
        r.haveNextNextGaussian = 0;
        return r.nextNextGaussian;
    }
    double v1, v2, s;
    do {
        v1 = 2 * jrNextDouble(r) - 1;
        v2 = 2 * jrNextDouble(r) - 1;
        s = v1 * v1 + v2 * v2;
    } while (s >= 1 || s == 0);
    double multiplier = sqrt(-2 * log(s) / s);
    r.nextNextGaussian = v2 * multiplier;
    r.haveNextNextGaussian = 1;
    return v1 * multiplier;
}

// RoundedRectangleEnclosure.inFreeSpace (src/arena/RoundedRectangleEnclosure.java:73-93)
PS_HD bool inFreeSpace(const Scene& sc, V2 v, float minDistance) {
    float ax = fabsf(v.x), ay = fabsf(v.y);
    float W = sc.W, H = sc.H, R = sc.R;
    if (ax < W / 2 - R - minDistance && ay < H / 2 - minDistance) return true;
    else if (ax > W / 2 - R && ax < W / 2 - minDistance && ay < H / 2 - R - minDistance) return true;
    else if (dist(v, mk(-sc.cornerX, sc.cornerY)) < R - minDistance) return true;
    else if (dist(v, mk(sc.cornerX, sc.cornerY)) < R - minDistance) return true;
    else if (dist(v, mk(-sc.cornerX, -sc.cornerY)) < R - minDistance) return true;
    else if (dist(v, mk(sc.cornerX, -sc.cornerY)) < R - minDistance) return true;
    return false;
}
// The same predicate with the constants hoisted and one corner test: of the four corner-arc centres the one in the
// point's own quadrant is the nearest (|v.x| - cornerX and -(|v.x| - cornerX) square to the same float, and rounding is
// monotonic), so "within R of any centre" == "within R of that one". Same result for every point.
struct FreeSpace {
    float x1, y1, x2lo, x2hi, y2, cx, cy, r;
};
PS_HD FreeSpace makeFreeSpace(const Scene& sc, float minDistance) {
    FreeSpace f;
    f.x1 = sc.W / 2 - sc.R - minDistance;
    f.y1 = sc.H / 2 - minDistance;
    f.x2lo = sc.W / 2 - sc.R;
    f.x2hi = sc.W / 2 - minDistance;
    f.y2 = sc.H / 2 - sc.R - minDistance;
    f.cx = sc.cornerX;
    f.cy = sc.cornerY;
    f.r = sc.R - minDistance;
    return f;
}
PS_HD bool inFreeSpaceFast(const FreeSpace& f, V2 v) {
    float ax = fabsf(v.x), ay = fabsf(v.y);
    if (ax < f.x1 && ay < f.y1) return true;
    if (ax > f.x2lo && ax < f.x2hi && ay < f.y2) return true;
    return dist(mk(ax, ay), mk(f.cx, f.cy)) < f.r;
}
PS_HD bool randomInside(const Scene& sc, JRandom& rng, float minDistance, V2& out) {
    for (int i = 0; i < 100000; i++) {
        float fx = jrNextFloat(rng);
        float fy = jrNextFloat(rng);
        V2 v = mk(fx * sc.W - sc.W / 2.0f, fy * sc.H - sc.H / 2.0f);
        if (inFreeSpace(sc, v, minDistance)) {
            out = v;
            return true;
        }
    }
    return false;
}

struct ResetOut {
    ps_robot_state* robot;   // [R]
    ps_arena_state* arena;
};

template <class Ctx>
PS_HD void arenaReset(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st, const ResetOut& out, int64_t seed,
                      const ps_robot_state* robotInit /* [2]: plain, informed */, int nInformed) {
    const Dims& d = e.cfg.d;
    const Scene& sc = *e.scene;
    int NB = d.NB, NP = d.NP;
    if (ctx.tid() == 0) {
        JRandom rng;
        jrSetSeed(rng, seed);
        const float PUCK_WIDTH = 6.3f, SRV_LENGTH = 22.0f;
        // pucks: Arena.createPuck (the minX/maxX band is the whole arena for modifiedPuckPlacement = false)
        for (int p = 0; p < d.P; ++p) {
            V2 pos = mk(0, 0);
            float closest = 0.0f;
            do {
                if (!randomInside(sc, rng, PUCK_WIDTH / 2, pos)) break;
                closest = INFINITY;
                for (int o = 0; o < p; ++o) {
                    float dd = dist(mk(wk.cx[o], wk.cy[o]), pos);
                    if (dd < closest) closest = dd;
                }
            } while (closest < PUCK_WIDTH);
            float ang = (float)(2.0 * 3.14159265358979323846 * (double)jrNextFloat(rng));
            wk.cx[p] = pos.x;
            wk.cy[p] = pos.y;
            wk.a[p] = ang;
        }
        // robots: Arena.createRobot (distance to other robots only, at most 1000 attempts)
        for (int r = 0; r < d.R; ++r) {
            V2 pos = mk(0, 0);
            float closest = 0.0f;
            int attempts = 0;
            do {
                if (!randomInside(sc, rng, SRV_LENGTH / 2, pos)) break;
                closest = INFINITY;
                for (int o = 0; o < r; ++o) {
                    float dd = dist(mk(wk.rpx[o], wk.rpy[o]), pos);
                    if (dd < closest) closest = dd;
                }
                attempts++;
            } while (attempts < 1000 && closest < 2 * SRV_LENGTH);
            float ang = (float)(2.0 * 3.14159265358979323846 * (double)jrNextFloat(rng));
            int b = d.P + r;
            // BodyDef -> Body: xf.position = pos, R.set(angle); resetMassData: sweep.c = xf * localCenter
            Xf T;
            T.px = pos.x;
            T.py = pos.y;
            T.c = psCos(e.trig, ang);
            T.s = psSin(e.trig, ang);
            wk.rpx[r] = T.px;
            wk.rpy[r] = T.py;
            wk.rc[r] = T.c;
            wk.rs[r] = T.s;
            V2 c = mulX(T, mk(sc.robot.lcx, sc.robot.lcy));
            wk.cx[b] = c.x;
            wk.cy[b] = c.y;
            wk.a[b] = ang;
        }
        ps_arena_state as;
        as.rng_next_gaussian = rng.nextNextGaussian;
        as.rng_seed = rng.seed;
        as.rng_have_gaussian = rng.haveNextNextGaussian;
        as.update_count = 0;
        as.step_count = 0;
        as.n_contacts = 0;
        as.inv_dt0 = 0.0f;
        as.error_flags = 0;
        *out.arena = as;
        for (int i = 0; i < 8; i++) wk.counters[i] = 0;
    }
    ctx.sync();
    PS_FOR(r, d.R) {
        ps_robot_state rs = robotInit[r < nInformed ? 1 : 0];
        // PS_RNG_PER_ROBOT streams: java.util.Random(seed * 1000003 + 7919 * (r + 1))
        JRandom rr;
        jrSetSeed(rr, seed * 1000003LL + 7919LL * (int64_t)(r + 1));
        rs.rng_seed = rr.seed;
        rs.rng_have_gaussian = 0;
        rs.rng_next_gaussian = 0.0;
        out.robot[r] = rs;
    }
    PS_FOR(b, NB) {
        st.body[0 * NB + b] = wk.cx[b];
        st.body[1 * NB + b] = wk.cy[b];
        st.body[2 * NB + b] = wk.a[b];
        st.body[3 * NB + b] = 0.0f;
        st.body[4 * NB + b] = 0.0f;
        st.body[5 * NB + b] = 0.0f;
    }
    // Fixture.createProxy: tight AABB at the creation transform, fattened by aabbExtension
    PS_FOR(i, NP) {
        int proxy = kWallFixtures + i;
        int b = proxyBody(d, proxy);
        const Shape& s = sc.shapes[proxyShape(d, proxy)];
        Aabb a = shapeAabb(s, bodyXf(e, wk, b));
        bool sensed = b < d.P ? (proxy - bodyFirstProxy(d, b)) == 1 : (proxy - bodyFirstProxy(d, b)) == 0;
        if (sensed) storeAabb(st.senseAabb, NB, b, a);
        a.lx -= kAabbExtension;
        a.ly -= kAabbExtension;
        a.ux += kAabbExtension;
        a.uy += kAabbExtension;
        storeAabb(st.fat, NP, i, a);
        wk.moved[i] = 1;
    }
    ctx.sync();
    physBodyFat(ctx, e, wk, st);
    physFindNewContacts(ctx, e, wk, st, true);
}

// ---- Arena.step's puck perturbations (src/arena/Arena.java:185-192, 318-373) ----
//
// Body.setTransform(pos, 0) of one puck (JBox2D 2.1.2.2 Body.setTransform: sweep collapsed onto the new transform, both
// fixtures re-synchronised with zero displacement, then ContactManager.findNewContacts), by the whole CTA. The puck keeps
// its velocity and its cached contacts: those whose fat AABBs separated die in the next World.step's collide(). Because
// of them "overlapped before" no longer means "has a contact", so the pairs go through the explicit cache check.
template <class Ctx>
PS_HD void teleportPuck(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st, int b, float x, float y) {
    const Dims& d = e.cfg.d;
    const Scene& sc = *e.scene;
    int NB = d.NB, NP = d.NP;
    int first = bodyFirstProxy(d, b);
    ctx.sync();
    if (ctx.tid() == 0) {
        st.body[0 * NB + b] = x;
        st.body[1 * NB + b] = y;
        st.body[2 * NB + b] = 0.0f;
        Xf T;
        T.px = x;
        T.py = y;
        T.c = 1.0f;
        T.s = 0.0f;
        for (int k = 0; k < kPuckFixtures; ++k) {
            int i = first - kWallFixtures + k;
            Aabb a = shapeAabb(sc.shapes[kShapePuck0 + k], T);   // combine(aabb1, aabb2) of two equal boxes
            if (k == 1) storeAabb(st.senseAabb, NB, b, a);
            Aabb fat = loadAabb(st.fat, NP, i);
            bool mv = !aabbContains(fat, a);
            if (mv) {
                a.lx -= kAabbExtension;
                a.ly -= kAabbExtension;
                a.ux += kAabbExtension;
                a.uy += kAabbExtension;   // the displacement prediction adds 2 * (0, 0)
                storeAabb(st.fat, NP, i, a);
            }
            wk.counters[4 + k] = mv ? 1 : 0;
        }
        wk.counters[3] = 0;
    }
    ctx.sync();
    // BroadPhase.updatePairs: each moved proxy queries every fat AABB; addPair's same-body and Filter tests are applied
    // here (they only drop pairs, so their place in the order is free)
    for (int k = 0; k < kPuckFixtures; ++k) {
        if (!wk.counters[4 + k]) continue;
        int q = first + k;
        Aabb fatQ = loadAabb(st.fat, NP, q - kWallFixtures);
        uint32_t filtQ = sc.filt[kShapePuck0 + k];
        PS_FOR(p, kWallFixtures + NP) {
            if (proxyBody(d, p) == b) continue;
            if (!filtCollide(filtQ, sc.filt[proxyShape(d, p)])) continue;
            if (!aabbOverlap(proxyFat(e, st, p), fatQ)) continue;
            int lo = p < q ? p : q, hi = p < q ? q : p;
            int slot = ctx.atomicAddI(&wk.counters[3], 1);
            if (slot < e.cfg.MPAIR) wk.pairs[slot] = ((uint32_t)lo << 16) | (uint32_t)hi;
        }
    }
    ctx.sync();
    if (wk.counters[3] > 0) physAppendPairs(ctx, e, wk, st, true);
}

// The head of Arena.step: shiftPucksThroughCentre when stepCount % puckShiftStep == 0, then shufflePucks when
// stepCount % puckShuffleStep == 0 (a step of -1 = never). Thread 0 draws the new positions in puck order from the arena's
// java.util.Random (a puck's new position depends on the robots and on its own old position only); the pucks are
// then moved one after the other, because each setTransform ends with its own findNewContacts.
template <class Ctx>
PS_HD void arenaPerturb(Ctx& ctx, const PhysEnv& e, const PhysWork& wk, const ArenaState& st, ps_arena_state* as, int shiftStep, int shuffleStep) {
    const Dims& d = e.cfg.d;
    const Scene& sc = *e.scene;
    int NB = d.NB;
    const float PUCK_WIDTH = 6.3f, MIN_BOUNDARY = PUCK_WIDTH / 2;
    int stepCount = as->step_count;
    for (int mode = 0; mode < 2; ++mode) {
        int every = mode == 0 ? shiftStep : shuffleStep;
        if (every == -1 || stepCount % every != 0) continue;
        ctx.sync();
        PS_FOR(r, d.R) {   // robot.body.m_xf.position
            int b = d.P + r;
            Xf T = makeXf(e.trig, st.body[0 * NB + b], st.body[1 * NB + b], st.body[2 * NB + b], sc.robot.lcx, sc.robot.lcy);
            wk.rpx[r] = T.px;
            wk.rpy[r] = T.py;
        }
        ctx.sync();
        if (ctx.tid() == 0) {
            JRandom rng;
            rng.seed = as->rng_seed;
            rng.haveNextNextGaussian = as->rng_have_gaussian;
            rng.nextNextGaussian = as->rng_next_gaussian;
            V2 shift = mk(0.0f, 0.0f);
            if (mode == 0) {
                for (int p = 0; p < d.P; ++p) shift = shift + mk(st.body[0 * NB + p], st.body[1 * NB + p]);
                float l = len(shift);   // Vec2.normalize
                if (!(l < kEpsilon)) {
                    float inv = 1.0f / l;
                    shift.x *= inv;
                    shift.y *= inv;
                }
                shift.x *= -100.0f;
                shift.y *= -100.0f;
            }
            for (int p = 0; p < d.P; ++p) {
                V2 pos = mk(0.0f, 0.0f);
                bool random = true, ok = true;
                if (mode == 0) {
                    pos = mk(st.body[0 * NB + p], st.body[1 * NB + p]) + shift;
                    if (inFreeSpace(sc, pos, MIN_BOUNDARY)) {
                        float closest = INFINITY;
                        for (int r = 0; r < d.R; ++r) {
                            float dd = dist(mk(wk.rpx[r], wk.rpy[r]), pos);
                            if (dd < closest) closest = dd;
                        }
                        random = closest < PUCK_WIDTH;
                    }
                }
                if (random) {   // Arena.getRandomPuckPos
                    float closest = 0.0f;
                    do {
                        if (!randomInside(sc, rng, MIN_BOUNDARY, pos)) {
                            ok = false;
                            break;
                        }
                        closest = INFINITY;
                        for (int r = 0; r < d.R; ++r) {
                            float dd = dist(mk(wk.rpx[r], wk.rpy[r]), pos);
                            if (dd < closest) closest = dd;
                        }
                    } while (closest < PUCK_WIDTH);
                }
                wk.cx[p] = pos.x;
                wk.cy[p] = pos.y;
                wk.bodyFlag[p] = ok ? 1 : 0;
            }
            as->rng_seed = rng.seed;
            as->rng_have_gaussian = rng.haveNextNextGaussian;
            as->rng_next_gaussian = rng.nextNextGaussian;
        }
        ctx.sync();
        for (int p = 0; p < d.P; ++p)
            if (wk.bodyFlag[p]) teleportPuck(ctx, e, wk, st, p, wk.cx[p], wk.cy[p]);
        ctx.sync();
    }
}

}  // namespace ps
