// Host emulation harness (test infrastructure only): instantiates the engine's per-arena device programs
// (puckswarm_b200/csrc/*.cuh) with HostCtx -- one virtual thread, no barriers -- so that the CPU-only test suite
// can check the device logic against the oracle without a GPU. It is NOT part of the product library and is not
// reachable from any ps_* entry point.
#define PS_HOST_EMU 1
#include "../../puckswarm_b200/csrc/ps_host.hpp"
#include "../../puckswarm_b200/csrc/ps_sense.cuh"
#include "../../puckswarm_b200/csrc/ps_pipeline.cuh"
#include <cstdlib>
#include <memory>

using namespace ps;

struct Emu {
    HostSetup hs;
    std::vector<float> body, senseAabb, fat, cimp;
    std::vector<uint32_t> ckey, cinfo;
    std::vector<ps_robot_state> robot;
    std::vector<ps_arena_state> arena;
    std::vector<char> slow;
    PhysWork wk;
    PhysEnv env;
    StatePtrs sp;
    SenseSetup sense;
    std::vector<char> senseBuf;
    std::vector<int32_t> lastStats;
    std::vector<std::vector<uint32_t>> lastOrder;
};

extern "C" {

Emu* emu_create(const ps_config* cfg) {
    std::unique_ptr<Emu> e(new Emu());
    if (!e->hs.init(*cfg)) return nullptr;
    const PhysCfg& pc = e->hs.phys;
    size_t A = cfg->n_arenas;
    e->body.assign(A * 6 * pc.d.NB, 0);
    e->senseAabb.assign(A * 4 * pc.d.NB, 0);
    e->fat.assign(A * 4 * pc.d.NP, 0);
    e->ckey.assign(A * pc.MC, 0);
    e->cinfo.assign(A * pc.MC, 0);
    e->cimp.assign(A * pc.MC * 4, 0);
    e->robot.assign(A * pc.d.R, ps_robot_state());
    e->arena.assign(A, ps_arena_state());
    Carver m = {nullptr, 0, 0, nullptr, 0};
    carveWork(pc, m);
    e->slow.assign(m.slowUsed + 64, 0);
    Carver cv = {nullptr, 0, 0, e->slow.data(), 0};
    e->wk = carveWork(pc, cv);
    e->env.scene = &e->hs.scene;
    e->env.trig.lut = e->hs.sinLut.data();
    e->env.trig.useLut = cfg->sincos_lut;
    e->env.cfg = pc;
    e->env.hot = e->hs.hot;
    e->sp.body = e->body.data();
    e->sp.senseAabb = e->senseAabb.data();
    e->sp.fat = e->fat.data();
    e->sp.ckey = e->ckey.data();
    e->sp.cinfo = e->cinfo.data();
    e->sp.cimp = e->cimp.data();
    e->sp.robot = e->robot.data();
    e->sp.arena = e->arena.data();
    e->lastStats.assign(A * 8, 0);
    e->lastOrder.resize(A);
    return e.release();
}
void emu_destroy(Emu* e) { delete e; }

int emu_get_dims(Emu* e, int32_t* nb, int32_t* np, int32_t* mc, int32_t* ow, int32_t* oh) {
    *nb = e->hs.phys.d.NB;
    *np = e->hs.phys.d.NP;
    *mc = e->hs.phys.MC;
    *ow = e->sense.occW;
    *oh = e->sense.occH;
    return 0;
}

int emu_load_calibration(Emu* e, const float* XrYr, const uint16_t* xy, const uint8_t* flags, int n, int w, int h) {
    if (!e->sense.build(e->hs.cfg, XrYr, xy, flags, n, w, h)) return PS_E_INVALID;
    return 0;
}

int emu_reset(Emu* e, const int64_t* seeds) {
    HostCtx ctx;
    for (int a = 0; a < e->hs.cfg.n_arenas; a++) {
        ArenaState st = arenaState(e->sp, e->hs.phys, a);
        ResetOut out = {e->sp.robot + (size_t)a * e->hs.phys.d.R, e->sp.arena + a};
        arenaReset(ctx, e->env, e->wk, st, out, seeds[a], e->hs.robotInit, e->hs.cfg.n_informed_robots);
    }
    return 0;
}

// RunOffline.step x nTicks: {Arena.step | Arena.coast} + World.step
int emu_step(Emu* e, int nTicks) {
    HostCtx ctx;
    int R = e->hs.phys.d.R;
    std::vector<float> fwd(R), tq(R);
    for (int a = 0; a < e->hs.cfg.n_arenas; a++) {
        ArenaState st = arenaState(e->sp, e->hs.phys, a);
        ps_robot_state* rs = e->sp.robot + (size_t)a * R;
        ps_arena_state* as = e->sp.arena + a;
        for (int t = 0; t < nTicks; t++) {
            if (as->update_count % e->hs.cfg.step_interval == 0) {
                arenaPerturb(ctx, e->env, e->wk, st, as, e->hs.cfg.puck_shift_step, e->hs.cfg.puck_shuffle_step);
                if (e->hs.cfg.controller != PS_CTRL_EXTERNAL) {
                    if (!e->sense.ready) return PS_E_NOCALIB;
                    senseAndThinkArena(ctx, e->env, e->sense, e->senseBuf, st, rs, as, nullptr);   // counts the think step
                } else {
                    as->step_count++;
                }
            }
            for (int r = 0; r < R; r++) {
                fwd[r] = rs[r].cmd_forwards;
                tq[r] = rs[r].cmd_torque;
            }
            as->update_count++;
            if (getenv("PS_EMU_FUSED")) {
                worldStep(ctx, e->env, e->wk, st, fwd.data(), tq.data());
            } else {
                // the three-stage pipeline of the engine, run for one arena: pre, every queued island, post
                std::vector<unsigned> qe((size_t)kQueueBuckets * e->hs.phys.d.NB);
                int qc[kQueueBuckets + 1] = {0};
                IslandQueue q = {qe.data(), qc, e->hs.phys.d.NB};
                physPre(ctx, e->env, e->wk, st, fwd.data(), tq.data(), q, 0);
                for (int bucket = 0; bucket < kQueueBuckets; bucket++)
                    for (int i = 0; i < qc[bucket]; i++) {
                        int k = (int)(qe[(size_t)bucket * q.cap + i] % (unsigned)e->hs.phys.d.NB);
                        IslandCursor cur;
                        islandBegin(cur, e->wk, k);
                        int iters = 0;
                        while (islandStep(e->env, e->wk, st, cur, &iters)) {
                        }
                        if (iters > e->wk.stats[0]) e->wk.stats[0] = iters;
                    }
                physPost(ctx, e->env, e->wk, st);
            }
            for (int i = 0; i < 8; i++) e->lastStats[a * 8 + i] = e->wk.stats[i];
            e->lastOrder[a].clear();
            for (int i = 0; i < e->wk.islandStart[e->wk.counters[1]]; i++) {
                const TC& c = e->wk.tc[e->wk.order[i]];
                e->lastOrder[a].push_back(c.key);
            }
        }
    }
    return 0;
}

int emu_sense(Emu* e, ps_obs_view* obs) {
    HostCtx ctx;
    int R = e->hs.phys.d.R;
    if (!e->sense.ready) return PS_E_NOCALIB;
    for (int a = 0; a < e->hs.cfg.n_arenas; a++) {
        ArenaState st = arenaState(e->sp, e->hs.phys, a);
        senseArenaToObs(ctx, e->env, e->sense, e->senseBuf, st, e->sp.robot + (size_t)a * R, a, obs);
    }
    return 0;
}

int emu_solve_order(Emu* e, int arena, uint32_t* out, int cap) {
    int n = (int)e->lastOrder[arena].size();
    for (int i = 0; i < n && i < cap; i++) out[i] = e->lastOrder[arena][i];
    return n;
}
int emu_step_info(Emu* e, int arena, int32_t* out) {
    for (int i = 0; i < 8; i++) out[i] = e->lastStats[arena * 8 + i];
    return 0;
}

static void copyState(Emu* e, ps_state_view* v, bool toView) {
    const PhysCfg& pc = e->hs.phys;
    size_t A = e->hs.cfg.n_arenas;
#define CP(field, vec, count)                                                                      \
    if (v->field) {                                                                                \
        if (toView) std::memcpy(v->field, e->vec.data(), (count) * sizeof(e->vec[0]));             \
        else std::memcpy(e->vec.data(), v->field, (count) * sizeof(e->vec[0]));                    \
    }
    CP(body, body, A * 6 * pc.d.NB)
    CP(sense_aabb, senseAabb, A * 4 * pc.d.NB)
    CP(fat_aabb, fat, A * 4 * pc.d.NP)
    CP(contact_key, ckey, A * pc.MC)
    CP(contact_info, cinfo, A * pc.MC)
    CP(contact_impulse, cimp, A * pc.MC * 4)
    CP(robot, robot, A * pc.d.R)
    CP(arena, arena, A)
#undef CP
}
int emu_get_state(Emu* e, ps_state_view* v) {
    copyState(e, v, true);
    return 0;
}
int emu_set_state(Emu* e, const ps_state_view* v) {
    copyState(e, const_cast<ps_state_view*>(v), false);
    return 0;
}

// unit entry: manifold between two proxies at explicit transforms (same output layout as orc_manifold)
int emu_manifold(Emu* e, int proxyA, float ax, float ay, float aa, int proxyB, float bx, float by, float ba, float* out) {
    const Dims& d = e->hs.phys.d;
    Xf xfA, xfB;
    xfA.px = ax;
    xfA.py = ay;
    xfA.c = psCos(e->env.trig, aa);
    xfA.s = psSin(e->env.trig, aa);
    xfB.px = bx;
    xfB.py = by;
    xfB.c = psCos(e->env.trig, ba);
    xfB.s = psSin(e->env.trig, ba);
    Manifold m;
    std::memset(&m, 0, sizeof(m));
    evaluateManifold(m, e->hs.scene.shapes[proxyShape(d, proxyA)], xfA, e->hs.scene.shapes[proxyShape(d, proxyB)], xfB);
    out[0] = (float)m.pointCount;
    out[1] = (float)m.type;
    out[2] = m.localNormal.x;
    out[3] = m.localNormal.y;
    out[4] = m.localPoint.x;
    out[5] = m.localPoint.y;
    out[6] = m.p[0].x;
    out[7] = m.p[0].y;
    out[8] = m.p[1].x;
    out[9] = m.p[1].y;
    out[10] = (float)m.id[0];
    out[11] = (float)m.id[1];
    return 0;
}
int emu_body_mass(Emu* e, int robot, float* out) {
    const BodyConst& k = robot ? e->hs.scene.robot : e->hs.scene.puck;
    out[0] = k.mass;
    out[1] = k.I;
    out[2] = k.lcx;
    out[3] = k.lcy;
    return 0;
}

}  // extern "C"
