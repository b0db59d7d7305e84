// ORACLE (test infrastructure only -- never linked into or called by the product path).
//
// C entry points over the CPU restatement (jb2d.h, puckswarm.h, jrandom.h). They mirror the product
// ABI of include/puckswarm_b200.h call for call (orc_X <-> ps_X, same structs, same array layouts) so
// that tests/ can run both from identical states and compare the results field by field, and so that
// bench.py's cpu_baseline / --impl reference leg can time the same work on the host cores.
// PARITY IS UNPINNED (see the headers): this is a restatement, not the JVM reference.
#include "puckswarm.h"
#include <thread>
#include <climits>
#include <limits>
#include <functional>
#include <atomic>
#include <map>

using namespace oracle;

struct orc_batch {
    ps_config cfg;
    Calibration calib;
    bool haveCalib = false;
    std::vector<std::unique_ptr<Arena>> arenas;
    int NB = 0, NP = 0, maxContacts = 0;
    float analysisThreshold = -1.0f, analysisTarget = 50.0f;   // < 0: cfg.cluster_distance_threshold
};

static int deriveMaxContacts(const ps_config& c) {
    if (c.max_contacts > 0) return c.max_contacts;
    int v = 24 * c.n_pucks + 256 * c.n_robots;
    return (v + 127) / 128 * 128;
}

static jb::Fixture* proxyFixture(jb::World& w, int id) { return w.proxies[id]; }

// Dynamic bodies in creation order: pucks then robots (world body 0 is the enclosure).
static jb::Body* dynBody(Arena& a, int i) {
    int P = (int)a.pucks.size();
    return i < P ? a.pucks[i].body : a.robots[i - P].body;
}
// The one fixture of a body that PointSensor can see (hull of a robot, inner disc of a puck).
static jb::Fixture* sensedFixture(jb::Body* b) {
    for (jb::Fixture* f = b->fixtureList; f; f = f->next)
        if (f->userKind != jb::U_HIDDEN) return f;
    return nullptr;
}

static void flushNewFixtures(Arena& a) {
    // World.step starts with findNewContacts when fixtures were added (A.8); doing it eagerly makes the
    // exported state self-contained and does not change any later result.
    if (a.world->newFixture) {
        a.world->findNewContacts();
        a.world->newFixture = false;
    }
}

extern "C" {

orc_batch* orc_create(const ps_config* cfg) {
    if (!cfg || cfg->abi_version != PS_ABI_VERSION || cfg->n_arenas < 1) return nullptr;
    orc_batch* b = new orc_batch();
    b->cfg = *cfg;
    b->NB = cfg->n_pucks / cfg->n_puck_types * cfg->n_puck_types + cfg->n_robots;
    b->NP = 2 * (b->NB - cfg->n_robots) + 14 * cfg->n_robots;
    b->maxContacts = deriveMaxContacts(*cfg);
    return b;
}
void orc_destroy(orc_batch* b) { delete b; }

int orc_load_calibration(orc_batch* b, const float* XrYr, const uint16_t* xy, const uint8_t* flags, int n, int w, int h) {
    b->calib.load(XrYr, xy, flags, n, w, h);
    b->haveCalib = true;
    return PS_OK;
}

int orc_get_dims(orc_batch* b, int32_t* nb, int32_t* np, int32_t* mc, int32_t* ow, int32_t* oh) {
    if (nb) *nb = b->NB;
    if (np) *np = b->NP;
    if (mc) *mc = b->maxContacts;
    if (b->haveCalib) {
        LocalMap lm;
        lm.init(&b->calib);
        if (ow) *ow = lm.width;
        if (oh) *oh = lm.height;
    }
    return PS_OK;
}

int orc_reset(orc_batch* b, const int64_t* seeds) {
    if (!b->haveCalib) return PS_E_NOCALIB;
    b->arenas.clear();
    for (int i = 0; i < b->cfg.n_arenas; i++) {
        b->arenas.emplace_back(new Arena(b->cfg, &b->calib));
        if (b->analysisThreshold >= 0.0f) b->arenas.back()->analysisThreshold = b->analysisThreshold;
        b->arenas.back()->analysisTarget = b->analysisTarget;
        b->arenas.back()->reset(seeds[i]);
        flushNewFixtures(*b->arenas.back());
    }
    return PS_OK;
}

static void parallelFor(int n, int nThreads, const std::function<void(int)>& fn) {
    if (nThreads <= 1 || n <= 1) {
        for (int i = 0; i < n; i++) fn(i);
        return;
    }
    std::atomic<int> next(0);
    std::vector<std::thread> ts;
    for (int t = 0; t < std::min(nThreads, n); t++)
        ts.emplace_back([&]() {
            for (;;) {
                int i = next.fetch_add(1);
                if (i >= n) break;
                fn(i);
            }
        });
    for (auto& t : ts) t.join();
}

int orc_step(orc_batch* b, int nTicks, int nThreads) {
    parallelFor((int)b->arenas.size(), nThreads, [&](int i) {
        for (int t = 0; t < nTicks; t++) b->arenas[i]->tick();
    });
    return PS_OK;
}

int orc_sense(orc_batch* b, int nThreads) {
    parallelFor((int)b->arenas.size(), nThreads, [&](int i) { b->arenas[i]->senseAll(); });
    return PS_OK;
}

// One think interval with host-supplied commands [n_arenas][n_robots]: Robot.move + World.step, step_interval times.
int orc_step_external(orc_batch* b, const float* fwd, const float* torque, int nThreads) {
    int R = b->cfg.n_robots;
    parallelFor((int)b->arenas.size(), nThreads, [&](int i) {
        Arena& a = *b->arenas[i];
        for (int r = 0; r < R; r++) {
            a.robots[r].controller->movement.forwards = fwd[(size_t)i * R + r];
            a.robots[r].controller->movement.torque = torque[(size_t)i * R + r];
        }
        a.perturb();
        if (a.stepCount % 100 == 0) a.storeRecord();
        a.stepCount++;
        for (int t = 0; t < a.cfg.step_interval; t++) {
            a.coast();
            a.updateCount++;
            a.world->step(a.cfg.dt, a.cfg.vel_iters, a.cfg.pos_iters);
        }
    });
    return PS_OK;
}

static void packRobot(const Robot& r, ps_robot_state* o, int controller) {
    std::memset(o, 0, sizeof(*o));
    const Controller* c = r.controller.get();
    o->ctrl_state = c->state;
    o->start_count = c->startCount;
    for (int i = 0; i < 8; i++) o->state_counts[i] = c->stateCounts[i];
    o->cmd_forwards = c->movement.forwards;
    o->cmd_torque = c->movement.torque;
    o->lm_carrying = r.localMap.carrying ? 1 : 0;
    for (int k = 0; k < VFHPlus::N; k++)
        if (c->vfh.binaryHistogram[k] != 0) o->vfh_binary[k >> 5] |= 1u << (k & 31);
    o->vfh_last_sector = c->vfh.lastTurnSector;
    o->vfh_last_angle = c->vfh.lastTurnAngle;
    o->last_carried_type = -1;
    if (controller == PS_CTRL_BHD) {
        const BHDController* bc = static_cast<const BHDController*>(c);
        o->bhd_turn_dir = bc->randomTurnDir;
        o->bhd_turn_time = bc->randomTurnTime;
    } else {
        const CacheConsController* cc = static_cast<const CacheConsController*>(c);
        for (int k = 0; k < PS_MAX_COLOURS; k++) {
            o->cache_valid[k] = cc->cacheValid[k];
            o->cache_x[k] = cc->cacheValid[k] ? cc->cachePoints[k].x : 0.0;
            o->cache_y[k] = cc->cacheValid[k] ? cc->cachePoints[k].y : 0.0;
            o->remembered[k] = cc->rememberedCacheSizes[k];
        }
        o->has_target = cc->haveTarget ? 1 : 0;
        o->target_type = cc->haveTarget ? cc->target.puckType : 0;
        o->target_x = cc->haveTarget ? cc->target.centroid.x : 0;
        o->target_y = cc->haveTarget ? cc->target.centroid.y : 0;
        o->last_carried_type = cc->lastCarriedType;
        if (controller == PS_CTRL_PROBSEEK) {
            const ProbSeekController* pc = static_cast<const ProbSeekController*>(c);
            o->bhd_turn_dir = pc->randomTurnDir;
            o->bhd_turn_time = pc->randomTurnTime;
        }
    }
}
static void unpackRobot(Robot& r, const ps_robot_state* o, int controller) {
    Controller* c = r.controller.get();
    c->state = o->ctrl_state;
    c->startCount = o->start_count;
    for (int i = 0; i < 8; i++) c->stateCounts[i] = o->state_counts[i];
    c->movement.forwards = o->cmd_forwards;
    c->movement.torque = o->cmd_torque;
    r.localMap.carrying = o->lm_carrying != 0;
    for (int k = 0; k < VFHPlus::N; k++) c->vfh.binaryHistogram[k] = (o->vfh_binary[k >> 5] >> (k & 31)) & 1 ? 1.0f : 0.0f;
    c->vfh.lastTurnSector = o->vfh_last_sector;
    c->vfh.lastTurnAngle = o->vfh_last_angle;
    if (controller == PS_CTRL_BHD) {
        BHDController* bc = static_cast<BHDController*>(c);
        bc->randomTurnDir = o->bhd_turn_dir;
        bc->randomTurnTime = o->bhd_turn_time;
    } else {
        CacheConsController* cc = static_cast<CacheConsController*>(c);
        for (int k = 0; k < PS_MAX_COLOURS; k++) {
            cc->cacheValid[k] = o->cache_valid[k] != 0;
            cc->cachePoints[k].x = o->cache_x[k];
            cc->cachePoints[k].y = o->cache_y[k];
            cc->cachePoints[k].theta = 0;
            cc->rememberedCacheSizes[k] = o->remembered[k];
        }
        cc->haveTarget = o->has_target != 0;
        if (cc->haveTarget) cc->target = Cluster::single(Vec2(o->target_x, o->target_y), o->target_type);
        cc->lastCarriedType = o->last_carried_type;
        if (controller == PS_CTRL_PROBSEEK) {
            ProbSeekController* pc = static_cast<ProbSeekController*>(c);
            pc->randomTurnDir = o->bhd_turn_dir;
            pc->randomTurnTime = o->bhd_turn_time;
        }
    }
}

int orc_get_state(orc_batch* b, ps_state_view* v) {
    int NB = b->NB, NP = b->NP, MC = b->maxContacts, R = b->cfg.n_robots;
    for (size_t ai = 0; ai < b->arenas.size(); ai++) {
        Arena& a = *b->arenas[ai];
        flushNewFixtures(a);
        jb::World& w = *a.world;
        if (v->body) {
            float* o = v->body + ai * 6 * NB;
            for (int i = 0; i < NB; i++) {
                jb::Body* bd = dynBody(a, i);
                o[0 * NB + i] = bd->sweep.c.x;
                o[1 * NB + i] = bd->sweep.c.y;
                o[2 * NB + i] = bd->sweep.a;
                o[3 * NB + i] = bd->linearVelocity.x;
                o[4 * NB + i] = bd->linearVelocity.y;
                o[5 * NB + i] = bd->angularVelocity;
            }
        }
        if (v->sense_aabb) {
            float* o = v->sense_aabb + ai * 4 * NB;
            for (int i = 0; i < NB; i++) {
                jb::Fixture* f = sensedFixture(dynBody(a, i));
                o[0 * NB + i] = f->aabb.lower.x;
                o[1 * NB + i] = f->aabb.lower.y;
                o[2 * NB + i] = f->aabb.upper.x;
                o[3 * NB + i] = f->aabb.upper.y;
            }
        }
        if (v->fat_aabb) {
            float* o = v->fat_aabb + ai * 4 * NP;
            for (int i = 0; i < NP; i++) {
                jb::Fixture* f = proxyFixture(w, PS_WALL_FIXTURES + i);
                o[0 * NP + i] = f->fat.lower.x;
                o[1 * NP + i] = f->fat.lower.y;
                o[2 * NP + i] = f->fat.upper.x;
                o[3 * NP + i] = f->fat.upper.y;
            }
        }
        int nc = 0;
        {
            // creation order = world contact list reversed
            std::vector<jb::Contact*> cs;
            for (jb::Contact* c = w.contactList; c; c = c->next) cs.push_back(c);
            std::reverse(cs.begin(), cs.end());
            nc = (int)cs.size();
            if (nc > MC) return PS_E_CAPACITY;
            for (int i = 0; i < nc; i++) {
                jb::Contact* c = cs[i];
                if (v->contact_key) v->contact_key[ai * MC + i] = ((uint32_t)c->fixtureA->proxyId << 16) | (uint32_t)c->fixtureB->proxyId;
                if (v->contact_info) {
                    const jb::Manifold& m = c->manifold;
                    uint32_t info = (uint32_t)m.pointCount;
                    if (m.pointCount > 0) info |= (uint32_t)m.points[0].id.pack() << 8;
                    if (m.pointCount > 1) info |= (uint32_t)m.points[1].id.pack() << 16;
                    v->contact_info[ai * MC + i] = info;
                }
                if (v->contact_impulse) {
                    float* o = v->contact_impulse + (ai * MC + i) * 4;
                    const jb::Manifold& m = c->manifold;
                    o[0] = m.pointCount > 0 ? m.points[0].normalImpulse : 0;
                    o[1] = m.pointCount > 0 ? m.points[0].tangentImpulse : 0;
                    o[2] = m.pointCount > 1 ? m.points[1].normalImpulse : 0;
                    o[3] = m.pointCount > 1 ? m.points[1].tangentImpulse : 0;
                }
            }
            for (int i = nc; i < MC; i++) {
                if (v->contact_key) v->contact_key[ai * MC + i] = 0;
                if (v->contact_info) v->contact_info[ai * MC + i] = 0;
                if (v->contact_impulse) std::memset(v->contact_impulse + (ai * MC + i) * 4, 0, 16);
            }
        }
        if (v->robot)
            for (int r = 0; r < R; r++) {
                ps_robot_state* o = v->robot + ai * R + r;
                packRobot(a.robots[r], o, b->cfg.controller);
                o->rng_seed = a.robots[r].rng->seed;
                o->rng_have_gaussian = a.robots[r].rng->haveNextNextGaussian ? 1 : 0;
                o->rng_next_gaussian = a.robots[r].rng->nextNextGaussian;
            }
        if (v->arena) {
            ps_arena_state* s = v->arena + ai;
            std::memset(s, 0, sizeof(*s));
            s->rng_seed = a.rng.seed;
            s->rng_have_gaussian = a.rng.haveNextNextGaussian ? 1 : 0;
            s->rng_next_gaussian = a.rng.nextNextGaussian;
            s->update_count = a.updateCount;
            s->step_count = a.stepCount;
            s->n_contacts = nc;
            s->inv_dt0 = w.inv_dt0;
        }
    }
    return PS_OK;
}

int orc_set_state(orc_batch* b, const ps_state_view* v) {
    int NB = b->NB, NP = b->NP, MC = b->maxContacts, R = b->cfg.n_robots;
    for (size_t ai = 0; ai < b->arenas.size(); ai++) {
        Arena& a = *b->arenas[ai];
        flushNewFixtures(a);
        jb::World& w = *a.world;
        if (v->body) {
            const float* o = v->body + ai * 6 * NB;
            for (int i = 0; i < NB; i++) {
                jb::Body* bd = dynBody(a, i);
                bd->sweep.c.x = o[0 * NB + i];
                bd->sweep.c.y = o[1 * NB + i];
                bd->sweep.a = o[2 * NB + i];
                bd->sweep.c0 = bd->sweep.c;
                bd->sweep.a0 = bd->sweep.a;
                bd->linearVelocity.x = o[3 * NB + i];
                bd->linearVelocity.y = o[4 * NB + i];
                bd->angularVelocity = o[5 * NB + i];
                bd->synchronizeTransform();
                bd->force = Vec2();
                bd->torque = 0;
            }
        }
        if (v->sense_aabb) {
            const float* o = v->sense_aabb + ai * 4 * NB;
            for (int i = 0; i < NB; i++) {
                jb::Fixture* f = sensedFixture(dynBody(a, i));
                f->aabb.lower.set(o[0 * NB + i], o[1 * NB + i]);
                f->aabb.upper.set(o[2 * NB + i], o[3 * NB + i]);
            }
        }
        if (v->fat_aabb) {
            const float* o = v->fat_aabb + ai * 4 * NP;
            for (int i = 0; i < NP; i++) {
                jb::Fixture* f = proxyFixture(w, PS_WALL_FIXTURES + i);
                f->fat.lower.set(o[0 * NP + i], o[1 * NP + i]);
                f->fat.upper.set(o[2 * NP + i], o[3 * NP + i]);
            }
        }
        if (v->contact_key && v->arena) {
            w.clearContacts();
            int nc = v->arena[ai].n_contacts;
            for (int i = 0; i < nc; i++) {
                uint32_t key = v->contact_key[ai * MC + i];
                jb::Contact* c = w.insertContact(proxyFixture(w, key >> 16), proxyFixture(w, key & 0xFFFF));
                uint32_t info = v->contact_info ? v->contact_info[ai * MC + i] : 0;
                c->manifold.pointCount = info & 3;
                c->touching = c->manifold.pointCount > 0;
                c->manifold.points[0].id = jb::ContactID::unpack((info >> 8) & 0xFF);
                c->manifold.points[1].id = jb::ContactID::unpack((info >> 16) & 0xFF);
                if (v->contact_impulse) {
                    const float* o = v->contact_impulse + (ai * MC + i) * 4;
                    c->manifold.points[0].normalImpulse = o[0];
                    c->manifold.points[0].tangentImpulse = o[1];
                    c->manifold.points[1].normalImpulse = o[2];
                    c->manifold.points[1].tangentImpulse = o[3];
                }
            }
        }
        if (v->robot)
            for (int r = 0; r < R; r++) {
                const ps_robot_state* o = v->robot + ai * R + r;
                unpackRobot(a.robots[r], o, b->cfg.controller);
                a.robots[r].rng->seed = o->rng_seed;
                a.robots[r].rng->haveNextNextGaussian = o->rng_have_gaussian != 0;
                a.robots[r].rng->nextNextGaussian = o->rng_next_gaussian;
            }
        if (v->arena) {
            const ps_arena_state* s = v->arena + ai;
            a.rng.seed = s->rng_seed;
            a.rng.haveNextNextGaussian = s->rng_have_gaussian != 0;
            a.rng.nextNextGaussian = s->rng_next_gaussian;
            a.updateCount = s->update_count;
            a.stepCount = s->step_count;
            w.inv_dt0 = s->inv_dt0;
        }
    }
    return PS_OK;
}

static void packLocalMap(const LocalMap& lm, ps_localmap* o) {
    std::memset(o, 0, sizeof(*o));
    o->carrying = lm.carrying ? 1 : 0;
    o->carried_type = lm.carriedType;
    o->carried_x = lm.carrying ? lm.carriedV.x : 0;
    o->carried_y = lm.carrying ? lm.carriedV.y : 0;
    o->carried_cluster = lm.carriedCluster;
    int n = 0;
    // raw clusters are emitted colour by colour and, within a colour, reference the de-duplicated puck list;
    // map every perceived puck to its cluster by bitwise-equal position within the colour.
    for (int k = 0; k < PS_MAX_COLOURS; k++)
        for (const Vec2& p : lm.pucks[k]) {
            if (n >= PS_MAX_LM_PUCKS) {
                o->overflow = 1;
                break;
            }
            o->puck_x[n] = p.x;
            o->puck_y[n] = p.y;
            o->puck_type[n] = (uint8_t)k;
            o->puck_cluster[n] = 255;
            for (size_t c = 0; c < lm.rawClusters.size(); c++) {
                const Cluster& cl = lm.rawClusters[c];
                if (cl.puckType != k) continue;
                for (size_t vi = 0; vi < cl.verts.size(); vi++)
                    if (vecEquals(cl.verts[vi], p)) {
                        o->puck_cluster[n] = (uint8_t)c;
                        o->puck_degree[n] = (uint8_t)cl.degree[vi];
                    }
            }
            n++;
        }
    o->n_pucks = n;
    int nc = (int)std::min<size_t>(lm.rawClusters.size(), PS_MAX_LM_PUCKS);
    o->n_clusters = nc;
    for (int c = 0; c < nc; c++) {
        const Cluster& cl = lm.rawClusters[c];
        o->cluster_x[c] = cl.centroid.x;
        o->cluster_y[c] = cl.centroid.y;
        o->cluster_type[c] = (uint8_t)cl.puckType;
        o->cluster_size[c] = (uint8_t)std::min(cl.size, 255);
        o->cluster_kept[c] = std::find(lm.clusters.begin(), lm.clusters.end(), c) != lm.clusters.end() ? 1 : 0;
    }
    lm.getFreeerSide(&o->left_sum, &o->right_sum);
}

int orc_get_observations(orc_batch* b, ps_obs_view* v) {
    int R = b->cfg.n_robots;
    int W = b->calib.imageWidth, H = b->calib.imageHeight;
    for (size_t ai = 0; ai < b->arenas.size(); ai++) {
        Arena& a = *b->arenas[ai];
        for (int r = 0; r < R; r++) {
            Robot& rb = a.robots[r];
            size_t ri = ai * R + r;
            if (v->camera) {
                uint8_t* o = v->camera + ri * (size_t)W * H;
                for (int y = 0; y < H; y++)
                    for (int x = 0; x < W; x++) o[(size_t)y * W + x] = rb.image[(size_t)x * H + y];
            }
            if (v->occupancy) std::memcpy(v->occupancy + ri * rb.localMap.occupancy.size(), rb.localMap.occupancy.data(), rb.localMap.occupancy.size());
            if (v->localmap) packLocalMap(rb.localMap, v->localmap + ri);
            if (v->pose) {
                Pose p = a.getPose(rb);
                v->pose[ri * 3 + 0] = (float)p.x;
                v->pose[ri * 3 + 1] = (float)p.y;
                v->pose[ri * 3 + 2] = (float)p.theta;
            }
        }
    }
    return PS_OK;
}

// PositionList.getClusterStats (src/arena/PositionList.java:85-96) over the true puck positions of one
// colour + the percent-completion definition of analysis.Analyze (src/analysis/Analyze.java:187-218).
int orc_compute_stats(orc_batch* b, float threshold, ps_stats* out) {
    std::memset(out, 0, sizeof(*out));
    out->n_arenas = (int64_t)b->arenas.size();
    out->min_steps_to_completion = INT64_MAX;
    out->max_steps_to_completion = -1;
    bool first = true;
    for (auto& ap : b->arenas) {
        Arena& a = *ap;
        double sumMax = 0;
        for (int k = 0; k < a.cfg.n_puck_types; k++) {
            std::vector<Vec2> pos;
            for (const Puck& p : a.pucks)
                if (p.colour == k) pos.push_back(p.body->xf.position);
            std::vector<Cluster> cl;
            LocalMap::extractClusters(pos, 0, cl, threshold);
            int mx = 0;
            for (const Cluster& c : cl) {
                out->cluster_size_hist[k][std::min(c.size, 63)]++;
                mx = std::max(mx, c.size);
            }
            sumMax += mx;
        }
        // Analyze.java:90,197 divides by the configured Arena.nPucks property (not by the number of pucks spawned)
        double pc = a.cfg.n_pucks > 0 ? 100.0 * sumMax / (double)a.cfg.n_pucks : 0.0;
        out->sum_completion += pc;
        double twc = a.analysis.twc_den > 0.0 ? a.analysis.twc_num / a.analysis.twc_den : 0.0;
        out->sum_twc += twc;
        out->n_snapshots += a.analysis.n_snapshots;
        if (a.analysis.steps_to_completion >= 0) {
            out->n_completed++;
            out->sum_steps_to_completion += a.analysis.steps_to_completion;
            out->min_steps_to_completion = std::min<int64_t>(out->min_steps_to_completion, a.analysis.steps_to_completion);
            out->max_steps_to_completion = std::max<int64_t>(out->max_steps_to_completion, a.analysis.steps_to_completion);
        }
        if (first) {
            out->min_completion = out->max_completion = pc;
            out->min_twc = out->max_twc = twc;
            first = false;
        } else {
            out->min_completion = std::min(out->min_completion, pc);
            out->max_completion = std::max(out->max_completion, pc);
            out->min_twc = std::min(out->min_twc, twc);
            out->max_twc = std::max(out->max_twc, twc);
        }
        for (Robot& r : a.robots) {
            out->ctrl_state_hist[r.controller->state & 7]++;
            if (r.localMap.carrying) out->n_carrying++;
            if (a.cfg.controller == PS_CTRL_CACHECONS) {
                CacheConsController* cc = static_cast<CacheConsController*>(r.controller.get());
                for (int k = 0; k < PS_MAX_COLOURS; k++)
                    if (cc->cacheValid[k]) out->n_caches++;
            }
        }
        for (jb::Contact* c = a.world->contactList; c; c = c->next) {
            out->n_contacts++;
            out->n_touching_points += c->manifold.pointCount;
        }
    }
    return PS_OK;
}

// The record the reference writes at the latest storage step (Arena.java:223-243 and the controllers' dumps), same layout as ps_get_snapshot
int orc_get_snapshot(orc_batch* b, ps_snapshot_view* v) {
    int R = b->cfg.n_robots;
    for (size_t i = 0; i < b->arenas.size(); i++) {
        Arena& a = *b->arenas[i];
        size_t P = a.pucks.size();
        if (v->step_count) v->step_count[i] = a.savedStep;
        if (a.savedStep < 0) continue;
        if (v->robot_pose) std::memcpy(v->robot_pose + i * R * 3, a.savedPoses.data(), sizeof(float) * R * 3);
        if (v->puck_xy && P) std::memcpy(v->puck_xy + i * P * 2, a.savedPucks.data(), sizeof(float) * P * 2);
        for (int r = 0; r < R; r++) {
            const Controller* c = a.robots[r].controller.get();
            if (!c || a.cfg.controller == PS_CTRL_EXTERNAL) continue;
            for (int k = 0; k < PS_MAX_COLOURS; k++) {
                bool cc = a.cfg.controller == PS_CTRL_CACHECONS;
                if (v->cache_xy) {
                    v->cache_xy[((i * R + r) * PS_MAX_COLOURS + k) * 2] = cc ? c->savedCache[k][0] : std::numeric_limits<float>::quiet_NaN();
                    v->cache_xy[((i * R + r) * PS_MAX_COLOURS + k) * 2 + 1] = cc ? c->savedCache[k][1] : std::numeric_limits<float>::quiet_NaN();
                }
            }
            if (v->state_counts)
                for (int k = 0; k < 8; k++) v->state_counts[(i * R + r) * 8 + k] = c->savedStateCounts[k];
        }
    }
    return PS_OK;
}

int orc_get_analysis(orc_batch* b, ps_arena_analysis* out) {
    for (size_t i = 0; i < b->arenas.size(); i++) out[i] = b->arenas[i]->analysis;
    return PS_OK;
}

int orc_analysis_config(orc_batch* b, float threshold, float target) {
    b->analysisThreshold = threshold;
    b->analysisTarget = target;
    for (auto& ap : b->arenas) {
        ap->analysisThreshold = threshold;
        ap->analysisTarget = target;
        ap->analysis = ps_arena_analysis{0, -1, 0.0, 0.0, 0.0};
        ap->savedStep = -1;
    }
    return PS_OK;
}

// ---- diagnostics for the parity tests ----
// out[0]=max position iterations of the last World.step, [1]=TOI events, [2]=islands, [3]=cached contacts,
// [4]=touching contacts, [5]=touching manifold points
int orc_step_info(orc_batch* b, int arena, int32_t* out) {
    jb::World& w = *b->arenas[arena]->world;
    out[0] = w.lastPositionIterations;
    out[1] = w.lastToiEvents;
    out[2] = w.lastIslands;
    int nc = 0, nt = 0, npnt = 0;
    for (jb::Contact* c = w.contactList; c; c = c->next) {
        nc++;
        if (c->manifold.pointCount > 0) nt++;
        npnt += c->manifold.pointCount;
    }
    out[3] = nc;
    out[4] = nt;
    out[5] = npnt;
    return PS_OK;
}

// Gauss-Seidel order of the last World.step: contact keys island after island; returns the count
int orc_solve_order(orc_batch* b, int arena, uint32_t* out, int cap) {
    jb::World& w = *b->arenas[arena]->world;
    int n = (int)w.lastOrder.size();
    for (int i = 0; i < n && i < cap; i++) out[i] = w.lastOrder[i];
    return n;
}

// Narrow-phase unit entry: manifold of two proxies of arena 0 at explicit transforms (x, y, angle).
// out: [0]=pointCount [1]=type [2..3]=localNormal [4..5]=localPoint [6..7]=p0.local [8..9]=p1.local [10]=id0 [11]=id1
int orc_manifold(orc_batch* b, int proxyA, float ax, float ay, float aa, int proxyB, float bx, float by, float ba, float* out) {
    jb::World& w = *b->arenas[0]->world;
    jb::Fixture *fA = w.proxies[proxyA], *fB = w.proxies[proxyB];
    jb::Transform xfA, xfB;
    xfA.R.set(aa, w.tr);
    xfA.position = Vec2(ax, ay);
    xfB.R.set(ba, w.tr);
    xfB.position = Vec2(bx, by);
    jb::Contact c;
    c.fixtureA = fA;
    c.fixtureB = fB;
    c.evaluate(xfA, xfB);
    const jb::Manifold& m = c.manifold;
    out[0] = (float)m.pointCount;
    out[1] = (float)m.type;
    out[2] = m.localNormal.x;
    out[3] = m.localNormal.y;
    out[4] = m.localPoint.x;
    out[5] = m.localPoint.y;
    out[6] = m.points[0].localPoint.x;
    out[7] = m.points[0].localPoint.y;
    out[8] = m.points[1].localPoint.x;
    out[9] = m.points[1].localPoint.y;
    out[10] = (float)m.points[0].id.pack();
    out[11] = (float)m.points[1].id.pack();
    return PS_OK;
}

// java.util.Random vectors for the known-answer tests
void orc_random_ints(int64_t seed, int n, int32_t* out) {
    JavaRandom r(seed);
    for (int i = 0; i < n; i++) out[i] = r.nextInt();
}
void orc_random_floats(int64_t seed, int n, float* out) {
    JavaRandom r(seed);
    for (int i = 0; i < n; i++) out[i] = r.nextFloat();
}
void orc_random_gaussians(int64_t seed, int n, double* out) {
    JavaRandom r(seed);
    for (int i = 0; i < n; i++) out[i] = r.nextGaussian();
}
void orc_random_bounded(int64_t seed, int bound, int n, int32_t* out) {
    JavaRandom r(seed);
    for (int i = 0; i < n; i++) out[i] = r.nextInt(bound);
}
float orc_sin_lut(float x) { return jb::sinLUT(x); }

// Mass data of the robot and puck bodies (known answers of SURVEY App. B): out = mass, I (about COM), localCenter.x, .y
int orc_body_mass(orc_batch* b, int dynIndex, float* out) {
    jb::Body* bd = dynBody(*b->arenas[0], dynIndex);
    out[0] = bd->mass;
    out[1] = bd->I;
    out[2] = bd->sweep.localCenter.x;
    out[3] = bd->sweep.localCenter.y;
    return PS_OK;
}

}  // extern "C"
