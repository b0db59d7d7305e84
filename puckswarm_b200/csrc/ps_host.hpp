// puckswarm_b200 -- host-side set-up shared by the engine (ps_engine.cu) and the host emulation harness of
// the unit tests: derives the kernel configuration from ps_config, builds the scene tables and the
// MathUtils sin table. Pure set-up: nothing here steps a simulation.
#pragma once
#include <vector>
#include <string>
#include <cstring>
#include <cstdlib>
#include "ps_layout.cuh"
#include "ps_pipeline.cuh"

namespace ps {

struct HostSetup {
    ps_config cfg;
    PhysCfg phys;
    Scene scene;
    std::vector<float> sinLut;     // (float)Math.sin(i * 0.00011f), i < 57120 (MathUtils static initialiser)
    ps_robot_state robotInit[2];   // initial controller state: plain, informed (preset caches)
    HotConsts hot;
    std::string error;

    bool init(const ps_config& c) {
        cfg = c;
        if (c.abi_version != PS_ABI_VERSION) return fail("abi_version mismatch");
        if (c.n_arenas < 1 || c.n_robots < 1 || c.n_pucks < 0 || c.n_puck_types < 1 || c.n_puck_types > PS_MAX_COLOURS)
            return fail("bad arena shape");
        if (c.step_interval < 1 || c.vel_iters < 0 || c.pos_iters < 0 || !(c.dt > 0)) return fail("bad step parameters");
        if (c.controller < PS_CTRL_CACHECONS || c.controller > PS_CTRL_PROBSEEK) return fail("bad controller");
        // stepCount % 0 throws in the reference (Arena.java:189-191); negative periods other than the "never" value are refused
        if (c.puck_shift_step == 0 || c.puck_shift_step < -1 || c.puck_shuffle_step == 0 || c.puck_shuffle_step < -1) return fail("bad perturbation period");
        int P = c.n_pucks / c.n_puck_types * c.n_puck_types;   // Arena.java:91-95: nPucks / nPuckTypes per colour
        Dims d;
        d.P = P;
        d.R = c.n_robots;
        d.NB = P + c.n_robots;
        d.NP = 2 * P + kRobotFixtures * c.n_robots;
        if (kWallFixtures + d.NP > 65535) return fail("too many proxies for 16-bit contact keys");
        phys.d = d;
        int mc = c.max_contacts > 0 ? c.max_contacts : (24 * c.n_pucks + 256 * c.n_robots + 127) / 128 * 128;
        phys.MC = mc;
        // touching contacts are a small fraction of the cached ones (a puck pair caches 4 contacts, at most 1 touches)
        phys.MT = mc / 4 < 65535 ? mc / 4 : 65535;
        if (phys.MT < 64) phys.MT = 64;
        int mp = 256;
        while (mp < mc / 2) mp <<= 1;
        phys.MPAIR = mp;
        phys.MW = mc / 4 > 256 ? mc / 4 : 256;
        {
            const char* wv = getenv("PS_WARP_MIN_CONTACTS");
            phys.warpMinContacts = wv && *wv ? atoi(wv) : kWarpMinContactsDefault;
            const char* lx = getenv("PS_LANEX_MAX_CONTACTS");
            phys.laneXMaxContacts = lx && *lx ? atoi(lx) : 0;   // off by default: a lane's chain of a 27-contact island is the longest of the tick (measured)
        }
        phys.velIters = c.vel_iters;
        phys.posIters = c.pos_iters;
        phys.toiEnabled = c.toi_enabled;
        phys.dt = c.dt;
        phys.damp = clampf(1.0f - c.dt * 10.0f, 0.0f, 1.0f);   // Puck.java:33-34, Robot.java:75-76: damping 10
        sinLut.resize(kLutLength);
        for (int i = 0; i < kLutLength; i++) sinLut[i] = (float)sin((double)(i * kLutPrecision));
        Trig tr;
        tr.lut = sinLut.data();
        tr.useLut = c.sincos_lut;
        build::scene(&scene, c.enclosure_scale, tr);
        hot.puck = scene.puck;
        hot.robot = scene.robot;
        hot.wallXf = scene.wallXf;
        hot.friction = scene.friction;
        hot.pmPuck = scene.puck.mass * scene.puck.invMass;
        hot.piPuck = scene.puck.mass * scene.puck.invI;
        hot.pmRobot = scene.robot.mass * scene.robot.invMass;
        hot.piRobot = scene.robot.mass * scene.robot.invI;
        // initial controller state (CacheConsController.java:205-243, BHDController.java:56-80)
        std::memset(robotInit, 0, sizeof(robotInit));
        for (int k = 0; k < 2; k++) robotInit[k].last_carried_type = -1;
        {
            // presetCaches (informed robots): nTypes = 2, L = 0.75 * 187 / 2
            float arenaWidth = 187.0f;
            int nTypes = 2;
            float L = 0.75f * arenaWidth / 2.0f;
            for (int type = 0; type < nTypes; type++) {
                float angle = (float)(type * (2 * M_PI) / nTypes + M_PI / 4.0f);
                float x = (float)(L * cos((double)angle));
                float y = (float)(L * sin((double)angle));
                robotInit[1].cache_x[type] = x;
                robotInit[1].cache_y[type] = y;
                robotInit[1].cache_valid[type] = 1;
                robotInit[1].remembered[type] = 1000;
            }
        }
        return true;
    }
    bool fail(const char* m) {
        error = m;
        return false;
    }
};

}  // namespace ps
