// Exercises the C++ host mirror (puckswarm_b200/host/puckswarm_host.hpp) the way the reference's own driver would use the
// Java classes: an Experiment from .properties files, an Arena batch, RunOffline.step in a loop, a Suite inspected, and a
// scripted host Controller through the batched callback path. Prints one line per arena for tests/test_host_cpp.py to compare
// with the Python binding. Usage: host_demo <properties dir> <code> <calibration.bin> <n arenas> <ticks> [external]
#include <cinttypes>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../puckswarm_b200/host/puckswarm_host.hpp"

using namespace puckswarm;

struct Spin : Controller {   // TestMovementController-style scripted controller
    int seen = 0;
    void computeDesired(Suite& suite, int stepCount) override {
        (void)stepCount;
        seen = suite.getLocalMap().n_pucks + (suite.getPose().getTheta() > 1e9 ? 1 : 0);
    }
    float getForwards() override { return 300000.0f; }
    float getTorque() override { return 1000000.0f; }
};

// --config <dir> <code>: the ps_config the Experiment maps to, one field per line (needs no GPU)
static int printConfig(const char* dir, const char* code) {
    ps_config c = Experiment::fromFiles(dir, code, 0).toConfig(1);
    std::printf("controller=%d n_robots=%d n_pucks=%d n_puck_types=%d n_informed_robots=%d puck_shift_step=%d puck_shuffle_step=%d\n", c.controller,
                c.n_robots, c.n_pucks, c.n_puck_types, c.n_informed_robots, c.puck_shift_step, c.puck_shuffle_step);
    std::printf("cc_k1=%.9g cc_k2=%.9g cc_pick_up_time=%d cc_cache_selection=%d cc_cache_separation=%d cc_ignore_pucks=%d cc_forget_prob=%.9g\n", c.cc_k1,
                c.cc_k2, c.cc_pick_up_time, c.cc_cache_selection, c.cc_cache_separation, c.cc_ignore_pucks, c.cc_forget_prob);
    std::printf("vfh_tau_lo=%.9g vfh_tau_hi=%.9g enclosure_scale=%.9g bhd_backup_time=%d ps_placement_time=%d\n", c.vfh_tau_lo, c.vfh_tau_hi,
                c.enclosure_scale, c.bhd_backup_time, c.ps_placement_time);
    return 0;
}

int main(int argc, char** argv) {
    if (argc == 4 && std::strcmp(argv[1], "--config") == 0) {
        try {
            return printConfig(argv[2], argv[3]);
        } catch (const std::exception& e) {
            std::fprintf(stderr, "error: %s\n", e.what());
            return 4;
        }
    }
    if (argc == 2 && std::strcmp(argv[1], "--format") == 0) {   // numbers on stdin -> Float.toString / Double.toString of the float
        double d;
        while (std::scanf("%lf", &d) == 1) std::printf("%s %s\n", javaFloatStr((float)d).c_str(), javaDoubleStr((double)(float)d).c_str());
        return 0;
    }
    if (argc < 6) {
        std::fprintf(stderr, "usage: host_demo <dir> <code> <calib.bin> <arenas> <ticks> [external]\n");
        return 2;
    }
    try {
        int nArenas = std::atoi(argv[4]), ticks = std::atoi(argv[5]);
        bool external = argc > 6 && std::strcmp(argv[6], "external") == 0;
        Experiment exp = Experiment::fromFiles(argv[1], argv[2], 0);
        Calibration calib = Calibration::load(argv[3]);
        std::vector<int64_t> seeds;
        for (int i = 0; i < nArenas; ++i) seeds.push_back(i);
        Arena::Controllers ctrls;
        if (external) {
            int R = exp.getProperty("Arena.nRobots", 4);
            ctrls.resize(nArenas);
            for (auto& row : ctrls)
                for (int r = 0; r < R; ++r) row.push_back(std::make_shared<Spin>());
        }
        Arena arena(exp, seeds, calib, 0, ctrls);
        RunOffline run(arena);
        for (int t = 0; t < ticks; ++t) run.step();
        if (const char* snap = std::getenv("HOST_DEMO_SNAPSHOT_DIR")) arena.saveSnapshot(snap);
        std::vector<float> b = arena.bodies();
        int NB = arena.nBodies();
        std::printf("code %s stepCount %d nBodies %d robot0 %s pucksSeen %d\n", exp.code().c_str(), arena.getStepCount(), NB,
                    arena.getSuite(0, 0).getRobotName().c_str(), arena.getSuite(0, 0).getLocalMap().n_pucks);
        for (int a = 0; a < nArenas; ++a) {
            // bit patterns of the first robot's pose and the first puck's position: compared exactly by the test
            const float* base = b.data() + (size_t)a * 6 * NB;
            int r0 = NB - arena.config().n_robots;
            uint32_t w[5];
            float v[5] = {base[0 * NB + r0], base[1 * NB + r0], base[2 * NB + r0], base[0 * NB + 0], base[1 * NB + 0]};
            std::memcpy(w, v, sizeof(w));
            std::printf("arena %d %08x %08x %08x %08x %08x\n", a, w[0], w[1], w[2], w[3], w[4]);
        }
        arena.dispose();
    } catch (const EngineError& e) {
        std::fprintf(stderr, "engine error %d: %s\n", e.code, e.what());
        return 3;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 4;
    }
    return 0;
}
