// puckswarm_b200 -- C++ host-side mirror of the reference's interfaces for the per-tick path, above the C-ABI of
// include/puckswarm_b200.h (header-only, C++17; link with -lpuckswarm_b200).
//
// The reference's host is Java and this image has no JDK, so the compiled-language host that can be built and run here is
// this header; java/ holds the JNI shim a maintainer compiles and puckswarm_b200/host.py is the same
// mirror for Python callers. Names, argument meaning and error behaviour follow the reference:
//   experiment.Experiment                      src/experiment/Experiment.java:21-90 (getProperty: missing key -> default,
//                                              malformed number -> NumberFormatException, here std::invalid_argument)
//   arena.Arena                                src/arena/Arena.java:45,183-267,310-316,392-411
//   arena.RunOffline.step                      src/arena/RunOffline.java:72-95
//   sensors.Suite / APS, localmap.Pose         src/sensors/Suite.java:5-43, src/sensors/SimAPS.java:14-18
//   controllers.Controller                     src/controllers/Controller.java:9-31
// One Arena object stands for a BATCH of independent arenas (one per seed); accessors take the arena index. All simulation
// work happens in the CUDA engine behind the C-ABI: nothing here computes physics, sensing or control, and there is no CPU
// path (ps_create fails with PS_E_NODEVICE without a B200).
#pragma once
#include <cctype>
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <sys/stat.h>
#include <map>
#include <algorithm>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/puckswarm_b200.h"

namespace puckswarm {

struct EngineError : std::runtime_error {   // the reference has no error returns (System.exit(-1), Arena.java:65-68); the engine reports
    int code;
    EngineError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
inline void check(int rc) {
    if (rc != PS_OK) throw EngineError(rc, ps_last_error());
}

// java.util.Properties.load, as far as the experiment files use it: key=value / key:value lines, '#' and '!' comments
inline std::map<std::string, std::string> loadProperties(const std::string& path) {
    std::map<std::string, std::string> out;
    std::ifstream f(path);
    std::string line;
    auto trim = [](std::string s) {
        size_t a = 0, b = s.size();
        while (a < b && std::isspace((unsigned char)s[a])) ++a;
        while (b > a && std::isspace((unsigned char)s[b - 1])) --b;
        return s.substr(a, b - a);
    };
    while (std::getline(f, line)) {
        std::string t = trim(line);
        if (t.empty() || t[0] == '#' || t[0] == '!') continue;
        size_t k = t.find_first_of("=:");
        if (k == std::string::npos) continue;
        out[trim(t.substr(0, k))] = trim(t.substr(k + 1));
    }
    return out;
}

// Float.toString(float) / Double.toString(double): the shortest decimal that round-trips, in Java's layout -- plain notation
// with at least one digit after the point for 1e-3 <= |x| < 1e7, otherwise d.dddE<exp> (FileUtils / PositionList.save /
// PoseList.save write their numbers this way: PositionList.java:36-49, PoseList.java:32-46)
template <class T>
inline std::string javaNumberStr(T x) {
    if (std::isnan(x)) return "NaN";
    if (std::isinf(x)) return x > 0 ? "Infinity" : "-Infinity";
    if (x == 0) return std::signbit(x) ? "-0.0" : "0.0";
    char buf[64];
    auto res = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);   // shortest round-trip digits: d[.ddd]e[+-]XX
    std::string sci(buf, res.ptr);
    size_t e = sci.find('e');
    std::string mant = sci.substr(0, e);
    int exp10 = std::stoi(sci.substr(e + 1));
    bool neg = mant[0] == '-';
    if (neg) mant = mant.substr(1);
    std::string digits;
    for (char c : mant)
        if (c != '.') digits.push_back(c);
    double a = std::fabs((double)x);
    std::string out;
    if (a >= 1e-3 && a < 1e7) {
        if (exp10 >= 0) {
            std::string ip = digits.substr(0, std::min(digits.size(), (size_t)exp10 + 1));
            while ((int)ip.size() < exp10 + 1) ip.push_back('0');
            std::string fp = digits.size() > (size_t)exp10 + 1 ? digits.substr(exp10 + 1) : "0";
            out = ip + "." + fp;
        } else {
            out = "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
        }
    } else {
        out = digits.substr(0, 1) + "." + (digits.size() > 1 ? digits.substr(1) : "0") + "E" + std::to_string(exp10);
    }
    return neg ? "-" + out : out;
}
inline std::string javaFloatStr(float x) { return javaNumberStr<float>(x); }
inline std::string javaDoubleStr(double x) { return javaNumberStr<double>(x); }

// experiment.Experiment: one (code, seed)
class Experiment {
public:
    Experiment(int64_t seed = 0, std::map<std::string, std::string> properties = {}, std::string code = "default")
        : seed_(seed), props_(std::move(properties)), code_(std::move(code)) {
        auto it = props_.find("code");
        if (it != props_.end()) code_ = it->second;
    }
    // common.properties then <code>.properties, the latter overriding (Experiment.java:26-33)
    static Experiment fromFiles(const std::string& outputDir, const std::string& code, int64_t seed) {
        auto p = loadProperties(outputDir + "/common.properties");
        for (auto& kv : loadProperties(outputDir + "/" + code + ".properties")) p[kv.first] = kv.second;
        return Experiment(seed, p, code);
    }
    int getProperty(const std::string& key, int dflt) const {
        auto it = props_.find(key);
        return it == props_.end() ? dflt : parseInt(key, it->second);
    }
    float getProperty(const std::string& key, float dflt) const {
        auto it = props_.find(key);
        return it == props_.end() ? dflt : parseFloat(key, it->second);
    }
    bool getProperty(const std::string& key, bool dflt) const {   // Boolean.parseBoolean
        auto it = props_.find(key);
        if (it == props_.end()) return dflt;
        std::string v = it->second;
        for (auto& c : v) c = (char)std::tolower((unsigned char)c);
        return v == "true";
    }
    std::string getProperty(const std::string& key, const std::string& dflt) const {
        auto it = props_.find(key);
        return it == props_.end() ? dflt : it->second;
    }
    int64_t getIndex() const { return seed_; }   // ExperimentBlock.java:27-28: the seed is the repetition index
    const std::string& code() const { return code_; }

    // ps_config carrying every property the hot path reads; unknown keys are ignored like in the reference
    ps_config toConfig(int nArenas) const {
        ps_config c;
        ps_config_defaults(&c);
        c.n_arenas = nArenas;
        std::string ctrl = getProperty("controllerType", std::string("CacheCons"));
        if (ctrl == "CacheCons") c.controller = PS_CTRL_CACHECONS;
        else if (ctrl == "BHD") c.controller = PS_CTRL_BHD;
        else if (ctrl == "ProbSeek") c.controller = PS_CTRL_PROBSEEK;
        else throw std::invalid_argument("controllerType: unknown value " + ctrl);
        c.n_pucks = getProperty("Arena.nPucks", c.n_pucks);
        c.n_puck_types = getProperty("Arena.nPuckTypes", c.n_puck_types);
        c.n_robots = getProperty("Arena.nRobots", c.n_robots);
        c.n_informed_robots = getProperty("Arena.nInformedRobots", c.n_informed_robots);
        c.puck_shift_step = getProperty("Arena.puckShiftStep", c.puck_shift_step);
        c.puck_shuffle_step = getProperty("Arena.puckShuffleStep", c.puck_shuffle_step);
        c.enclosure_scale = getProperty("RoundedRectangleEnclosure.scale", c.enclosure_scale);
        c.carry_threshold_lo = getProperty("LocalMap.CARRY_THRESHOLD_LO", c.carry_threshold_lo);
        c.carry_threshold_hi = getProperty("LocalMap.CARRY_THRESHOLD_HI", c.carry_threshold_hi);
        c.vfh_safety_gamma = getProperty("VFHPlus.SAFETY_DISTANCE_GAMMA", c.vfh_safety_gamma);
        c.vfh_safety_mask = getProperty("VFHPlus.SAFETY_DISTANCE_MASK", c.vfh_safety_mask);
        c.vfh_tau_lo = getProperty("VFHPlus.TAU_LO", c.vfh_tau_lo);
        c.vfh_tau_hi = getProperty("VFHPlus.TAU_HI", c.vfh_tau_hi);
        // ProbSeek reads the same key names under its own class prefix (ProbSeekController.java:98-121)
        std::string p = c.controller == PS_CTRL_PROBSEEK ? "ProbSeekController." : "CacheConsController.";
        c.cc_k1 = getProperty(p + "K1", c.cc_k1);
        c.cc_k2 = getProperty(p + "K2", c.cc_k2);
        c.cc_pick_up_time = getProperty(p + "PICK_UP_TIME", c.cc_pick_up_time);
        c.cc_push_in_time = getProperty(p + "PUSH_IN_TIME", c.cc_push_in_time);
        c.cc_backup_time = getProperty(p + "BACKUP_TIME", c.cc_backup_time);
        c.cc_prediction_distance_sqd = getProperty(p + "PREDICTION_DISTANCE_SQD", c.cc_prediction_distance_sqd);
        c.cc_st_dev = getProperty(p + "ST_DEV", c.cc_st_dev);
        c.ps_placement_time = getProperty("ProbSeekController.PLACEMENT_TIME", c.ps_placement_time);
        if (c.controller != PS_CTRL_PROBSEEK) {
            c.cc_exile_time = getProperty("CacheConsController.EXILE_TIME", c.cc_exile_time);
            c.cc_spit_out_time = getProperty("CacheConsController.SPIT_OUT_TIME", c.cc_spit_out_time);
            c.cc_k_absolute = getProperty("CacheConsController.K_ABSOLUTE", c.cc_k_absolute);
            c.cc_k_relative = getProperty("CacheConsController.K_RELATIVE", c.cc_k_relative);
            c.cc_forget_prob = getProperty("CacheConsController.FORGET_PROB", c.cc_forget_prob);
            c.cc_home_separation_distance = getProperty("CacheConsController.HOME_SEPARATION_DISTANCE", c.cc_home_separation_distance);
            c.cc_ignore_pucks = getProperty("CacheConsController.IGNORE_PUCKS", c.cc_ignore_pucks != 0) ? 1 : 0;
            c.cc_cache_selection = enumOf("CacheConsController.CACHE_SELECTION_OPTION", c.cc_cache_selection,
                                          {{"PROB_ABSOLUTE", PS_SEL_PROB_ABSOLUTE}, {"PROB_RELATIVE", PS_SEL_PROB_RELATIVE},
                                           {"RELATIVE", PS_SEL_RELATIVE}, {"RELATIVE_FORGET", PS_SEL_RELATIVE_FORGET}});
            c.cc_cache_separation = enumOf("CacheConsController.CACHE_SEPARATION_OPTION", c.cc_cache_separation,
                                           {{"IGNORE", PS_SEP_IGNORE}, {"NULLIFY_OLD", PS_SEP_NULLIFY_OLD},
                                            {"ACCEPT_ISOLATED", PS_SEP_ACCEPT_ISOLATED}, {"ACCEPT_LARGER", PS_SEP_ACCEPT_LARGER}});
        }
        c.bhd_push_in_time = getProperty("BHDController.PUSH_IN_TIME", c.bhd_push_in_time);
        c.bhd_backup_time = getProperty("BHDController.BACKUP_TIME", c.bhd_backup_time);
        return c;
    }

private:
    static int parseInt(const std::string& key, const std::string& v) {   // Integer.parseInt: the whole string or NumberFormatException
        size_t n = 0;
        int x = 0;
        try {
            x = std::stoi(v, &n);
        } catch (const std::exception&) {
            n = std::string::npos;
        }
        if (n != v.size()) throw std::invalid_argument(key + ": For input string: \"" + v + "\"");
        return x;
    }
    static float parseFloat(const std::string& key, const std::string& v) {
        size_t n = 0;
        float x = 0;
        try {
            x = std::stof(v, &n);
        } catch (const std::exception&) {
            n = std::string::npos;
        }
        if (n != v.size()) throw std::invalid_argument(key + ": For input string: \"" + v + "\"");
        return x;
    }
    int enumOf(const std::string& key, int dflt, const std::map<std::string, int>& names) const {
        auto it = props_.find(key);
        if (it == props_.end()) return dflt;
        auto e = names.find(it->second);
        if (e == names.end()) throw std::invalid_argument(key + ": unknown value " + it->second);   // Enum.valueOf: IllegalArgumentException
        return e->second;
    }
    int64_t seed_;
    std::map<std::string, std::string> props_;
    std::string code_;
};

// The packed grid-camera calibration (sensors.GridCalibration, GridCalibration.java:53-114): a flat file written by
// puckswarm_b200.abi.write_calibration_bin: int32 n, w, h; float XrYr[n][2]; uint16 xy[n][2]; uint8 flags[n]
struct Calibration {
    int n = 0, width = 0, height = 0;
    std::vector<float> XrYr;
    std::vector<uint16_t> xy;
    std::vector<uint8_t> flags;
    static Calibration load(const std::string& path) {
        Calibration c;
        std::ifstream f(path, std::ios::binary);
        int32_t hdr[3] = {0, 0, 0};
        f.read((char*)hdr, sizeof(hdr));
        if (!f || hdr[0] <= 0) throw std::runtime_error("calibration file not readable: " + path);
        c.n = hdr[0];
        c.width = hdr[1];
        c.height = hdr[2];
        c.XrYr.resize((size_t)c.n * 2);
        c.xy.resize((size_t)c.n * 2);
        c.flags.resize((size_t)c.n);
        f.read((char*)c.XrYr.data(), c.XrYr.size() * sizeof(float));
        f.read((char*)c.xy.data(), c.xy.size() * sizeof(uint16_t));
        f.read((char*)c.flags.data(), c.flags.size());
        if (!f) throw std::runtime_error("calibration file truncated: " + path);
        return c;
    }
};

struct Pose {   // localmap.Pose (Pose.java:31-55)
    double x = 0, y = 0, theta = 0;
    double getX() const { return x; }
    double getY() const { return y; }
    double getTheta() const { return theta; }
};

class Arena;
class Suite;

// controllers.Controller (Controller.java:9-31); getForwards / getTorque are valid after computeDesired
class Controller {
public:
    virtual ~Controller() {}
    virtual void computeDesired(Suite& suite, int stepCount) = 0;
    virtual float getForwards() = 0;
    virtual float getTorque() = 0;
    virtual std::string getInfoString() { return ""; }
};

// sensors.Suite of robot `robot` of arena `arena` (Suite.java:5-43); views into the arena's latest observations
class Suite {
public:
    Suite(Arena* a, int arena, int robot) : a_(a), arena_(arena), robot_(robot) {}
    std::string getRobotName() const { return "R" + std::to_string(robot_); }
    const ps_localmap& getLocalMap() const;
    const uint8_t* getOccupancy() const;       // [occ_w * occ_h], cell (i, j) at i * occ_h + j
    const uint8_t* getCameraImage() const;     // [img_h * img_w], pixel (x, y) at y * img_w + x
    Pose getPose() const;                      // getAPS().getPose(name)
    int getStorageInterval() const { return 100; }

private:
    Arena* a_;
    int arena_, robot_;
};

// arena.Arena for a batch of arenas that share one experiment shape; arena i is seeded with seeds[i].
// step() is one think step as RunOffline drives it: Arena.step (perturbations, sense, think, move) followed by the World.step
// calls up to the next think step (step_interval ticks, fused on the device).
class Arena {
public:
    typedef std::vector<std::vector<std::shared_ptr<Controller>>> Controllers;   // [n_arenas][n_robots]

    Arena(const Experiment& e, const std::vector<int64_t>& seeds, const Calibration& calib, int device = 0, Controllers controllers = {})
        : cfg_(e.toConfig((int)seeds.size())), controllers_(std::move(controllers)), seeds_(seeds), code_(e.code()), imgW_(calib.width),
          imgH_(calib.height) {
        if (!controllers_.empty()) cfg_.controller = PS_CTRL_EXTERNAL;   // host controllers: the batched per-think-step callback path
        check(ps_create(&cfg_, device, &h_));
        try {
            check(ps_load_calibration(h_, calib.XrYr.data(), calib.xy.data(), calib.flags.data(), calib.n, calib.width, calib.height));
            check(ps_reset(h_, seeds.data()));
            check(ps_get_dims(h_, &nBodies_, &nProxies_, &maxContacts_, &occW_, &occH_));
        } catch (...) {
            ps_destroy(h_);
            h_ = nullptr;
            throw;
        }
    }
    ~Arena() { dispose(); }
    Arena(const Arena&) = delete;
    Arena& operator=(const Arena&) = delete;

    // ticks: World.step calls that follow the think step (default: the whole interval; RunOffline's last step runs one)
    void step(bool allowThinking = true, bool allowForwards = true, bool allowTurning = true, bool forcedMarch = false, int forcedTurn = 0,
              bool showCamera = false, int ticks = 0) {
        (void)showCamera;
        if (ticks < 0 || ticks > cfg_.step_interval) throw std::invalid_argument("ticks must be in [1, step_interval]");
        if (forcedMarch || forcedTurn || !(allowThinking && allowForwards && allowTurning))
            throw std::logic_error("only the RunOffline call Arena.step(true, true, true, false, 0, false) is on the hot path");
        obsValid_ = false;
        if (controllers_.empty()) {
            check(ps_step(h_, ticks > 0 ? ticks : cfg_.step_interval));
            ++stepCount_;
            return;
        }
        check(ps_sense(h_));
        fetchObservations(true);
        int stepCount = getStepCount();
        size_t R = (size_t)cfg_.n_robots;
        std::vector<float> fwd((size_t)cfg_.n_arenas * R), tq(fwd.size());
        for (int a = 0; a < cfg_.n_arenas; ++a)
            for (int r = 0; r < cfg_.n_robots; ++r) {
                Suite s(this, a, r);
                Controller& c = *controllers_[a][r];
                c.computeDesired(s, stepCount);
                fwd[a * R + r] = c.getForwards();
                tq[a * R + r] = c.getTorque();
            }
        obsValid_ = false;
        check(ps_step_external(h_, fwd.data(), tq.data()));
        ++stepCount_;
    }
    void coast(bool = true, bool = true, bool = true) {
        throw std::logic_error("RunOffline's coast ticks are folded into step(): call step() once per think interval");
    }
    int getStepCount() const { return stepCount_; }   // Arena.getStepCount (Arena.java:392-394), tracked on the host: the batch is in lock step
    bool hasHostControllers() const { return !controllers_.empty(); }
    ps_handle getWorld() const { return h_; }          // Arena.getWorld (Arena.java:396-398): the engine handle stands for the World
    struct Enclosure {                                 // Arena.getEnclosure (Arena.java:400-402): RoundedRectangleEnclosure.java:37-44
        float WIDTH, HEIGHT, R;
    };
    Enclosure getEnclosure() const { return {187.0f * cfg_.enclosure_scale, 187.0f * cfg_.enclosure_scale, 15.5f * cfg_.enclosure_scale}; }
    void dispose() {
        if (h_) ps_destroy(h_);
        h_ = nullptr;
    }
    Suite getSuite(int arena, int robot) { return Suite(this, arena, robot); }
    const ps_config& config() const { return cfg_; }
    ps_handle handle() const { return h_; }
    int nBodies() const { return nBodies_; }
    // [n_arenas][6][NB]: sweep.c.x, sweep.c.y, sweep.a, v.x, v.y, omega (pucks first, then robots, creation order)
    std::vector<float> bodies() {
        std::vector<float> b((size_t)cfg_.n_arenas * 6 * (size_t)nBodies_);
        ps_state_view v = {};
        v.body = b.data();
        check(ps_get_state(h_, &v));
        return b;
    }

    // [n_arenas][n_robots][3]: body.m_xf.position.x, .y, body.getAngle() NOW (what Arena.java:231-233 stores at the head of
    // Arena.step). The poses the engine keeps are those of the last sense, one think interval old, so a sense is run on the
    // current state and the state put back (sensing moves nothing but the LocalMap.carrying hysteresis).
    std::vector<float> robotPoses() {
        size_t A = (size_t)cfg_.n_arenas, NB = (size_t)nBodies_, NP = (size_t)nProxies_, MC = (size_t)maxContacts_, R = (size_t)cfg_.n_robots;
        std::vector<float> body(A * 6 * NB), saabb(A * 4 * NB), fat(A * 4 * NP), cimp(A * MC * 4);
        std::vector<uint32_t> ckey(A * MC), cinfo(A * MC);
        std::vector<ps_robot_state> rs(A * R);
        std::vector<ps_arena_state> as(A);
        ps_state_view v = {};
        v.body = body.data();
        v.sense_aabb = saabb.data();
        v.fat_aabb = fat.data();
        v.contact_key = ckey.data();
        v.contact_info = cinfo.data();
        v.contact_impulse = cimp.data();
        v.robot = rs.data();
        v.arena = as.data();
        check(ps_get_state(h_, &v));
        check(ps_sense(h_));
        obsValid_ = false;
        fetchObservations(false);
        std::vector<float> pose = pose_;
        check(ps_set_state(h_, &v));
        obsValid_ = false;
        return pose;
    }
    // The files the reference stores at a storage step (stepCount % 100 == 0), from the record the engine kept on the device for
    // the latest such step (ps_get_snapshot), per arena under <outputDir>/<code>/<seed>/:
    //   step%07d_robots.txt ("x, y, theta" as doubles), step%07d_<COLOUR>_pucks.txt ("x, y" as floats)      Arena.java:223-243
    //   step%07d_R<i>_cachePoints.txt (8 lines, "NaN, NaN" = null)                                         CacheConsController.java:531-551
    //   step%07d_R<i>_stateCounts.txt (one count per State)                           CacheConsController.java:313-333, BHDController.java:103-123
    // so that analysis.Analyze reads GPU runs unchanged. Returns the files written.
    std::vector<std::string> saveSnapshot(const std::string& outputDir) {
        static const char* kColours[PS_MAX_COLOURS] = {"RED", "GREEN", "MAGENTA", "CYAN", "ORANGE", "GRAY", "YELLOW", "PINK"};   // SensedType.java:9-23
        int R = cfg_.n_robots, P = nBodies_ - R, K = cfg_.n_puck_types;
        size_t A = (size_t)cfg_.n_arenas;
        std::vector<int32_t> step(A), counts(A * R * 8);
        std::vector<float> pose(A * R * 3), puck(A * (size_t)std::max(1, P) * 2), cache(A * R * PS_MAX_COLOURS * 2);
        ps_snapshot_view sv = {};
        sv.step_count = step.data();
        sv.robot_pose = pose.data();
        sv.puck_xy = puck.data();
        sv.cache_xy = cache.data();
        sv.state_counts = counts.data();
        check(ps_get_snapshot(h_, &sv));
        int perType = std::max(1, P / std::max(1, K));
        int nStates = cfg_.controller == PS_CTRL_CACHECONS ? 7 : (cfg_.controller == PS_CTRL_BHD ? 4 : 0);   // State.values().length
        std::vector<std::string> written;
        ::mkdir(outputDir.c_str(), 0777);
        ::mkdir((outputDir + "/" + code_).c_str(), 0777);
        for (size_t a = 0; a < A; ++a) {
            if (step[a] < 0) continue;
            char name[32];
            std::snprintf(name, sizeof(name), "step%07d", step[a]);
            std::string d = outputDir + "/" + code_ + "/" + std::to_string(seeds_[a]);
            ::mkdir(d.c_str(), 0777);
            std::string base = d + "/" + name;
            {
                std::ofstream f(base + "_robots.txt");
                for (int r = 0; r < R; ++r) {
                    const float* p = pose.data() + (a * R + r) * 3;
                    f << javaDoubleStr(p[0]) << ", " << javaDoubleStr(p[1]) << ", " << javaDoubleStr(p[2]) << "\n";
                }
                written.push_back(base + "_robots.txt");
            }
            for (int k = 0; k < K; ++k) {
                int lo = k * perType, hi = std::min(P, (k + 1) * perType);
                if (hi <= lo) continue;
                std::string fn = base + "_" + kColours[k] + "_pucks.txt";
                std::ofstream f(fn);
                for (int p = lo; p < hi; ++p) f << javaFloatStr(puck[(a * P + p) * 2]) << ", " << javaFloatStr(puck[(a * P + p) * 2 + 1]) << "\n";
                written.push_back(fn);
            }
            if (!controllers_.empty()) continue;   // a host controller writes its own dumps
            for (int r = 0; r < R; ++r) {
                std::string rb = base + "_R" + std::to_string(r);
                if (cfg_.controller == PS_CTRL_CACHECONS) {
                    std::ofstream f(rb + "_cachePoints.txt");
                    for (int k = 0; k < PS_MAX_COLOURS; ++k) {
                        const float* c = cache.data() + ((a * R + r) * PS_MAX_COLOURS + k) * 2;
                        f << javaFloatStr(c[0]) << ", " << javaFloatStr(c[1]) << "\n";
                    }
                    written.push_back(rb + "_cachePoints.txt");
                }
                if (nStates) {
                    std::ofstream f(rb + "_stateCounts.txt");
                    for (int k = 0; k < nStates; ++k) f << counts[(a * R + r) * 8 + k] << "\n";   // FileUtils.saveArray (FileUtils.java:40-54)
                    written.push_back(rb + "_stateCounts.txt");
                }
            }
        }
        return written;
    }
    // analysis.Analyze's per-repetition statistics (Analyze.java:176-218), accumulated on the device at every storage step
    std::vector<ps_arena_analysis> analysis() {
        std::vector<ps_arena_analysis> an((size_t)cfg_.n_arenas);
        check(ps_get_analysis(h_, an.data()));
        return an;
    }

private:
    friend class Suite;
    void fetchObservations(bool camera) {
        if (obsValid_ && (!camera || obsCamera_)) return;
        size_t n = (size_t)cfg_.n_arenas * (size_t)cfg_.n_robots;
        lm_.resize(n);
        occ_.resize(n * (size_t)(occW_ * occH_));
        pose_.resize(n * 3);
        if (camera) cam_.resize(n * (size_t)(imgW_ * imgH_));
        ps_obs_view v = {};
        v.localmap = lm_.data();
        v.occupancy = occ_.data();
        v.pose = pose_.data();
        v.camera = camera ? cam_.data() : nullptr;
        check(ps_get_observations(h_, &v));
        obsValid_ = true;
        obsCamera_ = camera;
    }
    ps_config cfg_;
    Controllers controllers_;
    std::vector<int64_t> seeds_;
    std::string code_;
    ps_handle h_ = nullptr;
    int32_t nBodies_ = 0, nProxies_ = 0, maxContacts_ = 0, occW_ = 0, occH_ = 0;
    int imgW_ = 0, imgH_ = 0;
    bool obsValid_ = false, obsCamera_ = false;
    int stepCount_ = 0;
    std::vector<ps_localmap> lm_;
    std::vector<uint8_t> occ_, cam_;
    std::vector<float> pose_;
};

inline const ps_localmap& Suite::getLocalMap() const {
    a_->fetchObservations(false);
    return a_->lm_[(size_t)arena_ * a_->cfg_.n_robots + robot_];
}
inline const uint8_t* Suite::getOccupancy() const {
    a_->fetchObservations(false);
    return a_->occ_.data() + ((size_t)arena_ * a_->cfg_.n_robots + robot_) * (size_t)(a_->occW_ * a_->occH_);
}
inline const uint8_t* Suite::getCameraImage() const {
    a_->fetchObservations(true);
    return a_->cam_.data() + ((size_t)arena_ * a_->cfg_.n_robots + robot_) * (size_t)(a_->imgW_ * a_->imgH_);
}
inline Pose Suite::getPose() const {
    a_->fetchObservations(false);
    const float* p = a_->pose_.data() + ((size_t)arena_ * a_->cfg_.n_robots + robot_) * 3;
    Pose o;
    o.x = p[0];
    o.y = p[1];
    o.theta = p[2];
    return o;
}

// arena.RunOffline's loop (RunOffline.java:61-62,72-95) over a batch: every call of step() is one physics tick of the
// reference's driver; think steps (and with them the interval's World.step calls, fused on the device) happen on every
// step_interval-th tick. The stop check comes first (:75-86); the reference runs ONE World.step after the last think step
// (stepCount == maxStepCount) before the check ends the episode, and so does this.
class RunOffline {
public:
    explicit RunOffline(Arena& arena, std::string outputDir = "", int maxStepCount = 10000)
        : arena_(arena), outputDir_(std::move(outputDir)), maxStepCount_(maxStepCount) {}
    void step() {
        if (arena_.getStepCount() > maxStepCount_) {
            active_ = false;
            return;
        }
        if (updateCount_ % arena_.config().step_interval == 0) {
            bool storage = arena_.getStepCount() % 100 == 0;   // Arena.STORAGE_INTERVAL
            bool last = arena_.getStepCount() == maxStepCount_ && !arena_.hasHostControllers();
            arena_.step(true, true, true, false, 0, false, last ? 1 : 0);
            if (storage && !outputDir_.empty()) arena_.saveSnapshot(outputDir_);
        }
        ++updateCount_;
    }
    void run() {
        while (active_) step();
    }
    bool isActive() const { return active_; }
    int getUpdateCount() const { return updateCount_; }

private:
    Arena& arena_;
    std::string outputDir_;
    int maxStepCount_;
    int updateCount_ = 0;
    bool active_ = true;
};

// experiment.ExperimentBlock (ExperimentBlock.java:15-28) and ExperimentManager.readExistingExperiments (ExperimentManager.java:397-428):
// every <code>.properties beside common.properties is one block of `repetitions` experiments, seeds startIndex .. repetitions - 1.
// One arena = one (code, seed); a handle carries one ps_config, so a block is one batch.
struct ExperimentBlock {
    std::string code;
    std::vector<int64_t> seeds;
};
inline std::vector<ExperimentBlock> readExistingExperiments(const std::string& outputDir, const std::vector<std::string>& propertyFiles) {
    auto common = loadProperties(outputDir + "/common.properties");
    auto rep = common.find("repetitions");
    if (rep == common.end()) throw std::invalid_argument("ExperimentManager: Problem loading common properties (no `repetitions`)");
    int repetitions = std::stoi(rep->second), startIndex = common.count("startIndex") ? std::stoi(common["startIndex"]) : 0;
    std::vector<std::string> names;
    for (const std::string& f : propertyFiles) {
        const std::string suffix = ".properties";
        if (f == "common.properties" || f.size() <= suffix.size() || f.compare(f.size() - suffix.size(), suffix.size(), suffix) != 0) continue;
        names.push_back(f.substr(0, f.size() - suffix.size()));
    }
    std::sort(names.begin(), names.end());
    std::vector<ExperimentBlock> blocks;
    for (const std::string& n : names) {
        ExperimentBlock b;
        b.code = n;
        for (int i = startIndex; i < repetitions; ++i) b.seeds.push_back(i);
        blocks.push_back(b);
    }
    return blocks;
}

}  // namespace puckswarm
