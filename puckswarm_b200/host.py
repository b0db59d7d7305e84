"""Host-side mirror of the reference's interfaces for the per-tick path (Python, because no JDK exists in this image;
the JNI shim a maintainer would compile is in java/ and INTEGRATION.md).

Names, argument meaning and error behaviour follow the reference:
  experiment.Experiment / ExperimentManager  src/experiment/Experiment.java:21-90, ExperimentManager.java:397-420
  arena.Arena                                src/arena/Arena.java:45,183-267,310-316,392-411
  arena.RunOffline.step                      src/arena/RunOffline.java:72-95
  sensors.Suite / APS                        src/sensors/Suite.java:5-43, src/sensors/SimAPS.java:14-18
  controllers.Controller                     src/controllers/Controller.java:9-31
  snapshot files                             src/arena/Arena.java:223-243, PoseList.java:32-46, PositionList.java:36-49

One `Arena` object here stands for a BATCH of independent arenas (one per (experiment code, seed)); every accessor takes
the arena index. All simulation work happens in the CUDA engine behind include/puckswarm_b200.h; nothing here
computes physics, sensing or control.
"""
import math
import os
import re

import numpy as np

from . import abi

COLOUR_NAMES = ["RED", "GREEN", "MAGENTA", "CYAN", "ORANGE", "GRAY", "YELLOW", "PINK"]   # SensedType.java:9-23

# property key -> (ps_config field, type, enum map)   
_SEL = {"PROB_ABSOLUTE": abi.PS_SEL_PROB_ABSOLUTE, "PROB_RELATIVE": abi.PS_SEL_PROB_RELATIVE, "RELATIVE": abi.PS_SEL_RELATIVE,
        "RELATIVE_FORGET": abi.PS_SEL_RELATIVE_FORGET}
_SEP = {"IGNORE": abi.PS_SEP_IGNORE, "NULLIFY_OLD": abi.PS_SEP_NULLIFY_OLD, "ACCEPT_ISOLATED": abi.PS_SEP_ACCEPT_ISOLATED,
        "ACCEPT_LARGER": abi.PS_SEP_ACCEPT_LARGER}
_CTRL = {"CacheCons": abi.PS_CTRL_CACHECONS, "BHD": abi.PS_CTRL_BHD, "ProbSeek": abi.PS_CTRL_PROBSEEK}
PROPERTY_MAP = {
    "controllerType": ("controller", str, _CTRL),
    "Arena.nPucks": ("n_pucks", int, None), "Arena.nPuckTypes": ("n_puck_types", int, None),
    "Arena.nRobots": ("n_robots", int, None), "Arena.nInformedRobots": ("n_informed_robots", int, None),
    "Arena.puckShiftStep": ("puck_shift_step", int, None), "Arena.puckShuffleStep": ("puck_shuffle_step", int, None),
    "RoundedRectangleEnclosure.scale": ("enclosure_scale", float, None),
    "LocalMap.CARRY_THRESHOLD_LO": ("carry_threshold_lo", float, None), "LocalMap.CARRY_THRESHOLD_HI": ("carry_threshold_hi", float, None),
    "VFHPlus.SAFETY_DISTANCE_GAMMA": ("vfh_safety_gamma", float, None), "VFHPlus.SAFETY_DISTANCE_MASK": ("vfh_safety_mask", float, None),
    "VFHPlus.TAU_LO": ("vfh_tau_lo", float, None), "VFHPlus.TAU_HI": ("vfh_tau_hi", float, None),
    "CacheConsController.K1": ("cc_k1", float, None), "CacheConsController.K2": ("cc_k2", float, None),
    "CacheConsController.PICK_UP_TIME": ("cc_pick_up_time", int, None), "CacheConsController.PUSH_IN_TIME": ("cc_push_in_time", int, None),
    "CacheConsController.BACKUP_TIME": ("cc_backup_time", int, None),
    "CacheConsController.PREDICTION_DISTANCE_SQD": ("cc_prediction_distance_sqd", float, None),
    "CacheConsController.ST_DEV": ("cc_st_dev", float, None), "CacheConsController.EXILE_TIME": ("cc_exile_time", int, None),
    "CacheConsController.SPIT_OUT_TIME": ("cc_spit_out_time", int, None),
    "CacheConsController.CACHE_SELECTION_OPTION": ("cc_cache_selection", str, _SEL),
    "CacheConsController.K_ABSOLUTE": ("cc_k_absolute", float, None), "CacheConsController.K_RELATIVE": ("cc_k_relative", float, None),
    "CacheConsController.FORGET_PROB": ("cc_forget_prob", float, None),
    "CacheConsController.CACHE_SEPARATION_OPTION": ("cc_cache_separation", str, _SEP),
    "CacheConsController.HOME_SEPARATION_DISTANCE": ("cc_home_separation_distance", float, None),
    "CacheConsController.IGNORE_PUCKS": ("cc_ignore_pucks", bool, None),
    "BHDController.PUSH_IN_TIME": ("bhd_push_in_time", int, None), "BHDController.BACKUP_TIME": ("bhd_backup_time", int, None),
    "ProbSeekController.PLACEMENT_TIME": ("ps_placement_time", int, None),
}
# ProbSeekController's other keys have the CacheCons names and defaults (ProbSeekController.java:98-121) and use the same
# ps_config fields; they apply when controllerType = ProbSeek
PROBSEEK_SHARED = {
    "ProbSeekController.K1": ("cc_k1", float), "ProbSeekController.K2": ("cc_k2", float),
    "ProbSeekController.PICK_UP_TIME": ("cc_pick_up_time", int), "ProbSeekController.PUSH_IN_TIME": ("cc_push_in_time", int),
    "ProbSeekController.BACKUP_TIME": ("cc_backup_time", int),
    "ProbSeekController.PREDICTION_DISTANCE_SQD": ("cc_prediction_distance_sqd", float), "ProbSeekController.ST_DEV": ("cc_st_dev", float),
}


def load_properties(path):
    """java.util.Properties.load for the subset the reference writes: `key=value` / `key: value` lines, # and ! comments."""
    out = {}
    with open(path) as f:
        for line in f:
            s = line.strip()
            if not s or s[0] in "#!":
                continue
            m = re.match(r"([^=:\s]+)\s*[=:\s]\s*(.*)$", s)
            if m:
                out[m.group(1)] = m.group(2).strip()
    return out


class Experiment:
    """One (code, seed): properties with typed defaults, as experiment.Experiment.getProperty (Experiment.java:76-90)."""

    def __init__(self, seed=0, properties=None, code="default"):
        self.seed = int(seed)                       # ExperimentBlock.java:27-28: the seed is the repetition index
        self.properties = dict(properties or {})
        self.code = self.properties.get("code", code)

    @classmethod
    def from_files(cls, output_dir, code, seed):
        """common.properties then <code>.properties, the latter overriding (Experiment.java:26-33)."""
        props = load_properties(os.path.join(output_dir, "common.properties"))
        props.update(load_properties(os.path.join(output_dir, code + ".properties")))
        return cls(seed, props, code)

    def getProperty(self, key, default):
        """Missing key -> default; malformed number -> ValueError (Java: NumberFormatException)."""
        if key not in self.properties:
            return default
        v = self.properties[key]
        if isinstance(default, bool):
            return str(v).lower() == "true"           # Boolean.parseBoolean
        if isinstance(default, int):
            return int(v)
        if isinstance(default, float):
            return float(v)
        return v

    def getIndex(self):
        return self.seed

    def to_config(self, n_arenas=1, **overrides):
        """ps_config carrying every property the hot path reads (unknown keys are ignored like in the reference)."""
        cfg = abi.config_defaults(n_arenas=n_arenas)
        for key, (field, typ, enum) in PROPERTY_MAP.items():
            if key not in self.properties:
                continue
            raw = self.properties[key]
            if enum is not None:
                if raw not in enum:
                    raise ValueError("%s: unknown value %r" % (key, raw))
                val = enum[raw]
            elif typ is bool:
                val = 1 if str(raw).lower() == "true" else 0
            else:
                val = typ(raw)
            setattr(cfg, field, val)
        if cfg.controller == abi.PS_CTRL_PROBSEEK:
            d = abi.config_defaults()
            for key, (field, typ) in PROBSEEK_SHARED.items():
                setattr(cfg, field, typ(self.properties[key]) if key in self.properties else getattr(d, field))
        for k, v in overrides.items():
            setattr(cfg, k, v)
        return cfg


class ExperimentBlock:
    """experiment.ExperimentBlock (ExperimentBlock.java:15-28): one experiment code repeated for seeds startIndex .. repetitions-1;
    Experiment i of the block has index (= seed) i."""

    def __init__(self, size, start_index, output_dir, code):
        self.code, self.output_dir = code, output_dir
        self.seeds = list(range(int(start_index), int(size)))

    def experiments(self):
        return [Experiment.from_files(self.output_dir, self.code, seed) for seed in self.seeds]


class ExperimentManager:
    """experiment.ExperimentManager's view of an existing sweep directory (ExperimentManager.java:397-428): every <code>.properties
    beside common.properties is one experiment block; `repetitions` and the optional `startIndex` come from common.properties."""

    @staticmethod
    def readExistingExperiments(output_dir):
        names = sorted(f for f in os.listdir(output_dir) if f.endswith(".properties") and f != "common.properties")
        common = load_properties(os.path.join(output_dir, "common.properties"))
        if "repetitions" not in common:
            raise ValueError("ExperimentManager: Problem loading common properties (no `repetitions`)")
        repetitions = int(common["repetitions"])                 # Integer.valueOf: NumberFormatException -> ValueError
        start_index = int(common.get("startIndex", "0"))
        return [ExperimentBlock(repetitions, start_index, output_dir, n[:-len(".properties")]) for n in names]


class Sweep:
    """A whole sweep directory on the GPU: one arena per (code, seed). A handle carries one ps_config, so every experiment block
    (one code, its repetitions) is one batch; up to `concurrent` batches are stepped side by side (their kernels overlap on the
    device). Snapshot files go to <output_dir>/<code>/<seed>/ exactly where RunOffline puts them, so analysis.Analyze reads them."""

    def __init__(self, output_dir, device=0, concurrent=4, write_files=True):
        self.output_dir, self.device, self.concurrent, self.write_files = output_dir, device, max(1, int(concurrent)), write_files
        self.blocks = ExperimentManager.readExistingExperiments(output_dir)

    def pairs(self):
        """[(code, seed)] in the reference's execution order (block after block, repetition after repetition)."""
        return [(b.code, seed) for b in self.blocks for seed in b.seeds]

    def run(self, max_step_count=None):
        """Runs every block to Arena.maxStepCount (property, default 10000; RunOffline.java:75-76). Returns {code: summary} with
        Analyze's avgTWC / avgSTC / incompletions (Analyze.java:240-258) and the per-repetition arrays."""
        from . import dist as psd
        results = {}
        pending = [b for b in self.blocks if b.seeds]
        while pending:
            wave, pending = pending[:self.concurrent], pending[self.concurrent:]
            runs = []
            for b in wave:
                exp = Experiment.from_files(self.output_dir, b.code, b.seeds[0])
                arena = Arena(exp, b.seeds, device=self.device)
                limit = exp.getProperty("Arena.maxStepCount", Arena.maxStepCount) if max_step_count is None else max_step_count
                runs.append((b, arena, RunOffline(arena, self.output_dir if self.write_files else None, limit)))
            while any(r.active for _, _, r in runs):
                for _, _, r in runs:                      # round-robin: ps_step is asynchronous, the handles overlap on the GPU
                    if r.active:
                        r.step()
            for b, arena, _ in runs:
                stats = arena.engine.get_stats(arena.engine.analysis_threshold)
                summary = psd.ensemble_summary(stats)
                summary.update(arena.analysis())
                summary["seeds"] = list(b.seeds)
                results[b.code] = summary
                arena.dispose()
        return results


def java_float_str(x):
    """Float.toString(float): shortest decimal that round-trips, with Java's 1.0E-4 / 1.0E7 switch to E notation."""
    x = np.float32(x)
    if np.isnan(x):
        return "NaN"
    if np.isinf(x):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if np.signbit(x) else "0.0"
    a = abs(float(x))
    if 1e-3 <= a < 1e7:
        s = np.format_float_positional(x, unique=True, trim="0")
        return s if "." in s else s + ".0"
    s = np.format_float_scientific(x, unique=True, trim="0", exp_digits=1)
    mant, exp = s.split("e")
    if "." not in mant:
        mant += ".0"
    return mant + "E" + str(int(exp))


def java_double_str(x):
    """Double.toString(double) for the values PoseList writes (floats widened to double)."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if x == 0:
        return "-0.0" if math.copysign(1, x) < 0 else "0.0"
    a = abs(x)
    if 1e-3 <= a < 1e7:
        s = repr(x)
        if "e" in s or "E" in s:
            s = np.format_float_positional(x, unique=True, trim="0")
        return s if "." in s else s + ".0"
    s = np.format_float_scientific(x, unique=True, trim="0", exp_digits=1)
    mant, exp = s.split("e")
    if "." not in mant:
        mant += ".0"
    return mant + "E" + str(int(exp))


class Pose:
    """sensors.Pose (src/sensors/Pose.java:31-55): double x, y, theta."""

    def __init__(self, x, y, theta):
        self.x, self.y, self.theta = float(x), float(y), float(theta)

    def getX(self):
        return self.x

    def getY(self):
        return self.y

    def getTheta(self):
        return self.theta


class Controller:
    """controllers.Controller (Controller.java:9-31) for host-side controllers driven through ps_step_external.
    getForwards / getTorque are only valid after computeDesired, as in the reference."""

    def computeDesired(self, suite, stepCount):
        raise NotImplementedError

    def getForwards(self):
        raise NotImplementedError

    def getTorque(self):
        raise NotImplementedError

    def getInfoString(self):
        return ""


class Suite:
    """sensors.Suite of one robot of one arena: a view on the observations of the latest sense."""

    def __init__(self, arena, a, r):
        self._arena, self._a, self._r = arena, a, r

    def getRobotName(self):
        return "R%d" % self._r

    def getCameraImage(self):
        """160 x 120 SensedType ordinals, [y][x] (STCameraImage.pixels[x][y])."""
        return self._arena._obs(camera=True).camera[self._a, self._r]

    def getLocalMap(self):
        return self._arena._obs().localmap[self._a, self._r]

    def getOccupancy(self):
        return self._arena._obs().occupancy[self._a, self._r]

    def getAPS(self):
        return self

    def getPose(self, name=None):
        p = self._arena._obs().pose[self._a, self._r]
        return Pose(p[0], p[1], p[2])

    def getStorageInterval(self):
        return Arena.STORAGE_INTERVAL


class Arena:
    """arena.Arena for a batch of arenas that share one experiment shape; arena i is seeded with seeds[i].

    step() is one think step as RunOffline drives it: Arena.step (sense, think, move) followed by the World.step calls
    up to the next think step (stepInterval ticks); coast() advances one tick re-applying the last commands.
    """
    STORAGE_INTERVAL = 100      # Arena.java:39
    maxStepCount = 10000        # RunOffline.java:76

    def __init__(self, experiment, seeds, device=0, controllers=None, **cfg_overrides):
        from . import engine
        self.experiment = experiment
        self.seeds = np.ascontiguousarray(seeds, np.int64)
        self.cfg = experiment.to_config(n_arenas=len(self.seeds), **cfg_overrides)
        self.controllers = controllers      # [n_arenas][n_robots] host Controller objects, or None for the on-device ones
        if controllers is not None:
            self.cfg.controller = abi.PS_CTRL_EXTERNAL
        self.engine = engine.Engine(self.cfg, device=device)
        self.engine.reset(self.seeds)
        self._obs_cache = None
        self._obs_has_camera = False
        self._ticks_in_interval = 0
        self._step_count = 0

    # ---- Arena.step / coast ----
    def step(self, allowThinking=True, allowForwards=True, allowTurning=True, forcedMarch=False, forcedTurn=0, showCamera=False, ticks=None):
        """One think step and the World.step calls that follow it (`ticks`, default stepInterval: RunOffline's coast ticks are
        folded in). The periodic record of a storage step (stepCount % 100 == 0) is kept on the device: save_snapshot()."""
        if forcedMarch or forcedTurn or not (allowThinking and allowForwards and allowTurning):
            raise NotImplementedError("only the RunOffline call Arena.step(true, true, true, false, 0, false) is on the hot path")
        self._obs_cache = None
        interval = self.cfg.step_interval if ticks is None else int(ticks)
        if not (1 <= interval <= self.cfg.step_interval):
            raise ValueError("ticks must be in [1, stepInterval]")
        if self.controllers is None:
            self.engine.step(interval)
        else:
            if interval != self.cfg.step_interval:
                raise NotImplementedError("host controllers advance whole think intervals (ps_step_external)")
            self.engine.sense()
            fwd = np.zeros((self.cfg.n_arenas, self.cfg.n_robots), np.float32)
            tq = np.zeros_like(fwd)
            step_count = self.getStepCount()
            for a in range(self.cfg.n_arenas):
                for r in range(self.cfg.n_robots):
                    c = self.controllers[a][r]
                    c.computeDesired(Suite(self, a, r), step_count)
                    fwd[a, r], tq[a, r] = c.getForwards(), c.getTorque()
            self._obs_cache = None
            self.engine.step_external(fwd, tq)
        self._step_count += 1

    def coast(self, allowThinking=True, allowForwards=True, allowTurning=True):
        """One tick with the previous commands; only valid between think steps of an on-device controller."""
        raise NotImplementedError("RunOffline's coast ticks are folded into step(): call step() once per think interval")

    def getStepCount(self):
        """Arena.stepCount (think steps done), tracked on the host: every arena of the batch is at the same count."""
        return self._step_count

    def getWorld(self):
        return self.engine

    def getEnclosure(self):
        s = float(self.cfg.enclosure_scale)
        return {"WIDTH": 187.0 * s, "HEIGHT": 187.0 * s, "R": 15.5 * s}

    def dispose(self):
        self.engine.close()

    # ---- observations ----
    def _obs(self, camera=False):
        """Observations of the latest sense. The on-device path keeps the local map, occupancy grid and poses of the last
        think step; camera images are only produced by an explicit sense, which is done here without disturbing the
        carry hysteresis (LocalMap.carrying) the next think step depends on."""
        if self._obs_cache is None or (camera and not self._obs_has_camera):
            if camera and self.controllers is None:
                st = self.engine.get_state()
                self.engine.sense()
                self._obs_cache = self.engine.get_observations(camera=True)
                self.engine.set_state(st)
            else:
                self._obs_cache = self.engine.get_observations(camera=camera)
            self._obs_has_camera = camera
        return self._obs_cache

    def getSuite(self, arena, robot):
        return Suite(self, arena, robot)

    def robot_poses(self):
        """[n_arenas][n_robots][3]: body.m_xf.position.x, .y, body.getAngle() NOW (what Arena.java:231-233 stores at the head
        of Arena.step). The poses the engine keeps are those of the last sense, one think interval old, so a sense is run on
        the current state and the state put back (sensing moves nothing but the LocalMap.carrying hysteresis)."""
        st = self.engine.get_state()
        self.engine.sense()
        pose = self.engine.get_observations(camera=False).pose.copy()
        self.engine.set_state(st)
        self._obs_cache = None
        return pose

    def puck_positions(self):
        """[n_arenas][n_pucks][2] (pucks are circles about their origin: m_xf.position == sweep.c)."""
        st = self.engine.get_state()
        P = self.engine.NB - self.cfg.n_robots
        return np.stack([st.body[:, 0, :P], st.body[:, 1, :P]], axis=-1)

    # ---- the reference's snapshot files ----
    def save_snapshot(self, output_dir):
        """Write the files the reference stores at a storage step (stepCount % STORAGE_INTERVAL == 0), from the record the engine
        kept on the device for the latest such step (ps_get_snapshot), under <output_dir>/<code>/<seed>/:
          stepNNNNNNN_robots.txt, stepNNNNNNN_<COLOUR>_pucks.txt   Arena.java:223-243 (poses after the perturbations, before World.step)
          stepNNNNNNN_R<i>_cachePoints.txt                         CacheConsController.java:531-551 (8 lines, `NaN, NaN` = null)
          stepNNNNNNN_R<i>_stateCounts.txt                         CacheConsController.java:313-333, BHDController.java:103-123
        in the reference's text formats, so that the unmodified analysis.Analyze can read a GPU run. Returns the file names."""
        snap = self.engine.get_snapshot()
        per_type = max(1, snap.puck_xy.shape[1] // max(1, self.cfg.n_puck_types))
        n_states = {abi.PS_CTRL_CACHECONS: len(abi.CC_STATES), abi.PS_CTRL_BHD: len(abi.BHD_STATES)}.get(self.cfg.controller, 0)
        written = []
        for a, seed in enumerate(self.seeds):
            step = int(snap.step_count[a])
            if step < 0:
                continue
            d = os.path.join(output_dir, self.experiment.code, str(int(seed)))
            os.makedirs(d, exist_ok=True)
            base = os.path.join(d, "step%07d" % step)
            with open(base + "_robots.txt", "w") as f:
                for x, y, t in snap.robot_pose[a]:
                    f.write("%s, %s, %s\n" % (java_double_str(np.float32(x)), java_double_str(np.float32(y)), java_double_str(np.float32(t))))
            written.append(base + "_robots.txt")
            for k in range(self.cfg.n_puck_types):
                sel = snap.puck_xy[a, k * per_type:(k + 1) * per_type]
                if len(sel) == 0:
                    continue
                name = base + "_" + COLOUR_NAMES[k] + "_pucks.txt"
                with open(name, "w") as f:
                    for x, y in sel:
                        f.write("%s, %s\n" % (java_float_str(x), java_float_str(y)))
                written.append(name)
            if self.controllers is not None:
                continue      # a host controller writes its own dumps
            for r in range(self.cfg.n_robots):
                if self.cfg.controller == abi.PS_CTRL_CACHECONS:
                    name = base + "_R%d_cachePoints.txt" % r
                    with open(name, "w") as f:
                        for x, y in snap.cache_xy[a, r]:
                            f.write("%s, %s\n" % (java_float_str(x), java_float_str(y)))
                    written.append(name)
                if n_states:
                    name = base + "_R%d_stateCounts.txt" % r
                    with open(name, "w") as f:
                        for v in snap.state_counts[a, r, :n_states]:
                            f.write("%d\n" % int(v))       # FileUtils.saveArray (FileUtils.java:40-54)
                    written.append(name)
        return written

    def analysis(self):
        """analysis.Analyze's per-repetition statistics, accumulated on the device at every storage step (ps_get_analysis):
        per arena n_snapshots, steps_to_completion (-1 = target not reached), time-weighted completion, latest completion."""
        an = self.engine.get_analysis()
        twc = np.where(an["twc_den"] > 0, an["twc_num"] / np.maximum(an["twc_den"], 1e-300), 0.0)
        return {"n_snapshots": an["n_snapshots"].copy(), "stepsToCompletion": an["steps_to_completion"].copy(),
                "timeWeightedCompletion": twc, "percentCompletion": an["completion"].copy()}


class RunOffline:
    """arena.RunOffline's loop (RunOffline.java:61-62,72-95) over a batch: every call of step() is one physics tick of the
    reference's driver; think steps happen on every stepInterval-th tick and carry the interval's World.step calls with them."""

    def __init__(self, arena, output_dir=None, max_step_count=None):
        self.arena = arena
        self.output_dir = output_dir
        self.updateCount = 0
        self.maxStepCount = Arena.maxStepCount if max_step_count is None else int(max_step_count)   # Arena.maxStepCount property
        self.active = True

    def step(self):
        """RunOffline.step(): stop check first (:75-86), then Arena.step on think ticks. The reference runs ONE World.step after the
        last think step (stepCount == maxStepCount) before the check ends the episode; so does this."""
        if self.arena.getStepCount() > self.maxStepCount:
            self.active = False
            return
        interval = self.arena.cfg.step_interval
        if self.updateCount % interval == 0:
            storage = self.arena.getStepCount() % Arena.STORAGE_INTERVAL == 0
            last = self.arena.getStepCount() == self.maxStepCount
            self.arena.step(ticks=1 if (last and self.arena.controllers is None) else None)
            if storage and self.output_dir:
                self.arena.save_snapshot(self.output_dir)
        self.updateCount += 1

    def run(self, max_step_count=None):
        if max_step_count is not None:
            self.maxStepCount = int(max_step_count)
        while self.active:
            self.step()
