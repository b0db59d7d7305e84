"""Host-side mirror of the reference interfaces (puckswarm_b200/host.py): the parts that need no GPU."""
import numpy as np
import pytest

from puckswarm_b200 import abi, host


def test_java_number_formatting():
    # Float.toString / Double.toString as PositionList.save / PoseList.save write them
    assert host.java_float_str(np.float32(1.5)) == "1.5"
    assert host.java_float_str(np.float32(-38.847805)) == "-38.847805"
    assert host.java_float_str(np.float32(100.0)) == "100.0"
    assert host.java_float_str(np.float32(1e-4)) == "1.0E-4"
    assert host.java_float_str(np.float32(1.25e7)) == "1.25E7"
    assert host.java_float_str(np.float32(0.0)) == "0.0"
    assert host.java_double_str(np.float32(1.1)) == "1.100000023841858"
    assert host.java_double_str(2.0) == "2.0"
    assert host.java_double_str(np.float32(-0.5)) == "-0.5"


def test_properties_to_config(tmp_path):
    (tmp_path / "common.properties").write_text("#comment\nrepetitions=20\nArena.nRobots=4\nArena.nPucks=40\nArena.nPuckTypes=2\n")
    (tmp_path / "X.properties").write_text("code=X\ncontrollerType=BHD\nCacheConsController.CACHE_SEPARATION_OPTION=NULLIFY_OLD\n"
                                            "VFHPlus.TAU_LO=0.6\nCacheConsController.IGNORE_PUCKS=true\nArena.nRobots=6\n")
    e = host.Experiment.from_files(str(tmp_path), "X", seed=3)
    assert e.getIndex() == 3 and e.code == "X"
    assert e.getProperty("Arena.nRobots", 4) == 6 and e.getProperty("Arena.maxStepCount", 10000) == 10000
    assert e.getProperty("CacheConsController.IGNORE_PUCKS", False) is True
    cfg = e.to_config(n_arenas=5)
    assert (cfg.n_arenas, cfg.n_robots, cfg.n_pucks, cfg.n_puck_types) == (5, 6, 40, 2)
    assert cfg.controller == abi.PS_CTRL_BHD and cfg.cc_cache_separation == abi.PS_SEP_NULLIFY_OLD and cfg.cc_ignore_pucks == 1
    assert abs(cfg.vfh_tau_lo - 0.6) < 1e-7 and abs(cfg.vfh_tau_hi - 0.85) < 1e-7
    with pytest.raises(ValueError):
        host.Experiment(0, {"Arena.nRobots": "four"}).getProperty("Arena.nRobots", 4)
    with pytest.raises(ValueError):
        host.Experiment(0, {"controllerType": "Nope"}).to_config()
    # perturbation periods (Arena.java:187-188): absent = -1 = never
    assert (cfg.puck_shift_step, cfg.puck_shuffle_step, cfg.n_informed_robots) == (-1, -1, 0)
    c2 = host.Experiment(0, {"Arena.puckShiftStep": "5000", "Arena.puckShuffleStep": "2500", "Arena.nInformedRobots": "2"}).to_config()
    assert (c2.puck_shift_step, c2.puck_shuffle_step, c2.n_informed_robots) == (5000, 2500, 2)


def _write_sweep(d):
    (d / "common.properties").write_text("repetitions=3\nstartIndex=1\nArena.nRobots=3\nArena.nPucks=20\nArena.nPuckTypes=2\nArena.maxStepCount=100\n")
    (d / "K1_0.5.properties").write_text("code=K1_0.5\nCacheConsController.K1=0.5\n")
    (d / "BHD.properties").write_text("code=BHD\ncontrollerType=BHD\n")


def test_sweep_directory_enumeration(tmp_path):
    """ExperimentManager.readExistingExperiments (ExperimentManager.java:397-428) + ExperimentBlock (ExperimentBlock.java:15-28):
    one block per <code>.properties, repetitions / startIndex from common.properties, Experiment i has seed i."""
    _write_sweep(tmp_path)
    blocks = host.ExperimentManager.readExistingExperiments(str(tmp_path))
    assert [(b.code, b.seeds) for b in blocks] == [("BHD", [1, 2]), ("K1_0.5", [1, 2])]
    exps = blocks[1].experiments()
    assert [e.getIndex() for e in exps] == [1, 2] and exps[0].getProperty("CacheConsController.K1", 1.0) == 0.5
    assert exps[0].getProperty("Arena.maxStepCount", 10000) == 100
    assert exps[0].to_config(n_arenas=2).n_robots == 3 and blocks[0].experiments()[0].to_config().controller == abi.PS_CTRL_BHD
    (tmp_path / "common.properties").write_text("Arena.nRobots=3\n")
    with pytest.raises(ValueError):
        host.ExperimentManager.readExistingExperiments(str(tmp_path))


@pytest.mark.gpu
def test_sweep_runs_every_code_and_seed(tmp_path):
    """A sweep directory end to end: every (code, seed) is one arena, the reference's files appear under <code>/<seed>/, and
    Analyze's per-repetition numbers equal the oracle's restatement run on the same (code, seed) pairs."""
    import oracle_binding as oracle
    _write_sweep(tmp_path)
    sweep = host.Sweep(str(tmp_path), concurrent=2)
    assert sweep.pairs() == [("BHD", 1), ("BHD", 2), ("K1_0.5", 1), ("K1_0.5", 2)]
    res = sweep.run()
    for b in sweep.blocks:
        for seed in b.seeds:
            names = sorted(p.name for p in (tmp_path / b.code / str(seed)).iterdir())
            assert "step0000000_robots.txt" in names and "step0000100_RED_pucks.txt" in names
            assert ("step0000100_R2_cachePoints.txt" in names) == (b.code != "BHD")
            n_state_lines = len((tmp_path / b.code / str(seed) / "step0000100_R0_stateCounts.txt").read_text().split())
            assert n_state_lines == (4 if b.code == "BHD" else 7)
        cfg = b.experiments()[0].to_config(n_arenas=len(b.seeds))
        o = oracle.Oracle(cfg)
        o.reset(b.seeds)
        o.step(3 * 100 + 1)      # RunOffline: think steps 0..100, one World.step after the last
        an = o.get_analysis()
        assert np.array_equal(res[b.code]["stepsToCompletion"], an["steps_to_completion"])
        assert np.allclose(res[b.code]["timeWeightedCompletion"], an["twc_num"] / an["twc_den"], rtol=1e-12)
        assert (res[b.code]["n_snapshots"] == 2).all() and res[b.code]["repetitions"] == 2


@pytest.mark.gpu
def test_arena_runoffline_snapshot_and_host_controllers(tmp_path):
    """The mirror drives the engine like RunOffline drives Arena; snapshots come out in the reference's text format;
    a Java-style host controller runs through the batched external path."""
    import oracle_binding as oracle
    import parity
    exp = host.Experiment(0, {"Arena.nRobots": "3", "Arena.nPucks": "20", "code": "T"})
    seeds = [0, 1]
    arena = host.Arena(exp, seeds)
    run = host.RunOffline(arena, str(tmp_path))
    for _ in range(30):
        run.step()
    assert arena.getStepCount() == 10
    files = sorted(p.name for p in (tmp_path / "T" / "0").iterdir())
    per_robot = ["step0000000_R%d_%s.txt" % (r, what) for r in range(3) for what in ("cachePoints", "stateCounts")]
    assert files == sorted(["step0000000_GREEN_pucks.txt", "step0000000_RED_pucks.txt", "step0000000_robots.txt"] + per_robot)
    # CacheConsController.java:531-551: 8 lines, "NaN, NaN" for a null cache point; :313-333: one count per State, the step's own state counted
    cp = (tmp_path / "T" / "0" / "step0000000_R1_cachePoints.txt").read_text().strip().split("\n")
    assert len(cp) == 8 and cp[2:] == ["NaN, NaN"] * 6
    sc = [int(x) for x in (tmp_path / "T" / "0" / "step0000000_R1_stateCounts.txt").read_text().split()]
    assert len(sc) == len(abi.CC_STATES) and sum(sc) == 1
    rows = (tmp_path / "T" / "1" / "step0000000_robots.txt").read_text().strip().split("\n")
    assert len(rows) == 3 and all(len(r.split(", ")) == 3 for r in rows)
    # the poses of the snapshot are the spawn poses (Arena.java:231-233 stores them at the head of the first Arena.step)
    o0 = oracle.Oracle(arena.cfg)
    o0.reset(seeds)
    o0.sense()
    want = o0.get_observations().pose[1]
    got = np.array([[float(x) for x in r.split(", ")] for r in rows], np.float32)
    assert np.array_equal(got, want.astype(np.float32)), (got, want)
    # same trajectory as the oracle stepping the same experiment
    o = oracle.Oracle(arena.cfg)
    o.reset(seeds)
    o.step(30)
    assert parity.compare_states(arena.engine.get_state(), o.get_state(), 2, tol=1e-5) == []
    suite = arena.getSuite(1, 2)
    assert suite.getCameraImage().shape == (120, 160) and suite.getRobotName() == "R2"
    assert parity.compare_states(arena.engine.get_state(), o.get_state(), 2, tol=1e-5) == []   # inspection has no side effect
    arena.dispose()

    class Spin(host.Controller):          # TestMovementController-style scripted controller
        def computeDesired(self, suite, stepCount):
            self.pose = suite.getAPS().getPose()
            self.n = int(suite.getLocalMap()["n_pucks"])

        def getForwards(self):
            return 300000.0

        def getTorque(self):
            return 1000000.0

    ctrls = [[Spin() for _ in range(3)] for _ in seeds]
    a2 = host.Arena(exp, seeds, controllers=ctrls)
    p0 = a2.robot_poses().copy()
    for _ in range(5):
        a2.step()
    assert a2.getStepCount() == 5
    assert (np.abs(a2.robot_poses()[..., 2] - p0[..., 2]) > 0.05).all()   # every robot turned
    assert ctrls[1][2].n >= 0
    a2.dispose()
