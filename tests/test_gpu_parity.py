"""GPU parity tests: the CUDA engine (through the C-ABI) against the CPU oracle on identical seeded inputs.

Bar (BASELINE.json north_star): contact pair lists and discrete sensor outputs bit-exact, poses and velocities
within 1e-5 relative. The physics path and the sensing path use IEEE float arithmetic in JBox2D's operation
order on both sides, so they are asserted bit-exact here; only the controllers' double-precision libm calls
(atan2 / sin / cos / log on the device vs glibc) may differ in the last ulp, hence TOL for full steps.
"""
import numpy as np
import pytest

from puckswarm_b200 import abi

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _engine(cfg, **kw):
    from puckswarm_b200 import engine
    return engine.Engine(cfg, device=0, **kw)


def _oracle(cfg, threads=None):
    import os
    import oracle_binding
    return oracle_binding.Oracle(cfg, threads=threads or min(64, max(8, os.cpu_count() or 8)))


@pytest.mark.parametrize("shape", [(1, 2, 10, 2), (3, 4, 40, 2), (2, 16, 200, 2), (2, 5, 21, 3)])
def test_reset_matches_oracle(shape):
    import parity
    A, R, P, K = shape
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, n_puck_types=K)
    e, o = _engine(cfg), _oracle(cfg)
    seeds = np.arange(100, 100 + A)
    e.reset(seeds)
    o.reset(seeds)
    assert parity.compare_states(e.get_state(), o.get_state(), A) == []


@pytest.mark.parametrize("shape,steps", [((2, 4, 40), 200), ((2, 16, 200), 60)])
def test_world_step_bit_exact_with_external_commands(shape, steps):
    """Free-running World.step parity: random force/torque commands, no re-synchronisation, compared every think step."""
    import parity
    A, R, P = shape
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=abi.PS_CTRL_EXTERNAL)
    e, o = _engine(cfg), _oracle(cfg)
    seeds = np.arange(A)
    e.reset(seeds)
    o.reset(seeds)
    rng = np.random.RandomState(1)
    toi = 0
    for k in range(steps):
        fwd = ((rng.rand(A, R) * 2 - 0.5) * 750000).astype(np.float32)
        tq = ((rng.rand(A, R) * 2 - 1) * 2375000).astype(np.float32)
        e.step_external(fwd, tq)
        o.step_external(fwd, tq)
        bad = parity.compare_states(e.get_state(), o.get_state(), A, what=("body", "sense_aabb", "fat_aabb", "contacts", "arena"))
        assert bad == [], "think step %d: %s" % (k, bad[:5])
        toi += o.step_info(0)["toi_events"]
        assert e.step_info(0)["pos_iters"] == o.step_info(0)["pos_iters"]
    assert toi > 0 or steps < 100   # the long case exercises the TOI path


def test_camera_and_localmap_bit_exact():
    """Sense-only parity from identical states taken along an oracle trajectory: camera image, occupancy grid,
    perceived pucks, clusters, carry flag."""
    import parity
    A, R, P = 3, 6, 60
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, n_puck_types=3)
    e, o = _engine(cfg), _oracle(cfg)
    o.reset(np.arange(7, 7 + A))
    seen_puck = seen_robot = seen_carry = 0
    for k in range(40):
        o.step(9)
        st = o.get_state()
        e.set_state(st)
        o.sense()
        e.sense()
        oo, eo = o.get_observations(), e.get_observations()
        bad = parity.compare_observations(eo, oo)
        assert bad == [], "sample %d: %s" % (k, bad[:5])
        o.set_state(st)
        seen_puck += int((oo.camera < 8).sum())
        seen_robot += int((oo.camera == abi.PS_T_ROBOT).sum())
        seen_carry += int(oo.localmap["carrying"].sum())
    assert seen_puck > 0 and seen_robot > 0 and seen_carry > 0


@pytest.mark.parametrize("controller", [abi.PS_CTRL_CACHECONS, abi.PS_CTRL_BHD, abi.PS_CTRL_PROBSEEK])
def test_full_tick_lockstep(controller):
    """Stepping the reference restatement and the engine one think interval from identical states."""
    import parity
    A, R, P = 4, 4, 40
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=controller)
    e, o = _engine(cfg), _oracle(cfg)
    o.reset(np.arange(A))
    exact = 0
    for k in range(150):
        e.set_state(o.get_state())
        e.step(3)
        o.step(3)
        se, so = e.get_state(), o.get_state()
        bad = parity.compare_states(se, so, A, tol=TOL)
        assert bad == [], "think step %d: %s" % (k, bad[:5])
        exact += parity.compare_states(se, so, A) == []
    print("bit-exact think steps: %d / 150" % exact)


@pytest.mark.parametrize("controller", [abi.PS_CTRL_CACHECONS, abi.PS_CTRL_BHD, abi.PS_CTRL_PROBSEEK])
def test_full_tick_free_running(controller):
    """No re-synchronisation: both sides run 100 think steps (300 ticks) from the same seeds."""
    import parity
    A, R, P = 2, 4, 40
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=controller)
    e, o = _engine(cfg), _oracle(cfg)
    e.reset(np.arange(A))
    o.reset(np.arange(A))
    e.step(300)
    o.step(300)
    bad = parity.compare_states(e.get_state(), o.get_state(), A, tol=TOL)
    assert bad == [], bad[:5]


@pytest.mark.parametrize("controller", [abi.PS_CTRL_CACHECONS, abi.PS_CTRL_BHD, abi.PS_CTRL_PROBSEEK])
def test_lockstep_with_wide_cluster_threshold(controller):
    """LocalMap.CLUSTER_DISTANCE_THRESHOLD = 8 (touching pucks cluster): multi-puck clusters form, so the deposit states
    (PUSH / BACKUP / TURN, carried-cluster logic, the warp-built cluster lists) are exercised, which the default 1.5 never does."""
    import parity
    A, R, P = 4, 8, 120
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=controller, cluster_distance_threshold=8.0)
    e, o = _engine(cfg), _oracle(cfg)
    o.reset(np.arange(A))
    seen = np.zeros(8, int)
    for k in range(240):
        e.set_state(o.get_state())
        e.step(3)
        o.step(3)
        se, so = e.get_state(), o.get_state()
        bad = parity.compare_states(se, so, A, tol=TOL)
        assert bad == [], "think step %d: %s" % (k, bad[:5])
        for st in so.robot["ctrl_state"].ravel():
            seen[st] += 1
    n_states = {abi.PS_CTRL_CACHECONS: 5, abi.PS_CTRL_BHD: 4, abi.PS_CTRL_PROBSEEK: 7}[controller]
    assert (seen[:n_states] > 0).all(), seen


def test_dense_arena_big_islands():
    """A crowded arena (150 pucks, 6 robots, enclosure scale 0.6): puck clusters grow past the warp slot (48 bodies / 64
    contacts), so the big-island solver runs; lock-step parity with the oracle every think step."""
    import parity
    A = 2
    cfg = abi.config_defaults(n_arenas=A, n_robots=6, n_pucks=150, enclosure_scale=0.6)
    e, o = _engine(cfg), _oracle(cfg)
    o.reset(np.arange(A))
    e.island_hist()
    most_bodies = most_contacts = 0
    for k in range(100):
        e.set_state(o.get_state())
        e.step(3)
        o.step(3)
        bad = parity.compare_states(e.get_state(), o.get_state(), A, tol=TOL)
        assert bad == [], "think step %d: %s" % (k, bad[:5])
        if k % 10 == 9:
            w = int(e.island_hist()[7, 7])   # slowest warp-solved island since the last read: contacts << 20 | bodies << 10 | sweeps
            most_contacts, most_bodies = max(most_contacts, (w >> 20) & 0xFFF), max(most_bodies, (w >> 10) & 0x3FF)
    assert most_bodies > 48 or most_contacts > 64, (most_bodies, most_contacts)


def test_fast_mode_is_oracle_exact():
    """FAST mode = a lower position-iteration cap (ps_config.pos_iters = 10; the reference passes 100). It is a different
    experiment configuration, not a different engine: the oracle with the same cap must still be reproduced bit for bit."""
    import parity
    A = 6
    cfg = abi.config_defaults(n_arenas=A, n_robots=8, n_pucks=120, controller=abi.PS_CTRL_CACHECONS, pos_iters=10)
    e, o = _engine(cfg), _oracle(cfg)
    e.reset(np.arange(A))
    o.reset(np.arange(A))
    for k in range(4):
        e.step(90)
        o.step(90)
        bad = parity.compare_states(e.get_state(), o.get_state(), A, tol=TOL)
        assert bad == [], "tick %d: %s" % (90 * (k + 1), bad[:5])
    assert max(o.step_info(a)["pos_iters"] for a in range(A)) <= 10


def test_c3_shape_against_oracle():
    """16 robots / 200 pucks (the shape of the headline config), CacheCons on device, 20 think steps free-running."""
    import parity
    A = 4
    cfg = abi.config_defaults(n_arenas=A, n_robots=16, n_pucks=200, controller=abi.PS_CTRL_CACHECONS)
    e, o = _engine(cfg), _oracle(cfg)
    seeds = np.arange(40, 40 + A)
    e.reset(seeds)
    o.reset(seeds)
    e.step(60)
    o.step(60)
    bad = parity.compare_states(e.get_state(), o.get_state(), A, tol=TOL)
    assert bad == [], bad[:5]


def test_c3_shape_steady_state_free_running():
    """The regime the benchmark is defined on (after >= 300 warm-up ticks): 16 robots / 200 pucks, CacheCons on
    device, 900 ticks free-running with no re-synchronisation -- puck caches form, islands outgrow the lane and warp slots and
    run into the 100-sweep cap. Compared with the oracle at ticks 300, 600 and 900 (contact cache, bodies, controllers)."""
    import os
    import parity
    A = int(os.environ.get("PS_TEST_STEADY_ARENAS", "8"))
    cfg = abi.config_defaults(n_arenas=A, n_robots=16, n_pucks=200, controller=abi.PS_CTRL_CACHECONS)
    e, o = _engine(cfg), _oracle(cfg)
    seeds = np.arange(7, 7 + A)
    e.reset(seeds)
    o.reset(seeds)
    e.island_hist()
    capped = 0
    for k in range(3):
        e.step(300)
        o.step(300)
        bad = parity.compare_states(e.get_state(), o.get_state(), A, tol=TOL)
        assert bad == [], "tick %d: %s" % (300 * (k + 1), bad[:5])
        capped += int(max(o.step_info(a)["pos_iters"] for a in range(A)) >= 100)
    assert e.error_flags() == (0, 0)
    hist = e.island_hist(reset=False)
    assert hist[7, :7].sum() > 0, "no lane island ever ran into the 100-sweep cap: not the steady-state regime"
    assert capped > 0


def test_batch_position_independence_and_determinism():
    """Size-independent properties at a large batch: an arena's trajectory depends only on its seed, not on where it
    sits in the batch or how large the batch is; two runs give identical bits."""
    A = 512
    cfg = abi.config_defaults(n_arenas=A, n_robots=16, n_pucks=200, rng_mode=abi.PS_RNG_JAVA_SHARED)
    seeds = np.arange(A)
    e = _engine(cfg)
    e.reset(seeds)
    e.step(12)
    s1 = e.get_state()
    e2 = _engine(cfg)
    e2.reset(seeds[::-1].copy())
    e2.step(12)
    s2 = e2.get_state()
    assert np.array_equal(s1.body.view(np.uint32), s2.body[::-1].view(np.uint32))
    assert np.array_equal(s1.arena["n_contacts"], s2.arena["n_contacts"][::-1])
    assert np.array_equal(s1.contact_key, s2.contact_key[::-1])
    assert (s1.arena["error_flags"] == 0).all()
    small = abi.config_defaults(n_arenas=2, n_robots=16, n_pucks=200)
    e3 = _engine(small)
    e3.reset(seeds[5:7])
    e3.step(12)
    s3 = e3.get_state()
    assert np.array_equal(s3.body.view(np.uint32), s1.body[5:7].view(np.uint32))
    # physical sanity at scale: finite, inside the enclosure, contact cache == overlapping fat boxes invariant
    assert np.isfinite(s1.body).all()
    assert (np.abs(s1.body[:, 0:2, :]) < 187.0 / 2 + 1).all()


def test_contact_cache_invariant():
    """The cache must equal the set of filtered proxy pairs whose fat AABBs overlap (what findNewContacts / collide keep)."""
    cfg = abi.config_defaults(n_arenas=2, n_robots=8, n_pucks=100)
    e = _engine(cfg)
    e.reset([3, 4])
    e.step(30)
    st = e.get_state()
    P, R = 100, 8
    for a in range(2):
        fat = st.fat_aabb[a]   # [4][NP]
        n = int(st.arena[a]["n_contacts"])
        keys = st.contact_key[a, :n]
        dyn = set()
        for k in keys:
            pa, pb = int(k >> 16), int(k & 0xFFFF)
            if pa >= 84 and pb >= 84:
                dyn.add((min(pa, pb), max(pa, pb)))
        assert len(set(keys.tolist())) == n, "duplicate contacts"
        # brute force over dynamic proxies
        NP = fat.shape[1]
        body = np.array([(i // 2) if i < 2 * P else P + (i - 2 * P) // 14 for i in range(NP)])
        kind = np.array([0 if i < 2 * P else (2 if (i - 2 * P) % 14 == 13 else 1) for i in range(NP)])  # 0 puck, 1 robot poly, 2 bow
        lx, ly, ux, uy = fat
        ov = (lx[:, None] <= ux[None, :]) & (lx[None, :] <= ux[:, None]) & (ly[:, None] <= uy[None, :]) & (ly[None, :] <= uy[:, None])
        ok = body[:, None] != body[None, :]
        ok &= ~(((kind[:, None] == 2) & (kind[None, :] == 0)) | ((kind[:, None] == 0) & (kind[None, :] == 2)))
        ii, jj = np.nonzero(np.triu(ov & ok, 1))
        want = set((int(i) + 84, int(j) + 84) for i, j in zip(ii, jj))
        # every overlapping pair has a contact; extras are contacts whose boxes separated this tick (World.step
        # destroys them in the next ContactManager.collide)
        assert want <= dyn, sorted(want - dyn)[:5]
        assert len(dyn - want) < 0.2 * len(dyn) + 5


def test_stats_match_oracle():
    cfg = abi.config_defaults(n_arenas=6, n_robots=4, n_pucks=40)
    e, o = _engine(cfg), _oracle(cfg)
    e.reset(np.arange(6))
    o.reset(np.arange(6))
    e.step(90)
    o.step(90)
    for thr in (1.5, 9.45):
        se, so = e.compute_stats(thr), o.compute_stats(thr)
        assert np.array_equal(np.ctypeslib.as_array(se.cluster_size_hist), np.ctypeslib.as_array(so.cluster_size_hist))
        assert list(se.ctrl_state_hist) == list(so.ctrl_state_hist)
        assert (se.n_carrying, se.n_caches, se.n_contacts, se.n_touching_points) == (so.n_carrying, so.n_caches, so.n_contacts, so.n_touching_points)
        assert abs(se.sum_completion - so.sum_completion) < 1e-9 * max(1.0, so.sum_completion)


def test_errors_are_loud():
    from puckswarm_b200 import engine
    cfg = abi.config_defaults(n_arenas=1)
    e = engine.Engine(cfg, device=0, load_calibration=False)
    e.reset([0])
    with pytest.raises(engine.EngineError) as ei:
        e.step(3)          # CacheCons on device needs the calibration
    assert ei.value.code == abi.PS_E_NOCALIB
    bad = abi.config_defaults(n_arenas=0)
    with pytest.raises(engine.EngineError):
        engine.Engine(bad, device=0)
    with pytest.raises(engine.EngineError):      # stepCount % 0 throws in the reference
        engine.Engine(abi.config_defaults(n_arenas=1, puck_shift_step=0), device=0)


def test_baseline_config_1_and_2():
    """BASELINE configs[0] (1 arena, 2 robots, CacheCons, default pucks) and configs[1] (1 arena, 4 robots, BHD, 2 puck
    colours; seeds 0..19 as ExperimentManager.java:274-281,368): free-running against the oracle."""
    import parity
    c1 = abi.config_defaults(n_arenas=1, n_robots=2, n_pucks=10, n_puck_types=2, controller=abi.PS_CTRL_CACHECONS)
    e, o = _engine(c1), _oracle(c1)
    e.reset([0])
    o.reset([0])
    e.step(600)
    o.step(600)
    assert parity.compare_states(e.get_state(), o.get_state(), 1, tol=TOL) == []
    c2 = abi.config_defaults(n_arenas=20, n_robots=4, n_pucks=40, n_puck_types=2, controller=abi.PS_CTRL_BHD)
    e, o = _engine(c2), _oracle(c2)
    e.reset(np.arange(20))
    o.reset(np.arange(20))
    for _ in range(4):
        e.step(75)
        o.step(75)
        assert parity.compare_states(e.get_state(), o.get_state(), 20, tol=TOL) == []


def test_single_tick_parity_from_identical_states():
    """The north_star's correctness statement, literally: one tick from identical states, many states."""
    import parity
    A, R, P = 6, 4, 40
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=abi.PS_CTRL_BHD, step_interval=1)
    e, o = _engine(cfg), _oracle(cfg)
    o.reset(np.arange(A))
    for k in range(120):
        e.set_state(o.get_state())
        e.step(1)
        o.step(1)
        se, so = e.get_state(), o.get_state()
        assert parity.compare_states(se, so, A, tol=TOL) == [], k
        # touching pair lists (contacts with manifold points) identical, in order
        for a in range(A):
            ke, ie, _ = se.contacts(a)
            ko, io, _ = so.contacts(a)
            assert np.array_equal(ke[(ie & 3) > 0], ko[(io & 3) > 0])


def test_large_arena_shape_c5():
    """BASELINE configs[4] shape: 256 robots / 4000 pucks, enclosure scale 6 (one arena; the oracle needs ~0.2 s per tick
    and ~25 s per sense at this size): World.step under commands bit-exact, one sense bit-exact."""
    import parity
    cfg = abi.config_defaults(n_arenas=1, n_robots=256, n_pucks=4000, enclosure_scale=6.0, controller=abi.PS_CTRL_EXTERNAL)
    e, o = _engine(cfg), _oracle(cfg, threads=1)
    e.reset([0])
    o.reset([0])
    assert parity.compare_states(e.get_state(), o.get_state(), 1) == []
    rng = np.random.RandomState(5)
    for k in range(6):
        fwd = ((rng.rand(1, 256) * 1.5) * 750000).astype(np.float32)
        tq = ((rng.rand(1, 256) * 2 - 1) * 2375000).astype(np.float32)
        e.step_external(fwd, tq)
        o.step_external(fwd, tq)
        bad = parity.compare_states(e.get_state(), o.get_state(), 1, what=("body", "sense_aabb", "fat_aabb", "contacts", "arena"))
        assert bad == [], "think step %d: %s" % (k, bad[:5])
    o.sense()
    e.sense()
    bad = parity.compare_observations(e.get_observations(camera=False), o.get_observations(camera=False))
    assert bad == [], bad[:5]
    assert (e.get_state().arena["error_flags"] == 0).all()


def test_ensemble_statistics_match():
    """Ensemble statistics over seeded episodes (the north_star asks for matching distributions; here they are compared
    exactly): 64 episodes x 150 think steps, cluster-size histograms at the paper threshold and controller-state occupancy."""
    cfg = abi.config_defaults(n_arenas=64, n_robots=4, n_pucks=40, controller=abi.PS_CTRL_CACHECONS)
    e, o = _engine(cfg), _oracle(cfg)
    e.reset(np.arange(64))
    o.reset(np.arange(64))
    for _ in range(3):
        e.step(150)
        o.step(150)
        se, so = e.compute_stats(9.45), o.compute_stats(9.45)
        assert np.array_equal(np.ctypeslib.as_array(se.cluster_size_hist), np.ctypeslib.as_array(so.cluster_size_hist))
        assert list(se.ctrl_state_hist) == list(so.ctrl_state_hist)
        assert se.n_carrying == so.n_carrying and se.n_caches == so.n_caches
        assert abs(se.sum_completion - so.sum_completion) < 1e-9 * max(1.0, so.sum_completion)


@pytest.mark.parametrize("controller,extra", [(abi.PS_CTRL_CACHECONS, {}), (abi.PS_CTRL_BHD, {}), (abi.PS_CTRL_CACHECONS, dict(puck_shuffle_step=100)),
                                              (abi.PS_CTRL_PROBSEEK, dict(n_pucks=41))])
def test_periodic_record_and_analysis_match_oracle(controller, extra):
    """The record the reference writes every 100 think steps (Arena.java:223-243, CacheConsController.java:313-333,531-551) and
    analysis.Analyze's per-repetition statistics (Analyze.java:176-218), kept on the device, against the oracle's restatement:
    storage steps 0, 100 and 200; poses and pucks after the perturbation of the same step; counters before they are cleared."""
    A = 6
    kw = dict(n_arenas=A, n_robots=4, n_pucks=40, controller=controller)
    kw.update(extra)
    cfg = abi.config_defaults(**kw)
    e, o = _engine(cfg), _oracle(cfg)
    for x in (e, o):
        x.analysis_config(9.45, 12.0)
        x.reset(np.arange(A))
    for ticks in (1, 300, 300):       # just past think steps 0, 100, 200
        e.step(ticks)
        o.step(ticks)
        se, so = e.get_snapshot(), o.get_snapshot()
        assert np.array_equal(se.step_count, so.step_count) and se.step_count[0] in (0, 100, 200)
        assert np.array_equal(se.robot_pose.view(np.uint32), so.robot_pose.view(np.uint32))
        assert np.array_equal(se.puck_xy.view(np.uint32), so.puck_xy.view(np.uint32))
        assert np.array_equal(se.state_counts, so.state_counts)
        assert np.array_equal(np.isnan(se.cache_xy), np.isnan(so.cache_xy))
        assert np.allclose(np.nan_to_num(se.cache_xy), np.nan_to_num(so.cache_xy), rtol=TOL, atol=0)
        ae, ao = e.get_analysis(), o.get_analysis()
        for f in ("n_snapshots", "steps_to_completion"):
            assert np.array_equal(ae[f], ao[f]), f
        for f in ("twc_num", "twc_den", "completion"):
            assert np.allclose(ae[f], ao[f], rtol=1e-12, atol=0), f
    assert (ae["n_snapshots"] == 3).all() and (se.state_counts.sum(axis=-1) == 100).all()
    te, to = e.compute_stats(9.45), o.compute_stats(9.45)
    for f in ("n_completed", "sum_steps_to_completion", "n_snapshots", "min_steps_to_completion", "max_steps_to_completion"):
        assert getattr(te, f) == getattr(to, f), f
    for f in ("sum_completion", "sum_twc", "min_completion", "max_completion", "min_twc", "max_twc"):
        assert abs(getattr(te, f) - getattr(to, f)) <= 1e-9 * max(1.0, abs(getattr(to, f))), f
    assert to.n_snapshots == 3 * A and (to.n_completed > 0 or controller != abi.PS_CTRL_CACHECONS)


def test_ensemble_1000_episodes():
    """north_star: ensemble statistics over 1 000 seeded episodes must match the reference's distributions. Here 1 000 episodes
    (4 robots / 40 pucks / 2 colours, CacheCons, seeds 0..999) x 150 think steps are compared exactly: per-episode percent
    completion and controller states, and the ensemble's cluster-size histograms."""
    import os
    n = int(os.environ.get("PS_TEST_EPISODES", "1000"))
    cfg = abi.config_defaults(n_arenas=n, n_robots=4, n_pucks=40, controller=abi.PS_CTRL_CACHECONS)
    e, o = _engine(cfg), _oracle(cfg)
    e.reset(np.arange(n))
    o.reset(np.arange(n))
    e.step(450)
    o.step(450)
    se, so = e.compute_stats(9.45), o.compute_stats(9.45)
    assert np.array_equal(np.ctypeslib.as_array(se.cluster_size_hist), np.ctypeslib.as_array(so.cluster_size_hist))
    assert list(se.ctrl_state_hist) == list(so.ctrl_state_hist)
    assert se.n_carrying == so.n_carrying and se.n_caches == so.n_caches
    assert abs(se.sum_completion - so.sum_completion) < 1e-9 * max(1.0, so.sum_completion)
    ge, go = e.get_state(), o.get_state()
    # per episode: every puck and robot pose within TOL (bit-exact in practice), same controller states
    assert np.array_equal(ge.robot["ctrl_state"], go.robot["ctrl_state"])
    err = np.abs(ge.body - go.body) / np.maximum(1.0, np.abs(go.body))
    assert err.max() <= TOL, "worst episode %d" % int(np.unravel_index(err.argmax(), err.shape)[0])
    assert e.error_flags() == (0, 0)


@pytest.mark.parametrize("controller,shift,shuffle,shape", [(abi.PS_CTRL_EXTERNAL, 4, 6, (3, 4, 60)), (abi.PS_CTRL_CACHECONS, 5, -1, (5, 4, 60)),
                                                            (abi.PS_CTRL_BHD, 1, 1, (6, 3, 200)), (abi.PS_CTRL_CACHECONS, 3, 4, (4, 16, 200))])
def test_puck_shift_and_shuffle(controller, shift, shuffle, shape):
    """Arena.puckShiftStep / puckShuffleStep (Arena.java:185-192, 318-373) on the device: pucks teleported through
    Body.setTransform + findNewContacts at the head of the due think steps, arena groups perturbing concurrently; state
    incl. contact cache order and the shared RNG bit-exact. The (1, 1) case re-teleports 200 pucks twice per think step,
    which exercises addPair's "contact exists already" walk against stale contacts."""
    import parity
    A, R, P = shape
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=controller, puck_shift_step=shift, puck_shuffle_step=shuffle)
    e, o = _engine(cfg), _oracle(cfg)
    seeds = np.arange(20, 20 + A)
    e.reset(seeds)
    o.reset(seeds)
    rng = np.random.RandomState(1)
    for k in range(9):
        if controller == abi.PS_CTRL_EXTERNAL:
            fwd = ((rng.rand(A, R) * 1.5) * 750000).astype(np.float32)
            tq = ((rng.rand(A, R) * 2 - 1) * 2375000).astype(np.float32)
            o.step_external(fwd, tq)
            e.step_external(fwd, tq)
        else:
            o.step(3)
            e.step(3)
        # on-device controllers: double-precision libm may differ in the last ulp (see TOL above); the rest is bit-exact
        bad = parity.compare_states(e.get_state(), o.get_state(), A, tol=0.0 if controller == abi.PS_CTRL_EXTERNAL else TOL)
        assert bad == [], "think step %d: %s" % (k, bad[:4])
    assert (e.get_state().arena["error_flags"] == 0).all()


@pytest.mark.parametrize("controller", [abi.PS_CTRL_CACHECONS, abi.PS_CTRL_BHD])
def test_per_robot_rng_streams(controller):
    """PS_RNG_PER_ROBOT: one java.util.Random per robot (the engine's throughput option) -- the oracle runs the same streams."""
    import parity
    A, R, P = 4, 6, 80
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=controller, rng_mode=abi.PS_RNG_PER_ROBOT)
    e, o = _engine(cfg), _oracle(cfg)
    e.reset(np.arange(A))
    o.reset(np.arange(A))
    for k in range(40):
        o.step(3)
        e.step(3)
        bad = parity.compare_states(e.get_state(), o.get_state(), A, tol=TOL)
        assert bad == [], "think step %d: %s" % (k, bad[:4])


@pytest.mark.parametrize("threads", ["32", "64"])
def test_think_stage_cta_widths(threads, monkeypatch):
    """k_think picks its CTA width by batch size (one warp per arena for launches that fill the GPU); the narrow widths,
    which small test batches would not reach, forced here."""
    import parity
    monkeypatch.setenv("PS_THINK_THREADS", threads)
    A, R, P = 3, 8, 100
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=abi.PS_CTRL_CACHECONS)
    e, o = _engine(cfg), _oracle(cfg)
    e.reset(np.arange(A))
    o.reset(np.arange(A))
    for k in range(40):
        o.step(3)
        e.step(3)
        bad = parity.compare_states(e.get_state(), o.get_state(), A, tol=TOL)
        assert bad == [], "think step %d: %s" % (k, bad[:4])


@pytest.mark.parametrize("env", [{"PS_LANEX_MAX_CONTACTS": "27"}, {"PS_ISLAND_G": "8"}, {"PS_ISLAND_G": "4", "PS_ISLAND_SPLIT": "1"}, {"PS_HEAVY_ARENAS": "2"},
                                 {"PS_GROUPS": "3"}, {"PS_BATCH_THREADS": "32", "PS_MERGE_BATCH": "0"}, {"PS_PRE_THREADS": "128", "PS_PRE_SMEM_T": "64"},
                                 {"PS_PRE_THREADS": "32", "PS_POST_THREADS": "64"}])
def test_scheduling_options_keep_parity(env, monkeypatch):
    """The island / stage scheduling knobs only decide which kernel solves an island and which arenas share a stream; every
    one of them -- the measured-and-off ones included (wide lane slots, split solver, heavy / light arena split) -- must leave the
    trajectories bit-identical. 16 robots / 200 pucks, 240 ticks free-running (islands of every bucket form), compared every 60 ticks."""
    import parity
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    A = 6
    cfg = abi.config_defaults(n_arenas=A, n_robots=16, n_pucks=200, controller=abi.PS_CTRL_CACHECONS)
    e, o = _engine(cfg), _oracle(cfg)
    seeds = np.arange(11, 11 + A)
    e.reset(seeds)
    o.reset(seeds)
    for k in range(4):
        e.step(60)
        o.step(60)
        bad = parity.compare_states(e.get_state(), o.get_state(), A, tol=TOL)
        assert bad == [], "%s, tick %d: %s" % (env, 60 * (k + 1), bad[:4])
    assert e.error_flags() == (0, 0)


@pytest.mark.parametrize("shape", [(1, 2, 0, 2, abi.PS_CTRL_BHD), (2, 3, 7, 2, abi.PS_CTRL_CACHECONS), (1, 1, 5, 5, abi.PS_CTRL_PROBSEEK),
                                   (1, 6, 3, 3, abi.PS_CTRL_CACHECONS), (1, 2, 16, 8, abi.PS_CTRL_CACHECONS)])
def test_edge_shapes(shape):
    """Empty and ragged inputs on the device: no pucks, a puck count the colours do not divide, a single robot, more robots
    than pucks, all eight colours."""
    import parity
    A, R, P, K, ctrl = shape
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, n_puck_types=K, controller=ctrl)
    e, o = _engine(cfg), _oracle(cfg)
    e.reset(np.arange(A))
    o.reset(np.arange(A))
    for k in range(30):
        o.step(3)
        e.step(3)
        bad = parity.compare_states(e.get_state(), o.get_state(), A, tol=TOL)
        assert bad == [], "think step %d: %s" % (k, bad[:4])
    assert (e.get_state().arena["error_flags"] == 0).all()


def test_informed_robots():
    """Arena.nInformedRobots: the first robots run CacheCons with preset cache points (CacheConsController.java:209-243)."""
    import parity
    A, R, P = 3, 4, 40
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=abi.PS_CTRL_CACHECONS, n_informed_robots=2, cc_ignore_pucks=1)
    e, o = _engine(cfg), _oracle(cfg)
    e.reset(np.arange(5, 5 + A))
    o.reset(np.arange(5, 5 + A))
    assert (e.get_state().robot["cache_valid"][:, :2, :2] == 1).all()
    for k in range(60):
        o.step(3)
        e.step(3)
        bad = parity.compare_states(e.get_state(), o.get_state(), A, tol=TOL)
        assert bad == [], "think step %d: %s" % (k, bad[:4])


@pytest.mark.parametrize("kw", [dict(cc_cache_selection=abi.PS_SEL_PROB_ABSOLUTE, cc_cache_separation=abi.PS_SEP_NULLIFY_OLD, cc_forget_prob=0.05),
                                dict(cc_cache_selection=abi.PS_SEL_RELATIVE_FORGET, cc_cache_separation=abi.PS_SEP_ACCEPT_ISOLATED, cc_forget_prob=0.05),
                                dict(cc_cache_selection=abi.PS_SEL_PROB_RELATIVE, cc_cache_separation=abi.PS_SEP_IGNORE),
                                dict(sincos_lut=0), dict(toi_enabled=0), dict(n_puck_types=3, n_pucks=60), dict(step_interval=1),
                                dict(vel_iters=1, pos_iters=7), dict(enclosure_scale=2.0)],
                         ids=lambda kw: ",".join("%s=%s" % it for it in kw.items()))
def test_config_option_sweep(kw):
    """A sample of the option space of tests/test_hostemu_parity.py::test_config_option_sweep on the device."""
    import parity
    base = dict(n_arenas=3, n_robots=4, n_pucks=50, controller=abi.PS_CTRL_CACHECONS, cluster_distance_threshold=8.0)
    base.update(kw)
    cfg = abi.config_defaults(**base)
    e, o = _engine(cfg), _oracle(cfg)
    e.reset([1, 2, 3])
    o.reset([1, 2, 3])
    for k in range(50):
        o.step(cfg.step_interval)
        e.step(cfg.step_interval)
        bad = parity.compare_states(e.get_state(), o.get_state(), 3, tol=TOL)
        assert bad == [], "think step %d: %s" % (k, bad[:4])
