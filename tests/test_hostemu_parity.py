"""CPU-only check of the engine's device programs: the same templates the kernels instantiate with a CUDA block
context are instantiated with a one-thread host context (tests/hostemu) and compared with the oracle.
This validates the device logic (not the GPU execution -- that is tests/test_gpu_parity.py)."""
import numpy as np
import pytest

from puckswarm_b200 import abi
import emu_binding as emu
import oracle_binding as oracle
import parity


@pytest.mark.parametrize("shape", [(2, 2, 10, 2), (2, 16, 200, 2), (1, 5, 21, 3)])
def test_reset(shape):
    A, R, P, K = shape
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, n_puck_types=K)
    e, o = emu.Emu(cfg), oracle.Oracle(cfg)
    seeds = np.arange(11, 11 + A)
    e.reset(seeds)
    o.reset(seeds)
    assert parity.compare_states(e.get_state(), o.get_state(), A) == []


@pytest.mark.parametrize("shape,steps", [((2, 4, 40), 150), ((1, 16, 200), 40)])
def test_world_step_external(shape, steps):
    A, R, P = shape
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=abi.PS_CTRL_EXTERNAL)
    e, o = emu.Emu(cfg), oracle.Oracle(cfg)
    e.reset(np.arange(A))
    o.reset(np.arange(A))
    rng = np.random.RandomState(0)
    se = e.get_state()
    for k in range(steps):
        fwd = ((rng.rand(A, R) * 2 - 0.5) * 750000).astype(np.float32)
        tq = ((rng.rand(A, R) * 2 - 1) * 2375000).astype(np.float32)
        o.step_external(fwd, tq)
        se.robot["cmd_forwards"], se.robot["cmd_torque"] = fwd, tq
        e.set_state(se)
        e.step(3)
        se = e.get_state()
        bad = parity.compare_states(se, o.get_state(), A, what=("body", "sense_aabb", "fat_aabb", "contacts"))
        assert bad == [], "think step %d: %s" % (k, bad[:4])
        assert np.array_equal(e.solve_order(0), o.solve_order(0))


@pytest.mark.parametrize("controller", [abi.PS_CTRL_CACHECONS, abi.PS_CTRL_BHD, abi.PS_CTRL_PROBSEEK])
def test_sense_think_step(controller):
    A, R, P = 2, 4, 40
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=controller)
    e, o = emu.Emu(cfg), oracle.Oracle(cfg)
    e.reset([0, 1])
    o.reset([0, 1])
    for k in range(120):
        if k % 15 == 0:
            st = o.get_state()
            e.set_state(st)
            o.sense()
            bad = parity.compare_observations(e.sense(), o.get_observations())
            assert bad == [], "sense %d: %s" % (k, bad[:4])
            o.set_state(st)
            e.set_state(st)
        o.step(3)
        e.step(3)
        bad = parity.compare_states(e.get_state(), o.get_state(), A)
        assert bad == [], "think step %d: %s" % (k, bad[:4])


@pytest.mark.parametrize("controller", [abi.PS_CTRL_CACHECONS, abi.PS_CTRL_PROBSEEK])
def test_think_with_wide_cluster_threshold(controller):
    """Cluster threshold 8 (touching pucks cluster): the deposit states and the carried-cluster logic run, bit-exact."""
    A, R, P = 2, 6, 90
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=controller, cluster_distance_threshold=8.0)
    e, o = emu.Emu(cfg), oracle.Oracle(cfg)
    e.reset([0, 1])
    o.reset([0, 1])
    seen = np.zeros(8, int)
    for k in range(150):
        o.step(3)
        e.step(3)
        so = o.get_state()
        bad = parity.compare_states(e.get_state(), so, A)
        assert bad == [], "think step %d: %s" % (k, bad[:4])
        for st in so.robot["ctrl_state"].ravel():
            seen[st] += 1
    assert (seen[:4] > 0).all(), seen


def test_narrow_phase_samples():
    """Manifolds of every shape-pair kind at random relative poses: type, points, ids bit-exact."""
    cfg = abi.config_defaults(n_arenas=1, n_robots=2, n_pucks=2)
    e, o = emu.Emu(cfg), oracle.Oracle(cfg)
    e.reset([0])
    o.reset([0])
    rng = np.random.RandomState(3)
    r0, r1 = 84 + 4, 84 + 4 + 14
    pairs = [(84, 86), (r0, 84), (r0 + 5, 85), (0, 84), (83, 84), (r0, r1), (r0 + 1, r1 + 7), (r0 + 13, r1 + 13), (40, r0), (81, r0 + 13), (r0 + 3, r1 + 4)]
    hits = 0
    for pa, pb in pairs:
        for _ in range(300):
            if pa < 84:
                xa = (0.0, 0.0, 0.0)
                xb = (rng.uniform(-95, 95), rng.uniform(-95, 95), rng.uniform(-7, 7))
            else:
                xa = (rng.uniform(-5, 5), rng.uniform(-5, 5), rng.uniform(-7, 7))
                xb = (xa[0] + rng.uniform(-14, 14), xa[1] + rng.uniform(-14, 14), rng.uniform(-7, 7))
            me, mo = e.manifold(pa, xa, pb, xb), o.manifold(pa, xa, pb, xb)
            n = int(mo[0])
            assert me[0] == mo[0]
            if n:
                hits += 1
                assert np.array_equal(me[1:8].view(np.uint32), mo[1:8].view(np.uint32)) and me[10] == mo[10]
                if n == 2:
                    assert np.array_equal(me[8:10].view(np.uint32), mo[8:10].view(np.uint32)) and me[11] == mo[11]
    assert hits > 100


@pytest.mark.parametrize("controller,shift,shuffle", [(abi.PS_CTRL_EXTERNAL, 4, 6), (abi.PS_CTRL_CACHECONS, 5, -1), (abi.PS_CTRL_BHD, -1, 7)])
def test_puck_shift_and_shuffle(controller, shift, shuffle):
    """Arena.puckShiftStep / puckShuffleStep (Arena.java:185-192, 318-373): every puck teleported through
    Body.setTransform + findNewContacts at the head of the due think steps; state (incl. the contact cache, its order and
    the shared RNG) stays bit-exact through the perturbations and the World.steps that clean up after them."""
    A, R, P = 2, 4, 60
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=controller, puck_shift_step=shift, puck_shuffle_step=shuffle)
    e, o = emu.Emu(cfg), oracle.Oracle(cfg)
    e.reset([3, 4])
    o.reset([3, 4])
    rng = np.random.RandomState(1)
    moved = 0
    for k in range(16):
        before = o.get_state().body[:, 0:2, :P].copy()
        if controller == abi.PS_CTRL_EXTERNAL:
            fwd = ((rng.rand(A, R) * 1.5) * 750000).astype(np.float32)
            tq = ((rng.rand(A, R) * 2 - 1) * 2375000).astype(np.float32)
            o.step_external(fwd, tq)
            se = e.get_state()
            se.robot["cmd_forwards"], se.robot["cmd_torque"] = fwd, tq
            e.set_state(se)
            e.step(3)
        else:
            o.step(3)
            e.step(3)
        so = o.get_state()
        bad = parity.compare_states(e.get_state(), so, A)
        assert bad == [], "think step %d: %s" % (k, bad[:4])
        if np.abs(so.body[:, 0:2, :P] - before).max() > 20:
            moved += 1
    due = sum(1 for k in range(16) if (shift != -1 and k % shift == 0) or (shuffle != -1 and k % shuffle == 0))
    assert moved == due, (moved, due)


def test_puck_shuffle_every_step_dense():
    """200 pucks re-teleported twice per think step: stale contacts of earlier moves meet new overlaps, so addPair's
    "contact exists already" walk decides (without it the cache grows duplicates on the first step)."""
    A, R, P = 2, 3, 200
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=abi.PS_CTRL_BHD, puck_shift_step=1, puck_shuffle_step=1)
    e, o = emu.Emu(cfg), oracle.Oracle(cfg)
    e.reset([3, 4])
    o.reset([3, 4])
    for k in range(6):
        o.step(3)
        e.step(3)
        bad = parity.compare_states(e.get_state(), o.get_state(), A)
        assert bad == [], "think step %d: %s" % (k, bad[:4])


@pytest.mark.parametrize("controller", [abi.PS_CTRL_CACHECONS, abi.PS_CTRL_BHD, abi.PS_CTRL_PROBSEEK])
def test_per_robot_rng_streams(controller):
    """PS_RNG_PER_ROBOT (the engine's throughput option: one java.util.Random per robot instead of the experiment's shared
    one): same streams, same draws, the arena's generator untouched by the controllers."""
    A, R, P = 2, 5, 60
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=controller, rng_mode=abi.PS_RNG_PER_ROBOT)
    e, o = emu.Emu(cfg), oracle.Oracle(cfg)
    e.reset([7, 8])
    o.reset([7, 8])
    arena_seed = o.get_state().arena["rng_seed"].copy()
    for k in range(60):
        o.step(3)
        e.step(3)
        so = o.get_state()
        bad = parity.compare_states(e.get_state(), so, A)
        assert bad == [], "think step %d: %s" % (k, bad[:4])
    assert np.array_equal(so.arena["rng_seed"], arena_seed)
    assert len(np.unique(so.robot["rng_seed"])) == A * R


@pytest.mark.parametrize("shape", [(1, 2, 0, 2, abi.PS_CTRL_BHD), (2, 3, 7, 2, abi.PS_CTRL_CACHECONS), (1, 1, 5, 5, abi.PS_CTRL_PROBSEEK),
                                   (1, 6, 3, 3, abi.PS_CTRL_CACHECONS), (1, 2, 16, 8, abi.PS_CTRL_CACHECONS)])
def test_edge_shapes(shape):
    """Empty and ragged inputs: no pucks at all, a puck count the colours do not divide (Arena.java:91-95 drops the remainder),
    a single robot, more robots than pucks, all eight colours."""
    A, R, P, K, ctrl = shape
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, n_puck_types=K, controller=ctrl)
    e, o = emu.Emu(cfg), oracle.Oracle(cfg)
    e.reset(np.arange(A))
    o.reset(np.arange(A))
    assert e.get_state().body.shape[2] == P // K * K + R
    for k in range(30):
        o.step(3)
        e.step(3)
        bad = parity.compare_states(e.get_state(), o.get_state(), A)
        assert bad == [], "think step %d: %s" % (k, bad[:4])


@pytest.mark.parametrize("bad", [dict(puck_shift_step=0), dict(puck_shuffle_step=-3), dict(n_arenas=0), dict(n_puck_types=9), dict(step_interval=0),
                                 dict(controller=7)])
def test_config_validation(bad):
    """The host-side set-up shared by the engine and this harness (ps_host.hpp) refuses what the reference would crash on
    (stepCount % 0, Arena.java:189-191) or cannot represent."""
    with pytest.raises(RuntimeError):
        emu.Emu(abi.config_defaults(**bad))


def test_informed_robots():
    """Arena.nInformedRobots (Arena.java:101-106): the first robots get CacheConsController(presetCaches = true) -- two preset
    cache points, remembered sizes 1000, IGNORE_PUCKS forced off, no forgetting (CacheConsController.java:209-243, :558)."""
    A, R, P = 2, 4, 40
    cfg = abi.config_defaults(n_arenas=A, n_robots=R, n_pucks=P, controller=abi.PS_CTRL_CACHECONS, n_informed_robots=2, cc_ignore_pucks=1)
    e, o = emu.Emu(cfg), oracle.Oracle(cfg)
    e.reset([5, 6])
    o.reset([5, 6])
    s0 = o.get_state()
    assert (s0.robot["cache_valid"][:, :2, :2] == 1).all() and (s0.robot["cache_valid"][:, 2:, :] == 0).all()
    assert (s0.robot["remembered"][:, :2, :2] == 1000).all()
    for k in range(80):
        o.step(3)
        e.step(3)
        bad = parity.compare_states(e.get_state(), o.get_state(), A)
        assert bad == [], "think step %d: %s" % (k, bad[:4])


_OPTION_CASES = (
    [dict(cc_cache_selection=sel, cc_cache_separation=sep, cc_forget_prob=0.05)
     for sel in (abi.PS_SEL_PROB_ABSOLUTE, abi.PS_SEL_PROB_RELATIVE, abi.PS_SEL_RELATIVE, abi.PS_SEL_RELATIVE_FORGET)
     for sep in (abi.PS_SEP_IGNORE, abi.PS_SEP_NULLIFY_OLD, abi.PS_SEP_ACCEPT_ISOLATED, abi.PS_SEP_ACCEPT_LARGER)]
    + [dict(sincos_lut=0), dict(toi_enabled=0), dict(n_puck_types=3, n_pucks=60), dict(cc_ignore_pucks=1), dict(step_interval=1),
       dict(step_interval=5), dict(vel_iters=1, pos_iters=7), dict(enclosure_scale=2.0)])


@pytest.mark.parametrize("kw", _OPTION_CASES, ids=lambda kw: ",".join("%s=%s" % it for it in kw.items()))
def test_config_option_sweep(kw):
    """Every CacheCons cache-selection x cache-separation option (CacheConsController.java:136-177, with forgetting switched
    up so that it happens), the JBox2D switches (trig table, TOI), three colours, IGNORE_PUCKS, other think intervals, iteration
    counts and a larger enclosure: lock-step with the oracle at a cluster threshold that lets caches form."""
    base = dict(n_arenas=2, n_robots=4, n_pucks=50, controller=abi.PS_CTRL_CACHECONS, cluster_distance_threshold=8.0)
    base.update(kw)
    cfg = abi.config_defaults(**base)
    e, o = emu.Emu(cfg), oracle.Oracle(cfg)
    e.reset([1, 2])
    o.reset([1, 2])
    for k in range(60):
        o.step(cfg.step_interval)
        e.step(cfg.step_interval)
        bad = parity.compare_states(e.get_state(), o.get_state(), 2)
        assert bad == [], "think step %d: %s" % (k, bad[:4])


@pytest.mark.parametrize("ctrl,kw", [
    (abi.PS_CTRL_BHD, dict(bhd_push_in_time=3, bhd_backup_time=2)),
    (abi.PS_CTRL_PROBSEEK, dict(ps_placement_time=5, cc_k1=0.3, cc_k2=2.0, cc_pick_up_time=5, cc_backup_time=2)),
    (abi.PS_CTRL_CACHECONS, dict(cc_k1=0.2, cc_k2=3.0, cc_pick_up_time=4, cc_push_in_time=3, cc_backup_time=2, cc_exile_time=5, cc_spit_out_time=3)),
    (abi.PS_CTRL_CACHECONS, dict(cc_k_absolute=5.0, cc_k_relative=2.0, cc_cache_selection=abi.PS_SEL_PROB_ABSOLUTE, cc_home_separation_distance=20.0,
                                 cc_prediction_distance_sqd=25.0, cc_st_dev=2.0)),
    (abi.PS_CTRL_CACHECONS, dict(vfh_safety_gamma=4.0, vfh_safety_mask=2.0, vfh_tau_lo=0.5, vfh_tau_hi=0.6)),
    (abi.PS_CTRL_BHD, dict(carry_threshold_lo=0.05, carry_threshold_hi=0.2)),
    (abi.PS_CTRL_CACHECONS, dict(dt=1 / 30.0)),
    (abi.PS_CTRL_CACHECONS, dict(max_contacts=2048)),
], ids=lambda v: v if isinstance(v, int) else ",".join(v))
def test_numeric_property_keys(ctrl, kw):
    """The numeric keys of SURVEY App. C.8 away from their defaults (controller timers and constants, VFH+ safety distances and
    hysteresis, carry thresholds, time step, cache capacity): each reaches the device program and the oracle alike."""
    base = dict(n_arenas=2, n_robots=4, n_pucks=50, controller=ctrl, cluster_distance_threshold=8.0)
    base.update(kw)
    cfg = abi.config_defaults(**base)
    e, o = emu.Emu(cfg), oracle.Oracle(cfg)
    e.reset([1, 2])
    o.reset([1, 2])
    for k in range(60):
        o.step(3)
        e.step(3)
        bad = parity.compare_states(e.get_state(), o.get_state(), 2)
        assert bad == [], "think step %d: %s" % (k, bad[:4])
