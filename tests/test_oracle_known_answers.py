"""CPU tests that pin the oracle to every known answer available for this path:
java.util.Random vectors from the JDK specification, the scene's mass data, the two soft known
answers the reference itself states (about 30 cm per 21 full-force ticks, TestMovementController.java:26-27,81-82;
about 180 degrees per 30 full-torque ticks, Robot.java:48-49), calibration fixture counts, and analytic contacts.
PARITY IS UNPINNED beyond these: the reference ships no tests or golden vectors and cannot run here (no JVM)."""
import math
import numpy as np

from puckswarm_b200 import abi
import oracle_binding as oracle


def test_java_random_vectors():
    # well-known values of java.util.Random (JDK javadoc LCG): new Random(0).nextInt() etc.
    assert oracle.random_ints(0, 3).tolist() == [-1155484576, -723955400, 1033096058]
    assert oracle.random_ints(42, 2).tolist() == [-1170105035, 234785527]
    f = oracle.random_floats(0, 2)
    assert f[0] == np.float32(0.73096776) and f[1] == np.float32(0.831441)
    g = oracle.random_gaussians(0, 2)
    assert abs(g[0] - 0.8025330637390305) < 1e-15 and abs(g[1] - (-0.9015460884175122)) < 1e-15
    b = oracle.random_bounded(0, 6, 5).tolist()
    assert all(0 <= x < 6 for x in b)
    assert oracle.random_bounded(0, 10, 3).tolist() == [0, 8, 9]   # new Random(0).nextInt(10) x3


def test_sin_lut_definition():
    L = oracle.lib()
    for x in (0.0, 0.5, 1.0, -2.0, 7.0, 100.0):
        idx = int(np.float32(np.fmod(np.float32(x), np.float32(2 * math.pi)) + (np.float32(2 * math.pi) if x < 0 else 0)) / np.float32(0.00011) + np.float32(0.5))
        want = np.float32(math.sin(float(np.float32(idx % 57120) * np.float32(0.00011))))
        assert abs(L.orc_sin_lut(x) - want) < 2e-4
        assert abs(L.orc_sin_lut(x) - math.sin(x)) < 2e-4


def test_calibration_fixture_counts():
    XrYr, xy, flags, w, h = abi.load_calibration_arrays()
    assert (w, h) == (160, 120)
    assert len(flags) == 5763
    assert int((flags & 1).sum()) == 1207 and int(((flags & 2) > 0).sum()) == 731
    assert int(((flags & 1) & ((flags & 2) >> 1)).sum()) == 0          # gripperBody and gripperHole are disjoint
    assert xy[:, 1].min() == 64 and xy[:, 1].max() == 119 and xy[:, 0].min() == 8 and xy[:, 0].max() == 141
    cfg = abi.config_defaults()
    o = oracle.Oracle(cfg)
    assert (o.occ_w, o.occ_h) == (62, 37)                               # LocalMap 62 x 37


def test_mass_data_matches_survey():
    cfg = abi.config_defaults(n_arenas=1, n_robots=1, n_pucks=2)
    o = oracle.Oracle(cfg)
    o.reset([0])
    puck, robot = o.body_mass(0), o.body_mass(2)
    assert abs(puck[0] - 225.134) < 1e-2 and abs(puck[1] - 926.02) < 1e-1
    assert abs(robot[0] - 1066.375) < 1e-2 and abs(robot[1] - 46507.1) < 1.0
    assert abs(robot[2] - (-2.4431)) < 1e-3 and robot[3] == 0.0


def _free_robot():
    cfg = abi.config_defaults(n_arenas=1, n_robots=1, n_pucks=2, controller=abi.PS_CTRL_EXTERNAL, step_interval=1)
    o = oracle.Oracle(cfg)
    o.reset([0])
    st = o.get_state()
    # robot in the middle of the arena heading +x, pucks parked in a corner region away from it
    st.body[0, :, :] = 0
    st.body[0, 0, 0], st.body[0, 1, 0] = -60, -60
    st.body[0, 0, 1], st.body[0, 1, 1] = -60, -40
    st.arena[0]["n_contacts"] = 0
    o.set_state(st)
    return o


def test_known_answer_30cm_per_21_ticks():
    o = _free_robot()
    for _ in range(20):   # reach the steady speed first (the wall is 80 cm ahead)
        o.step_external(np.array([[750000.0]], np.float32), np.array([[0.0]], np.float32))
    x0 = o.get_state().body[0, 0, 2]
    for _ in range(21):
        o.step_external(np.array([[750000.0]], np.float32), np.array([[0.0]], np.float32))
    d = o.get_state().body[0, 0, 2] - x0
    assert 29.0 < d < 31.0, d          # "square with side length of approximately 30 cm"; derived value 30.14


def test_known_answer_180_degrees_per_30_ticks():
    o = _free_robot()
    for _ in range(60):
        o.step_external(np.array([[0.0]], np.float32), np.array([[2375000.0]], np.float32))
    a0 = o.get_state().body[0, 2, 2]
    for _ in range(30):
        o.step_external(np.array([[0.0]], np.float32), np.array([[2375000.0]], np.float32))
    da = o.get_state().body[0, 2, 2] - a0
    assert 3.0 < da < 3.25, da         # ONE_EIGHTY_TIME: ~pi; derived value 3.127


def test_puck_pair_contact_counts_and_symmetry():
    """Two overlapping pucks: 4 cached contacts (2 x 2 fixtures), only outer-outer touching; the position solver
    separates them symmetrically along x."""
    cfg = abi.config_defaults(n_arenas=1, n_robots=1, n_pucks=2, controller=abi.PS_CTRL_EXTERNAL, step_interval=1)
    o = oracle.Oracle(cfg)
    o.reset([0])
    st = o.get_state()
    st.body[0, :, :] = 0
    st.body[0, 0, 0], st.body[0, 0, 1] = -3.0, 3.0   # centres 6.0 apart < 6.3
    st.body[0, 0, 2], st.body[0, 1, 2] = 50, 50      # robot far away
    # fat boxes of the pucks around the new positions (robot proxies keep theirs)
    for p, x in ((0, -3.0), (1, 3.0)):
        for f, r in ((0, 3.15), (1, 2.1)):
            i = 2 * p + f
            st.fat_aabb[0, :, i] = [x - r - 0.1, -r - 0.1, x + r + 0.1, r + 0.1]
    keys = []
    for fa in (84, 85):
        for fb in (86, 87):
            keys.append((fa << 16) | fb)
    st.contact_key[0, :4] = sorted(keys)
    st.contact_info[0, :4] = 0
    st.contact_impulse[0, :4] = 0
    st.arena[0]["n_contacts"] = 4
    o.set_state(st)
    o.step_external(np.zeros((1, 1), np.float32), np.zeros((1, 1), np.float32))
    s2 = o.get_state()
    info = o.step_info(0)
    assert info["touching"] == 1 and info["points"] == 1
    assert s2.body[0, 0, 0] < -3.0 and s2.body[0, 0, 1] > 3.0
    assert abs(s2.body[0, 0, 0] + s2.body[0, 0, 1]) < 1e-6 and abs(s2.body[0, 1, 0]) < 1e-6


def test_in_free_space_regions_and_spawn_bounds():
    cfg = abi.config_defaults(n_arenas=3, n_robots=8, n_pucks=60)
    o = oracle.Oracle(cfg)
    o.reset([5, 6, 7])
    st = o.get_state()
    x, y = st.body[:, 0, :60], st.body[:, 1, :60]
    assert (np.abs(x) < 93.5 - 3.15).all() and (np.abs(y) < 93.5 - 3.15).all()
    # pucks are at least one puck width apart at spawn (Arena.java:143-161)
    for a in range(3):
        d = np.hypot(x[a][:, None] - x[a][None, :], y[a][:, None] - y[a][None, :]) + np.eye(60) * 100
        assert d.min() >= 6.3 - 1e-4
