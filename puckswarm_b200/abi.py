"""ctypes mirror of include/puckswarm_b200.h (the C-ABI of the engine).

Only declarations live here: struct layouts, constants and numpy-backed views of the caller-allocated
arrays the ABI exchanges. Nothing in this module computes simulation results.
"""
import ctypes as C
import numpy as np

PS_ABI_VERSION = 3
PS_MAX_COLOURS = 8
PS_ROBOT_FIXTURES = 14
PS_PUCK_FIXTURES = 2
PS_WALL_FIXTURES = 84
PS_MAX_LM_PUCKS = 96
PS_N_SECTORS = 73

PS_T_NOTHING, PS_T_WALL, PS_T_ROBOT, PS_T_HIDDEN = 8, 9, 10, 11

PS_OK = 0
PS_E_INVALID, PS_E_NODEVICE, PS_E_CUDA, PS_E_NOCALIB, PS_E_CAPACITY, PS_E_NOMEM = -1, -2, -3, -4, -5, -6

PS_CTRL_CACHECONS, PS_CTRL_BHD, PS_CTRL_EXTERNAL, PS_CTRL_PROBSEEK = 0, 1, 2, 3
PS_RNG_JAVA_SHARED, PS_RNG_PER_ROBOT = 0, 1
PS_SEL_PROB_ABSOLUTE, PS_SEL_PROB_RELATIVE, PS_SEL_RELATIVE, PS_SEL_RELATIVE_FORGET = 0, 1, 2, 3
PS_SEP_IGNORE, PS_SEP_NULLIFY_OLD, PS_SEP_ACCEPT_ISOLATED, PS_SEP_ACCEPT_LARGER = 0, 1, 2, 3

CC_STATES = ["PU_SCAN", "PU_TARGET", "HOMING", "DE_PUSH", "DE_BACKUP", "EXILE", "SPIT_OUT"]
BHD_STATES = ["FORWARD", "PUSH", "BACKUP", "TURN"]
PROBSEEK_STATES = ["PU_SCAN", "PU_TARGET", "DE_SCAN", "DE_TARGET", "DE_PUSH", "DE_BACKUP", "DE_TURN"]


class ps_config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("n_arenas", C.c_int32), ("n_robots", C.c_int32), ("n_pucks", C.c_int32),
        ("n_puck_types", C.c_int32), ("n_informed_robots", C.c_int32), ("controller", C.c_int32),
        ("rng_mode", C.c_int32), ("step_interval", C.c_int32), ("vel_iters", C.c_int32), ("pos_iters", C.c_int32),
        ("dt", C.c_float), ("enclosure_scale", C.c_float), ("sincos_lut", C.c_int32), ("toi_enabled", C.c_int32),
        ("max_contacts", C.c_int32), ("cluster_distance_threshold", C.c_float), ("carry_threshold_lo", C.c_float),
        ("carry_threshold_hi", C.c_float), ("vfh_safety_gamma", C.c_float), ("vfh_safety_mask", C.c_float),
        ("vfh_tau_lo", C.c_float), ("vfh_tau_hi", C.c_float), ("cc_k1", C.c_float), ("cc_k2", C.c_float),
        ("cc_pick_up_time", C.c_int32), ("cc_push_in_time", C.c_int32), ("cc_backup_time", C.c_int32),
        ("cc_prediction_distance_sqd", C.c_float), ("cc_st_dev", C.c_float), ("cc_exile_time", C.c_int32),
        ("cc_spit_out_time", C.c_int32), ("cc_cache_selection", C.c_int32), ("cc_k_absolute", C.c_float),
        ("cc_k_relative", C.c_float), ("cc_forget_prob", C.c_float), ("cc_cache_separation", C.c_int32),
        ("cc_home_separation_distance", C.c_float), ("cc_ignore_pucks", C.c_int32),
        ("bhd_push_in_time", C.c_int32), ("bhd_backup_time", C.c_int32), ("ps_placement_time", C.c_int32),
        ("puck_shift_step", C.c_int32), ("puck_shuffle_step", C.c_int32),
    ]


def config_defaults(**kw):
    """ps_config with every field at the reference default; keyword overrides."""
    c = ps_config()
    c.abi_version = PS_ABI_VERSION
    c.n_arenas = 1
    c.n_robots = 4
    c.n_pucks = 10
    c.n_puck_types = 2
    c.n_informed_robots = 0
    c.controller = PS_CTRL_CACHECONS
    c.rng_mode = PS_RNG_JAVA_SHARED
    c.step_interval = 3
    c.vel_iters = 3
    c.pos_iters = 100
    c.dt = np.float32(1.0) / np.float32(14.0)
    c.enclosure_scale = 1.0
    c.sincos_lut = 1
    c.toi_enabled = 1
    c.max_contacts = 0
    c.cluster_distance_threshold = 1.5
    c.carry_threshold_lo = 0.1
    c.carry_threshold_hi = 0.4
    c.vfh_safety_gamma = 2.0
    c.vfh_safety_mask = 5.0
    c.vfh_tau_lo = 0.7
    c.vfh_tau_hi = 0.85
    c.cc_k1 = 1.0
    c.cc_k2 = 8.0
    c.cc_pick_up_time = 20
    c.cc_push_in_time = 1
    c.cc_backup_time = 5
    c.cc_prediction_distance_sqd = 100.0
    c.cc_st_dev = 0.5
    c.cc_exile_time = 20
    c.cc_spit_out_time = 10
    c.cc_cache_selection = PS_SEL_RELATIVE
    c.cc_k_absolute = 40.0
    c.cc_k_relative = 8.0
    c.cc_forget_prob = 0.001
    c.cc_cache_separation = PS_SEP_ACCEPT_LARGER
    c.cc_home_separation_distance = 50.0
    c.cc_ignore_pucks = 0
    c.bhd_push_in_time = 1
    c.bhd_backup_time = 5
    c.ps_placement_time = 40
    c.puck_shift_step = -1
    c.puck_shuffle_step = -1
    for k, v in kw.items():
        if not hasattr(c, k):
            raise AttributeError("ps_config has no field %r" % k)
        setattr(c, k, v)
    return c


class ps_robot_state(C.Structure):
    _fields_ = [
        ("cache_x", C.c_double * PS_MAX_COLOURS), ("cache_y", C.c_double * PS_MAX_COLOURS),
        ("rng_next_gaussian", C.c_double), ("rng_seed", C.c_uint64),
        ("remembered", C.c_float * PS_MAX_COLOURS), ("cache_valid", C.c_uint8 * PS_MAX_COLOURS),
        ("rng_have_gaussian", C.c_int32), ("ctrl_state", C.c_int32), ("start_count", C.c_int32),
        ("has_target", C.c_int32), ("target_type", C.c_int32), ("target_x", C.c_float), ("target_y", C.c_float),
        ("last_carried_type", C.c_int32), ("lm_carrying", C.c_int32), ("vfh_binary", C.c_uint32 * 3),
        ("vfh_last_sector", C.c_int32), ("vfh_last_angle", C.c_float), ("bhd_turn_dir", C.c_int32),
        ("bhd_turn_time", C.c_int32), ("state_counts", C.c_int32 * 8), ("cmd_forwards", C.c_float),
        ("cmd_torque", C.c_float), ("pad_", C.c_int32),
    ]


class ps_arena_state(C.Structure):
    _fields_ = [
        ("rng_next_gaussian", C.c_double), ("rng_seed", C.c_uint64), ("rng_have_gaussian", C.c_int32),
        ("update_count", C.c_int32), ("step_count", C.c_int32), ("n_contacts", C.c_int32),
        ("inv_dt0", C.c_float), ("error_flags", C.c_int32),
    ]


class ps_state_view(C.Structure):
    _fields_ = [
        ("body", C.POINTER(C.c_float)), ("sense_aabb", C.POINTER(C.c_float)), ("fat_aabb", C.POINTER(C.c_float)),
        ("contact_key", C.POINTER(C.c_uint32)), ("contact_info", C.POINTER(C.c_uint32)),
        ("contact_impulse", C.POINTER(C.c_float)), ("robot", C.POINTER(ps_robot_state)),
        ("arena", C.POINTER(ps_arena_state)),
    ]


class ps_localmap(C.Structure):
    _fields_ = [
        ("carrying", C.c_int32), ("carried_type", C.c_int32), ("carried_x", C.c_float), ("carried_y", C.c_float),
        ("carried_cluster", C.c_int32), ("n_pucks", C.c_int32), ("n_clusters", C.c_int32),
        ("left_sum", C.c_int32), ("right_sum", C.c_int32), ("overflow", C.c_int32),
        ("puck_x", C.c_float * PS_MAX_LM_PUCKS), ("puck_y", C.c_float * PS_MAX_LM_PUCKS),
        ("cluster_x", C.c_float * PS_MAX_LM_PUCKS), ("cluster_y", C.c_float * PS_MAX_LM_PUCKS),
        ("puck_type", C.c_uint8 * PS_MAX_LM_PUCKS), ("puck_cluster", C.c_uint8 * PS_MAX_LM_PUCKS),
        ("puck_degree", C.c_uint8 * PS_MAX_LM_PUCKS), ("cluster_type", C.c_uint8 * PS_MAX_LM_PUCKS),
        ("cluster_size", C.c_uint8 * PS_MAX_LM_PUCKS), ("cluster_kept", C.c_uint8 * PS_MAX_LM_PUCKS),
    ]


class ps_obs_view(C.Structure):
    _fields_ = [
        ("camera", C.POINTER(C.c_uint8)), ("occupancy", C.POINTER(C.c_uint8)), ("localmap", C.POINTER(ps_localmap)),
        ("pose", C.POINTER(C.c_float)),
    ]


class ps_stats(C.Structure):
    """Grouped by the reduction that combines ranks: int64 sums | double sums | int64 min | int64 max | double minima | double maxima."""
    _fields_ = [
        ("n_arenas", C.c_int64), ("cluster_size_hist", (C.c_int64 * 64) * PS_MAX_COLOURS),
        ("ctrl_state_hist", C.c_int64 * 8), ("n_carrying", C.c_int64), ("n_caches", C.c_int64),
        ("n_touching_points", C.c_int64), ("n_contacts", C.c_int64), ("n_completed", C.c_int64),
        ("sum_steps_to_completion", C.c_int64), ("n_snapshots", C.c_int64),
        ("sum_completion", C.c_double), ("sum_twc", C.c_double),
        ("min_steps_to_completion", C.c_int64), ("max_steps_to_completion", C.c_int64),
        ("min_completion", C.c_double), ("min_twc", C.c_double),
        ("max_completion", C.c_double), ("max_twc", C.c_double),
    ]


STATS_SEGMENTS = [   # (first field, last field, dtype, reduction): one collective each
    ("n_arenas", "n_snapshots", np.int64, "sum"), ("sum_completion", "sum_twc", np.float64, "sum"),
    ("min_steps_to_completion", "min_steps_to_completion", np.int64, "min"), ("max_steps_to_completion", "max_steps_to_completion", np.int64, "max"),
    ("min_completion", "min_twc", np.float64, "min"), ("max_completion", "max_twc", np.float64, "max"),
]


class ps_snapshot_view(C.Structure):
    _fields_ = [
        ("step_count", C.POINTER(C.c_int32)), ("robot_pose", C.POINTER(C.c_float)), ("puck_xy", C.POINTER(C.c_float)),
        ("cache_xy", C.POINTER(C.c_float)), ("state_counts", C.POINTER(C.c_int32)),
    ]


class ps_arena_analysis(C.Structure):
    _fields_ = [
        ("n_snapshots", C.c_int32), ("steps_to_completion", C.c_int32), ("twc_num", C.c_double), ("twc_den", C.c_double),
        ("completion", C.c_double),
    ]


ROBOT_DTYPE = np.dtype(ps_robot_state)
ARENA_DTYPE = np.dtype(ps_arena_state)
LOCALMAP_DTYPE = np.dtype(ps_localmap)
STATS_DTYPE = np.dtype(ps_stats)
ANALYSIS_DTYPE = np.dtype(ps_arena_analysis)


def _ptr(a, ctype):
    return a.ctypes.data_as(C.POINTER(ctype))


class State:
    """Caller-allocated host arrays of one ps_state_view (layouts documented in the header)."""

    def __init__(self, n_arenas, n_robots, NB, NP, max_contacts):
        self.dims = (n_arenas, n_robots, NB, NP, max_contacts)
        self.body = np.zeros((n_arenas, 6, NB), np.float32)
        self.sense_aabb = np.zeros((n_arenas, 4, NB), np.float32)
        self.fat_aabb = np.zeros((n_arenas, 4, NP), np.float32)
        self.contact_key = np.zeros((n_arenas, max_contacts), np.uint32)
        self.contact_info = np.zeros((n_arenas, max_contacts), np.uint32)
        self.contact_impulse = np.zeros((n_arenas, max_contacts, 4), np.float32)
        self.robot = np.zeros((n_arenas, n_robots), ROBOT_DTYPE)
        self.arena = np.zeros((n_arenas,), ARENA_DTYPE)

    def view(self):
        v = ps_state_view()
        v.body = _ptr(self.body, C.c_float)
        v.sense_aabb = _ptr(self.sense_aabb, C.c_float)
        v.fat_aabb = _ptr(self.fat_aabb, C.c_float)
        v.contact_key = _ptr(self.contact_key, C.c_uint32)
        v.contact_info = _ptr(self.contact_info, C.c_uint32)
        v.contact_impulse = _ptr(self.contact_impulse, C.c_float)
        v.robot = _ptr(self.robot, ps_robot_state)
        v.arena = _ptr(self.arena, ps_arena_state)
        return v

    def contacts(self, a):
        """(key, info, impulse[4]) of the live cached contacts of arena a, creation order."""
        n = int(self.arena[a]["n_contacts"])
        return self.contact_key[a, :n], self.contact_info[a, :n], self.contact_impulse[a, :n]


class Snapshot:
    """Caller-allocated host arrays of one ps_snapshot_view (the reference's periodic record of the latest storage step)."""

    def __init__(self, n_arenas, n_robots, n_pucks_spawned):
        self.step_count = np.full((n_arenas,), -1, np.int32)
        self.robot_pose = np.zeros((n_arenas, n_robots, 3), np.float32)
        self.puck_xy = np.zeros((n_arenas, max(1, n_pucks_spawned), 2), np.float32)[:, :n_pucks_spawned]
        self.cache_xy = np.zeros((n_arenas, n_robots, PS_MAX_COLOURS, 2), np.float32)
        self.state_counts = np.zeros((n_arenas, n_robots, 8), np.int32)

    def view(self):
        v = ps_snapshot_view()
        v.step_count = _ptr(self.step_count, C.c_int32)
        v.robot_pose = _ptr(self.robot_pose, C.c_float)
        if self.puck_xy.size:
            v.puck_xy = _ptr(self.puck_xy, C.c_float)
        v.cache_xy = _ptr(self.cache_xy, C.c_float)
        v.state_counts = _ptr(self.state_counts, C.c_int32)
        return v


class Observations:
    """Caller-allocated host arrays of one ps_obs_view."""

    def __init__(self, n_arenas, n_robots, img_w, img_h, occ_w, occ_h, camera=True):
        self.camera = np.zeros((n_arenas, n_robots, img_h, img_w), np.uint8) if camera else None
        self.occupancy = np.zeros((n_arenas, n_robots, occ_w, occ_h), np.uint8)
        self.localmap = np.zeros((n_arenas, n_robots), LOCALMAP_DTYPE)
        self.pose = np.zeros((n_arenas, n_robots, 3), np.float32)

    def view(self):
        v = ps_obs_view()
        if self.camera is not None:
            v.camera = _ptr(self.camera, C.c_uint8)
        v.occupancy = _ptr(self.occupancy, C.c_uint8)
        v.localmap = _ptr(self.localmap, ps_localmap)
        v.pose = _ptr(self.pose, C.c_float)
        return v


def load_calibration_arrays(path=None):
    """The packed grid-camera calibration (tools/make_calibration.py): XrYr f32[n,2], xy u16[n,2], flags u8[n], (w, h)."""
    import os
    if path is None:
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "calib_114_160x120.npz")
    z = np.load(path)
    return (np.ascontiguousarray(z["XrYr"], np.float32), np.ascontiguousarray(z["xy"], np.uint16),
            np.ascontiguousarray(z["flags"], np.uint8), int(z["size"][0]), int(z["size"][1]))


def write_calibration_bin(path, calib=None):
    """The packed calibration as the flat file the C++ host reads (puckswarm_b200/host/puckswarm_host.hpp: Calibration::load):
    int32 n, w, h; float32 XrYr[n][2]; uint16 xy[n][2]; uint8 flags[n]."""
    XrYr, xy, flags, w, h = calib if calib is not None else load_calibration_arrays()
    with open(path, "wb") as f:
        f.write(np.array([len(flags), w, h], np.int32).tobytes())
        f.write(np.ascontiguousarray(XrYr, np.float32).tobytes())
        f.write(np.ascontiguousarray(xy, np.uint16).tobytes())
        f.write(np.ascontiguousarray(flags, np.uint8).tobytes())
    return path
