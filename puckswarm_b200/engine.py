"""ctypes binding of the engine's C-ABI (include/puckswarm_b200.h -> csrc/libpuckswarm_b200.so).

The library holds CUDA kernels for sm_100a only. There is no CPU path: if the library is missing or no
B200 is visible, construction fails loudly (EngineUnavailable / EngineError) instead of falling back.
"""
import ctypes as C
import os
import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PS_ENGINE_LIB") or os.path.join(_HERE, "csrc", "libpuckswarm_b200.so")   # PS_ENGINE_LIB: diagnostic builds

EXPORTS = [
    "ps_last_error", "ps_abi_version", "ps_struct_size", "ps_config_defaults", "ps_create", "ps_destroy", "ps_load_calibration",
    "ps_reset", "ps_set_state", "ps_get_state", "ps_step", "ps_sync", "ps_get_error_flags", "ps_join", "ps_sense", "ps_step_external",
    "ps_get_observations", "ps_compute_stats", "ps_stats_device_ptr", "ps_get_stats", "ps_get_snapshot", "ps_get_analysis", "ps_analysis_config", "ps_get_dims", "ps_get_step_stats", "ps_get_sense_stats", "ps_get_island_hist", "ps_get_island_census",
    "ps_profile", "ps_get_kernel_ms", "ps_num_groups", "ps_kernel_launches", "ps_stream",
]


class EngineUnavailable(RuntimeError):
    pass


class EngineError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("puckswarm_b200 error %d: %s" % (code, msg))
        self.code = code


_LIB = None


def load_library(path=None):
    """dlopen the engine; raises EngineUnavailable if it has not been built (python __graft_entry__.py build)."""
    global _LIB
    if _LIB is not None and path is None:
        return _LIB
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise EngineUnavailable("%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                                "there is no CPU fallback" % p)
    L = C.CDLL(p)
    L.ps_last_error.restype = C.c_char_p
    L.ps_abi_version.restype = C.c_int
    L.ps_struct_size.argtypes = [C.c_int]
    structs = [abi.ps_config, abi.ps_robot_state, abi.ps_arena_state, abi.ps_localmap, abi.ps_stats, abi.ps_state_view, abi.ps_obs_view,
               abi.ps_snapshot_view, abi.ps_arena_analysis]
    for i, t in enumerate(structs):
        if L.ps_struct_size(i) != C.sizeof(t):
            raise EngineUnavailable("ABI layout mismatch for %s: library %d bytes, abi.py %d bytes" % (t.__name__, L.ps_struct_size(i), C.sizeof(t)))
    L.ps_config_defaults.argtypes = [C.POINTER(abi.ps_config)]
    L.ps_create.argtypes = [C.POINTER(abi.ps_config), C.c_int, C.POINTER(C.c_void_p)]
    L.ps_destroy.argtypes = [C.c_void_p]
    L.ps_load_calibration.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.ps_reset.argtypes = [C.c_void_p, C.c_void_p]
    L.ps_set_state.argtypes = [C.c_void_p, C.POINTER(abi.ps_state_view)]
    L.ps_get_state.argtypes = [C.c_void_p, C.POINTER(abi.ps_state_view)]
    L.ps_step.argtypes = [C.c_void_p, C.c_int]
    L.ps_sync.argtypes = [C.c_void_p]
    L.ps_join.argtypes = [C.c_void_p]
    L.ps_get_error_flags.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.ps_sense.argtypes = [C.c_void_p]
    L.ps_step_external.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ps_get_observations.argtypes = [C.c_void_p, C.POINTER(abi.ps_obs_view)]
    L.ps_compute_stats.argtypes = [C.c_void_p, C.c_float, C.POINTER(abi.ps_stats)]
    L.ps_stats_device_ptr.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
    L.ps_get_stats.argtypes = [C.c_void_p, C.c_float, C.POINTER(abi.ps_stats), C.c_void_p]
    L.ps_get_snapshot.argtypes = [C.c_void_p, C.POINTER(abi.ps_snapshot_view)]
    L.ps_get_analysis.argtypes = [C.c_void_p, C.c_void_p]
    L.ps_analysis_config.argtypes = [C.c_void_p, C.c_float, C.c_float]
    L.ps_get_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 5
    L.ps_get_step_stats.argtypes = [C.c_void_p, C.c_void_p]
    L.ps_get_sense_stats.argtypes = [C.c_void_p, C.c_void_p]
    L.ps_get_island_hist.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.ps_get_island_census.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    L.ps_profile.argtypes = [C.c_void_p, C.c_int]
    L.ps_get_kernel_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.ps_num_groups.argtypes = [C.c_void_p]
    L.ps_kernel_launches.restype = C.c_int64
    L.ps_kernel_launches.argtypes = [C.c_void_p]
    L.ps_stream.restype = C.c_void_p
    L.ps_stream.argtypes = [C.c_void_p]
    if path is None:
        _LIB = L
    return L


class Engine:
    """One batch of arenas on one GPU. Thin, 1:1 over the C-ABI; host arrays in and out."""

    def __init__(self, cfg, device=0, calib=None, load_calibration=True):
        self.L = load_library()
        if self.L.ps_abi_version() != abi.PS_ABI_VERSION:
            raise EngineError(abi.PS_E_INVALID, "ABI version mismatch between abi.py and the library")
        self.cfg = cfg
        self.device = device
        h = C.c_void_p()
        self._check(self.L.ps_create(C.byref(cfg), device, C.byref(h)))
        self.h = h
        self.img_w = self.img_h = 0
        if load_calibration:
            XrYr, xy, flags, w, hh = calib if calib is not None else abi.load_calibration_arrays()
            self.img_w, self.img_h = w, hh
            self._check(self.L.ps_load_calibration(self.h, XrYr.ctypes.data, xy.ctypes.data, flags.ctypes.data, len(flags), w, hh))
        d = [C.c_int32() for _ in range(5)]
        self._check(self.L.ps_get_dims(self.h, *[C.byref(x) for x in d]))
        self.NB, self.NP, self.max_contacts, self.occ_w, self.occ_h = [x.value for x in d]
        self.analysis_threshold = float(cfg.cluster_distance_threshold)   # Analyze clusters with LocalMap's threshold

    def _check(self, rc):
        if rc != abi.PS_OK:
            raise EngineError(rc, (self.L.ps_last_error() or b"").decode("utf-8", "replace"))

    def close(self):
        if getattr(self, "h", None):
            self.L.ps_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- simulation ----
    def reset(self, seeds):
        s = np.ascontiguousarray(seeds, np.int64)
        if len(s) != self.cfg.n_arenas:
            raise ValueError("need one seed per arena")
        self._check(self.L.ps_reset(self.h, s.ctypes.data))

    def step(self, n_ticks=1, sync=True):
        self._check(self.L.ps_step(self.h, n_ticks))
        if sync:
            self.sync()

    def sync(self):
        self._check(self.L.ps_sync(self.h))

    def error_flags(self):
        """(OR of the arenas' capacity flags, number of flagged arenas); ps_sync raises EngineError(PS_E_CAPACITY) on the same condition."""
        f, n = C.c_int32(), C.c_int32()
        self._check(self.L.ps_get_error_flags(self.h, C.byref(f), C.byref(n)))
        return f.value, n.value

    def join(self):
        """Order the main stream after all queued steps (no host wait): events recorded on `stream` then cover them."""
        self._check(self.L.ps_join(self.h))

    def sense(self):
        self._check(self.L.ps_sense(self.h))
        self.sync()

    def step_external(self, fwd, torque, sync=True):
        f = np.ascontiguousarray(fwd, np.float32)
        t = np.ascontiguousarray(torque, np.float32)
        n = self.cfg.n_arenas * self.cfg.n_robots
        if f.size != n or t.size != n:
            raise ValueError("commands must be [n_arenas][n_robots]")
        self._check(self.L.ps_step_external(self.h, f.ctypes.data, t.ctypes.data))
        if sync:
            self.sync()

    # ---- state / observations ----
    def new_state(self):
        return abi.State(self.cfg.n_arenas, self.cfg.n_robots, self.NB, self.NP, self.max_contacts)

    def get_state(self, st=None):
        st = st or self.new_state()
        v = st.view()
        self._check(self.L.ps_get_state(self.h, C.byref(v)))
        return st

    def set_state(self, st):
        v = st.view()
        self._check(self.L.ps_set_state(self.h, C.byref(v)))

    def new_observations(self, camera=True):
        return abi.Observations(self.cfg.n_arenas, self.cfg.n_robots, self.img_w, self.img_h, self.occ_w, self.occ_h, camera)

    def get_observations(self, obs=None, camera=True):
        obs = obs or self.new_observations(camera)
        v = obs.view()
        self._check(self.L.ps_get_observations(self.h, C.byref(v)))
        return obs

    def compute_stats(self, threshold=None):
        s = abi.ps_stats()
        thr = self.cfg.cluster_distance_threshold if threshold is None else threshold
        self._check(self.L.ps_compute_stats(self.h, thr, C.byref(s)))
        return s

    def get_stats(self, threshold=None, nccl_comm=None):
        """ps_stats of this handle's arenas, reduced over the ranks of `nccl_comm` (an ncclComm_t as an integer / c_void_p) on the device."""
        s = abi.ps_stats()
        thr = self.cfg.cluster_distance_threshold if threshold is None else threshold
        self._check(self.L.ps_get_stats(self.h, thr, C.byref(s), C.c_void_p(nccl_comm) if nccl_comm else None))
        return s

    def get_snapshot(self, snap=None):
        """The reference's periodic record (Arena.java:223-243 and the controllers' dumps) of the latest storage step."""
        snap = snap or abi.Snapshot(self.cfg.n_arenas, self.cfg.n_robots, self.NB - self.cfg.n_robots)
        v = snap.view()
        self._check(self.L.ps_get_snapshot(self.h, C.byref(v)))
        return snap

    def get_analysis(self):
        """Per-arena Analyze accumulators (numpy structured array of ps_arena_analysis)."""
        out = np.zeros((self.cfg.n_arenas,), abi.ANALYSIS_DTYPE)
        self._check(self.L.ps_get_analysis(self.h, out.ctypes.data))
        return out

    def analysis_config(self, cluster_threshold, target_percent=50.0):
        self._check(self.L.ps_analysis_config(self.h, cluster_threshold, target_percent))
        self.analysis_threshold = float(cluster_threshold)

    def step_info(self, arena=0):
        out = np.zeros((self.cfg.n_arenas, 8), np.int32)
        self._check(self.L.ps_get_step_stats(self.h, out.ctypes.data))
        return dict(pos_iters=int(out[arena, 0]), toi_events=int(out[arena, 1]), islands=int(out[arena, 2]))

    def step_stats(self):
        out = np.zeros((self.cfg.n_arenas, 8), np.int32)
        self._check(self.L.ps_get_step_stats(self.h, out.ctypes.data))
        return out

    def sense_stats(self):
        out = np.zeros((self.cfg.n_arenas, self.cfg.n_robots, 8), np.int32)
        self._check(self.L.ps_get_sense_stats(self.h, out.ctypes.data))
        return out

    def island_hist(self, reset=True):
        out = np.zeros((8, 8), np.int64)
        self._check(self.L.ps_get_island_hist(self.h, out.ctypes.data, 1 if reset else 0))
        return out

    def island_census(self, reset=True):
        """Stage-B census since the last reset (include/puckswarm_b200.h: ps_get_island_census)."""
        out = np.zeros(128, np.int64)
        self._check(self.L.ps_get_island_census(self.h, out.ctypes.data, 1 if reset else 0))
        return out

    def profile(self, enable=True):
        self._check(self.L.ps_profile(self.h, 1 if enable else 0))

    def kernel_ms(self):
        """({'sense': ms, 'think': ms, 'physics': ms, 'other': ms}, counts) accumulated since the last call."""
        ms = np.zeros(12, np.float64)
        cnt = np.zeros(12, np.int64)
        self._check(self.L.ps_get_kernel_ms(self.h, ms.ctypes.data, cnt.ctypes.data))
        names = ("sense", "think", "physics", "other", "phys_pre", "phys_islands", "phys_post", "unused", "isl_big", "isl_warp", "isl_batch", "isl_lanex")
        return dict(zip(names, ms.tolist())), dict(zip(names, cnt.tolist()))

    @property
    def num_groups(self):
        return int(self.L.ps_num_groups(self.h))

    @property
    def kernel_launches(self):
        return int(self.L.ps_kernel_launches(self.h))

    @property
    def stream(self):
        return int(self.L.ps_stream(self.h) or 0)
