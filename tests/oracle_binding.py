"""ctypes binding of the CPU oracle (oracle/_build/liboracle.so) -- test infrastructure only.

Same method names and array layouts as puckswarm_b200.engine.Engine so a parity test drives both alike.
"""
import ctypes as C
import os
import subprocess
import numpy as np

from puckswarm_b200 import abi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])
    return os.path.join(_ROOT, "oracle", "_build", "liboracle.so")


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_ROOT, "oracle", "_build", "liboracle.so")
        srcs = [os.path.join(_ROOT, "oracle", f) for f in ("oracle_api.cpp", "jb2d.h", "puckswarm.h", "jrandom.h")]
        srcs.append(os.path.join(_ROOT, "include", "puckswarm_b200.h"))
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs if os.path.exists(s)):
            if all(os.path.exists(s) for s in srcs):
                build()
        L = C.CDLL(path)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(abi.ps_config)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_load_calibration.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.orc_get_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 5
        L.orc_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_step.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.orc_sense.argtypes = [C.c_void_p, C.c_int]
        L.orc_step_external.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_get_state.argtypes = [C.c_void_p, C.POINTER(abi.ps_state_view)]
        L.orc_set_state.argtypes = [C.c_void_p, C.POINTER(abi.ps_state_view)]
        L.orc_get_observations.argtypes = [C.c_void_p, C.POINTER(abi.ps_obs_view)]
        L.orc_compute_stats.argtypes = [C.c_void_p, C.c_float, C.POINTER(abi.ps_stats)]
        L.orc_step_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.orc_get_snapshot.argtypes = [C.c_void_p, C.POINTER(abi.ps_snapshot_view)]
        L.orc_get_analysis.argtypes = [C.c_void_p, C.c_void_p]
        L.orc_analysis_config.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.orc_solve_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.orc_manifold.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float,
                                   C.c_float, C.c_void_p]
        L.orc_random_ints.argtypes = [C.c_int64, C.c_int, C.c_void_p]
        L.orc_random_floats.argtypes = [C.c_int64, C.c_int, C.c_void_p]
        L.orc_random_gaussians.argtypes = [C.c_int64, C.c_int, C.c_void_p]
        L.orc_random_bounded.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_void_p]
        L.orc_sin_lut.restype = C.c_float
        L.orc_sin_lut.argtypes = [C.c_float]
        L.orc_body_mass.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


class Oracle:
    def __init__(self, cfg, calib=None, threads=1):
        self.L = lib()
        self.cfg = cfg
        self.threads = threads
        self.h = self.L.orc_create(C.byref(cfg))
        if not self.h:
            raise RuntimeError("orc_create failed")
        XrYr, xy, flags, w, h = calib if calib is not None else abi.load_calibration_arrays()
        self.img_w, self.img_h = w, h
        self._calib = (XrYr, xy, flags)
        self.L.orc_load_calibration(self.h, XrYr.ctypes.data, xy.ctypes.data, flags.ctypes.data, len(flags), w, h)
        d = [C.c_int32() for _ in range(5)]
        self.L.orc_get_dims(self.h, *[C.byref(x) for x in d])
        self.NB, self.NP, self.max_contacts, self.occ_w, self.occ_h = [x.value for x in d]

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, seeds):
        s = np.ascontiguousarray(seeds, np.int64)
        assert len(s) == self.cfg.n_arenas
        rc = self.L.orc_reset(self.h, s.ctypes.data)
        assert rc == 0, rc

    def step(self, n_ticks=1):
        assert self.L.orc_step(self.h, n_ticks, self.threads) == 0

    def sense(self):
        assert self.L.orc_sense(self.h, self.threads) == 0

    def step_external(self, fwd, torque):
        f = np.ascontiguousarray(fwd, np.float32)
        t = np.ascontiguousarray(torque, np.float32)
        assert self.L.orc_step_external(self.h, f.ctypes.data, t.ctypes.data, self.threads) == 0

    def new_state(self):
        return abi.State(self.cfg.n_arenas, self.cfg.n_robots, self.NB, self.NP, self.max_contacts)

    def get_state(self, st=None):
        st = st or self.new_state()
        v = st.view()
        rc = self.L.orc_get_state(self.h, C.byref(v))
        assert rc == 0, rc
        return st

    def set_state(self, st):
        v = st.view()
        assert self.L.orc_set_state(self.h, C.byref(v)) == 0

    def new_observations(self, camera=True):
        return abi.Observations(self.cfg.n_arenas, self.cfg.n_robots, self.img_w, self.img_h, self.occ_w, self.occ_h, camera)

    def get_observations(self, obs=None, camera=True):
        obs = obs or self.new_observations(camera)
        v = obs.view()
        assert self.L.orc_get_observations(self.h, C.byref(v)) == 0
        return obs

    def compute_stats(self, threshold=None):
        s = abi.ps_stats()
        thr = self.cfg.cluster_distance_threshold if threshold is None else threshold
        assert self.L.orc_compute_stats(self.h, thr, C.byref(s)) == 0
        return s

    def get_snapshot(self, snap=None):
        snap = snap or abi.Snapshot(self.cfg.n_arenas, self.cfg.n_robots, self.NB - self.cfg.n_robots)
        v = snap.view()
        assert self.L.orc_get_snapshot(self.h, C.byref(v)) == 0
        return snap

    def get_analysis(self):
        out = np.zeros((self.cfg.n_arenas,), abi.ANALYSIS_DTYPE)
        assert self.L.orc_get_analysis(self.h, out.ctypes.data) == 0
        return out

    def analysis_config(self, cluster_threshold, target_percent=50.0):
        assert self.L.orc_analysis_config(self.h, cluster_threshold, target_percent) == 0

    def step_info(self, arena=0):
        out = np.zeros(8, np.int32)
        self.L.orc_step_info(self.h, arena, out.ctypes.data)
        return dict(pos_iters=int(out[0]), toi_events=int(out[1]), islands=int(out[2]), contacts=int(out[3]),
                    touching=int(out[4]), points=int(out[5]))

    def solve_order(self, arena=0):
        out = np.zeros(65536, np.uint32)
        n = self.L.orc_solve_order(self.h, arena, out.ctypes.data, len(out))
        return out[:n].copy()

    def manifold(self, pa, xa, pb, xb):
        out = np.zeros(12, np.float32)
        self.L.orc_manifold(self.h, pa, xa[0], xa[1], xa[2], pb, xb[0], xb[1], xb[2], out.ctypes.data)
        return out

    def body_mass(self, i):
        out = np.zeros(4, np.float32)
        self.L.orc_body_mass(self.h, i, out.ctypes.data)
        return out


def random_ints(seed, n):
    out = np.zeros(n, np.int32)
    lib().orc_random_ints(seed, n, out.ctypes.data)
    return out


def random_floats(seed, n):
    out = np.zeros(n, np.float32)
    lib().orc_random_floats(seed, n, out.ctypes.data)
    return out


def random_gaussians(seed, n):
    out = np.zeros(n, np.float64)
    lib().orc_random_gaussians(seed, n, out.ctypes.data)
    return out


def random_bounded(seed, bound, n):
    out = np.zeros(n, np.int32)
    lib().orc_random_bounded(seed, bound, n, out.ctypes.data)
    return out
