"""ctypes binding of the host emulation harness (tests/hostemu/_build/libpsemu.so) -- tests only."""
import ctypes as C
import glob
import os
import subprocess
import numpy as np

from puckswarm_b200 import abi

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        d = os.path.join(_ROOT, "tests", "hostemu")
        path = os.path.join(d, "_build", "libpsemu.so")
        srcs = glob.glob(os.path.join(_ROOT, "puckswarm_b200", "csrc", "*.cuh")) + \
            glob.glob(os.path.join(_ROOT, "puckswarm_b200", "csrc", "*.hpp")) + [os.path.join(d, "emu.cpp")]
        if not os.path.exists(path) or any(os.path.getmtime(s) > os.path.getmtime(path) for s in srcs):
            subprocess.check_call(["make", "-s", "-C", d])
        L = C.CDLL(path)
        L.emu_create.restype = C.c_void_p
        L.emu_create.argtypes = [C.POINTER(abi.ps_config)]
        L.emu_destroy.argtypes = [C.c_void_p]
        L.emu_get_dims.argtypes = [C.c_void_p] + [C.POINTER(C.c_int32)] * 5
        L.emu_load_calibration.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.emu_reset.argtypes = [C.c_void_p, C.c_void_p]
        L.emu_step.argtypes = [C.c_void_p, C.c_int]
        L.emu_sense.argtypes = [C.c_void_p, C.POINTER(abi.ps_obs_view)]
        L.emu_step_info.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.emu_get_state.argtypes = [C.c_void_p, C.POINTER(abi.ps_state_view)]
        L.emu_set_state.argtypes = [C.c_void_p, C.POINTER(abi.ps_state_view)]
        L.emu_solve_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.emu_manifold.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int, C.c_float, C.c_float,
                                   C.c_float, C.c_void_p]
        L.emu_body_mass.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _LIB = L
    return _LIB


class Emu:
    def __init__(self, cfg, calib=None):
        self.L = lib()
        self.cfg = cfg
        self.h = self.L.emu_create(C.byref(cfg))
        if not self.h:
            raise RuntimeError("emu_create failed")
        XrYr, xy, flags, w, h = calib if calib is not None else abi.load_calibration_arrays()
        self.img_w, self.img_h = w, h
        rc = self.L.emu_load_calibration(self.h, XrYr.ctypes.data, xy.ctypes.data, flags.ctypes.data, len(flags), w, h)
        assert rc == 0
        d = [C.c_int32() for _ in range(5)]
        self.L.emu_get_dims(self.h, *[C.byref(x) for x in d])
        self.NB, self.NP, self.max_contacts, self.occ_w, self.occ_h = [x.value for x in d]

    def close(self):
        if self.h:
            self.L.emu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset(self, seeds):
        s = np.ascontiguousarray(seeds, np.int64)
        assert self.L.emu_reset(self.h, s.ctypes.data) == 0

    def step(self, n=1):
        rc = self.L.emu_step(self.h, n)
        assert rc == 0, rc

    def new_state(self):
        return abi.State(self.cfg.n_arenas, self.cfg.n_robots, self.NB, self.NP, self.max_contacts)

    def get_state(self, st=None):
        st = st or self.new_state()
        v = st.view()
        assert self.L.emu_get_state(self.h, C.byref(v)) == 0
        return st

    def set_state(self, st):
        v = st.view()
        assert self.L.emu_set_state(self.h, C.byref(v)) == 0

    def new_observations(self, camera=True):
        return abi.Observations(self.cfg.n_arenas, self.cfg.n_robots, self.img_w, self.img_h, self.occ_w, self.occ_h, camera)

    def sense(self, obs=None):
        obs = obs or self.new_observations()
        v = obs.view()
        assert self.L.emu_sense(self.h, C.byref(v)) == 0
        return obs

    def step_info(self, arena=0):
        out = np.zeros(8, np.int32)
        self.L.emu_step_info(self.h, arena, out.ctypes.data)
        return dict(pos_iters=int(out[0]), toi_events=int(out[1]), islands=int(out[2]))

    def solve_order(self, arena=0):
        out = np.zeros(65536, np.uint32)
        n = self.L.emu_solve_order(self.h, arena, out.ctypes.data, len(out))
        return out[:n].copy()

    def manifold(self, pa, xa, pb, xb):
        out = np.zeros(12, np.float32)
        self.L.emu_manifold(self.h, pa, xa[0], xa[1], xa[2], pb, xb[0], xb[1], xb[2], out.ctypes.data)
        return out

    def body_mass(self, robot):
        out = np.zeros(4, np.float32)
        self.L.emu_body_mass(self.h, robot, out.ctypes.data)
        return out
