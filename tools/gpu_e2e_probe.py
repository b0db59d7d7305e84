"""Where the end-to-end step goes: time in ps_get_state / ps_set_state / ps_step per handle, round-robin over N handles."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, '.')
import ctypes as C
import numpy as np
import torch
from puckswarm_b200 import abi, engine
A, NH = 4096, int(sys.argv[1]) if len(sys.argv) > 1 else 4
os.environ["PS_GROUPS"] = sys.argv[2] if len(sys.argv) > 2 else "1"
hs = []
for hi in range(NH):
    a0, a1 = hi * A // NH, (hi + 1) * A // NH
    cfg = abi.config_defaults(n_arenas=a1 - a0, n_robots=16, n_pucks=200)
    e = engine.Engine(cfg); e.reset(np.arange(a0, a1)); e.step(75, sync=False)
    st = e.new_state()
    pb = torch.empty(st.body.shape, dtype=torch.float32).pin_memory(); pa = torch.empty(st.arena.nbytes, dtype=torch.uint8).pin_memory()
    vin, vout = abi.ps_state_view(), abi.ps_state_view()
    vin.body = pb.numpy().ctypes.data_as(C.POINTER(C.c_float)); vout.body = vin.body
    vout.arena = pa.numpy().view(abi.ARENA_DTYPE).ctypes.data_as(C.POINTER(abi.ps_arena_state))
    e._check(e.L.ps_get_state(e.h, C.byref(vout)))
    hs.append((e, vin, vout, pb, pa))
MODE = sys.argv[4] if len(sys.argv) > 4 else "full"
STAGGER = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
T = {"get": 0.0, "set": 0.0, "step": 0.0}
def rnd(first, acc):
    for e, vin, vout, _, _ in hs:
        t0 = time.perf_counter()
        if not first:
            if MODE == "synconly": e.sync()
            else: e._check(e.L.ps_get_state(e.h, C.byref(vout)))
        t1 = time.perf_counter()
        if MODE != "noset" and MODE != "synconly": e._check(e.L.ps_set_state(e.h, C.byref(vin)))
        t2 = time.perf_counter(); e._check(e.L.ps_step(e.h, 3))
        t3 = time.perf_counter()
        if first and STAGGER > 0: time.sleep(STAGGER)
        if acc: T["get"] += t1 - t0; T["set"] += t2 - t1; T["step"] += t3 - t2
rnd(True, False)
for _ in range(4): rnd(False, False)
N = 20
t0 = time.perf_counter()
for _ in range(N): rnd(False, True)
for e, vin, vout, _, _ in hs: e._check(e.L.ps_get_state(e.h, C.byref(vout)))
dt = time.perf_counter() - t0
print("handles %d groups/handle %s: %.2f M robot-steps/s, %.2f ms per step; per step: get %.2f ms set %.2f ms step(launch) %.2f ms" % (
    NH, os.environ["PS_GROUPS"], A * 16 * 3 * N / dt / 1e6, dt / N * 1e3, T["get"] / N * 1e3, T["set"] / N * 1e3, T["step"] / N * 1e3))
# free-running reference: every handle gets all N steps at once
t0 = time.perf_counter()
for _ in range(N):
    for e, vin, vout, _, _ in hs: e._check(e.L.ps_step(e.h, 3))
for e, vin, vout, _, _ in hs: e.sync()
dt = time.perf_counter() - t0
print("  same handles free-running: %.2f ms per step" % (dt / N * 1e3))
