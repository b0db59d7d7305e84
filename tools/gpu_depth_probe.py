"""Does the queue depth matter? step(3) with at most D steps in flight per engine (events), 1 engine x 4 groups."""
import os, sys, time
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
sys.path.insert(0, '.')
import numpy as np
import torch
from puckswarm_b200 import abi, engine
A = 4096
D = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = abi.config_defaults(n_arenas=A, n_robots=16, n_pucks=200)
e = engine.Engine(cfg); e.reset(np.arange(A)); e.step(75)
s = torch.cuda.ExternalStream(e.stream)
N = 24
evs = []
t0 = time.perf_counter()
for k in range(N):
    e.step(3, sync=False)
    e.join()
    ev = torch.cuda.Event(); ev.record(s); evs.append(ev)
    if len(evs) > D - 1:
        evs[len(evs) - D].synchronize()
e.sync()
dt = time.perf_counter() - t0
print("depth %d: %.2f ms per step, %.2f M robot-steps/s" % (D, dt / N * 1e3, A * 48 * N / dt / 1e6))
