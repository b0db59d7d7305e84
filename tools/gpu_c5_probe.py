"""BASELINE configs[4] shape (256 robots / 4000 pucks, enclosure scale 6) on the device: throughput of a small batch and the
capacity flags (diagnostic; run on the GPU box).  python tools/gpu_c5_probe.py [arenas] [think_steps]"""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from puckswarm_b200 import abi, engine
A = int(sys.argv[1]) if len(sys.argv) > 1 else 32
K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
cfg = abi.config_defaults(n_arenas=A, n_robots=256, n_pucks=4000, enclosure_scale=6.0, controller=abi.PS_CTRL_CACHECONS)
t = time.time(); e = engine.Engine(cfg); e.reset(np.arange(A)); e.get_state(); print('create + reset s', round(time.time() - t, 2))
for rep in range(3):
    t = time.time(); e.step(3 * K); st = e.get_state(); dt = time.time() - t
    print('rep', rep, 'robot-steps/s', round(A * 256 * 3 * K / dt), 'ms per think step', round(1000 * dt / K, 1), 'error flags', np.unique(st.arena['error_flags']), 'contacts/arena', st.arena['n_contacts'].mean())
