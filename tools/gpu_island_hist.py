"""Island statistics of the steady state (diagnostic; run on the GPU box): islands per arena-tick by position iterations x contacts."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from puckswarm_b200 import abi, engine
A = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 300
cfg = abi.config_defaults(n_arenas=A, n_robots=16, n_pucks=200)
e = engine.Engine(cfg)
e.reset(np.arange(A))
e.step(warm)
e.island_hist()
t = time.time(); e.step(30); e.get_state(); dt = time.time() - t
print('robot-steps/s', A * 16 * 30 / dt)
h = e.island_hist()
worst = int(h[7, 7]); h[7, 7] = 0; h = h.astype(np.float64)
print('slowest warp / big island: kcycles', worst >> 32, 'contacts', (worst >> 20) & 0xFFF, 'bodies', (worst >> 10) & 0x3FF, 'sweeps', worst & 0x3FF, 'fit_slot', not bool(worst & (1 << 31)), 'unfit islands', int(h[7, 6]))
print('islands (>=3 contacts) per arena-tick by [iterations x contacts]; rows iters <=1,<=3,<=6,<=12,<=25,<=50,<100,100; cols contacts <=3,5,8,12,16,32,64,>64')
np.set_printoptions(linewidth=200, suppress=True)
print(np.round(h / (A * 30.0), 3))
w = np.array([1, 2.5, 5, 9, 18, 37, 75, 100])[:, None] * np.array([3, 4.5, 7, 10, 14, 24, 48, 96])[None, :]
hw = h * w
print('estimated position steps per arena-tick', hw.sum() / (A * 30.0))
print('share by iters row', np.round(hw.sum(axis=1) / hw.sum(), 3))
print('share by contacts col', np.round(hw.sum(axis=0) / hw.sum(), 3))
print('share matrix'); print(np.round(hw / hw.sum(), 3))
