"""Phase breakdown of k_physics from its in-kernel cycle counters (diagnostic; run on the GPU box)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from puckswarm_b200 import abi, engine
A = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = abi.config_defaults(n_arenas=A, n_robots=16, n_pucks=200)
e = engine.Engine(cfg)
e.reset(np.arange(A))
e.step(45)
e.island_hist()
e.profile(True); e.kernel_ms()
t = time.time(); e.step(30); dt = time.time() - t
kms, kc = e.kernel_ms()
print('robot-steps/s', A * 16 * 30 / dt, 'kernel ms', kms, kc)
st = e.step_stats().astype(np.float64)
names = ['pos_iters', 'toi', 'islands', 'toi_rounds', 'toi_evals', 'syncfix', 'findnew', 'toi_cycles']
for i, n in enumerate(names):
    scale = 16 if i >= 5 else 1
    print('%-12s mean %12.1f  max %12.1f' % (n, st[:, i].mean() * scale, st[:, i].max() * scale))
tot = st[:, 3:8].sum(axis=1) * 16
print('cycles per arena-tick mean', tot.mean(), 'max', tot.max())

w = np.argsort(-st[:, 7])[:8]
print('worst TOI arenas: idx, rounds, evals, toi_events, Mcycles')
for a in w:
    print(int(a), int(st[a, 3]), int(st[a, 4]), int(st[a, 1]), st[a, 7] * 16 / 1e6)

ss = e.sense_stats().reshape(-1, 8).astype(np.float64)
for i, n in enumerate(['candidates', 'camera', 'occupancy | blobs', '(unused)', 'bodies examined', 'coloured pixels', 'candidates', 'blobs']):
    sc = 16 if i < 4 else 1
    print('sense %-15s mean %10.1f max %10.1f' % (n, ss[:, i].mean() * sc, ss[:, i].max() * sc))

h = e.island_hist()
worst = int(h[7, 7]); h[7, 7] = 0
print('slowest oversized island: kcycles', worst >> 32, 'contacts', (worst >> 20) & 0x7FF, 'bodies', (worst >> 10) & 0x3FF, 'sweeps', worst & 0x3FF, 'fit_slot', not bool(worst & (1 << 31)))
print('islands (>=3 contacts) per arena-tick by [iterations x contacts]; rows iters <=1,<=3,<=6,<=12,<=25,<=50,<100,100; cols contacts <=3,5,8,12,16,32,64,>64')
print(np.round(h / (A * 30.0), 3))
w = np.array([1, 2.5, 5, 9, 18, 37, 75, 100])[:, None] * np.array([3, 4.5, 7, 10, 14, 24, 48, 96])[None, :]
print('estimated position steps per arena-tick', (h * w).sum() / (A * 30.0), 'share by iters row', np.round((h * w).sum(axis=1) / (h * w).sum(), 3))
