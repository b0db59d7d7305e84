"""Stage-B census of the steady state (diagnostic; run on the GPU box): islands per queue bucket, duration histograms of the
warp-slot and big islands, warp rounds and lane occupancy of the lane batches."""
import sys
sys.path.insert(0, '.')
import numpy as np
from puckswarm_b200 import abi, engine
A = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 300
ticks = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cfg = abi.config_defaults(n_arenas=A, n_robots=16, n_pucks=200)
e = engine.Engine(cfg)
e.reset(np.arange(A))
e.step(warm)
e.island_census()
e.step(ticks)
c = e.island_census().astype(np.float64)
T = max(1.0, c[10])
names = ['warp(0)', 'lane1 (1-2)', 'lane2 (3-5)', 'lane3 (6-12)', 'lane4 (13-16)', 'big', 'large', 'XL', 'XXL', 'laneX (17-27)']
print('arenas', A, 'ticks', int(T), 'after', warm)
print('islands per tick by bucket:', {n: round(c[i] / T, 1) for i, n in enumerate(names)})
print('per arena-tick:', {n: round(c[i] / T / A, 3) for i, n in enumerate(names)})
np.set_printoptions(linewidth=200, suppress=True)
print('warp-slot islands per tick by duration (bins of 262k cycles = 0.133 ms):', np.round(c[16:32] / T, 1))
print('big islands per tick by duration:', np.round(c[32:48] / T, 1))
print('warp-slot islands: mean kcycles', c[48] / max(1, c[49]), 'total Gcycles per tick', c[48] * 1024 / T / 1e9, '= warp-ms per tick', c[48] * 1024 / T / 1.965e6)
print('big islands: mean kcycles', c[50] / max(1, c[51]), 'total Gcycles per tick', c[50] * 1024 / T / 1e9, '= warp-ms per tick', c[50] * 1024 / T / 1.965e6)
for b in (0, 1, 2, 3, 5):
    print('lane bucket', b + 1 if b < 4 else 'X', ': warp rounds per tick', c[53 + 2 * b] / T, 'mean kcycles per round', c[52 + 2 * b] / max(1, c[53 + 2 * b]), 'warp-ms per tick', c[52 + 2 * b] * 1024 / T / 1.965e6)
print('lane occupancy of the lock-step position loop (sum of lane sweeps / 32 x warp sweeps):', c[60] / max(1, c[61]))
print('k_phys_pre cycles per arena-tick: load %.0f collide %.0f graph+islands %.0f integrate+queue %.0f' % tuple(c[11:15] / max(1, c[15])))
for name, o, cyc, n in (('big', 64, c[50], c[51]), ('warp-slot', 70, c[48], c[49])):
    print(name, 'islands: mean contacts %.1f bodies %.1f sweeps %.1f rows per sweep %.1f; cycles per row pass %.0f, per contact-sweep %.0f' % (
        c[o + 2] / max(1, n), c[o + 3] / max(1, n), c[o + 4] / max(1, n), c[o] / max(1, c[o + 4]), cyc * 1024 / max(1, c[o]), cyc * 1024 / max(1, c[o + 1])))
