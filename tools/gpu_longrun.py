"""Throughput and island sizes over a long run (clusters grow as CacheCons gathers pucks)."""
import sys, time
sys.path.insert(0, '.')
import numpy as np
from puckswarm_b200 import abi, engine
A = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
N = int(sys.argv[2]) if len(sys.argv) > 2 else 600
cfg = abi.config_defaults(n_arenas=A, n_robots=16, n_pucks=200)
e = engine.Engine(cfg)
e.reset(np.arange(A))
e.island_hist()
for blk in range(N // 50):
    t = time.time(); e.step(150); e.sync(); dt = time.time() - t
    h = e.island_hist()
    worst = int(h[7, 7]); nofit = int(h[7, 6])
    st = e.compute_stats()
    print('steps %4d..%4d  %.2f M robot-steps/s  no-fit islands/arena-tick %.4f  slowest warp island: kcycles %d contacts %d bodies %d sweeps %d fit %s  touching/arena %.0f' % (
        blk * 50, blk * 50 + 50, A * 16 * 150 / dt / 1e6, nofit / (A * 150.0), worst >> 32, (worst >> 20) & 0xFFF, (worst >> 10) & 0x3FF, worst & 0x3FF,
        not bool(worst & (1 << 31)), st.n_touching_points / A), flush=True)
