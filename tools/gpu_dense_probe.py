import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from puckswarm_b200 import abi, engine
for scale, P, R in ((0.55, 110, 4), (0.5, 100, 4), (0.6, 150, 6)):
    cfg = abi.config_defaults(n_arenas=4, n_robots=R, n_pucks=P, enclosure_scale=scale)
    e = engine.Engine(cfg)
    t = time.time(); e.reset(np.arange(4)); print('reset', time.time() - t)
    e.island_hist()
    for blk in range(4):
        e.step(60); e.sync()
        h = e.island_hist(); w = int(h[7, 7])
        print(scale, P, 'steps', blk * 20, 'slowest island contacts', (w >> 20) & 0xFFF, 'bodies', (w >> 10) & 0x3FF, 'sweeps', w & 0x3FF, 'nofit', int(h[7, 6]))
    e.close()
