"""Pair-search phase breakdown (needs a diagnostic build with the q0..q4 clocks patched into physFindNewContacts)."""
import sys
sys.path.insert(0, '.')
import numpy as np
from puckswarm_b200 import abi, engine
A = 4096
cfg = abi.config_defaults(n_arenas=A, n_robots=16, n_pucks=200)
e = engine.Engine(cfg)
e.reset(np.arange(A))
e.step(45)
e.step(3)
st = e.step_stats().astype(np.float64)
for i, n in enumerate(['bodypairs', 'proxypairs', 'wall', 'sort']):
    print('%-12s mean %10.1f max %10.1f' % (n, st[:, i].mean() * 16, st[:, i].max() * 16))
print('nMB mean', (st[:, 4] // 65536).mean(), 'nBP mean', (st[:, 4] % 65536).mean(), 'max', (st[:, 4] % 65536).max())
print('findnew total', st[:, 6].mean() * 16)
