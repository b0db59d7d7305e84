set -x
nvidia-smi --query-gpu=name,memory.total --format=csv
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
timeout 600 python - <<'PY' > gpurun_out/quick_bench.log 2>&1
import sys, time
sys.path.insert(0,'.')
import numpy as np
from puckswarm_b200 import abi, engine
for A in (256, 4096):
    cfg = abi.config_defaults(n_arenas=A, n_robots=16, n_pucks=200, rng_mode=abi.PS_RNG_JAVA_SHARED)
    e = engine.Engine(cfg)
    t=time.time(); e.reset(np.arange(A)); print('reset s', time.time()-t)
    e.step(30)
    for n in (30, 90):
        t=time.time(); e.step(n); dt=time.time()-t
        print('A',A,'ticks',n,'s',dt,'robot-steps/s', A*16*n/dt, flush=True)
    st = e.step_stats(); print('pos iters mean', st[:,0].mean(), 'max', st[:,0].max(), 'err', e.get_state().arena['error_flags'].max())
PY
cat gpurun_out/quick_bench.log
