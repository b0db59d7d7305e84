"""FAST mode gate (run on the GPU box): ps_config.pos_iters = 10 instead of the 100 position iterations RunOffline passes to
World.step (RunOffline.java:94). It is NOT the reference's configuration (the headline numbers never use it); it is the same
engine and the same oracle-exact arithmetic with a lower iteration cap -- capped islands (100 sweeps without meeting JBox2D's
tolerance) are half of all position-solver work in the steady state. The gate asks whether the experiment's OUTCOME statistics
change: 1 000 seeded episodes (4 robots / 40 pucks / 2 colours, CacheCons) are run in both modes and the distributions of
Analyze's per-episode numbers (percent completion at the end, time-weighted completion, steps to completion; clustering
threshold 9.45 = the paper's 1.5 * Puck.WIDTH) are compared with a two-sample Kolmogorov-Smirnov test.

  python tools/gpu_fast_mode_gate.py [episodes] [think_steps] [pos_iters] > profiles/r02_fast_mode_gate.json
"""
import json
import sys
import time

sys.path.insert(0, '.')
import numpy as np
from scipy import stats
from puckswarm_b200 import abi, engine

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
STEPS = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
FAST = int(sys.argv[3]) if len(sys.argv) > 3 else 10


def run(pos_iters):
    cfg = abi.config_defaults(n_arenas=N, n_robots=4, n_pucks=40, controller=abi.PS_CTRL_CACHECONS, pos_iters=pos_iters)
    e = engine.Engine(cfg)
    e.analysis_config(9.45, 50.0)
    e.reset(np.arange(N))
    t = time.time()
    for _ in range(STEPS // 100):
        e.step(300, sync=False)
    e.step(1)                      # the record of the last storage step
    dt = time.time() - t
    an = e.get_analysis()
    twc = an["twc_num"] / np.maximum(an["twc_den"], 1e-300)
    out = {"completion": an["completion"].copy(), "twc": twc, "stc": an["steps_to_completion"].astype(np.float64)}
    e.close()
    return out, N * 4 * 3 * STEPS / dt


ref, v_ref = run(100)
fast, v_fast = run(FAST)
res = {"episodes": N, "think_steps": STEPS, "shape": "4 robots / 40 pucks / 2 colours, CacheCons, seeds 0..%d" % (N - 1),
       "pos_iters_reference": 100, "pos_iters_fast": FAST, "robot_steps_per_s": {"pos_iters_100": v_ref, "pos_iters_%d" % FAST: v_fast}, "ks": {}}
for k in ("completion", "twc", "stc"):
    a, b = ref[k], fast[k]
    if k == "stc":   # episodes that never reached the target count as "later than any step"
        a = np.where(a < 0, STEPS + 100, a)
        b = np.where(b < 0, STEPS + 100, b)
    d, p = stats.ks_2samp(a, b)
    res["ks"][k] = {"D": float(d), "p_value": float(p), "mean_reference": float(np.mean(a)), "mean_fast": float(np.mean(b)),
                    "std_reference": float(np.std(a)), "std_fast": float(np.std(b))}
res["gate"] = "pass" if all(v["p_value"] > 0.01 for v in res["ks"].values()) else "fail"
res["note"] = "gate = no KS test rejects 'same distribution' at the 1 % level"
print(json.dumps(res, indent=1))
