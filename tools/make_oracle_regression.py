#!/usr/bin/env python3
"""Regression vectors of the CPU oracle (NOT reference golden vectors: the Java reference cannot run in this image).

They freeze the oracle's own output for a few seeded configurations so that an accidental change to oracle/ shows up in
the CPU test suite. Output: tests/golden/oracle_regression.json (sha256 over the raw state arrays + a few plain numbers).
"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from puckswarm_b200 import abi   # noqa: E402
import oracle_binding as oracle  # noqa: E402

CASES = [
    dict(name="C1_cachecons_2robots_10pucks", n_arenas=1, n_robots=2, n_pucks=10, controller=abi.PS_CTRL_CACHECONS, seeds=[0], ticks=300),
    dict(name="C2_bhd_4robots_40pucks", n_arenas=2, n_robots=4, n_pucks=40, controller=abi.PS_CTRL_BHD, seeds=[0, 1], ticks=300),
    dict(name="C3shape_cachecons_16robots_200pucks", n_arenas=1, n_robots=16, n_pucks=200, controller=abi.PS_CTRL_CACHECONS, seeds=[7], ticks=60),
    dict(name="probseek_4robots_40pucks", n_arenas=1, n_robots=4, n_pucks=40, controller=abi.PS_CTRL_PROBSEEK, seeds=[3], ticks=240),
    dict(name="shift_shuffle_informed", n_arenas=1, n_robots=4, n_pucks=40, controller=abi.PS_CTRL_CACHECONS, seeds=[2], ticks=150,
         extra=dict(puck_shift_step=10, puck_shuffle_step=25, n_informed_robots=1)),
    dict(name="per_robot_rng_bhd", n_arenas=1, n_robots=4, n_pucks=40, controller=abi.PS_CTRL_BHD, seeds=[4], ticks=150,
         extra=dict(rng_mode=abi.PS_RNG_PER_ROBOT)),
]


def digest(case):
    cfg = abi.config_defaults(n_arenas=case["n_arenas"], n_robots=case["n_robots"], n_pucks=case["n_pucks"], controller=case["controller"],
                              **case.get("extra", {}))
    o = oracle.Oracle(cfg)
    o.reset(case["seeds"])
    o.step(case["ticks"])
    st = o.get_state()
    h = hashlib.sha256()
    n = int(st.arena["n_contacts"].max())
    for a in (st.body, st.fat_aabb, st.sense_aabb, st.contact_key[:, :n], st.contact_info[:, :n], st.contact_impulse[:, :n]):
        h.update(np.ascontiguousarray(a).tobytes())
    h.update(np.ascontiguousarray(st.arena["rng_seed"]).tobytes())
    return {
        "sha256": h.hexdigest(),
        "n_contacts": st.arena["n_contacts"].tolist(),
        "step_count": st.arena["step_count"].tolist(),
        "rng_seed": [int(x) for x in st.arena["rng_seed"]],
        "ctrl_state": st.robot["ctrl_state"].tolist(),
        "body0": [float(x) for x in st.body[0, :, 0]],
    }


def main():
    out = {c["name"]: dict(case={k: v for k, v in c.items() if k != "name"}, expect=digest(c)) for c in CASES}
    path = os.path.join(ROOT, "tests", "golden", "oracle_regression.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
