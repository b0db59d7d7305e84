"""The oracle still produces the frozen vectors of tests/golden/oracle_regression.json (made by
tools/make_oracle_regression.py). These are regression vectors of the restatement itself, not reference goldens."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_oracle_regression_vectors():
    import make_oracle_regression as m
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "oracle_regression.json")))
    for c in m.CASES:
        got = m.digest(c)
        exp = want[c["name"]]["expect"]
        for k in ("n_contacts", "step_count", "rng_seed", "ctrl_state"):
            assert got[k] == exp[k], (c["name"], k)
        assert got["sha256"] == exp["sha256"], c["name"]
