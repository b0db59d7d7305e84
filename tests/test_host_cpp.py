"""The C++ host mirror (puckswarm_b200/host/puckswarm_host.hpp) above the C-ABI: compiled with g++ against the in-tree
library, its property mapping compared with the Python mirror's (no GPU), its loud failure without a device, and -- on a
B200 -- its Arena / RunOffline / Suite / Controller path compared bit for bit with the ctypes binding."""
import os
import subprocess
import sys

import numpy as np
import pytest

from puckswarm_b200 import abi, host

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "puckswarm_b200", "csrc")


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    import __graft_entry__ as g
    g.build_engine()
    exe = str(tmp_path_factory.mktemp("hostcpp") / "host_demo")
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-o", exe, os.path.join(ROOT, "tests", "hostcpp", "host_demo.cpp"),
                           "-L" + CSRC, "-lpuckswarm_b200", "-Wl,-rpath," + CSRC])
    return exe


def _write_props(d):
    (d / "common.properties").write_text("# sweep\nrepetitions=20\nArena.nRobots=4\nArena.nPucks=40\nArena.nPuckTypes=2\nVFHPlus.TAU_LO = 0.6\n")
    (d / "X.properties").write_text("code=X\ncontrollerType=CacheCons\nCacheConsController.CACHE_SEPARATION_OPTION=NULLIFY_OLD\n"
                                    "CacheConsController.K1=0.5\nCacheConsController.IGNORE_PUCKS=true\nArena.nRobots=3\nArena.puckShiftStep=7\n")
    (d / "P.properties").write_text("controllerType=ProbSeek\nProbSeekController.K2=3\nProbSeekController.PLACEMENT_TIME=12\nCacheConsController.K2=99\n")
    (d / "BAD.properties").write_text("Arena.nRobots=four\n")


def test_property_mapping_matches_python_mirror(demo, tmp_path):
    _write_props(tmp_path)
    for code in ("X", "P"):
        out = subprocess.run([demo, "--config", str(tmp_path), code], capture_output=True, text=True)
        assert out.returncode == 0, out.stderr
        got = dict(kv.split("=") for kv in out.stdout.split())
        cfg = host.Experiment.from_files(str(tmp_path), code, 0).to_config()
        for k, v in got.items():
            want = getattr(cfg, k)
            assert float(v) == pytest.approx(float(want), rel=1e-6), (code, k, v, want)
    bad = subprocess.run([demo, "--config", str(tmp_path), "BAD"], capture_output=True, text=True)
    assert bad.returncode == 4 and "four" in bad.stderr      # NumberFormatException of the reference


def test_java_number_formatting_matches_python_mirror(demo):
    """Float.toString / Double.toString as the snapshot files need them: the C++ and the Python mirror agree on 3 000 values."""
    rng = np.random.RandomState(0)
    vals = np.concatenate([rng.uniform(-100, 100, 2000), rng.uniform(-1, 1, 500) * 1e-4, rng.uniform(-1, 1, 500) * 1e8,
                           [0.0, 1.5, 100.0, 1e-4, 1.25e7, 1e-3, 1e7, 9999999.0, 123456.7, -38.847805, 3.0e-5, 1.0e10]]).astype(np.float32)
    out = subprocess.run([demo, "--format"], input="\n".join(repr(float(v)) for v in vals), capture_output=True, text=True)
    lines = out.stdout.strip().split("\n")
    assert len(lines) == len(vals)
    for v, l in zip(vals, lines):
        f, d = l.split()
        assert f == host.java_float_str(v) and d == host.java_double_str(np.float32(v)), (v, l)


def test_no_device_is_loud(demo, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    _write_props(tmp_path)
    calib = abi.write_calibration_bin(str(tmp_path / "calib.bin"))
    out = subprocess.run([demo, str(tmp_path), "X", calib, "2", "6"], capture_output=True, text=True)
    assert out.returncode == 3 and "engine error -2" in out.stderr   # PS_E_NODEVICE: there is no CPU path


@pytest.mark.gpu
@pytest.mark.parametrize("external", [False, True])
def test_cpp_host_matches_ctypes_binding(demo, tmp_path, external):
    from puckswarm_b200 import engine
    _write_props(tmp_path)
    calib = abi.write_calibration_bin(str(tmp_path / "calib.bin"))
    A, ticks = 3, 30
    env = dict(os.environ, HOST_DEMO_SNAPSHOT_DIR=str(tmp_path / "snap_cpp"))
    out = subprocess.run([demo, str(tmp_path), "X", calib, str(A), str(ticks)] + (["external"] if external else []), capture_output=True, text=True, env=env)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().split("\n")
    assert lines[0].startswith("code X stepCount %d nBodies 43 robot0 R0" % (ticks // 3))
    cfg = host.Experiment.from_files(str(tmp_path), "X", 0).to_config(n_arenas=A)
    if external:
        cfg.controller = abi.PS_CTRL_EXTERNAL
    e = engine.Engine(cfg, device=0)
    e.reset(np.arange(A))
    for _ in range(ticks // 3):
        if external:
            e.sense()
            e.step_external(np.full((A, 3), 300000.0, np.float32), np.full((A, 3), 1000000.0, np.float32))
        else:
            e.step(3)
    body = e.get_state().body
    NB = body.shape[2]
    for a in range(A):
        want = np.array([body[a, 0, NB - 3], body[a, 1, NB - 3], body[a, 2, NB - 3], body[a, 0, 0], body[a, 1, 0]], np.float32).view(np.uint32)
        got = [int(x, 16) for x in lines[1 + a].split()[2:]]
        assert got == [int(x) for x in want], (a, lines[1 + a])
    if not external:
        # the reference's snapshot files, written by the C++ mirror and by the Python mirror after the same run
        e.close()
        arena = host.Arena(host.Experiment.from_files(str(tmp_path), "X", 0), list(range(A)))
        for _ in range(ticks // 3):
            arena.step()
        written = arena.save_snapshot(str(tmp_path / "snap_py"))
        arena.dispose()
        assert len(written) == A * (3 + 3 * 2)     # robots + two colours + per robot cachePoints and stateCounts
        for f in written:
            rel = os.path.relpath(f, str(tmp_path / "snap_py"))
            assert open(f).read() == open(os.path.join(str(tmp_path / "snap_cpp"), rel)).read(), rel
