"""ps_get_stats(h, threshold, out, nccl_comm): the statistics reduce of SURVEY 8(b)/(e) inside the C-ABI -- ncclAllReduce on
the device copy of ps_stats over the caller's communicator (one collective per reduction group: sums, minima, maxima).
Runs with one rank per visible GPU (up to 2; a one-rank communicator on a one-GPU box still goes through NCCL), arenas
sharded by index, and compares every rank's result with the oracle over all arenas."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
N_TOTAL, TICKS = 6, 301


class NcclUniqueId(C.Structure):       # ncclUniqueId is passed BY VALUE to ncclCommInitRank
    _fields_ = [("internal", C.c_char * 128)]


def _nccl_lib():
    import torch  # noqa: F401  (its bundled libnccl.so.2 is then the copy both this test and the engine resolve by soname)
    lib = C.CDLL(os.environ.get("PS_NCCL_LIB", "libnccl.so.2"))
    lib.ncclGetUniqueId.argtypes = [C.c_void_p]
    lib.ncclCommInitRank.argtypes = [C.POINTER(C.c_void_p), C.c_int, NcclUniqueId, C.c_int]
    lib.ncclCommDestroy.argtypes = [C.c_void_p]
    return lib


def _worker(rank, world, uid_bytes, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    from puckswarm_b200 import abi, engine, dist as psd
    torch.cuda.set_device(rank)
    lib = _nccl_lib()
    uid = NcclUniqueId.from_buffer_copy(uid_bytes)
    comm = C.c_void_p()
    rc = lib.ncclCommInitRank(C.byref(comm), world, uid, rank)
    assert rc == 0, rc
    lo, hi = psd.shard_range(N_TOTAL, rank, world)
    cfg = abi.config_defaults(n_arenas=hi - lo, n_robots=3, n_pucks=20)
    e = engine.Engine(cfg, device=rank)
    e.analysis_config(9.45, 10.0)
    e.reset(np.arange(lo, hi))
    e.step(TICKS)
    local = e.compute_stats(9.45)
    total = e.get_stats(9.45, nccl_comm=comm.value)
    q.put((rank, int(local.n_arenas), [a.tolist() for a in psd.stats_to_arrays(total)]))
    e.close()
    lib.ncclCommDestroy(comm)


def test_stats_allreduce_through_the_c_abi():
    import torch
    import torch.multiprocessing as mp
    from puckswarm_b200 import abi, dist as psd
    import oracle_binding as oracle
    world = min(2, torch.cuda.device_count())
    lib = _nccl_lib()
    uid = NcclUniqueId()
    assert lib.ncclGetUniqueId(C.byref(uid)) == 0
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, bytes(uid), q)) for r in range(world)]
    for p in procs:
        p.start()
    res = []
    import queue
    import time
    deadline = time.time() + 240
    while len(res) < world:
        try:
            res.append(q.get(timeout=2))
        except queue.Empty:
            assert all(p.is_alive() or p.exitcode == 0 for p in procs), "a rank died: %r" % [p.exitcode for p in procs]
            assert time.time() < deadline, "timed out"
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    cfg = abi.config_defaults(n_arenas=N_TOTAL, n_robots=3, n_pucks=20)
    o = oracle.Oracle(cfg)
    o.analysis_config(9.45, 10.0)
    o.reset(np.arange(N_TOTAL))
    o.step(TICKS)
    truth = psd.stats_to_arrays(o.compute_stats(9.45))
    assert sum(n for _, n, _ in res) == N_TOTAL
    for rank, _, got in res:
        for (first, last, dt, red), g, t in zip(abi.STATS_SEGMENTS, got, truth):
            if dt == np.int64:
                assert g == t.tolist(), (rank, first, red)
            else:
                assert np.allclose(g, t, rtol=1e-9, atol=1e-12), (rank, first, red)
