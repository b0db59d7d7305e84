"""world_size-2 gloo test of the N>1 host logic: arenas shard by index (no data-path collective) and the per-arena
statistics are summed with one all-reduce. The engine itself needs a GPU, so the per-rank statistics here come from
the oracle stepping the rank's shard; the sharding and the reduction are exactly the code bench.py runs under NCCL."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    from puckswarm_b200 import abi, dist as psd
    import oracle_binding as oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_total = 5
    lo, hi = psd.shard_range(n_total, rank, world)
    cfg = abi.config_defaults(n_arenas=hi - lo, n_robots=2, n_pucks=10)
    o = oracle.Oracle(cfg)
    o.analysis_config(9.45, 10.0)
    o.reset(np.arange(lo, hi))
    o.step(330)   # storage steps 0 and 100: two Analyze rows per arena
    total = psd.allreduce_stats(o.compute_stats(9.45))
    q.put((rank, lo, hi, [a.tolist() for a in psd.stats_to_arrays(total)]))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_and_reduce_two_ranks():
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from puckswarm_b200 import abi, dist as psd
    import oracle_binding as oracle
    assert [psd.shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    assert [psd.shard_range(8, r, 4) for r in range(4)] == [(0, 2), (2, 4), (4, 6), (6, 8)]
    assert sum(hi - lo for lo, hi in (psd.shard_range(65536, r, 8) for r in range(8))) == 65536
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process truth over all 5 arenas
    cfg = abi.config_defaults(n_arenas=5, n_robots=2, n_pucks=10)
    o = oracle.Oracle(cfg)
    o.reset(np.arange(5))
    o.analysis_config(9.45, 10.0)
    o.reset(np.arange(5))
    o.step(330)
    truth = o.compute_stats(9.45)
    arrays = psd.stats_to_arrays(truth)
    assert truth.n_snapshots == 10 and truth.max_steps_to_completion >= 0 and truth.min_completion <= truth.max_completion
    for rank, lo, hi, got in res:
        for (first, last, dt, red), g, t in zip(abi.STATS_SEGMENTS, got, arrays):
            if dt == np.int64:
                assert g == t.tolist(), (first, red)            # sums, minimum and maximum of integers: exact
            else:
                assert np.allclose(g, t, rtol=1e-12, atol=1e-12), (first, red)
    # round trip of the struct <-> array mapping
    s = psd.arrays_to_stats(arrays)
    for a, b in zip(psd.stats_to_arrays(s), arrays):
        assert np.array_equal(a, b)
    assert psd.ensemble_summary(truth)["repetitions"] == 5
