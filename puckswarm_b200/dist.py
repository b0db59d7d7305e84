"""Multi-GPU plumbing: arenas shard by index, one process per GPU, no data-path collective.

The only cross-rank step is the reduction of the per-arena experiment statistics (ps_stats: cluster-size
histograms per colour, controller-state occupancy, carry / cache / contact counts, summed percent completion)
with torch.distributed -- NCCL on the GPUs (a few KB over NVLink), gloo in the CPU tests.
"""
import ctypes as C
import numpy as np

from . import abi


def shard_range(n_arenas, rank, world_size):
    """Contiguous block of arena indices owned by `rank`: the first (n % world) ranks take one extra arena."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank / world_size")
    base, extra = divmod(n_arenas, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _segment(s, first, last, dtype):
    """The bytes of ps_stats from field `first` to field `last` as a numpy view of `dtype` (fields of a segment are contiguous)."""
    lo = getattr(abi.ps_stats, first).offset
    hi = getattr(abi.ps_stats, last).offset + getattr(abi.ps_stats, last).size
    buf = (C.c_char * C.sizeof(s)).from_address(C.addressof(s))
    return np.frombuffer(buf, dtype=np.uint8)[lo:hi].view(dtype)


def stats_to_arrays(s):
    """ps_stats -> one array per reduction segment (abi.STATS_SEGMENTS: sums, minima, maxima over int64 and float64)."""
    return [_segment(s, a, b, dt).copy() for a, b, dt, _ in abi.STATS_SEGMENTS]


def arrays_to_stats(arrays):
    s = abi.ps_stats()
    for (a, b, dt, _), arr in zip(abi.STATS_SEGMENTS, arrays):
        _segment(s, a, b, dt)[:] = np.asarray(arr, dt)
    return s


def allreduce_stats(stats, device=None, group=None):
    """Reduce ps_stats over all ranks with torch.distributed (one collective per segment: SUM / MIN / MAX). `device` is a torch
    device ('cuda:N' for NCCL, None / 'cpu' for gloo). The C-ABI's ps_get_stats(h, thr, out, nccl_comm) does the same on the
    device copy with ncclAllReduce for hosts without torch (the Java shim)."""
    import torch
    import torch.distributed as dist
    arrays = stats_to_arrays(stats)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return arrays_to_stats(arrays)
    ops = {"sum": dist.ReduceOp.SUM, "min": dist.ReduceOp.MIN, "max": dist.ReduceOp.MAX}
    out = []
    for (_, _, _, red), arr in zip(abi.STATS_SEGMENTS, arrays):
        t = torch.from_numpy(arr.copy())
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=ops[red], group=group)
        out.append(t.cpu().numpy())
    return arrays_to_stats(out)


def mean_completion(stats):
    """Analyze.java's ensemble average of percent completion over the arenas of the (reduced) statistics."""
    return stats.sum_completion / max(1, stats.n_arenas)


def ensemble_summary(stats):
    """What Analyze.computeEnsembleData prints for an experiment code (Analyze.java:240-258): avgTWC, and avgSTC or the number of
    incompletions, from the (reduced) statistics."""
    n = max(1, stats.n_arenas)
    incompletions = int(stats.n_arenas - stats.n_completed)
    return {"avgTWC": stats.sum_twc / n, "avgSTC": (stats.sum_steps_to_completion / n) if incompletions == 0 else None,
            "incompletions": incompletions, "repetitions": int(stats.n_arenas)}
