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


def stats_to_arrays(s):
    """ps_stats -> (int64 vector, float64 vector) in a fixed field order."""
    ints = np.concatenate([
        np.array([s.n_arenas], np.int64),
        np.ctypeslib.as_array(s.cluster_size_hist).astype(np.int64).ravel(),
        np.ctypeslib.as_array(s.ctrl_state_hist).astype(np.int64).ravel(),
        np.array([s.n_carrying, s.n_caches, s.n_touching_points, s.n_contacts], np.int64),
    ])
    return ints, np.array([s.sum_completion], np.float64)


def arrays_to_stats(ints, floats):
    s = abi.ps_stats()
    ints = np.asarray(ints, np.int64)
    s.n_arenas = int(ints[0])
    h = ints[1:1 + abi.PS_MAX_COLOURS * 64].reshape(abi.PS_MAX_COLOURS, 64)
    for k in range(abi.PS_MAX_COLOURS):
        for i in range(64):
            s.cluster_size_hist[k][i] = int(h[k, i])
    o = 1 + abi.PS_MAX_COLOURS * 64
    for i in range(8):
        s.ctrl_state_hist[i] = int(ints[o + i])
    s.n_carrying, s.n_caches, s.n_touching_points, s.n_contacts = [int(x) for x in ints[o + 8:o + 12]]
    s.sum_completion = float(np.asarray(floats)[0])
    return s


def allreduce_stats(stats, device=None, group=None):
    """Sum ps_stats over all ranks. `device` is a torch device ('cuda:N' for NCCL, None / 'cpu' for gloo)."""
    import torch
    import torch.distributed as dist
    ints, floats = stats_to_arrays(stats)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return arrays_to_stats(ints, floats)
    ti = torch.from_numpy(ints.copy())
    tf = torch.from_numpy(floats.copy())
    if device is not None:
        ti, tf = ti.to(device), tf.to(device)
    dist.all_reduce(ti, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(tf, op=dist.ReduceOp.SUM, group=group)
    return arrays_to_stats(ti.cpu().numpy(), tf.cpu().numpy())


def mean_completion(stats):
    """Analyze.java's ensemble average of percent completion over the arenas of the (reduced) statistics."""
    return stats.sum_completion / max(1, stats.n_arenas)
