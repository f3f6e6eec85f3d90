"""Host-side fleet logic for N > 1 GPUs: instance-index sharding and the end-of-run merge of the
ensemble statistics.  Filter instances share nothing (Usckf.hpp:77-78 / Msckf.hpp:74-75: every object
holds its state by value), so the step path has NO collective; the only exchange is one all-reduce of
`1 + N + N*N` doubles per fleet: (count, sum x, sum x x^T) as produced on each device by
slb_ensemble_stats (include/slb.h).  Backend-agnostic: NCCL on the GPUs, gloo in the CPU tests."""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous shard [lo, hi) of `total` instances for `rank` of `world`; the remainder goes to the
    first ranks so shard sizes differ by at most one."""
    if not (0 <= rank < world) or total < 0:
        raise ValueError("bad shard request")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def stats_from_vectors(x):
    """(count, sum x, sum x x^T) flattened like slb_ensemble_stats, from vectorised means x[B, N]."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[1]
    out = np.empty(1 + n + n * n)
    out[0] = x.shape[0]
    out[1:1 + n] = x.sum(axis=0)
    out[1 + n:] = (x.T @ x).ravel()
    return out


def merge_stats(stats, group=None):
    """All-reduce (SUM) a local statistics tensor in place across the process group and return it.
    A no-op when torch.distributed is not initialised (single GPU)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def make_nccl_comm(engine, rank, world):
    """An NCCL communicator for the C-level gather (slb_gather_stats): rank 0's unique id travels through the
    already-initialised torch.distributed group (any backend), every rank then joins through the C ABI."""
    import torch.distributed as dist

    def exchange(raw):
        box = [raw]
        dist.broadcast_object_list(box, src=0)
        return box[0]
    return engine.NcclComm(rank, world, exchange)


def moments(stats, n):
    """Ensemble mean and covariance from merged (count, sum x, sum x x^T)."""
    s = np.asarray(stats, dtype=np.float64)
    cnt = s[0]
    mean = s[1:1 + n] / cnt
    second = s[1 + n:].reshape(n, n) / cnt
    return cnt, mean, second - np.outer(mean, mean)
