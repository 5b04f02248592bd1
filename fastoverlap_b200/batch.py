"""Multi-GPU partitioning: independent alignment pairs are sharded over GPUs, every shard runs
the same single-GPU hot path, results are concatenated on the host.  There is no data-path
collective (BASELINE.json north_star: "no NCCL on the hot path"); torch.distributed is only used
to gather the small per-pair results when one process per GPU is used.

Two ways to drive N GPUs of one box:
  * one process per GPU (torchrun): shard_bounds + gather_results
  * one process, one host thread + one Context per GPU: MultiGPU (ctypes releases the GIL)
Per-pair results are bit-identical for any shard count (each pair is computed independently).
"""
import threading

import numpy as np


def shard_bounds(n, rank, world):
    """Contiguous, balanced [lo, hi) of n items for `rank` of `world`."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def run_sharded(fn, arrays, rank, world):
    """Apply fn to this rank's shard of every array in `arrays` (sharded along axis 0)."""
    lo, hi = shard_bounds(len(arrays[0]), rank, world)
    return fn(*[a[lo:hi] for a in arrays]), (lo, hi)


def gather_results(local, dist=None, dst=0):
    """Concatenate per-rank tuples of numpy arrays on rank `dst` (None elsewhere)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return tuple(local)
    objs = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(tuple(local), objs, dst=dst)
    if dist.get_rank() != dst:
        return None
    return tuple(np.concatenate([o[i] for o in objs]) for i in range(len(local)))


class MultiGPU(object):
    """One Context per GPU, one host thread per GPU."""

    def __init__(self, devices):
        from . import _lib
        self.ctxs = [_lib.Context(d) for d in devices]

    def map_pairs(self, fn, *arrays):
        """fn(ctx, *shard) -> tuple of arrays; shards along axis 0; results concatenated in order."""
        world = len(self.ctxs)
        out = [None] * world
        err = []

        def work(r):
            try:
                lo, hi = shard_bounds(len(arrays[0]), r, world)
                out[r] = fn(self.ctxs[r], *[a[lo:hi] for a in arrays])
            except Exception as exc:  # surfaced to the caller below
                err.append(exc)

        ts = [threading.Thread(target=work, args=(r,)) for r in range(world)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        if err:
            raise err[0]
        return tuple(np.concatenate([o[i] for o in out]) for i in range(len(out[0])))
