"""CPU, world_size 2, gloo: the N>1 host path (sharding + gather).  The per-shard compute here is
the CPU oracle standing in for the GPU call -- what is tested is the partitioning logic:
results for any shard count are identical to the unsharded run."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

from conftest import ROOT, golden, groups_from


def test_shard_bounds():
    from fastoverlap_b200.batch import shard_bounds
    for n in (0, 1, 7, 8, 100003):
        for w in (1, 2, 3, 8):
            b = [shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch.distributed as dist
    import oracle
    from fastoverlap_b200.batch import run_sharded, gather_results
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    g = np.load(os.path.join(ROOT, "tests", "golden", "periodic_synth.npz"))
    k = "c1_"
    sizes = g[k + "gsizes"]
    flat = g[k + "groups"]
    perm, o = [], 0
    for s in sizes:
        perm.append(flat[o:o + int(s)])
        o += int(s)
    rng = np.random.default_rng(9)
    A = np.stack([g[k + "pos1"] + rng.normal(scale=0.01, size=g[k + "pos1"].shape) for _ in range(7)])
    B = np.stack([g[k + "pos2"]] * 7)
    fn = lambda a, b: oracle.per_align_pairs(a, b, g[k + "box"], int(g[k + "n"]), int(g[k + "F"]),
                                             float(g[k + "scale"]), perm)[:3]
    local, (lo, hi) = run_sharded(fn, (A, B), rank, world)
    res = gather_results(local, dist)
    dist.barrier()
    if rank == 0:
        full = fn(A, B)
        ok = all(np.array_equal(r, f) for r, f in zip(res, full))
        q.put((ok, len(res[0])))
    dist.destroy_process_group()


def test_two_rank_sharding_matches_unsharded():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    ok, n = q.get(timeout=120)
    [p.join(timeout=60) for p in procs]
    assert ok and n == 7
    assert all(p.exitcode == 0 for p in procs)
