"""The N > 1 path on the CPU: two gloo ranks shard a batch by galaxy, each analyses its contiguous
slice (through the emulator build here; through the CUDA library on a GPU box) and all-gathers the
result rows.  The gathered table must be bitwise identical to the single-process table."""
import os
import sys

import numpy as np
import torch
import torch.multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, B, n, q):
    sys.path.insert(0, REPO)
    import torch.distributed as dist

    import pawlikmorph_lsst_b200 as pm
    import synth
    from tests import emul_harness as H
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    x = synth.make_batch(B, n, 4242).numpy()          # every rank can regenerate the seeded batch
    lo, hi = pm.batch.shard_bounds(B, world, rank)
    rows_local = torch.from_numpy(H.analyse(x[lo:hi], 100.0, 5.0, 3, 15, want=())["rows"])
    table = pm.batch.gather_rows(rows_local, B)
    if rank == 0:
        q.put(table.numpy())
    dist.barrier()
    dist.destroy_process_group()


def _run(B, world=2):
    from tests import emul_harness as H
    import synth
    H.build()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, 48, q)) for r in range(world)]
    for p in procs:
        p.start()
    table = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    x = synth.make_batch(B, 48, 4242).numpy()
    single = H.analyse(x, 100.0, 5.0, 3, 15, want=())["rows"]
    assert np.array_equal(table, single, equal_nan=True)


def test_shard_bounds_cover_the_batch():
    import pawlikmorph_lsst_b200 as pm
    for B in (0, 1, 7, 8, 1000):
        for W in (1, 2, 3, 8):
            b = [pm.batch.shard_bounds(B, W, r) for r in range(W)]
            assert b[0][0] == 0 and b[-1][1] == B and all(b[i][1] == b[i + 1][0] for i in range(W - 1))
            assert max(h - l for l, h in b) - min(h - l for l, h in b) <= 1


def test_two_ranks_even_split():
    _run(8)


def test_two_ranks_ragged_split():
    _run(7)
