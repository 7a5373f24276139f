"""N>1 host logic on CPU: two processes over gloo shard a batch round-robin, "align" their
units with the CPU oracle standing in for the GPU, gather on rank 0; the column-slice handle
exchange is exercised with dummy handles.  (The data path itself has no collective.)"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_shard_indices_partition():
    from tsta_b200 import shard
    for n in (0, 1, 7, 64, 1001):
        for world in (1, 2, 3, 8):
            parts = [shard.shard_indices(n, r, world) for r in range(world)]
            allidx = np.sort(np.concatenate(parts)) if n else np.array([], dtype=np.int64)
            np.testing.assert_array_equal(allidx, np.arange(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        shard.shard_indices(4, 2, 2)


def _worker(rank, world, port, q):
    sys.path.insert(0, str(ROOT))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from oracle import binding as ob
    from tsta_b200 import shard, synth, api
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        npairs = 23
        pairs = [synth.read_pair(40 + 3 * k, 35 + 2 * k, 500 + k, None if k % 2 else 0.1) for k in range(npairs)]
        mine = shard.shard_indices(npairs, rank, world)
        vals = np.array([[r.maxsorce, r.corner] for r in (ob.psa(*pairs[k], trace=False) for k in mine)], dtype=np.int32)
        table = shard.gather_results(dist, mine, vals.reshape(len(mine), 2), npairs, world)
        # column slices: every rank derives the same slices and learns its right neighbour's handle
        slices = api.slice_columns(10_000, world)
        handle = (b"H%d" % rank).ljust(64, b"\0") if rank > 0 else b""
        handles = shard.exchange_inbox_handles(dist, handle, world)
        right = handles[rank + 1] if rank + 1 < len(slices) else None
        q.put((rank, None if table is None else table.tolist(), slices, right))
    finally:
        dist.destroy_process_group()


def test_two_ranks_gloo_round_robin_and_gather():
    import torch.multiprocessing as mp
    from oracle import binding as ob
    from tsta_b200 import synth
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in procs:
        rank, table, slices, right = q.get(timeout=120)
        got[rank] = (table, slices, right)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[1][0] is None
    table = np.array(got[0][0])
    want = np.array([[r.maxsorce, r.corner] for r in
                     (ob.psa(*synth.read_pair(40 + 3 * k, 35 + 2 * k, 500 + k, None if k % 2 else 0.1), trace=False)
                      for k in range(23))])
    np.testing.assert_array_equal(table, want)
    assert got[0][1] == got[1][1] == [(0, 6144), (6144, 3856)]
    assert got[0][2] == b"H1".ljust(64, b"\0") and got[1][2] is None
