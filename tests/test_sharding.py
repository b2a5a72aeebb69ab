"""N>1 path on CPU: two gloo ranks shard a batch, run their shard (through the CPU oracle, which is
the checker here), and the gathered result equals the single-process result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from weiner_slamit_v2_b200.sharding import gather_counters, max_over_ranks, shard_range


def test_shard_range_partitions_exactly():
    for total in (0, 1, 7, 256, 1000, 1024):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import oracle_lib as O
    from weiner_slamit_v2_b200.frames import synthetic_frame
    lo, hi = shard_range(total, rank, world)
    orc = O.OracleExtractor(300, 1.2, 4, 20, 7)
    nk, digest = 0, 0
    for i in range(lo, hi):
        k, d = orc(synthetic_frame(i, 320, 240))
        nk += len(k)
        digest += int(d.astype(np.int64).sum())
    tot = gather_counters({"frames": hi - lo, "keypoints": nk, "digest": digest})
    tmax = max_over_ranks([float(rank + 1), 10.0 - rank])
    if rank == 0:
        q.put((tot, tmax))
    dist.destroy_process_group()


def test_two_rank_gloo_shards_equal_single_process():
    import oracle_lib as O
    from weiner_slamit_v2_b200.frames import synthetic_frame
    total = 5
    orc = O.OracleExtractor(300, 1.2, 4, 20, 7)
    nk, digest = 0, 0
    for i in range(total):
        k, d = orc(synthetic_frame(i, 320, 240))
        nk += len(k); digest += int(d.astype(np.int64).sum())
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    [p.start() for p in procs]
    tot, tmax = q.get(timeout=120)
    [p.join(60) for p in procs]
    assert tot == {"frames": total, "keypoints": nk, "digest": digest}
    assert tmax == [2.0, 10.0]
