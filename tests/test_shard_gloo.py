"""The N > 1 path on CPU: world_size-2 gloo processes shard a batch of frames (disjoint, complete), run the hot path's
CPU stand-in (the oracle) on their share, and the end-of-run collectives (max of timers, all-gather of counters)
reproduce the single-process totals."""
import os
import socket
import sys
import zlib

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from orbx.shard import sequence_owner, shard_range
    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    assert [sequence_owner(s, 8) for s in range(10)] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _frame_record(i):
    """what one unit contributes to the counters: keypoints and a checksum of keypoints + descriptors"""
    from oracle import oracle_py as O
    from orbx import synth
    kp, de = O.Extractor(300, 1.2, 4, 20, 7)(synth.g_rect(i, 320, 240, nrect=200))
    return len(kp), zlib.crc32(kp.tobytes() + de.tobytes())


def _worker(rank, world, port, n_units, q):
    for p in (ROOT, os.path.join(ROOT, "active-orb-slam2_b200")):
        sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from orbx.shard import gather_counters, max_over_ranks, shard_range
    lo, hi = shard_range(n_units, rank, world)
    nkp, chk = 0, 0
    for i in range(lo, hi):
        n, c = _frame_record(i)
        nkp += n
        chk ^= c
    elapsed = 1.0 + rank                       # stand-in timers: the slowest rank defines the step
    tmax = max_over_ranks([elapsed, 10.0 - rank])
    recs = gather_counters([hi - lo, nkp, chk])
    dist.barrier()
    if rank == 0:
        q.put((tmax, recs))
    dist.destroy_process_group()


def test_two_rank_gloo_shards_match_single_process():
    n_units, world = 5, 2
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_units, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    tmax, recs = q.get()
    assert tmax == [2.0, 10.0]                                  # max over ranks of each timer
    assert [r[0] for r in recs] == [3, 2]                       # 5 units -> 3 + 2
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "active-orb-slam2_b200"))
    single = [_frame_record(i) for i in range(n_units)]
    assert sum(r[1] for r in recs) == sum(n for n, _ in single)
    chk = 0
    for _, c in single:
        chk ^= c
    assert recs[0][2] ^ recs[1][2] == chk
