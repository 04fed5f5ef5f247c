"""Multi-GPU plumbing of the batched many-sequence mode: frames / sequences / BA windows are
independent units, so ranks take disjoint contiguous shares and the data path has no collective.  The only
communication is one all-reduce(MAX) of the timers and one all-gather of fixed-size per-rank counters at the end.
Backend-agnostic (NCCL on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def shard_range(n_units, rank, world):
    """contiguous share [lo, hi) of rank `rank`; shares differ by at most one unit and cover [0, n_units) exactly"""
    if world < 1 or not (0 <= rank < world) or n_units < 0:
        raise ValueError("bad shard arguments")
    base, extra = divmod(n_units, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sequence_owner(seq_id, world):
    """a whole sequence stays on one GPU (tracking is a serial chain): sequence s -> rank s mod world"""
    return seq_id % world


def max_over_ranks(values, device="cpu"):
    """element-wise maximum of a list of floats over all ranks (the bench's elapsed times)"""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def gather_counters(counters, device="cpu"):
    """all-gather of a fixed-size int64 record per rank (units done, keypoints, matches, checksum) -> [world][len]"""
    t = torch.tensor(list(counters), dtype=torch.int64, device=device)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [t.tolist()]
    out = [torch.zeros_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return [o.tolist() for o in out]
