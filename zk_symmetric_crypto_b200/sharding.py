"""Multi-GPU host logic: independent proofs shard across ranks with no data-path collective.

One process per GPU (torchrun); `torch.distributed` is plumbing only: a barrier, a max-over-ranks of the device-timed step and
a gather of the finished proofs to rank 0.  Works with the nccl backend on GPUs and with gloo on CPU (tests)."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_indices(n_items, world_size, rank):
    """Round-robin assignment of proof instances to ranks (proofs of one batch have equal size, so this balances)."""
    return list(range(rank, n_items, world_size))


def max_over_ranks(values, device="cpu"):
    """Element-wise max of a list of floats over all ranks (multi-GPU timings are the max over ranks, never wall clock)."""
    w, _ = world()
    if w == 1:
        return [float(v) for v in values]
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def prove_batch(jobs, prove_fn):
    """jobs: list of picklable job descriptions, identical on every rank.  Each rank proves its round-robin share with
    prove_fn(job) -> bytes; rank 0 returns the proofs in job order, other ranks return None."""
    w, r = world()
    mine = {i: prove_fn(jobs[i]) for i in shard_indices(len(jobs), w, r)}
    if w == 1:
        return [mine[i] for i in range(len(jobs))]
    gathered = [None] * w if r == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if r != 0:
        return None
    merged = {}
    for part in gathered:
        merged.update(part)
    assert sorted(merged) == list(range(len(jobs))), "every job must be proved exactly once"
    return [merged[i] for i in range(len(jobs))]
