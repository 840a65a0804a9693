"""Multi-GPU plumbing: MPC instances are independent, so a batch is partitioned contiguously over
the ranks of one node (one process per GPU, launched by torchrun) with NO collective on the solve path; the only
exchange is one gather of plans + statistics to rank 0 at the end (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def shard_range(total: int, world: int, rank: int):
    """Contiguous, balanced partition of `total` instances: returns (offset, count) of `rank`."""
    base, rem = divmod(int(total), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def gather_to_root(tensors, counts=None, dst: int = 0):
    """Gather per-rank tensors (same trailing shape, possibly different leading size) to `dst`.
    Returns the concatenation on `dst`, None elsewhere. `counts[r]` = leading size on rank r (needed only when
    the shards are unequal); works for CUDA tensors with the nccl backend and CPU tensors with gloo."""
    import torch
    import torch.distributed as dist

    single = not isinstance(tensors, (list, tuple))
    tensors = [tensors] if single else list(tensors)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return tensors[0] if single else tensors
    world, rank = dist.get_world_size(), dist.get_rank()
    out = []
    for t in tensors:
        lead = [t.shape[0]] * world if counts is None else list(counts)
        pad = max(lead)
        send = t if t.shape[0] == pad else torch.cat([t, t.new_zeros((pad - t.shape[0],) + tuple(t.shape[1:]))])
        bufs = [torch.empty_like(send) for _ in range(world)] if rank == dst else None
        dist.gather(send.contiguous(), bufs, dst=dst)
        out.append(torch.cat([b[:n] for b, n in zip(bufs, lead)]) if rank == dst else None)
    return out[0] if single else out
