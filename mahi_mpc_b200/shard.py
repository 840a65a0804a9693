"""Multi-GPU plumbing: MPC instances are independent, so a batch is partitioned contiguously over
the ranks of one node (one process per GPU, launched by torchrun) with NO collective on the solve path; the only
exchange is one gather of plans + statistics to rank 0 at the end (one all-gather per array: NCCL over NVLink / NVSwitch on GPUs,
gloo in the CPU tests)."""
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
        # ONE all-gather into a contiguous [world * pad, ...] buffer (NCCL on GPUs, gloo in the CPU tests: the same code path).
        # dist.gather is world - 1 send/recv pairs into rank 0 plus a concatenation copy there: 0.36 / 0.53 ms for the bench's plans +
        # statistics at 2 / 8 GPUs, longer than the solve they follow; the all-gather runs at NVSwitch bandwidth on every rank and
        # needs no copy when the shards are equal (2 GPUs: 0.185 ms, 4: 0.265 ms).
        buf = torch.empty((world * pad,) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_gather_into_tensor(buf, send.contiguous())
        if rank != dst:
            out.append(None)
        elif all(n == pad for n in lead):
            out.append(buf)
        else:
            out.append(torch.cat([buf[r * pad:r * pad + n] for r, n in enumerate(lead)]))
    return out[0] if single else out
