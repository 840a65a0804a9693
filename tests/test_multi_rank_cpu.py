"""world_size-2 gloo test of the multi-GPU host logic: contiguous sharding of the synthetic batch,
independent per-rank solves (here by the ORACLE, there is no GPU in this suite), one gather to rank 0, and equality
with the un-sharded result."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mahi_mpc_b200 import shard, workload


def test_shard_range_partitions():
    for total, world in [(4096, 1), (4096, 8), (16384, 4), (10, 3), (2, 4)]:
        parts = [shard.shard_range(total, world, r) for r in range(world)]
        assert parts[0][0] == 0 and sum(c for _, c in parts) == total
        for (o, c), (o2, _) in zip(parts, parts[1:]):
            assert o + c == o2
        assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, N, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import coracle, nlp
        off, cnt = shard.shard_range(total, world, rank)
        x0, up, traj = workload.synthetic_batch(4, 2, N, 0.002, cnt, offset=off)
        Q, R, Rm = workload.default_weights(4, 2)
        P = nlp.Problem(N=N)
        p = np.stack([P.pack_p(traj[b], Q, R, Rm, up[b]) for b in range(cnt)])
        V, lam, st, bad = coracle.COracle("O3").solve_batch("double_pendulum", N, P.h, False, x0, p, tol=1e-10, nthreads=1)
        assert bad == 0
        counts = [shard.shard_range(total, world, r)[1] for r in range(world)]
        Vg, itg = shard.gather_to_root([torch.from_numpy(V), torch.from_numpy(st["iters"].astype(np.int64))], counts=counts)
        if rank == 0:
            q.put((Vg.numpy(), itg.numpy()))
        else:
            assert Vg is None
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_shard_solve_gather(coracle):
    from oracle import nlp
    total, N, world = 7, 10, 2  # odd total: unequal shards
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, N, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    Vg, itg = q.get(timeout=150)
    for p_ in procs:
        p_.join(timeout=60)
        assert p_.exitcode == 0
    x0, up, traj = workload.synthetic_batch(4, 2, N, 0.002, total)
    Q, R, Rm = workload.default_weights(4, 2)
    P = nlp.Problem(N=N)
    p = np.stack([P.pack_p(traj[b], Q, R, Rm, up[b]) for b in range(total)])
    V, _, st, bad = coracle.solve_batch("double_pendulum", N, P.h, False, x0, p, tol=1e-10, nthreads=1)
    assert bad == 0 and Vg.shape == (total, P.NV)
    assert np.array_equal(Vg, V) and np.array_equal(itg, st["iters"])
