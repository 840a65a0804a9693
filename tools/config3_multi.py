"""BASELINE config 3 on several GPUs: 4-DOF arm, N = 50, batch 16384 in total, sharded contiguously over the ranks
(strong scaling; one process per GPU, no collective on the solve path -- mahi_mpc_b200/shard.py).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 --master-port P -m tools.config3_multi [out.json]
Timing: barrier + CUDA events around each cold-start batched solve of the local shard, 3 warm-up calls, median of 5,
MAX over ranks; rank 0 prints one JSON object."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

import mahi_mpc_b200.api as api
from mahi_mpc_b200 import shard, workload


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B_total, N, nx, nu = 16384, 50, 8, 4
    lo, cnt = shard.shard_range(B_total, world, rank)  # (offset, count)
    hi = lo + cnt
    x0, up, traj = workload.synthetic_batch(nx, nu, N, 0.002, B_total)
    traj *= 0.5
    x0, up, traj = x0[lo:hi], up[lo:hi], traj[lo:hi]
    B = hi - lo
    Q, R, Rm = workload.default_weights(nx, nu)
    m = api.Model("arm4", N=N)
    s = api.Solver(m, B, warm_start=0, device=local)
    s.set_weights(Q, R, Rm)
    d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x0, up, traj)]
    V = torch.empty(B, m.NV, dtype=torch.float64, device=dev)
    st = torch.empty(B, 5, dtype=torch.float64, device=dev)
    for _ in range(3):
        s.solve(*d, V_out=V, stats_out=st, want_lam=False)
    ts = []
    for _ in range(5):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s.solve(*d, V_out=V, stats_out=st, want_lam=False); b.record(); torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ts.append(t.item())
    stats = api.stats_from_tensor(st)
    conv = torch.tensor([float((stats["status"] == 0).sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(conv)
    if rank == 0:
        ms = float(np.median(ts))
        out = dict(workload="BASELINE configs[2]: 4-DOF arm, N=50, batch 16384 in total (strong scaling)", n_gpus=world, batch_per_gpu=B,
                   ms=ms, solves_per_s=conv.item() / ms * 1e3, converged=conv.item() / B_total,
                   iterations_mean_rank0=float(stats["iterations"].mean()))
        print(json.dumps(out))
        if len(sys.argv) > 1:
            json.dump(out, open(sys.argv[1], "w"), indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
