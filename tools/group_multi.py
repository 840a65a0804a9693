"""Developer tool: mmpc_group_* over ALL GPUs of the box in ONE process (the product's multi-GPU path; bench.py --gpus N is one
process per GPU).  Headline workload (double pendulum, N = 50), 4096 instances per device: checks the sharded solve against a
single-device solve of one shard, times mmpc_group_solve (host buffers, wall clock) and the NVLink gather to device 0.
    python -m tools.group_multi [out.json]"""
import json
import sys
import time

import numpy as np
import torch

import mahi_mpc_b200.api as api
from mahi_mpc_b200 import workload

n = api.device_count()
N, per = 50, 4096
B = per * n
x0, up, traj = workload.synthetic_batch(4, 2, N, 0.002, B)
Q, R, Rm = workload.default_weights(4, 2)
m = api.Model("double_pendulum", N=N)
g = api.Group(m, B, list(range(n)), warm_start=0)
g.set_weights(Q, R, Rm)
V, lam, st = g.solve(x0, up, traj, want_lam=False)
assert (st["status"] == 0).all()
s = api.Solver(m, per, device=n - 1, warm_start=0); s.set_weights(Q, R, Rm)
V1, _, st1 = s.solve(x0[-per:], up[-per:], traj[-per:], want_lam=False)
assert np.array_equal(V1, V[-per:]) and np.array_equal(st1["iterations"], st["iterations"][-per:])
ts = []
for _ in range(10):
    t0 = time.perf_counter(); g.solve(x0, up, traj, want_lam=False); ts.append(time.perf_counter() - t0)
Vd = torch.empty(B, m.NV, dtype=torch.float64, device="cuda:0"); sd = torch.empty(B, 5, dtype=torch.float64, device="cuda:0")
gm = [g.gather(0, Vd.data_ptr(), sd.data_ptr()) for _ in range(5)]
assert np.array_equal(Vd.cpu().numpy(), V)
out = dict(workload="mmpc_group_solve: double pendulum, N=50, 4096 instances per device, host (pageable) buffers, one process",
           n_gpus=n, batch_total=B, shards=g.shards(), solve_ms_median=1e3 * float(np.median(ts)), e2e_solves_per_s=B / float(np.median(ts)),
           gather_to_device0_ms=float(np.median(gm)), gathered_bytes=int(B * (m.NV * 8 + 40)))
print(json.dumps(out))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
