"""Developer tool: iteration histogram of one kernel variant for several batch sizes.
    python -m tools.it_hist variant B1 B2 ..."""
import os
import numpy as np, sys
import mahi_mpc_b200.api as api
from mahi_mpc_b200 import workload
N = 50; v = int(sys.argv[1])
Q, R, Rm = workload.default_weights(4, 2)
m = api.Model("double_pendulum", N=N)
for B in [int(a) for a in sys.argv[2:]]:
    x0, up, traj = workload.synthetic_batch(4, 2, N, 0.002, B)
    s = api.Solver(m, B, warm_start=0, kernel_variant=v); s.set_weights(Q, R, Rm)
    if os.environ.get("MMPC_UBOUND"):
        ub = float(os.environ["MMPC_UBOUND"]); s.set_bounds(u_min=[-ub, -ub], u_max=[ub, ub])
    for rep in range(2):
        if os.environ.get("MMPC_DEV"):  # device-resident inputs (single launch, no sub-batches)
            import torch
            d = [torch.from_numpy(a).cuda() for a in (x0, up, traj)]
            Vt, lt, stt = s.solve(*d); torch.cuda.synchronize()
            st = api.stats_from_tensor(stt)
        else:
            V, lam, st = s.solve(x0, up, traj)
        print("variant", v, "B", B, "iters hist", np.bincount(st["iterations"]), "first 8", st["iterations"][:8], "ls", st["ls_steps"].sum(), "reg", st["reg_steps"].sum(),
              "obj sum %.12e" % st["objective"].sum(), "kkt max %.3e" % st["kkt_error"].max(), "status", np.bincount(st["status"]))
