"""Developer tool: end-to-end (pinned host buffers) timing of mmpc_solve for the bench workload."""
import os, time, sys
import numpy as np, torch
import mahi_mpc_b200.api as api
from mahi_mpc_b200 import workload
N, B = 50, int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x0, up, traj = workload.synthetic_batch(4, 2, N, 0.002, B)
Q, R, Rm = workload.default_weights(4, 2)
m = api.Model("double_pendulum", N=N)
s = api.Solver(m, B, warm_start=0, kernel_variant=int(os.environ.get("MMPC_VARIANT", "0")))
s.set_weights(Q, R, Rm)
h = [torch.from_numpy(a).pin_memory() for a in (x0, up, traj)]
V = torch.empty(B, m.NV, dtype=torch.float64).pin_memory(); st = torch.empty(B, 5, dtype=torch.float64).pin_memory()
for _ in range(5):
    s.solve(*h, V_out=V, stats_out=st, want_lam=False)
ts = []
for _ in range(30):
    t0 = time.perf_counter(); s.solve(*h, V_out=V, stats_out=st, want_lam=False); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print(f"B={B} e2e median {np.median(ts):.3f} ms min {ts.min():.3f} ms  -> {B/np.median(ts)*1e3:.3e} solves/s   device span {s.last_timing()}")
