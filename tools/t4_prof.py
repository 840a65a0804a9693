"""Developer tool: ONE device-resident solve (after one warm-up) of a kernel variant, for the in-kernel timers of profile builds.
    python -m tools.t4_prof variant N B"""
import sys
import numpy as np, torch
import mahi_mpc_b200.api as api
from mahi_mpc_b200 import workload
v = int(sys.argv[1], 0); N = int(sys.argv[2]); B = int(sys.argv[3]); dev = torch.device("cuda", 0)
x0, up, traj = workload.synthetic_batch(4, 2, N, 0.002, B)
Q, R, Rm = workload.default_weights(4, 2)
m = api.Model("double_pendulum", N=N)
s = api.Solver(m, B, warm_start=0, kernel_variant=v); s.set_weights(Q, R, Rm)
d = [torch.from_numpy(a).to(dev) for a in (x0, up, traj)]
V = torch.empty(B, m.NV, dtype=torch.float64, device=dev); st = torch.empty(B, 5, dtype=torch.float64, device=dev)
for rep in range(2):
    print(f"---- solve {rep}", flush=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s.solve(*d, V_out=V, stats_out=st, want_lam=False); b.record(); torch.cuda.synchronize()
    print(f"---- {a.elapsed_time(b):.3f} ms", flush=True)
