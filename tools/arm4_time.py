"""Developer tool: device-resident timing of the 4-DOF arm (BASELINE config 3 shape) on one GPU.  python -m tools.arm4_time B [variant]"""
import sys
import numpy as np, torch
import mahi_mpc_b200.api as api
from mahi_mpc_b200 import workload
B = int(sys.argv[1]); v = int(sys.argv[2], 0) if len(sys.argv) > 2 else 0
N, nx, nu = 50, 8, 4
x0, up, traj = workload.synthetic_batch(nx, nu, N, 0.002, B); traj *= 0.5
Q, R, Rm = workload.default_weights(nx, nu)
m = api.Model("arm4", N=N)
s = api.Solver(m, B, warm_start=0, kernel_variant=v); s.set_weights(Q, R, Rm)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x0, up, traj)]
V = torch.empty(B, m.NV, dtype=torch.float64, device=dev); st = torch.empty(B, 5, dtype=torch.float64, device=dev)
for _ in range(2): s.solve(*d, V_out=V, stats_out=st, want_lam=False)
ts = []
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s.solve(*d, V_out=V, stats_out=st, want_lam=False); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
it = api.stats_from_tensor(st)
print(f"arm4 N=50 B={B} var={v}: {np.median(ts):.2f} ms  {B/np.median(ts)*1e3:.3e} solves/s  iters mean {it['iterations'].mean():.2f} conv {(it['status']==0).mean():.3f}")
