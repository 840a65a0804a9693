"""Developer tool: device-resident timing (CUDA events, inputs in HBM) of the batched solve for several batch sizes.
    [MMPC_RK4=substeps] python -m tools.dev_time N variant B1 B2 ..."""
import os, sys
import numpy as np, torch
import mahi_mpc_b200.api as api
from mahi_mpc_b200 import workload
N = int(sys.argv[1]); variant = int(sys.argv[2], 0); dev = torch.device("cuda", 0)
for B in [int(v) for v in sys.argv[3:]]:
    x0, up, traj = workload.synthetic_batch(4, 2, N, 0.002, B)
    Q, R, Rm = workload.default_weights(4, 2)
    m = api.Model("double_pendulum", N=N); rk = int(os.environ.get("MMPC_RK4", "0"))
    kw = dict(integrator=api.RK4, substeps=rk) if rk else {}
    s = api.Solver(m, B, warm_start=0, kernel_variant=variant, **kw)
    if os.environ.get("MMPC_UBOUND"):
        ub = float(os.environ["MMPC_UBOUND"]); s.set_bounds(u_min=[-ub, -ub], u_max=[ub, ub])
    s.set_weights(Q, R, Rm)
    d = [torch.from_numpy(a).to(dev) for a in (x0, up, traj)]
    V = torch.empty(B, m.NV, dtype=torch.float64, device=dev); st = torch.empty(B, 5, dtype=torch.float64, device=dev)
    for _ in range(3): s.solve(*d, V_out=V, stats_out=st, want_lam=False)
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); s.solve(*d, V_out=V, stats_out=st, want_lam=False); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ms = float(np.median(ts)); it = api.stats_from_tensor(st)["iterations"]
    print(f"N={N} var={variant} B={B:6d} {ms:8.3f} ms  {B/ms*1e3:.3e} solves/s  iters max {it.max()} mean {it.mean():.2f}")
