"""Developer tool: one T4 solve with a given max_iter (bisects which sweep misbehaves); prints status/iters, no oracle.
   python -m tools.t4_dbg variant N B max_iter"""
import sys
import numpy as np
import mahi_mpc_b200.api as api
from mahi_mpc_b200 import workload
v = int(sys.argv[1], 0); N = int(sys.argv[2]); B = int(sys.argv[3]); mi = int(sys.argv[4])
Q, R, Rm = workload.default_weights(4, 2)
x0, up, traj = workload.synthetic_batch(4, 2, N, 0.002, B)
m = api.Model("double_pendulum", N=N)
s = api.Solver(m, B, tol=1e-10, kernel_variant=v, warm_start=0, max_iter=mi); s.set_weights(Q, R, Rm)
V, lam, st = s.solve(x0, up, traj)
print(f"variant {v:#x} N={N} B={B} max_iter={mi}: status {st['status'][:6]} iters {st['iterations'][:6]} kkt {st['kkt_error'][:3]} V finite {np.isfinite(V).all()}", flush=True)
