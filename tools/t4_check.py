"""Developer tool: quick parity of one kernel variant against the oracle on a small batch (run with a hard timeout: a new
kernel may hang).   python -m tools.t4_check variant N B [tol]"""
import sys
import numpy as np
import mahi_mpc_b200.api as api
from oracle import coracle, nlp
v = int(sys.argv[1], 0); N = int(sys.argv[2]); B = int(sys.argv[3]); tol = float(sys.argv[4]) if len(sys.argv) > 4 else 1e-10
P = nlp.Problem(N=N); Q, R, Rm = nlp.default_weights(4, 2)
x0, up, traj = nlp.synthetic_batch(4, 2, N, P.h, B)
p = np.stack([P.pack_p(traj[b], Q, R, Rm, up[b]) for b in range(B)])
co = coracle.COracle("O3")
Vo, lamo, sto, bad = co.solve_batch("double_pendulum", N, P.h, False, x0, p, tol=tol)
m = api.Model("double_pendulum", N=N)
s = api.Solver(m, B, tol=tol, kernel_variant=v, warm_start=0); s.set_weights(Q, R, Rm)
V, lam, st = s.solve(x0, up, traj)
print(f"variant {v:#x} N={N} B={B}: status {np.bincount(st['status'])} iters {np.bincount(st['iterations'])} oracle iters {np.bincount(sto['iters'])} "
      f"max|V-Vo| {np.abs(V - Vo).max():.3e} max|lam-lamo| {np.abs(lam - lamo).max():.3e} obj rel {np.abs(st['objective'] / sto['f'] - 1).max():.3e}", flush=True)
