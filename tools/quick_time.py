"""Developer tool: quick device timing of the batched solve and node evaluation (not the bench)."""
import sys
import time

import numpy as np

import mahi_mpc_b200.api as api
from oracle import nlp

N = int(sys.argv[1]) if len(sys.argv) > 1 else 50
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
variant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
P = nlp.Problem(N=N)
Q, R, Rm = nlp.default_weights(4, 2)
x0, up, traj = nlp.synthetic_batch(4, 2, N, P.h, B)
m = api.Model("double_pendulum", N=N)
s = api.Solver(m, B, warm_start=0, kernel_variant=variant)
s.set_weights(Q, R, Rm)
for i in range(5):
    t0 = time.perf_counter()
    V, lam, st = s.solve(x0, up, traj)
    t1 = time.perf_counter()
    tm = s.last_timing()
    print(f"N={N} B={B} var={variant} wall {1e3*(t1-t0):.3f} ms  kernel {tm['kernel_ms']:.3f} ms  h2d {tm['h2d_ms']:.3f} d2h {tm['d2h_ms']:.3f}  "
          f"iters {st['iterations'].min()}-{st['iterations'].max()} mean {st['iterations'].mean():.2f} conv {(st['status']==0).mean():.3f} "
          f"ls {st['ls_steps'].sum()} reg {st['reg_steps'].sum()} -> {B/tm['kernel_ms']*1e3:.3e} solves/s (kernel)")
