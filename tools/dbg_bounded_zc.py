"""Diagnostic: bounded (T3) solves at B = 4096 -- device-resident twice, then page-locked host buffers, compared bit for bit."""
import sys
import numpy as np, torch
import mahi_mpc_b200.api as api
from mahi_mpc_b200 import workload as w
N, B = 50, 4096
from oracle import nlp
P = nlp.Problem(N=N)
Q, R, Rm = nlp.default_weights(4, 2)
x0, up, traj = nlp.synthetic_batch(4, 2, N, P.h, B, offset=0)
print("h", P.h)
WANT_LAM = len(sys.argv) > 1
m = api.Model("double_pendulum", N=N)
s = api.Solver(m, B, tol=1e-8, warm_start=0); s.set_weights(Q, R, Rm)
s.set_bounds(u_min=[-2.0, -2.0], u_max=[2.0, 2.0])
d = [torch.from_numpy(a).cuda() for a in (x0, up, traj)]
outs = []
for rep in range(3):
    V, lam, st = s.solve(*d, want_lam=WANT_LAM); torch.cuda.synchronize()
    outs.append((V.cpu().clone(), api.stats_from_tensor(st).copy()))
for rep in (1, 2):
    same = torch.equal(outs[0][0], outs[rep][0])
    bad = (outs[0][0] != outs[rep][0]).any(dim=1).nonzero().flatten().tolist()
    print(f"device solve 0 vs {rep}: identical {same}; differing instances {bad[:10]} status {outs[0][1]['status'][bad[:10]]} iters {outs[0][1]['iterations'][bad[:10]]} / {outs[rep][1]['iterations'][bad[:10]]}")
print("status hist", np.bincount(outs[0][1]["status"]), "iters hist", np.bincount(outs[0][1]["iterations"]))
h = [torch.from_numpy(a).pin_memory() for a in (x0, up, traj)]
Vh = torch.empty(B, m.NV, dtype=torch.float64).pin_memory(); sth = torch.empty(B, 5, dtype=torch.float64).pin_memory()
lamh = torch.empty(B, m.M, dtype=torch.float64).pin_memory() if WANT_LAM else None
for rep in range(2):
    s.solve(*h, V_out=Vh, lam_out=lamh, stats_out=sth, want_lam=WANT_LAM)
    bad = (outs[0][0] != Vh).any(dim=1).nonzero().flatten().tolist()
    a = api.stats_from_tensor(sth)
    print(f"pinned solve {rep} vs device 0: identical {len(bad) == 0}; differing instances {len(bad)}: {bad[:10]} status {a['status'][bad[:10]]} iters {a['iterations'][bad[:10]]} vs {outs[0][1]['iterations'][bad[:10]]}")
Vp, _, stp = s.solve(x0, up, traj, want_lam=False)
bad = (outs[0][0].numpy() != Vp).any(axis=1).nonzero()[0].tolist()
print(f"pageable vs device 0: differing instances {len(bad)}: {bad[:10]}")
