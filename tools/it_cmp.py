"""Developer tool: compare the iterates of two kernel variants after k Newton iterations (max_iter = k)."""
import numpy as np, sys
import mahi_mpc_b200.api as api
from mahi_mpc_b200 import workload
N = 50; B = 16; va, vb = int(sys.argv[1]), int(sys.argv[2])
Q, R, Rm = workload.default_weights(4, 2)
m = api.Model("double_pendulum", N=N)
x0, up, traj = workload.synthetic_batch(4, 2, N, 0.002, B)
for mi in range(0, 5):
    out = []
    for v in (va, vb):
        s = api.Solver(m, B, warm_start=0, kernel_variant=v, max_iter=mi); s.set_weights(Q, R, Rm)
        out.append(s.solve(x0, up, traj))
    (Va, la, sa), (Vb, lb, sb) = out
    print("max_iter", mi, "dV %.3e dlam %.3e" % (np.abs(Va - Vb).max(), np.abs(la - lb).max()), "kkt a", sa["kkt_error"][:3], "kkt b", sb["kkt_error"][:3], "it", sa["iterations"][:3], sb["iterations"][:3])
