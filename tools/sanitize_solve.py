"""One small batched solve per kernel family (unbounded + bounded), meant to be run under compute-sanitizer:

    compute-sanitizer --tool racecheck  python tools/sanitize_solve.py
    compute-sanitizer --tool initcheck  python tools/sanitize_solve.py
    compute-sanitizer --tool memcheck   python tools/sanitize_solve.py

Exits non-zero if a solve does not converge; the sanitizer's own summary is the evidence (profiles/r2_sanitizer_*.txt)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mahi_mpc_b200 import api, workload  # noqa: E402


def main():
    variants = [int(v, 0) for v in sys.argv[1:]] or [3 + 256 * 2, 3 + 256 * 1]
    N, B = 25, 40  # ragged: the second tile of the wide configuration is partly empty
    h = 0.002
    Q, R, Rm = workload.default_weights(4, 2)
    x0, up, traj = workload.synthetic_batch(4, 2, N, h, B)
    m = api.Model("double_pendulum", N=N)
    for variant in variants:
        for bounded in (False, True):
            s = api.Solver(m, B, tol=1e-9, kernel_variant=variant)
            s.set_weights(Q, R, Rm)
            if bounded:
                s.set_bounds(u_min=[-2.5, -2.5], u_max=[2.5, 2.5])
            V, lam, st = s.solve(x0, up, traj)
            ok = bool((st["status"] == 0).all())
            print(f"variant {variant:#x} bounded={bounded}: converged={ok} iterations {st['iterations'].min()}-{st['iterations'].max()}", flush=True)
            if not ok:
                sys.exit(1)
            s.close()


if __name__ == "__main__":
    main()
