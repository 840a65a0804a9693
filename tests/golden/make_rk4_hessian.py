"""Generates tests/golden/rk4_hessian.json: the RK4 x M interval map F of the double pendulum (workingMPC_8_30_21.m:46-59:
classical RK4, `number_of_finite_elements` sub-steps per shooting interval), its Jacobian and the contraction
sum_i w_i d2F_i/dz2 over z = (x, u), evaluated with mpmath at 60 digits (derivatives by mpmath's high-order numerical
differentiation of the 60-digit rollout: independent of every differentiation code in this repository).
    python tests/golden/make_rk4_hessian.py          (needs sympy + mpmath; ~1 minute)"""
import json
import os
import sys

import mpmath as mp
import numpy as np
import sympy as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import models  # noqa: E402  (the sympy statement of the model: oracle/models.py)

mp.mp.dps = 60
mdl = models.derive("double_pendulum")["model"]
f_mp = sp.lambdify(mdl.z, list(mdl.f), "mpmath")


def rk4(z, h, substeps):
    x, u = list(z[:4]), list(z[4:])
    hs = mp.mpf(h) / substeps
    for _ in range(substeps):
        k1 = f_mp(*x, *u)
        k2 = f_mp(*[a + hs / 2 * b for a, b in zip(x, k1)], *u)
        k3 = f_mp(*[a + hs / 2 * b for a, b in zip(x, k2)], *u)
        k4 = f_mp(*[a + hs * b for a, b in zip(x, k3)], *u)
        x = [a + hs / 6 * (p + 2 * q + 2 * r + s) for a, p, q, r, s in zip(x, k1, k2, k3, k4)]
    return x


def main():
    rng = np.random.default_rng(2021)
    cases = []
    for substeps in (1, 2, 4):
        for _ in range(3):
            x = rng.uniform(-1.5, 1.5, 4); u = rng.uniform(-4, 4, 2); w = rng.uniform(-1, 1, 4)
            h = 0.1
            z0 = [mp.mpf(float(v)) for v in list(x) + list(u)]
            F = rk4(z0, h, substeps)
            J = [[None] * 6 for _ in range(4)]
            for i in range(4):
                for j in range(6):
                    order = tuple(1 if q == j else 0 for q in range(6))
                    J[i][j] = mp.diff(lambda *z: rk4(list(z), h, substeps)[i], tuple(z0), order)
            phi = lambda *z: sum(mp.mpf(float(wi)) * Fi for wi, Fi in zip(w, rk4(list(z), h, substeps)))  # noqa: E731
            H = [[None] * 6 for _ in range(6)]
            for a in range(6):
                for b in range(a, 6):
                    order = tuple((1 if q == a else 0) + (1 if q == b else 0) for q in range(6))
                    H[a][b] = H[b][a] = mp.diff(phi, tuple(z0), order)
            cases.append(dict(h=h, substeps=substeps, x=x.tolist(), u=u.tolist(), w=w.tolist(), F=[float(v) for v in F],
                              Fz=[[float(v) for v in row] for row in J], H=[[float(v) for v in row] for row in H]))
            print(substeps, "done", flush=True)
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rk4_hessian.json"), "w") as fh:
        json.dump(dict(model="double_pendulum", digits=60, cases=cases), fh, indent=1)


if __name__ == "__main__":
    main()
