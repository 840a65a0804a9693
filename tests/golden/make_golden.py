"""Regenerates the committed golden fixtures (run in the build container, where /root/reference exists):

    python -m tests.golden.make_golden

  dp_nodes.json   64 random shooting nodes of the double pendulum with
                    * f, A, B from the REFERENCE'S OWN closed forms (old/Models/DoublePendulumModel.hpp:15-132,
                      compiled in place into oracle/_ref/libdp_old_ref.so)            -> "ref_old"
                    * f, A, B, Hc from 50-digit mpmath evaluation of the sympy model  -> "mp"
                  plus the hand-derivable known answers of SURVEY.md section 4.
  dp_solves.json  converged NLP solutions (numpy oracle, tol 1e-12) for N = 10, 25, 50 on seeded inputs,
                  the N=10 one cross-checked here against scipy SLSQP (an unrelated solver).
The fixtures travel to the GPU box; /root/reference does not.
"""
import json
import os

import numpy as np

from oracle import coracle, models, nlp

HERE = os.path.dirname(os.path.abspath(__file__))


def nodes():
    rng = np.random.default_rng(20261018)
    n = 64
    x = np.c_[rng.uniform(-np.pi, np.pi, (n, 2)), rng.uniform(-5, 5, (n, 2))]
    u = rng.uniform(-10, 10, (n, 2))
    w = rng.uniform(-2, 2, (n, 4))
    x[0] = 0; u[0] = 0  # known answer: f = [0,0,-9.81,9.81], B[2:,:] = [[1,-2],[-2,5]]
    x[1] = [0.3, -0.7, 1.1, -0.9]; u[1] = [2.5, -1.5]  # SURVEY 8 a-1 20-digit value
    ref = coracle.RefOld()
    f_r, A_r, B_r = ref.eval(x, u)
    mp = [models.mp_eval("double_pendulum", x[i], u[i], w[i]) for i in range(n)]
    out = dict(x=x.tolist(), u=u.tolist(), w=w.tolist(),
               ref_old=dict(f=f_r.tolist(), A=A_r.tolist(), B=B_r.tolist()),
               mp=dict(f=[m["f"].tolist() for m in mp], A=[m["A"].tolist() for m in mp],
                       B=[m["B"].tolist() for m in mp], Hc=[m["Hc"].tolist() for m in mp]),
               known=dict(f00=[0.0, 0.0, -9.81, 9.81], B00_rows23=[[1.0, -2.0], [-2.0, 5.0]],
                          qdd_survey=[-5.1643145815526023, -0.64190470767822423]))
    with open(os.path.join(HERE, "dp_nodes.json"), "w") as fh:
        json.dump(out, fh)
    print("dp_nodes.json: ref_old vs mp  f", np.abs(f_r - np.array(out["mp"]["f"])).max(),
          "A", np.abs(A_r - np.array(out["mp"]["A"])).max(), "B", np.abs(B_r - np.array(out["mp"]["B"])).max())


def solves():
    from scipy.optimize import minimize

    Q, R, Rm = nlp.default_weights(4, 2)
    cases = []
    for N, seed_off, u_init, bounded in [(10, 0, [0.0, 0.0], None), (25, 1, [0.2, -0.1], None), (50, 2, [0.0, 0.0], None),
                                         (25, 3, [0.0, 0.0], 3.0), (50, 4, [0.1, 0.1], 2.0)]:
        P = nlp.Problem(N=N)
        x0, _, traj = nlp.synthetic_batch(4, 2, N, P.h, 1, offset=seed_off)
        p = P.pack_p(traj[0], Q, R, Rm, np.array(u_init))
        lb = ub = None
        if bounded:
            lb, ub = nlp.bounds_vectors(P, x0[0], u_min=[-bounded] * 2, u_max=[bounded] * 2)
        r = nlp.solve(P, x0[0], p, tol=1e-12 if not bounded else 1e-11, lb=lb, ub=ub)
        assert r.status == 0, (N, r)
        case = dict(N=N, h=P.h, x0=x0[0].tolist(), p=p.tolist(), u_bound=bounded, V=r.V.tolist(),
                    lam_g=r.lam_g.tolist(), f=r.f, iters=r.iters)
        if N == 10:  # independent solver cross-check
            cons = dict(type="eq", fun=lambda v: P.nlp_g(np.r_[x0[0], v], p),
                        jac=lambda v: P.nlp_jac_g(np.r_[x0[0], v], p)[1][:, 4:])
            s = minimize(lambda v: P.nlp_grad_f(np.r_[x0[0], v], p)[0], np.zeros(P.NV - 4),
                         jac=lambda v: P.nlp_grad_f(np.r_[x0[0], v], p)[1][4:], constraints=[cons],
                         method="SLSQP", options=dict(ftol=1e-15, maxiter=500))
            case["slsqp_f"] = float(s.fun)
            case["slsqp_V_maxdiff"] = float(np.abs(s.x - r.V[4:]).max())
            print("SLSQP cross-check N=10: f rel", abs(s.fun - r.f) / r.f, "V maxdiff", case["slsqp_V_maxdiff"])
        cases.append(case)
        print(f"N={N} bounded={bounded} iters={r.iters} f={r.f!r} u0={r.V[4:6]}")
    with open(os.path.join(HERE, "dp_solves.json"), "w") as fh:
        json.dump(dict(Q=Q.tolist(), R=R.tolist(), Rm=Rm.tolist(), cases=cases), fh)


if __name__ == "__main__":
    nodes()
    solves()
