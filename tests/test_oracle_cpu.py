"""CPU suite: pins the oracle (numpy + C restatement) against every anchor the reference offers for
this path -- it has no tests of its own, so the anchors are:
  * the reference's own hand-expanded closed forms old/Models/DoublePendulumModel.hpp:15-132,
    compiled in place into oracle/_ref (when /root/reference exists) and frozen in tests/golden/dp_nodes.json,
  * 50-digit mpmath evaluation of the model expression of examples/ex_model_generate.cpp:36-37,
  * hand-derivable known answers,
  * an unrelated solver (scipy SLSQP) for one converged solution, frozen in tests/golden/dp_solves.json."""
import os

import numpy as np
import pytest

from oracle import models, nlp


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(np.asarray(b)).max())


def test_factored_model_equals_reference_expression():
    import sympy as sp
    m = models.double_pendulum_sym()
    ref = models.double_pendulum_expanded_reference_text()
    for a, b in zip(m.f, ref):
        assert sp.simplify(a - b) == 0


def test_c_oracle_nodes_vs_golden(coracle, golden):
    g_ = golden["nodes"]
    x, u, w = np.array(g_["x"]), np.array(g_["u"]), np.array(g_["w"])
    h = 0.002
    g, Fx, Fu, H = coracle.node_eval("double_pendulum", h, x, u, np.zeros_like(x), w)
    f = (g - x) / h
    A = (Fx - np.eye(4)) / h
    B = Fu / h
    for src in ("ref_old", "mp"):
        assert rel(f, g_[src]["f"]) < 5e-12
        assert rel(A, g_[src]["A"]) < 5e-12
        assert rel(B, g_[src]["B"]) < 5e-12
    assert rel(H / h, g_["mp"]["Hc"]) < 1e-12
    np.testing.assert_allclose(f[0], g_["known"]["f00"], atol=1e-13)
    np.testing.assert_allclose(B[0][2:], g_["known"]["B00_rows23"], atol=1e-12)
    np.testing.assert_allclose(f[1][2:], g_["known"]["qdd_survey"], rtol=1e-12)


def test_numpy_oracle_nodes_vs_golden(golden):
    g_ = golden["nodes"]
    x, u, w = np.array(g_["x"]).T, np.array(g_["u"]).T, np.array(g_["w"]).T
    m = models.numeric("double_pendulum")
    assert rel(m["f"](x, u).T, g_["mp"]["f"]) < 1e-13
    assert rel(np.moveaxis(m["A"](x, u), 2, 0), g_["ref_old"]["A"]) < 1e-13
    assert rel(np.moveaxis(m["B"](x, u), 2, 0), g_["ref_old"]["B"]) < 1e-13
    assert rel(np.moveaxis(m["Hc"](x, u, w), 2, 0), g_["mp"]["Hc"]) < 1e-13


@pytest.mark.skipif(not os.path.exists("/root/reference/old/Models/DoublePendulumModel.hpp"), reason="reference tree absent")
def test_ref_old_library_matches_fixture(coracle, golden):
    """oracle/_ref (the reference's own code, rebuilt here) reproduces the committed fixture bit for bit."""
    from oracle import coracle as co
    g_ = golden["nodes"]
    f, A, B = co.RefOld().eval(np.array(g_["x"]), np.array(g_["u"]))
    assert np.array_equal(f, np.array(g_["ref_old"]["f"]))
    assert np.array_equal(A, np.array(g_["ref_old"]["A"]))
    assert np.array_equal(B, np.array(g_["ref_old"]["B"]))


def test_layouts_match_reference_definition():
    """V, p, g sizes (ModelGenerator.cpp:61,129-143) and SURVEY section 8 table."""
    for N, NV, M, np_ in [(25, 154, 100, 110), (50, 304, 200, 210), (10, 64, 40, 50), (100, 604, 400, 410), (200, 1204, 800, 810)]:
        P = nlp.Problem(N=N)
        assert (P.NV, P.M, P.np_) == (NV, M, np_)
    P = nlp.Problem(N=25, is_linear=True)
    assert P.np_ == 110 + 16 + 8 + 4 + 4
    # jac_g structural pattern: 20 nnz per node
    P = nlp.Problem(N=5)
    colind, rows = P.jac_g_ccs()
    assert rows.size == 20 * 5 and colind.size == P.NV + 1


def test_nlp_derivatives_by_finite_differences():
    rng = np.random.default_rng(3)
    for lin in (False, True):
        P = nlp.Problem(N=6, is_linear=lin)
        Q, R, Rm = nlp.default_weights(4, 2)
        x0, _, traj = nlp.synthetic_batch(4, 2, P.N, P.h, 1)
        ui = np.array([0.3, -0.4])
        p = P.pack_p(traj[0], Q, R, Rm, ui, P.linearisation(x0[0], ui) if lin else None)
        V = rng.uniform(-0.5, 0.5, P.NV)
        lam = rng.uniform(-1, 1, P.M)
        f, gf = P.nlp_grad_f(V, p)
        g, J = P.nlp_jac_g(V, p)
        H = P.nlp_hess_l(V, p, 1.0, lam)
        eps = 1e-6
        for i in rng.choice(P.NV, 12, replace=False):
            d = np.zeros(P.NV); d[i] = eps
            assert abs((P.nlp_f(V + d, p) - P.nlp_f(V - d, p)) / (2 * eps) - gf[i]) < 1e-5 * max(1, abs(gf[i]))
            assert np.abs((P.nlp_g(V + d, p) - P.nlp_g(V - d, p)) / (2 * eps) - J[:, i]).max() < 1e-6
            gl = lambda v: P.nlp_grad_f(v, p)[1] + P.nlp_jac_g(v, p)[1].T @ lam
            assert np.abs((gl(V + d) - gl(V - d)) / (2 * eps) - H[:, i]).max() < 2e-4 * max(1, np.abs(H[:, i]).max())


def test_c_oracle_nlp_vs_numpy(coracle):
    rng = np.random.default_rng(4)
    for N, lin in [(1, False), (7, False), (25, True)]:
        P = nlp.Problem(N=N, is_linear=lin)
        Q, R, Rm = nlp.default_weights(4, 2)
        x0, _, traj = nlp.synthetic_batch(4, 2, N, P.h, 1)
        ui = rng.uniform(-1, 1, 2)
        p = P.pack_p(traj[0], Q, R, Rm, ui, P.linearisation(x0[0], ui) if lin else None)
        V = rng.uniform(-0.5, 0.5, P.NV); lam = rng.uniform(-1, 1, P.M)
        o = coracle.nlp_eval("double_pendulum", N, P.h, lin, V, p, lam, 0.6)
        f, gf = P.nlp_grad_f(V, p)
        g, J = P.nlp_jac_g(V, p)
        H = P.nlp_hess_l(V, p, 0.6, lam)
        assert abs(o["f"] - f) < 1e-12 * abs(f) and rel(o["g"], g) < 1e-13 and rel(o["grad_f"], gf) < 1e-13
        for k in range(N):
            assert rel(o["Hzz"][k], H[k * 6:(k + 1) * 6, k * 6:(k + 1) * 6]) < 1e-12
            assert rel(o["Fx"][k], J[k * 4:(k + 1) * 4, k * 6:k * 6 + 4]) < 1e-13


def test_oracle_solves_vs_golden(coracle, golden):
    gs = golden["solves"]
    Q, R, Rm = np.array(gs["Q"]), np.array(gs["R"]), np.array(gs["Rm"])
    for case in gs["cases"]:
        N = case["N"]
        P = nlp.Problem(N=N)
        x0, p = np.array(case["x0"]), np.array(case["p"])
        lb = ub = None
        if case["u_bound"]:
            lb, ub = nlp.bounds_vectors(P, x0, u_min=[-case["u_bound"]] * 2, u_max=[case["u_bound"]] * 2)
        Vc, lamc, st, bad = coracle.solve_batch("double_pendulum", N, P.h, False, x0[None], p[None], lb=lb, ub=ub,
                                                tol=1e-11 if case["u_bound"] else 1e-12)
        assert bad == 0, st
        assert np.abs(Vc[0] - np.array(case["V"])).max() < 1e-7
        assert abs(st["f"][0] - case["f"]) < 1e-9 * abs(case["f"])
        if "slsqp_f" in case:  # unrelated solver agrees
            assert abs(case["slsqp_f"] - case["f"]) < 1e-10 * abs(case["f"]) and case["slsqp_V_maxdiff"] < 1e-9
    # numpy oracle on the first case (slow path, once)
    case = gs["cases"][0]
    P = nlp.Problem(N=case["N"])
    r = nlp.solve(P, np.array(case["x0"]), np.array(case["p"]), tol=1e-12)
    assert r.status == 0 and np.abs(r.V - np.array(case["V"])).max() < 1e-10


def test_survey_scratch_values():
    """BASELINE.md section 5 / SURVEY section 6: zero start, +-sin(2 pi t) reference."""
    Q, R, Rm = nlp.default_weights(4, 2)
    for N, fstar, u0 in [(10, 4.012520669434680e3, [0.55741378, -1.29984066]), (50, 1.496847277541389e4, [2.2361408, -5.12057289])]:
        P = nlp.Problem(N=N)
        t = np.arange(N) * P.h
        s, c = np.sin(2 * np.pi * t), np.cos(2 * np.pi * t)
        p = P.pack_p(np.stack([s, -s, 2 * np.pi * c, -2 * np.pi * c], 1), Q, R, Rm, np.zeros(2))
        r = nlp.solve(P, np.zeros(4), p, tol=1e-10)
        assert r.status == 0 and abs(r.f - fstar) < 1e-9 * fstar
        np.testing.assert_allclose(r.V[4:6], u0, atol=1e-7)


def test_synthetic_batch_is_shardable():
    a = nlp.synthetic_batch(4, 2, 50, 0.002, 64)
    b = nlp.synthetic_batch(4, 2, 50, 0.002, 16, offset=32)
    for x, y in zip(a, b):
        assert np.array_equal(x[32:48], y)
    assert np.abs(a[0][:, :2]).max() <= 0.5 and np.abs(a[0][:, 2:]).max() <= 1.0


def test_rk4_hessian_oracle_vs_mpmath_golden():
    """The oracle's RK4 x M interval map (oracle/nlp.py: step_jac, complex-step step_hess) against 60-digit mpmath values
    (tests/golden/rk4_hessian.json; integrator of workingMPC_8_30_21.m:46-59): pins the oracle for SURVEY 8 f-4."""
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "rk4_hessian.json")) as fh:
        cases = json.load(fh)["cases"]
    assert len(cases) >= 9
    for c in cases:
        P = nlp.Problem(N=2, h=c["h"], integrator="rk4", substeps=c["substeps"])
        x, u, w = np.array([c["x"]]), np.array([c["u"]]), np.array([c["w"]])
        F, Fx, Fu = P.step_jac(x, u, {})
        H = P.step_hess(x, u, {}, w)[0]
        Fz = np.concatenate([Fx, Fu], axis=2)[0]
        assert np.abs(F[0] - c["F"]).max() < 1e-14
        assert np.abs(Fz - np.array(c["Fz"])).max() <= 1e-13 * np.abs(c["Fz"]).max()
        assert np.abs(H - np.array(c["H"])).max() <= 1e-13 * np.abs(c["H"]).max()
        # Gauss-Newton option: no integrator second derivatives
        Pg = nlp.Problem(N=2, h=c["h"], integrator="rk4", substeps=c["substeps"], hessian="gauss_newton")
        assert not Pg.step_hess(x, u, {}, w).any()


def test_node_cost_oracle_derivatives_by_finite_differences():
    """oracle/nlp.py with cost = "node" (the MATLAB prototypes' objective, workingMPC_8_30_21.m:90-106): gradient and
    Lagrangian Hessian against central finite differences, Euler and RK4 x 2; a solve converges to a KKT point."""
    for integ, sub in (("euler", 1), ("rk4", 2)):
        P = nlp.Problem(N=6, h=0.05, integrator=integ, substeps=sub, cost="node")
        Q, R, Rm = nlp.default_weights(4, 2)
        x0, up, traj = nlp.synthetic_batch(4, 2, 6, 0.05, 1)
        p = P.pack_p(traj[0], Q, R, Rm, up[0])
        rng = np.random.default_rng(0)
        V = rng.uniform(-.5, .5, P.NV); lam = rng.uniform(-1, 1, P.M)
        f, g = P.nlp_grad_f(V, p)
        assert abs(f - P.nlp_f(V, p)) < 1e-12 * abs(f)
        eps = 1e-6

        def grad_l(Vq):
            _, gq = P.nlp_grad_f(Vq, p)
            _, J = P.nlp_jac_g(Vq, p)
            return gq + J.T @ lam
        H = P.nlp_hess_l(V, p, 1.0, lam)
        for i in range(P.NV):
            Vp, Vm = V.copy(), V.copy()
            Vp[i] += eps; Vm[i] -= eps
            assert abs((P.nlp_f(Vp, p) - P.nlp_f(Vm, p)) / (2 * eps) - g[i]) < 1e-6
            assert np.abs((grad_l(Vp) - grad_l(Vm)) / (2 * eps) - H[:, i]).max() < 1e-6
        r = nlp.solve(P, x0[0], p, tol=1e-10, max_iter=100)
        assert r.status == 0 and r.kkt_error < 1e-10
