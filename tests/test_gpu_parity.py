"""Parity tests proper: the CUDA path, called through the C ABI, against the oracle and the golden
fixtures.  Tolerances are north_star's: 1e-12 relative for rollout/Jacobian/Hessian evaluations,
1e-6 on controls and 1e-8 relative on the objective for converged solves."""
import os

import numpy as np
import pytest

from oracle import nlp

pytestmark = pytest.mark.gpu

EVAL_RTOL = 1e-12   # north_star: "within 1e-12 relative error in fp64"
U_ATOL = 1e-6       # north_star: "within 1e-6 on controls"
OBJ_RTOL = 1e-8     # north_star: "1e-8 relative on the objective"
NB = 10e30      # the reference's "no bound" sentinel (ModelParameters.cpp:21-24)


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(b).max())


def rel_rows(a, b, axis=0):
    """Relative error of every ENTRY POSITION against that position's own magnitude over the sample axis (max-norm per
    position): an h*A entry of size 1e-3 is judged against 1e-3, not against the 1 on the diagonal of Fx = I + h A."""
    a, b = np.asarray(a), np.asarray(b)
    scale = np.abs(b).max(axis=axis, keepdims=True)
    err = np.abs(a - b).max(axis=axis, keepdims=True)
    nz = scale > 0
    assert np.all(err[~nz] == 0.0)  # structural zeros are exact zeros
    return (err[nz] / scale[nz]).max() if nz.any() else 0.0


@pytest.fixture(scope="module")
def api(native_lib):
    from mahi_mpc_b200 import api as a
    assert a.device_count() >= 1
    return a


def random_nodes(n, seed):
    rng = np.random.default_rng(seed)
    x = np.c_[rng.uniform(-np.pi, np.pi, (n, 2)), rng.uniform(-5, 5, (n, 2))]
    u = rng.uniform(-10, 10, (n, 2))
    xn = rng.uniform(-3, 3, (n, 4))
    w = rng.uniform(-2, 2, (n, 4))
    return x, u, xn, w


def test_eval_nodes_vs_golden(api, golden):
    g_ = golden["nodes"]
    x, u, w = np.array(g_["x"]), np.array(g_["u"]), np.array(g_["w"])
    n = x.shape[0]
    h = 0.002
    m = api.Model("double_pendulum", N=25)
    xn = np.zeros_like(x)
    g, Fx, Fu, H = m.eval_nodes(x.T.copy(), u.T.copy(), xn.T.copy(), w.T.copy())
    f_gpu = (g.T - x) / h
    A_gpu = (Fx.T.reshape(n, 4, 4) - np.eye(4)) / h
    B_gpu = Fu.T.reshape(n, 4, 2) / h
    for src in ("ref_old", "mp"):  # the reference's own closed forms, and 50-digit mpmath
        assert rel(f_gpu, g_[src]["f"]) < 5e-12  # (g - x)/h amplifies rounding of x + h f by 1/h
        assert rel(A_gpu, g_[src]["A"]) < 5e-12
        assert rel(B_gpu, g_[src]["B"]) < 5e-12
    Hm = np.array(g_["mp"]["Hc"])
    Hu = np.array([[Hm[i, a, b] for a in range(6) for b in range(a, 6)] for i in range(n)])
    assert rel(H.T / h, Hu) < EVAL_RTOL
    # known answers: f(0,0) = [0,0,-9.81,9.81]; B(0,0) rows 2,3 = [[1,-2],[-2,5]]
    np.testing.assert_allclose(f_gpu[0], g_["known"]["f00"], atol=1e-12)
    np.testing.assert_allclose(B_gpu[0][2:], g_["known"]["B00_rows23"], atol=1e-12)
    np.testing.assert_allclose(f_gpu[1][2:], g_["known"]["qdd_survey"], rtol=1e-11)


@pytest.mark.parametrize("n", [1, 31, 4097, 200_000])
def test_eval_nodes_vs_oracle(api, coracle, n):
    x, u, xn, w = random_nodes(n, 100 + n)
    m = api.Model("double_pendulum", N=25)
    g, Fx, Fu, H = m.eval_nodes(x.T.copy(), u.T.copy(), xn.T.copy(), w.T.copy())
    go, Fxo, Fuo, Ho = coracle.node_eval("double_pendulum", m.h, x, u, xn, w)
    assert rel(g.T, go) < EVAL_RTOL
    assert rel(Fx.T.reshape(n, 4, 4), Fxo) < EVAL_RTOL
    assert rel(Fu.T.reshape(n, 4, 2), Fuo) < EVAL_RTOL
    iu = np.triu_indices(6)
    assert rel(H.T, Ho[:, iu[0], iu[1]]) < EVAL_RTOL
    # entry-wise: h A = Fx - I, h B = Fu and every Hessian entry against ITS OWN scale (the diagonal ones of Fx hide
    # the 1e-3-sized h A entries in a whole-array norm).  Fx - I loses log10(1/(h |A|)) digits to the addition of 1 --
    # the bound is eps / (h |A|_pos) per position, i.e. still <= 1e-12 relative to the position's largest entry.
    if n >= 31:
        Fxg = Fx.T.reshape(n, 4, 4)
        assert rel_rows(Fxg - np.eye(4), Fxo - np.eye(4)) < 2e-12
        assert rel_rows(Fu.T.reshape(n, 4, 2), Fuo) < EVAL_RTOL
        assert rel_rows(H.T, Ho[:, iu[0], iu[1]]) < EVAL_RTOL
        assert rel_rows(g.T, go) < EVAL_RTOL
    # without w: H untouched, same g/Fx/Fu
    g2, Fx2, Fu2, H2 = m.eval_nodes(x.T.copy(), u.T.copy(), xn.T.copy())
    assert H2 is None and np.array_equal(g2, g) and np.array_equal(Fx2, Fx)


def test_eval_nodes_empty(api):
    m = api.Model("double_pendulum", N=25)
    z = np.zeros((4, 0))
    g, Fx, Fu, H = m.eval_nodes(z, np.zeros((2, 0)), z)
    assert g.shape == (4, 0)


def test_linearise_vs_oracle(api, coracle):
    x, u, _, _ = random_nodes(257, 5)
    m = api.Model("double_pendulum", N=25)
    xd, A, B = m.linearise(x, u)
    for i in (0, 100, 256):
        f, Ac, Bc = coracle.linearise("double_pendulum", x[i], u[i])
        assert rel(xd[i], f) < EVAL_RTOL and rel(A[i], Ac) < EVAL_RTOL and rel(B[i], Bc) < EVAL_RTOL


@pytest.mark.parametrize("N,is_linear", [(1, False), (10, False), (25, True), (50, False), (200, False)])
def test_nlp_eval_vs_oracle(api, coracle, N, is_linear):
    rng = np.random.default_rng(N)
    B = 5
    P = nlp.Problem(N=N, is_linear=is_linear)
    m = api.Model("double_pendulum", N=N, is_linear=is_linear)
    assert (m.NV, m.M, m.np) == (P.NV, P.M, P.np_)
    Q, R, Rm = nlp.default_weights(4, 2)
    x0, _, traj = nlp.synthetic_batch(4, 2, N, P.h, B)
    ps = []
    for b in range(B):
        ui = rng.uniform(-1, 1, 2)
        ps.append(P.pack_p(traj[b], Q, R, Rm, ui, P.linearisation(x0[b], ui) if is_linear else None))
    p = np.stack(ps)
    V = rng.uniform(-0.5, 0.5, (B, P.NV))
    lam = rng.uniform(-2, 2, (B, P.M))
    out = m.nlp_eval(V, p, lam, lam_f=0.7)
    for b in range(B):
        o = coracle.nlp_eval("double_pendulum", N, P.h, is_linear, V[b], p[b], lam[b], 0.7)
        assert abs(out["f"][b] - o["f"]) <= EVAL_RTOL * abs(o["f"])
        assert rel(out["g"][b], o["g"]) < EVAL_RTOL
        assert rel(out["grad_f"][b], o["grad_f"]) < EVAL_RTOL
        assert rel(out["Fx"][b], o["Fx"]) < EVAL_RTOL
        assert rel(out["Fu"][b], o["Fu"]) < EVAL_RTOL
        assert rel(out["Hzz"][b], o["Hzz"]) < EVAL_RTOL
    # numpy oracle (dense jac / hess in the reference's CCS-able layout) on instance 0
    f, gf = P.nlp_grad_f(V[0], p[0])
    _, J = P.nlp_jac_g(V[0], p[0])
    H = P.nlp_hess_l(V[0], p[0], 0.7, lam[0])
    s = 6
    for k in range(N):
        assert rel(out["Fx"][0, k], J[k * 4:(k + 1) * 4, k * s:k * s + 4]) < EVAL_RTOL
        assert rel(out["Fu"][0, k], J[k * 4:(k + 1) * 4, k * s + 4:(k + 1) * s]) < EVAL_RTOL
        assert rel(out["Hzz"][0, k], H[k * s:(k + 1) * s, k * s:(k + 1) * s]) < EVAL_RTOL
    assert rel(out["grad_f"][0], gf) < EVAL_RTOL


def _solve_inputs(N, B, seed_offset=0):
    P = nlp.Problem(N=N)
    Q, R, Rm = nlp.default_weights(4, 2)
    x0, up, traj = nlp.synthetic_batch(4, 2, N, P.h, B, offset=seed_offset)
    p = np.stack([P.pack_p(traj[b], Q, R, Rm, up[b]) for b in range(B)])
    return P, (Q, R, Rm), x0, up, traj, p


def test_solve_vs_golden(api, golden):
    gs = golden["solves"]
    for case in gs["cases"]:
        if case["u_bound"]:
            continue
        N = case["N"]
        P = nlp.Problem(N=N)
        pv = P.split_p(np.array(case["p"]))
        m = api.Model("double_pendulum", N=N)
        s = api.Solver(m, 1, tol=1e-10)
        s.set_weights(pv["Q"], pv["R"], pv["Rm"])
        V, lam, st = s.solve(np.array(case["x0"])[None], pv["u_init"][None].copy(), pv["r"][None].copy())
        assert st["status"][0] == 0
        Vg = np.array(case["V"])
        Ug = Vg[:N * 6].reshape(N, 6)[:, 4:]
        U = V[0, :N * 6].reshape(N, 6)[:, 4:]
        assert np.abs(U - Ug).max() < U_ATOL
        assert np.abs(V[0] - Vg).max() < U_ATOL
        assert abs(st["objective"][0] - case["f"]) <= OBJ_RTOL * abs(case["f"])
        assert rel(lam[0], case["lam_g"]) < 1e-6


# 3 + 256*2 forces T3's 32-instance tile configuration (chosen automatically only for batches beyond 1184) on ragged small batches;
# 4 + 256*k forces T4's tile width 8 / 28 / 32 / 4 (chosen from the batch size otherwise: 4 here)
SOLVE_VARIANTS = [1, 2, 3, 3 + 256 * 2, 4 + 256 * 1, 4 + 256 * 2, 4 + 256 * 3, 4 + 256 * 4]


@pytest.mark.parametrize("variant", SOLVE_VARIANTS)
@pytest.mark.parametrize("N,B", [(1, 3), (10, 7), (25, 33), (50, 64), (100, 8), (200, 4)])
def test_solve_vs_oracle(api, coracle, N, B, variant):
    P, (Q, R, Rm), x0, up, traj, p = _solve_inputs(N, B)
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B, tol=1e-10, kernel_variant=variant)
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    Vo, lamo, sto, bad = coracle.solve_batch("double_pendulum", N, P.h, False, x0, p, tol=1e-10)
    assert bad == 0 and (st["status"] == 0).all()
    s6 = 6
    U = V[:, :N * s6].reshape(B, N, s6)[:, :, 4:]
    Uo = Vo[:, :N * s6].reshape(B, N, s6)[:, :, 4:]
    assert np.abs(U - Uo).max() < U_ATOL
    assert np.abs(V - Vo).max() < U_ATOL
    assert (np.abs(st["objective"] - sto["f"]) <= OBJ_RTOL * np.abs(sto["f"])).all()
    # the returned point satisfies the KKT conditions as judged by the ORACLE's evaluator
    for b in range(min(B, 4)):
        o = coracle.nlp_eval("double_pendulum", N, P.h, False, V[b], p[b], lam[b], 1.0)
        assert np.abs(o["g"]).max() < 1e-9
        # stationarity: grad_f + J' lam = 0 on the free variables
        r = o["grad_f"].copy()
        for k in range(N):
            r[k * s6:k * s6 + 4] += o["Fx"][k].T @ lam[b, k * 4:(k + 1) * 4]
            r[k * s6 + 4:(k + 1) * s6] += o["Fu"][k].T @ lam[b, k * 4:(k + 1) * 4]
            r[(k + 1) * s6:(k + 1) * s6 + 4] -= lam[b, k * 4:(k + 1) * 4]
        assert np.abs(r[4:]).max() < 1e-7


def test_solve_reference_tolerance_and_warm_start(api, coracle):
    """tol = 1e-5 (ModelControl.cpp:54): still inside the north_star ball of the tight solution; the second
    call starts from the previous solution (ModelControl.cpp:161) and needs fewer iterations."""
    N, B = 50, 128
    P, (Q, R, Rm), x0, up, traj, p = _solve_inputs(N, B, seed_offset=1000)
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B)  # reference defaults: tol 1e-5, max_iter 200
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    Vo, _, sto, bad = coracle.solve_batch("double_pendulum", N, P.h, False, x0, p, tol=1e-11)
    assert bad == 0 and (st["status"] == 0).all()
    assert np.abs(V - Vo).max() < U_ATOL
    assert (np.abs(st["objective"] - sto["f"]) <= OBJ_RTOL * np.abs(sto["f"])).all()
    it_cold = st["iterations"].copy()
    V2, _, st2 = s.solve(x0, up, traj)
    assert (st2["iterations"] <= 1).all() and (it_cold >= 2).all()
    assert np.abs(V2 - V).max() < 1e-6
    s.reset_warm_start()
    V3, _, st3 = s.solve(x0, up, traj)
    assert np.array_equal(st3["iterations"], it_cold) and np.array_equal(V3, V)


def test_solve_device_pointers_match_host(api):
    import torch
    N, B = 25, 40
    P, (Q, R, Rm), x0, up, traj, p = _solve_inputs(N, B, seed_offset=50)
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B, tol=1e-9, warm_start=0)
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    dev = torch.device("cuda:0")
    Vt, lamt, stt = s.solve(torch.from_numpy(x0).to(dev), torch.from_numpy(up).to(dev), torch.from_numpy(traj).to(dev))
    torch.cuda.synchronize()
    assert np.array_equal(Vt.cpu().numpy(), V) and np.array_equal(lamt.cpu().numpy(), lam)
    st2 = api.stats_from_tensor(stt)
    assert np.array_equal(st2["iterations"], st["iterations"])


def test_solve_full_size_properties(api):
    """BASELINE config 2 size (B=4096, N=50): size-independent checks -- every instance converged, continuity
    holds, x_0 is pinned, shards reproduce the un-sharded result bit for bit, and re-solving from the solution is
    a fixed point."""
    N, B = 50, 4096
    P, (Q, R, Rm), x0, up, traj, _ = _solve_inputs(N, B)
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B, warm_start=0)
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    assert (st["status"] == 0).all() and (st["kkt_error"] <= 1e-5).all()
    assert np.array_equal(V[:, :4], x0)
    out = m.nlp_eval(V[:64], np.stack([P.pack_p(traj[b], Q, R, Rm, up[b]) for b in range(64)]), lam[:64], want=("f", "g"))
    assert np.abs(out["g"]).max() < 1e-5
    assert rel(out["f"], st["objective"][:64]) < 1e-12
    # shard [1024, 2048) solved alone == the same rows of the full batch
    s2 = api.Solver(m, 1024, warm_start=0)
    s2.set_weights(Q, R, Rm)
    Vs, _, _ = s2.solve(x0[1024:2048].copy(), up[1024:2048].copy(), traj[1024:2048].copy())
    assert np.array_equal(Vs, V[1024:2048])


def test_solve_headline_config_vs_oracle(api, coracle):
    """BASELINE config 2 exactly as bench.py times it (B = 4096, N = 50, default kernel variant = the wide-tile
    configuration, one CTA per tile on every SM): ALL 4096 instances against the oracle, at the parity tolerance and at
    the reference's tol = 1e-5 (ModelControl.cpp:54)."""
    N, B = 50, 4096
    P, (Q, R, Rm), x0, up, traj, p = _solve_inputs(N, B)
    Vt, _, stt, bad = coracle.solve_batch("double_pendulum", N, P.h, False, x0, p, tol=1e-10)
    assert bad == 0
    m = api.Model("double_pendulum", N=N)
    for tol in (1e-10, 1e-5):
        # north_star: "... at the same tolerance settings": the oracle stops by the same rule at the same tol
        Vo, lamo, sto, bad = coracle.solve_batch("double_pendulum", N, P.h, False, x0, p, tol=tol)
        assert bad == 0
        s = api.Solver(m, B, tol=tol, warm_start=0)
        s.set_weights(Q, R, Rm)
        V, lam, st = s.solve(x0, up, traj)
        assert (st["status"] == 0).all()
        assert np.abs(V - Vo).max() < U_ATOL and np.abs(V - Vt).max() < U_ATOL  # ... and inside the ball of the tight solution
        assert (np.abs(st["objective"] - sto["f"]) <= OBJ_RTOL * np.abs(sto["f"])).all()
        if tol == 1e-10:
            assert np.abs(V - Vo).max() < 1e-8 and rel(lam, lamo) < 1e-7
            # same Newton iterates as the oracle's band-LU solver: identical iteration counts on (nearly) every instance
            assert (st["iterations"] == sto["iters"]).mean() > 0.99, np.bincount(st["iterations"] - sto["iters"] + 5)
        s.close()


@pytest.mark.parametrize("variant", [1, 3, 3 + 256 * 2, 4 + 256 * 1, 4 + 256 * 3])
@pytest.mark.parametrize("N,B", [(25, 40), (50, 64)])
def test_linearised_solve_vs_oracle(api, coracle, N, B, variant):
    """is_linear models (ModelGenerator.cpp:38-58) -- the kernel linearises f at (state, control) itself
    (ModelControl.cpp:125-135) and solves the resulting QP; the oracle gets A, B, xdot_init, x_init through p."""
    P = nlp.Problem(N=N, is_linear=True)
    Q, R, Rm = nlp.default_weights(4, 2)
    x0, up, traj = nlp.synthetic_batch(4, 2, N, P.h, B, offset=400)
    up = np.random.default_rng(3).uniform(-1, 1, up.shape)  # a non-trivial linearisation point in u as well
    p = np.stack([P.pack_p(traj[b], Q, R, Rm, up[b], P.linearisation(x0[b], up[b])) for b in range(B)])
    Vo, lamo, sto, bad = coracle.solve_batch("double_pendulum", N, P.h, True, x0, p, tol=1e-10)
    assert bad == 0
    m = api.Model("double_pendulum", N=N, is_linear=True)
    s = api.Solver(m, B, tol=1e-10, kernel_variant=variant)
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    assert (st["status"] == 0).all() and (st["iterations"] <= 2).all()  # a QP: one Newton step (+ the test of the new point)
    assert np.abs(V - Vo).max() < U_ATOL and np.abs(V - Vo).max() < 1e-9
    assert (np.abs(st["objective"] - sto["f"]) <= OBJ_RTOL * np.abs(sto["f"])).all()
    assert rel(lam, lamo) < 1e-8
    r = nlp.solve(P, x0[0], p[0], tol=1e-11)  # numpy oracle (dense KKT) on instance 0
    assert np.abs(V[0] - r.V).max() < 1e-9 and abs(st["objective"][0] - r.f) <= OBJ_RTOL * abs(r.f)


def test_solve_per_instance_weights_and_errors(api):
    N, B = 10, 3
    P, (Q, R, Rm), x0, up, traj, p = _solve_inputs(N, B)
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B, tol=1e-10)
    Qb = np.stack([Q, 2 * Q, 0.5 * Q]); Rb = np.stack([R, R, 3 * R]); Rmb = np.stack([Rm, 2 * Rm, Rm])
    s.set_weights(Qb, Rb, Rmb, per_instance=True)
    V, _, st = s.solve(x0, up, traj)
    for b in range(B):
        pb = P.pack_p(traj[b], Qb[b], Rb[b], Rmb[b], up[b])
        r = nlp.solve(P, x0[b], pb, tol=1e-11)
        assert np.abs(V[b] - r.V).max() < U_ATOL and abs(st["objective"][b] - r.f) <= OBJ_RTOL * abs(r.f)
    with pytest.raises(api.MmpcError):
        api.Solver(m, 0)
    with pytest.raises(api.MmpcError):
        api.Model("no_such_model", N=5, nx=3, nu=1)


# interior point: T1 (thread per instance), the default (T3 for the pendulum) and T3's 32-instance tile configuration
BOUNDED_VARIANTS = [1, 0, 3 + 256 * 2]


@pytest.mark.parametrize("variant", BOUNDED_VARIANTS)
def test_bounded_solve_vs_golden(api, golden, variant):
    """Finite control limits (ModelControl::update_control_limits, ModelControl.cpp:205-209): interior-point path."""
    gs = golden["solves"]
    n = 0
    for case in gs["cases"]:
        if not case["u_bound"]:
            continue
        n += 1
        N, ub = case["N"], case["u_bound"]
        P = nlp.Problem(N=N)
        pv = P.split_p(np.array(case["p"]))
        m = api.Model("double_pendulum", N=N)
        s = api.Solver(m, 1, tol=1e-9, kernel_variant=variant)
        s.set_weights(pv["Q"], pv["R"], pv["Rm"])
        s.set_bounds(u_min=[-ub, -ub], u_max=[ub, ub])
        V, lam, st = s.solve(np.array(case["x0"])[None], pv["u_init"][None].copy(), pv["r"][None].copy())
        assert st["status"][0] == 0, st
        Vg = np.array(case["V"])
        U = V[0, :N * 6].reshape(N, 6)[:, 4:]
        assert np.abs(U).max() <= ub and np.abs(U).max() > ub - 1e-6  # the limit is active and respected
        assert np.abs(V[0] - Vg).max() < U_ATOL
        assert abs(st["objective"][0] - case["f"]) <= OBJ_RTOL * abs(case["f"])
    assert n == 2


@pytest.mark.parametrize("variant", BOUNDED_VARIANTS)
@pytest.mark.parametrize("N,B,ub,xb", [(25, 16, 2.5, None), (50, 40, 1.5, None), (25, 8, 4.0, 0.9)])
def test_bounded_solve_vs_oracle(api, coracle, N, B, ub, xb, variant):
    P, (Q, R, Rm), x0, up, traj, p = _solve_inputs(N, B, seed_offset=300)
    x_min = x_max = None
    if xb is not None:  # also a state bound on the first joint velocity-like state
        x_min, x_max = [-NB, -NB, -NB, -NB], [xb, NB, NB, NB]
    m = api.Model("double_pendulum", N=N, u_min=[-ub, -ub], u_max=[ub, ub], x_min=x_min, x_max=x_max)
    s = api.Solver(m, B, tol=1e-10, kernel_variant=variant)  # same barrier end point (mu = tol/10) on both sides
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    lb, ubv = nlp.bounds_vectors(P, x0[0], x_min=x_min, x_max=x_max, u_min=[-ub] * 2, u_max=[ub] * 2)
    Vo, _, sto, bad = coracle.solve_batch("double_pendulum", N, P.h, False, x0, p, lb=lb, ub=ubv, tol=1e-10)
    good = (sto["status"] == 0)
    assert bad == 0 and good.all() and (st["status"] == 0).all(), (st["status"], sto["status"])  # no allowance: every instance
    assert np.abs(V[good] - Vo[good]).max() < U_ATOL
    assert (np.abs(st["objective"] - sto["f"])[good] <= OBJ_RTOL * np.abs(sto["f"][good])).all()
    U = V[:, :N * 6].reshape(B, N, 6)[:, :, 4:]
    assert np.abs(U).max() <= ub
    if xb is not None:
        assert V[:, 6::6][:, :N].max() <= xb + 1e-12
    # unbounded handle switched to bounds later (update_control_limits) gives the same answer
    m2 = api.Model("double_pendulum", N=N)
    s2 = api.Solver(m2, B, tol=1e-10, kernel_variant=variant)
    s2.set_weights(Q, R, Rm)
    s2.solve(x0, up, traj)
    s2.set_bounds(x_min=x_min, x_max=x_max, u_min=[-ub, -ub], u_max=[ub, ub])
    V2, _, st2 = s2.solve(x0, up, traj)
    assert np.abs(V2[good] - V[good]).max() < 1e-7


# ---------------------------------------------------------------------------------------------------------------
# BASELINE config 3: 4-DOF arm (builder-defined).  Oracle = oracle/arm4_jets.c (forward-mode AD over the
# closed-form mass matrix), product = symbolic engine + generated CUDA: two independent derivations.
# ---------------------------------------------------------------------------------------------------------------
def test_arm4_eval_nodes_vs_oracle(api, coracle):
    rng = np.random.default_rng(44)
    n = 3000
    x = np.c_[rng.uniform(-1.5, 1.5, (n, 4)), rng.uniform(-2, 2, (n, 4))]
    u = rng.uniform(-3, 3, (n, 4)); xn = rng.uniform(-1, 1, (n, 8)); w = rng.uniform(-1, 1, (n, 8))
    m = api.Model("arm4", N=10)
    assert (m.nx, m.nu, m.NV, m.np) == (8, 4, 128, 10 * 8 + 8 + 12)
    g, Fx, Fu, H = m.eval_nodes(x.T.copy(), u.T.copy(), xn.T.copy(), w.T.copy())
    go, Fxo, Fuo, Ho = coracle.node_eval("arm4", m.h, x, u, xn, w)
    iu = np.triu_indices(12)
    assert rel(g.T, go) < EVAL_RTOL and rel(Fx.T.reshape(n, 8, 8), Fxo) < EVAL_RTOL and rel(Fu.T.reshape(n, 8, 4), Fuo) < EVAL_RTOL
    assert rel(H.T, Ho[:, iu[0], iu[1]]) < 1e-11  # sums of ~1e4 terms: one more digit of slack than the 1e-12 of the pendulum


def _arm_inputs(N, B):
    from mahi_mpc_b200 import workload
    P = nlp.Problem(model="arm4", N=N)
    Q, R, Rm = workload.default_weights(8, 4)
    x0, up, traj = workload.synthetic_batch(8, 4, N, P.h, B)
    traj *= 0.5
    p = np.stack([P.pack_p(traj[b], Q, R, Rm, up[b]) for b in range(B)])
    return P, (Q, R, Rm), x0, up, traj, p


def test_arm4_nlp_eval_vs_oracle(api, coracle):
    N, B = 12, 3
    P, (Q, R, Rm), x0, up, traj, p = _arm_inputs(N, B)
    rng = np.random.default_rng(45)
    V = rng.uniform(-0.5, 0.5, (B, P.NV)); lam = rng.uniform(-1, 1, (B, P.M))
    m = api.Model("arm4", N=N)
    out = m.nlp_eval(V, p, lam, lam_f=1.0)
    for b in range(B):
        o = coracle.nlp_eval("arm4", N, P.h, False, V[b], p[b], lam[b], 1.0)
        assert abs(out["f"][b] - o["f"]) <= EVAL_RTOL * abs(o["f"])
        for k in ("g", "grad_f", "Fx", "Fu"):
            assert rel(out[k][b], o[k]) < EVAL_RTOL, k
        assert rel(out["Hzz"][b], o["Hzz"]) < 1e-11


@pytest.mark.parametrize("variant", [1, 2, 3])
def test_arm4_solve_vs_oracle(api, coracle, variant):
    N, B = 20, 9
    P, (Q, R, Rm), x0, up, traj, p = _arm_inputs(N, B)
    m = api.Model("arm4", N=N)
    s = api.Solver(m, B, tol=1e-10, kernel_variant=variant)
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    Vo, lamo, sto, bad = coracle.solve_batch("arm4", N, P.h, False, x0, p, tol=1e-10)
    assert bad == 0 and (st["status"] == 0).all(), (st, sto)
    assert np.abs(V - Vo).max() < U_ATOL
    assert (np.abs(st["objective"] - sto["f"]) <= OBJ_RTOL * np.abs(sto["f"])).all()


def test_arm4_config3_shape_vs_oracle(api, coracle):
    """BASELINE config 3's horizon (N = 50) on the variant the ABI picks for the arm by default, 256 instances."""
    N, B = 50, 256
    P, (Q, R, Rm), x0, up, traj, p = _arm_inputs(N, B)
    m = api.Model("arm4", N=N)
    s = api.Solver(m, B, tol=1e-10)
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    Vo, lamo, sto, bad = coracle.solve_batch("arm4", N, P.h, False, x0, p, tol=1e-10)
    assert bad == 0 and (st["status"] == 0).all(), (st["status"], sto["status"])
    assert np.abs(V - Vo).max() < U_ATOL
    assert (np.abs(st["objective"] - sto["f"]) <= OBJ_RTOL * np.abs(sto["f"])).all()
    assert rel(lam, lamo) < 1e-6


# ---------------------------------------------------------------------------------------------------------------
# RK4 x M shooting integrator (workingMPC_8_30_21.m:46-59): rollout + exact forward sensitivities;
# the solve uses the Gauss-Newton part of the Hessian and must reach the same KKT point as the oracle.
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("substeps", [1, 4])
def test_rk4_rollout_and_jacobian_vs_oracle(api, substeps):
    n = 513
    x, u, xn, _ = random_nodes(n, 7 + substeps)
    m = api.Model("double_pendulum", N=10, step_size_us=100000)  # dt = 0.1 s as in the MATLAB prototypes
    P = nlp.Problem(N=10, h=0.1, integrator="rk4", substeps=substeps)
    o = api.default_options(integrator=api.RK4, substeps=substeps)
    g, Fx, Fu, H = m.eval_nodes(x.T.copy(), u.T.copy(), xn.T.copy(), opts=o)
    F, Fxo, Fuo = P.step_jac(x, u, {})
    assert H is None
    assert rel(g.T, F - xn) < EVAL_RTOL and rel(Fx.T.reshape(n, 4, 4), Fxo) < EVAL_RTOL and rel(Fu.T.reshape(n, 4, 2), Fuo) < EVAL_RTOL
    # exact second derivatives through the integrator (forward-over-adjoint on the GPU, complex step in the oracle)
    w = np.random.default_rng(3).uniform(-1, 1, (n, 4))
    g2, Fx2, Fu2, H = m.eval_nodes(x.T.copy(), u.T.copy(), xn.T.copy(), w.T.copy(), opts=o)
    Ho = P.step_hess(x, u, {}, w)
    iu = np.triu_indices(6)
    assert np.array_equal(g2, g) or rel(g2, g) < 1e-14
    assert (np.abs(H.T - Ho[:, iu[0], iu[1]]).max(axis=1) <= 1e-11 * np.abs(Ho).max(axis=(1, 2))).all()
    with pytest.raises(api.MmpcError):  # more sub-steps than the stage store holds
        m.eval_nodes(x.T.copy(), u.T.copy(), xn.T.copy(), w.T.copy(), opts=api.default_options(integrator=api.RK4, substeps=5))


def test_rk4_hessian_vs_mpmath_golden(api):
    """F, dF/dz and sum_i w_i d2F_i/dz2 of the RK4 x M interval map against 60-digit mpmath values
    (tests/golden/rk4_hessian.json, tests/golden/make_rk4_hessian.py; integrator of workingMPC_8_30_21.m:46-59)."""
    import json
    with open(os.path.join(os.path.dirname(__file__), "golden", "rk4_hessian.json")) as fh:
        cases = json.load(fh)["cases"]
    iu = np.triu_indices(6)
    for c in cases:
        m = api.Model("double_pendulum", N=4, step_size_us=int(round(c["h"] * 1e6)))
        o = api.default_options(integrator=api.RK4, substeps=c["substeps"])
        x = np.array([c["x"]]).T.copy(); u = np.array([c["u"]]).T.copy(); w = np.array([c["w"]]).T.copy()
        g, Fx, Fu, H = m.eval_nodes(x, u, np.zeros((4, 1)), w, opts=o)
        Fz = np.concatenate([Fx[:, 0].reshape(4, 4), Fu[:, 0].reshape(4, 2)], axis=1)
        Hg = np.array(c["H"])
        assert np.abs(g[:, 0] - c["F"]).max() < 1e-14
        assert np.abs(Fz - np.array(c["Fz"])).max() <= 1e-13 * np.abs(c["Fz"]).max()
        assert np.abs(H[:, 0] - Hg[iu]).max() <= 1e-12 * np.abs(Hg).max()


def test_rk4_nlp_and_solve_vs_oracle(api):
    N, B, h = 10, 3, 0.05
    P = nlp.Problem(N=N, h=h, integrator="rk4", substeps=2)
    Q, R, Rm = nlp.default_weights(4, 2)
    x0, up, traj = nlp.synthetic_batch(4, 2, N, h, B)
    p = np.stack([P.pack_p(traj[b], Q, R, Rm, up[b]) for b in range(B)])
    m = api.Model("double_pendulum", N=N, step_size_us=int(h * 1e6))
    rng = np.random.default_rng(11)
    V = rng.uniform(-0.5, 0.5, (B, P.NV))
    out = m.nlp_eval(V, p, None, opts=api.default_options(integrator=api.RK4, substeps=2), want=("f", "g", "grad_f", "Fx", "Fu"))
    for b in range(B):
        f, gf = P.nlp_grad_f(V[b], p[b])
        g, J = P.nlp_jac_g(V[b], p[b])
        assert abs(out["f"][b] - f) <= EVAL_RTOL * abs(f) and rel(out["g"][b], g) < EVAL_RTOL and rel(out["grad_f"][b], gf) < EVAL_RTOL
        for k in range(N):
            assert rel(out["Fx"][b, k], J[k * 4:(k + 1) * 4, k * 6:k * 6 + 4]) < EVAL_RTOL
    # exact Lagrangian Hessian through the integrator (nlp_hess_l of the prototypes' NLP)
    lam_g = rng.uniform(-1, 1, (B, P.M))
    outh = m.nlp_eval(V, p, lam_g, opts=api.default_options(integrator=api.RK4, substeps=2), want=("Hzz",))
    for b in range(B):
        Hd = P.nlp_hess_l(V[b], p[b], 1.0, lam_g[b])
        for k in range(N):  # diagonal node blocks, as test_nlp_eval_vs_oracle maps them for Euler
            blk = Hd[k * 6:(k + 1) * 6, k * 6:(k + 1) * 6]
            assert np.abs(outh["Hzz"][b, k] - blk).max() <= 1e-10 * max(1.0, np.abs(blk).max()), (b, k)
    ref = [nlp.solve(P, x0[b], p[b], tol=1e-10, max_iter=400) for b in range(B)]
    assert all(r.status == 0 for r in ref)
    # default (T4), T1 (thread per instance), T3 / T4 with 32-instance tiles; exact Newton: the oracle's iteration counts
    for variant in (0, 1, 3 + 256 * 2, 4 + 256 * 3):
        s = api.Solver(m, B, tol=1e-9, integrator=api.RK4, substeps=2, kernel_variant=variant)
        s.set_weights(Q, R, Rm)
        Vs, lam, st = s.solve(x0, up, traj)
        assert (st["status"] == 0).all(), (variant, st)
        for b, r in enumerate(ref):
            assert np.abs(Vs[b] - r.V).max() < U_ATOL and abs(st["objective"][b] - r.f) <= OBJ_RTOL * abs(r.f), variant
            assert abs(int(st["iterations"][b]) - r.iters) <= 1, (variant, st["iterations"], [q.iters for q in ref])
    # Gauss-Newton on request: the same KKT point, more iterations
    s = api.Solver(m, B, tol=1e-9, integrator=api.RK4, substeps=2, hessian=api.HESSIAN_GAUSS_NEWTON)
    s.set_weights(Q, R, Rm)
    Vg, _, stg = s.solve(x0, up, traj)
    assert (stg["status"] == 0).all() and stg["iterations"].mean() > st["iterations"].mean()
    for b, r in enumerate(ref):
        assert np.abs(Vg[b] - r.V).max() < U_ATOL
    with pytest.raises(api.MmpcError):
        api.Solver(m, B, integrator=api.RK4, substeps=8)  # exact Hessian: up to 4 sub-steps
    api.Solver(m, B, integrator=api.RK4, substeps=8, hessian=api.HESSIAN_GAUSS_NEWTON).close()


@pytest.mark.parametrize("plant", ["rk4", "euler"])
def test_closed_loop_batch_vs_oracle(api, coracle, plant):
    """device-resident closed loop (solve -> apply u_0 -> plant step) for a batch, against the same loop
    driven on the host with the ORACLE's solver and plant."""
    from oracle import models
    N, B, cycles, h = 25, 6, 15, 0.002
    P = nlp.Problem(N=N)
    Q, R, Rm = nlp.default_weights(4, 2)
    x0, up, traj_long = nlp.synthetic_batch(4, 2, cycles + N - 1, h, B, offset=77)
    mdl = models.numeric("double_pendulum")
    f = lambda x, u: mdl["f"](x.T, u.T).T
    # oracle loop (host): warm start carried like ModelControl.cpp:161
    x, u, V0 = x0.copy(), up.copy(), None
    Xo, Uo = np.empty((B, cycles, 4)), np.empty((B, cycles, 2))
    for c in range(cycles):
        p = np.stack([P.pack_p(traj_long[b, c:c + N], Q, R, Rm, u[b]) for b in range(B)])
        V0, _, st, bad = coracle.solve_batch("double_pendulum", N, h, False, x, p, V0=V0, tol=1e-11)
        assert bad == 0
        u = V0[:, 4:6].copy()
        Xo[:, c], Uo[:, c] = x, u
        if plant == "rk4":
            k1 = f(x, u); k2 = f(x + h / 2 * k1, u); k3 = f(x + h / 2 * k2, u); k4 = f(x + h * k3, u)
            x = x + h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        else:
            x = x + h * f(x, u)
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B, tol=1e-10)
    s.set_weights(Q, R, Rm)
    xs, us = x0.copy(), up.copy()
    X, U, it = s.closed_loop(xs, us, traj_long, cycles, plant=api.RK4 if plant == "rk4" else api.EULER)
    assert np.abs(U - Uo).max() < U_ATOL and np.abs(X - Xo).max() < 1e-8
    assert np.abs(xs - x).max() < 1e-8 and np.abs(us - u).max() < U_ATOL  # final state / control written back
    assert (it[:, 1:] <= it[:, :1]).all() and (it >= 0).all()             # warm starts never need more than the cold solve


@pytest.mark.parametrize("variant", SOLVE_VARIANTS)
def test_failed_instance_does_not_poison_the_batch(api, coracle, variant):
    """SURVEY section 5 (failure detection): a non-finite input makes THAT instance report MMPC_SOLVE_NUMERICAL; its
    neighbours (same warp / same CTA tile) still converge to the oracle's solution."""
    N, B = 25, 40
    P, (Q, R, Rm), x0, up, traj, p = _solve_inputs(N, B, seed_offset=9)
    bad = [3, 17, 39]
    x0b, trajb = x0.copy(), traj.copy()
    x0b[3, 1] = np.nan
    trajb[17, 5, 2] = np.inf
    x0b[39, 0] = 1e308
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B, tol=1e-10, max_iter=50, kernel_variant=variant)
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0b, up, trajb)
    good = np.setdiff1d(np.arange(B), bad)
    assert (st["status"][bad] != 0).all() and (st["status"][good] == 0).all(), st["status"]
    Vo, _, sto, nbad = coracle.solve_batch("double_pendulum", N, P.h, False, x0[good], p[good], tol=1e-10)
    assert nbad == 0 and np.abs(V[good] - Vo).max() < U_ATOL


def test_max_iter_is_honoured_and_reported(api):
    N, B = 50, 33
    P, (Q, R, Rm), x0, up, traj, p = _solve_inputs(N, B)
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B, tol=1e-12, max_iter=1)
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    assert (st["status"] == 1).all() and (st["iterations"] == 1).all() and (st["kkt_error"] > 1e-12).all()
    s.set_options(max_iter=200, tol=1e-5)
    V, lam, st = s.solve(x0, up, traj)
    assert (st["status"] == 0).all()


# ---------------------------------------------------------------------------------------------------------------
# several devices in one process (mmpc_group_*), the packed small-batch staging path, observability
# ---------------------------------------------------------------------------------------------------------------
def test_group_solve_shards_and_gathers(api):
    """mmpc_group_solve: the batch cut into contiguous shards (here three handles on device 0 -- the driver box of the
    parity tests has one GPU; tools/config3_multi.py and bench.py --gpus N exercise one process per GPU), every shard's plans
    land in the caller's arrays, and mmpc_group_gather collects them on one device.  Same results as ONE handle."""
    import torch
    N, B = 25, 70
    P, (Q, R, Rm), x0, up, traj, p = _solve_inputs(N, B)
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B, tol=1e-10, warm_start=0); s.set_weights(Q, R, Rm)
    V1, lam1, st1 = s.solve(x0, up, traj)
    g = api.Group(m, B, [0, 0, 0], tol=1e-10, warm_start=0); g.set_weights(Q, R, Rm)
    assert g.shards() == [(0, 0, 24), (0, 24, 23), (0, 47, 23)]  # mahi_mpc_b200/shard.py's partition
    Vg, lamg, stg = g.solve(x0, up, traj)
    assert (stg["status"] == 0).all() and np.array_equal(stg["iterations"], st1["iterations"])
    assert np.abs(Vg - V1).max() < 1e-12 and np.abs(lamg - lam1).max() < 1e-9
    Vd = torch.zeros(B, m.NV, dtype=torch.float64, device="cuda:0")
    std = torch.zeros(B, 5, dtype=torch.float64, device="cuda:0")
    ms = g.gather(0, Vd.data_ptr(), std.data_ptr())
    assert ms >= 0.0 and np.array_equal(Vd.cpu().numpy(), Vg)
    assert np.array_equal(api.stats_from_tensor(std)["iterations"], stg["iterations"])
    with pytest.raises(api.MmpcError):
        api.Group(m, 2, [0, 0, 0])  # fewer instances than shards


def test_small_batch_packed_staging_and_stats_jsonl(api, coracle, tmp_path):
    """B = 1 (one ModelControl::calc_u): the zero-copy path (the kernel reads / writes the page-locked staging buffers in place) gives
    the oracle's plan, and so does the packed-copy path (MMPC_ZERO_COPY=0) -- bit for bit the same one; mmpc_last_timing reports
    the phases; MMPC_STATS_JSONL appends one line per solve."""
    import json, subprocess, sys
    N = 25
    P, (Q, R, Rm), x0, up, traj, p = _solve_inputs(N, 1)
    Vo, _, sto, bad = coracle.solve_batch("double_pendulum", N, P.h, False, x0, p, tol=1e-10)
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, 1, tol=1e-10, warm_start=0); s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    assert st["status"][0] == 0 and np.abs(V - Vo).max() < U_ATOL
    assert abs(st["objective"][0] - sto["f"][0]) <= 1e-8 * abs(sto["f"][0]) and abs(int(st["iterations"][0]) - int(sto["iters"][0])) <= 1
    t = s.last_timing()
    assert t["h2d_ms"] == 0 and t["kernel_ms"] > 0 and t["d2h_ms"] == 0  # zero-copy: the transfers are the kernel's own loads / stores
    # ragged small batches through the zero-copy path (tile width 4: full tiles, a partial tile, several CTAs), multipliers included
    for B in (3, 4, 9, 33):
        P, (Q, R, Rm), x0b, upb, trajb, pb = _solve_inputs(N, B)
        Vob, lamo, stob, bad = coracle.solve_batch("double_pendulum", N, P.h, False, x0b, pb, tol=1e-10)
        sb = api.Solver(m, B, tol=1e-10, warm_start=0); sb.set_weights(Q, R, Rm)
        Vb, lamb, stb = sb.solve(x0b, upb, trajb)
        assert bad == 0 and (stb["status"] == 0).all() and np.abs(stb["iterations"] - stob["iters"]).max() <= 1
        assert np.abs(Vb - Vob).max() < U_ATOL and np.abs(stb["objective"] - stob["f"]).max() <= 1e-8 * np.abs(stob["f"]).max()
        V2, lam2, st2 = sb.solve(x0b, upb, trajb)  # the staging buffers are reused: the same bits again
        assert np.array_equal(V2, Vb) and np.array_equal(lam2, lamb)
    # the packed-copy path in a child process (the switch is read once): the same bits, three phases
    np.save(tmp_path / "V_zc.npy", V)
    code = ("import numpy as np, mahi_mpc_b200.api as api; from oracle import nlp;"
            "P = nlp.Problem(N=25); Q, R, Rm = nlp.default_weights(4, 2); x0, up, traj = nlp.synthetic_batch(4, 2, 25, P.h, 1);"
            "m = api.Model('double_pendulum', N=25); s = api.Solver(m, 1, tol=1e-10, warm_start=0); s.set_weights(Q, R, Rm);"
            "V, lam, st = s.solve(x0, up, traj); t = s.last_timing();"
            "assert t['h2d_ms'] > 0 and t['kernel_ms'] > 0 and t['d2h_ms'] > 0, t;"
            f"assert np.array_equal(V, np.load(r'{tmp_path / 'V_zc.npy'}'))")
    subprocess.run([sys.executable, "-c", code], check=True, env=dict(os.environ, MMPC_ZERO_COPY="0"),
                   cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    # the environment variable is read once per process: a child process writes the file
    out = tmp_path / "stats.jsonl"
    code = ("import numpy as np, mahi_mpc_b200.api as api; from mahi_mpc_b200 import workload as w;"
            "x0, up, traj = w.synthetic_batch(4, 2, 25, 0.002, 5); Q, R, Rm = w.default_weights(4, 2);"
            "m = api.Model('double_pendulum', N=25); s = api.Solver(m, 5, warm_start=0); s.set_weights(Q, R, Rm);"
            "s.solve(x0, up, traj); s.solve(x0, up, traj)")
    env = dict(os.environ, MMPC_STATS_JSONL=str(out))
    subprocess.run([sys.executable, "-c", code], check=True, env=env, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    lines = [json.loads(l) for l in out.read_text().splitlines()]
    assert len(lines) == 2 and lines[0]["batch"] == 5 and lines[0]["converged"] == 5 and lines[0]["kernel_ms"] > 0 and lines[0]["N"] == 25


def test_pinned_host_buffers_zero_copy_matches_device_resident(api):
    """Page-locked caller buffers (torch pin_memory) are read and written by the tile kernels in place (zero-copy inputs staged per
    tile, transposing epilogue for the outputs, sub-batches on their own streams): the plans, multipliers and statistics must be
    the bits of the device-resident solve -- full batches, a ragged one (partial last tile and sub-batch) and a bounded one (T3)."""
    import torch
    N = 50
    m = api.Model("double_pendulum", N=N)
    for B, bounded in ((4096, False), (4096 + 37, False), (1500, False), (4096, True), (8192 + 5, False)):  # the last: two waves, copy engines, 8 pieces
        P, (Q, R, Rm), x0, up, traj, p = _solve_inputs(N, B)
        s = api.Solver(m, B, tol=1e-8, warm_start=0); s.set_weights(Q, R, Rm)
        if bounded:
            s.set_bounds(u_min=[-2.0, -2.0], u_max=[2.0, 2.0])
        d = [torch.from_numpy(a).cuda() for a in (x0, up, traj)]
        Vd, lamd, std = s.solve(*d)  # stream-ordered on torch's current stream; the host-buffer call below waits for it (ev[3])
        h = [torch.from_numpy(a).pin_memory() for a in (x0, up, traj)]
        Vh = torch.empty(B, m.NV, dtype=torch.float64).pin_memory(); lamh = torch.empty(B, m.M, dtype=torch.float64).pin_memory()
        sth = torch.empty(B, 5, dtype=torch.float64).pin_memory()
        Vh.fill_(float("nan")); lamh.fill_(float("nan"))
        s.solve(*h, V_out=Vh, lam_out=lamh, stats_out=sth)
        assert torch.equal(Vh, Vd.cpu()) and torch.equal(lamh, lamd.cpu()), (B, bounded)
        a, b = api.stats_from_tensor(sth), api.stats_from_tensor(std)
        assert np.array_equal(a["iterations"], b["iterations"]) and np.array_equal(a["status"], b["status"]) and np.array_equal(a["objective"], b["objective"])
        assert (a["status"] == 0).mean() > (0.95 if bounded else 0.999)
        # pageable buffers (copy engines) give the same bits too
        Vp, lamp, stp = s.solve(x0, up, traj)
        assert np.array_equal(Vp, Vd.cpu().numpy()) and np.array_equal(lamp, lamd.cpu().numpy())
        s.close()


# ---------------------------------------------------------------------------------------------------------------
# MMPC_COST_NODE: the MATLAB prototypes' objective (workingMPC_8_30_21.m:90-106) -- tracking error on the decision
# variable x_{k+1} instead of F(x_k, u_k); with RK4 x 4, Q <- 1000 Q, R <- 0, Rm <- R this is the prototypes' NLP.
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("proto", [False, True])
def test_node_cost_nlp_eval_vs_oracle(api, proto):
    N, B = (10, 3) if proto else (25, 2)
    h = 0.1 if proto else 0.002
    kw = dict(integrator="rk4", substeps=4) if proto else {}
    P = nlp.Problem(N=N, h=h, cost="node", **kw)
    Q, R, Rm = (np.full(4, 1000.0), np.zeros(2), np.ones(2)) if proto else nlp.default_weights(4, 2)
    x0, up, traj = nlp.synthetic_batch(4, 2, N, h, B)
    p = np.stack([P.pack_p(traj[b], Q, R, Rm, up[b]) for b in range(B)])
    m = api.Model("double_pendulum", N=N, step_size_us=int(round(h * 1e6)))
    rng = np.random.default_rng(21)
    V = rng.uniform(-0.5, 0.5, (B, P.NV)); lam = rng.uniform(-1, 1, (B, P.M))
    o = api.default_options(cost=api.COST_NODE, **(dict(integrator=api.RK4, substeps=4) if proto else {}))
    out = m.nlp_eval(V, p, lam, opts=o)
    for b in range(B):
        f, gf = P.nlp_grad_f(V[b], p[b])
        H = P.nlp_hess_l(V[b], p[b], 1.0, lam[b])
        assert abs(out["f"][b] - f) <= EVAL_RTOL * abs(f) and rel(out["g"][b], P.nlp_g(V[b], p[b])) < EVAL_RTOL
        assert np.abs(out["grad_f"][b] - gf).max() <= 1e-11 * np.abs(gf).max()
        for k in range(N):
            blk = H[k * 6:(k + 1) * 6, k * 6:(k + 1) * 6]
            assert np.abs(out["Hzz"][b, k] - blk).max() <= 1e-10 * max(1.0, np.abs(blk).max()), (b, k)
        assert np.allclose(np.diag(H[N * 6:, N * 6:]), 2 * Q)  # the implicit constant block of x_N


@pytest.mark.parametrize("variant", [0, 1, 4 + 256 * 1, 4 + 256 * 2, 4 + 256 * 3])
@pytest.mark.parametrize("proto", [False, True])
def test_node_cost_solve_vs_oracle(api, variant, proto):
    N, B = (10, 5) if proto else (25, 9)
    h = 0.1 if proto else 0.002
    kw = dict(integrator="rk4", substeps=4) if proto else {}
    P = nlp.Problem(N=N, h=h, cost="node", **kw)
    Q, R, Rm = (np.full(4, 1000.0), np.zeros(2), np.ones(2)) if proto else nlp.default_weights(4, 2)
    x0, up, traj = nlp.synthetic_batch(4, 2, N, h, B)
    if proto:
        x0 = x0 * 0.0  # the prototypes start at rest
    p = np.stack([P.pack_p(traj[b], Q, R, Rm, up[b]) for b in range(B)])
    ref = [nlp.solve(P, x0[b], p[b], tol=1e-10, max_iter=200) for b in range(B)]
    assert all(r.status == 0 for r in ref)
    m = api.Model("double_pendulum", N=N, step_size_us=int(round(h * 1e6)))
    s = api.Solver(m, B, tol=1e-9, warm_start=0, kernel_variant=variant, cost=api.COST_NODE,
                   **(dict(integrator=api.RK4, substeps=4) if proto else {}))
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    assert (st["status"] == 0).all(), st
    for b, r in enumerate(ref):
        assert np.abs(V[b] - r.V).max() < U_ATOL and abs(st["objective"][b] - r.f) <= OBJ_RTOL * max(1.0, abs(r.f)), (variant, b)
        assert abs(int(st["iterations"][b]) - r.iters) <= 1
        assert np.abs(lam[b] - r.lam_g).max() <= 1e-6 * max(1.0, np.abs(r.lam_g).max())


def test_node_cost_with_control_limits_uses_t1(api):
    N, B = 25, 6
    P = nlp.Problem(N=N, cost="node")
    Q, R, Rm = nlp.default_weights(4, 2)
    x0, up, traj = nlp.synthetic_batch(4, 2, N, P.h, B)
    p = np.stack([P.pack_p(traj[b], Q, R, Rm, up[b]) for b in range(B)])
    ub = 2.0
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B, tol=1e-8, warm_start=0, cost=api.COST_NODE)
    s.set_weights(Q, R, Rm)
    s.set_bounds(u_min=[-ub, -ub], u_max=[ub, ub])
    V, lam, st = s.solve(x0, up, traj)
    assert (st["status"] == 0).all()
    for b in range(B):
        lb_, ub_ = nlp.bounds_vectors(P, x0[b], u_min=[-ub] * 2, u_max=[ub] * 2)
        r = nlp.solve(P, x0[b], p[b], lb=lb_, ub=ub_, tol=1e-9, max_iter=300)
        assert r.status == 0 and np.abs(V[b] - r.V).max() < U_ATOL


@pytest.mark.parametrize("N", [10, 100, 200])
def test_horizon_sweep_full_batch_vs_oracle(api, coracle, N):
    """BASELINE config 4 (horizon sweep at batch 4096) as bench.py's `other_configs.horizon_sweep` runs it: ALL 4096 instances of
    every horizon against the oracle at the reference's tol = 1e-5 (N = 50 is test_solve_headline_config_vs_oracle)."""
    B = 4096
    P, (Q, R, Rm), x0, up, traj, p = _solve_inputs(N, B)
    Vo, lamo, sto, bad = coracle.solve_batch("double_pendulum", N, P.h, False, x0, p, tol=1e-5)
    assert bad == 0
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B, warm_start=0)
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    assert (st["status"] == 0).all() and (st["kkt_error"] <= 1e-5).all()
    assert np.abs(V - Vo).max() < U_ATOL
    assert (np.abs(st["objective"] - sto["f"]) <= OBJ_RTOL * np.abs(sto["f"])).all()
    assert (st["iterations"] == sto["iters"]).mean() > 0.99, np.bincount(np.abs(st["iterations"] - sto["iters"]))
