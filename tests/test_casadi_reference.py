"""Parity against the TRUE reference path -- CasADi + IPOPT + MUMPS -- rebuilt with the CasADi Python API
(oracle/casadi_reference.py restates ModelGenerator.cpp:23-233 and ModelControl.cpp:54-59,116-172 call for call).
Skipped where `import casadi` fails (this image: no network, no wheel); on a machine with `pip install casadi` these
tests turn SURVEY 8c's "parity unpinned" into a pinned oracle and a pinned CUDA path."""
import numpy as np
import pytest

from oracle import casadi_reference as cref
from oracle import nlp

needs_casadi = pytest.mark.skipif(not cref.HAVE_CASADI, reason="casadi (with its bundled IPOPT/MUMPS) is not installed")


def test_hook_is_importable_and_reports_availability():
    assert isinstance(cref.HAVE_CASADI, bool)
    if not cref.HAVE_CASADI:
        with pytest.raises(ImportError):
            cref.ReferenceNLP()


@needs_casadi
def test_oracle_functions_and_solutions_vs_casadi_ipopt():
    dev = cref.self_check(N=25, B=4, tol=1e-9)
    assert dev["f"] < 1e-12 and dev["g"] < 1e-12 and dev["grad_f"] < 1e-12 and dev["jac_g"] < 1e-12 and dev["hess_l"] < 1e-11, dev
    assert dev["V"] < 1e-6 and dev["obj"] < 1e-8, dev  # north_star's solve tolerances


@needs_casadi
@pytest.mark.gpu
def test_cuda_solve_vs_casadi_ipopt(native_lib):
    import mahi_mpc_b200.api as api
    N, B = 25, 8
    P = nlp.Problem(N=N)
    ref = cref.ReferenceNLP(N=N, h=P.h)
    Q, R, Rm = nlp.default_weights(4, 2)
    x0, up, traj = nlp.synthetic_batch(4, 2, N, P.h, B)
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B, warm_start=0)  # the reference's tol = 1e-5, max_iter = 200
    s.set_weights(Q, R, Rm)
    V, lam, st = s.solve(x0, up, traj)
    assert (st["status"] == 0).all()
    for b in range(B):
        r = ref.solve(x0[b], ref.pack_p(traj[b], Q, R, Rm, up[b]))
        assert r["success"], r
        U = V[b, :N * 6].reshape(N, 6)[:, 4:]
        Ur = r["V"][:N * 6].reshape(N, 6)[:, 4:]
        assert np.abs(U - Ur).max() < 1e-6 and abs(st["objective"][b] - r["f"]) <= 1e-8 * abs(r["f"])
