"""ORACLE (test infrastructure): the TRUE reference path, rebuilt with the CasADi Python API.

mahi-mpc defines its NLP in C++ against CasADi (SX/MX) and solves it with IPOPT + MUMPS; neither library is in this image
or installable offline (CMakeLists.txt:34-45: CasADi @ fadc8644, IPOPT 3.14.3), so every parity claim of this repository is
pinned against the restatements in oracle/nlp.py / oracle/mpc_oracle.c ("parity unpinned" for the converged solutions,
DESIGN.md 5).  This module closes that gap on any machine that has `pip install casadi` (its wheels bundle IPOPT + MUMPS):
it restates src/Mahi/Mpc/ModelGenerator.cpp:23-233 and ModelControl.cpp:28-68,116-172 line for line with `casadi` calls
-- the same symbols, the same V / p layouts, the same objective and constraints, the same solver options -- so that

  * nlp_functions()   gives CasADi's own nlp_f / nlp_g / nlp_grad_f / nlp_jac_g / nlp_hess_l (what generate_dependencies
                      emits as C, ModelGenerator.cpp:235-239) to compare oracle/nlp.py and the CUDA mmpc_nlp_eval against,
  * solve()           runs IPOPT (tol 1e-5, max_iter 200, mumps: ModelControl.cpp:54-59) from the reference's cold start
                      to compare converged controls (1e-6) and objectives (1e-8 relative) against.

tests/test_casadi_reference.py uses it and SKIPS when `import casadi` fails (as it does in this image).
    python -m oracle.casadi_reference            # self-check against oracle/nlp.py (needs casadi)
"""
from __future__ import annotations

import numpy as np

try:  # optional dependency
    import casadi as ca
    HAVE_CASADI = True
except Exception:  # pragma: no cover - absent in the build image
    ca = None
    HAVE_CASADI = False

NO_BOUND = 10e30  # ModelParameters.cpp:21-24


def double_pendulum_sx():
    """examples/model_generate_example.cpp:66-131 / ex_model_generate.cpp: the double pendulum's equations of motion
    (L = m = 1, g = 9.81; the same closed form as workingMPC_8_30_21.m:27-38)."""
    qA, qB, qAd, qBd = (ca.SX.sym(n) for n in ("qA", "qB", "qA_dot", "qB_dot"))
    TA, TB = ca.SX.sym("TA"), ca.SX.sym("TB")
    L, m, g = 1.0, 1.0, 9.81
    den = L ** 2 * m * (ca.cos(qB) ** 2 - 2)
    qA_dd = -(TA - TB - TB * ca.cos(qB) + L ** 2 * m * qAd ** 2 * ca.sin(qB) + L ** 2 * m * qBd ** 2 * ca.sin(qB) - 2 * L * g * m * ca.cos(qA)
              + L ** 2 * m * qAd ** 2 * ca.cos(qB) * ca.sin(qB) + 2 * L ** 2 * m * qAd * qBd * ca.sin(qB) + L * g * m * ca.cos(qA + qB) * ca.cos(qB)) / den
    qB_dd = (TA - 3 * TB + TA * ca.cos(qB) - 2 * TB * ca.cos(qB) + 2 * L * g * m * ca.cos(qA + qB) + 3 * L ** 2 * m * qAd ** 2 * ca.sin(qB)
             + L ** 2 * m * qBd ** 2 * ca.sin(qB) - 2 * L * g * m * ca.cos(qA) + 2 * L ** 2 * m * qAd ** 2 * ca.cos(qB) * ca.sin(qB)
             + L ** 2 * m * qBd ** 2 * ca.cos(qB) * ca.sin(qB) - 2 * L * g * m * ca.cos(qA) * ca.cos(qB) + 2 * L ** 2 * m * qAd * qBd * ca.sin(qB)
             + L * g * m * ca.cos(qA + qB) * ca.cos(qB) + 2 * L ** 2 * m * qAd * qBd * ca.cos(qB) * ca.sin(qB)) / den
    x = ca.vertcat(qA, qB, qAd, qBd)
    u = ca.vertcat(TA, TB)
    return x, u, ca.vertcat(qAd, qBd, qA_dd, qB_dd)


class ReferenceNLP:
    """ModelGenerator::create_model (ModelGenerator.cpp:23-233) for one (x, x_dot, u) triple."""

    def __init__(self, N=25, h=0.002, is_linear=False, x_min=None, x_max=None, u_min=None, u_max=None, model=double_pendulum_sx):
        if not HAVE_CASADI:
            raise ImportError("casadi is not installed")
        x, u, x_dot = model()
        nx, nu = x.shape[0], u.shape[0]
        self.nx, self.nu, self.N, self.h, self.is_linear = nx, nu, N, h, is_linear
        F = ca.Function("F", [x, u], [x + x_dot * h], ["x", "u"], ["x_next"])                    # :33-34
        A_sym, B_sym = ca.SX.sym("A_sym", nx, nx), ca.SX.sym("B_sym", nx, nu)                    # :38-42
        x_init_sym, u_init_sym, x_dot_init_sym = ca.SX.sym("x_init_sym", nx), ca.SX.sym("u_init_sym", nu), ca.SX.sym("x_dot_init_sym", nx)
        A, B = ca.jacobian(x_dot, x), ca.jacobian(x_dot, u)                                       # :45-46
        x_dot_lin = ca.mtimes(A_sym, x - x_init_sym) + ca.mtimes(B_sym, u - u_init_sym) + x_dot_init_sym  # :47
        self.get_A = ca.Function("get_A", [x, u], [A])                                            # :51-53
        self.get_B = ca.Function("get_B", [x, u], [B])
        self.get_x_dot_init = ca.Function("get_x_dot_init", [x, u], [x_dot])
        F_lin = ca.Function("F_lin", [A_sym, B_sym, x, u, x_dot_init_sym, x_init_sym, u_init_sym], [x + x_dot_lin * h])  # :58
        NV = nx * (N + 1) + nu * N                                                                # :61
        self.NV = NV
        x_min = np.full(nx, -NO_BOUND) if x_min is None else np.asarray(x_min, float)
        x_max = np.full(nx, NO_BOUND) if x_max is None else np.asarray(x_max, float)
        u_min = np.full(nu, -NO_BOUND) if u_min is None else np.asarray(u_min, float)
        u_max = np.full(nu, NO_BOUND) if u_max is None else np.asarray(u_max, float)
        V = ca.MX.sym("V", NV)                                                                    # :76
        v_min, v_max, X, U, off = [], [], [], [], 0
        for k in range(N):                                                                        # :86-105
            X.append(V[off:off + nx]); off += nx
            v_min += ([0.0] * nx if k == 0 else list(x_min)); v_max += ([0.0] * nx if k == 0 else list(x_max))
            U.append(V[off:off + nu]); off += nu
            v_min += list(u_min); v_max += list(u_max)
        X.append(V[off:off + nx])                                                                 # :108-112
        v_min += [-np.inf] * nx; v_max += [np.inf] * nx
        self.v_min, self.v_max = np.array(v_min), np.array(v_max)
        np_ = N * nx + nx + 2 * nu + (nx * nx + nx * nu + 2 * nx if is_linear else 0) + nu          # :129-141
        self.np_ = np_
        traj = ca.MX.sym("traj", np_)                                                             # :145
        sQ = N * nx; sR = sQ + nx; sRm = sR + nu; eRm = sRm + nu                                   # :147-150
        Q, R, Rm = ca.diag(traj[sQ:sR]), ca.diag(traj[sR:sRm]), ca.diag(traj[sRm:eRm])             # :152-157
        if is_linear:                                                                             # :167-180 (reshape = column-major)
            sA = eRm; sB = sA + nx * nx; sxd = sB + nx * nu; sxi = sxd + nx; sui = sxi + nx
            lin_A, lin_B = ca.reshape(traj[sA:sB], nx, nx), ca.reshape(traj[sB:sxd], nx, nu)
            x_dot_init, x_init_in, u_init_in = traj[sxd:sxi], traj[sxi:sui], traj[sui:sui + nu]
        else:
            u_init_in = traj[eRm:eRm + nu]                                                        # :182-186
        J, g = 0, []
        for k in range(N):                                                                        # :190-222
            cur = F_lin(lin_A, lin_B, X[k], U[k], x_dot_init, x_init_in, u_init_in) if is_linear else F(X[k], U[k])
            g.append(cur - X[k + 1])                                                              # :206
            err = cur - traj[k * nx:(k + 1) * nx]                                                 # :207-211
            J = J + ca.mtimes(err.T, ca.mtimes(Q, err))                                           # :214
            dU = U[k] - (u_init_in if k == 0 else U[k - 1])                                       # :217
            J = J + ca.mtimes(dU.T, ca.mtimes(R, dU)) + ca.mtimes(U[k].T, ca.mtimes(Rm, U[k]))    # :218-221
        self.nlp = {"x": V, "f": J, "g": ca.vertcat(*g), "p": traj}                               # :225-226
        self._V, self._p, self._J, self._g = V, traj, J, ca.vertcat(*g)
        self._solver = None

    # ---- the five functions generate_dependencies() emits (ModelGenerator.cpp:235-239) ----
    def nlp_functions(self):
        V, p, J, g = self._V, self._p, self._J, self._g
        lam_f, lam_g = ca.MX.sym("lam_f"), ca.MX.sym("lam_g", g.shape[0])
        Lg = lam_f * J + ca.dot(lam_g, g)
        H = ca.triu(ca.hessian(Lg, V)[0])
        return dict(nlp_f=ca.Function("nlp_f", [V, p], [J]), nlp_g=ca.Function("nlp_g", [V, p], [g]),
                    nlp_grad_f=ca.Function("nlp_grad_f", [V, p], [J, ca.gradient(J, V)]),
                    nlp_jac_g=ca.Function("nlp_jac_g", [V, p], [g, ca.jacobian(g, V)]),
                    nlp_hess_l=ca.Function("nlp_hess_l", [V, p, lam_f, lam_g], [H]))

    def pack_p(self, traj, Q, R, Rm, u_init, lin=None):
        parts = [np.asarray(traj, float).ravel(), Q, R, Rm]
        if self.is_linear:
            A, B, xd, xi = lin  # A, B column-major flat, as ModelControl.cpp:127-143 inserts them
            parts += [A, B, xd, xi]
        parts.append(u_init)
        return np.concatenate([np.asarray(a, float).ravel() for a in parts])

    def solve(self, state, p, V0=None, tol=1e-5, max_iter=200):
        """ModelControl::calc_u (ModelControl.cpp:116-172): pin x_0, solve, return V, lam_g, f, IPOPT's iteration count."""
        if self._solver is None or self._solver[0] != (tol, max_iter):
            opts = {"ipopt.tol": tol, "ipopt.max_iter": max_iter, "ipopt.linear_solver": "mumps", "ipopt.print_level": 0,
                    "print_time": 0, "ipopt.sb": "yes"}                                          # ModelControl.cpp:54-59
            self._solver = ((tol, max_iter), ca.nlpsol("nlpsol", "ipopt", self.nlp, opts))
        lbx, ubx = self.v_min.copy(), self.v_max.copy()
        lbx[:self.nx] = state; ubx[:self.nx] = state                                              # :144-157
        x0 = np.zeros(self.NV) if V0 is None else np.asarray(V0, float)                            # :68 cold / :161 warm
        r = self._solver[1](x0=x0, lbx=lbx, ubx=ubx, lbg=0, ubg=0, p=p)
        st = self._solver[1].stats()
        return dict(V=np.array(r["x"]).ravel(), lam_g=np.array(r["lam_g"]).ravel(), f=float(r["f"]), iters=int(st.get("iter_count", -1)),
                    success=bool(st.get("success", False)), return_status=st.get("return_status"))


def self_check(N=25, B=4, tol=1e-9):
    """CasADi's functions and IPOPT's solutions against oracle/nlp.py on the bench workload; returns the largest deviations."""
    from oracle import nlp
    P = nlp.Problem(N=N)
    ref = ReferenceNLP(N=N, h=P.h)
    fn = ref.nlp_functions()
    Q, R, Rm = nlp.default_weights(4, 2)
    x0, up, traj = nlp.synthetic_batch(4, 2, N, P.h, B)
    rng = np.random.default_rng(0)
    dev = dict(f=0.0, g=0.0, grad_f=0.0, jac_g=0.0, hess_l=0.0, V=0.0, obj=0.0)
    for b in range(B):
        p = P.pack_p(traj[b], Q, R, Rm, up[b])
        assert np.array_equal(p, ref.pack_p(traj[b], Q, R, Rm, up[b]))
        V = rng.uniform(-0.5, 0.5, P.NV); lam = rng.uniform(-1, 1, P.M)
        f, gf = P.nlp_grad_f(V, p); g, Jg = P.nlp_jac_g(V, p); H = P.nlp_hess_l(V, p, 1.0, lam)
        dev["f"] = max(dev["f"], abs(float(fn["nlp_f"](V, p)) - f) / abs(f))
        dev["g"] = max(dev["g"], np.abs(np.array(fn["nlp_g"](V, p)).ravel() - g).max())
        dev["grad_f"] = max(dev["grad_f"], np.abs(np.array(fn["nlp_grad_f"](V, p)[1]).ravel() - gf).max() / np.abs(gf).max())
        dev["jac_g"] = max(dev["jac_g"], np.abs(np.array(fn["nlp_jac_g"](V, p)[1]) - Jg).max())
        Hc = np.array(fn["nlp_hess_l"](V, p, 1.0, lam)); Hc = Hc + np.triu(Hc, 1).T
        dev["hess_l"] = max(dev["hess_l"], np.abs(Hc - H).max() / np.abs(H).max())
        s = ref.solve(x0[b], p, tol=tol)
        o = nlp.solve(P, x0[b], p, tol=tol, max_iter=400)
        dev["V"] = max(dev["V"], np.abs(s["V"] - o.V).max())
        dev["obj"] = max(dev["obj"], abs(s["f"] - o.f) / abs(o.f))
    return dev


if __name__ == "__main__":
    print(self_check())
