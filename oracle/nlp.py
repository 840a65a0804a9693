"""ORACLE (test infrastructure, never shipped, never on the product path).

numpy restatement of the multiple-shooting NLP that mahi::mpc::ModelGenerator::create_model
builds symbolically (src/Mahi/Mpc/ModelGenerator.cpp:23-233) and that ModelControl::calc_u
hands to IPOPT each cycle (src/Mahi/Mpc/ModelControl.cpp:116-172).

Layouts (all fp64, exactly the reference's):
  V = [x_0,u_0,x_1,u_1,...,x_{N-1},u_{N-1},x_N]                 ModelGenerator.cpp:61,76-112
  p = [r_0..r_{N-1} | Q diag | R diag | Rm diag | {A colmaj | B colmaj | xdot_init | x_init}
       only if is_linear | u_init]                              ModelGenerator.cpp:129-187
  g = [F(x_k,u_k) - x_{k+1}]_{k<N}                              ModelGenerator.cpp:206,225
  J = sum_k e_k'Qe_k + du_k'R du_k + u_k'Rm u_k,  e_k = F(x_k,u_k) - r_k,
      du_k = u_k - (k==0 ? u_init : u_{k-1})                    ModelGenerator.cpp:207-221
  F(x,u) = x + h f(x,u)  (explicit Euler)                       ModelGenerator.cpp:33-34
  F_lin  = x + h (A (x-x_init) + B (u-u_init) + xdot_init)      ModelGenerator.cpp:47-48,58

The solver below is NOT IPOPT: it is a plain full-space Newton / primal-dual interior point
iteration on the same KKT conditions with a dense/sparse LU (scipy), i.e. an independent
route to the same KKT point.  Termination restates IPOPT's scaled optimality error
(Waechter & Biegler 2006, eq. 5-6) [upstream].  Parity vs CasADi/IPOPT itself: UNPINNED.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import models as _models

NO_BOUND = 10e30  # ModelParameters.cpp:21-24 sentinel
INF_THRESH = 1e19  # IPOPT nlp_{lower,upper}_bound_inf [upstream]


@dataclass
class Problem:
    model: str = "double_pendulum"
    N: int = 25  # ex_model_generate.cpp:57
    h: float = 0.002  # ex_model_generate.cpp:56
    is_linear: bool = False
    integrator: str = "euler"  # "euler" (library parity) | "rk4"
    hessian: str = "exact"     # "exact" | "gauss_newton" (RK4 only: drop the integrator's second derivatives)
    # "shooting": tracking error on F(x_k, u_k) - r_k, the C++ library's objective (ModelGenerator.cpp:207-214);
    # "node": on the decision variable, x_{k+1} - r_k, the MATLAB prototypes' (workingMPC_8_30_21.m:90-106: the error of x(:,k)
    # against the reference at its own time, with Q <- 1000 Q, no du term, Rm <- R there; x_0 is pinned, its term is a constant)
    cost: str = "shooting"
    substeps: int = 1

    def __post_init__(self):
        m = _models.numeric(self.model)
        self.m = m
        self.nx, self.nu = m["nx"], m["nu"]
        self.s = self.nx + self.nu
        self.NV = self.nx * (self.N + 1) + self.nu * self.N
        self.M = self.N * self.nx
        self.np_ = self.N * self.nx + self.nx + 2 * self.nu + self.nu
        if self.is_linear:
            self.np_ += self.nx * self.nx + self.nx * self.nu + 2 * self.nx

    # ---- layout helpers ---------------------------------------------------------------
    def split_V(self, V):
        nx, nu, N, s = self.nx, self.nu, self.N, self.s
        body = V[: N * s].reshape(N, s)
        X = np.vstack([body[:, :nx], V[N * s:][None, :]])
        U = body[:, nx:]
        return X, U

    def join_V(self, X, U):
        return np.concatenate([np.hstack([X[:-1], U]).ravel(), X[-1]])

    def pack_p(self, traj, Q, R, Rm, u_init, lin=None):
        parts = [np.asarray(traj, float).ravel(), Q, R, Rm]
        if self.is_linear:
            A, B, xdot0, x_init = lin
            parts += [np.asarray(A).ravel(order="F"), np.asarray(B).ravel(order="F"), xdot0, x_init]
        parts.append(u_init)
        p = np.concatenate([np.asarray(a, float).ravel() for a in parts])
        assert p.size == self.np_, (p.size, self.np_)
        return p

    def split_p(self, p):
        nx, nu, N = self.nx, self.nu, self.N
        o = N * nx
        d = dict(r=p[:o].reshape(N, nx), Q=p[o:o + nx], R=p[o + nx:o + nx + nu],
                 Rm=p[o + nx + nu:o + nx + 2 * nu])
        o += nx + 2 * nu
        if self.is_linear:
            d["A"] = p[o:o + nx * nx].reshape(nx, nx, order="F"); o += nx * nx
            d["B"] = p[o:o + nx * nu].reshape(nx, nu, order="F"); o += nx * nu
            d["xdot_init"] = p[o:o + nx]; o += nx
            d["x_init"] = p[o:o + nx]; o += nx
        d["u_init"] = p[o:o + nu]
        return d

    def linearisation(self, state, control):
        """<name>_get_A/_get_B/_get_x_dot_init (ModelGenerator.cpp:51-53; used ModelControl.cpp:125-135)."""
        x = np.asarray(state, float)[:, None]
        u = np.asarray(control, float)[:, None]
        return (self.m["A"](x, u)[:, :, 0], self.m["B"](x, u)[:, :, 0], self.m["f"](x, u)[:, 0],
                np.asarray(state, float))

    # ---- one-step map and its derivatives, all nodes at once ---------------------------
    def _ode(self, X, U, P, want, W=None):
        """xdot, A, B, Hc for columns of X (nx,n), U (nu,n)."""
        if self.is_linear:
            A, B = P["A"], P["B"]
            n = X.shape[1]
            f = A @ (X - P["x_init"][:, None]) + B @ (U - P["u_init"][:, None]) + P["xdot_init"][:, None]
            out = [f]
            if want >= 1:
                out += [np.repeat(A[:, :, None], n, 2), np.repeat(B[:, :, None], n, 2)]
            if want >= 2:
                out.append(np.zeros((self.s, self.s, n)))
            return out
        out = [self.m["f"](X, U)]
        if want >= 1:
            out += [self.m["A"](X, U), self.m["B"](X, U)]
        if want >= 2:
            out.append(self.m["Hc"](X, U, W))
        return out

    def step(self, X, U, P):
        """F for every node; X (N,nx), U (N,nu) -> (N,nx)."""
        if self.integrator == "euler":
            (f,) = self._ode(X.T, U.T, P, 0)
            return X + self.h * f.T
        hs = self.h / self.substeps
        x = X.T.copy()
        for _ in range(self.substeps):
            k1 = self._ode(x, U.T, P, 0)[0]
            k2 = self._ode(x + 0.5 * hs * k1, U.T, P, 0)[0]
            k3 = self._ode(x + 0.5 * hs * k2, U.T, P, 0)[0]
            k4 = self._ode(x + hs * k3, U.T, P, 0)[0]
            x = x + hs / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
        return x.T

    def step_jac(self, X, U, P):
        """F, Fx (N,nx,nx), Fu (N,nx,nu)."""
        nx = self.nx
        if self.integrator == "euler":
            f, A, B = self._ode(X.T, U.T, P, 1)
            F = X + self.h * f.T
            Fx = np.eye(nx)[None] + self.h * np.moveaxis(A, 2, 0)
            Fu = self.h * np.moveaxis(B, 2, 0)
            return F, Fx, Fu
        # RK4 with forward sensitivities
        n = X.shape[0]
        hs = self.h / self.substeps
        x = X.T.copy()
        Sx = np.repeat(np.eye(nx)[None], n, 0)
        Su = np.zeros((n, nx, self.nu))
        Ut = U.T

        def stage(xs, dSx, dSu):
            f, A, B = self._ode(xs, Ut, P, 1)
            A = np.moveaxis(A, 2, 0); B = np.moveaxis(B, 2, 0)
            return f, A @ dSx, A @ dSu + B

        for _ in range(self.substeps):
            k1, k1x, k1u = stage(x, Sx, Su)
            k2, k2x, k2u = stage(x + 0.5 * hs * k1, Sx + 0.5 * hs * k1x, Su + 0.5 * hs * k1u)
            k3, k3x, k3u = stage(x + 0.5 * hs * k2, Sx + 0.5 * hs * k2x, Su + 0.5 * hs * k2u)
            k4, k4x, k4u = stage(x + hs * k3, Sx + hs * k3x, Su + hs * k3u)
            x = x + hs / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
            Sx = Sx + hs / 6.0 * (k1x + 2 * k2x + 2 * k3x + k4x)
            Su = Su + hs / 6.0 * (k1u + 2 * k2u + 2 * k3u + k4u)
        return x.T, Sx, Su

    def step_hess(self, X, U, P, W):
        """sum_i W[k,i] grad^2_zz F_i for every node -> (N,s,s).
        Euler: h times the model's Hessian contraction (library parity).  RK4 (workingMPC_8_30_21.m:46-59): the COMPLEX-STEP
        derivative of the first-order sensitivities, H[:, j] = Im(Fz(z + i eps e_j)' w) / eps -- exact to rounding (no
        subtraction), and a different route from the CUDA path's forward-over-adjoint sweep (mmpc_kernels.cuh: step_rk4_hess)."""
        if self.integrator == "euler":
            f, A, B, Hc = self._ode(X.T, U.T, P, 2, W.T)
            return self.h * np.moveaxis(Hc, 2, 0)
        if self.hessian == "gauss_newton":
            return np.zeros((X.shape[0], self.s, self.s))
        n, nx, s = X.shape[0], self.nx, self.s
        eps = 1e-30
        H = np.zeros((n, s, s))
        Z = np.concatenate([X, U], axis=1).astype(complex)
        for j in range(s):
            Zp = Z.copy()
            Zp[:, j] += 1j * eps
            _, Fx, Fu = self.step_jac(Zp[:, :nx], Zp[:, nx:], P)
            Fz = np.concatenate([Fx, Fu], axis=2)  # (n, nx, s)
            H[:, :, j] = np.einsum("kis,ki->ks", Fz, W).imag / eps
        return 0.5 * (H + np.swapaxes(H, 1, 2))

    # ---- the five CasADi NLP functions ------------------------------------------------
    def nlp_g(self, V, p):
        X, U = self.split_V(V)
        P = self.split_p(p)
        return (self.step(X[:-1], U, P) - X[1:]).ravel()

    def nlp_f(self, V, p):
        X, U = self.split_V(V)
        P = self.split_p(p)
        e = (X[1:] if self.cost == "node" else self.step(X[:-1], U, P)) - P["r"]
        dU = U - np.vstack([P["u_init"][None], U[:-1]])
        return float((e * e * P["Q"]).sum() + (dU * dU * P["R"]).sum() + (U * U * P["Rm"]).sum())

    def nlp_grad_f(self, V, p):
        nx, N = self.nx, self.N
        X, U = self.split_V(V)
        P = self.split_p(p)
        dU = U - np.vstack([P["u_init"][None], U[:-1]])
        if self.cost == "node":
            e = X[1:] - P["r"]
            gx = np.vstack([np.zeros((1, nx)), 2.0 * P["Q"] * e])  # d/dx_{k+1}, k = 0..N-1; x_0 carries none
            gu = 2 * P["R"] * dU + 2 * P["Rm"] * U
            gu[:-1] -= 2 * P["R"] * dU[1:]
            f = float((e * e * P["Q"]).sum() + (dU * dU * P["R"]).sum() + (U * U * P["Rm"]).sum())
            return f, self.join_V(gx, gu)
        F, Fx, Fu = self.step_jac(X[:-1], U, P)
        e = F - P["r"]
        qe = 2.0 * P["Q"] * e
        gx = np.einsum("kij,ki->kj", Fx, qe)
        gu = np.einsum("kij,ki->kj", Fu, qe) + 2 * P["R"] * dU + 2 * P["Rm"] * U
        gu[:-1] -= 2 * P["R"] * dU[1:]
        f = float((e * e * P["Q"]).sum() + (dU * dU * P["R"]).sum() + (U * U * P["Rm"]).sum())
        return f, self.join_V(np.vstack([gx, np.zeros((1, nx))]), gu)

    def nlp_jac_g(self, V, p):
        """dense (M, NV) Jacobian of g."""
        nx, nu, N, s = self.nx, self.nu, self.N, self.s
        X, U = self.split_V(V)
        P = self.split_p(p)
        F, Fx, Fu = self.step_jac(X[:-1], U, P)
        J = np.zeros((self.M, self.NV))
        for k in range(N):
            r0, c0 = k * nx, k * s
            J[r0:r0 + nx, c0:c0 + nx] = Fx[k]
            J[r0:r0 + nx, c0 + nx:c0 + s] = Fu[k]
            c1 = (k + 1) * s
            J[r0:r0 + nx, c1:c1 + nx] -= np.eye(nx)
        return (F - X[1:]).ravel(), J

    def nlp_hess_l(self, V, p, lam_f, lam_g):
        """dense symmetric (NV,NV) Hessian of lam_f*f + lam_g'g  (SURVEY.md 8 a-8)."""
        nx, nu, N, s = self.nx, self.nu, self.N, self.s
        X, U = self.split_V(V)
        P = self.split_p(p)
        F, Fx, Fu = self.step_jac(X[:-1], U, P)
        e = F - P["r"]
        lam = np.asarray(lam_g, float).reshape(N, nx)
        node = self.cost == "node"
        W = lam if node else 2.0 * lam_f * P["Q"] * e + lam
        H = np.zeros((self.NV, self.NV))
        Hc = self.step_hess(X[:-1], U, P, W)
        Fz = np.concatenate([Fx, Fu], axis=2)  # (N,nx,s)
        for k in range(N):
            c0 = k * s
            blk = Hc[k] if node else 2.0 * lam_f * (Fz[k].T * P["Q"]) @ Fz[k] + Hc[k]
            H[c0:c0 + s, c0:c0 + s] += blk
            if node:  # the tracking term sits on the decision variable x_{k+1}: a constant diagonal block
                ix = np.arange((k + 1) * s, (k + 1) * s + nx)
                H[ix, ix] += 2.0 * lam_f * P["Q"]
            iu = np.arange(c0 + nx, c0 + s)
            H[iu, iu] += 2 * lam_f * (P["R"] + P["Rm"])
            if k < N - 1:
                H[iu, iu] += 2 * lam_f * P["R"]
                ju = iu + s
                H[iu, ju] -= 2 * lam_f * P["R"]
                H[ju, iu] -= 2 * lam_f * P["R"]
        return H

    # ---- CCS sparsity, the way CasADi hands it to IPOPT (src/codegen_usage.cpp:118-132) --
    def jac_g_ccs(self):
        """(colind, row) of jac_g; structural pattern = union over random points."""
        return _ccs_of(self._struct_jac())

    def hess_l_ccs(self):
        return _ccs_of(np.triu(self._struct_hess()))

    def _struct_jac(self):
        rng = np.random.default_rng(7)
        pat = np.zeros((self.M, self.NV), bool)
        for _ in range(3):
            V = rng.uniform(-1, 1, self.NV); p = rng.uniform(0.5, 1.5, self.np_)
            pat |= self.nlp_jac_g(V, p)[1] != 0
        return pat

    def _struct_hess(self):
        rng = np.random.default_rng(8)
        pat = np.zeros((self.NV, self.NV), bool)
        for _ in range(3):
            V = rng.uniform(-1, 1, self.NV); p = rng.uniform(0.5, 1.5, self.np_)
            pat |= self.nlp_hess_l(V, p, 1.0, rng.uniform(-1, 1, self.M)) != 0
        return pat


def _ccs_of(pat):
    colind = [0]
    rows = []
    for j in range(pat.shape[1]):
        r = np.nonzero(pat[:, j])[0]
        rows.extend(r.tolist())
        colind.append(len(rows))
    return np.array(colind), np.array(rows)


# -------------------------------------------------------------------------------------------------
# solver: full-space Newton / primal-dual IPM on the KKT system (independent of the GPU's Riccati)
# -------------------------------------------------------------------------------------------------
@dataclass
class SolveResult:
    V: np.ndarray
    lam_g: np.ndarray
    f: float
    iters: int
    kkt_error: float
    status: int  # 0 converged, 1 max_iter, 2 numerical failure
    z_L: np.ndarray = None
    z_U: np.ndarray = None


def bounds_vectors(prob: Problem, state, x_min=None, x_max=None, u_min=None, u_max=None):
    """lbx/ubx the way ModelControl::load_model + calc_u build them (ModelControl.cpp:29-50,144-157)."""
    nx, nu, N, s = prob.nx, prob.nu, prob.N, prob.s
    x_min = np.full(nx, -NO_BOUND) if x_min is None else np.asarray(x_min, float)
    x_max = np.full(nx, NO_BOUND) if x_max is None else np.asarray(x_max, float)
    u_min = np.full(nu, -NO_BOUND) if u_min is None else np.asarray(u_min, float)
    u_max = np.full(nu, NO_BOUND) if u_max is None else np.asarray(u_max, float)
    lb = np.concatenate([np.tile(np.concatenate([x_min, u_min]), N), x_min])
    ub = np.concatenate([np.tile(np.concatenate([x_max, u_max]), N), x_max])
    lb[:nx] = state
    ub[:nx] = state
    return lb, ub


def kkt_error(prob, V, p, lam, lb, ub, zL, zU, mu, obj_scale=1.0):
    """IPOPT's scaled optimality error E_mu [upstream: Waechter&Biegler 2006 eq. (5)]."""
    nx = prob.nx
    f, gf = prob.nlp_grad_f(V, p)
    g, J = prob.nlp_jac_g(V, p)
    free = slice(nx, None)
    dual = obj_scale * gf[free] + J[:, free].T @ lam - zL[free] + zU[free]
    hasL = lb > -INF_THRESH
    hasU = ub < INF_THRESH
    compL = np.where(hasL, (V - lb) * zL - mu, 0.0)[free]
    compU = np.where(hasU, (ub - V) * zU - mu, 0.0)[free]
    n = V.size - nx
    m = g.size
    s_max = 100.0
    zsum = np.abs(zL[free]).sum() + np.abs(zU[free]).sum()
    s_d = max(s_max, (np.abs(lam).sum() + zsum) / (m + n)) / s_max
    s_c = max(s_max, zsum / n) / s_max if n else 1.0
    e_d = np.abs(dual).max() / s_d
    e_p = np.abs(g).max()
    e_c = max(np.abs(compL).max(initial=0.0), np.abs(compU).max(initial=0.0)) / s_c
    return max(e_d, e_p, e_c), (e_d, e_p, e_c)


def solve(prob: Problem, state, p, V0=None, lam0=None, tol=1e-5, max_iter=200,
          lb=None, ub=None, mu_init=0.1, verbose=False) -> SolveResult:
    """Newton / IPM from V0 (zeros = the reference's cold start, ModelControl.cpp:28-50,68)."""
    nx, NV, M = prob.nx, prob.NV, prob.M
    if lb is None:
        lb, ub = bounds_vectors(prob, state)
    V = np.zeros(NV) if V0 is None else np.array(V0, float)
    V[:nx] = state
    lam = np.zeros(M) if lam0 is None else np.array(lam0, float)
    free = np.arange(nx, NV)
    n = free.size
    hasL = (lb > -INF_THRESH); hasL[:nx] = False
    hasU = (ub < INF_THRESH); hasU[:nx] = False
    bounded = bool(hasL.any() or hasU.any())
    mu = mu_init if bounded else 0.0
    zL = np.zeros(NV); zU = np.zeros(NV)
    if bounded:
        # push the start into the interior (IPOPT bound_push/bound_frac = 1e-2) [upstream]
        k1 = k2 = 1e-2
        both = hasL & hasU
        pl = np.where(hasL, np.minimum(k1 * np.maximum(1, np.abs(lb)), np.where(both, k2 * (ub - lb), np.inf)), 0)
        pu = np.where(hasU, np.minimum(k1 * np.maximum(1, np.abs(ub)), np.where(both, k2 * (ub - lb), np.inf)), 0)
        V = np.where(hasL, np.maximum(V, lb + pl), V)
        V = np.where(hasU, np.minimum(V, ub - pu), V)
        zL[hasL] = 1.0; zU[hasU] = 1.0
    nu_merit = 1.0
    status, it = 1, 0
    err = np.inf
    for it in range(max_iter + 1):
        err0, _ = kkt_error(prob, V, p, lam, lb, ub, zL, zU, 0.0)
        if verbose:
            print(f"it {it:3d} f={prob.nlp_f(V, p):.12e} E0={err0:.3e} mu={mu:.2e}")
        err = err0
        if err0 <= tol:
            status = 0
            break
        if it == max_iter:
            break
        if bounded:
            # monotone barrier update (Fiacco-McCormick, IPOPT defaults kappa_eps=10, kappa_mu=.2, theta_mu=1.5)
            while True:
                emu, _ = kkt_error(prob, V, p, lam, lb, ub, zL, zU, mu)
                if emu <= 10.0 * mu and mu > tol / 10.0:
                    mu = max(tol / 10.0, min(0.2 * mu, mu ** 1.5))
                else:
                    break
        f, gf = prob.nlp_grad_f(V, p)
        g, J = prob.nlp_jac_g(V, p)
        H = prob.nlp_hess_l(V, p, 1.0, lam)
        dL = np.where(hasL, V - lb, 1.0); dU = np.where(hasU, ub - V, 1.0)
        Sigma = np.where(hasL, zL / dL, 0.0) + np.where(hasU, zU / dU, 0.0)
        gphi = gf - np.where(hasL, mu / dL, 0.0) + np.where(hasU, mu / dU, 0.0)
        Hf = H[np.ix_(free, free)] + np.diag(Sigma[free])
        Jf = J[:, free]
        delta = 0.0
        for _try in range(30):
            K = np.block([[Hf + delta * np.eye(n), Jf.T], [Jf, np.zeros((M, M))]])
            rhs = -np.concatenate([gphi[free], g])
            try:
                sol = np.linalg.solve(K, rhs)
            except np.linalg.LinAlgError:
                sol = None
            ok = sol is not None and np.all(np.isfinite(sol))
            if ok:
                # inertia check on the reduced Hessian via the curvature along d (cheap surrogate)
                d = sol[:n]
                curv = d @ (Hf + delta * np.eye(n)) @ d
                ok = curv > 0 or np.abs(d).max() == 0
            if ok:
                break
            delta = 1e-4 if delta == 0 else delta * 8
        else:
            status = 2
            break
        d = np.zeros(NV); d[free] = sol[:n]
        lam_new = sol[n:]
        dzL = np.where(hasL, (mu - zL * d) / dL - zL, 0.0) if bounded else zL * 0
        dzU = np.where(hasU, (mu + zU * d) / dU - zU, 0.0) if bounded else zU * 0
        a_max = 1.0; a_z = 1.0
        if bounded:
            tau = max(0.99, 1.0 - mu)
            with np.errstate(divide="ignore", invalid="ignore"):
                cand = np.concatenate([np.where(hasL & (d < 0), -tau * dL / d, np.inf),
                                       np.where(hasU & (d > 0), tau * dU / d, np.inf)])
                a_max = min(1.0, cand.min())
                candz = np.concatenate([np.where(hasL & (dzL < 0), -tau * zL / dzL, np.inf),
                                        np.where(hasU & (dzU < 0), -tau * zU / dzU, np.inf)])
                a_z = min(1.0, candz.min())

        def phi(Vt):
            b = 0.0
            if bounded:
                b = -mu * (np.log(np.where(hasL, Vt - lb, 1.0)).sum() + np.log(np.where(hasU, ub - Vt, 1.0)).sum())
            return prob.nlp_f(Vt, p) + b

        nu_merit = max(nu_merit, 1.1 * np.abs(lam_new).max() + 1e-8)
        theta0 = np.abs(g).sum()
        m0 = phi(V) + nu_merit * theta0
        D = gphi @ d - nu_merit * theta0
        a = a_max
        for _ls in range(40):
            Vt = V + a * d
            mt = phi(Vt) + nu_merit * np.abs(prob.nlp_g(Vt, p)).sum()
            if mt <= m0 + 1e-4 * a * D + 1e-12 * abs(m0):
                break
            a *= 0.5
        V = V + a * d
        lam = lam + a * (lam_new - lam)
        if bounded:
            zL = zL + a_z * dzL; zU = zU + a_z * dzU
            # IPOPT's kappa_sigma safeguard (eq. 16) [upstream]
            ks = 1e10
            dL = np.where(hasL, V - lb, 1.0); dU = np.where(hasU, ub - V, 1.0)
            zL = np.where(hasL, np.clip(zL, mu / (ks * dL), ks * mu / dL), 0.0)
            zU = np.where(hasU, np.clip(zU, mu / (ks * dU), ks * mu / dU), 0.0)
    return SolveResult(V=V, lam_g=lam, f=prob.nlp_f(V, p), iters=it, kkt_error=float(err), status=status,
                       z_L=zL, z_U=zU)


# -------------------------------------------------------------------------------------------------
# synthetic workload (mirrors examples/model_control_example.cpp:60-68, randomised)
# -------------------------------------------------------------------------------------------------
DEFAULT_Q = {4: [10.0, 1.0, 5.0, 5.0]}  # examples/thread_model_control_example.cpp:24
DEFAULT_R = {2: [5.0, 5.0]}  # :25


def default_weights(nx, nu):
    Q = np.array(DEFAULT_Q.get(nx, ([10.0, 1.0] * nx)[: nx // 2] + [5.0] * (nx - nx // 2)), float)
    R = np.array(DEFAULT_R.get(nu, [5.0] * nu), float)
    Rm = np.ones(nu)
    return Q, R, Rm


def synthetic_batch(nx, nu, N, h, B, seed=0x5EED, offset=0):
    """x0 (B,nx), u_prev (B,nu), traj (B,N,nx). Philox counter RNG, one independent stream
    position per instance so a shard [offset, offset+B) reproduces the same instances."""
    half = nx // 2
    assert half <= 6
    x0 = np.empty((B, nx))
    bg = np.random.Philox(key=seed)
    bg.advance(offset * 4)  # 4 counter blocks (=16 doubles worth of state) reserved per instance
    rng = np.random.Generator(bg)
    draws = rng.random((B, 16))
    x0[:, :half] = draws[:, :half] - 0.5
    x0[:, half:] = 2.0 * draws[:, 6:6 + nx - half] - 1.0
    amp = 0.25 + 0.75 * draws[:, 12]
    frq = 0.25 + 0.75 * draws[:, 13]
    phs = 2 * np.pi * draws[:, 14]
    t = np.arange(N) * h
    arg = 2 * np.pi * frq[:, None] * t[None, :] + phs[:, None]
    s, c = np.sin(arg), np.cos(arg)
    traj = np.empty((B, N, nx))
    for j in range(nx):
        sign = 1.0 if ((j - half) if j >= half else j) % 2 == 0 else -1.0
        traj[:, :, j] = sign * amp[:, None] * (s if j < half else 2 * np.pi * frq[:, None] * c)
    return x0, np.zeros((B, nu)), traj
