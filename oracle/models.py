"""ORACLE (test infrastructure, never shipped, never on the product path).

Symbolic (sympy) statement of the dynamics models the reference's drivers use, and the
derivative objects the reference obtains from CasADi AD:

  f(x,u)                       user ODE right-hand side   examples/ex_model_generate.cpp:28-43
  A = d f/d x, B = d f/d u     casadi::jacobian           src/Mahi/Mpc/ModelGenerator.cpp:45-46
  Hc(w) = sum_i w_i grad^2 f_i  the part of nlp_hess_l that comes from the dynamics
                               (CasADi derives it inside nlpsol)

Parity status: UNPINNED against CasADi/IPOPT (neither is installable here).
What *is* pinned: f, A, B agree with the reference's own hand-expanded closed forms in
old/Models/DoublePendulumModel.hpp:15-132 (compiled from where they lie into oracle/_ref by
oracle/ref_build/Makefile) and with 50-digit mpmath evaluation of the same expressions.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package.
"""
from __future__ import annotations

import functools
from dataclasses import dataclass, field

import numpy as np
import sympy as sp


@dataclass
class SymModel:
    name: str
    x: list
    u: list
    f: list  # xdot expressions
    nx: int = field(init=False)
    nu: int = field(init=False)

    def __post_init__(self):
        self.nx = len(self.x)
        self.nu = len(self.u)

    @property
    def z(self):
        return list(self.x) + list(self.u)


def double_pendulum_sym() -> SymModel:
    """Double pendulum, L = m = 1, g = 9.81 (examples/ex_model_generate.cpp:24-43).

    Written in a factored form (torque terms, velocity terms and gravity terms grouped);
    tests/test_oracle_models.py checks it is symbolically identical to the expanded
    expression text at ex_model_generate.cpp:36-37.
    """
    qA, qB, qAd, qBd, TA, TB = sp.symbols("qA qB qA_dot qB_dot TA TB", real=True)
    L, m, g = sp.Integer(1), sp.Integer(1), sp.Float("9.81", 17)
    cA, cB, sB, cAB = sp.cos(qA), sp.cos(qB), sp.sin(qB), sp.cos(qA + qB)
    mL2 = m * L * L
    gmL = g * m * L
    den = mL2 * (cB * cB - 2)
    numA = (TA - TB * (1 + cB)
            + mL2 * sB * (qAd**2 * (1 + cB) + qBd**2 + 2 * qAd * qBd)
            - gmL * (2 * cA - cAB * cB))
    numB = (TA * (1 + cB) - TB * (3 + 2 * cB)
            + mL2 * sB * (qAd**2 * (3 + 2 * cB) + (qBd**2 + 2 * qAd * qBd) * (1 + cB))
            + gmL * (cAB * (2 + cB) - 2 * cA * (1 + cB)))
    return SymModel("double_pendulum", [qA, qB, qAd, qBd], [TA, TB],
                    [qAd, qBd, -numA / den, numB / den])


def double_pendulum_expanded_reference_text():
    """The expression exactly as the reference driver spells it
    (examples/ex_model_generate.cpp:36-37), used only to prove the factored form equal."""
    qA, qB, qA_dot, qB_dot, TA, TB = sp.symbols("qA qB qA_dot qB_dot TA TB", real=True)
    L, m, g = sp.Integer(1), sp.Integer(1), sp.Float("9.81", 17)
    cos, sin = sp.cos, sp.sin
    qA_ddot = -(TA - TB - TB*cos(qB) + L*L*m*qA_dot*qA_dot*sin(qB) + L*L*m*qB_dot*qB_dot*sin(qB) - 2*L*g*m*cos(qA) + L*L*m*qA_dot*qA_dot*cos(qB)*sin(qB) + 2*L*L*m*qA_dot*qB_dot*sin(qB) + L*g*m*cos(qA + qB)*cos(qB))/(L*L*m*(cos(qB)*cos(qB) - 2))
    qB_ddot = (TA - 3*TB + TA*cos(qB) - 2*TB*cos(qB) + 2*L*g*m*cos(qA + qB) + 3*L*L*m*qA_dot*qA_dot*sin(qB) + L*L*m*qB_dot*qB_dot*sin(qB) - 2*L*g*m*cos(qA) + 2*L*L*m*qA_dot*qA_dot*cos(qB)*sin(qB) + L*L*m*qB_dot*qB_dot*cos(qB)*sin(qB) - 2*L*g*m*cos(qA)*cos(qB) + 2*L*L*m*qA_dot*qB_dot*sin(qB) + L*g*m*cos(qA + qB)*cos(qB) + 2*L*L*m*qA_dot*qB_dot*cos(qB)*sin(qB))/(L*L*m*(cos(qB)*cos(qB) - 2))
    return [qA_dot, qB_dot, qA_ddot, qB_ddot]


# ---------------------------------------------------------------------------------------------
# 4-DOF arm (BASELINE config 3).  NOT in the reference as a runnable model: only the symbolic
# mass matrix exists (src/inverseTest.cpp:59-83).  This is the builder-defined stand-in
# documented in DESIGN.md: a planar 4-link serial chain with point masses at the link tips,
# x = [q; qdot], u = tau, qddot = M(q)^-1 (tau - C(q,qdot) qdot - G(q)), C from the
# Christoffel symbols of M.  Same definition is restated in C++ in the product's built-in
# model (mahi_mpc_b200/csrc/models) through the product's own symbolic engine.
# ---------------------------------------------------------------------------------------------
ARM4_LINK_LENGTH = (0.30, 0.25, 0.20, 0.15)
ARM4_LINK_MASS = (2.0, 1.5, 1.0, 0.5)
ARM4_GRAVITY = 9.81
ARM4_DAMPING = (0.5, 0.4, 0.3, 0.2)


def arm4_sym() -> SymModel:
    n = 4
    q = sp.symbols("q0:4", real=True)
    qd = sp.symbols("qd0:4", real=True)
    tau = sp.symbols("tau0:4", real=True)
    l = [sp.Float(v, 17) for v in ARM4_LINK_LENGTH]
    m = [sp.Float(v, 17) for v in ARM4_LINK_MASS]
    d = [sp.Float(v, 17) for v in ARM4_DAMPING]
    g = sp.Float(ARM4_GRAVITY, 17)
    # absolute link angles and tip positions
    th = [sum(q[: i + 1]) for i in range(n)]
    px, py = [], []
    cx, cy = 0, 0
    for i in range(n):
        cx = cx + l[i] * sp.cos(th[i])
        cy = cy + l[i] * sp.sin(th[i])
        px.append(cx)
        py.append(cy)
    # kinetic / potential energy of tip point masses
    M = sp.zeros(n, n)
    V = 0
    for i in range(n):
        Jx = sp.Matrix([[sp.diff(px[i], qj) for qj in q]])
        Jy = sp.Matrix([[sp.diff(py[i], qj) for qj in q]])
        M += m[i] * (Jx.T * Jx + Jy.T * Jy)
        V += m[i] * g * py[i]
    M = sp.simplify(M)
    G = [sp.diff(V, qj) for qj in q]
    # Coriolis through Christoffel symbols
    Cq = []
    for i in range(n):
        acc = 0
        for j in range(n):
            for k in range(n):
                c = (sp.diff(M[i, j], q[k]) + sp.diff(M[i, k], q[j]) - sp.diff(M[j, k], q[i])) / 2
                acc += c * qd[j] * qd[k]
        Cq.append(acc)
    rhs = sp.Matrix([tau[i] - Cq[i] - G[i] - d[i] * qd[i] for i in range(n)])
    qdd = M.LUsolve(rhs)
    return SymModel("arm4", list(q) + list(qd), list(tau), list(qd) + [qdd[i] for i in range(n)])


MODELS = {"double_pendulum": double_pendulum_sym, "arm4": arm4_sym}


# ---------------------------------------------------------------------------------------------
# derivative objects
# ---------------------------------------------------------------------------------------------
@functools.lru_cache(maxsize=None)
def derive(name: str):
    """Return dict(model, A, B, w, Hc) of sympy objects for model `name`."""
    mdl = MODELS[name]()
    f = sp.Matrix(mdl.f)
    A = f.jacobian(mdl.x)
    B = f.jacobian(mdl.u)
    w = sp.symbols(f"w0:{mdl.nx}", real=True)
    lag = sum(w[i] * mdl.f[i] for i in range(mdl.nx))
    Hc = sp.hessian(lag, mdl.z)
    return dict(model=mdl, A=A, B=B, w=list(w), Hc=Hc)


def _numeric_arm4():
    """arm4: sympy is far too slow for its second derivatives (expression swell), so the numpy oracle borrows the C
    oracle's forward-mode AD restatement (oracle/arm4_jets.c) through oracle/coracle.py."""
    from . import coracle
    co = coracle.COracle("O3")

    def _call(x, u, w=None):
        x = np.atleast_2d(np.asarray(x, float).T); u = np.atleast_2d(np.asarray(u, float).T)
        wt = None if w is None else np.atleast_2d(np.asarray(w, float).T)
        return co.node_eval("arm4", 1.0, x, u, np.zeros_like(x), wt), x

    def f_np(x, u):
        (g, _, _, _), xt = _call(x, u)
        return (g - xt).T

    def A_np(x, u):
        (_, Fx, _, _), _ = _call(x, u)
        return np.moveaxis(Fx - np.eye(8), 0, 2)

    def B_np(x, u):
        (_, _, Fu, _), _ = _call(x, u)
        return np.moveaxis(Fu, 0, 2)

    def Hc_np(x, u, w):
        (_, _, _, H), _ = _call(x, u, w)
        return np.moveaxis(H, 0, 2)

    return dict(nx=8, nu=4, f=f_np, A=A_np, B=B_np, Hc=Hc_np)


@functools.lru_cache(maxsize=None)
def numeric(name: str):
    """numpy-callable f, A, B, Hc (vectorised over a trailing batch axis)."""
    if name == "arm4":
        return _numeric_arm4()
    d = derive(name)
    mdl = d["model"]
    args = mdl.z
    f = sp.lambdify(args, mdl.f, "numpy", cse=True)
    A = sp.lambdify(args, d["A"].tolist(), "numpy", cse=True)
    B = sp.lambdify(args, d["B"].tolist(), "numpy", cse=True)
    Hc = sp.lambdify(args + d["w"], d["Hc"].tolist(), "numpy", cse=True)

    # complex arguments pass through (oracle/nlp.py differentiates the RK4 sensitivities by the complex-step method)
    def _arr(v, shape_like):
        dt = np.complex128 if np.iscomplexobj(shape_like) else np.float64
        return np.array([[np.broadcast_to(np.asarray(e, dtype=dt), shape_like.shape)
                          for e in row] for row in v]) if isinstance(v[0], (list, tuple)) else \
            np.array([np.broadcast_to(np.asarray(e, dtype=dt), shape_like.shape) for e in v])

    def _in(*vs):
        cplx = any(np.iscomplexobj(v) for v in vs)
        return [np.asarray(v, np.complex128 if cplx else np.float64) for v in vs]

    def f_np(x, u):
        x, u = _in(x, u)
        return _arr(f(*x, *u), x[0])

    def A_np(x, u):
        x, u = _in(x, u)
        return _arr(A(*x, *u), x[0])

    def B_np(x, u):
        x, u = _in(x, u)
        return _arr(B(*x, *u), x[0])

    def Hc_np(x, u, w):
        x, u, w = _in(x, u, w)
        return _arr(Hc(*x, *u, *w), x[0])

    return dict(nx=mdl.nx, nu=mdl.nu, f=f_np, A=A_np, B=B_np, Hc=Hc_np)


def mp_eval(name: str, x, u, w=None, digits: int = 50):
    """High-precision (mpmath) evaluation of f, A, B and Hc at one point -> python floats."""
    import mpmath as mp

    d = derive(name)
    mdl = d["model"]
    with mp.workdps(digits):
        sub = {s: sp.Float(repr(float(v)), digits) if False else sp.Float(float(v), digits)
               for s, v in zip(mdl.z, list(x) + list(u))}
        # exact binary value of the doubles: sp.Float(float, digits) keeps the double exactly
        fv = [float(sp.N(e.subs(sub), digits)) for e in mdl.f]
        Av = [[float(sp.N(e.subs(sub), digits)) for e in row] for row in d["A"].tolist()]
        Bv = [[float(sp.N(e.subs(sub), digits)) for e in row] for row in d["B"].tolist()]
        Hv = None
        if w is not None:
            subw = dict(sub)
            subw.update({s: sp.Float(float(v), digits) for s, v in zip(d["w"], w)})
            Hv = [[float(sp.N(e.subs(subw), digits)) for e in row] for row in d["Hc"].tolist()]
    return dict(f=np.array(fv), A=np.array(Av), B=np.array(Bv), Hc=None if Hv is None else np.array(Hv))
