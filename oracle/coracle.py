"""ORACLE (test infrastructure). ctypes access to oracle/_build/liboracle_*.so (the C restatement,
oracle/mpc_oracle.c) and oracle/_ref/libdp_old_ref.so (the reference's own closed forms of the
double pendulum, old/Models/DoublePendulumModel.hpp:15-132, compiled in place)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)


class Stats(C.Structure):
    _fields_ = [("f", C.c_double), ("kkt_error", C.c_double), ("mu", C.c_double), ("iters", C.c_int),
                ("status", C.c_int)]


def build(quiet=True):
    """Compile the C restatement (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


class COracle:
    def __init__(self, opt="O3"):
        path = os.path.join(HERE, "_build", f"liboracle_{opt}.so")
        if not os.path.exists(path):
            build()
        self.lib = L = C.CDLL(path)
        L.oracle_model_index.argtypes = [C.c_char_p]
        L.oracle_np.argtypes = [C.c_int, C.c_int, C.c_int]
        L.oracle_node_eval.argtypes = [C.c_int, C.c_int, C.c_double] + [_dp] * 8
        L.oracle_linearise.argtypes = [C.c_int] + [_dp] * 5
        L.oracle_nlp_eval.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, _dp, _dp, _dp, C.c_double] + [_dp] * 5
        L.oracle_nlp_eval.restype = C.c_double
        L.oracle_solve_batch.argtypes = [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, _dp, _dp, _dp, _dp, _dp, _dp,
                                         C.c_double, C.c_int, C.c_double, C.POINTER(Stats), C.c_int]

    def model(self, name):
        i = self.lib.oracle_model_index(name.encode())
        if i < 0:
            raise KeyError(name)
        return i, self.lib.oracle_model_nx(i), self.lib.oracle_model_nu(i)

    def node_eval(self, name, h, x, u, xnext, w=None):
        m, nx, nu = self.model(name)
        x = np.ascontiguousarray(x, np.float64); u = np.ascontiguousarray(u, np.float64)
        xnext = np.ascontiguousarray(xnext, np.float64)
        n, s = x.shape[0], nx + nu
        g = np.empty((n, nx)); Fx = np.empty((n, nx, nx)); Fu = np.empty((n, nx, nu))
        H = np.empty((n, s, s)) if w is not None else None
        if w is not None:
            w = np.ascontiguousarray(w, np.float64)
        self.lib.oracle_node_eval(m, n, h, _p(x), _p(u), _p(xnext), _p(w), _p(g), _p(Fx), _p(Fu), _p(H))
        return g, Fx, Fu, H

    def linearise(self, name, x, u):
        m, nx, nu = self.model(name)
        x = np.ascontiguousarray(x, np.float64); u = np.ascontiguousarray(u, np.float64)
        f = np.empty(nx); A = np.empty(nx * nx); B = np.empty(nx * nu)
        self.lib.oracle_linearise(m, _p(x), _p(u), _p(f), _p(A), _p(B))
        return f, A, B  # A, B column-major flat (ModelControl.cpp:127-133)

    def nlp_eval(self, name, N, h, is_linear, V, p, lam_g=None, lam_f=1.0, hess=True):
        m, nx, nu = self.model(name)
        s = nx + nu
        V = np.ascontiguousarray(V, np.float64); p = np.ascontiguousarray(p, np.float64)
        lam = np.zeros(N * nx) if lam_g is None else np.ascontiguousarray(lam_g, np.float64)
        g = np.empty(N * nx); gf = np.empty(V.size); Fx = np.empty((N, nx, nx)); Fu = np.empty((N, nx, nu))
        H = np.empty((N, s, s)) if hess else None
        f = self.lib.oracle_nlp_eval(m, N, h, int(is_linear), _p(V), _p(p), _p(lam), lam_f, _p(g), _p(gf), _p(Fx),
                                     _p(Fu), _p(H))
        return dict(f=f, g=g, grad_f=gf, Fx=Fx, Fu=Fu, Hzz=H)

    def solve_batch(self, name, N, h, is_linear, x0, p, V0=None, lam0=None, lb=None, ub=None, tol=1e-5,
                    max_iter=200, mu_init=0.1, nthreads=0):
        m, nx, nu = self.model(name)
        x0 = np.ascontiguousarray(x0, np.float64); p = np.ascontiguousarray(p, np.float64)
        B = x0.shape[0]
        NV = nx * (N + 1) + nu * N
        V = np.zeros((B, NV)) if V0 is None else np.array(V0, np.float64, order="C")
        lam = np.zeros((B, N * nx)) if lam0 is None else np.array(lam0, np.float64, order="C")
        lb = None if lb is None else np.ascontiguousarray(lb, np.float64)
        ub = None if ub is None else np.ascontiguousarray(ub, np.float64)
        st = (Stats * B)()
        bad = self.lib.oracle_solve_batch(m, N, h, int(is_linear), B, _p(x0), _p(p), _p(V), _p(lam), _p(lb), _p(ub),
                                          tol, max_iter, mu_init, st, nthreads)
        stats = np.array([(s_.f, s_.kkt_error, s_.mu, s_.iters, s_.status) for s_ in st],
                         dtype=[("f", "f8"), ("kkt_error", "f8"), ("mu", "f8"), ("iters", "i4"), ("status", "i4")])
        return V, lam, stats, bad

    def max_threads(self):
        return self.lib.oracle_max_threads()


class RefOld:
    """oracle/_ref/libdp_old_ref.so -- f, A, B from the reference's own old/Models closed forms."""

    def __init__(self):
        path = os.path.join(HERE, "_ref", "libdp_old_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.lib.dp_old_ref_eval.argtypes = [C.c_int] + [_dp] * 5

    def eval(self, x, u):
        x = np.ascontiguousarray(x, np.float64); u = np.ascontiguousarray(u, np.float64)
        n = x.shape[0]
        f = np.empty((n, 4)); A = np.empty((n, 4, 4)); B = np.empty((n, 4, 2))
        self.lib.dp_old_ref_eval(n, _p(x), _p(u), _p(f), _p(A), _p(B))
        return f, A, B
