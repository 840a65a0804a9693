"""CPU suite for the product's own ModelGenerator core (csrc/expr + csrc/modelgen): the generated
device functor, compiled for the host, must match the golden vectors (the reference's closed forms and
50-digit mpmath), and the generated <name>_linear_functions.c must carry CasADi's generated-code ABI
(src/codegen_usage.cpp:61-70,118-132)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GEN = os.path.join(ROOT, "mahi_mpc_b200", "csrc", "generated")
dp = C.POINTER(C.c_double)


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(1.0, np.abs(np.asarray(b)).max())


@pytest.fixture(scope="module")
def generated(native_lib):
    from mahi_mpc_b200 import build
    build.generate_builtin()
    return GEN


def test_generated_functor_vs_golden(generated, golden, tmp_path):
    src = tmp_path / "chk.cpp"
    src.write_text('#include "mmpc_model_double_pendulum.cuh"\n'
                   'extern "C" void chk(const double* x,const double* u,const double* wa,const double* wb,double* f,double* A,double* B,double* H){'
                   ' mmpc_model_double_pendulum::fABH(x,u,wa,wb,f,A,B,H);}\n'
                   'extern "C" void chk_f(const double* x,const double* u,double* f){ mmpc_model_double_pendulum::f(x,u,f);}\n'
                   'extern "C" void chk_fab(const double* x,const double* u,double* f,double* A,double* B){ mmpc_model_double_pendulum::fAB(x,u,f,A,B);}\n')
    so = tmp_path / "chk.so"
    subprocess.run(["/usr/bin/g++", "-O1", "-shared", "-fPIC", "-I", generated, str(src), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    g_ = golden["nodes"]
    x, u, w = np.array(g_["x"]), np.array(g_["u"]), np.array(g_["w"])
    iu = np.triu_indices(6)
    for i in range(x.shape[0]):
        f = np.empty(4); A = np.empty(16); B = np.empty(8); H = np.empty(21); wa = np.zeros(4)
        lib.chk(*[a.ctypes.data_as(dp) for a in (x[i], u[i], wa, w[i], f, A, B, H)])
        assert rel(f, g_["mp"]["f"][i]) < 1e-13 and rel(f, g_["ref_old"]["f"][i]) < 1e-13
        assert rel(A.reshape(4, 4), g_["ref_old"]["A"][i]) < 1e-13
        assert rel(B.reshape(4, 2), g_["ref_old"]["B"][i]) < 1e-13
        assert rel(H, np.array(g_["mp"]["Hc"][i])[iu]) < 1e-13
        f2 = np.empty(4); lib.chk_f(x[i].ctypes.data_as(dp), u[i].ctypes.data_as(dp), f2.ctypes.data_as(dp))
        f3 = np.empty(4); A3 = np.empty(16); B3 = np.empty(8)
        lib.chk_fab(*[a.ctypes.data_as(dp) for a in (x[i], u[i], f3, A3, B3)])
        assert np.array_equal(f2, f) and np.array_equal(A3, A)
    # affine weights w = wa*f + wb: H(wa, wb) == H(0, wa*f + wb)
    i = 5
    f = np.empty(4); A = np.empty(16); B = np.empty(8); H1 = np.empty(21); H2 = np.empty(21)
    wa = np.array([0.3, -0.2, 0.7, 1.1]); wb = w[i]
    lib.chk(*[a.ctypes.data_as(dp) for a in (x[i], u[i], wa, wb, f, A, B, H1)])
    lib.chk(*[a.ctypes.data_as(dp) for a in (x[i], u[i], np.zeros(4), wa * f + wb, f, A, B, H2)])
    assert rel(H1, H2) < 1e-13


def test_linear_functions_casadi_abi(generated, golden, tmp_path):
    so = tmp_path / "lin.so"
    # the reference compiles this file with exactly: gcc -fPIC -shared -O3 (ModelGenerator.cpp:247-249)
    subprocess.run(["/usr/bin/gcc", "-fPIC", "-shared", "-O3", os.path.join(generated, "double_pendulum_linear_functions.c"),
                    "-o", str(so), "-lm"], check=True)
    lib = C.CDLL(str(so))
    ll = C.c_longlong
    for fn, shape, key in [("double_pendulum_get_A", (4, 4), "A"), ("double_pendulum_get_B", (4, 2), "B"),
                           ("double_pendulum_get_x_dot_init", (4, 1), "f")]:
        f = getattr(lib, fn)
        n_in = getattr(lib, fn + "_n_in"); n_in.restype = ll
        n_out = getattr(lib, fn + "_n_out"); n_out.restype = ll
        assert n_in() == 2 and n_out() == 1
        sp = getattr(lib, fn + "_sparsity_out"); sp.restype = C.POINTER(ll); sp.argtypes = [ll]
        s = sp(0)
        nrow, ncol = s[0], s[1]
        assert (nrow, ncol) == shape
        colind = [s[2 + j] for j in range(ncol + 1)]
        rows = [s[3 + ncol + k] for k in range(colind[-1])]
        work = getattr(lib, fn + "_work")
        sz = [ll() for _ in range(4)]
        assert work(*[C.byref(v) for v in sz]) == 0 and sz[0].value == 2 and sz[1].value == 1
        g_ = golden["nodes"]
        for i in (0, 1, 17):
            x = np.array(g_["x"][i]); u = np.array(g_["u"][i]); nz = np.zeros(colind[-1])
            arg = (dp * 2)(x.ctypes.data_as(dp), u.ctypes.data_as(dp))
            res = (dp * 1)(nz.ctypes.data_as(dp))
            assert f(arg, res, None, None, 0) == 0
            dense = np.zeros(shape)
            for j in range(ncol):
                for k in range(colind[j], colind[j + 1]):
                    dense[rows[k], j] = nz[k]
            want = np.array(g_["ref_old"][key][i]).reshape(shape)
            assert rel(dense, want) < 1e-13
    # structural pattern CasADi would report for the pendulum: A has 10 nnz, B has 4
    spA = lib.double_pendulum_get_A_sparsity_out; spA.restype = C.POINTER(ll); spA.argtypes = [ll]
    assert spA(0)[2 + 4] == 10


def test_symbolic_engine_unit(native_lib, tmp_path):
    """Derivative rules of csrc/expr/sx.cpp against finite differences on a synthetic expression."""
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include "expr/sx.hpp"
#include <cstdio>
#include <cmath>
using namespace mahi::mpc::sym;
int main() {
    SX x = SX::sym("x", 3);
    SX e = SX::vertcat({ sin(x(0)) * exp(x(1)) / (1.0 + x(2) * x(2)) + sqrt(x(0) * x(0) + 2.0) - tanh(x(1)) * log(3.0 + x(2)),
                         pow(x(0) * x(0) + 1.0, SX(1.5)) * cos(x(1) - x(2)) + atan(x(0) * x(2)) + tan(0.3 * x(1)) });
    SX J = jacobian(e, x);
    SX H = hessian(e(0) + 2.0 * e(1), x);
    std::vector<double> v = {0.4, -0.7, 1.3};
    auto f = [&](std::vector<double> p) { return evaluate(e, x, p); };
    auto Jv = evaluate(J, x, v);
    double worst = 0;
    for (int j = 0; j < 3; ++j) {
        auto p = v, m = v; p[j] += 1e-6; m[j] -= 1e-6;
        auto fp = f(p), fm = f(m);
        for (int i = 0; i < 2; ++i) worst = std::fmax(worst, std::fabs((fp[i] - fm[i]) / 2e-6 - Jv[j * 2 + i]));
    }
    SX gsym = gradient(e(0) + 2.0 * e(1), x);
    auto Hv = evaluate(H, x, v);
    for (int j = 0; j < 3; ++j) {
        auto p = v, m = v; p[j] += 1e-6; m[j] -= 1e-6;
        auto gp = evaluate(gsym, x, p), gm = evaluate(gsym, x, m);
        for (int i = 0; i < 3; ++i) worst = std::fmax(worst, std::fabs((gp[i] - gm[i]) / 2e-6 - Hv[j * 3 + i]));
    }
    // hash-consing: identical sub-expressions are one node; simplifications fold
    SX a = x(0) * x(1) + 0.0, b = x(1) * x(0) * 1.0;
    int shared = (a.ids()[0] == b.ids()[0]);
    SX z = x(0) - x(0);
    std::printf("%.3e %d %d\n", worst, shared, (int)graph().is_zero(z.ids()[0]));
    return 0;
}''')
    exe = tmp_path / "t"
    csrc = os.path.join(ROOT, "mahi_mpc_b200", "csrc")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-I", csrc, str(src), os.path.join(csrc, "expr", "sx.cpp"), "-o", str(exe)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert float(out[0]) < 1e-7 and out[1] == "1" and out[2] == "1"
