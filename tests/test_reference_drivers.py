"""The reference's OWN example drivers, unmodified, against this repo's include/ tree and libraries.

CPU (needs /root/reference, i.e. the build container): examples/ex_model_generate.cpp, model_control_example.cpp and
thread_model_control_example.cpp compile and link as they are; the generate driver runs and leaves the artefact triple.
The binaries are kept under build/refex/ (git-ignored, shipped to the GPU box), so the GPU test can run the reference's
closed-loop driver there, where /root/reference does not exist, and compare its result file with the oracle's closed loop.
"""
import json
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_EX = "/root/reference/examples"
OUT = os.path.join(ROOT, "build", "refex")
LIB = os.path.join(ROOT, "mahi_mpc_b200", "lib")
DRIVERS = ["ex_model_generate", "model_control_example", "thread_model_control_example"]


def _compile(name):
    os.makedirs(OUT, exist_ok=True)
    exe = os.path.join(OUT, name)
    cmd = ["/usr/bin/g++", "-std=c++17", "-O1", "-pthread", "-I" + os.path.join(ROOT, "include"), os.path.join(REF_EX, name + ".cpp"),
           "-o", exe, "-L" + LIB, "-lmahi_mpc_facade", "-lmahi_mpc_b200", "-Wl,-rpath," + LIB, "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, f"{name}.cpp does not compile unmodified against include/:\n" + r.stderr[-4000:]
    return exe


@pytest.fixture(scope="module")
def ref_drivers(native_lib):
    if not os.path.isdir(REF_EX):
        pytest.skip("/root/reference is not present on this machine")
    from mahi_mpc_b200 import build
    build.build_facade()
    return {n: _compile(n) for n in DRIVERS}


def test_reference_examples_compile_unmodified(ref_drivers):
    for n in DRIVERS:
        assert os.access(ref_drivers[n], os.X_OK)


def test_reference_generate_driver_runs_unmodified(ref_drivers, tmp_path):
    """examples/ex_model_generate.cpp:8-60 as shipped: ModelParameters -> ModelGenerator::create_model / generate_c_code /
    compile_model; here the model library is CUDA for sm_100a (nvcc cross-compiles without a GPU)."""
    env = dict(os.environ, MAHI_MPC_B200_HOME=ROOT)
    r = subprocess.run([ref_drivers["ex_model_generate"]], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    name = "nonlinear_double_pendulum"  # ex_model_generate.cpp:44-45
    for ext in (".json", ".so", "_linear_functions.c", "_linear_functions.so"):
        assert (tmp_path / (name + ext)).exists(), name + ext
    j = json.loads((tmp_path / (name + ".json")).read_text())["model"]
    assert (j["num_x"], j["num_u"], j["num_shooting_nodes"], j["step_size"], j["is_linear"]) == (4, 2, 25, 2000, False)
    from mahi_mpc_b200 import api
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        m = api.Model(json_name=name)  # the generated library exports the kernel set through the C ABI
        assert (m.nx, m.nu, m.N, m.NV) == (4, 2, 25, 154)
    finally:
        os.chdir(cwd)


def _parse_m_file(path):
    out = {}
    for line in open(path):
        m = re.match(r"^(\w+) = \[(.*)\];", line)
        if m:
            out[m.group(1)] = np.array([float(v) for v in m.group(2).split(",") if v.strip()])
    return out


def _oracle_driver_loop(coracle, cycles, amp, freq, Q, R, Rm, every):
    """The loop of examples/model_control_example.cpp:54-93 restated on the oracle: reference over the horizon, solve every
    `every`-th cycle (primal warm start, ModelControl.cpp:161), zero-order hold of the plan (ModelControl.cpp:192-197:
    last entry with time < t), explicit-Euler plant from <name>_get_x_dot_init."""
    from oracle import models, nlp
    N, h_us = 25, 2000
    h = h_us * 1e-6
    P = nlp.Problem(N=N)
    mdl = models.numeric("double_pendulum")
    f = lambda x, u: mdl["f"](x[:, None], u[:, None])[:, 0]
    x, u, t, V0 = np.zeros(4), np.zeros(2), 0.0, None
    plan_t0_us, plan_u = 0, None
    X, U = [], []
    for c in range(cycles):
        tt = t + h * np.arange(N)  # curr_traj_time accumulates by repeated addition in the driver; same to ~1e-16
        tt = np.cumsum(np.r_[t, np.full(N - 1, h)])
        s, co = amp * np.sin(2 * np.pi * freq * tt), amp * 2 * np.pi * freq * np.cos(2 * np.pi * freq * tt)
        traj = np.stack([s, -s, co, -co], 1)
        X.append(x.copy()); U.append(u.copy())
        t_us = int(round(t * 1e6))
        if c % every == 0:
            p = P.pack_p(traj, Q, R, Rm, u)
            V, _, st, bad = coracle.solve_batch("double_pendulum", N, h, False, x[None], p[None], V0=V0, tol=1e-5)
            assert bad == 0
            V0 = V
            plan_t0_us, plan_u = t_us, V[0, :N * 6].reshape(N, 6)[:, 4:].copy()
        times = plan_t0_us + h_us * np.arange(N)  # seconds(curr_time + h i) rounded to microseconds
        i = int((times < t_us).sum())
        u = plan_u[0 if i == 0 else i - 1].copy()
        x = x + f(x, u) * h
        t += h
    return np.array(X), np.array(U)


@pytest.mark.gpu
def test_reference_control_driver_runs_unmodified_vs_oracle(coracle, native_lib, tmp_path):
    """examples/model_control_example.cpp as shipped (built here against include/), run on the GPU box: ModelControl over
    mmpc_solve with B = 1, 100 solves in a 1 s closed loop; its result file against the oracle's restatement of the loop."""
    exe = os.path.join(OUT, "model_control_example")
    gen = os.path.join(ROOT, "build", "examples", "ex_model_generate")
    if not (os.path.exists(exe) and os.path.exists(gen)):
        pytest.skip("build/refex/model_control_example was not built (needs /root/reference at build time)")
    env = dict(os.environ, MAHI_MPC_B200_HOME=ROOT)
    # the driver loads "double_pendulum" (model_control_example.cpp:17), the name its generator does not write
    r = subprocess.run([gen, "--name", "double_pendulum"], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    r = subprocess.run([exe], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "avg solve time" in r.stdout
    got = _parse_m_file(tmp_path / "double_pendulum_results.m")
    cycles = got["t"].size
    assert cycles in (500, 501)  # 1 s of 2 ms cycles (accumulated floating-point time)
    # the driver passes no weights: Q = R = Rm = identity (ModelGenerator.cpp:122-124)
    X, U = _oracle_driver_loop(coracle, cycles, 0.5, 0.25, np.ones(4), np.ones(2), np.ones(2), 5)
    Xg = np.stack([got[f"q{i}"] for i in range(4)], 1)
    Ug = np.stack([got[f"T{i}"] for i in range(2)], 1)
    # the result file carries 6 significant digits (default ostream precision)
    assert np.abs(Ug - U).max() < 2e-5 * max(1.0, np.abs(U).max()), np.abs(Ug - U).max()
    assert np.abs(Xg - X).max() < 2e-5, np.abs(Xg - X).max()
    assert np.abs(U).max() > 0.1  # a real control signal, not zeros
    shutil.rmtree(tmp_path, ignore_errors=True)
