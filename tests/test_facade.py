"""The drop-in C++ layer: mahi::mpc::ModelParameters / ModelGenerator / ModelControl with the reference's class
and method names (include/Mahi/Mpc/*.hpp) over the C ABI.  CPU part: the generate driver runs end to end here
(nvcc cross-compiles the generated CUDA model); GPU part: the closed-loop control driver (BASELINE config 1)
against the oracle's closed loop."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EX = os.path.join(ROOT, "build", "examples")


@pytest.fixture(scope="module")
def examples(native_lib):
    from mahi_mpc_b200 import build
    build.build_facade()
    exes = build.build_examples()
    assert {os.path.basename(e) for e in exes} >= {"ex_model_generate", "ex_model_control"}
    return EX


@pytest.fixture(scope="module")
def generated_model(examples, tmp_path_factory):
    d = tmp_path_factory.mktemp("model")
    env = dict(os.environ, MAHI_MPC_B200_HOME=ROOT)
    for flag in ([], ["-l"]):
        r = subprocess.run([os.path.join(examples, "ex_model_generate")] + flag, cwd=d, env=env, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return d


def test_generate_driver_writes_reference_artefacts(generated_model):
    d = generated_model
    for name, lin in (("nonlinear_double_pendulum", False), ("linear_double_pendulum", True)):
        for ext in (".json", ".cu", ".so", "_linear_functions.c", "_linear_functions.so"):
            assert (d / (name + ext)).exists(), name + ext
        j = json.loads((d / (name + ".json")).read_text())["model"]
        # schema of ModelParameters.cpp:37-50
        assert sorted(j) == sorted(["name", "timespan", "step_size", "num_x", "num_u", "num_shooting_nodes", "x_min", "u_min",
                                    "x_max", "u_max", "dll_filepath", "is_linear"])
        assert (j["name"], j["num_x"], j["num_u"], j["num_shooting_nodes"], j["step_size"], j["timespan"]) == (name, 4, 2, 25, 2000, 50000)
        assert j["is_linear"] is lin and j["dll_filepath"] == name + ".so" and j["x_min"] == [-1e31] * 4 and j["u_max"] == [1e31] * 2


def test_generated_model_library_loads_through_the_abi(generated_model, native_lib):
    from mahi_mpc_b200 import api
    cwd = os.getcwd()
    os.chdir(generated_model)
    try:
        m = api.Model(json_name="nonlinear_double_pendulum")
        assert (m.nx, m.nu, m.N, m.NV, m.np, m.h, m.is_linear) == (4, 2, 25, 154, 110, 0.002, False)
        ml = api.Model(json_name="linear_double_pendulum")
        assert ml.is_linear and ml.np == 142
    finally:
        os.chdir(cwd)


def _oracle_closed_loop(cycles, tol):
    """numpy restatement of the loop of examples/model_generate_example.cpp:156-230 (solve every cycle, apply u_0,
    RK4 plant), with the primal warm start of ModelControl.cpp:161."""
    from oracle import models, nlp
    P = nlp.Problem(N=25)
    m = models.numeric("double_pendulum")
    Q, R, Rm = np.array([10.0, 1, 5, 5]), np.array([5.0, 5]), np.array([1.0, 1])
    f = lambda x, u: m["f"](x[:, None], u[:, None])[:, 0]
    x, u, t, V0 = np.zeros(4), np.zeros(2), 0.0, None
    rows = []
    for _ in range(cycles):
        tt = t + np.arange(25) * P.h
        s, c = np.sin(2 * np.pi * tt), 2 * np.pi * np.cos(2 * np.pi * tt)
        p = P.pack_p(np.stack([s, -s, c, -c], 1), Q, R, Rm, u)
        r = nlp.solve(P, x, p, V0=V0, tol=tol)
        assert r.status == 0
        V0, u = r.V, r.V[4:6].copy()
        rows.append(np.r_[t, x, u])
        k1 = f(x, u); k2 = f(x + P.h / 2 * k1, u); k3 = f(x + P.h / 2 * k2, u); k4 = f(x + P.h * k3, u)
        x = x + P.h / 6 * (k1 + 2 * k2 + 2 * k3 + k4)
        t += P.h
    return np.array(rows)


@pytest.mark.gpu
def test_control_driver_closed_loop_vs_oracle(generated_model, examples):
    """BASELINE config 1: one closed-loop run through the C++ ModelControl facade on the GPU vs the oracle."""
    cycles = 40
    env = dict(os.environ, MAHI_MPC_B200_HOME=ROOT)
    r = subprocess.run([os.path.join(examples, "ex_model_control"), "-t", str(cycles * 0.002), "-e", "1e-10"], cwd=generated_model,
                       env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "avg solve time" in r.stdout
    got = np.loadtxt(generated_model / "nonlinear_double_pendulum_results.csv", delimiter=",", skiprows=1)
    want = _oracle_closed_loop(cycles, 1e-11)
    assert got.shape == (cycles, 8)
    assert np.abs(got[:, :5] - want[:, :5]).max() < 1e-8       # time and state trajectory
    assert np.abs(got[:, 5:7] - want[:, 5:7]).max() < 1e-6     # applied controls (north_star: 1e-6 on controls)
    assert (got[1:, 7] <= got[0, 7]).all()                     # warm starts never need more iterations than the cold solve
    # the reference's default tolerance keeps the loop inside the same ball
    r = subprocess.run([os.path.join(examples, "ex_model_control"), "-t", str(cycles * 0.002)], cwd=generated_model, env=env,
                       capture_output=True, text=True)
    assert r.returncode == 0
    got5 = np.loadtxt(generated_model / "nonlinear_double_pendulum_results.csv", delimiter=",", skiprows=1)
    assert np.abs(got5[:, 5:7] - want[:, 5:7]).max() < 1e-4
    # linearised model (-l): runs and tracks; no oracle value asserted beyond finiteness and boundedness
    r = subprocess.run([os.path.join(examples, "ex_model_control"), "-l", "-t", "0.04"], cwd=generated_model, env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    gl = np.loadtxt(generated_model / "linear_double_pendulum_results.csv", delimiter=",", skiprows=1)
    assert np.isfinite(gl).all() and np.abs(gl[:, 5:7]).max() < 1e3


@pytest.mark.gpu
def test_facade_threaded_and_batched(generated_model, examples, tmp_path):
    """ModelControl's solver thread, control_at_time, update_weights, update_control_limits, BatchModelControl (C++)."""
    exe = tmp_path / "facade_threaded"
    lib = os.path.join(ROOT, "mahi_mpc_b200", "lib")
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-pthread", "-I" + os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "cpp", "facade_threaded.cpp"), "-o", str(exe), "-L" + lib, "-lmahi_mpc_facade",
                    "-lmahi_mpc_b200", "-Wl,-rpath," + lib, "-ldl"], check=True)
    r = subprocess.run([str(exe)], cwd=generated_model, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "FACADE-THREADED-OK" in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_prototype_closed_loop_example(examples):
    """examples/ex_prototype_mpc.cpp: the MATLAB prototype's closed loop (workingMPC_8_30_21.m: RK4 x 4, exact Hessian, objective on
    the decision variables with weight 1000) through the C ABI -- every cycle converges and the loop tracks the reference."""
    import re
    r = subprocess.run([os.path.join(examples, "ex_prototype_mpc"), "40"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "status" not in r.stderr, r.stdout[-2000:] + r.stderr[-2000:]
    m = re.search(r"mean Newton iterations ([0-9.]+), max \|qA - reference\| after the first second ([0-9.e+-]+) rad", r.stdout)
    assert m, r.stdout[-1000:]
    assert float(m.group(1)) < 12 and float(m.group(2)) < 0.5, r.stdout[-1000:]  # 1000 e^2 against u^2 with |u| ~ 20 Nm: ~0.3 rad
