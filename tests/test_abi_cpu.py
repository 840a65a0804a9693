"""CPU suite for the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, handles model artefacts, and FAILS LOUDLY (no fallback) when there is no CUDA device."""
import ctypes as C
import json
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api(native_lib):
    from mahi_mpc_b200 import api as a
    a.lib()
    return a


def _has_gpu(api):
    try:
        return api.device_count() > 0
    except api.MmpcError:
        return False


def test_header_symbols_all_exported(api):
    hdr = open(os.path.join(ROOT, "include", "mahi_mpc_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mmpc_[a-z_0-9]+)\s*\(", hdr)))
    assert len(declared) >= 20
    lib = C.CDLL(api.LIB_PATH)
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(api.ABI_SYMBOLS) == declared
    assert lib.mmpc_abi_version() == 1
    # ... and nothing else of the library's own: the per-model launchers / vtable getters of the kernel units are internal
    import shutil, subprocess
    if shutil.which("nm"):
        out = subprocess.run(["nm", "-D", "--defined-only", api.LIB_PATH], capture_output=True, text=True).stdout
        exported = sorted({ln.split()[-1] for ln in out.splitlines() if " T " in ln and ln.split()[-1].startswith("mmpc_")})
        assert exported == declared, sorted(set(exported) ^ set(declared))


def test_options_default_are_the_reference_ipopt_options(api):
    o = api.default_options()  # ModelControl.cpp:54-55
    assert o.tol == 1e-5 and o.max_iter == 200 and o.integrator == api.EULER and o.warm_start == 1 and o.warm_start_duals == 0


def test_model_dims_and_errors(api):
    for N, NV, M, np_ in [(25, 154, 100, 110), (50, 304, 200, 210), (200, 1204, 800, 810)]:
        m = api.Model("double_pendulum", N=N)
        assert (m.nx, m.nu, m.N, m.NV, m.M, m.np) == (4, 2, N, NV, M, np_)
        assert m.h == 0.002 and not m.is_linear
    m = api.Model("linear_double_pendulum", N=25, is_linear=True)
    assert m.np == 110 + 32 and m.is_linear
    assert api.Model("nonlinear_double_pendulum", N=25).nx == 4  # ex_model_generate.cpp:44-45 naming
    with pytest.raises(api.MmpcError) as e:
        api.Model("mystery_plant", N=5, nx=3, nu=1)
    assert e.value.code == -5
    with pytest.raises(api.MmpcError):
        api.Model("double_pendulum", N=0)
    with pytest.raises(api.MmpcError):
        api.Model("double_pendulum", N=5, nx=5, nu=2)  # dims disagree with the kernel set
    c = api.Model("double_pendulum", N=5).op_counts()
    assert c["f"][0] > 0 and c["fABH"][0] > c["fAB"][0] > c["f"][0] and c["fABH"][1] == 6


def test_model_load_json_reference_schema(api, tmp_path):
    """<name>.json exactly as ModelParameters.cpp:37-50 / ModelGenerator.cpp:261-270 write it."""
    j = {"model": {"name": "nonlinear_double_pendulum", "timespan": 50000, "step_size": 2000, "num_x": 4, "num_u": 2,
                   "num_shooting_nodes": 25, "x_min": [-1e31] * 4, "u_min": [-1e31] * 2, "x_max": [1e31] * 4,
                   "u_max": [1e31] * 2, "dll_filepath": "nonlinear_double_pendulum.so", "is_linear": False}}
    path = tmp_path / "nonlinear_double_pendulum.json"
    path.write_text(json.dumps(j))
    # the library the artefact names is absent (or is the reference's CasADi CPU library): an error, never a silent switch to
    # the built-in L = m = 1 dynamics
    with pytest.raises(api.MmpcError) as e:
        api.Model(json_name=str(tmp_path / "nonlinear_double_pendulum"))
    assert e.value.code == -5 and "nonlinear_double_pendulum.so" in str(e.value)
    (tmp_path / "cpu_only.c").write_text("int nlp_f(void) { return 0; }\n")
    import subprocess
    subprocess.run(["gcc", "-shared", "-fPIC", str(tmp_path / "cpu_only.c"), "-o", str(tmp_path / "cpu_only.so")], check=True)
    j["model"]["dll_filepath"] = str(tmp_path / "cpu_only.so")
    path.write_text(json.dumps(j))
    with pytest.raises(api.MmpcError) as e:
        api.Model(json_name=str(tmp_path / "nonlinear_double_pendulum"))
    assert e.value.code == -5 and "not a mahi-mpc-b200 CUDA model library" in str(e.value)
    j["model"]["dll_filepath"] = "builtin:double_pendulum"  # explicit request for a built-in kernel set
    path.write_text(json.dumps(j))
    m = api.Model(json_name=str(tmp_path / "nonlinear_double_pendulum"))
    assert (m.nx, m.nu, m.N, m.h) == (4, 2, 25, 0.002)
    with pytest.raises(api.MmpcError) as e:
        api.Model(json_name=str(tmp_path / "absent"))
    assert e.value.code == -4
    (tmp_path / "broken.json").write_text('{"model": {"name": "x"')
    with pytest.raises(api.MmpcError):
        api.Model(json_name=str(tmp_path / "broken"))
    del j["model"]["num_x"]
    (tmp_path / "partial.json").write_text(json.dumps(j))
    with pytest.raises(api.MmpcError):
        api.Model(json_name=str(tmp_path / "partial"))


def test_no_cpu_fallback(api):
    """Without a CUDA device every compute entry point must refuse (MMPC_ERR_NO_DEVICE = -2)."""
    if _has_gpu(api):
        pytest.skip("a GPU is present")
    m = api.Model("double_pendulum", N=5)
    with pytest.raises(api.MmpcError) as e:
        api.Solver(m, 4)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)
    z = np.zeros((4, 3))
    with pytest.raises(api.MmpcError) as e:
        m.eval_nodes(z, np.zeros((2, 3)), z)
    assert e.value.code == -2
    with pytest.raises(api.MmpcError) as e:
        m.nlp_eval(np.zeros((1, m.NV)), np.zeros((1, m.np)))
    assert e.value.code == -2
    with pytest.raises(api.MmpcError):
        m.linearise(np.zeros((1, 4)), np.zeros((1, 2)))


def test_product_never_imports_oracle():
    """The product path must not route through oracle/ (only tests/, smoke() and bench.py's baseline legs may)."""
    bad = []
    for d, _, files in os.walk(os.path.join(ROOT, "mahi_mpc_b200")):
        for f in files:
            if f.endswith((".py", ".cpp", ".hpp", ".cu", ".cuh", ".h")):
                text = open(os.path.join(d, f), errors="replace").read()
                if re.search(r"^\s*(from|import)\s+oracle\b|#include\s+[\"<].*oracle", text, flags=re.M):
                    bad.append(os.path.join(d, f))
    for d, _, files in os.walk(os.path.join(ROOT, "include")):
        for f in files:
            if "oracle" in open(os.path.join(d, f), errors="replace").read():
                bad.append(os.path.join(d, f))
    assert not bad, bad
