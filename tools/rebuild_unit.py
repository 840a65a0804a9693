"""Developer tool: recompile ONE generated unit (e.g. builtin_double_pendulum_t3) with extra defines and relink the
C-ABI library from the existing objects -- seconds instead of the full build.
    python -m tools.rebuild_unit builtin_double_pendulum_t3 [-DMMPC_T3_PROFILE ...]
Run `python -m mahi_mpc_b200.build --force` afterwards to get back to a clean build."""
import os, sys
from mahi_mpc_b200 import build as B

unit, extra = sys.argv[1], sys.argv[2:]
objdir = os.path.join(B.BUILD, "obj")
src = os.path.join(B.GEN, unit + ".cu")
defs = [f"-DMMPC_HAVE_{m.upper()}" for m in B.BUILTIN_MODELS if m != "double_pendulum"] + extra
_, log = B._compile_unit((src, os.path.join(objdir, unit + ".o"), defs))
print("\n".join(l for l in log.splitlines() if "error" in l or "stack frame" in l or "Used" in l))
objs = [os.path.join(objdir, f) for f in sorted(os.listdir(objdir)) if f.endswith(".o")]
so = os.path.join(B.LIB, "libmahi_mpc_b200.so")
B._run([B._nvcc()] + B.NVCC_ARCH + ["-shared", "-cudart", "shared", "-ccbin", B._cxx()] + objs + ["-o", so, "-ldl"])
print(so)
