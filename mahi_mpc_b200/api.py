"""ctypes binding of the C ABI (include/mahi_mpc_b200.h) for the Python test/bench harness.

This is harness plumbing, not the product: every call goes straight into libmahi_mpc_b200.so
(hand-written sm_100a kernels).  There is no Python or CPU implementation behind these functions;
if the library has not been built the import of the native handle raises.

numpy arrays are passed as HOST pointers (MMPC_MEM_HOST), torch CUDA tensors as DEVICE pointers
(MMPC_MEM_DEVICE) on the tensor's device and the current torch stream.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
# MMPC_LIB_PATH: developer aid for A/B timing of two builds of the library on one GPU box (tools/gpu_ab_lib.sh)
LIB_PATH = os.environ.get("MMPC_LIB_PATH") or os.path.join(PKG, "lib", "libmahi_mpc_b200.so")

MEM_HOST, MEM_DEVICE = 0, 1
EULER, RK4 = 0, 1
HESSIAN_EXACT, HESSIAN_GAUSS_NEWTON = 0, 1
COST_SHOOTING, COST_NODE = 0, 1
_dp = C.POINTER(C.c_double)


class MmpcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mmpc error {code}: {msg}")
        self.code = code


class ModelParams(C.Structure):
    _fields_ = [("name", C.c_char_p), ("timespan_us", C.c_longlong), ("step_size_us", C.c_longlong), ("num_x", C.c_int),
                ("num_u", C.c_int), ("num_shooting_nodes", C.c_int), ("x_min", _dp), ("u_min", _dp), ("x_max", _dp),
                ("u_max", _dp), ("dll_filepath", C.c_char_p), ("is_linear", C.c_int)]


class Options(C.Structure):
    _fields_ = [("tol", C.c_double), ("max_iter", C.c_int), ("integrator", C.c_int), ("substeps", C.c_int),
                ("mu_init", C.c_double), ("warm_start", C.c_int), ("warm_start_duals", C.c_int),
                ("kernel_variant", C.c_int), ("hessian", C.c_int), ("cost", C.c_int), ("reserved", C.c_int * 6)]


class Stats(C.Structure):
    _fields_ = [("objective", C.c_double), ("kkt_error", C.c_double), ("mu", C.c_double), ("iterations", C.c_int),
                ("status", C.c_int), ("ls_steps", C.c_int), ("reg_steps", C.c_int)]


STATS_DTYPE = np.dtype([("objective", "f8"), ("kkt_error", "f8"), ("mu", "f8"), ("iterations", "i4"), ("status", "i4"),
                        ("ls_steps", "i4"), ("reg_steps", "i4")])
assert STATS_DTYPE.itemsize == C.sizeof(Stats) == 40

# every symbol include/mahi_mpc_b200.h declares (tests check the library exports them all)
ABI_SYMBOLS = [
    "mmpc_last_error", "mmpc_abi_version", "mmpc_device_count", "mmpc_options_default", "mmpc_model_create",
    "mmpc_model_load", "mmpc_model_free", "mmpc_model_dims", "mmpc_model_step_size", "mmpc_model_is_linear",
    "mmpc_model_op_counts", "mmpc_linearise", "mmpc_eval_nodes", "mmpc_nlp_eval", "mmpc_solver_create",
    "mmpc_solver_free", "mmpc_solver_batch", "mmpc_set_options", "mmpc_set_weights", "mmpc_set_bounds",
    "mmpc_reset_warm_start", "mmpc_solve", "mmpc_last_timing", "mmpc_last_launch_count", "mmpc_measure_fp64_peak",
    "mmpc_closed_loop", "mmpc_group_create", "mmpc_group_free", "mmpc_group_shard", "mmpc_group_size", "mmpc_group_set_weights",
    "mmpc_group_set_bounds", "mmpc_group_solve", "mmpc_group_gather",
]

_lib = None


def lib():
    """The native library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -m mahi_mpc_b200.build` (there is no CPU fallback)")
        L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
        L.mmpc_last_error.restype = C.c_char_p
        vp = C.c_void_p
        L.mmpc_model_create.argtypes = [C.POINTER(ModelParams), C.POINTER(vp)]
        L.mmpc_group_create.argtypes = [vp, C.POINTER(C.c_int), C.c_int, C.c_int, C.POINTER(Options), C.POINTER(vp)]
        L.mmpc_group_free.argtypes = [vp]
        L.mmpc_group_size.argtypes = [vp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.mmpc_group_shard.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(vp)]
        L.mmpc_group_set_weights.argtypes = [vp, _dp, _dp, _dp]
        L.mmpc_group_set_bounds.argtypes = [vp, _dp, _dp, _dp, _dp]
        L.mmpc_group_solve.argtypes = [vp, _dp, _dp, _dp, _dp, _dp, C.POINTER(Stats)]
        L.mmpc_group_gather.argtypes = [vp, C.c_int, vp, vp, C.POINTER(C.c_double)]
        L.mmpc_model_load.argtypes = [C.c_char_p, C.POINTER(vp)]
        L.mmpc_model_free.argtypes = [vp]
        L.mmpc_model_dims.argtypes = [vp] + [C.POINTER(C.c_int)] * 6
        L.mmpc_model_step_size.argtypes = [vp, _dp]
        L.mmpc_model_is_linear.argtypes = [vp, C.POINTER(C.c_int)]
        L.mmpc_model_op_counts.argtypes = [vp, C.POINTER(C.c_int * 6)]
        L.mmpc_options_default.argtypes = [C.POINTER(Options)]
        L.mmpc_linearise.argtypes = [vp, C.c_int, C.c_int] + [vp] * 5 + [C.c_int, vp]
        L.mmpc_eval_nodes.argtypes = [vp, C.c_int, C.POINTER(Options), C.c_longlong] + [vp] * 8 + [C.c_int, vp]
        L.mmpc_nlp_eval.argtypes = [vp, C.c_int, C.POINTER(Options), C.c_int, vp, vp, vp, C.c_double] + [vp] * 6 + [C.c_int, vp]
        L.mmpc_solver_create.argtypes = [vp, C.c_int, C.c_int, C.POINTER(Options), C.POINTER(vp)]
        L.mmpc_solver_free.argtypes = [vp]
        L.mmpc_solver_batch.argtypes = [vp, C.POINTER(C.c_int)]
        L.mmpc_set_options.argtypes = [vp, C.POINTER(Options)]
        L.mmpc_set_weights.argtypes = [vp, vp, vp, vp, C.c_int]
        L.mmpc_set_bounds.argtypes = [vp, vp, vp, vp, vp]
        L.mmpc_reset_warm_start.argtypes = [vp]
        L.mmpc_solve.argtypes = [vp] + [vp] * 6 + [C.c_int, vp]
        L.mmpc_last_timing.argtypes = [vp, C.POINTER(C.c_double * 3)]
        L.mmpc_last_launch_count.argtypes = [vp, C.POINTER(C.c_int)]
        L.mmpc_device_count.argtypes = [C.POINTER(C.c_int)]
        L.mmpc_measure_fp64_peak.argtypes = [C.c_int, _dp]
        L.mmpc_closed_loop.argtypes = [vp, C.c_int, C.c_int] + [vp] * 6 + [C.c_int, vp]
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise MmpcError(rc, lib().mmpc_last_error().decode(errors="replace"))


def default_options(**kw) -> Options:
    o = Options()
    _ck(lib().mmpc_options_default(C.byref(o)))
    for k, v in kw.items():
        if not hasattr(o, k):
            raise AttributeError(k)
        setattr(o, k, v)
    return o


def _is_torch(a):
    return a is not None and type(a).__module__.startswith("torch")


def _ptr(a):
    """(pointer, mem kind) of a numpy array / torch tensor / None."""
    if a is None:
        return None, None
    if _is_torch(a):
        assert a.is_contiguous() and a.dtype.is_floating_point and a.element_size() == 8, "need contiguous float64"
        return C.c_void_p(a.data_ptr()), (MEM_DEVICE if a.is_cuda else MEM_HOST)
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"] and a.dtype in (np.float64, STATS_DTYPE), \
        "need C-contiguous float64 numpy array"
    return C.c_void_p(a.ctypes.data), MEM_HOST


def _mem_of(*arrs):
    kinds = {k for k in (_ptr(a)[1] for a in arrs) if k is not None}
    if len(kinds) != 1:
        raise ValueError("all buffers of one call must live in the same memory space")
    return kinds.pop()


def _stream_of(*arrs):
    for a in arrs:
        if _is_torch(a) and a.is_cuda:
            import torch
            # torch's default stream is the legacy NULL stream; the ABI reads NULL as "the handle's own
            # stream", so name it explicitly (cudaStreamLegacy == (cudaStream_t)0x1)
            return C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream or 1), a.device.index or 0
    return None, None


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


class Model:
    """mmpc_model_t: a model artefact (ModelParameters + kernel set)."""

    def __init__(self, name="double_pendulum", N=25, step_size_us=2000, is_linear=False, nx=None, nu=None, x_min=None,
                 x_max=None, u_min=None, u_max=None, dll_filepath=None, json_name=None):
        self._h = C.c_void_p()
        if json_name is not None:
            _ck(lib().mmpc_model_load(json_name.encode(), C.byref(self._h)))
        else:
            builtin_dims = {"double_pendulum": (4, 2), "arm4": (8, 4)}
            key = next((k for k in builtin_dims if name.endswith(k)), None)
            if nx is None or nu is None:
                if key is None:
                    raise ValueError("nx/nu required for non-built-in models")
                nx, nu = builtin_dims[key]
            keep = [_f64(v) for v in (x_min, u_min, x_max, u_max)]
            p = ModelParams(name.encode(), 0, int(step_size_us), nx, nu, int(N),
                            *[None if v is None else v.ctypes.data_as(_dp) for v in keep],
                            (dll_filepath or (f"builtin:{key}" if key else "")).encode(), int(bool(is_linear)))
            _ck(lib().mmpc_model_create(C.byref(p), C.byref(self._h)))
        d = [C.c_int() for _ in range(6)]
        _ck(lib().mmpc_model_dims(self._h, *[C.byref(v) for v in d]))
        self.nx, self.nu, self.N, self.NV, self.M, self.np = [v.value for v in d]
        self.s = self.nx + self.nu
        self.nh = self.s * (self.s + 1) // 2
        h = C.c_double()
        _ck(lib().mmpc_model_step_size(self._h, C.byref(h)))
        self.h = h.value
        lin = C.c_int()
        _ck(lib().mmpc_model_is_linear(self._h, C.byref(lin)))
        self.is_linear = bool(lin.value)

    def op_counts(self):
        c = (C.c_int * 6)()
        _ck(lib().mmpc_model_op_counts(self._h, C.byref(c)))
        return dict(f=(c[0], c[1]), fAB=(c[2], c[3]), fABH=(c[4], c[5]))

    def close(self):
        if self._h:
            lib().mmpc_model_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- stateless evaluators -------------------------------------------------------------
    def linearise(self, x, u, device=0):
        """AoS x (n,nx), u (n,nu) -> xdot (n,nx), A (n,nx*nx) col-major, B (n,nx*nu) col-major."""
        n = x.shape[0]
        if _is_torch(x):
            import torch
            mk = lambda *s: torch.empty(*s, dtype=torch.float64, device=x.device)
        else:
            x, u = _f64(x), _f64(u)
            mk = lambda *s: np.empty(s)
        xd, A, B = mk(n, self.nx), mk(n, self.nx * self.nx), mk(n, self.nx * self.nu)
        mem = _mem_of(x, u, xd)
        st, dev = _stream_of(x)
        _ck(lib().mmpc_linearise(self._h, dev if dev is not None else device, n, _ptr(x)[0], _ptr(u)[0], _ptr(xd)[0],
                                 _ptr(A)[0], _ptr(B)[0], mem, st))
        return xd, A, B

    def eval_nodes(self, x, u, x_next, w=None, out=None, device=0, opts=None):
        """SoA inputs x (nx,n), u (nu,n), x_next (nx,n), w (nx,n)|None -> g (nx,n), Fx (nx*nx,n), Fu (nx*nu,n), H (nh,n)|None."""
        n = x.shape[1]
        if _is_torch(x):
            import torch
            mk = lambda r: torch.empty(r, n, dtype=torch.float64, device=x.device)
        else:
            x, u, x_next, w = _f64(x), _f64(u), _f64(x_next), _f64(w)
            mk = lambda r: np.empty((r, n))
        if out is None:
            out = (mk(self.nx), mk(self.nx * self.nx), mk(self.nx * self.nu), mk(self.nh) if w is not None else None)
        g, Fx, Fu, H = out
        mem = _mem_of(x, u, x_next, g, Fx, Fu)
        st, dev = _stream_of(x)
        _ck(lib().mmpc_eval_nodes(self._h, dev if dev is not None else device, C.byref(opts) if opts else None, n,
                                  _ptr(x)[0], _ptr(u)[0], _ptr(x_next)[0], _ptr(w)[0], _ptr(g)[0], _ptr(Fx)[0],
                                  _ptr(Fu)[0], _ptr(H)[0], mem, st))
        return g, Fx, Fu, H

    def nlp_eval(self, V, p, lam_g=None, lam_f=1.0, device=0, want=("f", "g", "grad_f", "Fx", "Fu", "Hzz"), opts=None):
        """Batched nlp_f/nlp_g/nlp_grad_f/nlp_jac_g/nlp_hess_l. V (B,NV), p (B,np), lam_g (B,M)."""
        V, p, lam_g = _f64(V), _f64(p), _f64(lam_g)
        B = V.shape[0]
        assert V.shape == (B, self.NV) and p.shape == (B, self.np), (V.shape, p.shape, self.NV, self.np)
        shapes = dict(f=(B,), g=(B, self.M), grad_f=(B, self.NV), Fx=(B, self.N, self.nx, self.nx),
                      Fu=(B, self.N, self.nx, self.nu), Hzz=(B, self.N, self.s, self.s))
        out = {k: (np.empty(shapes[k]) if k in want else None) for k in shapes}
        _ck(lib().mmpc_nlp_eval(self._h, device, C.byref(opts) if opts else None, B, _ptr(V)[0], _ptr(p)[0], _ptr(lam_g)[0], float(lam_f),
                                *[_ptr(out[k])[0] for k in ("f", "g", "grad_f", "Fx", "Fu", "Hzz")], MEM_HOST, None))
        return {k: v for k, v in out.items() if v is not None}


class Solver:
    """mmpc_solver_t: `batch` independent ModelControl::calc_u problems solved per call."""

    def __init__(self, model: Model, batch: int, device: int = 0, **opts):
        self.model, self.batch, self.device = model, int(batch), int(device)
        self.opts = default_options(**opts)
        self._h = C.c_void_p()
        _ck(lib().mmpc_solver_create(model._h, self.device, self.batch, C.byref(self.opts), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().mmpc_solver_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, **kw):
        for k, v in kw.items():
            if not hasattr(self.opts, k):
                raise AttributeError(k)
            setattr(self.opts, k, v)
        _ck(lib().mmpc_set_options(self._h, C.byref(self.opts)))

    def set_weights(self, Q=None, R=None, Rm=None, per_instance=False):
        Q, R, Rm = _f64(Q), _f64(R), _f64(Rm)
        _ck(lib().mmpc_set_weights(self._h, _ptr(Q)[0], _ptr(R)[0], _ptr(Rm)[0], int(per_instance)))

    def set_bounds(self, x_min=None, x_max=None, u_min=None, u_max=None):
        a = [_f64(v) for v in (x_min, x_max, u_min, u_max)]
        _ck(lib().mmpc_set_bounds(self._h, *[_ptr(v)[0] for v in a]))

    def reset_warm_start(self):
        _ck(lib().mmpc_reset_warm_start(self._h))

    def solve(self, state, control, traj, V_out=None, lam_out=None, stats_out=None, want_lam=True, want_stats=True):
        """state (B,nx), control (B,nu), traj (B,N,nx) -> V (B,NV), lam (B,M)|None, stats (B,) structured|None.
        numpy in -> numpy out through HOST pointers; torch CUDA in -> torch out through DEVICE pointers."""
        m, B = self.model, self.batch
        if _is_torch(state):
            import torch
            dev = state.device
            V = V_out if V_out is not None else torch.empty(B, m.NV, dtype=torch.float64, device=dev)
            lam = lam_out if lam_out is not None else (torch.empty(B, m.M, dtype=torch.float64, device=dev) if want_lam else None)
            st_t = stats_out if stats_out is not None else (torch.empty(B, 5, dtype=torch.float64, device=dev) if want_stats else None)
            st_ptr = None if st_t is None else C.c_void_p(st_t.data_ptr())
            mem = MEM_DEVICE if state.is_cuda else MEM_HOST
            stream, _ = _stream_of(state)
            _ck(lib().mmpc_solve(self._h, _ptr(state)[0], _ptr(control)[0], _ptr(traj)[0], _ptr(V)[0], _ptr(lam)[0], st_ptr,
                                 mem, stream))
            return V, lam, st_t
        state, control, traj = _f64(state), _f64(control), _f64(traj)
        assert state.shape == (B, m.nx) and control.shape == (B, m.nu) and traj.shape == (B, m.N, m.nx), \
            (state.shape, control.shape, traj.shape)
        V = V_out if V_out is not None else np.empty((B, m.NV))
        lam = lam_out if lam_out is not None else (np.empty((B, m.M)) if want_lam else None)
        st = stats_out if stats_out is not None else (np.empty(B, dtype=STATS_DTYPE) if want_stats else None)
        _ck(lib().mmpc_solve(self._h, _ptr(state)[0], _ptr(control)[0], _ptr(traj)[0], _ptr(V)[0], _ptr(lam)[0],
                             _ptr(st)[0], MEM_HOST, None))
        return V, lam, st

    def closed_loop(self, state, control, traj_long, cycles, plant=RK4, want_logs=True):
        """`cycles` closed-loop steps of every instance on the device. numpy (host) or torch CUDA (device) buffers.
        state (B,nx), control (B,nu) are updated in place; traj_long (B, cycles+N-1, nx).
        Returns X_log (B,cycles,nx), U_log (B,cycles,nu), iters (B,cycles) int32."""
        m, B = self.model, self.batch
        if _is_torch(state):
            import torch
            dev = state.device
            X = torch.empty(B, cycles, m.nx, dtype=torch.float64, device=dev) if want_logs else None
            U = torch.empty(B, cycles, m.nu, dtype=torch.float64, device=dev) if want_logs else None
            it = torch.empty(B, cycles, dtype=torch.int32, device=dev) if want_logs else None
            stream, _ = _stream_of(state)
            _ck(lib().mmpc_closed_loop(self._h, int(cycles), int(plant), _ptr(state)[0], _ptr(control)[0], _ptr(traj_long)[0],
                                       _ptr(X)[0], _ptr(U)[0], None if it is None else C.c_void_p(it.data_ptr()),
                                       MEM_DEVICE if state.is_cuda else MEM_HOST, stream))
            return X, U, it
        assert state.dtype == np.float64 and state.flags["C_CONTIGUOUS"] and control.flags["C_CONTIGUOUS"]
        traj_long = _f64(traj_long)
        assert traj_long.shape == (B, cycles + m.N - 1, m.nx), traj_long.shape
        X = np.empty((B, cycles, m.nx)) if want_logs else None
        U = np.empty((B, cycles, m.nu)) if want_logs else None
        it = np.empty((B, cycles), dtype=np.int32) if want_logs else None
        _ck(lib().mmpc_closed_loop(self._h, int(cycles), int(plant), _ptr(state)[0], _ptr(control)[0], _ptr(traj_long)[0],
                                   _ptr(X)[0], _ptr(U)[0], None if it is None else C.c_void_p(it.ctypes.data), MEM_HOST, None))
        return X, U, it

    def last_timing(self):
        t = (C.c_double * 3)()
        _ck(lib().mmpc_last_timing(self._h, C.byref(t)))
        return dict(h2d_ms=t[0], kernel_ms=t[1], d2h_ms=t[2])

    def last_launch_count(self):
        n = C.c_int()
        _ck(lib().mmpc_last_launch_count(self._h, C.byref(n)))
        return n.value


class Group:
    """mmpc_group_*: one batch sharded over several devices of one process (contiguous shards, no collective on the solve path)."""

    def __init__(self, model: Model, batch_total: int, devices, **opts):
        self.model, self.batch = model, int(batch_total)
        self.opts = default_options(**opts)
        self._h = C.c_void_p()
        dv = (C.c_int * len(devices))(*devices)
        _ck(lib().mmpc_group_create(model._h, dv, len(devices), self.batch, C.byref(self.opts), C.byref(self._h)))

    def close(self):
        if self._h:
            lib().mmpc_group_free(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shards(self):
        n, tot = C.c_int(), C.c_int()
        _ck(lib().mmpc_group_size(self._h, C.byref(n), C.byref(tot)))
        out = []
        for i in range(n.value):
            d, o, c = C.c_int(), C.c_int(), C.c_int()
            _ck(lib().mmpc_group_shard(self._h, i, C.byref(d), C.byref(o), C.byref(c), None))
            out.append((d.value, o.value, c.value))
        return out

    def set_weights(self, Q, R, Rm):
        q, r, rm = (np.ascontiguousarray(a, np.float64) for a in (Q, R, Rm))
        _ck(lib().mmpc_group_set_weights(self._h, q.ctypes.data_as(_dp), r.ctypes.data_as(_dp), rm.ctypes.data_as(_dp)))

    def solve(self, state, control, traj, want_lam=True):
        m, B = self.model, self.batch
        state, control, traj = (np.ascontiguousarray(a, np.float64) for a in (state, control, traj))
        V = np.empty((B, m.NV)); lam = np.empty((B, m.M)) if want_lam else None
        st = np.empty(B, STATS_DTYPE)
        _ck(lib().mmpc_group_solve(self._h, state.ctypes.data_as(_dp), control.ctypes.data_as(_dp), traj.ctypes.data_as(_dp), V.ctypes.data_as(_dp),
                                   lam.ctypes.data_as(_dp) if want_lam else None, st.ctypes.data_as(C.POINTER(Stats))))
        return V, lam, st

    def gather(self, root_device, V_root_ptr, stats_root_ptr=None):
        ms = C.c_double()
        _ck(lib().mmpc_group_gather(self._h, int(root_device), C.c_void_p(V_root_ptr), C.c_void_p(stats_root_ptr) if stats_root_ptr else None, C.byref(ms)))
        return ms.value


def stats_from_tensor(t):
    """View a (B,5) float64 torch stats tensor (device or host) as the structured numpy dtype."""
    return np.frombuffer(t.detach().cpu().numpy().tobytes(), dtype=STATS_DTYPE)


def device_count() -> int:
    n = C.c_int()
    _ck(lib().mmpc_device_count(C.byref(n)))
    return n.value


def measure_fp64_peak(device=0) -> float:
    """Measured DFMA throughput of the device in TFLOP/s."""
    v = C.c_double()
    _ck(lib().mmpc_measure_fp64_peak(device, C.byref(v)))
    return v.value
