#!/usr/bin/env python
"""bench.py -- BASELINE.json's headline metric on B200: converged MPC solves/sec at batch 4096, N = 50.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this framework (CUDA path through the C ABI)
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # the CPU path, on the box's host cores

One "step" = one pass of the hot path over one batch: `batch` independent cold-start
ModelControl::calc_u problems (double pendulum of examples/ex_model_generate.cpp, N = 50, h = 2 ms,
Q={10,1,5,5}, R={5,5}, Rm={1,1}, synthetic x0 / sinusoid references, IPOPT's tol 1e-5 / max_iter 200 as in
ModelControl.cpp:54-55).  value = converged solves / s, whole job (all ranks), inputs resident in HBM;
e2e = the same through mmpc_solve with pinned HOST buffers (H2D of state/control/traj and D2H of the plans and
statistics inside the timed region).  With --gpus N>1 (torchrun) every rank solves its own shard of
N*batch instances (weak scaling, no collective on the solve path); a final NCCL gather of plans + statistics
to rank 0 is timed and reported separately.

The reference arm times the CPU restatement of the reference's path (oracle/mpc_oracle.c: generated C
callbacks + Newton/KKT with a band LU, one instance per OpenMP thread) because CasADi/IPOPT/MUMPS are not
installable in this image.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "converged MPC solves/sec at batch 4096, N=50"
UNIT = "solves/s"
N_NODES, BATCH, NX, NU, H_STEP = 50, 4096, 4, 2, 0.002
WORKLOAD = "BASELINE configs[1]: double pendulum (ex_model_generate), N=50, batch 4096 per GPU, cold start, tol 1e-5"


print_json = lambda obj: print(json.dumps(obj))


def config_of(n_gpus):
    """The workload description: IDENTICAL in both arms (the driver compares the two dicts)."""
    return {"workload": WORKLOAD, "model": "double_pendulum", "N": N_NODES, "batch_per_gpu": BATCH, "global_batch": BATCH * n_gpus,
            "start": "cold (zeros, x0 pinned)", "tol": 1e-5, "max_iter": 200}


def solve_flops(ops, nx, nu, N, iters_mean):
    """Algorithmic fp64 flops of one solve: per Newton iteration and node one evaluation of f, A, B and
    the Hessian contraction (backward sweep), one Riccati stage on the augmented state (x_k, u_{k-1}) and the linear roll-out
    (forward sweep), one model evaluation for the merit / Armijo test; plus the final convergence sweep."""
    fl_eval = sum(ops["fABH"]) + sum(ops["fAB"]) + sum(ops["f"])
    nxa = nx + nu
    fl_ric = 7 * nxa ** 3 / 3 + 4 * nxa ** 2 * nu + 2 * nxa * nu ** 2 + nu ** 3 / 3
    fl_fwd = 2 * (nu * (nx + nu) + nx * (nx + nu) + nx * (nx + nu))
    return iters_mean * N * (fl_eval + fl_ric + fl_fwd) + N * sum(ops["fABH"])


def workload(batch, offset=0):
    from mahi_mpc_b200 import workload as w  # synthetic-input generator; identical to the oracle's copy
    x0, up, traj = w.synthetic_batch(NX, NU, N_NODES, H_STEP, batch, offset=offset)
    Q, R, Rm = w.default_weights(NX, NU)
    return x0, up, traj, (Q, R, Rm)


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md "clocks DURING the timed region")
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self, interval_ms=20):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", str(interval_ms),
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def halt(self):
        """Stop the running nvidia-smi loop; the rows read so far are kept (start() may follow with another interval)."""
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            self.proc = None

    def stop(self):
        self.halt()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        busy = sorted(sm)[len(sm) // 2:]  # the upper half = samples under load
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the committed `ncu --set full` capture
    (profiles/ncu_traffic.json names the capture each figure was read from); None if there is none."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh)[kernel]["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# CPU legs (the only places allowed to execute oracle/)
# ------------------------------------------------------------------------------------------------
def cpu_solve_rate(sample, opt="O3", threads=0, repeats=1, model="double_pendulum", N=N_NODES, traj_scale=1.0, mean=False):
    """solves/s of the CPU restatement on `sample` instances of a bench workload with all host threads (threads = 1: one core)."""
    from mahi_mpc_b200 import workload as w
    from oracle import coracle, nlp
    co = coracle.COracle(opt)
    nx, nu = (8, 4) if model == "arm4" else (NX, NU)
    x0, up, traj = w.synthetic_batch(nx, nu, N, H_STEP, sample)
    traj = traj * traj_scale
    Q, R, Rm = w.default_weights(nx, nu)
    P = nlp.Problem(model=model, N=N)
    p = np.stack([P.pack_p(traj[b], Q, R, Rm, up[b]) for b in range(sample)])
    nthr = threads or (os.cpu_count() or 1)
    best, conv, total = None, 0, 0.0
    for _ in range(repeats):
        t0 = time.perf_counter()
        _, _, st, bad = co.solve_batch(model, N, H_STEP, False, x0, p, tol=1e-5, max_iter=200, nthreads=nthr)
        dt = time.perf_counter() - t0
        conv = int((st["status"] == 0).sum())
        best = dt if best is None else min(best, dt)
        total += dt
    if mean:  # sustained rate over all passes (the cpu_baseline leg: ~10 s of CPU work)
        return conv * repeats / total, nthr, total, conv
    return conv / best, nthr, best, conv


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = 2048
    rate0, nthr, dt, _ = cpu_solve_rate(256, "O3")  # calibrate so one step is a few seconds at most
    sample = int(min(BATCH, max(256, rate0 * 4.0)))
    for _ in range(args.warmup):
        cpu_solve_rate(min(sample, 256), "O3")
    times, convs = [], []
    for _ in range(args.steps):
        r, nthr, dt, conv = cpu_solve_rate(sample, "O3")
        times.append(dt); convs.append(conv)
    value = sum(convs) / sum(times)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_of(args.gpus), "run": {"sample_per_step": sample, "threads": nthr},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": nthr, "kind": "port",
                         "sample": f"{sample} of the {BATCH} instances per step, oracle/mpc_oracle.c compiled gcc -O3 "
                                   f"(the reference builds its callbacks -O1, ModelGenerator.cpp:256; -O3 is the faster, conservative baseline), OpenMP over instances; CasADi/IPOPT/MUMPS not installable here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print_json(line)
    return 0


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist

    import mahi_mpc_b200.api as api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, N = BATCH, N_NODES
    x0, up, traj, (Q, R, Rm) = workload(B, offset=rank * B)
    model = api.Model("double_pendulum", N=N)
    solver = api.Solver(model, B, device=local, warm_start=0, kernel_variant=args.variant)  # tol/max_iter = reference defaults
    solver.set_weights(Q, R, Rm)
    d_x0, d_up, d_traj = (torch.from_numpy(a).to(dev) for a in (x0, up, traj))
    d_V = torch.empty(B, model.NV, dtype=torch.float64, device=dev)
    d_st = torch.empty(B, 5, dtype=torch.float64, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)  # > 126 MB L2

    def step_resident():
        solver.solve(d_x0, d_up, d_traj, V_out=d_V, stats_out=d_st, want_lam=False)

    # the clock sampler is started BEFORE the warm-up: nvidia-smi needs ~0.1-0.3 s to deliver its first row, longer than the
    # timed region of a short run; the warm-up keeps the GPU under load while it spins up
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_wait = time.perf_counter()
        while not sampler.rows and time.perf_counter() - t_wait < 1.0:
            step_resident()
    for _ in range(max(args.warmup, 3)):
        step_resident()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    barrier()
    for k in range(args.steps):
        flush.fill_(float(k))  # evict the previous step's workspace from L2 (outside the timed events)
        ev[k][0].record()
        step_resident()
        ev[k][1].record()
        launches += solver.last_launch_count()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    st = api.stats_from_tensor(d_st)
    conv_frac = float((st["status"] == 0).mean())
    iters_mean = float(st["iterations"].mean())
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    conv_count = torch.tensor([float((st["status"] == 0).sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(conv_count, op=dist.ReduceOp.SUM)
    total_s = total_ms.item() * 1e-3
    value = conv_count.item() * args.steps / total_s

    # ---- e2e: pinned host buffers through the C ABI, copies inside the timed region ----------------
    h_x0, h_up, h_traj = (torch.from_numpy(a).pin_memory() for a in (x0, up, traj))
    h_V = torch.empty(B, model.NV, dtype=torch.float64).pin_memory()
    h_st = torch.empty(B, 5, dtype=torch.float64).pin_memory()
    for _ in range(3):
        solver.solve(h_x0, h_up, h_traj, V_out=h_V, stats_out=h_st, want_lam=False)
    # This leg is timed on the HOST clock, and every nvidia-smi query holds driver locks that the calls of the step wait for (measured:
    # the same loop reads 0.55 ms per step without the sampler and 0.60-0.74 ms beside a 20 ms loop).  The clocks are still sampled
    # during the leg, every 50 ms instead of the 20 ms of the device-timed leg (the profiling recipe's own period is 200 ms).
    if rank == 0:
        n_rows = len(sampler.rows)
        sampler.halt()
        sampler.start(interval_ms=50)
        t_wait = time.perf_counter()
        while len(sampler.rows) == n_rows and time.perf_counter() - t_wait < 1.0:
            solver.solve(h_x0, h_up, h_traj, V_out=h_V, stats_out=h_st, want_lam=False)
    barrier()
    t_e2e = []
    for k in range(args.steps):
        flush.fill_(float(k))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        solver.solve(h_x0, h_up, h_traj, V_out=h_V, stats_out=h_st, want_lam=False)  # synchronous for HOST memory
        t_e2e.append(time.perf_counter() - t0)
        launches += solver.last_launch_count()
    e2e_t = torch.tensor([sum(t_e2e)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    st_h = api.stats_from_tensor(h_st)
    e2e_value = float((st_h["status"] == 0).sum()) * world * args.steps / e2e_t.item()
    clocks = sampler.stop() if rank == 0 else None
    h2d = (h_x0.numel() + h_up.numel() + h_traj.numel()) * 8
    d2h = (h_V.numel() + h_st.numel()) * 8

    # ---- final gather of plans + statistics to rank 0 (the only NVLink traffic) -----------
    gather_ms = None
    if world > 1:
        from mahi_mpc_b200 import shard
        shard.gather_to_root([d_V, d_st])  # warm-up (NCCL communicator setup)
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        gathered = shard.gather_to_root([d_V, d_st])
        g1.record()
        barrier()
        if rank == 0:  # rank r's block is rows [r B, (r + 1) B)
            assert gathered[0].shape[0] == world * B and gathered[1].shape[0] == world * B and torch.equal(gathered[0][:B], d_V)
        t = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        gather_ms = t.item()

    if rank == 0:
        ms_per_step = total_s * 1e3 / args.steps
        hbm_peak, peak_src = measured_peaks()
        kern_s = (sum(step_ms) / len(step_ms)) * 1e-3
        # algorithmic bytes per solve: inputs (nx + nu + N nx) + outputs (NV + 5 stats words)
        bytes_per_solve = 8 * (NX + NU + N * NX) + 8 * (model.NV + 5)
        ops = model.op_counts()
        fp64_peak = api.measure_fp64_peak(local)
        flops_per_solve = solve_flops(ops, NX, NU, N, iters_mean)
        hbm = {"bound": "hbm", "achieved": B * bytes_per_solve / kern_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
               "frac": B * bytes_per_solve / kern_s / 1e9 / hbm_peak, "peak_source": peak_src, "algorithmic_bytes_per_solve": bytes_per_solve}
        # the solve is bound by the fp64 pipe / its dependent chains, not by HBM: THE roofline of this kernel is fp64
        roofline = {"bound": "fp64", "achieved": B * flops_per_solve / kern_s / 1e12, "peak": fp64_peak, "unit": "TFLOP/s",
                    "frac": B * flops_per_solve / kern_s / 1e12 / fp64_peak if fp64_peak else None,
                    "traffic": ncu_traffic("solve_kernel_t4"), "kernel": "solve_kernel_t4<double_pendulum, 32, 7, 5>",
                    "peak_source": "measured in this run (mmpc_measure_fp64_peak: independent DFMA chains on every SM)",
                    "algorithmic_flops_per_solve": flops_per_solve, "hbm": hbm}
        fp64 = dict(roofline)  # round-1 key, kept for continuity
        extras = node_eval_bench(api, model, dev, hbm_peak) if world == 1 else None
        lat = latency_bench(api, dev, with_cpu=not args.no_cpu) if world == 1 else None
        others = other_configs(api, dev, fp64_peak, with_cpu=not args.no_cpu) if world == 1 and not args.headline_only else None
        cpu = None
        if world == 1 and not args.no_cpu:
            r0, nthr, _, _ = cpu_solve_rate(128, "O3")
            sample = int(min(BATCH, max(128, r0 * 12.0)))
            r1, _, _, _ = cpu_solve_rate(sample, "O3")  # warm pass: the rate the passes below will run at
            passes = max(1, int(10.0 * r1 / sample))  # about ten seconds of CPU work
            rate, nthr, dt, conv = cpu_solve_rate(sample, "O3", repeats=passes, mean=True)
            cpu = {"value": rate, "unit": UNIT, "cores": nthr, "kind": "port",
                   "sample": f"{passes} passes over {sample} of the {BATCH} instances (same seeds), oracle/mpc_oracle.c gcc -O3, OpenMP over "
                             f"instances, {dt:.1f} s in all; restated CPU baseline -- CasADi/IPOPT/MUMPS not installable in this image"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config_of(world),
            "run": {"parallelism": f"instances sharded over {world} GPU(s), no collective on the solve path",
                    "l2": "flushed between timed steps (256 MB write outside the event pair)", "kernel_variant": args.variant},
            "converged_fraction": conv_frac, "iterations_mean": iters_mean,
            "node_evals_per_s": value * N * (2 * iters_mean + 1),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "path": "mmpc_solve(MMPC_MEM_HOST) on page-locked host buffers: the tile kernels read the inputs and write the plans / "
                            "statistics in place over PCIe (zero-copy, 8 sub-batches on their own streams); every byte crosses PCIe inside "
                            "the timed region, MMPC_ZC_BIG=0 selects the copy engines instead"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "fp64_roofline": fp64,
            "cpu_baseline": cpu, "node_eval": extras, "latency": lat, "gather_ms": gather_ms, "other_configs": others,
        }
        print_json(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def node_eval_bench(api, model, dev, hbm_peak):
    """BASELINE config 5: 10^7 shooting-node evaluations (F, g, dF/dx, dF/du), the HBM-bound streaming kernel."""
    import torch
    n = 10_000_000
    g = torch.Generator(device=dev); g.manual_seed(5)
    x = torch.rand(NX, n, dtype=torch.float64, device=dev, generator=g) * 4 - 2
    u = torch.rand(NU, n, dtype=torch.float64, device=dev, generator=g) * 20 - 10
    xn = torch.rand(NX, n, dtype=torch.float64, device=dev, generator=g)
    out = (torch.empty(NX, n, dtype=torch.float64, device=dev), torch.empty(NX * NX, n, dtype=torch.float64, device=dev),
           torch.empty(NX * NU, n, dtype=torch.float64, device=dev), None)
    for _ in range(3):
        model.eval_nodes(x, u, xn, out=out)
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); model.eval_nodes(x, u, xn, out=out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e-3)
    t = statistics.median(ts)
    bytes_per_node = 8 * (2 * NX + NU) + 8 * (NX + NX * NX + NX * NU)  # 80 read + 224 written = 304 B
    return {"workload": "BASELINE configs[4]: 1e7 node evaluations (g, dF/dx, dF/du dense), inputs 800 MB > L2", "nodes": n,
            "node_evals_per_s": n / t, "ms": t * 1e3,
            "roofline": {"bound": "hbm", "achieved": n * bytes_per_node / t / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": n * bytes_per_node / t / 1e9 / hbm_peak, "traffic": ncu_traffic("eval_nodes_kernel"), "bytes_per_node": bytes_per_node}}


def latency_bench(api, dev, with_cpu=True):
    """p50 single-solve latency, B = 1, N = 25 closed-loop setting of config 1 (host pointers, cold start), next to the CPU
    restatement's single-thread latency on this box (the reference's budget is a 1 kHz loop,
    thread_model_control_example.cpp:70-71); the same at the benchmark horizon N = 50 under "N50"."""
    out = _latency_one(api, dev, 25, with_cpu)
    out["N50"] = _latency_one(api, dev, 50, with_cpu)
    return out


def _latency_one(api, dev, N, with_cpu):
    from mahi_mpc_b200 import workload as w
    x0, up, traj = w.synthetic_batch(NX, NU, N, H_STEP, 1)
    Q, R, Rm = w.default_weights(NX, NU)
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, 1, device=dev.index or 0, warm_start=0)
    s.set_weights(Q, R, Rm)
    ts = []
    for i in range(110):
        t0 = time.perf_counter()
        s.solve(x0, up, traj, want_lam=False)
        ts.append(time.perf_counter() - t0)
    ts = sorted(ts[10:])
    out = {"workload": f"B=1, N={N}, cold start, host pointers (one ModelControl::calc_u)", "p50_us": 1e6 * ts[len(ts) // 2],
           "p99_us": 1e6 * ts[-1], "kernel_us": 1e3 * s.last_timing()["kernel_ms"],
           "note": "zero-copy staging (no copy-engine transfers); kernel_us = launch .. synchronised on the host clock"}
    if with_cpu:
        from oracle import coracle, nlp
        co = coracle.COracle("O3")
        P = nlp.Problem(N=N)
        p = np.stack([P.pack_p(traj[0], Q, R, Rm, up[0])])
        tc = []
        for i in range(110):
            t0 = time.perf_counter()
            co.solve_batch("double_pendulum", N, H_STEP, False, x0, p, tol=1e-5, max_iter=200, nthreads=1)
            tc.append(time.perf_counter() - t0)
        tc = sorted(tc[10:])
        out["cpu_port_p50_us"] = 1e6 * tc[len(tc) // 2]
        out["cpu_port_p99_us"] = 1e6 * tc[-1]
        out["faster_for_one_solve"] = "cpu port (one core)" if out["cpu_port_p50_us"] < out["p50_us"] else "gpu"
    return out


def other_configs(api, dev, fp64_peak, with_cpu=True):
    """BASELINE configs[2] (4-DOF arm, N = 50, batch 16384: here on ONE GPU, the multi-GPU runs are tools/config3_multi.py) and
    configs[3] (horizon sweep N = 10/50/100/200 at batch 4096), each with its fp64 roofline fraction; bounded to a few seconds."""
    import torch
    from mahi_mpc_b200 import workload as w

    def timed(model_name, nx, nu, N, B, reps, traj_scale=1.0):
        x0, up, traj = w.synthetic_batch(nx, nu, N, H_STEP, B)
        traj = traj * traj_scale
        Q, R, Rm = w.default_weights(nx, nu)
        m = api.Model(model_name, N=N)
        s = api.Solver(m, B, device=dev.index or 0, warm_start=0)
        s.set_weights(Q, R, Rm)
        d = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x0, up, traj)]
        V = torch.empty(B, m.NV, dtype=torch.float64, device=dev)
        st = torch.empty(B, 5, dtype=torch.float64, device=dev)
        for _ in range(2):
            s.solve(*d, V_out=V, stats_out=st, want_lam=False)
        ts = []
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); s.solve(*d, V_out=V, stats_out=st, want_lam=False); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        stats = api.stats_from_tensor(st)
        ms = statistics.median(ts)
        conv = float((stats["status"] == 0).sum())
        it = float(stats["iterations"].mean())
        fl = solve_flops(m.op_counts(), nx, nu, N, it)
        ach = B * fl / (ms * 1e-3) / 1e12
        out = {"model": model_name, "N": N, "batch": B, "ms": ms, "solves_per_s": conv / ms * 1e3, "converged_fraction": conv / B,
               "iterations_mean": it, "node_evals_per_s": conv / ms * 1e3 * N * (2 * it + 1),
               "roofline": {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak if fp64_peak else None,
                            "algorithmic_flops_per_solve": fl}}
        s.close()
        return out

    res = {"horizon_sweep": [timed("double_pendulum", NX, NU, N, BATCH, 5) for N in (10, 50, 100, 200)]}
    c3 = timed("arm4", 8, 4, 50, 16384, 3, traj_scale=0.5)
    c3["workload"] = "BASELINE configs[2]: 4-DOF arm with torque cost, N=50, batch 16384 on ONE GPU (2/4/8 GPUs: tools/config3_multi.py, profiles/)"
    if with_cpu:
        rate, nthr, dt, conv = cpu_solve_rate(256, "O3", model="arm4", N=50, traj_scale=0.5)
        c3["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": nthr, "kind": "port",
                              "sample": f"256 of the 16384 instances, oracle/mpc_oracle.c + oracle/arm4_jets.c gcc -O3, {dt:.1f} s"}
    res["config3_arm4"] = c3
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--variant", type=int, default=0, help="solve kernel variant (0 = auto)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--headline-only", action="store_true", help="skip the other BASELINE configs (horizon sweep, 4-DOF arm)")
    args = ap.parse_args()
    # Libraries (NCCL prints its version banner) write to fd 1; the contract wants exactly ONE JSON line on stdout.
    # Route fd 1 to stderr for the duration of the run and emit the JSON line on the real stdout.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    lines = []
    global print_json
    print_json = lambda obj: lines.append(json.dumps(obj))
    try:
        rc = run_reference(args) if args.impl == "reference" else run_gpu(args)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
    for ln in lines:
        os.write(1, (ln + "\n").encode())
    return rc


if __name__ == "__main__":
    sys.exit(main())
