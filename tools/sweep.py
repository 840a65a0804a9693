"""Measures the BASELINE configs that are not the headline bench line and writes profiles/<round>_sweep.json:
  config 3  4-DOF arm, N=50, batch 16384 sharded over the visible GPUs (run under torchrun for >1 GPU)
  config 4  horizon sweep N = 10/50/100/200 at batch 4096 (double pendulum)
  config 5  1e7 node evaluations (rollout+Jacobian, Euler and RK4x4) vs the CPU restatement on the host cores
  plus      closed-loop steps/s of the device-resident loop and single-solve latency
    python -m tools.sweep [out.json]
All timings: CUDA events on the launching stream after 3 warm-up calls, inputs resident in HBM, cold-start solves.
"""
import json
import os
import sys
import time

import numpy as np
import torch

import mahi_mpc_b200.api as api
from mahi_mpc_b200 import workload


def timed(fn, reps=5, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def solve_rate(model_name, nx, nu, N, B, dev, scale=1.0, **opts):
    x0, up, traj = workload.synthetic_batch(nx, nu, N, 0.002, B)
    traj *= scale
    Q, R, Rm = workload.default_weights(nx, nu)
    m = api.Model(model_name, N=N)
    s = api.Solver(m, B, warm_start=0, **opts)
    s.set_weights(Q, R, Rm)
    d = [torch.from_numpy(a).to(dev) for a in (x0, up, traj)]
    V = torch.empty(B, m.NV, dtype=torch.float64, device=dev); st = torch.empty(B, 5, dtype=torch.float64, device=dev)
    ms = timed(lambda: s.solve(*d, V_out=V, stats_out=st, want_lam=False))
    stats = api.stats_from_tensor(st)
    return dict(model=model_name, N=N, batch=B, ms=ms, solves_per_s=B / ms * 1e3, converged=float((stats["status"] == 0).mean()),
                iterations_mean=float(stats["iterations"].mean()), node_evals_per_s=B * N * (2 * stats["iterations"].mean() + 1) / ms * 1e3)


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sweep.json"
    dev = torch.device("cuda", 0)
    res = {"gpu": torch.cuda.get_device_name(0)}
    res["config4_horizon_sweep"] = [solve_rate("double_pendulum", 4, 2, N, 4096, dev) for N in (10, 50, 100, 200)]
    res["config3_arm4_single_gpu"] = [solve_rate("arm4", 8, 4, 50, B, dev, scale=0.5) for B in (2048, 16384)]
    # config 5
    n = 10_000_000
    m = api.Model("double_pendulum", N=25)
    g = torch.Generator(device=dev); g.manual_seed(5)
    x = torch.rand(4, n, dtype=torch.float64, device=dev, generator=g) * 4 - 2
    u = torch.rand(2, n, dtype=torch.float64, device=dev, generator=g) * 20 - 10
    xn = torch.rand(4, n, dtype=torch.float64, device=dev, generator=g)
    w = torch.rand(4, n, dtype=torch.float64, device=dev, generator=g)
    out = (torch.empty(4, n, dtype=torch.float64, device=dev), torch.empty(16, n, dtype=torch.float64, device=dev),
           torch.empty(8, n, dtype=torch.float64, device=dev), None)
    outH = out[:3] + (torch.empty(21, n, dtype=torch.float64, device=dev),)
    hbm = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0
    c5 = {}
    for name, fn, byts in (("euler_g_Fx_Fu", lambda: m.eval_nodes(x, u, xn, out=out), 304),
                           ("euler_g_Fx_Fu_H", lambda: m.eval_nodes(x, u, xn, w, out=outH), 304 + 32 + 168),
                           ("rk4x4_g_Fx_Fu", lambda: m.eval_nodes(x, u, xn, out=out, opts=api.default_options(integrator=api.RK4, substeps=4)), 304)):
        ms = timed(fn)
        c5[name] = dict(ms=ms, node_evals_per_s=n / ms * 1e3, gbs=n * byts / ms / 1e6, hbm_frac=n * byts / ms / 1e6 / hbm, bytes_per_node=byts)
    from oracle import coracle  # CPU leg (the checker, timed as the host baseline)
    co = coracle.COracle("O3")
    ns = 400_000
    xs, us, xns = (a[:, :ns].T.contiguous().cpu().numpy() for a in (x, u, xn))
    t0 = time.perf_counter(); co.node_eval("double_pendulum", m.h, xs, us, xns); dt = time.perf_counter() - t0
    c5["cpu_port"] = dict(node_evals_per_s=ns / dt, cores=os.cpu_count(), sample=f"{ns} nodes, oracle/mpc_oracle.c -O3, OpenMP")
    res["config5_node_eval_1e7"] = c5
    # closed loop
    N, B, cycles = 25, 4096, 100
    x0, up, tl = workload.synthetic_batch(4, 2, cycles + N - 1, 0.002, B)
    Q, R, Rm = workload.default_weights(4, 2)
    m2 = api.Model("double_pendulum", N=N)
    s = api.Solver(m2, B)
    s.set_weights(Q, R, Rm)
    dx, du, dtl = (torch.from_numpy(a).to(dev) for a in (x0, up, tl))
    def loop():
        s.reset_warm_start()
        s.closed_loop(dx.clone(), du.clone(), dtl, cycles, plant=api.RK4, want_logs=False)
    ms = timed(loop, reps=3, warm=1)
    res["closed_loop"] = dict(N=N, batch=B, cycles=cycles, ms=ms, steps_per_s=B * cycles / ms * 1e3,
                              note="config-1 loop (solve every 2 ms cycle, RK4 plant), 4096 loops in parallel, warm started")
    os.makedirs(os.path.dirname(out_path) or ".", exist_ok=True)
    with open(out_path, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
