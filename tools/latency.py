"""Single-solve latency through mmpc_solve with host pointers: zero-copy path (page-locked staging read / written by the kernel in
place) against the packed-copy path (MMPC_ZERO_COPY=0), plus the CPU restatement on one core.  The two paths must return the
same bits.  One process per setting because the environment switch is read once.

    python -m tools.latency            # driver: runs both settings, prints one JSON line
"""
import json
import os
import subprocess
import sys
import time

import numpy as np


def one(setting):
    import mahi_mpc_b200.api as api
    from mahi_mpc_b200 import workload as w
    out = {}
    for (B, N) in (((1, 25), (1, 50), (4, 25), (32, 25), (32, 50), (64, 25)) if not os.environ.get('MMPC_LAT_SIZES') else [tuple(int(v) for v in t.split('x')) for t in os.environ['MMPC_LAT_SIZES'].split(',')]):
        x0, up, traj = w.synthetic_batch(4, 2, N, 0.002, B)
        Q, R, Rm = w.default_weights(4, 2)
        m = api.Model("double_pendulum", N=N)
        s = api.Solver(m, B, warm_start=0)
        s.set_weights(Q, R, Rm)
        ts = []
        for i in range(220):
            t0 = time.perf_counter()
            V, lam, st = s.solve(x0, up, traj, want_lam=False)
            ts.append(time.perf_counter() - t0)
        ts = sorted(ts[20:])
        V, lam, st = s.solve(x0, up, traj)
        out[f"B{B}_N{N}"] = {"p50_us": 1e6 * ts[len(ts) // 2], "p99_us": 1e6 * ts[int(len(ts) * 0.99)], "kernel_us": 1e3 * s.last_timing()["kernel_ms"],
                             "iters": [int(st["iterations"].min()), int(st["iterations"].max())], "conv": int((st["status"] == 0).sum()),
                             "V_sum": float(np.abs(V).sum()), "lam_sum": float(np.abs(lam).sum()), "obj_sum": float(st["objective"].sum())}
        # bounded solve (T3) through the same path
        s.set_bounds(u_min=[-2.5, -2.5], u_max=[2.5, 2.5])
        s.reset_warm_start()
        tb = []
        for i in range(60):
            t0 = time.perf_counter()
            Vb, _, stb = s.solve(x0, up, traj, want_lam=False)
            tb.append(time.perf_counter() - t0)
        tb = sorted(tb[10:])
        out[f"B{B}_N{N}"]["bounded"] = {"p50_us": 1e6 * tb[len(tb) // 2], "V_sum": float(np.abs(Vb).sum()), "conv": int((stb["status"] == 0).sum())}
    if setting == "1":
        from oracle import coracle, nlp
        co = coracle.COracle("O3")
        for N in (25, 50):
            x0, up, traj = w.synthetic_batch(4, 2, N, 0.002, 1)
            P = nlp.Problem(N=N)
            p = np.stack([P.pack_p(traj[0], Q, R, Rm, up[0])])
            tc = []
            for i in range(120):
                t0 = time.perf_counter()
                Vo, _, sto, bad = co.solve_batch("double_pendulum", N, 0.002, False, x0, p, tol=1e-5, max_iter=200, nthreads=1)
                tc.append(time.perf_counter() - t0)
            tc = sorted(tc[20:])
            out[f"cpu_port_N{N}"] = {"p50_us": 1e6 * tc[len(tc) // 2], "p99_us": 1e6 * tc[int(len(tc) * 0.99)]}
            m = api.Model("double_pendulum", N=N)
            s = api.Solver(m, 1, warm_start=0, tol=1e-10)
            s.set_weights(Q, R, Rm)
            V, _, st = s.solve(x0, up, traj)
            Vo, _, sto, bad = co.solve_batch("double_pendulum", N, 0.002, False, x0, p, tol=1e-10)
            out[f"cpu_port_N{N}"]["max_abs_V_diff_tol1e-10"] = float(np.abs(V - Vo).max())
    print("LATENCY " + json.dumps(out))


def main():
    if len(sys.argv) > 1:
        return one(sys.argv[1])
    res = {}
    for setting in ("1", "0"):
        env = dict(os.environ, MMPC_ZERO_COPY=setting)
        r = subprocess.run([sys.executable, "-m", "tools.latency", setting], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("LATENCY ")]
        if not line:
            print(r.stdout[-3000:])
            raise SystemExit(1)
        res["zero_copy" if setting == "1" else "packed_copies"] = json.loads(line[0][8:])
    same = all(res["zero_copy"][k][f] == res["packed_copies"][k][f] for k in res["packed_copies"] if k.startswith("B") for f in ("V_sum", "lam_sum", "obj_sum", "iters"))
    same_b = all(res["zero_copy"][k]["bounded"]["V_sum"] == res["packed_copies"][k]["bounded"]["V_sum"] for k in res["packed_copies"] if k.startswith("B"))
    res["identical_results"] = bool(same and same_b)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
