"""Developer tool: end-to-end (pinned host buffers) timing of mmpc_solve for the bench workload.

    python -m tools.e2e_time [B]          one setting (MMPC_ZC_BIG / MMPC_PIPE / MMPC_VARIANT from the environment)
    python -m tools.e2e_time sweep [B]    copy engines (MMPC_ZC_BIG=0) vs zero-copy inputs (1) vs inputs + outputs (2), each with
                                          1 / 2 / 4 / 8 sub-batches; the results of all settings must be the same bits
"""
import json
import os
import subprocess
import sys
import time

import numpy as np


def one(B):
    import torch
    import mahi_mpc_b200.api as api
    from mahi_mpc_b200 import workload
    N = 50
    x0, up, traj = workload.synthetic_batch(4, 2, N, 0.002, B)
    Q, R, Rm = workload.default_weights(4, 2)
    m = api.Model("double_pendulum", N=N)
    s = api.Solver(m, B, warm_start=0, kernel_variant=int(os.environ.get("MMPC_VARIANT", "0")))
    s.set_weights(Q, R, Rm)
    h = [torch.from_numpy(a).pin_memory() for a in (x0, up, traj)]
    V = torch.empty(B, m.NV, dtype=torch.float64).pin_memory(); st = torch.empty(B, 5, dtype=torch.float64).pin_memory()
    lam = torch.empty(B, m.M, dtype=torch.float64).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5):
        s.solve(*h, V_out=V, stats_out=st, want_lam=False)
    ts = []
    for _ in range(40):
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter(); s.solve(*h, V_out=V, stats_out=st, want_lam=False); ts.append(time.perf_counter() - t0)
    ts = np.array(ts) * 1e3
    s.solve(*h, V_out=V, lam_out=lam, stats_out=st)
    stats = api.stats_from_tensor(st)
    out = {"B": B, "median_ms": float(np.median(ts)), "min_ms": float(ts.min()), "solves_per_s": B / float(np.median(ts)) * 1e3,
           "span": s.last_timing(), "conv": int((stats["status"] == 0).sum()), "V_sum": float(V.abs().sum()), "lam_sum": float(lam.abs().sum()),
           "obj_sum": float(stats["objective"].sum()), "iters": int(stats["iterations"].sum())}
    print("E2E " + json.dumps(out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "sweep":
        B = sys.argv[2] if len(sys.argv) > 2 else "4096"
        res = {}
        for zc in os.environ.get("MMPC_SWEEP_ZC", "0,1,2,3").split(","):
            for pipe in os.environ.get("MMPC_SWEEP_PIPE", "1,2,4,8").split(","):
                env = dict(os.environ, MMPC_ZC_BIG=zc, MMPC_PIPE=pipe)
                r = subprocess.run([sys.executable, "-m", "tools.e2e_time", B], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
                line = [l for l in r.stdout.splitlines() if l.startswith("E2E ")]
                if not line:
                    print(r.stdout[-2000:])
                    continue
                d = json.loads(line[0][4:])
                res[f"zc{zc}_pipe{pipe}"] = d
                print(f"zc={zc} pipe={pipe}: median {d['median_ms']:.3f} ms min {d['min_ms']:.3f} -> {d['solves_per_s']:.3e} /s", flush=True)
        keys = list(res)
        same = all(res[k][f] == res[keys[0]][f] for k in keys for f in ("V_sum", "lam_sum", "obj_sum", "iters", "conv"))
        print(json.dumps({"identical_results": bool(same), "settings": res}))
        return
    one(int(sys.argv[1]) if len(sys.argv) > 1 else 4096)


if __name__ == "__main__":
    main()
