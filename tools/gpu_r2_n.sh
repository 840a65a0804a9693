mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "group or packed or facade or reference_drivers or warm_start" > gpurun_out/n_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/n_pytest.log
tail -15 gpurun_out/n_pytest.log
python bench.py --steps 30 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc $?"; tail -3 gpurun_out/n_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/n_bench.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "e2e", "roofline", "latency", "cpu_baseline"):
    print(k, d.get(k))
for r in d["other_configs"]["horizon_sweep"]:
    print("sweep", r["N"], r["ms"], r["solves_per_s"], r["roofline"]["frac"])
print("config3", d["other_configs"]["config3_arm4"])
PY
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/n_bench_ref.json 2>/dev/null; head -c 600 gpurun_out/n_bench_ref.json
