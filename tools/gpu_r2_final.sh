mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/final_pytest.log
tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/final_bench.err; echo "bench rc $?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; echo "ref rc $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["latency"]["p50_us"], d["latency"]["cpu_port_p50_us"])
print([ (r["N"], round(r["ms"],3), round(r["roofline"]["frac"],3)) for r in d["other_configs"]["horizon_sweep"]], d["other_configs"]["config3_arm4"]["ms"])
PY
