# Round-2 final measurements on ONE box: tests, smoke, bench (both arms), ncu launch list and one full capture of the bench kernel.
bash tools/gpu_ab_lib.sh
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/final_pytest.log
tail -3 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1

timeout 600 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/final_bench.err; echo "bench rc $?"
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; echo "ref rc $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["roofline"]["traffic"], d["latency"]["p50_us"], d["latency"]["cpu_port_p50_us"], d["latency"]["N50"]["p50_us"], d["latency"]["N50"]["cpu_port_p50_us"])
print([ (r["N"], round(r["ms"],3), round(r["roofline"]["frac"],3)) for r in d["other_configs"]["horizon_sweep"]], d["other_configs"]["config3_arm4"]["ms"])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --headline-only --no-cpu > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc $?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"solve_kernel_t4<mmpc_model_double_pendulum, .int.32" -s 4 -c 1 -f -o gpurun_out/r2_t4_full python bench.py --steps 2 --warmup 1 --headline-only --no-cpu > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc $?"
ls -la gpurun_out/r2_t4_full.ncu-rep
