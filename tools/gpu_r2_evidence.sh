mkdir -p gpurun_out
python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/ev_bench.err; echo "bench rc $?"
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2_bench_reference_arm.json 2>/dev/null; echo "ref rc $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --headline-only --no-cpu > gpurun_out/ev_ncu_launch.log 2>&1; echo "ncu launches rc $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel_t4 -s 4 -c 1 -o gpurun_out/r2_bench_solve_t4 python bench.py --steps 4 --warmup 3 --headline-only --no-cpu > gpurun_out/ev_ncu_full.log 2>&1; echo "ncu full rc $?"
python -m tools.rebuild_unit builtin_double_pendulum_t4 -DMMPC_T4_PROFILE > /dev/null 2>&1
timeout 100 python -m tools.t4_prof 0 50 4096 2>&1 | tail -4 > gpurun_out/r2_t4_timers_4096.txt; cat gpurun_out/r2_t4_timers_4096.txt
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["latency"]["p50_us"], d["latency"]["cpu_port_p50_us"])
PY
