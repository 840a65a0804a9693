mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/final_bench.err; echo "bench rc $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_1gpu.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["frac"], d["clocks"], d["latency"]["p50_us"], d["latency"]["cpu_port_p50_us"], d["latency"]["N50"]["p50_us"], d["latency"]["N50"]["cpu_port_p50_us"], d["cpu_baseline"]["value"], d["cpu_baseline"]["sample"][:60])
PY
timeout 600 python -m pytest tests -m gpu -q -k "pinned_host or small_batch_packed" 2>&1 | tail -2
