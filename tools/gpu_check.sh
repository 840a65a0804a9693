mkdir -p gpurun_out
MMPC_SWEEP_ZC=0,2,3 MMPC_SWEEP_PIPE=4,8 timeout 600 python -m tools.e2e_time sweep 4096 2>&1 | grep "^zc="
for m in 2 3; do
MMPC_ZC_BIG=$m timeout 600 python bench.py --headline-only --no-cpu --steps 30 > gpurun_out/check_bench_$m.json 2> gpurun_out/check_bench.err; echo "bench mode $m rc $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/check_bench_$m.json").read().strip().splitlines()[-1])
print("mode $m value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
PY
done
