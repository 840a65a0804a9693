mkdir -p gpurun_out
timeout 600 python -m tools.e2e_time sweep 4096 > gpurun_out/r2_e2e_sweep.txt 2>&1; echo "e2e rc $?"
grep "^zc=" gpurun_out/r2_e2e_sweep.txt; grep -o '"identical_results": [a-z]*' gpurun_out/r2_e2e_sweep.txt
timeout 300 python -m tools.latency > gpurun_out/r2_latency.json 2> gpurun_out/latency.err; echo "latency rc $?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_latency.json'))
print("identical", d["identical_results"])
for k in d['zero_copy']:
    z=d['zero_copy'][k]; p=d['packed_copies'].get(k)
    if k.startswith('B'):
        print(k, 'zc p50 %.1f p99 %.1f span %.1f | packed p50 %.1f p99 %.1f kern %.1f | bounded zc %.0f packed %.0f'%(z['p50_us'],z['p99_us'],z['kernel_us'],p['p50_us'],p['p99_us'],p['kernel_us'],z['bounded']['p50_us'],p['bounded']['p50_us']))
    else: print(k,z)
PY
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/lat_pytest.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/lat_pytest.log
