mkdir -p gpurun_out
bash tools/gpu_ab_lib.sh
echo "--- pytest"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
MMPC_LAT_SIZES=1x25,1x50,32x25,256x25 timeout 300 python -m tools.latency 1 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('LATENCY '):
        d = json.loads(l[8:])
        print({k: (round(v['p50_us'], 1), round(v['kernel_us'], 1)) if k.startswith('B') else v for k, v in d.items()})
"
