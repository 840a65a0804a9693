for cfg in "0x504 25 3" "0x504 50 20" "0x404 50 40" "0x404 200 33" "0x404 50 4096 1e-5"; do timeout 120 python -m tools.t4_check $cfg 2>&1 | grep -E "variant|rror" || echo "FAILED/TIMEOUT $cfg rc=$?"; done
timeout 300 python -m tools.dev_time 50 0x404 32 4096 8192 2>&1 | tail -3
timeout 300 python -m tools.dev_time 50 0x304 4096 2>&1 | tail -1
timeout 300 python -m tools.dev_time 50 0x504 32 1184 2>&1 | tail -2
timeout 300 python -m tools.dev_time 50 0x104 32 1184 2>&1 | tail -2
timeout 300 python -m tools.dev_time 25 0x504 1 2>&1 | tail -1
timeout 300 python -m tools.dev_time 25 0x104 1 2>&1 | tail -1
