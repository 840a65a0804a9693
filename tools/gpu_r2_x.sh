for cfg in "0x404 25 3" "0x404 50 21" "0x404 200 9"; do timeout 120 python -m tools.t4_check $cfg 2>&1 | grep -E "variant|rror" || echo "FAILED/TIMEOUT $cfg rc=$?"; done
for v in 0x104 0x404; do timeout 300 python -m tools.dev_time 50 $v 4 32 296 592 1184 2368 2>&1 | tail -6; timeout 300 python -m tools.dev_time 25 $v 1 2>&1 | tail -1; done
for v in 0x104 0x404; do MMPC_RK4=4 timeout 300 python -m tools.dev_time 10 $v 1 2>&1 | tail -1; done
