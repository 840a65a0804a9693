for v in 3 4 0x304 0x404 0x504; do timeout 300 python -m tools.dev_time 50 $v 4096 2>&1 | tail -1; done
timeout 300 python -m tools.dev_time 50 0x104 32 1184 2>&1 | tail -2
timeout 300 python -m tools.dev_time 25 0x104 1 2>&1 | tail -2
timeout 300 python -m tools.quick_time 25 1 0 2>&1 | tail -2
