mkdir -p gpurun_out
for cfg in "0x404 50 40" "0x504 50 40" "0x204 50 64"; do timeout 120 python -m tools.t4_check $cfg 2>&1 | grep -E "variant|rror" || echo "FAILED/TIMEOUT $cfg rc=$?"; done
for v in 3 4 0x304 0x404 0x504; do timeout 300 python -m tools.dev_time 50 $v 4096 2>&1 | tail -1; done
timeout 300 python -m tools.dev_time 50 0x104 32 1184 2>&1 | tail -2
python -m tools.rebuild_unit builtin_double_pendulum_t4 -DMMPC_T4_PROFILE -DMMPC_T3_PROFILE_SWEEPS > /dev/null 2>&1
for v in 4 0x504; do timeout 100 python -m tools.t4_prof $v 50 4096 2>&1 | tail -9; done
