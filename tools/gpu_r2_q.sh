for cfg in "0x304 50 40" "0x104 25 3"; do timeout 120 python -m tools.t4_check $cfg 2>&1 | grep -E "variant|rror" || echo "FAILED/TIMEOUT $cfg rc=$?"; done
timeout 300 python -m tools.dev_time 50 0 32 4096 8192 2>&1 | tail -3
python -m tools.rebuild_unit builtin_double_pendulum_t4 -DMMPC_T4_PROFILE > /dev/null 2>&1
timeout 100 python -m tools.t4_prof 0 50 4096 2>&1 | tail -3
