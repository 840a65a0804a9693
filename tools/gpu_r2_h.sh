for cfg in "0x304 50 40" "0x304 200 33"; do timeout 120 python -m tools.t4_check $cfg 2>&1 | grep -E "variant|rror" || echo "FAILED/TIMEOUT $cfg rc=$?"; done
for v in 0x304; do timeout 300 python -m tools.dev_time 50 $v 4096 8192 2>&1 | tail -2; done
python -m tools.rebuild_unit builtin_double_pendulum_t4 -DMMPC_T4_NO_DISCARD > /dev/null 2>&1
for v in 0x304; do timeout 300 python -m tools.dev_time 50 $v 4096 8192 2>&1 | tail -2; done
