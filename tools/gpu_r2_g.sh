python -m tools.rebuild_unit builtin_double_pendulum_t4 -DMMPC_T4_PROFILE > /dev/null 2>&1
for v in 0x404 0x504; do timeout 100 python -m tools.t4_prof $v 50 32 2>&1 | tail -4; done
