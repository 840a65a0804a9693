python -m tools.rebuild_unit builtin_double_pendulum_t4 -DMMPC_T4_PROFILE -DMMPC_T3_PROFILE_SWEEPS > /dev/null 2>&1
for v in 0x304 0x404; do echo "== $v"; timeout 100 python -m tools.t4_prof $v 50 4096 2>&1 | tail -8; done
python -m tools.rebuild_unit builtin_double_pendulum_t4 -DMMPC_T4_PROFILE -DMMPC_T3_PROFILE_SWEEPS -DMMPC_T3_ST=3 > /dev/null 2>&1
for v in 0x304 0x404; do echo "== $v no record stores"; timeout 100 python -m tools.t4_prof $v 50 4096 2>&1 | tail -8; done
for v in 0x304; do echo "== $v no record stores B=32"; timeout 100 python -m tools.t4_prof $v 50 32 2>&1 | tail -8; done
