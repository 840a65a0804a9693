for fr in 0.25 0.5 0.75 1.0; do
python -m tools.rebuild_unit builtin_double_pendulum_t4 -DMMPC_T3_ST=2 -DMMPC_T3_ST_FRACTION=$fr > /dev/null 2>&1
echo "evict_last fraction $fr"; for v in 0x304; do timeout 300 python -m tools.dev_time 50 $v 4096 2>&1 | tail -1; done
done
python -m tools.rebuild_unit builtin_double_pendulum_t4 -DMMPC_T3_ST=1 > /dev/null 2>&1
echo "evict first"; for v in 0x304; do timeout 300 python -m tools.dev_time 50 $v 4096 2>&1 | tail -1; done
