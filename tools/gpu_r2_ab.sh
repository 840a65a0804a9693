# A/B on ONE box: $1 = extra define of variant B
for rep in 1 2 3; do
python -m tools.rebuild_unit builtin_double_pendulum_t4 > /dev/null 2>&1
echo "A (tree)"; timeout 300 python -m tools.dev_time 50 0 32 4096 8192 2>&1 | tail -3
python -m tools.rebuild_unit builtin_double_pendulum_t4 $1 > /dev/null 2>&1
echo "B ($1)"; timeout 300 python -m tools.dev_time 50 0 32 4096 8192 2>&1 | tail -3
done
