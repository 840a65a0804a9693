for rep in 1 2; do
for w in dfs creation; do
cp gpurun_ab_$w.so mahi_mpc_b200/lib/libmahi_mpc_b200.so
echo "== $w"; timeout 300 python -m tools.dev_time 50 0 32 4096 2>&1 | tail -2; timeout 300 python -m tools.dev_time 50 3 4096 2>&1 | tail -1; timeout 300 python -m tools.arm4_time 4096 2>&1 | tail -1
done; done
MMPC_UBOUND=1.5 timeout 300 python -m tools.dev_time 50 0 4096 2>&1 | tail -1
