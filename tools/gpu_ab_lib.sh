# A/B on ONE box: tools/_build/libmahi_mpc_b200_old.so (an earlier build, not tracked; MMPC_LIB_PATH selects it) against the tree's library
for rep in 1 2; do
echo "A (old)"; MMPC_LIB_PATH=$PWD/tools/_build/libmahi_mpc_b200_old.so timeout 300 python -m tools.dev_time 50 0 32 4096 2>&1 | tail -2
echo "B (tree)"; timeout 300 python -m tools.dev_time 50 0 32 4096 2>&1 | tail -2
done
