# A/B on ONE box: variant builds of the library under tools/_build (not tracked) against the tree's library
L=$PWD/tools/_build/lib_lean.so
for rep in 1 2; do
echo "tree: device-resident"; timeout 300 python -m tools.dev_time 50 0 4096 2>&1 | tail -1
echo "lean: device-resident"; MMPC_LIB_PATH=$L timeout 300 python -m tools.dev_time 50 0 4096 2>&1 | tail -1
done
for rep in 1 2; do
echo "tree, in-kernel epilogue: e2e"; MMPC_PIPE=8 timeout 300 python -m tools.e2e_time 4096 | tail -1 | cut -c1-120
echo "tree, copy kernel: e2e"; MMPC_ZC_OUT_KERNEL=1 MMPC_PIPE=8 timeout 300 python -m tools.e2e_time 4096 | tail -1 | cut -c1-120
echo "lean, copy kernel: e2e"; MMPC_LIB_PATH=$L MMPC_ZC_OUT_KERNEL=1 MMPC_PIPE=8 timeout 300 python -m tools.e2e_time 4096 | tail -1 | cut -c1-120
done
echo "lean + copy kernel: pinned test"; MMPC_LIB_PATH=$L MMPC_ZC_OUT_KERNEL=1 timeout 600 python -m pytest tests -m gpu -q -k "pinned_host or small_batch_packed" 2>&1 | tail -2
