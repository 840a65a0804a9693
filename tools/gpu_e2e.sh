mkdir -p gpurun_out
for B in 4096 16384; do
echo "B=$B"; MMPC_SWEEP_ZC=0,2,3 MMPC_SWEEP_PIPE=4,8 timeout 600 python -m tools.e2e_time sweep $B > gpurun_out/r2_e2e_sweep_$B.txt 2>&1
grep "^zc=" gpurun_out/r2_e2e_sweep_$B.txt; grep -o '"identical_results": [a-z]*' gpurun_out/r2_e2e_sweep_$B.txt
done
