mkdir -p gpurun_out; rm -f gpurun_out/g6_*.log
export CUDA_LAUNCH_BLOCKING=1
for cfg in "0x104 25 3 1" "0x104 10 5 1" "0x304 25 33 1" "0x304 25 28 1" "0x304 25 5 1"; do
  echo "=== $cfg" >> gpurun_out/g6_dbg.log
  timeout 60 python -m tools.t4_dbg $cfg 2>&1 | python tools/t4_dbg_summary.py >> gpurun_out/g6_dbg.log
done
cat gpurun_out/g6_dbg.log
