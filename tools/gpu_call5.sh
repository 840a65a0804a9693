mkdir -p gpurun_out; rm -f gpurun_out/g5_*.log
export CUDA_LAUNCH_BLOCKING=1
for cfg in "0x104 25 3" "0x104 50 7" "0x104 10 5" "0x204 50 20" "0x304 25 33" "0x304 50 40" "0x104 200 4" "0x304 200 30"; do
  echo "=== $cfg" >> gpurun_out/g5_dbg.log
  timeout 60 python -m tools.t4_check $cfg 2>&1 | grep -v "^  File\|^    \|Traceback" | head -8 >> gpurun_out/g5_dbg.log
done
cat gpurun_out/g5_dbg.log
