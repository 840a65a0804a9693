mkdir -p gpurun_out; rm -f gpurun_out/g4_*.log
export CUDA_LAUNCH_BLOCKING=1
for cfg in "0x104 25 3 0" "0x104 25 3 1" "0x104 25 3 2" "0x104 10 5 0" "0x104 10 5 1" "0x304 25 33 0" "0x304 25 33 1" "0x304 25 33 2" "0x304 50 40 3" "0x104 200 4 2"; do
  echo "=== $cfg" >> gpurun_out/g4_dbg.log
  timeout 60 python -m tools.t4_dbg $cfg 2>&1 | grep -v "^  File\|^    \|Traceback" | sort | uniq -c | sort -rn | head -12 >> gpurun_out/g4_dbg.log
done
cat gpurun_out/g4_dbg.log
