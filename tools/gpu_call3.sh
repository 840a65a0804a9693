mkdir -p gpurun_out; rm -f gpurun_out/g3_*.log
export CUDA_LAUNCH_BLOCKING=1
for cfg in "0x104 50 7" "0x104 25 3" "0x104 25 7" "0x104 50 3" "0x104 10 5" "0x204 50 20" "0x304 50 40" "0x304 50 64" "0x104 200 4" "0x304 200 30" "0x304 25 33"; do
  timeout 40 python -m tools.t4_check $cfg >> gpurun_out/g3_check.log 2>&1 || echo "FAILED/TIMEOUT $cfg rc=$?" >> gpurun_out/g3_check.log
done
grep -v "^  File\|^    \|Traceback" gpurun_out/g3_check.log
unset CUDA_LAUNCH_BLOCKING
timeout 100 python -m tools.dev_time 50 4 32 512 2048 4096 8192 >> gpurun_out/g3_time.log 2>&1
timeout 100 python -m tools.dev_time 25 4 1 >> gpurun_out/g3_time.log 2>&1
timeout 100 python -m tools.dev_time 200 4 4096 >> gpurun_out/g3_time.log 2>&1
grep -E "^N=|rror" gpurun_out/g3_time.log
