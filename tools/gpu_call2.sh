set -x
mkdir -p gpurun_out
for cfg in "0x104 25 3" "0x104 50 7" "0x204 50 20" "0x304 50 40" "0x304 50 64" "0x104 200 4"; do
  timeout 120 python -m tools.t4_check $cfg >> gpurun_out/g2_check.log 2>&1 || echo "FAILED/TIMEOUT $cfg rc=$?" >> gpurun_out/g2_check.log
done
cat gpurun_out/g2_check.log
timeout 300 python -m tools.dev_time 50 3 4096 > gpurun_out/g2_time.log 2>&1
timeout 300 python -m tools.dev_time 50 4 32 512 2048 4096 8192 >> gpurun_out/g2_time.log 2>&1
timeout 300 python -m tools.dev_time 25 4 1 >> gpurun_out/g2_time.log 2>&1
timeout 300 python -m tools.dev_time 25 3 1 >> gpurun_out/g2_time.log 2>&1
timeout 300 python -m tools.dev_time 200 4 4096 >> gpurun_out/g2_time.log 2>&1
cat gpurun_out/g2_time.log
