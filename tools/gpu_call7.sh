mkdir -p gpurun_out; rm -f gpurun_out/g7_*.log
for cfg in "0x104 25 3" "0x104 50 7" "0x104 10 5" "0x204 50 20" "0x304 25 33" "0x304 50 40" "0x104 200 4" "0x304 200 30" "0x304 50 300"; do
  echo "=== $cfg" >> gpurun_out/g7_dbg.log
  timeout 60 python -m tools.t4_check $cfg 2>&1 | python tools/t4_dbg_summary.py | head -12 >> gpurun_out/g7_dbg.log
done
cat gpurun_out/g7_dbg.log
timeout 100 python -m tools.dev_time 50 4 32 512 2048 4096 8192 2>&1 | grep -E "^N=|rror|timeout" | head -20
timeout 100 python -m tools.dev_time 25 4 1 2>&1 | grep -E "^N=|rror|timeout" | head
