mkdir -p gpurun_out; rm -f gpurun_out/g8_*.log
for cfg in "0x104 25 3" "0x104 50 7" "0x104 10 5" "0x204 50 20" "0x304 25 33" "0x304 50 40" "0x104 200 4" "0x304 200 30" "0x304 50 300" "0x204 100 100" "0x104 1 3"; do
  echo "=== $cfg" >> gpurun_out/g8_dbg.log
  timeout 60 python -m tools.t4_check $cfg > gpurun_out/g8_one.log 2>&1; python tools/t4_dbg_summary.py < gpurun_out/g8_one.log | head -30 >> gpurun_out/g8_dbg.log
done
cat gpurun_out/g8_dbg.log
timeout 100 python -m tools.dev_time 50 4 32 512 2048 4096 8192 > gpurun_out/g8_t.log 2>&1; grep -E "^N=|rror" gpurun_out/g8_t.log | head -20
timeout 100 python -m tools.dev_time 25 4 1 > gpurun_out/g8_t2.log 2>&1;  grep -E "^N=|rror" gpurun_out/g8_t2.log | head
