mkdir -p gpurun_out
for cfg in "0x104 25 3" "0x204 50 20" "0x304 50 40" "0x304 200 30"; do timeout 60 python -m tools.t4_check $cfg 2>&1 | grep -E "variant|rror" ; done
timeout 100 python -m tools.t4_prof 4 50 4096 2>&1 | tail -14
timeout 100 python -m tools.t4_prof 4 50 32 2>&1 | tail -12
