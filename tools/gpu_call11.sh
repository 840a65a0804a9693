for cfg in "0x304 50 40"; do timeout 60 python -m tools.t4_check $cfg 2>&1 | grep -E "variant|rror" ; done
timeout 100 python -m tools.t4_prof 4 50 4096 2>&1 | tail -8
timeout 100 python -m tools.t4_prof 4 50 32 2>&1 | tail -7
