for cfg in "0x304 50 40" "0x304 25 33" "4 50 4096 1e-5"; do timeout 120 python -m tools.t4_check $cfg 2>&1 | grep -E "variant|rror" || echo "FAILED/TIMEOUT $cfg rc=$?"; done
python -m pytest tests -m gpu -x -q -k "full_size or headline or device_pointers" 2>&1 | tail -2
bash tools/gpu_r2_ab.sh -DMMPC_T4_PLAIN_EPILOGUE
