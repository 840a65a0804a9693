timeout 100 python -m tools.t4_prof 0x104 50 32 2>&1 | grep -E "serial|kernel|ms" | head -6
