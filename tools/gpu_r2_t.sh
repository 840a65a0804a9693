for p in 1 2 3 4 6 8; do echo "pipe $p"; MMPC_PIPE=$p python -m tools.e2e_time 4096 2>&1 | tail -1; done
echo trace; MMPC_TRACE=1 python -m tools.e2e_time 4096 2>&1 | grep "mmpc trace" | tail -4
