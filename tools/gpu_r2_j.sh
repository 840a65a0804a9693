mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/j_pytest.log
tail -4 gpurun_out/j_pytest.log
timeout 300 python -m tools.dev_time 50 0 32 1184 4096 8192 37888 2>&1 | tail -5
timeout 300 python -m tools.dev_time 50 3 4096 2>&1 | tail -1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:solve_kernel_t4 -s 1 -c 1 -o gpurun_out/r2_t4_a python -m tools.t4_prof 0 50 4096 > gpurun_out/j_ncu.log 2>&1; tail -3 gpurun_out/j_ncu.log
