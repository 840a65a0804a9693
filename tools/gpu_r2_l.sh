mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/l_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/l_pytest.log
tail -4 gpurun_out/l_pytest.log
timeout 300 python -m tools.dev_time 50 0 32 1184 4096 8192 37888 2>&1 | tail -5
MMPC_RK4=1 timeout 300 python -m tools.dev_time 50 0 4096 2>&1 | tail -1
MMPC_RK4=1 timeout 300 python -m tools.dev_time 50 3 4096 2>&1 | tail -1
MMPC_RK4=4 timeout 300 python -m tools.dev_time 50 0 4096 2>&1 | tail -1
python bench.py --steps 30 > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; tail -c 1500 gpurun_out/l_bench.json
