mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/a_pytest.log
tail -5 gpurun_out/a_pytest.log
python -m tools.quick_time 50 4096 0 2>&1 | tail -2
python -m tools.quick_time 50 4096 4 2>&1 | tail -2
python -m tools.quick_time 50 4096 3 2>&1 | tail -1
python -m tools.dev_time 50 0 4096 2>&1 | tail -1; python -m tools.dev_time 50 4 32 4096 2>&1 | tail -2; python -m tools.dev_time 50 0x304 4096 2>&1 | tail -1
