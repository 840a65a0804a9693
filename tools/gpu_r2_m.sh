mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "rk4 or closed_loop" > gpurun_out/m_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/m_pytest.log
tail -25 gpurun_out/m_pytest.log
MMPC_RK4=1 timeout 300 python -m tools.dev_time 50 0 4096 2>&1 | tail -1
MMPC_RK4=4 timeout 300 python -m tools.dev_time 50 0 4096 2>&1 | tail -1
