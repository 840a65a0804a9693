mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "horizon_sweep or headline" > gpurun_out/u_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/u_pytest.log
tail -12 gpurun_out/u_pytest.log
