mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/s_pytest.log
tail -6 gpurun_out/s_pytest.log
