mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "node_cost" > gpurun_out/r_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r_pytest.log
tail -30 gpurun_out/r_pytest.log
