mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/s_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/s_pytest.log
tail -4 gpurun_out/s_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
