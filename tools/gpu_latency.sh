mkdir -p gpurun_out
timeout 300 python -m tools.latency > gpurun_out/r2_latency.json 2> gpurun_out/latency.err; echo "latency rc $?"
tail -c 3000 gpurun_out/r2_latency.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/lat_pytest.log 2>&1; echo "pytest rc $?"
tail -3 gpurun_out/lat_pytest.log
