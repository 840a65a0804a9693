set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/g1_smi.txt
python -m pytest tests -m gpu -x -q > gpurun_out/g1_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/g1_pytest.log
tail -15 gpurun_out/g1_pytest.log
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 30 python tools/sanitize_solve.py > gpurun_out/g1_san_$tool.txt 2>&1; echo "rc $?" >> gpurun_out/g1_san_$tool.txt
  tail -8 gpurun_out/g1_san_$tool.txt
done
python bench.py > gpurun_out/g1_bench.json 2> gpurun_out/g1_bench.err; tail -c 3000 gpurun_out/g1_bench.json
