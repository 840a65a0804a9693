mkdir -p gpurun_out
nvidia-smi -L | wc -l
for g in 2 4 8; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 -m tools.config3_multi gpurun_out/r2_config3_${g}gpu.json > gpurun_out/q_c3_$g.log 2>&1; echo "config3 $g rc $?"; tail -1 gpurun_out/q_c3_$g.log | cut -c1-300
done
python -m tools.config3_multi gpurun_out/r2_config3_1gpu.json 2>&1 | tail -1 | cut -c1-300
python -m tools.group_multi gpurun_out/r2_group_8gpu.json 2>&1 | tail -2 | cut -c1-600
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 50 > gpurun_out/r2_bench_8gpu.json 2> gpurun_out/q_b8.err; echo "bench8 rc $?"; head -c 400 gpurun_out/r2_bench_8gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 50 > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/q_b2.err; echo "bench2 rc $?"; head -c 300 gpurun_out/r2_bench_2gpu.json
