mkdir -p gpurun_out
timeout 100 python -m tools.t4_prof 4 50 4096 > gpurun_out/g9_prof.log 2>&1; cat gpurun_out/g9_prof.log | tail -40
timeout 100 python -m tools.t4_prof 4 50 32 > gpurun_out/g9_prof32.log 2>&1; cat gpurun_out/g9_prof32.log | tail -24
