#!/bin/bash
# full ncu capture of the conv kernel: one steady-state step (19 launches) of the default workload
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/status.log
tail -4 gpurun_out/pytest_gpu.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:conv1d_umma -s 38 -c 19 -f -o gpurun_out/prof_conv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1; echo "ncu-full rc=$?" | tee -a gpurun_out/status.log
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out/
