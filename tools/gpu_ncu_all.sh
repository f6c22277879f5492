#!/bin/bash
# full ncu capture of every kernel of one steady-state step (eager launches, no CUDA graph)
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'absmax|logmel|instnorm|conv1d_umma|ctc_|log_softmax|greedy|topk|entropy|grouped' -s 64 -c 32 -f -o gpurun_out/prof_step python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cuda-graphs > gpurun_out/ncu_step.log 2>&1; echo "ncu rc=$?"
tail -c 400 gpurun_out/ncu_step.log; ls -la gpurun_out/prof_step.ncu-rep
