#!/bin/bash
# full ncu capture of the tensor-pipe kernels of one eager training step (37 conv + 19 wgrad launches);
# the raw page is exported on the box and the (large) report dropped
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none -k regex:'conv1d_umma|wgrad_umma' -s 168 -c 56 -f -o /tmp/prof_train python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-cuda-graphs > gpurun_out/ncu_train.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_train.ncu-rep --page raw --csv > gpurun_out/prof_train_raw.csv 2>/dev/null; ls -la gpurun_out/prof_train_raw.csv
