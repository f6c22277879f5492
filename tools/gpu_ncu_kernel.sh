#!/bin/bash
# usage: gpu_ncu_kernel.sh <kernel-regex> <skip> <count> <workload> [extra bench args]: full ncu capture of selected launches,
# raw + details pages exported on the box (the report itself is dropped)
mkdir -p gpurun_out
K=$1; S=$2; C=$3; W=$4; shift 4
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$K" -s $S -c $C -f -o /tmp/prof_k python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-cuda-graphs "$@" > gpurun_out/ncu_kernel.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_k.ncu-rep --page raw --csv > gpurun_out/prof_k_raw.csv 2>/dev/null
ncu -i /tmp/prof_k.ncu-rep --page details > gpurun_out/prof_k_details.txt 2>/dev/null
ncu -i /tmp/prof_k.ncu-rep --page source --csv --print-source sass > gpurun_out/prof_k_source.csv 2>/dev/null
ls -la gpurun_out/prof_k_raw.csv gpurun_out/prof_k_details.txt; tail -3 gpurun_out/ncu_kernel.log | cut -c1-300
