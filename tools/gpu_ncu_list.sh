#!/bin/bash
# ncu launch list (durations only) of the default bench command's workload, graphs on
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_bench_train.log 2>&1; echo "ncu-list rc=$?"
wc -l gpurun_out/launches_train.csv; tail -2 gpurun_out/ncu_bench_train.log
python tools/summarize_launches.py gpurun_out/launches_train.csv > gpurun_out/launches_train_summary.md; head -40 gpurun_out/launches_train_summary.md
