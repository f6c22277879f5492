#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/time_train_kernels.py 2>&1 | tail -6
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_quick.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['clocks'])"
tail -3 gpurun_out/bench_quick.err
