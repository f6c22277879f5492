#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/pytest_skip.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_skip.log
for flag in 0 1; do
  CONVASR_B200_SKIP_PADDING=$flag timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_skip$flag.log 2> gpurun_out/bench_skip$flag.err; echo "bench skip=$flag rc=$?"
  tail -1 gpurun_out/bench_skip$flag.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['roofline']['kernel_ms_per_step'], d['clocks']['sm_mhz'], d['also'])"
  tail -3 gpurun_out/bench_skip$flag.err
done
