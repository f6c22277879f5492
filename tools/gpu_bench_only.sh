#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_only.log 2> gpurun_out/bench_only.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_only.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e'], d['roofline']['kernel_ms_per_step'], d['clocks'], d['also'])"
tail -3 gpurun_out/bench_only.err
