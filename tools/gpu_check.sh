#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for w in wav2letter_char_fwd_ctc_B8x10s_fp32; do
  timeout 600 python bench.py --workload $w --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  tail -1 gpurun_out/bench_$w.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'], {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['ms_per_step'])"
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_quick.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['clocks'])"
