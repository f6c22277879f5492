#!/bin/bash
mkdir -p gpurun_out
for w in jasper_separable_fwd_ctc_B256x20s_bf16 wav2letter_bpe5000_fwd_ctc_B64x15s_bf16 wav2letter_char_fwd_ctc_B8x10s_fp32; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  tail -1 gpurun_out/bench_$w.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'], {k:d[k] for k in ('value','ms_per_step')}, 'e2e', d['e2e']['ms_per_step'], 'roofline', d['roofline']['achieved'], d['roofline']['frac'])"
  tail -2 gpurun_out/bench_$w.err
done
