#!/bin/bash
# configs C3 / C4 at full size: throughput + ncu launch list
mkdir -p gpurun_out
for w in jasper_separable_fwd_ctc_B256x20s_bf16 wav2letter_bpe5000_fwd_ctc_B64x15s_bf16; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  tail -1 gpurun_out/bench_$w.log | cut -c1-400; tail -3 gpurun_out/bench_$w.err
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_$w.log 2>&1; echo "ncu $w rc=$?"
  python tools/summarize_launches.py gpurun_out/launches_$w.csv ctc_grad_scatter > gpurun_out/launches_${w}_summary.md; head -30 gpurun_out/launches_${w}_summary.md
done
