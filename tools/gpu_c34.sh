#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/pytest_gc.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gc.log
for w in jasper_separable_fwd_ctc_B256x20s_bf16 wav2letter_bpe5000_fwd_ctc_B64x15s_bf16; do
  timeout 600 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err; echo "bench $w rc=$?"
  tail -1 gpurun_out/bench_$w.log | cut -c1-200; tail -3 gpurun_out/bench_$w.err
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_$w.log 2>&1; echo "ncu $w rc=$?"
  python tools/summarize_launches.py gpurun_out/launches_$w.csv ctc_grad_scatter > gpurun_out/launches_${w}_summary.md; head -14 gpurun_out/launches_${w}_summary.md
done
