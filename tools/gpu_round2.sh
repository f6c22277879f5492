#!/bin/bash
# two-GPU box visit: data-parallel parity test, 2-rank bench (graphed, then eager), 1-rank bench for the ratio
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.log
timeout 600 python -m pytest tests/test_gpu_training.py -m gpu -x -q -k "two_gpus or graphed" > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest2 rc=$?" | tee -a gpurun_out/status2.log
tail -15 gpurun_out/pytest_2gpu.log
for mode in "" "--no-cuda-graphs"; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-secondary $mode > gpurun_out/bench_2gpu$mode.log 2> gpurun_out/bench_2gpu$mode.err; echo "bench2 $mode rc=$?" | tee -a gpurun_out/status2.log
  tail -1 gpurun_out/bench_2gpu$mode.log; tail -3 gpurun_out/bench_2gpu$mode.err
done
timeout 600 python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/bench_1gpu.log 2> gpurun_out/bench_1gpu.err; echo "bench1 rc=$?" | tee -a gpurun_out/status2.log
tail -1 gpurun_out/bench_1gpu.log
