#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 4 --steps 10 --warmup 3 --no-secondary > gpurun_out/bench_4gpu.log 2> gpurun_out/bench_4gpu.err; echo "bench4 rc=$?"
tail -1 gpurun_out/bench_4gpu.log | cut -c1-300; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_4gpu.err | tail -5
