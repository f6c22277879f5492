#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --steps 10 --warmup 3 --no-secondary > gpurun_out/bench_8gpu.log 2> gpurun_out/bench_8gpu.err; echo "bench8 rc=$?"
tail -1 gpurun_out/bench_8gpu.log | cut -c1-300; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_8gpu.err | tail -5
