#!/bin/bash
# usage (under `gpurun --gpus N`): gpu_scale.sh N -- the default bench workload on N GPUs of one box, as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --steps 10 --warmup 3 --no-secondary > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "bench$N rc=$?"
tail -1 gpurun_out/bench_${N}gpu.log | cut -c1-300; grep -v "OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_${N}gpu.err | tail -5
