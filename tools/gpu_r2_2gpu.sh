#!/bin/bash
# round 2: the data-parallel GradSync test on two GPUs (kept log), then the 2-GPU bench line (replicas_identical)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_training.py -m gpu -q -s --timeout 200 -k "two_gpus" > gpurun_out/r02_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; grep -E "rank [01]:|passed|failed|skipped|Error" gpurun_out/r02_pytest_2gpu.log | head
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; echo "bench2 rc=$?"; grep -E "Error" gpurun_out/r02_bench_2gpu.err | head -5
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_2gpu.json').read().strip().splitlines()[-1])
print('N=2', d['value'], d['ms_per_step'], 'replicas', d.get('replicas'), 'kernel ms', d['roofline']['kernel_ms_per_step'], d['clocks'])
for n, e in d['also'].items():
    print(n, round(e['value']), round(e['ms_per_step'], 3), e.get('e2e', {}).get('value'))
PY
