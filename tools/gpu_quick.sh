#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q --timeout 120 -x > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02_pytest_gpu.log | head -12
timeout 120 python tools/time_frontend.py 2>&1 | grep "^\[" | tee gpurun_out/r02_frontend_ab.log
CONVASR_B200_FRONTEND=radix2 timeout 120 python tools/time_frontend.py 2>&1 | grep "^\[" | tee -a gpurun_out/r02_frontend_ab.log
bash tools/gpu_ncu_frontend.sh 2>&1 | tail -8
timeout 200 python bench.py --no-cpu-baseline --no-secondary --steps 20 > gpurun_out/r02_bench_quick.json 2>gpurun_out/r02_bench_quick.err; python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_quick.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['clocks'])
for k in d['kernels']: print(k['entry'], k['launches_per_step'], round(k['ms_per_step'],3), k.get('bound'), round(k.get('frac',0),3))
"
