#!/bin/bash
# round 2: GPU test-suite (tight per-test timeout), frontend A/B, then the default bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -s --timeout 150 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02_pytest_gpu.log | head -20
timeout 120 python tools/time_frontend.py 2>&1 | grep "^\[" | tee gpurun_out/r02_frontend_ab.log
CONVASR_B200_FRONTEND=radix2 timeout 120 python tools/time_frontend.py 2>&1 | grep "^\[" | tee -a gpurun_out/r02_frontend_ab.log
SECONDS=0; timeout 600 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$? wall ${SECONDS}s"; grep -E "Error|error" gpurun_out/r02_bench.err | head
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print('primary', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], d['roofline']['kernel_ms_per_step'], d['clocks'])
for k in d['kernels']:
    print('  ', k['entry'], k['launches_per_step'], round(k['ms_per_step'], 3), k.get('bound'), round(k.get('frac', 0), 3))
print('cpu', d['cpu_baseline'])
for n, e in d['also'].items():
    if 'value' in e:
        print(n, round(e['value']), round(e['ms_per_step'], 3), 'roofline', e['roofline'] and round(e['roofline']['frac'], 3), 'e2e', e.get('e2e', {}).get('value'))
        for k in (e.get('kernels') or [])[:8]:
            print('     ', k['entry'], k['launches_per_step'], round(k['ms_per_step'], 3), k.get('bound'), round(k.get('frac', 0), 3))
    else:
        print(n, e)
PY
