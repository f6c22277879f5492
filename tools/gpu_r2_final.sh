#!/bin/bash
# round 2 final single-GPU visit: test-suite, smoke, the default bench line, a launch list of the training step
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -s --timeout 150 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02_pytest_gpu.log | head -12
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
SECONDS=0; timeout 500 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$? wall ${SECONDS}s"; grep -E "Error|error" gpurun_out/r02_bench.err | head -5
timeout 120 python bench.py --impl reference --steps 2 > gpurun_out/r02_bench_reference.json 2>/dev/null; tail -c 600 gpurun_out/r02_bench_reference.json
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_train.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --no-cuda-graphs > gpurun_out/ncu_bench_train.log 2>&1; echo "ncu-list rc=$?"
python tools/summarize_launches.py gpurun_out/r02_launches_train.csv > gpurun_out/r02_launches_train_summary.md 2>&1; head -40 gpurun_out/r02_launches_train_summary.md
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print('primary', round(d['value']), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'], 3), round(d['roofline']['kernel_ms_per_step'], 2), d['clocks'])
for n, e in d['also'].items():
    if 'value' in e: print(n, round(e['value']), round(e['ms_per_step'], 3), 'e2e', e.get('e2e', {}).get('value'))
    else: print(n, {k: v for k, v in e.items() if k != 'what'})
PY
