#!/bin/bash
# round 2 final single-GPU visit: test-suite, smoke, the default bench line, a launch list of the training step, full ncu captures
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q -s --timeout 150 > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02_pytest_gpu.log | head -12
timeout 300 python __graft_entry__.py smoke > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
SECONDS=0; timeout 500 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$? wall ${SECONDS}s"; grep -E "Error|error" gpurun_out/r02_bench.err | head -5
timeout 120 python bench.py --impl reference --steps 2 > gpurun_out/r02_bench_reference.json 2>/dev/null; tail -c 300 gpurun_out/r02_bench_reference.json
SECONDS=0
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_train.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-cuda-graphs > gpurun_out/ncu_bench_train.log 2>&1; echo "ncu-list rc=$? ${SECONDS}s"
python tools/summarize_launches.py gpurun_out/r02_launches_train.csv > gpurun_out/r02_launches_train_summary.md 2>&1; head -32 gpurun_out/r02_launches_train_summary.md
# full captures: the BatchNorm stream kernels (18 forward + 18 backward-apply launches per step: a few of each) and dgrad launches with the folded reduction
SECONDS=0
timeout 300 ncu --set full --clock-control none -k regex:'bn_stream' -s 156 -c 12 -f -o /tmp/prof_bn python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-cuda-graphs > gpurun_out/ncu_bn.log 2>&1; echo "ncu-bn rc=$? ${SECONDS}s"
ncu -i /tmp/prof_bn.ncu-rep --page raw --csv > gpurun_out/r02b_bn_stream_ncu_raw.csv 2>/dev/null
python tools/summarize_ncu_raw.py gpurun_out/r02b_bn_stream_ncu_raw.csv MEASURED_PEAKS.json > gpurun_out/r02b_bn_stream_ncu.md 2>&1; cat gpurun_out/r02b_bn_stream_ncu.md | cut -c1-220
# dgrad launches with the folded reduction: decoder dgrad (K = 64), 1x1 conv dgrad, the k = 29 dgrad, one 768-channel dgrad
SECONDS=0
timeout 240 ncu --set full --clock-control none -k regex:'conv1d_umma' -s 204 -c 4 -f -o /tmp/prof_dg python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-cuda-graphs > gpurun_out/ncu_dg.log 2>&1; echo "ncu-dgrad rc=$? ${SECONDS}s"
ncu -i /tmp/prof_dg.ncu-rep --page raw --csv > gpurun_out/r02b_dgrad_fold_ncu_raw.csv 2>/dev/null
python tools/summarize_ncu_raw.py gpurun_out/r02b_dgrad_fold_ncu_raw.csv MEASURED_PEAKS.json > gpurun_out/r02b_dgrad_fold_ncu.md 2>&1; cat gpurun_out/r02b_dgrad_fold_ncu.md | cut -c1-220
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
print('primary', round(d['value']), round(d['ms_per_step'], 3), 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'], 3), round(d['roofline']['kernel_ms_per_step'], 2), d['clocks'])
for k in d['kernels']: print('  ', k['entry'], k['launches_per_step'], round(k['ms_per_step'], 3), k.get('bound'), round(k.get('frac', 0), 3))
for n, e in d['also'].items():
    if 'value' in e: print(n, round(e['value']), round(e['ms_per_step'], 3), 'e2e', e.get('e2e', {}).get('value'))
    else: print(n, {k: v for k, v in e.items() if k != 'what'})
PY
