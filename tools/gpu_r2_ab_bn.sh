#!/bin/bash
# round 2: GPU tests + same-box A/B of the BatchNorm-path changes (templated stream kernels are always on;
# folded backward reduction, L2-aware chunk order, column-parallel forward statistics are switchable)
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -s --timeout 150 > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|^FAILED|^E  " gpurun_out/r02b_pytest_gpu.log | head -20
run() {
	name=$1; shift
	env "$@" timeout 200 python bench.py --no-cpu-baseline --no-secondary --steps 20 > gpurun_out/r02b_ab_$name.json 2> gpurun_out/r02b_ab_$name.err
	python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
	d = json.loads(open(f'gpurun_out/r02b_ab_{name}.json').read().strip().splitlines()[-1])
	k = {e['entry']: round(e['ms_per_step'], 3) for e in d['kernels']}
	print(name, 'ms/step', round(d['ms_per_step'], 3), 'tensor', round(d['roofline']['kernel_ms_per_step'], 2), 'clk', d['clocks']['sm_mhz'],
		'| conv', k.get('cab_conv1d_fused'), 'wgrad', k.get('cab_conv1d_wgrad'), 'bn_bwd', k.get('cab_bn_act_mask_bwd'), 'bn_bwd_apply', k.get('cab_bn_act_mask_bwd_apply'), 'bn_fwd', k.get('cab_bn_act_mask_fwd_stats'))
except Exception as e:
	print(name, 'FAILED', e); print(open(f'gpurun_out/r02b_ab_{name}.err').read()[-1500:])
PY
}
run all_on A=1
run fold_off CONVASR_B200_FOLD_BN_REDUCE=0
run order_off CONVASR_B200_BN_ORDER=0
run statscols_off CONVASR_B200_STATS_COLUMNS=0
run all_off CONVASR_B200_FOLD_BN_REDUCE=0 CONVASR_B200_BN_ORDER=0 CONVASR_B200_STATS_COLUMNS=0
run all_on_2 A=1
