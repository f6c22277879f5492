#!/bin/bash
# round 2: same-box weak scaling of the headline training step, N = 1 and N = 8 (dynamic vs static tile schedule at N = 8)
mkdir -p gpurun_out
run() { # name, nproc, extra env
  env $3 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2951$2 bench.py --gpus $2 --steps 30 --warmup 5 --no-cpu-baseline $4 > gpurun_out/r02_scale_$1.json 2> gpurun_out/r02_scale_$1.err; echo "$1 rc=$?"
  python - "$1" <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r02_scale_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'N', d['n_gpus'], round(d['value']), 'ms', round(d['ms_per_step'], 3), 'kernel ms', round(d['roofline']['kernel_ms_per_step'], 3), 'replicas', d.get('replicas_identical'), d['clocks'])
for n, e in (d.get('also') or {}).items():
    if 'value' in e: print('   ', n, round(e['value']), round(e['ms_per_step'], 3))
PY
}
run n1 1 "A=1" "--no-secondary"
run n8 8 "A=1" ""
run n8_static 8 "CONVASR_B200_STATIC_TILES=1" "--no-secondary"
