#!/bin/bash
# same-box A/B of the gradient exchange: N = 1, then N = 8 with 64 MB buckets vs one collective per layer, interleaved, twice
mkdir -p gpurun_out
run() { # name, nproc, env
  env $3 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port 2952$2 bench.py --gpus $2 --steps 40 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/r02_ab_$1.json 2> gpurun_out/r02_ab_$1.err || echo "$1 failed"
  python - "$1" <<'PY'
import json, sys
d = json.loads(open(f'gpurun_out/r02_ab_{sys.argv[1]}.json').read().strip().splitlines()[-1])
print(sys.argv[1], 'N', d['n_gpus'], round(d['value']), 'ms', round(d['ms_per_step'], 3), 'kernel ms', round(d['roofline']['kernel_ms_per_step'], 3), 'replicas', d.get('replicas_identical'), d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
}
run n1 1 "A=1"
for i in 1 2; do
  run n8_bucket64_$i 8 "CONVASR_B200_GRAD_BUCKET_MB=64"
  run n8_perlayer_$i 8 "CONVASR_B200_GRAD_BUCKET_MB=1"
done
run n8_bucket256 8 "CONVASR_B200_GRAD_BUCKET_MB=256"
