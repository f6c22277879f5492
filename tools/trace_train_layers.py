"""Per-call device times of one eager training step of the headline workload (C2: Wav2Letter-char, B = 80 x 15 s, bf16), with
the BatchNorm-backward reduction folded into the dgrad epilogues and without: which layers pay more in the GEMM than the
separate reduce pass would cost?  CUDA events around every C-ABI call (_lib.trace); minimum over a few steps per call index."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from convasr_b200 import _lib, models, optimizers, training
from oracle import oracle as O  # seeded synthetic weights only

dev = torch.device('cuda:0')
name = bench.DEFAULT_WORKLOAD
model_name, C, B, seconds, precision, kind, lengths = bench.WORKLOADS[name]
model = getattr(models, model_name)(64, [C], frontend = models.LogFilterBankFrontend(64, bench.SAMPLE_RATE, .02, .01, 'hann_window'), dropout = 0., check_time_dim_padded = False)
shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if not k.startswith('frontend.')}
model.load_state_dict(O.synth_state_dict(shapes, seed = 0), strict = False)
model = model.to(dev).set_precision(precision).train()
sig, xlen, y, ylen = [t.to(dev) for t in bench.synth_batch(B, seconds, C, seed = 1000, lengths = lengths)]
opt = optimizers.SGD([p for p in model.parameters() if p.requires_grad], lr = 1e-6, momentum = 0.9, weight_decay = 1e-3)
flush = torch.empty(256 << 20, dtype = torch.uint8, device = dev)

def step():
	opt.zero_grad(set_to_none = True)
	out = model(sig, xlen, y = y, ylen = ylen)
	(out['loss'] * ylen[:, 0]).mean().backward()
	opt.step(max_grad_norm = 100.0)

res = {}
for fold in (True, False, True, False):
	training._FOLD_BN_REDUCE = fold
	for _ in range(2):
		step()
	per_call = None
	for _ in range(4):
		flush.zero_()
		with _lib.trace() as tr:
			step()
		torch.cuda.synchronize()
		times = [(n, e0.elapsed_time(e1) * 1e3) for n, e0, e1 in tr.events]
		per_call = times if per_call is None else [(n, min(a, b)) for (n, a), (_, b) in zip(per_call, times)]
	key = 'fold' if fold else 'nofold'
	res[key] = per_call if key not in res else [(n, min(a, b)) for (n, a), (_, b) in zip(res[key], per_call)]

def pick(calls, entry):
	return [t for n, t in calls if n == entry]

conv_f, conv_n = pick(res['fold'], 'cab_conv1d_fused'), pick(res['nofold'], 'cab_conv1d_fused')
bn_f, bn_n = pick(res['fold'], 'cab_bn_act_mask_bwd_apply'), pick(res['nofold'], 'cab_bn_act_mask_bwd')
wg = pick(res['fold'], 'cab_conv1d_wgrad')
print('conv1d_fused calls', len(conv_f), 'bn bwd calls', len(bn_f), len(bn_n))
# backward order: conv call 19 is the decoder dgrad (-> gradient of the last repeat's output), then one dgrad per repeat going down
n_fwd = 19
print('| backward launch (dgrad producing dL/d out of repeat r) | dgrad us, fold | dgrad us, no fold | BN bwd us, apply only | BN bwd us, reduce + apply | fold gain us |')
print('|---|---|---|---|---|---|')
tot = 0.0
rows = []
for i in range(len(bn_f)):
	r = len(bn_f) - 1 - i
	df, dn = conv_f[n_fwd + i], conv_n[n_fwd + i]
	gain = (dn + bn_n[i]) - (df + bn_f[i])
	tot += gain
	rows.append(dict(repeat = r, dgrad_fold = df, dgrad_nofold = dn, bn_apply = bn_f[i], bn_full = bn_n[i], gain = gain))
	print(f'| {r} | {df:.1f} | {dn:.1f} | {bn_f[i]:.1f} | {bn_n[i]:.1f} | {gain:+.1f} |')
print(f'total fold gain {tot:.1f} us; forward conv launches (us):', ' '.join(f'{t:.0f}' for t in conv_f[:n_fwd]))
print('wgrad launches (us):', ' '.join(f'{t:.0f}' for t in wg))
os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok = True)
json.dump(dict(rows = rows, conv_fold = conv_f, conv_nofold = conv_n, wgrad = wg, bn_fwd = pick(res['fold'], 'cab_bn_act_mask_fwd_stats')), open(os.path.join(ROOT, 'gpurun_out', 'r02b_trace_layers.json'), 'w'))
