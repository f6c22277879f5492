"""Headline benchmark: audio-seconds per second (RTFx) of the Wav2Letter forward + CTC path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference] [--workload NAME]

Default workload = BASELINE.json configs[1]: Wav2Letter-char (66.5 M params, 38 classes) TRAINING
STEP, batch 80 x 15 s of synthetic 8 kHz int16 PCM, every utterance full length (xlen = 1, as the
reference's own benchmark.py feeds it), bf16.  One step = log-mel
frontend -> instance norm -> 18 conv + batch-statistics BatchNorm + hardtanh + mask layers -> decoder
+ log_softmax -> CTC loss -> backward through everything (CTC gradient, log_softmax, decoder, BN,
dgrad and wgrad of every conv) -> clip_grad_norm(100) + SGD(momentum 0.9, weight decay 1e-3) update (the
reference's train.py defaults, train.py:657-662,776-779) as one native multi-tensor step.  The whole step
runs on this repo's kernels.  N > 1: data-parallel replicas; every layer's weight gradient is
all-reduced (NCCL) from inside the native backward as soon as it exists, inside the same CUDA graph.

Secondary workloads (reported under "also", selectable with --workload): the same training step on a
RAGGED batch (xlen ~ U(0.5, 1], SURVEY.md 8d: masks exercised; tiles of pure padding are structural zeros
and are left out of the GEMMs) and the inference path of the same shape -- eval-mode forward with folded
BatchNorm (CUDA-graph replay) + CTC loss + CTC gradient.

value  = whole-job audio-seconds / second with the PCM already resident in HBM (CUDA events).
e2e    = same metric through the public module API with HOST buffers: pinned int16 PCM + targets
         -> H2D through convasr_b200.feed.DeviceFeeder (every step uploads its own batch; the copy of step
         i+1 overlaps the compute of step i), the step, D2H of the per-utterance loss (and the greedy
         ids for inference).
N > 1  = per-GPU batch fixed (weak scaling); time = max over ranks.
--impl reference = the CPU oracle port of the reference path (torch CPU ops, all host threads) on a
         bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
	sys.path.insert(0, ROOT)

import torch

WORKLOADS = {
	# name: (model, num_classes, batch, seconds, precision, kind, lengths)
	#   lengths 'full'  : every utterance fills its row, xlen = 1 (the reference's own benchmark.py:126 style; BASELINE "80 x 15 s")
	#   lengths 'ragged': xlen ~ U(0.5, 1] (SURVEY.md 8d "recommended additionally": exercises masks); on average 25 % of every row
	#                     is padding, whose tiles the conv / wgrad kernels leave out (exact: they are structural zeros)
	'wav2letter_char_train_step_B80x15s_bf16': ('Wav2Letter', 38, 80, 15.0, 'bf16', 'train', 'full'),
	'wav2letter_char_train_step_B80x15s_bf16_ragged': ('Wav2Letter', 38, 80, 15.0, 'bf16', 'train', 'ragged'),
	'wav2letter_char_fwd_ctc_B80x15s_bf16': ('Wav2Letter', 38, 80, 15.0, 'bf16', 'infer', 'full'),
	'wav2letter_char_fwd_ctc_B80x15s_bf16_ragged': ('Wav2Letter', 38, 80, 15.0, 'bf16', 'infer', 'ragged'),
	'wav2letter_char_fwd_ctc_B8x10s_fp32': ('Wav2Letter', 38, 8, 10.0, 'fp32', 'infer', 'ragged'),
	'jasper_separable_fwd_ctc_B256x20s_bf16': ('JasperNetSeparable', 38, 256, 20.0, 'bf16', 'infer', 'full'),
	'wav2letter_bpe5000_fwd_ctc_B64x15s_bf16': ('Wav2Letter', 5000, 64, 15.0, 'bf16', 'infer', 'full'),
}
DEFAULT_WORKLOAD = 'wav2letter_char_train_step_B80x15s_bf16'
SECONDARY_WORKLOADS = ['wav2letter_char_train_step_B80x15s_bf16_ragged', 'wav2letter_char_fwd_ctc_B80x15s_bf16']
# mean DRAM bytes per launch of the tensor-pipe kernels, from the committed ncu --set full captures
NCU_DRAM_BYTES_PER_LAUNCH = {
	'wav2letter_char_train_step_B80x15s_bf16': 131.2e6,  # profiles/r01_train_step_tensor_kernels_ncu_full.csv (56 launches)
	'wav2letter_char_train_step_B80x15s_bf16_ragged': 131.2e6,
	'wav2letter_char_fwd_ctc_B80x15s_bf16': 105.0e6,  # profiles/r01_step_kernels_ncu_full.csv (conv1d_umma_kernel launches)
}
STEP_DESC = {
	'train': 'frontend+instnorm+18x(conv, batch-stat BN, hardtanh, mask)+decoder/log_softmax+CTC loss+full backward (CTC grad, BN bwd, dgrad, wgrad)+clip_grad_norm+SGD(momentum,wd) update',
	'infer': 'frontend+instnorm+conv stack (BN folded)+decoder/log_softmax/argmax+CTC loss+CTC grad (no conv backward)',
}
SAMPLE_RATE = 8000
LENGTHS_DESC = {
	'full': 'xlen = 1: every utterance fills its row (benchmark.py:126 style)',
	'ragged': 'xlen ~ U(0.5, 1]: ~25 % of each row is padding; value counts PADDED seconds; tiles of pure padding are left out of the GEMMs (structural zeros, exact)',
}


def synth_batch(B, seconds, C, seed, lengths = 'ragged'):
	"""int16 PCM round(3000*N(0,1)); targets never the blank, lengths such that an alignment exists.
	'ragged': xlen ~ U(0.5, 1] with one full-length row; 'full': xlen = 1 (benchmark.py:126), target lengths up to
	0.3 * t_out (L <~ 225 of t = 753)."""
	g = torch.Generator().manual_seed(seed)
	T = int(seconds * SAMPLE_RATE)
	sig = (torch.randn(B, T, generator = g) * 3000).round().clamp(-32767, 32767).to(torch.int16)
	xlen = torch.rand(B, generator = g) * 0.5 + 0.5
	xlen[0] = 1.0
	t_out = (T // 80 + 1 - 1) // 2 + 1 + 2
	if lengths == 'full':
		xlen = torch.ones(B)
		L = int(0.3 * t_out)
		lo = max(1, L // 3)
	else:
		t_min = int((xlen.min() * t_out).ceil())
		L = int(0.45 * t_min)
		lo = max(1, int(0.2 * t_min / 2))
	ylen = torch.randint(lo, L + 1, (B, ), generator = g)
	y = torch.randint(0, C - 1, (B, 1, L), generator = g)
	return sig, xlen, y, ylen.unsqueeze(1)


def model_shapes(model_name, C):
	from convasr_b200 import models
	m = getattr(models, model_name)(64, [C], dropout = 0.)
	return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def conv_flops_per_step(model, B, F):
	"""algorithmic 2*MAC of every conv (real channels, real frames, no padding), SURVEY.md 8(d)"""
	import torch.nn as nn
	total = 0
	t = F
	for block in model.backbone:
		for seq in block.conv:
			for conv in seq:
				if isinstance(conv, nn.Conv1d):
					t = (t + 2 * conv.padding[0] - conv.dilation[0] * (conv.kernel_size[0] - 1) - 1) // conv.stride[0] + 1
					total += 2 * B * t * conv.out_channels * (conv.in_channels // conv.groups) * conv.kernel_size[0]
		for rc in block.conv_residual:
			if isinstance(rc, nn.Conv1d):
				total += 2 * B * t * rc.out_channels * rc.in_channels
	d = model.decoder[0]
	total += 2 * B * t * d.out_channels * d.in_channels
	return total, t


class ClockSampler:
	"""nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)"""

	def __init__(self, index):
		self.index, self.rows, self.proc = index, [], None

	def start(self):
		q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
		try:
			self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '100'], stdout = subprocess.PIPE, stderr = subprocess.DEVNULL, text = True)
			threading.Thread(target = self._read, daemon = True).start()
		except Exception:
			self.proc = None

	def _read(self):
		for line in self.proc.stdout:
			self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

	def stop(self, t0, t1):
		if self.proc is None:
			return None
		time.sleep(0.15)
		self.proc.terminate()
		rows = [r for ts, r in self.rows if t0 <= ts <= t1 + 0.2] or [r for _, r in self.rows[-3:]]
		if not rows:
			return None
		try:
			sm = sorted(float(r[0]) for r in rows)
			reasons = set()
			for r in rows:
				for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], r[3:7]):
					if v.lower().startswith('active'):
						reasons.add(name)
			return dict(sm_mhz = sm[len(sm) // 2], sm_max_mhz = float(rows[0][1]), power_w_max = max(float(r[2]) for r in rows), reasons = sorted(reasons), samples = len(rows))
		except Exception:
			return None


def measured_peaks():
	try:
		return json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
	except Exception:
		return None


# --------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference path (oracle/ is the checker AND the CPU baseline)
# --------------------------------------------------------------------------------------------
def cpu_reference_step(sd, model_name, sig, xlen, y, ylen, C, kind):
	from oracle import oracle as O
	if kind == 'train':
		leaf = {k: (v.detach().clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in sd.items()}
		logits, log_probs, olen = O.model_forward(leaf, sig, xlen, model = model_name, training = True)
		loss = O.ctc_loss_torch(log_probs[0].permute(2, 0, 1), y[:, 0], olen[0], ylen[:, 0], C - 1).mean()
		loss.backward()  # backward through the whole stack, as train.py:770-774
		with torch.no_grad():
			for k, v in leaf.items():
				if v.is_floating_point() and v.grad is not None:
					sd[k] = sd[k] - 1e-6 * v.grad  # plain SGD update
		return loss
	logits, log_probs, olen = O.model_forward(sd, sig, xlen, model = model_name)
	lp = log_probs[0].permute(2, 0, 1).detach().requires_grad_(True)
	loss = O.ctc_loss_torch(lp, y[:, 0], olen[0], ylen[:, 0], C - 1)
	loss.sum().backward()  # CTC gradient w.r.t. the log-probs: the same work the native step does
	return loss


def run_cpu_baseline(model_name, C, seconds, sample_B, steps, warmup, kind, lengths = 'ragged'):
	from oracle import oracle as O
	torch.set_num_threads(os.cpu_count())
	shapes = {k: s for k, s in model_shapes(model_name, C).items() if not k.startswith('frontend.')}
	sd = O.synth_state_dict(shapes, seed = 0)
	sig, xlen, y, ylen = synth_batch(sample_B, seconds, C, seed = 0, lengths = lengths)
	for _ in range(warmup):
		cpu_reference_step(sd, model_name, sig, xlen, y, ylen, C, kind)
	t0 = time.perf_counter()
	for _ in range(steps):
		cpu_reference_step(sd, model_name, sig, xlen, y, ylen, C, kind)
	dt = time.perf_counter() - t0
	return sample_B * seconds * steps / dt, dt / steps


def measure(name, args, rank, world, local_rank, dev, steps, warmup, with_cpu_baseline, full):
	"""one workload on this rank's GPU; `full` adds roofline / e2e / clocks (primary line)"""
	from convasr_b200 import _lib, models, ops
	from oracle import oracle as O  # only for the seeded synthetic weights + the cpu_baseline leg
	model_name, C, B, seconds, precision, kind, lengths = WORKLOADS[name]
	config = dict(workload = name, lengths = LENGTHS_DESC[lengths], model = model_name, num_classes = C, batch_per_gpu = B, seconds_per_utterance = seconds, sample_rate = SAMPLE_RATE, precision = precision, kind = kind,
				step = STEP_DESC[kind], parallelism = (f'data-parallel replicas x{world} (per-layer NCCL gradient all-reduce overlapped with the backward)' if kind == 'train' else f'utterance-sharded replicas x{world}'),
				l2 = 'flushed between timed steps (256 MiB memset)', cuda_graphs = not args.no_cuda_graphs)
	cpu_baseline = None
	if with_cpu_baseline:
		sB = args.cpu_sample_batch if kind == 'infer' else max(2, args.cpu_sample_batch // 2)
		v, sec = run_cpu_baseline(model_name, C, seconds, sB, 3, 1, kind, lengths)
		cpu_baseline = dict(value = v, unit = 'audio-s/s', cores = torch.get_num_threads(), kind = 'port', sample = f'{sB} x {seconds:g} s utterances per step, 3 steps after 1 warm-up ({sec:.2f} s/step; oracle port: torch {torch.__version__} CPU ops, {kind} step)')

	frontend = models.LogFilterBankFrontend(64, SAMPLE_RATE, .02, .01, 'hann_window')
	model = getattr(models, model_name)(64, [C], frontend = frontend, dropout = 0., check_time_dim_padded = False)
	shapes = {k: tuple(v.shape) for k, v in model.state_dict().items() if not k.startswith('frontend.')}
	model.load_state_dict(O.synth_state_dict(shapes, seed = 0), strict = False)
	model = model.to(dev)
	sig, xlen, y, ylen = synth_batch(B, seconds, C, seed = 1000 + rank, lengths = lengths)
	sig_pin, xlen_pin, y_pin, ylen_pin = [t.pin_memory() for t in (sig, xlen, y, ylen)]
	sig_d, xlen_d, y_d, ylen_d = [t.to(dev) for t in (sig, xlen, y, ylen)]
	flush = torch.empty(256 << 20, dtype = torch.uint8, device = dev)
	# e2e leg: the public batch feed (convasr_b200/feed.py) -- every step uploads its own pinned host batch; the copy of
	# step i+1 is issued on a copy stream while step i computes (the timed region still contains one full H2D per step)
	from convasr_b200 import feed as feed_mod
	feeder = iter(feed_mod.DeviceFeeder(((None, None, sig_pin, xlen_pin, y_pin, ylen_pin) for _ in range(steps + 16)), dev))
	Fr = sig.shape[1] // 80 + 1
	flops_fwd, t_out = conv_flops_per_step(model, B, Fr)
	first = model.backbone[0].conv[0][0]
	flops_first = 2 * B * ((Fr + 2 * first.padding[0] - first.kernel_size[0]) // first.stride[0] + 1) * first.out_channels * first.in_channels * first.kernel_size[0]

	if kind == 'train':
		from convasr_b200 import training
		model.train()
		assert training.supported(model), 'native training path does not cover this topology'
		net = model
		if world > 1:
			net, _ = models.distributed_data_parallel_and_autocast(model, local_rank)
		# train.py defaults: SGD(momentum 0.9, weight_decay 1e-3) after clip_grad_norm_(max_norm 100) (train.py:657-662,776-779)
		from convasr_b200 import optimizers
		train_params = [p for p in model.parameters() if p.requires_grad]
		optimizer = optimizers.SGD(train_params, lr = 1e-6, momentum = 0.9, weight_decay = 1e-3)  # native multi-tensor step
		flops = 3 * flops_fwd - flops_first  # forward + dgrad (all but the first layer) + wgrad

		def run_eager(s, xl, yy, yl):
			optimizer.zero_grad(set_to_none = True)
			out = net(s, xl, y = yy, ylen = yl)
			loss = (out['loss'] * yl[:, 0]).mean()  # train.py:754-755
			loss.backward()
			optimizer.step(max_grad_norm = 100.0)  # clip_grad_norm_ folded into the native step
			return out['loss']

		run = [run_eager]  # swapped for the CUDA-graph replay after the eager launch count

		def step_device():
			return run[0](sig_d, xlen_d, y_d, ylen_d)

		def step_e2e():
			_, _, s, xl, yy, yl = next(feeder)  # this step's H2D was issued on the copy stream during the previous step
			per_utt = run[0](s, xl, yy, yl)
			return per_utt.detach().cpu()

		d2h = B * 4
		traced_names = ['conv1d_fused', 'conv1d_wgrad']
		kernel_label = 'conv1d_umma_kernel (forward + dgrad) + wgrad_umma_kernel'
	else:
		model.eval().set_precision(precision)
		flops = flops_fwd

		def step_device():
			# grad w.r.t. the log-probs/logits only: CTC alpha || beta -> gradient scatter
			with torch.no_grad():
				out = model(sig_d, xlen_d)
			lp = out['log_probs'][0].requires_grad_(True)
			nll = ops.ctc_loss(lp.permute(2, 0, 1), y_d[:, 0], out['olen'][0], ylen_d[:, 0], blank = C - 1)
			nll.sum().backward()
			return nll

		def step_e2e():
			_, _, s, xl, yy, yl = next(feeder)
			out = model(s, xl, y = yy, ylen = yl)
			return out['loss'].cpu(), out['log_probs'][0]._convasr_argmax.cpu()

		d2h = B * 4 + B * t_out * 4
		traced_names = ['conv1d_fused']
		kernel_label = 'conv1d_umma_kernel'

	def barrier():
		if world > 1:
			torch.distributed.barrier()
		torch.cuda.synchronize()

	def timed(fn, n):
		evs = []
		for _ in range(n):
			flush.zero_()
			e0, e1 = torch.cuda.Event(enable_timing = True), torch.cuda.Event(enable_timing = True)
			e0.record()
			fn()
			e1.record()
			evs.append((e0, e1))
		torch.cuda.synchronize()
		return sum(a.elapsed_time(b) for a, b in evs)  # ms

	# one eager step counts this repo's kernel launches per step (graph replays bypass the counter)
	step_device()
	torch.cuda.synchronize()
	l0 = _lib.launch_count()
	step_device()
	torch.cuda.synchronize()
	launches_per_step = _lib.launch_count() - l0
	if config['cuda_graphs'] and kind == 'infer':
		model.enable_cuda_graphs(True)  # forward = one graph replay; CTC loss/grad stay eager launches
	if config['cuda_graphs'] and kind == 'train':
		try:
			run[0] = training.GraphedTrainStep(net, optimizer, sig_d, xlen_d, y_d, ylen_d, max_grad_norm = 100.0)  # whole step = one replay
		except Exception as e:  # data-parallel capture includes the NCCL collectives; keep the eager step if a stack refuses it
			if world == 1:
				raise
			print(f'[bench] rank {rank}: CUDA-graph capture of the data-parallel step failed ({e!r}); eager step instead', file = sys.stderr)
			config['cuda_graphs'] = False
	for _ in range(warmup):
		nll = step_device()
	torch.cuda.synchronize()
	assert bool(torch.isfinite(nll).all()), 'synthetic targets must admit an alignment'
	sampler = ClockSampler(local_rank)
	if rank == 0:
		sampler.start()
	barrier()
	t_wall0 = time.time()
	ms = timed(step_device, steps)
	barrier()
	t_wall1 = time.time()
	clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
	result = dict(ms = ms, ms_e2e = None)
	roofline = None
	if full:
		# dominant (tensor-pipe) kernels: per-launch CUDA events on the launching stream
		events = []
		originals = {n: getattr(ops, n) for n in traced_names}

		def make_traced(fn):
			def traced(*a, **k):
				e0, e1 = torch.cuda.Event(enable_timing = True), torch.cuda.Event(enable_timing = True)
				e0.record()
				r = fn(*a, **k)
				e1.record()
				events.append((e0, e1))
				return r
			return traced

		for n_, fn in originals.items():
			setattr(ops, n_, make_traced(fn))
		model.enable_cuda_graphs(False)  # per-launch events need the eager launch path
		graphed = run[0] if kind == 'train' else None
		if kind == 'train':
			run[0] = run_eager
		n_prof = min(steps, 5)
		for _ in range(n_prof):
			flush.zero_()
			step_device()
		torch.cuda.synchronize()
		for n_, fn in originals.items():
			setattr(ops, n_, fn)
		if config['cuda_graphs'] and kind == 'infer':
			model.enable_cuda_graphs(True)
		if kind == 'train':
			run[0] = graphed
		kern_ms = sum(a.elapsed_time(b) for a, b in events) / n_prof
		n_kern = len(events) // n_prof
		# end to end through the public API with host buffers
		for _ in range(3):
			step_e2e()
		barrier()
		result['ms_e2e'] = timed(step_e2e, steps)
		barrier()
		peaks = measured_peaks()
		peak = peaks['bf16_tflops_sustained'] if peaks else 1590.0
		achieved = flops / (kern_ms / 1e3) / 1e12
		roofline = dict(
			bound = 'tensor', kernel = kernel_label, achieved = achieved, peak = peak, unit = 'TFLOP/s', frac = achieved / peak,
			peak_source = 'MEASURED_PEAKS.json bf16_tflops_sustained (of measured)' if peaks else 'fallback 1.59 PFLOP/s (of fallback)',
			peak_burst = peaks['bf16_tflops'] if peaks else None, frac_of_burst = (achieved / peaks['bf16_tflops']) if peaks else None, traffic = NCU_DRAM_BYTES_PER_LAUNCH.get(name), traffic_source = 'ncu --set full dram__bytes_read.sum + dram__bytes_write.sum, mean over the launches of one step (profiles/r01_*_ncu_full.csv); the kernels are L2-fed, not HBM-fed',
			launches_per_step = n_kern, kernel_ms_per_step = kern_ms,
			algorithmic_gflop_per_step = flops / 1e9, mma_passes_per_flop = 3 if precision == 'fp32' else 1, share_of_step = kern_ms / (ms / steps)
		)
	result.update(valid_fraction = float(xlen.mean()), config = config, cpu_baseline = cpu_baseline, clocks = clocks, roofline = roofline, launches = launches_per_step * steps, B = B, seconds = seconds, precision = precision,
					h2d = sum(t.numel() * t.element_size() for t in (sig, xlen, y, ylen)), d2h = d2h)
	del model, flush
	torch.cuda.empty_cache()
	return result


def main():
	ap = argparse.ArgumentParser()
	ap.add_argument('--gpus', type = int, default = 1)
	ap.add_argument('--steps', type = int, default = 30)
	ap.add_argument('--warmup', type = int, default = 5)
	ap.add_argument('--impl', default = 'native', choices = ['native', 'reference'])
	ap.add_argument('--workload', default = DEFAULT_WORKLOAD, choices = sorted(WORKLOADS))
	ap.add_argument('--no-cpu-baseline', action = 'store_true')
	ap.add_argument('--no-secondary', action = 'store_true')
	ap.add_argument('--cpu-sample-batch', type = int, default = 8)
	ap.add_argument('--no-cuda-graphs', action = 'store_true')
	args = ap.parse_args()
	model_name, C, B, seconds, precision, kind, lengths = WORKLOADS[args.workload]
	rank = int(os.environ.get('RANK', 0))
	world = int(os.environ.get('WORLD_SIZE', 1))
	local_rank = int(os.environ.get('LOCAL_RANK', 0))
	steps, warmup = args.steps, max(args.warmup, 3)

	if args.impl == 'reference':
		if rank != 0:
			return
		sB = args.cpu_sample_batch if kind == 'infer' else max(2, args.cpu_sample_batch // 2)
		n_steps = max(1, min(steps, 3))
		value, sec = run_cpu_baseline(model_name, C, seconds, sB, n_steps, 1, kind, lengths)
		config = dict(workload = args.workload, lengths = LENGTHS_DESC[lengths], model = model_name, num_classes = C, batch_per_gpu = B, seconds_per_utterance = seconds, sample_rate = SAMPLE_RATE, precision = 'fp32', kind = kind, step = STEP_DESC[kind])
		line = dict(
			impl = 'reference', metric = 'audio_seconds_per_second', value = value, unit = 'audio-s/s', n_gpus = args.gpus, steps = n_steps, warmup = 1, ms_per_step = sec * 1e3,
			higher_is_better = True, scaling = 'weak', vs_baseline = None, dtype = 'fp32', data = 'synthetic', config = config,
			cpu_baseline = dict(value = value, unit = 'audio-s/s', cores = torch.get_num_threads(), kind = 'port', sample = f'{sB} x {seconds:g} s utterances per step, {n_steps} steps (oracle port: torch {torch.__version__} CPU ops, {kind} step)'),
			e2e = dict(value = value, unit = 'audio-s/s', h2d_bytes_per_step = 0, d2h_bytes_per_step = 0), gpu_launches = 0
		)
		print(json.dumps(line))
		return

	# ------------------------------------------------------------------ native arm
	assert torch.cuda.is_available(), 'bench.py --impl native needs a CUDA device'
	torch.cuda.set_device(local_rank)
	dev = torch.device('cuda', local_rank)
	if world > 1:
		import torch.distributed as dist
		dist.init_process_group('nccl', device_id = dev)

	def reduce_max(vals):
		if world == 1:
			return vals
		t = torch.tensor(vals, device = dev, dtype = torch.float64)
		torch.distributed.all_reduce(t, op = torch.distributed.ReduceOp.MAX)
		return t.tolist()

	r = measure(args.workload, args, rank, world, local_rank, dev, steps, warmup, with_cpu_baseline = rank == 0 and world == 1 and not args.no_cpu_baseline, full = True)
	ms, ms_e2e = reduce_max([r['ms'], r['ms_e2e']])
	also = None
	if not args.no_secondary and args.workload == DEFAULT_WORKLOAD and world == 1:  # the scaling runs (N > 1) measure the headline workload only
		also = {}
		for name2 in SECONDARY_WORKLOADS:
			r2 = measure(name2, args, rank, world, local_rank, dev, steps, warmup, with_cpu_baseline = False, full = False)
			ms2, = reduce_max([r2['ms']])
			also[name2] = dict(value = r2['B'] * r2['seconds'] * world * steps / (ms2 / 1e3), unit = 'audio-s/s', ms_per_step = ms2 / steps, lengths = r2['config']['lengths'], valid_audio_fraction = r2['valid_fraction'], step = r2['config']['step'], cuda_graphs = r2['config']['cuda_graphs'], gpu_launches = r2['launches'], clocks = r2['clocks'])
	if rank == 0:
		audio_s = r['B'] * r['seconds'] * world
		line = dict(
			metric = 'audio_seconds_per_second', value = audio_s * steps / (ms / 1e3), unit = 'audio-s/s', n_gpus = world, steps = steps, warmup = warmup, ms_per_step = ms / steps,
			higher_is_better = True, scaling = 'weak', vs_baseline = None, dtype = 'bf16' if r['precision'] == 'bf16' else 'bf16x3 (split-bf16, fp32 accumulate)', data = 'synthetic',
			config = r['config'],
			e2e = dict(value = audio_s * steps / (ms_e2e / 1e3), unit = 'audio-s/s', ms_per_step = ms_e2e / steps, h2d_bytes_per_step = r['h2d'], d2h_bytes_per_step = r['d2h']),
			gpu_launches = r['launches'], roofline = r['roofline'], cpu_baseline = r['cpu_baseline'], clocks = r['clocks'], also = also
		)
		print(json.dumps(line))
	if world > 1:
		torch.distributed.destroy_process_group()


if __name__ == '__main__':
	main()
