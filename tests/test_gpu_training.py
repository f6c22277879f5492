"""Native training step (conv/BN forward with batch statistics + dgrad/wgrad/BN backward on this
repo's kernels) against CPU fp32 references.

Precision note.  Two tiers (model.set_precision): 'fp32' keeps every activation, gradient activation and operand
copy as a split-bf16 (hi, lo) pair with three accumulated MMA passes -- the exactness harness, pinned against the
reference's own TRAIN-mode outputs (tests/golden/train.pt) at 1e-3; 'bf16' is the speed tier: a randomly initialised
Wav2Letter in TRAINING mode amplifies a perturbation ~1.2x per conv layer (batch-statistics BatchNorm renormalises
every layer), so its end-to-end error is reported as a measured number next to the fp32-tier one.  Independently of
precision, every nonlinearity of the reference has a kink and a gate that flips between two implementations moves the
gradient by O(1 / sqrt(elements)) -- the reference vs its own fp32 CPU restatement differ by 1.7e-2 at full depth for
that reason alone (tests/test_oracle_golden.py); full-depth exactness is therefore asserted on the smooth-regime case
(oracle.smooth_regime) and on the shallow / residual / dense cases where no gate sits within rounding of a kink.
"""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import oracle as O

BF16 = torch.bfloat16


def rel(a, b):
	a, b = a.detach().double().cpu(), b.detach().double().cpu()
	return float((a - b).norm() / (b.norm() + 1e-30))


def cl(x):
	"""[B, C, T] float -> bf16 channels-last [B, T, C]"""
	return x.permute(0, 2, 1).contiguous().to(BF16)


# ------------------------------------------------------------------------------------------ kernels
def test_bn_act_mask_forward_backward_kernels():
	from convasr_b200 import _lib, ops
	dev = torch.device('cuda:0')
	lib = _lib.load()
	g = torch.Generator().manual_seed(0)
	for act_name, code, a, b in (('hardtanh', _lib.ACT_HARDTANH, 0.0, 20.0), ('relu', _lib.ACT_RELU, 0.0, 0.0), ('leaky_relu', _lib.ACT_LEAKY_RELU, 0.01, 0.0)):
		B, C, T = 3, 96, 77
		y = (torch.randn(B, C, T, generator = g) * 3 + torch.randn(1, C, 1, generator = g)).to(BF16).float()
		gamma, beta = torch.rand(C, generator = g) + 0.5, torch.randn(C, generator = g)
		rm, rv = torch.randn(C, generator = g), torch.rand(C, generator = g) + 0.5
		xlen = torch.tensor([1.0, 0.6, 0.35])
		go = torch.randn(B, C, T, generator = g).to(BF16).float()
		# reference: torch fp32 autograd
		yr = y.clone().requires_grad_(True)
		gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
		rm_ref, rv_ref = rm.clone(), rv.clone()
		z = F.batch_norm(yr, rm_ref, rv_ref, gr, br, True, 0.1, 1e-5)
		o = O._activation(z, (act_name, a, b) if act_name == 'hardtanh' else (act_name, a) if act_name == 'leaky_relu' else (act_name, ))
		o = o * O.temporal_mask(T, O.output_lengths(T, xlen))
		o.backward(go)
		# native
		yd = cl(y).to(dev)
		ws = torch.empty(2, C, dtype = torch.float64, device = dev); ss = torch.empty(4, C, device = dev)
		rm_d, rv_d, gamma_d, beta_d, go_d = rm.to(dev), rv.to(dev), gamma.to(dev), beta.to(dev), cl(go).to(dev)  # keep alive: raw pointers below
		_lib.check(lib.cab_bn_batch_stats(ops._p(yd), B, T, C, C, ops._p(gamma_d), ops._p(beta_d), 1e-5, 0.1, ops._p(rm_d), ops._p(rv_d), ops._p(ws), ops._p(ss), ops._stream()), 'stats')
		out = torch.empty_like(yd)
		xl = xlen.to(dev)
		_lib.check(lib.cab_bn_act_mask_fwd(ops._p(yd), None, ops._p(ss), B, T, C, C, code, a, b, ops._p(xl), ops._p(out), None, 0.0, None, 0, ops._stream()), 'fwd')
		sums = torch.empty(2, C, device = dev); dy = torch.empty_like(yd); part = torch.empty(_lib.BN_SUM_REPLICAS, 2, C, dtype = torch.float64, device = dev)
		_lib.check(lib.cab_bn_act_mask_bwd(ops._p(yd), None, ops._p(go_d), None, ops._p(ss), B, T, C, C, code, a, b, ops._p(xl), ops._p(sums), ops._p(dy), None, 0.0, None, 0, 0, ops._p(part), ops._stream()), 'bwd')
		# without the replica scratch the row-walker variant runs: same results
		sums2 = torch.empty(2, C, device = dev); dy2 = torch.empty_like(yd)
		_lib.check(lib.cab_bn_act_mask_bwd(ops._p(yd), None, ops._p(go_d), None, ops._p(ss), B, T, C, C, code, a, b, ops._p(xl), ops._p(sums2), ops._p(dy2), None, 0.0, None, 0, 0, None, ops._stream()), 'bwd (row walker)')
		torch.cuda.synchronize()
		assert rel(sums2, sums) < 1e-5 and rel(dy2.float(), dy.float()) < 1e-3, act_name
		assert rel(out.float().permute(0, 2, 1), o) < 4e-3, act_name  # bf16 output rounding
		assert torch.allclose(rm_d.cpu(), rm_ref, atol = 1e-5) and torch.allclose(rv_d.cpu(), rv_ref, rtol = 1e-4, atol = 1e-5)
		assert rel(sums[0], br.grad) < 1e-4 and rel(sums[1], gr.grad) < 1e-4, act_name
		assert rel(dy.float().permute(0, 2, 1), yr.grad) < 4e-3, act_name


def test_dropout_in_bn_act_kernels():
	"""F.dropout after the activation (models.py:357-371): keep/(1-p) scaling, rate ~ p, and a backward
	that uses exactly the forward's mask (recomputed from the counter, not stored)."""
	from convasr_b200 import _lib, ops
	dev = torch.device('cuda:0')
	lib = _lib.load()
	g = torch.Generator().manual_seed(5)
	B, C, T, p = 4, 128, 200, 0.2
	y = (torch.randn(B, C, T, generator = g) * 2).to(BF16).float()
	go = torch.randn(B, C, T, generator = g).to(BF16).float()
	gamma, beta = torch.rand(C, generator = g) + 0.5, torch.randn(C, generator = g) * 0.1
	yd, go_d, gamma_d, beta_d = cl(y).to(dev), cl(go).to(dev), gamma.to(dev), beta.to(dev)
	ws = torch.empty(2, C, dtype = torch.float64, device = dev); ss = torch.empty(4, C, device = dev)
	_lib.check(lib.cab_bn_batch_stats(ops._p(yd), B, T, C, C, ops._p(gamma_d), ops._p(beta_d), 1e-5, 0.1, None, None, ops._p(ws), ops._p(ss), ops._stream()), 'stats')
	seed = torch.tensor([1234], dtype = torch.int64, device = dev)
	outs = []
	for salt in (3, 3, 4):
		out = torch.empty_like(yd)
		_lib.check(lib.cab_bn_act_mask_fwd(ops._p(yd), None, ops._p(ss), B, T, C, C, _lib.ACT_RELU, 0.0, 0.0, None, ops._p(out), None, p, ops._p(seed), salt, ops._stream()), 'fwd')
		outs.append(out.float().permute(0, 2, 1).cpu())
	assert torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], outs[2])  # deterministic in (seed, salt)
	# reference without dropout
	yr = y.clone().requires_grad_(True)
	z = F.batch_norm(yr, None, None, gamma, beta, True, 0.1, 1e-5).relu()
	pos = z > 1e-3
	kept = outs[0][pos] != 0
	rate = 1.0 - float(kept.float().mean())
	assert abs(rate - p) < 0.01, rate
	assert torch.allclose(outs[0][pos][kept], (z[pos][kept] / (1 - p)).detach(), rtol = 1e-2, atol = 1e-2)
	# backward against autograd with the recovered mask
	mask = torch.where(pos, (outs[0] != 0).float(), torch.ones_like(z)) / (1 - p)  # where z ~ 0 the mask is irrelevant
	(z * mask).backward(go)
	sums = torch.empty(2, C, device = dev); dy = torch.empty_like(yd); part = torch.empty(_lib.BN_SUM_REPLICAS, 2, C, dtype = torch.float64, device = dev)
	_lib.check(lib.cab_bn_act_mask_bwd(ops._p(yd), None, ops._p(go_d), None, ops._p(ss), B, T, C, C, _lib.ACT_RELU, 0.0, 0.0, None, ops._p(sums), ops._p(dy), None, p, ops._p(seed), 3, 0, ops._p(part), ops._stream()), 'bwd')
	torch.cuda.synchronize()
	assert rel(dy.float().permute(0, 2, 1), yr.grad) < 2e-2


def test_training_step_with_dropout_runs_natively():
	from convasr_b200 import training
	dev = torch.device('cuda:0')
	C = 38
	from convasr_b200 import models
	m = models.Wav2Letter(64, [C], frontend = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window'), dropout = 0.2, check_time_dim_padded = False, base_width = 32).to(dev).train().set_precision('bf16')
	assert training.supported(m) and m.backbone[1].activation.dropout == 0.2
	sig, xlen, y, ylen = [t.to(dev) for t in _batch(C)]
	losses = []
	for _ in range(2):
		m.zero_grad()
		out = m(sig, xlen, y = y, ylen = ylen)
		(out['loss'] * ylen[:, 0]).mean().backward()
		losses.append(out['loss'].detach().clone())
		assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters() if p.requires_grad)
	assert not torch.equal(losses[0], losses[1])  # a fresh dropout mask every step


def test_wgrad_and_dgrad_kernels_against_cpu_autograd():
	from convasr_b200 import engine, ops, training
	dev = torch.device('cuda:0')
	g = torch.Generator().manual_seed(1)
	for (B, T_in, Ci, Co, k, dil, pad) in [(3, 131, 64, 192, 11, 1, 5), (2, 200, 320, 128, 29, 2, 29), (2, 97, 128, 128, 1, 1, 0)]:
		x = torch.randn(B, Ci, T_in, generator = g).to(BF16).float().requires_grad_(True)
		w = (torch.randn(Co, Ci, k, generator = g) / (Ci * k)**0.5).to(BF16).float().requires_grad_(True)
		y = F.conv1d(x, w, None, padding = pad, dilation = dil)
		dy = torch.randn(y.shape, generator = g).to(BF16).float()
		y.backward(dy)
		T_out = y.shape[-1]
		a = cl(dy).to(dev)
		bx = cl(x.detach()).to(dev)
		packed = ops.conv1d_wgrad(a, T_out, Co, bx, T_in, Ci, k, dil, pad)
		dw = training._unpack(packed, k, Co, Ci, transposed = False)
		assert rel(dw, w.grad) < 1e-5, (Ci, Co, k)
		# dgrad = the forward kernel on flipped, transposed weights
		_, w_dgr = training._pack(w.detach().to(dev), Ci, Co, want_dgrad = True)
		gx = torch.empty(B, T_in, Ci, dtype = BF16, device = dev)
		ops.conv1d_fused([ops.Source(a, w_dgr, Co, k, dil, dil * (k - 1) - pad, T_in = T_out)], B, T_in, Ci, out_hi = gx)
		assert rel(gx.float().permute(0, 2, 1), x.grad) < 4e-3, (Ci, Co, k)  # bf16 output rounding


# ------------------------------------------------------------------------------------------ end to end
def _model(dev, kwargs, C = 38, seed = 7, precision = 'bf16'):
	from convasr_b200 import models, training
	m = models.Wav2Letter(64, [C], frontend = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window'), dropout = 0., check_time_dim_padded = False, **kwargs)
	shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith('frontend.')}
	sd = O.synth_state_dict(shapes, seed = seed)
	m.load_state_dict(sd, strict = False)
	m = m.to(dev).train().set_precision(precision)
	assert training.supported(m)
	return m, sd


def _batch(C, seed = 3):
	g = torch.Generator().manual_seed(seed)
	sig = (torch.randn(4, 12000, generator = g) * 3000).round().clamp(-32767, 32767).to(torch.int16)
	xlen = torch.tensor([1.0, 0.9, 0.66, 0.5])
	y = torch.randint(0, C - 1, (4, 1, 14), generator = g)
	ylen = torch.tensor([[14], [11], [8], [5]])
	return sig, xlen, y, ylen


def _oracle_step(sd, sig, xlen, y, ylen, C, **kw):
	ref_sd = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v.clone()) for k, v in sd.items()}
	logits, log_probs, olen = O.model_forward(ref_sd, sig, xlen, model = 'Wav2Letter', training = True, **kw)
	nll = O.ctc_loss_torch(log_probs[0].permute(2, 0, 1), y[:, 0], olen[0], ylen[:, 0], C - 1)
	loss = nll.mean()
	loss.backward()
	return ref_sd, logits[0], loss


def test_training_step_shallow_model_matches_oracle():
	"""6 conv layers (num_blocks = 1): prologue + 1 block x 3 repeats + 2 epilogue layers."""
	dev = torch.device('cuda:0')
	C = 38
	m, sd = _model(dev, dict(base_width = 32, num_blocks = 1))
	assert sum(len(b.conv) for b in m.backbone) == 6
	sig, xlen, y, ylen = _batch(C)
	out = m(sig.to(dev), xlen.to(dev), y = y.to(dev), ylen = ylen.to(dev))
	loss = (out['loss'] * ylen[:, 0].to(dev)).mean()  # train.py:754-755: mean raw CTC NLL
	loss.backward()
	ref_sd, ref_logits, ref_loss = _oracle_step(sd, sig, xlen, y, ylen, C, round_bf16 = True)
	assert rel(out['logits'][0], ref_logits) < 2e-2
	assert abs(float(loss) - float(ref_loss)) / abs(float(ref_loss)) < 2e-2
	named = dict(m.named_parameters())
	keys = [k for k, v in ref_sd.items() if v.is_floating_point() and v.requires_grad]
	errs = {}
	for k in keys:
		assert named[k].grad is not None and named[k].grad.shape == ref_sd[k].grad.shape, k
		errs[k] = rel(named[k].grad, ref_sd[k].grad)
	flat = torch.cat([named[k].grad.flatten().cpu() for k in keys])
	flat_ref = torch.cat([ref_sd[k].grad.flatten() for k in keys])
	print('shallow grads: total rel', rel(flat, flat_ref), 'worst', max(errs.items(), key = lambda kv: kv[1]))
	# The CTC gradient p - occupancy cancels to O(1e-2) of its terms, so the 1 % logits noise of bf16 is
	# a several-% perturbation of d(logits) already, and back-propagation amplifies it like the forward
	# pass does (measured on B200: 8.7 % total, 12 % worst tensor; native vs the fp32 ATen path: 12 %).
	# The exactness of every backward kernel is pinned separately above at 1e-5 .. 4e-3.
	assert max(errs.values()) < 0.25, max(errs.items(), key = lambda kv: kv[1])
	assert rel(flat, flat_ref) < 0.15, rel(flat, flat_ref)
	# running statistics moved exactly like nn.BatchNorm1d(momentum=0.1) on the first layer
	feats = O.masked_instance_norm(O.frontend_logmel(sig, xlen), xlen).to(BF16).float()
	y0 = F.conv1d(feats, sd['backbone.0.conv.0.0.weight'].to(BF16).float(), stride = 2, padding = 5)
	bn = m.backbone[0].bn[0]
	assert torch.allclose(bn.running_mean.cpu(), 0.9 * sd['backbone.0.bn.0.running_mean'] + 0.1 * y0.mean(dim = (0, 2)), atol = 2e-3)
	var_u = y0.transpose(0, 1).reshape(y0.shape[1], -1).var(dim = 1, unbiased = True)
	assert torch.allclose(bn.running_var.cpu(), 0.9 * sd['backbone.0.bn.0.running_var'] + 0.1 * var_u, rtol = 2e-2, atol = 1e-3)
	assert int(bn.num_batches_tracked) == 101


def test_training_step_full_depth_loss_and_gradient_direction():
	dev = torch.device('cuda:0')
	C = 38
	m, sd = _model(dev, dict(base_width = 32))
	assert sum(len(b.conv) for b in m.backbone) == 18
	sig, xlen, y, ylen = _batch(C, seed = 4)
	out = m(sig.to(dev), xlen.to(dev), y = y.to(dev), ylen = ylen.to(dev))
	loss = (out['loss'] * ylen[:, 0].to(dev)).mean()
	loss.backward()
	ref_sd, ref_logits, ref_loss = _oracle_step(sd, sig, xlen, y, ylen, C)  # pure fp32 oracle
	assert abs(float(loss) - float(ref_loss)) / abs(float(ref_loss)) < 0.1
	named = dict(m.named_parameters())
	keys = [k for k, v in ref_sd.items() if v.is_floating_point() and v.requires_grad]
	flat = torch.cat([named[k].grad.flatten().cpu() for k in keys]).double()
	flat_ref = torch.cat([ref_sd[k].grad.flatten() for k in keys]).double()
	cos = float((flat * flat_ref).sum() / (flat.norm() * flat_ref.norm()))
	print('full depth: loss', float(loss), float(ref_loss), 'cos', cos)
	assert cos > 0.7, cos  # 18 chaotic layers each way at bf16 (see module docstring)
	assert all(torch.isfinite(named[k].grad).all() for k in keys)


def test_graphed_train_step_equals_eager_step():
	"""training.GraphedTrainStep (whole step as one CUDA-graph replay) == the eager step."""
	from convasr_b200 import training
	dev = torch.device('cuda:0')
	C = 38
	sig, xlen, y, ylen = [t.to(dev) for t in _batch(C)]
	losses = []
	for graphed in (False, True):
		m, sd = _model(dev, dict(base_width = 32, num_blocks = 1))
		opt = torch.optim.SGD(m.parameters(), lr = 1e-3, momentum = 0.9)
		if graphed:
			step = training.GraphedTrainStep(m, opt, sig, xlen, y, ylen, warmup = 2)
			m.load_state_dict(sd, strict = False)  # warm-up + capture stepped the weights: restart from the same point
			for st in opt.state.values():  # the graph holds these buffers: reset their contents, keep the tensors
				st['momentum_buffer'].zero_()
			out = [step(sig, xlen, y, ylen) for _ in range(3)]
		else:
			out = []
			for _ in range(3):
				opt.zero_grad(set_to_none = True)
				o = m(sig, xlen, y = y, ylen = ylen)
				(o['loss'] * ylen[:, 0]).mean().backward()
				opt.step()
				out.append(o['loss'].detach().clone())
		losses.append(torch.stack(out).cpu())
	# step 1 sees identical weights; later steps differ only through atomics ordering in the reductions
	assert torch.allclose(losses[0][0], losses[1][0], rtol = 1e-4, atol = 1e-4)
	assert torch.allclose(losses[0], losses[1], rtol = 2e-2, atol = 2e-2)


# ------------------------------------------------------------------------------------------ data parallel
def _grad_sync_worker(rank, world, port, tmp):
	import os
	import torch.distributed as dist
	from convasr_b200 import models, optimizers, parallel, training
	os.environ.update(RANK = str(rank), WORLD_SIZE = str(world), LOCAL_RANK = str(rank), MASTER_ADDR = '127.0.0.1', MASTER_PORT = str(port))
	parallel.init_from_env()
	dev = torch.device('cuda', rank)
	C = 38
	batches = [[t.to(dev) for t in _batch(C, seed = 3 + k)] for k in range(world)]
	# per-rank gradients without any exchange, from identical weights and BN state; then the data-parallel replica: same
	# module, gradients averaged inside the native backward.  Both precision tiers at 1e-5: the BatchNorm statistics (forward
	# and backward) are accumulated in fp64, so the forward pass is bit-reproducible run to run and GPU to GPU and the only
	# difference left is the order of the fp32 split-K reductions of the weight gradients (measured 6e-8 on 2 x B200).  With
	# the fp32 atomics of round 1 a flipped bf16 rounding / activation gate cost 2e-3 .. 9e-3 here.
	measured = {}
	for precision, bound in (('fp32', 1e-5), ('bf16', 1e-5)):
		m, sd = _model(dev, dict(base_width = 32, num_blocks = 1), precision = precision)
		expect = None
		for k in range(world):
			m.load_state_dict(sd, strict = False)
			m.zero_grad(set_to_none = True)
			sig, xlen, y, ylen = batches[k]
			out = m(sig, xlen, y = y, ylen = ylen)
			(out['loss'] * ylen[:, 0]).mean().backward()
			g = {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}
			expect = g if expect is None else {n: expect[n] + g[n] for n in g}
		m.load_state_dict(sd, strict = False)
		m.zero_grad(set_to_none = True)
		net, _ = models.distributed_data_parallel_and_autocast(m, rank, opt_level = 'O2' if precision == 'bf16' else None)
		assert net is m and isinstance(m._grad_sync, parallel.GradSync) and m._active_precision() == precision
		sig, xlen, y, ylen = batches[rank]
		out = net(sig, xlen, y = y, ylen = ylen)
		(out['loss'] * ylen[:, 0]).mean().backward()
		torch.cuda.synchronize()
		assert m._grad_sync.n_collectives == 1 + 1  # one bucket holds every packed weight gradient of this small model, plus the flat small-tensor buffer
		errs = sorted(((rel(p.grad, expect[n] / world), n) for n, p in m.named_parameters() if n in expect), reverse = True)
		worst = errs[0][0]
		measured[precision] = worst
		assert worst < bound, (precision, errs[:4])
	print(f'rank {rank}: averaged gradients vs mean of per-rank gradients, worst tensor: {measured}', flush = True)
	# whole step as a CUDA graph with the all-reduces captured inside: replicas stay bit-identical
	opt = optimizers.SGD([p for p in m.parameters() if p.requires_grad], lr = 1e-3, momentum = 0.9)
	step = training.GraphedTrainStep(net, opt, sig, xlen, y, ylen, warmup = 2, max_grad_norm = 100.0)
	for _ in range(3):
		loss = step(sig, xlen, y, ylen)
	assert bool(torch.isfinite(loss).all())
	flat = torch.cat([p.detach().flatten() for p in m.parameters() if p.requires_grad])
	both = [torch.empty_like(flat) for _ in range(world)]
	dist.all_gather(both, flat)
	assert all(torch.equal(both[0], b) for b in both[1:])
	open(os.path.join(tmp, f'ok{rank}'), 'w').write('ok')
	# tear down in this order: a live CUDA graph that holds captured NCCL kernels makes destroy_process_group() wait forever
	del step
	import gc
	gc.collect()
	torch.cuda.synchronize()
	dist.barrier()
	os._exit(0)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason = 'needs two GPUs')
def test_data_parallel_native_grad_sync_two_gpus(tmp_path):
	import socket
	import torch.multiprocessing as mp
	s = socket.socket()
	s.bind(('127.0.0.1', 0))
	port = s.getsockname()[1]
	s.close()
	mp.spawn(_grad_sync_worker, args = (2, port, str(tmp_path)), nprocs = 2, join = True)
	assert all((tmp_path / f'ok{r}').exists() for r in range(2))


# ------------------------------------------------------------------------------------------ ragged batches
def test_padding_tiles_are_skipped_exactly():
	"""Tiles that lie entirely in the padding of a ragged batch are not computed (cab_conv_epilogue_t.skip_frac,
	cab_conv1d_wgrad skip_*): the forward result is bit-identical, its BN statistics and the weight gradient equal
	up to the order of the fp32 reductions."""
	from convasr_b200 import ops, training
	dev = torch.device('cuda:0')
	g = torch.Generator().manual_seed(11)
	B, T, Ci, Co, K, pad = 4, 700, 128, 192, 11, 5
	xlen = torch.tensor([1.0, 0.55, 0.3, 0.05]).to(dev)
	lens = ops.frac_lengths(xlen, T)
	x = torch.randn(B, T, Ci, generator = g).to(BF16).to(dev)
	x = x * (torch.arange(T, device = dev)[None, :, None] < lens[:, None, None])  # zero past the valid frames
	w = (torch.randn(Co, Ci, K, generator = g) / (Ci * K) ** 0.5).to(dev)
	w_fwd, _ = training._pack(w, Ci, Co, want_dgrad = False)
	outs = []
	for skip in (None, (xlen, T, pad)):
		y = torch.full((B, T, Co), float('nan'), dtype = BF16, device = dev)
		st = torch.empty(2, Co, dtype = torch.float64, device = dev)
		ops.conv1d_fused([ops.Source(x, w_fwd, Ci, K, 1, pad, T_in = T)], B, T, Co, out_hi = y, stats = st, skip = skip)
		outs.append((y, st))
	torch.cuda.synchronize()
	assert torch.equal(outs[0][0], outs[1][0])
	assert torch.allclose(outs[0][1], outs[1][1], rtol = 1e-12, atol = 1e-9)  # fp64 accumulators: only the order of the partial sums differs
	assert bool((outs[1][0][3, 200:] == 0).all())  # utterance 3: 35 valid frames, tiles 2.. are pure padding
	# with the launch's own temporal mask the skip is implied
	ym = [torch.empty(B, T, Co, dtype = BF16, device = dev) for _ in range(2)]
	ops.conv1d_fused([ops.Source(x, w_fwd, Ci, K, 1, pad, T_in = T)], B, T, Co, out_hi = ym[0], xlen = xlen, act = 1)
	ref = torch.relu(outs[0][0].float()) * (torch.arange(T, device = dev)[None, :, None] < lens[:, None, None])
	assert torch.equal(ym[0].float(), ref)
	# weight gradient: both operand orientations
	dy = torch.randn(B, T, Co, generator = g).to(BF16).to(dev)
	for a, a_C, bx, b_C, dil, pd, margin in ((dy, Co, x, Ci, 1, pad, pad), (x, Ci, dy, Co, -1, -pad, 0)):
		full = ops.conv1d_wgrad(a, T, a_C, bx, T, b_C, K, dil, pd)
		part = ops.conv1d_wgrad(a, T, a_C, bx, T, b_C, K, dil, pd, skip = (xlen, T, margin))
		assert rel(part, full) < 1e-6, rel(part, full)


def test_bn_backward_reduction_folded_into_the_dgrad_epilogue():
	"""cab_conv_epilogue_t.bnr_*: the dgrad GEMM accumulates the BatchNorm-backward channel sums of the repeat whose output
	gradient it produces, cab_bn_act_mask_bwd_apply finishes.  Against the unfused pair (same GEMM + cab_bn_act_mask_bwd): the
	gradient activation is bit-identical, dbeta / dgamma / grad_y agree up to the order of the fp32 partial sums."""
	from convasr_b200 import _lib, ops, training
	dev = torch.device('cuda:0')
	lib = _lib.load()
	g = torch.Generator().manual_seed(5)
	cases = [
		# B, T, C_up (channels of the layer above), C (this layer), k, dil, pad, act, ragged mask, skip padding tiles
		(3, 300, 128, 96, 7, 1, 3, (_lib.ACT_HARDTANH, 0.0, 20.0), True, True),
		(2, 129, 64, 320, 1, 1, 0, (_lib.ACT_RELU, 0.0, 0.0), False, False),
		(4, 257, 192, 256, 11, 2, 10, (_lib.ACT_LEAKY_RELU, 0.01, 0.0), True, False),
		(2, 140, 64, 40, 3, 1, 1, (_lib.ACT_HARDTANH, -1.0, 1.0), True, True),  # C not a multiple of the allocation (40 -> 64)
	]
	for B, T, C_up, C, k, dil, pad, (code, a, b), ragged, skip_tiles in cases:
		ld = (C + 63) // 64 * 64
		xlen = (torch.tensor([1.0, 0.52, 0.3, 0.07])[:B] if ragged else torch.ones(B)).to(dev)
		# the layer above: dy_up [B, T, C_up] and its dgrad weights [k, C (ld), C_up]
		dy_up = torch.randn(B, T, C_up, generator = g).to(BF16).to(dev)
		w = (torch.randn(C_up, C, k, generator = g) / (C * k) ** 0.5).to(dev)
		_, w_dgr = training._pack(w, ld, C_up, want_dgrad = True)
		y = torch.zeros(B, T, ld, dtype = BF16, device = dev)
		y[:, :, :C] = (torch.randn(B, T, C, generator = g) * 2 + torch.randn(1, 1, C, generator = g)).to(BF16).to(dev)
		gamma, beta = (torch.rand(C, generator = g) + 0.5).to(dev), (torch.randn(C, generator = g) + (8.0 if code == _lib.ACT_HARDTANH and b == 20.0 else 0.0)).to(dev)
		gamma[::7] *= -1  # negative scales flip the side of the gate
		rm, rv = torch.zeros(C, device = dev), torch.ones(C, device = dev)
		ws, ss = torch.empty(2, C, dtype = torch.float64, device = dev), torch.empty(4, C, device = dev)
		_lib.check(lib.cab_bn_batch_stats(ops._p(y), B, T, C, ld, ops._p(gamma), ops._p(beta), 1e-5, 0.1, ops._p(rm), ops._p(rv), ops._p(ws), ops._p(ss), ops._stream()), 'stats')
		src = [ops.Source(dy_up, w_dgr, C_up, k, dil, dil * (k - 1) - pad, T_in = T)]
		skip = (xlen, T, 0) if (skip_tiles and ragged) else None
		mask = xlen if ragged else None
		out = {}
		for fold in (False, True):
			gx = torch.full((B, T, ld), float('nan'), dtype = BF16, device = dev)
			partials = torch.full((_lib.BN_SUM_REPLICAS * 2 * C, ), float('nan'), dtype = torch.float64, device = dev)
			sums = torch.empty(2, C, device = dev)
			dyl = torch.empty(B, T, ld, dtype = BF16, device = dev)
			ops.conv1d_fused(src, B, T, ld, out_hi = gx, skip = skip, bn_reduce = (y, ss, C, code, a, b, mask, partials) if fold else None)
			fn = lib.cab_bn_act_mask_bwd_apply if fold else lib.cab_bn_act_mask_bwd
			_lib.check(fn(ops._p(y), None, ops._p(gx), None, ops._p(ss), B, T, C, ld, code, a, b, ops._p(mask), ops._p(sums), ops._p(dyl), None, 0.0, None, 0, 0, ops._p(partials), ops._stream()), 'bn bwd')
			torch.cuda.synchronize()
			out[fold] = (gx, sums, dyl)
		lens = ops.frac_lengths(xlen, T)
		valid = (torch.arange(T, device = dev)[None, :, None] < lens[:, None, None])
		if skip is None:
			assert torch.equal(out[False][0], out[True][0]), 'the folded launch must store the same gradient activation'
		else:
			assert torch.equal(out[False][0] * valid, out[True][0] * valid)
		assert float(out[False][1].abs().max()) > 0
		assert rel(out[True][1], out[False][1]) < 1e-4, (C, rel(out[True][1], out[False][1]))
		assert rel(out[True][2][:, :, :C].float(), out[False][2][:, :, :C].float()) < 2e-3, (C, rel(out[True][2].float(), out[False][2].float()))  # bf16 outputs: a last-bit difference of a coefficient moves roundings


def test_training_step_same_with_and_without_folded_bn_reduction():
	from convasr_b200 import training
	dev = torch.device('cuda:0')
	sig, xlen, y, ylen = _batch(38)
	grads = []
	for fold in (False, True):
		old, training._FOLD_BN_REDUCE = training._FOLD_BN_REDUCE, fold
		try:
			m, _ = _model(dev, dict(base_width = 32, num_blocks = 2))
			out = m(sig.to(dev), xlen.to(dev), y = y.to(dev), ylen = ylen.to(dev))
			out['loss'].mean().backward()
			grads.append({k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None})
		finally:
			training._FOLD_BN_REDUCE = old
	assert grads[0].keys() == grads[1].keys() and len(grads[0]) > 10
	worst = max(rel(grads[1][k], grads[0][k]) for k in grads[0])
	assert worst < 2e-2, worst  # bf16 tier: an ulp in a BatchNorm-backward coefficient moves bf16 roundings of the gradient activations downstream


def test_training_step_same_with_and_without_padding_skip():
	from convasr_b200 import training
	dev = torch.device('cuda:0')
	C = 38
	sig, xlen, y, ylen = _batch(C)
	xlen = torch.tensor([1.0, 0.9, 0.4, 0.2])
	res = []
	for flag in (False, True):
		training._SKIP_PADDING = flag
		try:
			m, sd = _model(dev, dict(base_width = 32, num_blocks = 1))
			out = m(sig.to(dev), xlen.to(dev), y = y.to(dev), ylen = torch.tensor([[14], [11], [5], [2]]).to(dev))
			out['loss'].sum().backward()
			res.append((out['logits'][0].detach().clone(), torch.cat([p.grad.flatten() for p in m.parameters() if p.grad is not None])))
		finally:
			training._SKIP_PADDING = True
	assert rel(res[1][0], res[0][0]) < 1e-3, rel(res[1][0], res[0][0])  # forward: same numbers up to the order of the fp32 atomics behind the BN statistics (measured 1e-5)
	assert rel(res[1][1], res[0][1]) < 1e-2, rel(res[1][1], res[0][1])  # run-to-run atomics noise amplified by the backward (see above)


def test_batched_weight_pack_equals_per_layer_pack():
	from convasr_b200 import training
	dev = torch.device('cuda:0')
	g = torch.Generator().manual_seed(2)
	shapes = [(256, 64, 11), (384, 256, 11), (896, 768, 29), (1024, 896, 1), (38, 1024, 1), (70, 50, 5)]
	ws = [torch.randn(*s, generator = g).to(dev) for s in shapes]
	specs = [(w, (w.shape[1] + 63) // 64 * 64, (w.shape[0] + 63) // 64 * 64, i != 0) for i, w in enumerate(ws)]
	batched = training._pack_all(specs)
	for (w, ci_ld, co_ld, want_dgrad), (fwd, dgr, _, _) in zip(specs, batched):
		fwd1, dgr1 = training._pack(w, ci_ld, co_ld, want_dgrad)
		assert torch.equal(fwd, fwd1) and (dgr is None) == (dgr1 is None) and (dgr is None or torch.equal(dgr, dgr1))
		Co, Ci, K = w.shape
		assert torch.equal(fwd[:, :, :Ci], w.permute(2, 0, 1).to(BF16))
		if dgr is not None:
			assert torch.equal(dgr[:, :, :Co], w.flip(2).permute(2, 1, 0).to(BF16))


# ------------------------------------------------------------------------------------------ golden-backed (reference TRAIN mode)
def _golden_model(dev, c, precision):
	from convasr_b200 import models, training
	from test_oracle_golden import golden_state_dict
	kw = {k: v for k, v in c['kwargs'].items() if k != 'freeze'}
	m = getattr(models, c['model'])(64, [c['num_classes']], frontend = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window'), dropout = 0., check_time_dim_padded = False, **kw)
	m.load_state_dict(golden_state_dict(c), strict = False)
	m = m.to(dev).train().set_precision(precision)
	if c['kwargs'].get('freeze'):
		m.freeze(backbone = c['kwargs']['freeze'])
	assert training.unsupported_reason(m) is None, training.unsupported_reason(m)
	return m


def _native_train_step(dev, c, precision):
	m = _golden_model(dev, c, precision)
	ylen = c['ylen'].to(dev)
	out = m(c['signal'].to(dev), c['xlen'].to(dev), y = c['y'].to(dev), ylen = ylen)
	(out['loss'] * ylen[:, 0]).mean().backward()  # train.py:754-755
	grads = {k: p.grad for k, p in m.named_parameters() if p.grad is not None and not k.startswith('frontend.')}
	stats = {k: v.detach().cpu() for k, v in m.state_dict().items() if k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))}
	return out, grads, stats


def test_training_step_matches_reference_train_mode_golden(golden, capsys):
	"""The native training step against the REFERENCE's own train-mode outputs (oracle/make_golden.py: golden_train):
	logits, per-utterance loss, every parameter gradient, BatchNorm running statistics after the step -- plain, residual,
	dense, separable, mask-free and frozen-backbone families.
	fp32 (split-bf16) tier: every FORWARD quantity (logits, loss, running statistics) within 1e-3 of the reference for every
	family (measured 2e-5 .. 2e-4); gradients within 1e-3 where no activation gate can flip (the full-depth smooth case,
	measured 7e-5 total / 1.2e-4 worst tensor).  Where the nonlinearity has reachable kinks, the tier's 2e-5 forward noise
	flips a handful of gates (an element within 2e-5 of a kink: ~2 of the 120 k pre-activations of the shallow model) and
	each flip moves its layer's gradient by ~1 / sqrt(elements) = 7e-3 -- the same mechanism that separates the reference
	from its own fp32 CPU restatement by 1.6e-2 at full depth (tests/test_oracle_golden.py); those cases are bounded at 1e-1
	per tensor (measured 8e-3 .. 7e-2).  bf16 tier: measured and reported next to it, bounded loosely."""
	from test_oracle_golden import check_grads_against_golden
	dev = torch.device('cuda:0')
	report, failures = [], []
	for c in golden('train')['cases']:
		deep_kinked = 'nonlinearity' not in c['kwargs'] and c['kwargs'].get('num_blocks', 5) > 1 and c['model'] in ('Wav2Letter', 'JasperNetSeparable')
		for precision in ('fp32', 'bf16'):
			out, grads, stats = _native_train_step(dev, c, precision)
			assert torch.equal(out['olen'][0].cpu(), c['olen'])
			e_logits = rel(out['logits'][0], c['logits'])
			e_loss = float(((out['loss'].cpu() - c['loss']).abs() / c['loss'].abs()).max())
			kinked = 'nonlinearity' not in c['kwargs']
			tol_fwd, tol_grad = (1e-3, 1e-1 if kinked else 1e-3) if precision == 'fp32' else (0.2 if deep_kinked or c['model'] == 'Wav2LetterResidual' else 5e-2, 1.0)
			total, worst = check_grads_against_golden(grads, c['grads'], 1e9, (c['model'], precision))
			e_stats = 0.0
			for k, v in c['stats'].items():
				if k.endswith('num_batches_tracked'):
					assert int(stats[k]) == int(v), k
				else:
					e_stats = max(e_stats, float((stats[k] - v).abs().max() / (v.abs().max() + 1e-6)))
			line = f'{c["model"]:22s} {str(c["kwargs"]):64s} {precision}: logits {e_logits:.2e} loss {e_loss:.2e} grads total {total:.2e} worst tensor {worst:.2e} running stats {e_stats:.2e}'
			report.append(line)
			ok = e_logits < tol_fwd and e_loss < max(tol_fwd, 1e-4) * (1 if precision == 'fp32' else 5) and worst < tol_grad and e_stats < (2e-3 if precision == 'fp32' else 5e-2)
			if not ok:
				failures.append(line)
	with capsys.disabled():
		print('\nnative training step vs reference TRAIN-mode golden:\n  ' + '\n  '.join(report))
	assert not failures, failures


def test_inplace_family_trains_natively_like_the_reference_twin(golden, capsys):
	"""Wav2LetterDenseNoDilationInplace (section 8f rank 4) on the native training step: InplaceBatchNorm1d / the invertible
	leaky_relu (models.py:376-433) compute BatchNorm1d / leaky_relu, so the golden is the reference's non-in-place twin
	(tests/golden/train_inplace.pt; the in-place forward itself needs CUDA-only ATen operators).  Same bars as the other
	kinked families."""
	from test_oracle_golden import check_grads_against_golden
	from convasr_b200 import training
	dev = torch.device('cuda:0')
	lines = []
	for c in golden('train_inplace')['cases']:
		for precision in ('fp32', 'bf16'):
			out, grads, stats = _native_train_step(dev, c, precision)
			assert torch.equal(out['olen'][0].cpu(), c['olen'])
			e_logits = rel(out['logits'][0], c['logits'])
			e_loss = float(((out['loss'].cpu() - c['loss']).abs() / c['loss'].abs()).max())
			total, worst = check_grads_against_golden(grads, c['grads'], 1e9, (c['model'], precision))
			e_stats = max(float((stats[k] - v).abs().max() / (v.abs().max() + 1e-6)) for k, v in c['stats'].items() if not k.endswith('num_batches_tracked'))
			lines.append(f'{c["model"]} {precision}: logits {e_logits:.2e} loss {e_loss:.2e} grads total {total:.2e} worst tensor {worst:.2e} running stats {e_stats:.2e}')
			if precision == 'fp32':
				assert e_logits < 1e-3 and e_loss < 1e-3 and worst < 1e-1 and e_stats < 2e-3, lines[-1]
			else:
				assert e_logits < 0.2 and e_loss < 5e-2 and worst < 1.0 and e_stats < 5e-2, lines[-1]
	with capsys.disabled():
		print('\nnative training step, *Inplace family vs the reference twin:\n  ' + '\n  '.join(lines))


def test_multi_branch_bn_kernels_against_autograd():
	"""cab_bn_multi_act_mask_fwd / cab_act_mask_bwd_dz + per-branch BatchNorm backward == torch autograd of
	act(BN0(y0) + BN1(y1) + y2) * mask (two BatchNorm branches and an identity branch), both tiers"""
	import ctypes
	from convasr_b200 import _lib, ops
	dev = torch.device('cuda:0')
	lib = _lib.load()
	g = torch.Generator().manual_seed(3)
	B, C, T = 3, 96, 77
	xlen = torch.tensor([1.0, 0.6, 0.35])
	for act_name, code, a, b in (('hardtanh', _lib.ACT_HARDTANH, 0.0, 20.0), ('relu', _lib.ACT_RELU, 0.0, 0.0), ('leaky_relu', _lib.ACT_LEAKY_RELU, 0.01, 0.0)):
		for split in (False, True):
			q = (lambda t: t) if split else (lambda t: t.to(BF16).float())
			ys = [q(torch.randn(B, C, T, generator = g) * 2 + torch.randn(1, C, 1, generator = g)) for _ in range(3)]
			gammas = [torch.rand(C, generator = g) + 0.5 for _ in range(2)]
			betas = [torch.randn(C, generator = g) + (5.0 if act_name == 'hardtanh' else 0.0) for _ in range(2)]
			go = q(torch.randn(B, C, T, generator = g))
			yr = [y.clone().requires_grad_(True) for y in ys]
			gr, br = [x.clone().requires_grad_(True) for x in gammas], [x.clone().requires_grad_(True) for x in betas]
			z = F.batch_norm(yr[0], None, None, gr[0], br[0], True, 0.1, 1e-5) + F.batch_norm(yr[1], None, None, gr[1], br[1], True, 0.1, 1e-5) + yr[2]
			o = O._activation(z, (act_name, a, b) if act_name == 'hardtanh' else (act_name, a) if act_name == 'leaky_relu' else (act_name, ))
			o = o * O.temporal_mask(T, O.output_lengths(T, xlen))
			o.backward(go)

			def dv(t):  # [B, C, T] fp32 -> (hi, lo) channels-last on the device
				x = t.permute(0, 2, 1).contiguous()
				hi = x.to(BF16)
				return hi.to(dev), ((x - hi.float()).to(BF16).to(dev) if split else None)

			y_d = [dv(y) for y in ys]
			go_d = dv(go)
			xl = xlen.to(dev)
			sss = []
			for i in range(2):
				ws = torch.empty(2, C, dtype = torch.float64, device = dev); ss = torch.empty(4, C, device = dev)
				# statistics from the fp32 values (the conv epilogue accumulates hi + lo in the split tier)
				yf = (y_d[i][0].float() + (y_d[i][1].float() if split else 0)).reshape(-1, C)
				sums = torch.stack([yf.double().sum(0), (yf.double() * yf.double()).sum(0)]).contiguous()
				gd, bd = gammas[i].to(dev), betas[i].to(dev)
				_lib.check(lib.cab_bn_finalize(ops._p(sums), B * T, C, ops._p(gd), ops._p(bd), 1e-5, 0.1, None, None, ops._p(ss), ops._stream()), 'finalize')
				sss.append(ss)
			branches = [(y_d[0], sss[0]), (y_d[1], sss[1]), (y_d[2], None)]
			arr = (_lib.BnBranch * 3)(*[_lib.BnBranch(t[0].data_ptr(), t[1].data_ptr() if split else None, s.data_ptr() if s is not None else None) for t, s in branches])
			out_hi = torch.empty(B, T, C, dtype = BF16, device = dev); out_lo = torch.empty_like(out_hi) if split else None
			_lib.check(lib.cab_bn_multi_act_mask_fwd(arr, 3, B, T, C, C, code, a, b, ops._p(xl), ops._p(out_hi), ops._p(out_lo), 0.0, None, 0, ops._stream()), 'multi fwd')
			dz_hi = torch.empty_like(out_hi); dz_lo = torch.empty_like(out_hi) if split else None
			_lib.check(lib.cab_act_mask_bwd_dz(ops._p(out_hi), ops._p(out_lo), ops._p(go_d[0]), ops._p(go_d[1]), B, T, C, C, code, a, b, ops._p(xl), ops._p(dz_hi), ops._p(dz_lo), 0.0, None, 0, ops._stream()), 'dz')
			tol = 2e-5 if split else 4e-3
			val = lambda hi, lo: (hi.float() + (lo.float() if lo is not None else 0)).permute(0, 2, 1)
			assert rel(val(out_hi, out_lo), o) < tol, (act_name, split, rel(val(out_hi, out_lo), o))
			assert rel(val(dz_hi, dz_lo), yr[2].grad) < tol, (act_name, split)  # the identity branch receives dz itself
			part = torch.empty(_lib.BN_SUM_REPLICAS, 2, C, dtype = torch.float64, device = dev)
			for i in range(2):
				sums = torch.empty(2, C, device = dev); dy_hi = torch.empty_like(out_hi); dy_lo = torch.empty_like(out_hi) if split else None
				_lib.check(lib.cab_bn_act_mask_bwd(ops._p(y_d[i][0]), ops._p(y_d[i][1]), ops._p(dz_hi), ops._p(dz_lo), ops._p(sss[i]), B, T, C, C, _lib.ACT_NONE, 0.0, 0.0, None, ops._p(sums), ops._p(dy_hi), ops._p(dy_lo),
													0.0, None, 0, 0, ops._p(part), ops._stream()), 'branch bwd')
				torch.cuda.synchronize()
				assert rel(val(dy_hi, dy_lo), yr[i].grad) < (1e-4 if split else 6e-3), (act_name, split, i, rel(val(dy_hi, dy_lo), yr[i].grad))
				assert rel(sums[0], br[i].grad) < (1e-4 if split else 5e-3) and rel(sums[1], gr[i].grad) < (1e-4 if split else 5e-3), (act_name, split, i)


def test_grouped_conv_training_kernels_against_autograd():
	"""separable blocks: the grouped conv's weight / bias gradient (cab_grouped_conv1d_wgrad) and input gradient
	(cab_grouped_conv1d with in-group transposed, tap-flipped weights, no ReLU) vs torch autograd, both tiers"""
	from convasr_b200 import engine, ops
	dev = torch.device('cuda:0')
	g = torch.Generator().manual_seed(4)
	for (B, T, C_in, C_out, groups, k) in [(3, 150, 64, 96, 32, 11), (2, 90, 256, 384, 128, 13), (2, 70, 96, 96, 32, 25), (2, 65, 224, 224, 32, 5)]:
		for split in (False, True):
			q = (lambda t: t) if split else (lambda t: t.to(BF16).float())
			x = q(torch.randn(B, C_in, T, generator = g)).requires_grad_(True)
			w = (torch.randn(C_out, C_in // groups, k, generator = g) / (k * C_in / groups)**0.5).requires_grad_(True)
			bias = torch.randn(C_out, generator = g).requires_grad_(True)
			y = F.conv1d(x, w, bias, padding = k // 2, groups = groups)
			dy = q(torch.randn(y.shape, generator = g))
			y.backward(dy)

			def dv(t, ld):
				xcl = torch.zeros(t.shape[0], t.shape[2], ld)
				xcl[:, :, :t.shape[1]] = t.detach().permute(0, 2, 1)
				hi = xcl.to(BF16)
				return engine._Act(hi.to(dev), (xcl - hi.float()).to(BF16).to(dev) if split else None, t.shape[2], t.shape[1])

			ld_in, ld_out = engine._ceil_to(C_in, 64), engine._ceil_to(C_out, 64)
			xd, dyd = dv(x, ld_in), dv(dy, ld_out)
			db = torch.empty(C_out, device = dev)
			dw = ops.grouped_conv1d_wgrad(dyd, T, xd, T, C_in, C_out, groups, k, k // 2, db)
			tol = 2e-5 if split else 1e-5
			assert rel(dw, w.grad) < tol, (C_in, C_out, k, split, rel(dw, w.grad))
			assert rel(db, bias.grad) < tol
			cin_g, cout_g = C_in // groups, C_out // groups
			wt = w.detach().view(groups, cout_g, cin_g, k).permute(0, 2, 1, 3).flip(3).reshape(C_in, cout_g, k).contiguous().to(dev)
			dx_hi, dx_lo = ops.grouped_conv1d(dyd.hi, T, C_out, wt, None, groups, k - 1 - k // 2, ld_out = ld_in, act_lo = dyd.lo, want_lo = split, relu = False)
			dx = dx_hi.float() + (dx_lo.float() if split else 0)
			assert rel(dx[:, :, :C_in].permute(0, 2, 1), x.grad) < (2e-5 if split else 4e-3), (C_in, C_out, k, split)


def test_eval_plan_is_rebuilt_after_graphed_training_steps():
	"""the native kernels write parameters / running statistics through raw pointers; a CUDA-graph replay
	runs no Python, so the cached eval plan (folded BN, packed weights) must be invalidated explicitly.  Also: the learning
	rate a scheduler sets between replays reaches the device cell the replay reads."""
	from convasr_b200 import optimizers, training
	dev = torch.device('cuda:0')
	C = 38
	m, sd = _model(dev, dict(base_width = 32, num_blocks = 1))
	sig, xlen, y, ylen = [t.to(dev) for t in _batch(C)]
	opt = optimizers.SGD([p for p in m.parameters() if p.requires_grad], lr = 1e-2, momentum = 0.9)
	sched = optimizers.MultiStepLR(opt, gamma = 0.1, milestones = [2])
	step = training.GraphedTrainStep(m, opt, sig, xlen, y, ylen, warmup = 2)

	def eval_logits():
		m.eval()
		with torch.no_grad():
			lg = m(sig, xlen)['logits'][0].clone()
		m.train()
		return lg

	first = eval_logits()
	w0 = m.backbone[1].conv[0][0].weight.detach().clone()
	step(sig, xlen, y, ylen)
	w1 = m.backbone[1].conv[0][0].weight.detach().clone()
	sched.step(2)  # lr 1e-2 -> 1e-3 from here on
	step(sig, xlen, y, ylen)
	w2 = m.backbone[1].conv[0][0].weight.detach().clone()
	second = eval_logits()
	assert not torch.equal(first, second), 'eval after graph replays must see the trained weights and moved running statistics'
	m._plan = None  # a plan built from scratch gives the same numbers as the one the signature refreshed
	assert torch.equal(second, eval_logits())
	d1, d2 = float((w1 - w0).norm()), float((w2 - w1).norm())
	assert d2 < 0.5 * d1, (d1, d2)  # the 10x smaller learning rate took effect inside the replay (momentum carries some of the old step)
