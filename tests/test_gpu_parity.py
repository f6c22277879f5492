"""Parity tests proper: the CUDA path (through the drop-in modules and the C ABI) against the
committed golden vectors of the real reference and against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star): logits within 1e-3 relative in the fp32 tier and 2e-2 in bf16;
CTC loss and gradients within 1e-4 relative; alignments / greedy ids / collapsed text bit-exact.
Relative error = Frobenius norm of the difference over the norm of the reference.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O


def rel(a, b):
	a, b = a.detach().double().cpu(), b.detach().double().cpu()
	return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope = 'module')
def dev():
	return torch.device('cuda:0')


# ---------------------------------------------------------------- frontend / instance norm
def test_frontend_against_golden(golden, dev):
	from convasr_b200 import models
	g = golden('frontend')
	fe = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window').to(dev)
	norm = models.MaskedInstanceNorm1d(64, affine = False, eps = 2.0**-14, track_running_stats = False, temporal_mask = True, legacy = True)
	for c in g['cases']:
		sig = c['signal'].to(dev)
		xlen = c['xlen'].to(dev) if c['xlen'] is not None else None
		logmel = fe(sig, xlen = xlen)
		assert logmel.shape == c['logmel'].shape and logmel.dtype == torch.float32
		assert float((logmel.cpu() - c['logmel']).abs().max()) < 2e-4, c['name']
		# the mask-based entry point of the reference (models.py:290) gives the same result
		if xlen is not None:
			T = sig.shape[-1]
			mask = torch.arange(T, device = dev)[None] < (xlen * T).ceil().long()[:, None]
			assert torch.equal(fe(sig, mask = mask), logmel)
		feats = norm(c['logmel'].to(dev), xlen = xlen)
		assert float((feats.cpu() - c['feats']).abs().max()) < 2e-5, c['name']


def test_frontend_against_oracle_at_bench_shape(dev):
	"""config C2 row shape (15 s of 8 kHz audio), int16 and fp32 input, ragged lengths."""
	from convasr_b200 import models
	g = torch.Generator().manual_seed(0)
	sig = (torch.randn(4, 120000, generator = g) * 3000).round().clamp(-32767, 32767)
	xlen = torch.tensor([1.0, 0.5, 0.777, 0.9001])
	ref = O.frontend_logmel(sig, xlen)
	fe = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window').to(dev)
	for s in (sig.to(dev), sig.to(torch.int16).to(dev)):
		got = fe(s, xlen = xlen.to(dev))
		assert got.shape == (4, 64, 1501)
		assert float((got.cpu() - ref).abs().max()) < 2e-4
	# scale invariance up to the +1e-5 (quirk 1): int16 without /32768 == float / 32768
	assert float((fe(sig.to(dev) / 32768.0, xlen = xlen.to(dev)) - got).abs().max()) < 1e-3


# ---------------------------------------------------------------- conv stack
def _build(c, dev, precision):
	from convasr_b200 import models
	m = getattr(models, c['model'])(64, [c['num_classes']], frontend = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window'), dropout = 0., check_time_dim_padded = False, **c['kwargs'])
	m.load_state_dict(O.synth_state_dict(c['shapes'], seed = c['seed']), strict = False)
	m = m.to(dev).eval().set_precision(precision)
	if c['fused']:
		m.fuse_conv_bn_eval()
	return m


@pytest.mark.parametrize('precision,tol', [('fp32', 1e-3), ('bf16', 2e-2)])
def test_models_against_golden(golden, dev, precision, tol):
	for c in golden('models')['cases']:
		m = _build(c, dev, precision)
		with torch.no_grad():
			out = m(c['signal'].to(dev), c['xlen'].to(dev), y = c['y'].to(dev), ylen = c['ylen'].to(dev))
		logits, log_probs, olen = out['logits'][0], out['log_probs'][0], out['olen'][0]
		assert logits.shape == c['logits'].shape and log_probs.dtype == torch.float32
		assert torch.equal(olen.cpu(), c['olen']), c['model']
		r = rel(logits, c['logits'])
		assert r < tol, (c['model'], precision, r)
		assert rel(log_probs, c['log_probs']) < tol
		if precision == 'fp32':
			# greedy ids must agree wherever the reference's own top-2 margin is not a numerical tie
			ids = log_probs.argmax(1).cpu()
			top2 = c['log_probs'].topk(2, dim = 1).values
			clear = (top2[:, 0] - top2[:, 1]) > 1e-3
			assert torch.equal(ids[clear], c['log_probs'].argmax(1)[clear])
			assert torch.allclose(out['loss'].cpu(), c['loss'], rtol = 2e-3, atol = 1e-3), c['model']
		# the fused epilogue's argmax is the argmax of the log_probs it wrote (ties -> lowest id)
		assert torch.equal(log_probs._convasr_argmax.long(), log_probs.argmax(1))


def test_padded_frames_equal_decoder_bias(golden, dev):
	"""quirk 6: frames past olen are zeroed after every layer, so logits there are the decoder bias."""
	c = golden('models')['cases'][0]
	m = _build(c, dev, 'fp32')
	with torch.no_grad():
		out = m(c['signal'].to(dev), c['xlen'].to(dev))
	logits, olen = out['logits'][0], out['olen'][0]
	bias = m.decoder[0].bias.detach()
	b = 2
	assert olen[b] < logits.shape[-1]
	assert torch.allclose(logits[b, :, int(olen[b]):], bias[:, None].expand(-1, logits.shape[-1] - int(olen[b])), atol = 1e-6)


def test_full_width_wav2letter_against_oracle(dev):
	"""The real Wav2Letter (base_width 128, 66.5 M params) on a small batch, both tiers."""
	from convasr_b200 import models
	m = models.Wav2Letter(64, [38], frontend = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window'), dropout = 0., check_time_dim_padded = False)
	shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith('frontend.')}
	assert 66e6 < sum(p.numel() for p in m.parameters()) < 67e6  # 66.53 M
	sd = O.synth_state_dict(shapes, seed = 3)
	m.load_state_dict(sd, strict = False)
	m = m.to(dev).eval()
	g = torch.Generator().manual_seed(5)
	sig = (torch.randn(2, 24000, generator = g) * 2000).round().to(torch.int16)
	xlen = torch.tensor([1.0, 0.7])
	ref_logits, ref_lp, ref_olen = O.model_forward(sd, sig, xlen, model = 'Wav2Letter')
	for precision, tol in (('fp32', 1e-3), ('bf16', 2e-2)):
		m.set_precision(precision)
		with torch.no_grad():
			out = m(sig.to(dev), xlen.to(dev))
		assert out['logits'][0].shape == (2, 38, 153)  # t = 151 + 2 (quirk 7)
		assert torch.equal(out['olen'][0].cpu(), ref_olen[0])
		r = rel(out['logits'][0], ref_logits[0])
		assert r < tol, (precision, r)


# ---------------------------------------------------------------- CTC loss / alignment
def test_ctc_loss_and_grad_against_golden(golden, dev):
	from convasr_b200 import ctc
	for c in golden('ctc')['cases']:
		lp = c['log_probs'].to(dev).requires_grad_(True)
		loss = ctc.ctc_loss(lp, c['targets'].to(dev), c['input_lengths'].to(dev), c['target_lengths'].to(dev), blank = c['blank'])
		finite = torch.isfinite(c['loss'])
		assert torch.equal(torch.isinf(loss).cpu(), ~finite)
		assert torch.allclose(loss.cpu()[finite], c['loss'][finite], rtol = 1e-4)
		if c['grad'] is not None:
			loss.sum().backward()
			assert rel(lp.grad, c['grad']) < 1e-4
			il = c['input_lengths']
			for b in range(lp.shape[1]):  # exactly zero past the input length
				assert float(lp.grad[int(il[b]):, b].abs().max() if il[b] < lp.shape[0] else 0.0) == 0.0


def test_ctc_loss_bench_shape_against_float64_oracle(dev):
	"""C2 shape slice: t=753, C=38, L up to 225, through the [B,C,t]-permuted view the model uses."""
	from convasr_b200 import ctc
	g = torch.Generator().manual_seed(1)
	B, C, T, L = 3, 38, 753, 225
	lp_bct = torch.randn(B, C, T, generator = g).log_softmax(1)
	y = torch.randint(0, C - 1, (B, L), generator = g)
	ylen = torch.tensor([225, 150, 3])
	olen = torch.tensor([753, 600, 410])
	# (a) log-probs of a model that has learnt something: the target path is likely, nll = O(100)
	path = torch.full((B, T), C - 1)
	for b in range(B):
		pos = torch.linspace(0, int(olen[b]) - 1, int(ylen[b])).long()
		path[b, pos] = y[b, :int(ylen[b])]
	peaked = (torch.randn(B, C, T, generator = g) + 6.0 * torch.nn.functional.one_hot(path, C).permute(0, 2, 1)).log_softmax(1)
	# (b) uniform-ish random log-probs: nll = O(3000), where fp32 alpha+beta-nll carries ~ulp(3000)=2.4e-4
	#     of rounding noise in ANY fp32 implementation (ATen's included) -- looser bar, documented in DESIGN.md
	for lp_cpu, tol in ((peaked, 1e-4), (lp_bct, 1e-3)):
		nll, grad = O.ctc_loss_np(lp_cpu.permute(2, 0, 1).numpy(), y.numpy(), olen.numpy(), ylen.numpy(), C - 1)
		lp = lp_cpu.to(dev).requires_grad_(True)
		loss = ctc.ctc_loss(lp.permute(2, 0, 1), y.to(dev), olen.to(dev), ylen.to(dev), blank = C - 1)
		assert torch.allclose(loss.double().cpu(), torch.from_numpy(nll), rtol = 1e-4)
		loss.sum().backward()
		assert rel(lp.grad.permute(2, 0, 1), torch.from_numpy(grad)) < tol, tol
		assert float(lp.grad.sum(1).abs().max()) < 2e-3  # softmax already folded in: sums to 0 over C
	# forward-only call (no gradient requested) gives the same loss
	with torch.no_grad():
		loss2 = ctc.ctc_loss(lp.detach().permute(2, 0, 1), y.to(dev), olen.to(dev), ylen.to(dev), blank = C - 1)
	assert torch.equal(loss2, loss.detach())


def test_fp16_alignment_against_reference_golden(golden, dev):
	"""fp16 log_probs (ctc.py:29): the native kernel rounds to fp16 after every operation of the recursion, like the reference
	on an fp16 tensor; bit-exact against the reference's own fp16 outputs (tests/golden/ctc_fp16.pt) and the oracle"""
	from convasr_b200 import ctc
	for c in golden('ctc_fp16')['cases']:
		al = ctc.alignment(c['log_probs'].to(dev), c['targets'].to(dev), c['input_lengths'].to(dev), c['target_lengths'].to(dev), blank = c['blank'])
		assert al.dtype == torch.int64 and torch.equal(al.cpu(), c['alignment']), int((al.cpu() != c['alignment']).sum())
	g = torch.Generator().manual_seed(4)
	B, C, T, L = 5, 38, 400, 80
	lp = (torch.randn(B, C, T, generator = g) * 3).log_softmax(1).half()
	y = torch.randint(0, C - 1, (B, L), generator = g)
	ylen, olen = torch.tensor([80, 41, 17, 0, 33]), torch.tensor([400, 350, 399, 400, 170])
	ref = O.ctc_alignment(lp.permute(2, 0, 1), y, olen, ylen, C - 1)
	got = ctc.alignment(lp.to(dev).permute(2, 0, 1), y.to(dev), olen.to(dev), ylen.to(dev), blank = C - 1)
	assert torch.equal(got.cpu(), ref), int((got.cpu() != ref).sum())


def test_alignment_against_golden_and_oracle(golden, dev):
	from convasr_b200 import ctc
	for c in golden('ctc')['cases']:
		if c['alignment'] is None:
			continue
		for packed in (False, True):
			al = ctc.alignment(c['log_probs'].to(dev), c['targets'].to(dev), c['input_lengths'].to(dev), c['target_lengths'].to(dev), blank = c['blank'], pack_backpointers = packed)
			assert al.dtype == torch.int64 and torch.equal(al.cpu(), c['alignment'])
	# larger seeded case incl. the padded-batch quirk (input_lengths < T) and the [B,C,T] view
	g = torch.Generator().manual_seed(2)
	B, C, T, L = 6, 38, 300, 60
	lp = (torch.randn(B, C, T, generator = g) * 2).log_softmax(1)
	y = torch.randint(0, C - 1, (B, L), generator = g)
	ylen = torch.tensor([60, 41, 17, 1, 0, 33])
	olen = torch.tensor([300, 250, 299, 120, 300, 70])
	ref = O.ctc_alignment(lp.permute(2, 0, 1), y, olen, ylen, C - 1)
	got = ctc.alignment(lp.to(dev).permute(2, 0, 1), y.to(dev), olen.to(dev), ylen.to(dev), blank = C - 1)
	assert torch.equal(got.cpu(), ref)
	# size-independent properties: frames are increasing along the target and inside the input
	for b in range(B):
		a = got[b, :int(ylen[b])].cpu()
		assert bool((a[1:] > a[:-1]).all()) and (len(a) == 0 or int(a.max()) < int(olen[b]))


# ---------------------------------------------------------------- decoders / generator
def test_log_softmax_argmax_and_topk(dev):
	from convasr_b200 import ops
	g = torch.Generator().manual_seed(3)
	for C in (38, 5000):
		x = torch.randn(2, C, 77, generator = g) * 3
		lp, am = ops.log_softmax_argmax(x.to(dev))
		assert rel(lp, x.log_softmax(1)) < 1e-6
		assert torch.equal(am.long().cpu(), x.argmax(1))
	# exact ties resolve to the lowest index (SURVEY.md hard parts)
	x = torch.zeros(1, 9, 5)
	_, am = ops.log_softmax_argmax(x.to(dev))
	assert am.tolist() == [[0] * 5]
	x = torch.randn(3, 38, 40, generator = g)
	assert torch.equal(ops.topk_ids(x.to(dev), 4).long().cpu(), x.topk(4, dim = 1).indices)
	# backward of log_softmax
	xg = x.to(dev).requires_grad_(True)
	w = torch.randn(3, 38, 40, generator = g)
	(ops.log_softmax_dim1(xg) * w.to(dev)).sum().backward()
	xr = x.clone().requires_grad_(True)
	(xr.log_softmax(1) * w).sum().backward()
	assert rel(xg.grad, xr.grad) < 1e-5


def test_greedy_decoder_and_generator_against_golden(golden, dev):
	from convasr_b200 import decoders, transcript_generators
	g = golden('decode')
	tok = O.CharTokenizer(g['alphabet'])
	dec = decoders.GreedyDecoder()
	d = g['decode']
	assert dec.decode(d['lp1'].to(dev), [4]) == d['d1']
	assert dec.decode(d['lp2'].to(dev), K = 2) == d['d2']
	assert dec.decode(d['lp3'].to(dev), [33, 20, 1, 7]) == d['d3']
	assert dec.decode(d['lp3'].to(dev), [33, 20, 1, 7], K = 3) == d['d4']
	gen = transcript_generators.GreedyCTCGenerator()
	for c in g['generate']:
		ids = torch.tensor(c['ids'])
		B, T = ids.shape
		lp = torch.full((B, 38, T), -10.0).scatter_(1, ids[:, None, :], 0.0).to(dev)
		ts = (torch.arange(T)[None].float() * 0.02).expand(B, -1).to(dev) if c['with_ts'] else None
		tr = gen.generate(tok, lp, begin = torch.zeros(B, device = dev), end = torch.full((B, ), T * 0.02, device = dev), output_lengths = c['olens'], time_stamps = ts)
		segs = [[(float(s['begin']), float(s['end']), s['hyp']) for s in t[0]] for t in tr]
		assert segs == c['segments']


def test_entropy_reductions(dev):
	from convasr_b200 import models
	g = torch.Generator().manual_seed(4)
	lp = torch.randn(4, 38, 90, generator = g).log_softmax(1)
	lens = torch.tensor([90, 45, 1, 77])
	assert rel(models.entropy(lp.to(dev), lens.to(dev)), O.entropy(lp, lens)) < 1e-5
	assert rel(models.entropy(lp.to(dev)), O.entropy(lp)) < 1e-5
	assert rel(models.weighted_mean_entropy(lp.to(dev), lens.to(dev)), O.weighted_mean_entropy(lp, lens)) < 1e-5


# ---------------------------------------------------------------- training-mode surface
def test_training_step_runs_and_matches_oracle_loss(golden, dev):
	"""model.train(): differentiable ATen conv stack + native frontend / log_softmax / CTC."""
	from convasr_b200 import models
	c = golden('models')['cases'][0]
	m = _build(c, dev, None).train()
	out = m(c['signal'].to(dev), c['xlen'].to(dev), y = c['y'].to(dev), ylen = c['ylen'].to(dev))
	loss = (out['loss'] * c['ylen'][:, 0].to(dev)).mean()
	loss.backward()
	grads = [p.grad for p in m.parameters() if p.requires_grad]
	assert all(g is not None and torch.isfinite(g).all() for g in grads)


def test_feature_input_and_no_lengths_paths(dev):
	"""model(x[B,64,F]) without a frontend (features computed elsewhere, datasets.py:269-277) and
	xlen=None (no masks anywhere, full-length olen) -- models.py:282-326."""
	from convasr_b200 import models
	m = models.Wav2Letter(64, [38], base_width = 32, dropout = 0., check_time_dim_padded = False)
	shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
	sd = O.synth_state_dict(shapes, seed = 9)
	m.load_state_dict(sd)
	m = m.to(dev).eval()
	g = torch.Generator().manual_seed(6)
	sig = (torch.randn(2, 9600, generator = g) * 1000).round().to(torch.int16)
	feats = O.frontend_logmel(sig, None)  # [2, 64, 121], un-normalised
	ref_logits, ref_lp, ref_olen = O.conv_stack_forward(sd, O.masked_instance_norm(feats, None), None, **O.MODEL_CONFIGS['Wav2Letter'])
	for precision, tol in (('fp32', 1e-3), ('bf16', 2e-2)):
		m.set_precision(precision)
		with torch.no_grad():
			out = m(feats.to(dev))
		assert out['logits'][0].shape == ref_logits[0].shape
		assert out['olen'][0].tolist() == [ref_logits[0].shape[-1]] * 2 == ref_olen[0].tolist()
		assert rel(out['logits'][0], ref_logits[0]) < tol, precision
	# lengths given, features in: masks applied from the fractions
	xlen = torch.tensor([1.0, 0.52])
	ref2, _, olen2 = O.conv_stack_forward(sd, O.masked_instance_norm(feats, xlen), xlen, **O.MODEL_CONFIGS['Wav2Letter'])
	m.set_precision('fp32')
	with torch.no_grad():
		out2 = m(feats.to(dev), xlen.to(dev))
	assert torch.equal(out2['olen'][0].cpu(), olen2[0]) and rel(out2['logits'][0], ref2[0]) < 1e-3


def test_cuda_graph_forward_equals_eager(golden, dev):
	c = golden('models')['cases'][0]
	m = _build(c, dev, 'bf16')
	sig, xlen = c['signal'].to(dev), c['xlen'].to(dev)
	with torch.no_grad():
		eager = m(sig, xlen)
		m.enable_cuda_graphs(True)
		g1 = m(sig, xlen)
		g2 = m(sig.clone(), xlen.clone())  # replay with fresh input tensors
		other = m(sig.roll(1, 0), xlen.roll(1, 0))
	assert torch.equal(eager['logits'][0], g1['logits'][0]) and torch.equal(g1['logits'][0], g2['logits'][0])
	assert torch.equal(other['logits'][0], eager['logits'][0].roll(1, 0))
	assert torch.equal(g1['log_probs'][0]._convasr_argmax, eager['log_probs'][0]._convasr_argmax)


@pytest.mark.parametrize('C_in,C_out,groups,k', [(256, 256, 128, 11), (256, 384, 128, 13), (384, 384, 128, 13), (640, 768, 128, 25), (96, 160, 32, 17), (64, 64, 32, 5), (40, 40, 8, 19), (64, 64, 16, 31)])
def test_grouped_conv_relu_kernel(dev, C_in, C_out, groups, k):
	"""First stage of the separable ConvSamePadding (models.py:50-64): grouped Conv1d + bias + ReLU, both
	precision tiers, odd channels per group, padded row pitches, ragged tile ends; k = 31 takes the generic kernel."""
	from convasr_b200 import ops
	g = torch.Generator().manual_seed(C_in + k)
	B, T = 3, 77
	ld_in, ld_out = (C_in + 63) // 64 * 64, (C_out + 63) // 64 * 64
	x = torch.randn(B, C_in, T, generator = g)
	w = torch.randn(C_out, C_in // groups, k, generator = g) / (k * C_in / groups) ** 0.5
	bias = torch.randn(C_out, generator = g) * 0.1
	hi = x.to(torch.bfloat16)
	lo = (x - hi.float()).to(torch.bfloat16)

	def cl(t):  # [B, C, T] -> channels-last with a zero-padded pitch and two spare rows
		out = torch.zeros(B, T + 2, ld_in, dtype = torch.bfloat16)
		out[:, :T, :C_in] = t.permute(0, 2, 1)
		return out.to(dev)

	for tier, xin, tol in (('bf16', hi.float(), 5e-3), ('fp32', hi.float() + lo.float(), 2e-5)):
		ref = torch.nn.functional.conv1d(xin.double(), w.double(), bias.double(), padding = k // 2, groups = groups).relu()
		o_hi, o_lo = ops.grouped_conv1d_relu(cl(hi), T, C_in, w.to(dev), bias.to(dev), groups, k // 2, ld_out = ld_out, act_lo = cl(lo) if tier == 'fp32' else None, want_lo = tier == 'fp32')
		got = o_hi.float() + (o_lo.float() if o_lo is not None else 0)
		assert got.shape == (B, T, ld_out)
		assert rel(got[:, :, :C_out].permute(0, 2, 1), ref) < tol, (tier, rel(got[:, :, :C_out].permute(0, 2, 1), ref))
		assert float(got[:, :, C_out:].abs().max()) == 0.0 if ld_out > C_out else True


def test_device_feeder_prefetch(dev):
	"""feed.DeviceFeeder: batches arrive on the GPU unchanged and in order while the next upload is already in
	flight on the copy stream; pinned staging buffers are recycled (train.py:745)."""
	from convasr_b200 import feed
	from oracle import make_golden
	pool = feed.PinnedPool()
	name, items, multiple = make_golden.feed_batches()[0]
	want = [feed.collate(items[k:] + items[:k], multiple) for k in range(4)]  # 4 different batches (rotations)

	def host_batches():
		for k in range(4):
			yield feed.collate(items[k:] + items[:k], multiple, pool = pool)

	feeder = feed.DeviceFeeder(host_batches(), dev, pool = pool)
	n = 0
	for (meta, s, x, xlen, y, ylen), ref_b in zip(feeder, want):
		assert x.is_cuda and x.dtype == torch.int16 and xlen.dtype == torch.float32
		_ = (x.float() * 2).sum()  # some work on the consumer stream
		for got, exp in ((x, ref_b[2]), (xlen, ref_b[3]), (y, ref_b[4]), (ylen, ref_b[5])):
			assert torch.equal(got.cpu(), exp)
		assert [m['example_id'] for m in meta] == [m['example_id'] for m in ref_b[0]]
		n += 1
	assert n == 4 and feeder.bytes_uploaded == sum(t.numel() * t.element_size() for b in want for t in b[2:])
	assert sum(len(v) for v in pool.free.values()) >= 4  # staging buffers came back


def test_tiny_utterances_in_a_ragged_batch(dev):
	"""Edge case of a ragged batch: an utterance shorter than one frame hop (one valid frame; every other tile of
	it is padding and is skipped) must not disturb its neighbours and equals its own stand-alone result; padded
	frames carry the decoder bias (quirk 6)."""
	from convasr_b200 import models
	torch.manual_seed(0)
	m = models.Wav2Letter(64, [38], frontend = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window'), dropout = 0., check_time_dim_padded = False, base_width = 32)
	shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith('frontend.')}
	m.load_state_dict(O.synth_state_dict(shapes, seed = 3), strict = False)
	m = m.to(dev).eval().set_precision('fp32')
	g = torch.Generator().manual_seed(5)
	T = 24000
	sig = (torch.randn(4, T, generator = g) * 3000).round().clamp(-32767, 32767).to(torch.int16)
	xlen = torch.tensor([1.0, 30 / T, 0.37, 0.5])
	with torch.no_grad():
		out = m(sig.to(dev), xlen.to(dev))
		solo = [m(sig[k:k + 1].to(dev), xlen[k:k + 1].to(dev)) for k in (1, 2)]
	logits, olen = out['logits'][0], out['olen'][0]
	assert int(olen[1]) == 1 and bool(torch.isfinite(logits).all())
	for k, s1 in zip((1, 2), solo):
		assert rel(logits[k], s1['logits'][0][0]) < 1e-5, k
		n = int(olen[k])
		bias = m.decoder[0].bias.detach()
		assert torch.allclose(logits[k][:, n:], bias[:, None].expand(-1, logits.shape[2] - n), atol = 1e-6)


def test_greedy_collapse_warp_scan_equals_the_serial_state_machine(dev):
	"""the segmented warp scan of cab_greedy_collapse against a line-by-line transcription of the reference's loop
	(transcript_generators.py:32-83) on random id rows: blank / space / repeat densities, ragged lengths, frame counts that are
	not multiples of 32, word-start tables with extra tokens (BPE-style), tiny output capacity"""
	from convasr_b200 import ops
	g = torch.Generator().manual_seed(12)
	for C, T, B, K in ((38, 97, 300, 10), (12, 64, 200, 3), (50, 333, 100, 10), (38, 31, 50, 2)):
		eps, space = C - 1, C - 2
		pool = torch.tensor([0, 1, 2, 3, space, space] + [eps] * 14)
		ids = pool[torch.randint(0, len(pool), (B, T), generator = g)]
		ids[: B // 4] = torch.randint(0, C, (B // 4, T), generator = g)  # dense tokens, few blanks
		ids[B // 4] = eps  # nothing but silence
		lens = torch.randint(0, T + 1, (B, ), generator = g)
		lens[0] = T
		sil = torch.zeros(C, dtype = torch.uint8); sil[eps] = sil[space] = 1
		ws = torch.zeros(C, dtype = torch.uint8); ws[space] = 1; ws[1] = 1  # token 1 also starts words
		tok, frm, cnt = ops.greedy_collapse(ids.int().to(dev), lens.to(dev), C, eps, space, sil.to(dev), ws.to(dev), K)
		tok, frm, cnt = tok.cpu(), frm.cpu(), cnt.cpu()
		for b in range(B):
			row = ids[b].tolist()
			t = 0
			while t < T and sil[row[t]]:
				t += 1
			if t >= T:
				assert int(cnt[b]) == -1, b
				continue
			tokens, frames, allow_repeat, count_eps = [eps], [None], False, 0
			for t in range(t, int(lens[b])):
				x = row[t]
				if x == eps and tokens[-1] == space:
					continue
				if x == eps:
					allow_repeat = True
					count_eps += 1
					if count_eps >= K and not ws[tokens[-1]]:
						tokens.append(space); frames.append(-(t + 1))
					continue
				elif x == tokens[-1] and not allow_repeat:
					continue
				allow_repeat = False
				tokens.append(x); frames.append(t)
				count_eps = 0
			n = len(tokens) - 1
			assert int(cnt[b]) == n, (b, int(cnt[b]), n)
			assert tok[b, :n].tolist() == tokens[1:] and frm[b, :n].tolist() == frames[1:], b


# ---------------------------------------------------------------- small parity holes closed in round 2
def test_uncertainty_reductions_against_reference_golden(golden, dev):
	"""entropy / weighted_mean_entropy / margin (models.py:645-678) against outputs of the reference itself"""
	from convasr_b200 import models
	for c in golden('misc')['uncertainty']:
		lp, lens = c['log_probs'].to(dev), c['lengths'].to(dev)
		C = lp.shape[1]
		assert torch.allclose(models.entropy(lp, lens, dim = 1).cpu(), c['entropy'], rtol = 1e-5, atol = 1e-6)
		assert torch.allclose(models.entropy(lp, None, dim = 1).cpu(), c['entropy_nolen'], rtol = 1e-5, atol = 1e-6)
		assert torch.allclose(models.weighted_mean_entropy(lp, lens, dim = -2, eps_id = C - 1).cpu(), c['wme'], rtol = 1e-5, atol = 1e-6)
		assert torch.allclose(models.weighted_mean_entropy(lp, None, dim = -2, eps_id = C - 1).cpu(), c['wme_nolen'], rtol = 1e-5, atol = 1e-6)
		assert torch.allclose(models.margin(lp[:2], dim = 1).cpu(), c['margin'], rtol = 1e-5, atol = 1e-7)
		with pytest.raises(TypeError):  # the reference unpacks the batch dimension: only B == 2 is defined
			models.margin(torch.cat([lp, lp])[:3], dim = 1)
		# the permuted view the large-vocabulary head returns gives the same numbers
		lpv = lp.permute(0, 2, 1).contiguous().permute(0, 2, 1)
		assert torch.allclose(models.margin(lpv[:2]).cpu(), c['margin'], rtol = 1e-5, atol = 1e-7)


def test_two_head_bpe_decoder_against_reference_golden(golden, dev):
	"""Decoder(type = 'bpe') (models.py:27-44): head 0 = 1x1 conv, head 1 = two ConvBn1d(k = 15, relu, no mask); loss = sum of the
	heads' CTC losses, every head divided by ylen[:, 0] (models.py:323)"""
	from convasr_b200 import models
	c = golden('misc')['bpe']
	m = models.Wav2Letter(64, list(c['num_classes']), frontend = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window'), dropout = 0., check_time_dim_padded = False, **c['kwargs'])
	assert {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith('frontend.')} == c['shapes']
	m.load_state_dict(O.synth_state_dict(c['shapes'], seed = c['seed']), strict = False)
	m = m.to(dev).eval()
	for precision, tol in (('fp32', 1e-3), ('bf16', 2e-2)):
		m.set_precision(precision)
		with torch.no_grad():
			out = m(c['signal'].to(dev), c['xlen'].to(dev), y = c['y'].to(dev), ylen = c['ylen'].to(dev))
		assert len(out['logits']) == 2
		for h in range(2):
			assert torch.equal(out['olen'][h].cpu(), c['olen'][h])
			assert rel(out['logits'][h], c['logits'][h]) < tol, (precision, h, rel(out['logits'][h], c['logits'][h]))
			assert rel(out['log_probs'][h], c['log_probs'][h]) < tol
		assert torch.allclose(out['loss'].cpu(), c['loss'], rtol = 5 * tol, atol = 5 * tol)
	m.bpe_only = True
	with torch.no_grad():
		only = m(c['signal'].to(dev), c['xlen'].to(dev), y = c['y'].to(dev), ylen = c['ylen'].to(dev))
	assert float(only['loss'].sum()) < float(out['loss'].sum())


def test_bf16_tier_greedy_id_and_text_agreement_is_reported(golden, dev, capsys):
	"""bf16 tier vs the reference's fp32 outputs on the golden model cases: fraction of valid frames whose greedy id agrees, and
	the character error rate between the collapsed texts (ids are only guaranteed bit-exact on identical logits; this is the
	measured cost of the speed tier)"""
	from convasr_b200 import models, transcript_generators
	tok = O.CharTokenizer('абвгдеёжзийклмнопрстуфхцчшщъыьэюя')
	gen = transcript_generators.GreedyCTCGenerator()

	def edit(a, b):
		d = list(range(len(b) + 1))
		for i, ca in enumerate(a, 1):
			p, d[0] = d[0], i
			for j, cb in enumerate(b, 1):
				p, d[j] = d[j], min(d[j] + 1, d[j - 1] + 1, p + (ca != cb))
		return d[-1]

	lines = []
	for c in golden('models')['cases']:
		if c['num_classes'] != 38:
			continue
		m = getattr(models, c['model'])(64, [38], frontend = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window'), dropout = 0., check_time_dim_padded = False, **c['kwargs'])
		m.load_state_dict(O.synth_state_dict(c['shapes'], seed = c['seed']), strict = False)
		m = m.to(dev).eval()
		ref_ids = c['log_probs'].argmax(1)
		B = ref_ids.shape[0]
		ref_txt = O.greedy_generate(tok, ref_ids.tolist(), c['olen'].tolist(), None, [0.0] * B, [1.0] * B)
		for precision in ('fp32', 'bf16'):
			m.set_precision(precision)
			with torch.no_grad():
				out = m(c['signal'].to(dev), c['xlen'].to(dev))
			ids = out['log_probs'][0]._convasr_argmax.long().cpu()
			valid = torch.arange(ids.shape[1])[None] < c['olen'][:, None]
			agree = float((ids == ref_ids)[valid].float().mean())
			tr = gen.generate(tok, out['log_probs'][0], begin = torch.zeros(B, device = dev), end = torch.ones(B, device = dev), output_lengths = out['olen'][0])
			hyp = [' '.join(s['hyp'] for s in t[0]) for t in tr]
			ref = [' '.join(s[2] for s in t) for t in ref_txt]
			cer = sum(edit(h, r) for h, r in zip(hyp, ref)) / max(1, sum(len(r) for r in ref))
			lines.append(f'{c["model"]:20s} fused={c["fused"]} {precision}: greedy-id agreement {agree:.4f}, CER vs reference text {cer:.4f}')
			if precision == 'fp32':
				assert agree > 0.999 and cer < 1e-3, lines[-1]
			else:
				assert agree > 0.9, lines[-1]
	with capsys.disabled():
		print('\ngreedy decode agreement with the reference (random-init weights: near-uniform posteriors, worst case for argmax ties):\n  ' + '\n  '.join(lines))


def test_dropin_modules_run_a_transcribe_shaped_flow(golden, dev):
	"""`PYTHONPATH=convasr_b200/dropin`: the bare-name modules the reference's callers import (transcribe.py:27-52,140-196;
	train.py:428-435) resolve to this repo's implementation and run the caller's flow: frontend + model factory with the
	`dict =` callable and check_time_dim_padded, fuse_conv_bn_eval, forward, GreedyCTCGenerator.generate, ctc.alignment"""
	import importlib
	import os
	import sys
	root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
	dropin = os.path.join(root, 'convasr_b200', 'dropin')
	saved = {n: sys.modules.pop(n, None) for n in ('models', 'ctc', 'decoders', 'transcript_generators', 'optimizers')}
	sys.path.insert(0, dropin)
	try:
		models, ctc, decoders, transcript_generators, optimizers = [importlib.import_module(n) for n in ('models', 'ctc', 'decoders', 'transcript_generators', 'optimizers')]
		assert os.path.dirname(models.__file__) == dropin
		c = golden('models')['cases'][0]
		frontend = models.LogFilterBankFrontend(out_channels = 64, sample_rate = 8000, window_size = .02, window_stride = .01, window = 'hann_window', dither = 0.0, dither0 = 0.0, stft_mode = None)
		model = getattr(models, 'Wav2Letter')(num_input_features = 64, num_classes = [38], frontend = frontend, check_time_dim_padded = False,
												dict = lambda logits, log_probs, olen, **kwargs: (log_probs[0], logits[0], olen[0]), **c['kwargs'])  # transcribe.py:44-52
		model.load_state_dict(O.synth_state_dict(c['shapes'], seed = c['seed']), strict = False)
		model.to(dev)
		model.eval()
		model.fuse_conv_bn_eval()  # transcribe.py:56
		model, _ = models.data_parallel_and_autocast(model, opt_level = None, data_parallel = False)
		x, xlen = c['signal'].unsqueeze(1), c['xlen']
		with torch.no_grad():
			log_probs, logits, olen = model(x.squeeze(1).to(dev), xlen.to(dev))  # transcribe.py:140
		assert rel(logits, c['logits']) < 1e-3 and torch.equal(olen.cpu(), c['olen'])
		tok = O.CharTokenizer('абвгдеёжзийклмнопрстуфхцчшщъыьэюя')
		B = len(x)
		tr = transcript_generators.GreedyCTCGenerator().generate(tokenizer = tok, log_probs = log_probs, begin = torch.zeros(B, device = dev), end = torch.ones(B, device = dev), output_lengths = olen, time_stamps = None, segment_text_key = 'hyp')
		ref = O.greedy_generate(tok, log_probs.argmax(1).cpu().tolist(), olen.tolist(), None, [0.0] * B, [1.0] * B)
		assert [[s['hyp'] for s in t[0]] for t in tr] == [[s[2] for s in t] for t in ref]
		assert decoders.GreedyDecoder().decode(log_probs, olen) == O.greedy_decode(log_probs.cpu(), olen.cpu())
		y, ylen = c['y'][:, 0].to(dev), c['ylen'][:, 0].to(dev)
		al = ctc.alignment(log_probs.permute(2, 0, 1), y, olen, ylen, blank = 37, pack_backpointers = False)  # transcribe.py:176-183
		assert torch.equal(al.cpu(), O.ctc_alignment(log_probs.cpu().permute(2, 0, 1), y.cpu(), olen.cpu(), ylen.cpu(), 37))
		assert hasattr(optimizers, 'NovoGrad') and hasattr(models, 'entropy') and hasattr(models, 'margin')
	finally:
		sys.path.remove(dropin)
		for n, mod in saved.items():
			sys.modules.pop(n, None)
			if mod is not None:
				sys.modules[n] = mod


def test_dense_big_model_beyond_18_k_segments_per_launch(dev):
	"""JasperNetBig (dense residual, 2 sub-blocks per block): the last blocks join 10 branches, i.e. 30 operand pairs in the split
	tier -- more than one launch has K segments.  Eval engine and training path both chain launches (the partial sum re-enters
	as an identity segment); logits against the oracle in both tiers, and a native training step in the fp32 tier."""
	from convasr_b200 import models, training
	g = torch.Generator().manual_seed(21)
	m = models.JasperNetBig(64, [38], base_width = 16, frontend = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window'), dropout = 0., check_time_dim_padded = False)
	shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith('frontend.')}
	sd = O.synth_state_dict(shapes, seed = 31)
	m.load_state_dict(sd, strict = False)
	m = m.to(dev).eval()
	sig = (torch.randn(3, 6400, generator = g) * 3000).round().clamp(-32767, 32767).to(torch.int16)
	xlen = torch.tensor([1.0, 0.7, 0.45])
	ref_logits, _, ref_olen = O.model_forward(sd, sig, xlen, model = 'JasperNetBig')
	for precision, tol in (('fp32', 1e-3), ('bf16', 3e-2)):
		m.set_precision(precision)
		with torch.no_grad():
			out = m(sig.to(dev), xlen.to(dev))
		assert torch.equal(out['olen'][0].cpu(), ref_olen[0])
		assert rel(out['logits'][0], ref_logits[0]) < tol, (precision, rel(out['logits'][0], ref_logits[0]))
	# training, split tier: forward quantities against the oracle's train-mode forward, finite gradients for every parameter
	m.train().set_precision('fp32')
	assert training.unsupported_reason(m) is None
	y = torch.randint(0, 37, (3, 1, 6), generator = g)
	ylen = torch.tensor([[6], [4], [2]])
	out = m(sig.to(dev), xlen.to(dev), y = y.to(dev), ylen = ylen.to(dev))
	(out['loss'] * ylen[:, 0].to(dev)).mean().backward()
	ref_train, _, _ = O.model_forward(sd, sig, xlen, model = 'JasperNetBig', training = True)
	assert rel(out['logits'][0], ref_train[0]) < 1e-3, rel(out['logits'][0], ref_train[0])
	assert all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for n, p in m.named_parameters() if not n.startswith('frontend.'))
