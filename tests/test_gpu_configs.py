"""BASELINE.json configs 3-5 at their full sizes, through size-independent properties
(utterance independence, shard / micro-batch invariance, normalisation identities) plus oracle
parity on a small slice of the same configuration."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O

RU = 'абвгдеёжзийклмнопрстуфхцчшщъыьэюя'


def rel(a, b):
	a, b = a.detach().double().cpu(), b.detach().double().cpu()
	return float((a - b).norm() / (b.norm() + 1e-30))


def make_model(name, C, dev, precision = 'bf16', seed = 0, **kw):
	from convasr_b200 import models
	m = getattr(models, name)(64, [C], frontend = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window'), dropout = 0., check_time_dim_padded = False, **kw)
	shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith('frontend.')}
	sd = O.synth_state_dict(shapes, seed = seed)
	m.load_state_dict(sd, strict = False)
	return m.to(dev).eval().set_precision(precision), sd


def pcm(B, T, seed):
	g = torch.Generator().manual_seed(seed)
	sig = (torch.randn(B, T, generator = g) * 3000).round().clamp(-32767, 32767).to(torch.int16)
	xlen = torch.rand(B, generator = g) * 0.5 + 0.5
	xlen[0] = 1.0
	return sig, xlen


def test_config3_jasper_separable_B256x20s():
	dev = torch.device('cuda:0')
	m, sd = make_model('JasperNetSeparable', 38, dev)
	sig, xlen = pcm(256, 160000, 3)
	with torch.no_grad():
		out = m(sig.to(dev), xlen.to(dev))
	logits, lp, olen = out['logits'][0], out['log_probs'][0], out['olen'][0]
	assert logits.shape == (256, 38, 1001)  # dilation 1: no +2 frames
	assert torch.equal(olen.cpu(), (xlen * 1001).ceil().long())
	assert bool(torch.isfinite(logits).all())
	assert float(lp.exp().sum(1).sub(1).abs().max()) < 1e-4
	# utterance independence: the same rows in a batch of 3 give bit-identical results
	with torch.no_grad():
		small = m(sig[:3].to(dev), xlen[:3].to(dev))
	assert torch.equal(small['logits'][0], logits[:3])
	# oracle parity on that slice (bf16 tier bar)
	ref_logits, _, ref_olen = O.model_forward(sd, sig[:2], xlen[:2], model = 'JasperNetSeparable')
	assert torch.equal(ref_olen[0], olen[:2].cpu())
	assert rel(logits[:2], ref_logits[0]) < 2e-2


def test_config4_wav2letter_bpe5000_head_B64x15s():
	from convasr_b200 import ctc
	dev = torch.device('cuda:0')
	C = 5000
	m, sd = make_model('Wav2Letter', C, dev)
	sig, xlen = pcm(64, 120000, 4)
	g = torch.Generator().manual_seed(44)
	L = 60
	y = torch.randint(0, C - 1, (64, 1, L), generator = g)
	ylen = torch.randint(20, L + 1, (64, 1), generator = g)
	with torch.no_grad():
		out = m(sig.to(dev), xlen.to(dev), y = y.to(dev), ylen = ylen.to(dev))
	logits, lp, olen = out['logits'][0], out['log_probs'][0], out['olen'][0]
	assert logits.shape == (64, C, 753)
	assert float(lp.exp().sum(1).sub(1).abs().max()) < 1e-3
	assert torch.equal(lp._convasr_argmax.long(), lp.argmax(1))
	assert bool(torch.isfinite(out['loss']).all())
	with torch.no_grad():
		small = m(sig[:2].to(dev), xlen[:2].to(dev))
	assert torch.equal(small['logits'][0], logits[:2])
	ref_logits, ref_lp, ref_olen = O.model_forward(sd, sig[:2], xlen[:2], model = 'Wav2Letter')
	assert rel(logits[:2], ref_logits[0]) < 2e-2
	# CTC at C = 5000 on the GPU's own log-probs: loss vs torch CPU arithmetic, gradient identities
	lp2 = lp[:2].detach().clone().requires_grad_(True)
	nll = ctc.ctc_loss(lp2.permute(2, 0, 1), y[:2, 0].to(dev), olen[:2], ylen[:2, 0].to(dev), blank = C - 1)
	ref_nll = O.ctc_loss_torch(lp[:2].cpu().permute(2, 0, 1), y[:2, 0], olen[:2].cpu(), ylen[:2, 0], C - 1)
	assert torch.allclose(nll.cpu(), ref_nll, rtol = 1e-4)
	nll.sum().backward()
	assert float(lp2.grad.sum(1).abs().max()) < 2e-3
	for b in range(2):
		assert float(lp2.grad[b, :, int(olen[b]):].abs().max() if olen[b] < 753 else 0.0) == 0.0


def test_config5_shard_and_microbatch_invariance_1024x10s():
	"""config 5 (8192 utt x 10 s over 1-8 GPUs) scaled to one GPU's test budget: 1024 utterances;
	the property that makes sharding communication-free is that every transcript is independent of
	which shard / micro-batch its utterance lands in."""
	from convasr_b200 import parallel, transcript_generators
	dev = torch.device('cuda:0')
	m, _ = make_model('Wav2Letter', 38, dev)
	sig, xlen = pcm(1024, 80000, 5)
	sig, xlen = sig.to(dev), xlen.to(dev)
	tok = O.CharTokenizer(RU)
	gen = transcript_generators.GreedyCTCGenerator()
	full = parallel.transcribe_sharded(m, gen, tok, sig, xlen, micro_batch = 256)
	assert len(full) == 1024
	# 8 "ranks", length-bucketed shards, different micro-batch size
	shards = parallel.shard_by_length(xlen.tolist(), 8)
	merged = [None] * 1024
	for idx in shards:
		ii = torch.tensor(idx, device = dev)
		for i, hyp in zip(idx, parallel.transcribe_sharded(m, gen, tok, sig[ii], xlen[ii], micro_batch = 96)):
			merged[i] = hyp
	assert merged == full
