"""CPU oracle: a restatement of the reference's acoustic-model hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under convasr_b200/ may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg use it, and only as
the checker or as the timed CPU baseline -- never as the product path.

Parity status: the reference ships no golden vectors, known-answer tests or fixtures
(SURVEY.md section 4), so this oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, run in
the build container through oracle/reference_shim.py; the vectors are committed under
tests/golden/ together with the generator oracle/make_golden.py.  The CTC *loss* arithmetic is
not in the reference at all: it is torch.nn.functional.ctc_loss (PyTorch, unpinned by the
reference, 2.11.0 in this image; call site models.py:323) -- restated here as the textbook
log-space alpha/beta recursion in float64 and anchored on F.ctc_loss outputs.

Every function cites the reference file:line it follows (paths relative to the reference root).
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------
# lengths / masks
# --------------------------------------------------------------------------------------------


def output_lengths(T, xlen):
	"""models.py:611-614 compute_output_lengths: ceil(fp32(xlen) * T).long(); full T when xlen is None."""
	if xlen is None:
		return None
	return (xlen.to(torch.float32) * T).ceil().long()


def temporal_mask(T, lengths):
	"""models.py:617-619: arange(T) < lengths, shaped [B, 1, T]."""
	return (torch.arange(T)[None, :] < lengths[:, None]).unsqueeze(1)


# --------------------------------------------------------------------------------------------
# frontend (models.py:486-603, 684-686), closed form of SURVEY.md appendix C
# --------------------------------------------------------------------------------------------


def slaney_mel_filterbank(sample_rate, n_fft, n_mels, fmin = 0.0, fmax = None):
	"""librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax) as called at models.py:521-523
	(librosa is absent from the image; this is its published Slaney-scale, Slaney-normalised
	triangular filterbank, computed in float64 and rounded to float32)."""
	fmax = sample_rate / 2.0 if fmax is None else float(fmax)
	f_sp = 200.0 / 3
	min_log_hz = 1000.0
	min_log_mel = min_log_hz / f_sp
	logstep = math.log(6.4) / 27.0

	def hz_to_mel(f):
		f = np.asarray(f, dtype = np.float64)
		return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, f / f_sp)

	def mel_to_hz(m):
		m = np.asarray(m, dtype = np.float64)
		return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

	n_freq = 1 + n_fft // 2
	fftfreqs = np.linspace(0.0, sample_rate / 2.0, n_freq)
	mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
	fdiff = np.diff(mel_f)
	ramps = mel_f[:, None] - fftfreqs[None, :]
	weights = np.zeros((n_mels, n_freq))
	for i in range(n_mels):
		lower = -ramps[i] / fdiff[i]
		upper = ramps[i + 2] / fdiff[i + 1]
		weights[i] = np.maximum(0, np.minimum(lower, upper))
	enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
	weights *= enorm[:, None]
	return torch.from_numpy(weights.astype(np.float32))


def normalize_signal(signal, eps = 1e-5, denom_multiplier = 1.0):
	"""models.py:684-686."""
	if signal.numel() == 0:
		return signal
	signal_max = signal.abs().max(dim = -1, keepdim = True).values + eps
	return signal / (signal_max * denom_multiplier)


def frontend_logmel(
	signal, xlen = None, sample_rate = 8000, window_size = 0.02, window_stride = 0.01, n_mels = 64, preemphasis = 0.97,
	eps = 2.0**-14, normalize = True, window = 'hann_window', mel = None
):
	"""LogFilterBankFrontend.forward, models.py:565-597, with the mask of JasperNet.forward :290.

	signal [B, T] int16 or float; returns fp32 [B, n_mels, T // hop + 1]."""
	win_length = int(window_size * sample_rate)  # models.py:514
	hop = int(window_stride * sample_rate)  # :515
	nfft = 2**math.ceil(math.log2(win_length))  # :516
	n_freq = nfft // 2 + 1  # :517
	w = getattr(torch, window)(win_length, periodic = True).float()  # :519
	if mel is None:
		mel = slaney_mel_filterbank(sample_rate, nfft, n_mels, 0.0, int(sample_rate / 2))
	x = signal if signal.is_floating_point() else signal.to(torch.float32)  # :568 (no /32768)
	if normalize:
		x = normalize_signal(x)  # :570
	if preemphasis > 0:
		x = torch.cat([x[..., :1], x[..., 1:] - preemphasis * x[..., :-1]], dim = -1)  # :572
	T = x.shape[-1]
	if xlen is not None:
		x = x * (torch.arange(T)[None, :] < output_lengths(T, xlen)[:, None])  # :575 with mask of :290
	pad = n_freq - 1
	# :577-582: reflect pad left (constant if the signal is shorter than the pad), zero pad right
	if pad < T:
		left = x[..., 1:pad + 1].flip(-1)
	else:
		left = torch.zeros(x.shape[:-1] + (pad, ), dtype = x.dtype)
	padded = torch.cat([left, x, torch.zeros(x.shape[:-1] + (pad, ), dtype = x.dtype)], dim = -1)
	# :590 stft(center=False): frames of nfft samples every hop, window centred in nfft
	n_frames = (padded.shape[-1] - nfft) // hop + 1
	frames = padded.unfold(-1, nfft, hop)[..., :n_frames, :]  # [B, F, nfft]
	wpad = torch.zeros(nfft)
	off = (nfft - win_length) // 2
	wpad[off:off + win_length] = w
	spec = torch.fft.rfft(frames * wpad, dim = -1)  # [B, F, n_freq]
	power = spec.real**2 + spec.imag**2  # :592-594
	logmel = (torch.einsum('mk,bfk->bmf', mel, power) + eps).log()  # :595 (eps is the mel conv's bias)
	return logmel


def masked_instance_norm(x, xlen = None, eps = 2.0**-14):
	"""MaskedInstanceNorm1d.forward models.py:694-719 (legacy=True, affine=False), in fp32 as :301."""
	x = x.to(torch.float32)
	if xlen is None:
		mean = x.mean(dim = -1, keepdim = True)  # :704
		xm = x - mean
		std = (xm * xm).mean(dim = -1, keepdim = True).add(eps).sqrt()  # :709
		return xm / std
	mask = temporal_mask(x.shape[-1], output_lengths(x.shape[-1], xlen))
	n = mask.int().sum(dim = -1, keepdim = True)  # :715
	mean = (x * mask).sum(dim = -1, keepdim = True) / n  # :716
	z = mask * (x - mean)  # :717
	std = ((z * z).sum(dim = -1, keepdim = True) / n).add(eps).sqrt()  # :718
	return z / std  # :719


# --------------------------------------------------------------------------------------------
# conv stack (models.py:47-151, 158-326), driven by a reference-format state_dict
# --------------------------------------------------------------------------------------------

MODEL_CONFIGS = {
	# class name -> hyper-parameters that are not recoverable from tensor shapes
	# (models.py:819-1442).  `mask` = temporal_mask of the ConvBn1d blocks.
	'Wav2Letter': dict(act = ('hardtanh', 0, 20), residual = False, dilation = 2, mask = True),
	'Wav2LetterResidual': dict(act = ('hardtanh', 0, 20), residual = True, dilation = 2, mask = True),
	'Wav2LetterDense': dict(act = ('hardtanh', 0, 20), residual = 'dense', dilation = 2, mask = True),
	'JasperNet': dict(act = ('relu', ), residual = 'dense', dilation = 1, mask = True),
	'JasperNetSeparable': dict(act = ('relu', ), residual = 'dense', dilation = 1, mask = True, groups = 128),
	'JasperNetSmall': dict(act = ('relu', ), residual = 'dense', dilation = 1, mask = False),
	'JasperNetBig': dict(act = ('relu', ), residual = 'dense', dilation = 1, mask = False),
	# the *Inplace families compute BatchNorm1d + leaky_relu (models.py:376-433 only change how the backward gets its inputs)
	'Wav2LetterDenseNoDilationInplace': dict(act = ('leaky_relu', 0.01), residual = 'dense', dilation = 1, mask = True),
}


def _activation(y, act):
	if act[0] == 'hardtanh':
		return y.clamp(act[1], act[2])
	if act[0] == 'relu':
		return y.relu()
	if act[0] == 'leaky_relu':
		return F.leaky_relu(y, act[1])
	raise ValueError(act)


def _bn_eval(y, sd, prefix, eps = 1e-5, training = False, stats_out = None):
	"""nn.BatchNorm1d (models.py:112-113): eval mode uses the running statistics; training mode the
	biased batch statistics over all B*t positions (padded frames included -- the mask comes after the
	activation, models.py:135-138) and moves the running statistics with momentum 0.1 (running_var with the
	unbiased batch variance); `stats_out` receives the updated buffers.  Absent keys = already fused (Identity)."""
	if prefix + '.running_mean' not in sd:
		return y
	if training:
		rm = rv = None
		if stats_out is not None:
			rm, rv = sd[prefix + '.running_mean'].detach().clone().float(), sd[prefix + '.running_var'].detach().clone().float()
			stats_out[prefix + '.running_mean'], stats_out[prefix + '.running_var'] = rm, rv
			stats_out[prefix + '.num_batches_tracked'] = sd[prefix + '.num_batches_tracked'] + 1
		return F.batch_norm(y, rm, rv, sd[prefix + '.weight'], sd[prefix + '.bias'], True, 0.1, eps)
	g, b = sd[prefix + '.weight'], sd[prefix + '.bias']
	m, v = sd[prefix + '.running_mean'], sd[prefix + '.running_var']
	return (y - m[None, :, None]) / torch.sqrt(v[None, :, None] + eps) * g[None, :, None] + b[None, :, None]


def _count(sd, fmt):
	n = 0
	while any(k.startswith(fmt.format(n)) for k in sd):
		n += 1
	return n


def _round_bf16(t):
	"""bf16 rounding with a straight-through gradient: emulates bf16 storage of a tensor"""
	return t + (t.to(torch.bfloat16).to(t.dtype) - t).detach()


def conv_stack_forward(sd, x, xlen, act, residual, dilation, mask = True, stride1 = 2, groups = 1, num_epilogue = 2, dtype = torch.float32, training = False, round_bf16 = False, stats_out = None, frozen_blocks = 0):
	"""JasperNet.forward models.py:303-317 (backbone, decoder, log_softmax) in eval mode.

	x: normalised features [B, C, F].  Returns (logits list, log_probs list, olen list)."""
	sd = {k: v.to(dtype) if v.is_floating_point() else v for k, v in sd.items()}
	x = x.to(dtype)
	# round_bf16: the reference under bf16 mixed precision (apex O2 / autocast): conv operands and conv
	# outputs live in bf16, BatchNorm arithmetic and accumulation in fp32
	rb = _round_bf16 if round_bf16 else (lambda t: t)
	x = rb(x)
	n_blocks = _count(sd, 'backbone.{}.')
	res = []
	train_all = training
	for i in range(n_blocks):
		training = train_all and i >= frozen_blocks  # model.freeze(backbone = N) pins the BatchNorms of the first N blocks to eval (models.py:328-339)
		reps = _count(sd, 'backbone.%d.conv.{}.' % i)
		n_res = _count(sd, 'backbone.%d.conv_residual.{}.' % i) if residual else 0
		for j in range(reps):
			p = f'backbone.{i}.conv.{j}'
			w = sd[p + '.0.weight']
			k = w.shape[-1]
			stride = stride1 if i == 0 else 1
			dil = dilation if i == n_blocks - num_epilogue else 1
			pad = dil * k // 2  # models.py:49
			if p + '.2.weight' in sd:  # separable: grouped conv (+bias) -> ReLU -> pointwise (models.py:50-64)
				y = F.conv1d(x, w, sd.get(p + '.0.bias'), stride = stride, padding = pad, dilation = dil, groups = groups)
				y = F.conv1d(y.relu(), sd[p + '.2.weight'], sd.get(p + '.2.bias'))
			else:
				y = rb(F.conv1d(x, rb(w), sd.get(p + '.0.bias'), stride = stride, padding = pad, dilation = dil))
			y = _bn_eval(y, sd, f'backbone.{i}.bn.{j}', training = training, stats_out = stats_out)
			if j == reps - 1:  # residuals join on the last repeat only (models.py:129-133)
				assert n_res == len(res) or not residual
				for r, rx in enumerate(res[:n_res] if residual else []):
					pr = f'backbone.{i}.conv_residual.{r}'
					if pr + '.weight' in sd:
						ry = F.conv1d(rx, sd[pr + '.weight'], sd.get(pr + '.bias'))
						ry = _bn_eval(ry, sd, f'backbone.{i}.bn_residual.{r}', training = training, stats_out = stats_out)
					else:  # 'flat' residual: Identity (models.py:117,121)
						ry = rx
					y = y + ry
			x = _activation(y, act)
			if mask and xlen is not None:  # models.py:136-138
				x = x * temporal_mask(x.shape[-1], output_lengths(x.shape[-1], xlen))
			x = rb(x)
		# residual bookkeeping, models.py:306-313
		if i >= n_blocks - num_epilogue - 1:
			res = []
		elif residual == 'dense':
			res.append(x)
		elif residual:
			res = [x]
		else:
			res = []
	training = train_all
	logits = [F.conv1d(x, rb(sd['decoder.0.weight']), sd['decoder.0.bias'])]  # models.py:26
	if 'decoder.1.0.conv.0.0.weight' in sd:  # Decoder(type = 'bpe') models.py:27-33: two ConvBn1d(k = 15, relu, no lengths -> no mask)
		h = x
		for m in range(2):
			w = sd[f'decoder.1.{m}.conv.0.0.weight']
			h = F.conv1d(h, w, sd.get(f'decoder.1.{m}.conv.0.0.bias'), padding = w.shape[-1] // 2)
			h = _bn_eval(h, sd, f'decoder.1.{m}.bn.0', training = training, stats_out = stats_out).relu()
		logits.append(h)
	log_probs = [F.log_softmax(l, dim = 1).to(torch.float32) for l in logits]  # :316
	olen = [output_lengths(l.shape[-1], xlen) if xlen is not None else torch.full((len(l), ), l.shape[-1], dtype = torch.long) for l in logits]  # :317
	return logits, log_probs, olen


def model_forward(sd, signal, xlen, model = 'Wav2Letter', frontend = True, dtype = torch.float32, **overrides):
	"""frontend -> instance norm -> conv stack (JasperNet.forward models.py:282-326), eval mode."""
	cfg = dict(MODEL_CONFIGS[model])
	cfg.update(overrides)
	feats = frontend_logmel(signal, xlen) if frontend else signal
	feats = masked_instance_norm(feats, xlen)
	return conv_stack_forward(sd, feats, xlen, dtype = dtype, **cfg)


# --------------------------------------------------------------------------------------------
# CTC loss (call site models.py:320-324; arithmetic = torch F.ctc_loss, see module docstring)
# --------------------------------------------------------------------------------------------


def _lse(*xs):
	m = max(xs)
	if m == -math.inf:
		return -math.inf
	return m + math.log(sum(math.exp(x - m) for x in xs))


def ctc_loss_np(log_probs, targets, input_lengths, target_lengths, blank):
	"""Textbook CTC in float64.  log_probs [T, B, C] (numpy), targets [B, L].

	Returns (nll [B], grad [T, B, C]) with grad in ATen's convention exp(lp) - occupancy (for
	grad_out = 1), zero for t >= input_length -- SURVEY.md A11."""
	lp = np.asarray(log_probs, dtype = np.float64)
	T, B, C = lp.shape
	nll = np.zeros(B)
	grad = np.zeros_like(lp)
	for b in range(B):
		il, tl = int(input_lengths[b]), int(target_lengths[b])
		ext = [blank] * (2 * tl + 1)
		ext[1::2] = [int(v) for v in targets[b][:tl]]
		S = len(ext)
		alpha = np.full((il, S), -math.inf)
		beta = np.full((il, S), -math.inf)
		alpha[0, 0] = lp[0, b, blank]
		if S > 1:
			alpha[0, 1] = lp[0, b, ext[1]]
		for t in range(1, il):
			for s in range(S):
				a = [alpha[t - 1, s]]
				if s >= 1:
					a.append(alpha[t - 1, s - 1])
				if s >= 2 and ext[s] != ext[s - 2]:
					a.append(alpha[t - 1, s - 2])
				alpha[t, s] = _lse(*a) + lp[t, b, ext[s]]
		ll = _lse(alpha[il - 1, S - 1], alpha[il - 1, S - 2]) if S > 1 else alpha[il - 1, S - 1]
		nll[b] = -ll
		beta[il - 1, S - 1] = lp[il - 1, b, ext[S - 1]]
		if S > 1:
			beta[il - 1, S - 2] = lp[il - 1, b, ext[S - 2]]
		for t in range(il - 2, -1, -1):
			for s in range(S):
				a = [beta[t + 1, s]]
				if s + 1 < S:
					a.append(beta[t + 1, s + 1])
				if s + 2 < S and ext[s] != ext[s + 2]:
					a.append(beta[t + 1, s + 2])
				beta[t, s] = _lse(*a) + lp[t, b, ext[s]]
		if ll == -math.inf:
			grad[:il, b, :] = np.nan
			continue
		for t in range(il):
			occ = np.zeros(C)
			for s in range(S):
				occ[ext[s]] += math.exp(alpha[t, s] + beta[t, s] - ll - lp[t, b, ext[s]])
			grad[t, b, :] = np.exp(lp[t, b, :]) - occ
	return nll, grad


def ctc_loss_torch(log_probs, targets, input_lengths, target_lengths, blank):
	"""The third-party arithmetic itself (torch CPU), as invoked at models.py:323."""
	return F.ctc_loss(log_probs, targets, input_lengths, target_lengths, blank = blank, reduction = 'none')


# --------------------------------------------------------------------------------------------
# ctc.alignment (ctc.py:6-75)
# --------------------------------------------------------------------------------------------


def ctc_alignment(log_probs, targets, input_lengths, target_lengths, blank = 0):
	"""Restatement of ctc.alignment in float32 numpy, one utterance at a time but keeping the
	reference's batch-coupled quirks: the recursion runs over all T frames of the padded batch
	(ctc.py:47), "zero" is finfo.min (ctc.py:29), alpha accumulates with logsumexp while the
	back-pointer is the argmax of the three predecessors with stay < step < skip tie order
	(ctc.py:48-50), the terminal state is read after the last global frame (ctc.py:56-61) and the
	back-trace starts at input_length-1 (ctc.py:61).  Output: last frame index per label
	(ctc.py:72-75).  log_probs [T, B, C] torch fp32; returns int64 [B, L]."""
	half = log_probs.dtype == torch.float16
	if half:
		return _ctc_alignment_fp16(log_probs, targets, input_lengths, target_lengths, blank)
	lp = log_probs.detach().to(torch.float32).numpy()
	T, B, C = lp.shape
	L = targets.shape[1]
	zero = np.float32(np.finfo(np.float32).min)
	out = np.zeros((B, L), dtype = np.int64)
	for b in range(B):
		tl, il = int(target_lengths[b]), int(input_lengths[b])
		ext = np.full(2 * L + 1, blank, dtype = np.int64)
		ext[1::2] = targets[b].numpy()
		S = 2 * L + 1
		diff = np.zeros(S, dtype = bool)
		diff[2:] = ext[2:] != ext[:-2]  # ctc.py:23-27
		alpha = np.full(S + 2, zero, dtype = np.float32)  # two zero-padding cells in front, ctc.py:30-31
		alpha[2] = lp[0, b, blank]
		if S > 1:
			alpha[3] = lp[0, b, ext[1]]
		bp = np.zeros((T, S + 2), dtype = np.uint8)
		for t in range(1, T):
			prev = np.stack([alpha[2:], alpha[1:-1], np.where(diff, alpha[:-2], zero)])  # ctc.py:48
			m = prev.max(axis = 0)
			ms = np.where(np.isinf(m), np.float32(0), m)
			lse = np.log(np.exp(prev - ms).sum(axis = 0, dtype = np.float32)) + ms  # torch.logsumexp
			bp[t, 2:] = prev.argmax(axis = 0)  # first max wins
			alpha = np.concatenate([alpha[:2], (lp[t, b, ext] + lse).astype(np.float32)])
		l1l2 = alpha[[2 + 2 * tl - 1, 2 + 2 * tl]]
		path = np.zeros(T, dtype = np.int64)
		path[il - 1] = 2 + 2 * tl - 1 + int(l1l2.argmax())
		for t in range(T - 1, 0, -1):
			idx = path[t]
			path[t - 1] += idx - int(bp[t, idx])
		full = np.zeros(S, dtype = np.int64)
		for t in range(T):  # scatter_: the last write wins
			full[max(path[t] - 2, 0)] = t
		out[b] = full[1::2]
	return torch.from_numpy(out)


def _ctc_alignment_fp16(log_probs, targets, input_lengths, target_lengths, blank):
	"""ctc.alignment on fp16 log_probs (ctc.py:29 picks finfo(float16).min as "zero"): every tensor of the reference's
	recursion is fp16, so every operation rounds to fp16 -- torch.logsumexp is the composite amax / sub / exp / sum / log /
	add, each with an fp16 result (the 3-term sum accumulates in fp32 and rounds once); checked against the reference on
	CPU (tests/test_oracle_golden.py).  Same structure as the fp32 restatement above."""
	f16, f32 = np.float16, np.float32
	lp = log_probs.detach().numpy().astype(f16)
	T, B, C = lp.shape
	L = targets.shape[1]
	zero = f16(np.finfo(f16).min)
	out = np.zeros((B, L), dtype = np.int64)
	for b in range(B):
		tl, il = int(target_lengths[b]), int(input_lengths[b])
		ext = np.full(2 * L + 1, blank, dtype = np.int64)
		ext[1::2] = targets[b].numpy()
		S = 2 * L + 1
		diff = np.zeros(S, dtype = bool)
		diff[2:] = ext[2:] != ext[:-2]
		alpha = np.full(S + 2, zero, dtype = f16)
		alpha[2] = lp[0, b, blank]
		if S > 1:
			alpha[3] = lp[0, b, ext[1]]
		bp = np.zeros((T, S + 2), dtype = np.uint8)
		for t in range(1, T):
			prev = np.stack([alpha[2:], alpha[1:-1], np.where(diff, alpha[:-2], zero)])
			m = prev.max(axis = 0)
			ms = np.where(np.isinf(m), f16(0), m)
			np.seterr(over = 'ignore')  # zero + log-prob may leave the fp16 range: -inf, as in the reference
			d = (prev.astype(f32) - ms.astype(f32)).astype(f16)
			e = np.exp(d.astype(f32)).astype(f16)
			ssum = ((e[0].astype(f32) + e[1].astype(f32)) + e[2].astype(f32)).astype(f16)
			lse = (np.log(ssum.astype(f32)).astype(f16).astype(f32) + ms.astype(f32)).astype(f16)
			bp[t, 2:] = prev.argmax(axis = 0)
			alpha = np.concatenate([alpha[:2], (lp[t, b, ext].astype(f32) + lse.astype(f32)).astype(f16)])
		l1l2 = alpha[[2 + 2 * tl - 1, 2 + 2 * tl]]
		path = np.zeros(T, dtype = np.int64)
		path[il - 1] = 2 + 2 * tl - 1 + int(l1l2.argmax())
		for t in range(T - 1, 0, -1):
			idx = path[t]
			path[t - 1] += idx - int(bp[t, idx])
		full = np.zeros(S, dtype = np.int64)
		for t in range(T):
			full[max(path[t] - 2, 0)] = t
		out[b] = full[1::2]
	return torch.from_numpy(out)


# --------------------------------------------------------------------------------------------
# decoders.GreedyDecoder (decoders.py:5-16) and GreedyCTCGenerator (transcript_generators.py:8-93)
# --------------------------------------------------------------------------------------------


def greedy_decode(log_probs, output_lengths = None, K = 1):
	"""decoders.py:5-16: per-frame top-K ids, truncated to the output length, NOT collapsed."""
	B, C, T = log_probs.shape
	lens = [T] * B if output_lengths is None else [int(v) for v in output_lengths]
	idx = log_probs.topk(K, dim = 1).indices
	return [(l[:, :o] if K > 1 else l[0, :o]).tolist() for o, l in zip(lens, idx)]


class CharTokenizer:
	"""text_tokenizers.py:7-51 CharTokenizerLegacy: alphabet + ['*', '.', '2', ' ', '|']."""
	RU = 'абвгдежзийклмнопрстуфхцчшщъыьэюя'  # 33 letters minus 'ё' = 32?  see make(): alphabet is passed in

	def __init__(self, alphabet):
		self.idx2char = list(alphabet) + ['*', '.', '2', ' ', '|']
		self.char2idx = {c: i for i, c in enumerate(self.idx2char)}
		self.space_id = self.char2idx[' ']
		self.eps_id = self.char2idx['|']
		self.silence_tokens_ids = {self.eps_id, self.space_id}

	def is_start_word_token(self, idx):
		return idx == self.space_id

	def decode(self, tokens):
		return [''.join(self.idx2char[i] for i in t) for t in tokens]


def greedy_generate(tokenizer, ids, output_lengths = None, time_stamps = None, begin = None, end = None, blank_amount_to_space = 10):
	"""GreedyCTCGenerator.generate transcript_generators.py:27-92 on precomputed argmax ids.

	ids: list of per-utterance id lists.  Returns per utterance a list of segments
	(begin, end, text) exactly as the reference builds them."""
	out = []
	for i, row in enumerate(ids):
		n = int(output_lengths[i]) if output_lengths is not None else len(row)
		ts = time_stamps[i] if time_stamps is not None else None
		b0 = (max(float(begin[i]), 0.0) if ts is not None else float(begin[i])) if begin is not None else 0.0
		e0 = float(end[i]) if end is not None else 0.0
		segs = []
		t = 0
		while t < len(row) and row[t] in tokenizer.silence_tokens_ids:  # :38-40 scans the FULL row
			t += 1
		if t >= len(row):
			out.append(segs)
			continue
		tokens = [tokenizer.eps_id]
		time_begin = b0 + ts[t] if ts is not None else b0
		time_end = e0
		allow_repeat, count_eps = False, 0
		for t in range(t, n):
			x = row[t]
			if x == tokenizer.eps_id and tokens[-1] == tokenizer.space_id:
				continue
			if x == tokenizer.eps_id:
				allow_repeat = True
				count_eps += 1
				if count_eps >= blank_amount_to_space and not tokenizer.is_start_word_token(tokens[-1]):
					tokens.append(tokenizer.space_id)
				continue
			elif x == tokens[-1] and not allow_repeat:
				continue
			if tokenizer.is_start_word_token(x) and ts is not None:
				segs.append((time_begin, time_end, tokenizer.decode([tokens[1:]])[0]))
				tokens = [tokenizer.eps_id, x]
				time_begin = b0 + ts[t]
			allow_repeat = False
			tokens.append(x)
			time_end = b0 + ts[t] if ts is not None else e0
			count_eps = 0
		if len(tokens) > 1:
			segs.append((time_begin, time_end, tokenizer.decode([tokens[1:]])[0]))
		out.append(segs)
	return out


# --------------------------------------------------------------------------------------------
# uncertainty reductions (models.py:645-673)
# --------------------------------------------------------------------------------------------


def entropy(log_probs, lengths = None, eps = 1e-9):
	"""models.py:645-658 (dim = 1, sum = True)"""
	e = -(log_probs.exp() * log_probs).sum(dim = 1)
	if lengths is None:
		return e.mean(dim = -1)
	e = e * (torch.arange(e.shape[-1])[None] < lengths[:, None])
	return e.sum(dim = -1) / (eps + lengths.type_as(log_probs))


def margin(log_probs):
	"""models.py:676-677: `torch.sub(*probs.topk(2, dim = 1).values)` unpacks the BATCH dimension, so it is only
	defined for B == 2 and returns top2(utt 0) - top2(utt 1) [2, T]; any other B raises TypeError.  Kept as is."""
	vals = log_probs.exp().topk(2, dim = 1).values
	if vals.shape[0] != 2:
		raise TypeError(f'sub() takes 2 positional arguments but {vals.shape[0]} were given')
	return vals[0] - vals[1]


def weighted_mean_entropy(log_probs, lengths = None, eps = 1e-9, eps_id = -1):
	prob = log_probs.exp()
	e = -(prob * log_probs).sum(dim = 1)
	w = 1 - prob[:, eps_id]
	if lengths is not None:
		w = w * (torch.arange(e.shape[-1])[None] < lengths[:, None])
	return (e * w).sum(dim = -1) / (eps + w.sum(dim = -1))


# --------------------------------------------------------------------------------------------
# deterministic synthetic weights (shared by make_golden.py, tests and bench.py)
# --------------------------------------------------------------------------------------------


def synth_state_dict(shapes, seed = 0):
	"""Fill a {key: shape} dict with seeded values: conv weights ~ N(0, 1/fan_in) scaled, BN
	statistics randomised as SURVEY.md section 8(d) recommends so that folding is non-trivial."""
	g = torch.Generator().manual_seed(seed)
	sd = {}
	for k, shape in shapes.items():
		shape = tuple(shape)
		if k.endswith('num_batches_tracked'):
			sd[k] = torch.tensor(100, dtype = torch.long)
		elif k.endswith('running_var'):
			sd[k] = torch.rand(shape, generator = g) + 0.5
		elif k.endswith('running_mean'):
			sd[k] = torch.randn(shape, generator = g) * 0.1
		elif '.bn' in k and k.endswith('.weight'):
			sd[k] = torch.rand(shape, generator = g) + 0.5
		elif k.endswith('.bias'):
			sd[k] = torch.randn(shape, generator = g) * 0.1
		elif k.endswith('.weight') and len(shape) == 3:
			fan_in = shape[1] * shape[2]
			sd[k] = torch.randn(shape, generator = g) * (2.0 / fan_in)**0.5
		else:
			sd[k] = torch.randn(shape, generator = g)
	return sd


# Exactness-harness nonlinearity: hardtanh with both kinks far outside the reachable range, so that no activation gate
# can flip between two implementations.  Every nonlinearity of the reference has a kink; in TRAIN mode a gate that flips
# changes the gradient by O(1 / sqrt(elements)) -- measured here: the fp32 CPU oracle vs the reference itself differ by
# 1.7e-2 in the full-depth Wav2Letter gradients (3 flips out of 59 k elements at conv layer 15) while every forward
# quantity agrees to 3e-5.  With this nonlinearity (a legal constructor argument, models.py:825) the
# conv -> batch-statistics BatchNorm -> mask chain is smooth and kernel-chain exactness at full depth can be asserted at 1e-3.
SMOOTH_NONLINEARITY = ('hardtanh', -1000.0, 1000.0)


# --------------------------------------------------------------------------------------------
# optimizer step (section 8f "next" #3): optimizers.NovoGrad.step, optimizers.py:72-90
# --------------------------------------------------------------------------------------------


def novograd_step(params, grads, state, lr = 1.0, betas = (0.95, 0.98), eps = 1e-8, weight_decay = 0.0, dampening = False):
	"""One NovoGrad step on lists of tensors; `state` is a list of dicts that persists across calls."""
	for p, g, st in zip(params, grads, state):
		g_2 = (g**2).sum()  # :77
		st['_grads_ema'] = g_2 if '_grads_ema' not in st else st['_grads_ema'] * betas[1] + g_2 * (1. - betas[1])  # :78-79
		grad = g / (st['_grads_ema'] + eps).sqrt()  # :81
		if weight_decay > 0:
			grad = grad + weight_decay * p  # :82-83
		if dampening:
			grad = grad * (1 - betas[0])  # :84-85
		st['momentum_buffer'] = st['momentum_buffer'] * betas[0] + grad if 'momentum_buffer' in st else grad  # :87-89
		p.sub_(lr * st['momentum_buffer'])  # :90


def clip_grad_norm(grads, max_norm):
	"""torch.nn.utils.clip_grad_norm_ as called at train.py:776-779: returns (scaled grads, total norm)."""
	total = torch.sqrt(sum((g.double()**2).sum() for g in grads)).float()
	coef = torch.clamp(max_norm / (total + 1e-6), max = 1.0)
	return [g * coef for g in grads], total
