"""Native training step of the conv stack (forward with batch-statistics BatchNorm + backward).

What `model.train(); loss.backward()` does in the reference through autograd + cuDNN/ATen
(ConvBn1d.forward models.py:127-139, train.py:748-774) runs here on this repo's kernels:

  step start                 bf16 operand copies of all weights cab_pack_weights_batched (one launch)
  forward, per conv repeat   y = conv(x) + batch sums in the epilogue cab_conv1d_fused (tcgen05 implicit GEMM)
                             mean / invstd, running stats    cab_bn_finalize
                             x' = dropout(act(BN(y))) * mask cab_bn_act_mask_fwd (bulk-copy streamed)
  decoder                    logits, log_probs               cab_conv1d_fused   (LOGSOFTMAX epilogue)
  backward, per conv repeat  dz -> (dgamma, dbeta), dy       cab_bn_act_mask_bwd
                             dW = dy (x) x over time         cab_conv1d_wgrad   (tcgen05, MN-major operands)
                             dx = conv(dy, flipped W^T)      cab_conv1d_fused   (same kernel as forward)
                             all-reduce(dW) on the comm stream parallel.GradSync (data-parallel replicas)

On ragged batches the three GEMMs leave out tiles that lie entirely in an utterance's padding (the
previous layer's mask makes the inputs exact zeros there; masked gradient rows are never read).

Activations are bf16 channels-last, parameters stay fp32 masters (bf16 operand copies are re-packed
every step), all gradients are fp32.  Supported topologies: dense (non-separable)
blocks without residual branches -- the Wav2Letter family, with or without dropout (dropout masks come
from a counter-based generator of this repo, not from torch's Philox stream); anything else keeps using
the ATen path in models.JasperNet._forward_training.
"""
import torch
import torch.nn as nn

from . import _lib, engine, ops

BF16 = torch.bfloat16
_SEPARATE_BN_STATS = __import__('os').environ.get('CONVASR_B200_SEPARATE_BN_STATS', '0') == '1'
_SKIP_PADDING = __import__('os').environ.get('CONVASR_B200_SKIP_PADDING', '1') == '1'  # A/B switch: leave tiles of pure padding out


def supported(model):
	"""can this module tree train on the native kernels?"""
	if getattr(model.decoder, 'type', None) is not None:
		return False
	if model.decoder[0].out_channels > 256:
		return False
	for i, block in enumerate(model.backbone):
		if len(block.conv_residual) > 0 or block.activation.invertible:
			return False
		if block.activation.nonlinearity[0] not in ('relu', 'hardtanh', 'leaky_relu'):
			return False
		for seq, bn in zip(block.conv, block.bn):
			if len(seq) != 1 or not isinstance(bn, nn.BatchNorm1d) or type(bn) is not nn.BatchNorm1d:
				return False
			conv = seq[0]
			if conv.groups != 1 or conv.bias is not None or conv.stride[0] not in (1, 2) or (conv.stride[0] == 2 and (i != 0 or conv.dilation[0] != 1)):
				return False
			if bn.weight is None or not bn.track_running_stats or bn.momentum is None:
				return False
	return all(p.dtype == torch.float32 for p in model.parameters())


class _Layer:
	"""static description of one conv + BN + activation repeat"""

	def __init__(self, conv, bn, act, mask, dropout):
		self.conv, self.bn, self.dropout = conv, bn, float(dropout)
		self.k, self.stride, self.dil, self.pad = conv.kernel_size[0], conv.stride[0], conv.dilation[0], conv.padding[0]
		self.C_in, self.C_out = conv.in_channels, conv.out_channels
		self.ci_alloc, self.co_alloc = engine._ceil_to(self.C_in, 64), engine._ceil_to(self.C_out, 64)
		self.act = engine.act_code(act)
		self.mask = mask


def _layers(model):
	out = []
	for block in model.backbone:
		for seq, bn in zip(block.conv, block.bn):
			out.append(_Layer(seq[0], bn, block.activation.nonlinearity, bool(block.temporal_mask), block.activation.dropout))
	return out


def _pack(w, ci_ld, co_ld, want_dgrad):
	Co, Ci, K = w.shape
	fwd = torch.zeros(K, Co, ci_ld, dtype = BF16, device = w.device) if ci_ld != Ci else torch.empty(K, Co, ci_ld, dtype = BF16, device = w.device)
	dgr = None
	if want_dgrad:
		dgr = torch.zeros(K, Ci, co_ld, dtype = BF16, device = w.device) if co_ld != Co else torch.empty(K, Ci, co_ld, dtype = BF16, device = w.device)
	rc = _lib.load().cab_pack_weight(ops._p(w), Co, Ci, K, ops._p(fwd), ci_ld, ops._p(dgr), co_ld, ops._stream())
	_lib.check(rc, 'cab_pack_weight')
	return fwd, dgr


def _pack_all(specs):
	"""specs: [(w fp32 [Co, Ci, K], ci_ld, co_ld, want_dgrad)] -> [(fwd bf16 [K, Co, ci_ld], dgrad bf16 [K, Ci, co_ld] or None)]
	in ONE launch (cab_pack_weights_batched)"""
	out, items = [], (_lib.PackItem * len(specs))()
	for i, (w, ci_ld, co_ld, want_dgrad) in enumerate(specs):
		Co, Ci, K = w.shape
		fwd = torch.zeros(K, Co, ci_ld, dtype = BF16, device = w.device) if ci_ld != Ci else torch.empty(K, Co, ci_ld, dtype = BF16, device = w.device)
		dgr = None
		if want_dgrad:
			dgr = torch.zeros(K, Ci, co_ld, dtype = BF16, device = w.device) if co_ld != Co else torch.empty(K, Ci, co_ld, dtype = BF16, device = w.device)
		items[i] = _lib.PackItem(w.data_ptr(), fwd.data_ptr(), dgr.data_ptr() if dgr is not None else None, Co, Ci, K, ci_ld, co_ld)
		out.append((fwd, dgr))
	_lib.check(_lib.load().cab_pack_weights_batched(items, len(specs), ops._stream()), 'cab_pack_weights_batched')
	return out


def _bn_finalize(sums, n_rows, layer):
	"""sums: fp32 [2, co_alloc] accumulated by the conv epilogue -> [scale, shift, mean, invstd] + running stats"""
	bn = layer.bn
	ss = torch.empty(4, layer.C_out, dtype = torch.float32, device = sums.device)
	# the epilogue indexes the statistics with the padded channel count; BatchNorm sees the real channels
	if sums.shape[1] != layer.C_out:
		sums = sums[:, :layer.C_out].contiguous()
	rc = _lib.load().cab_bn_finalize(
		ops._p(sums), n_rows, layer.C_out, ops._p(bn.weight), ops._p(bn.bias), float(bn.eps), float(bn.momentum),
		ops._p(bn.running_mean), ops._p(bn.running_var), ops._p(ss), ops._stream()
	)
	_lib.check(rc, 'cab_bn_finalize')
	return ss


class NativeStack(torch.autograd.Function):
	"""feats (bf16 channels-last, no grad) -> (logits, log_probs); parameters enter as explicit inputs
	so autograd routes their gradients; `holder` carries the module tree and launch geometry."""

	@staticmethod
	def forward(ctx, holder, feats, xlen, *params):
		model = holder['model']
		layers = holder['layers']
		lib = _lib.load()
		B = feats.shape[0]
		T = holder['n_frames']
		x, x_T = feats, T
		saved = []
		# bf16 operand copies of every stride-1 conv weight and of the decoder, in one launch
		dec = model.decoder[0]
		plain = [li for li, L in enumerate(layers) if L.stride != 2]
		packed_w = _pack_all([(layers[li].conv.weight.detach(), layers[li].ci_alloc, layers[li].co_alloc, li > 0) for li in plain] + [(dec.weight.detach(), engine._ceil_to(dec.in_channels, 64), engine._ceil_to(dec.out_channels, 64), True)])
		packed_w, (w_dec, w_dec_dgr) = dict(zip(plain, packed_w[:-1])), packed_w[-1]
		for li, L in enumerate(layers):
			w = L.conv.weight
			if L.stride == 2:
				wp, taps, pad_left = engine.pack_taps_stride2(w.detach(), L.pad, L.ci_alloc)
				w_fwd, w_dgr = wp.to(BF16).contiguous(), None
				src_view = x.view(B, x.shape[1] // 2, 2 * x.shape[2])
				src = ops.Source(src_view, w_fwd, 2 * L.ci_alloc, taps, 1, pad_left, T_in = src_view.shape[1])
				T_out = (x_T + 2 * L.pad - (L.k - 1) - 1) // 2 + 1
				geom = ('pair', taps, pad_left)
				skip = None
			else:
				skip = None
				w_fwd, w_dgr = packed_w[li]
				src = ops.Source(x, w_fwd, L.ci_alloc, L.k, L.dil, L.pad, T_in = x_T)
				T_out = x_T + 2 * L.pad - L.dil * (L.k - 1)
				geom = ('plain', L.k, L.pad)
				if _SKIP_PADDING and xlen is not None and li > 0 and layers[li - 1].mask:
					# the input is exactly zero from frame ceil(xlen*x_T) on (the previous layer's mask), the conv has no
					# bias: output rows >= that + pad are zeros -- tiles of pure padding are not computed
					skip = (xlen, x_T, L.pad)
			y = torch.empty(B, T_out, L.co_alloc, dtype = BF16, device = x.device)
			if _SEPARATE_BN_STATS:  # A/B switch: statistics as a separate pass over y
				ops.conv1d_fused([src], B, T_out, L.co_alloc, out_hi = y, skip = skip)
				ws = torch.empty(2, L.C_out, dtype = torch.float32, device = x.device)
				ss = torch.empty(4, L.C_out, dtype = torch.float32, device = x.device)
				_lib.check(lib.cab_bn_batch_stats(ops._p(y), B, T_out, L.C_out, L.co_alloc, ops._p(L.bn.weight), ops._p(L.bn.bias), float(L.bn.eps), float(L.bn.momentum), ops._p(L.bn.running_mean), ops._p(L.bn.running_var), ops._p(ws), ops._p(ss), ops._stream()), 'cab_bn_batch_stats')
			else:
				sums = torch.empty(2, L.co_alloc, dtype = torch.float32, device = x.device)
				ops.conv1d_fused([src], B, T_out, L.co_alloc, out_hi = y, stats = sums, skip = skip)  # batch statistics in the epilogue
				ss = _bn_finalize(sums, B * T_out, L)
			out = torch.empty_like(y)
			code, a, b = L.act
			rc = lib.cab_bn_act_mask_fwd(ops._p(y), ops._p(ss), B, T_out, L.C_out, L.co_alloc, code, a, b, ops._p(xlen if L.mask else None), ops._p(out), L.dropout, ops._p(holder['seed']), li, ops._stream())
			_lib.check(rc, 'cab_bn_act_mask_fwd')
			saved.append((x, x_T, y, T_out, ss, w_dgr, geom))
			x, x_T = out, T_out
		torch._foreach_add_([L.bn.num_batches_tracked for L in layers], 1)
		C = dec.out_channels
		logits = torch.empty(B, C, x_T, dtype = torch.float32, device = x.device)
		log_probs = torch.empty_like(logits)
		argmax = torch.empty(B, x_T, dtype = torch.int32, device = x.device)
		ops.conv1d_fused([ops.Source(x, w_dec, w_dec.shape[2], 1, 1, 0, T_in = x_T)], B, x_T, C, bias = dec.bias.detach() if dec.bias is not None else None, logits = logits, log_probs = log_probs, argmax = argmax, epilogue = _lib.EPI_LOGSOFTMAX)
		ctx.holder, ctx.saved, ctx.xlen = holder, saved, xlen
		ctx.seed = holder['seed'].clone()  # the value this forward used (the live counter advances every step)
		holder['seed'].add_(len(layers) + 1)
		ctx.x_last, ctx.T_last, ctx.w_dec_dgr = x, x_T, w_dec_dgr
		ctx.save_for_backward(log_probs)
		ctx.mark_non_differentiable(argmax)
		return logits, log_probs, argmax

	@staticmethod
	def backward(ctx, g_logits, g_log_probs, _g_argmax):
		holder, saved, xlen = ctx.holder, ctx.saved, ctx.xlen
		model, layers = holder['model'], holder['layers']
		log_probs, = ctx.saved_tensors
		lib = _lib.load()
		B, C, T = log_probs.shape
		dev = log_probs.device
		# gradient w.r.t. the logits
		g = None
		if g_log_probs is not None:
			g = torch.empty_like(log_probs)
			rc = lib.cab_log_softmax_bwd(ops._p(log_probs), ops._p(g_log_probs.to(torch.float32).contiguous()), B, C, T, ops._p(g), ops._stream())
			_lib.check(rc, 'cab_log_softmax_bwd')
		if g_logits is not None:
			g = g_logits.to(torch.float32).contiguous() if g is None else g + g_logits
		dec = model.decoder[0]
		c_ld = engine._ceil_to(C, 64)
		g_cl = torch.empty(B, T, c_ld, dtype = BF16, device = dev)
		# all the small gradients (decoder bias, BN gamma / beta) live in one flat buffer: one all-reduce
		sync = getattr(model, '_grad_sync', None)
		small = torch.zeros(c_ld + sum(2 * L.C_out for L in layers), dtype = torch.float32, device = dev)
		small_off = c_ld
		partials = torch.empty(_lib.BN_SUM_REPLICAS * 2 * max(L.C_out for L in layers), dtype = torch.float32, device = dev)  # scratch of the BN backward sums
		d_bias = small[:C] if dec.bias is not None else None
		rc = lib.cab_bct_to_btc(ops._p(g), B, C, T, c_ld, ops._p(g_cl), ops._p(d_bias), ops._stream())
		_lib.check(rc, 'cab_bct_to_btc')
		grads = {}
		# decoder: the wide side (input channels) sits on the 128-row M side -> packed gradient is [1, Ci, Co]
		x_last, T_last = ctx.x_last, ctx.T_last
		masked_in = lambda li: _SKIP_PADDING and xlen is not None and (layers[li - 1].mask if li > 0 else False)  # is layer li's input zero past ceil(xlen*T)?
		skip_last = (xlen, T_last, 0) if _SKIP_PADDING and xlen is not None and layers[-1].mask else None
		packed = ops.conv1d_wgrad(x_last, T_last, dec.in_channels, g_cl, T, C, 1, 1, 0, skip = skip_last)
		grads[dec.weight] = _unpack(packed, 1, C, dec.in_channels, transposed = True)
		if sync is not None:
			sync.reduce(grads[dec.weight])
		if dec.bias is not None:
			grads[dec.bias] = d_bias
		ci_alloc = x_last.shape[2]
		gx = torch.empty(B, T_last, ci_alloc, dtype = BF16, device = dev)
		# gradient rows of masked frames are never read (the mask's backward selects, it does not multiply)
		ops.conv1d_fused([ops.Source(g_cl, ctx.w_dec_dgr, c_ld, 1, 1, 0, T_in = T)], B, T_last, ci_alloc, out_hi = gx, skip = skip_last)

		for li in range(len(layers) - 1, -1, -1):
			L = layers[li]
			x, x_T, y, T_out, ss, w_dgr, geom = saved[li]
			code, a, b = L.act
			sums = small[small_off:small_off + 2 * L.C_out].view(2, L.C_out)
			small_off += 2 * L.C_out
			dy = torch.empty_like(y)
			rc = lib.cab_bn_act_mask_bwd(ops._p(y), ops._p(gx), ops._p(ss), B, T_out, L.C_out, L.co_alloc, code, a, b, ops._p(xlen if L.mask else None), ops._p(sums), ops._p(dy), L.dropout, ops._p(ctx.seed), li, ops._p(partials), ops._stream())
			_lib.check(rc, 'cab_bn_act_mask_bwd')
			grads[L.bn.bias] = sums[0]
			grads[L.bn.weight] = sums[1]
			if geom[0] == 'pair':
				_, taps, pad_left = geom
				xv = x.view(B, x.shape[1] // 2, 2 * x.shape[2])
				packed = ops.conv1d_wgrad(dy, T_out, L.C_out, xv, xv.shape[1], 2 * L.ci_alloc, taps, 1, pad_left)
				grads[L.conv.weight] = _unpack_stride2(packed, L)
			else:
				grads[L.conv.weight] = _wgrad(dy, T_out, L.C_out, x, x_T, L.C_in, L.k, L.dil, L.pad, xlen if masked_in(li) else None)
			if sync is not None:
				sync.reduce(grads[L.conv.weight])  # overlaps the dgrad / wgrad of the layers still to come
			if li > 0:
				gx = torch.empty(B, x_T, L.ci_alloc, dtype = BF16, device = dev)
				# dx[j] = sum_k' W'[k'] dy[j + k'*d - (d*(K-1) - pad)], W' = flipped, transposed weights
				ops.conv1d_fused([ops.Source(dy, w_dgr, L.co_alloc, L.k, L.dil, L.dil * (L.k - 1) - L.pad, T_in = T_out)], B, x_T, L.ci_alloc, out_hi = gx, skip = (xlen, x_T, 0) if masked_in(li) else None)
		if sync is not None:
			sync.reduce(small)
			sync.finish()
		return (None, None, None) + tuple(grads.get(p) for p in holder['params'])


def _padded_work(M, N):
	n_nt = (N + 255) // 256
	bn = ((N + n_nt - 1) // n_nt + 63) // 64 * 64
	return ((M + 127) // 128 * 128) * n_nt * bn


def _wgrad(dy, T_out, C_out, x, x_T, C_in, k, dil, pad, xlen_zero = None):
	"""dW[co, ci, tap] = sum_{b,t} dy[b,t,co] * x[b, t + tap*dil - pad, ci].  Either tensor can sit on the
	128-row M side of the GEMM; pick the orientation with less tile padding (e.g. 640 -> 768 wastes 20 %
	one way and nothing the other way).  Swapping sides negates the frame shift."""
	# xlen_zero: x is exactly zero from frame ceil(xlen*x_T) on, so products with t + tap*dil - pad >= that vanish
	if _padded_work(C_out, C_in) <= _padded_work(C_in, C_out):
		packed = ops.conv1d_wgrad(dy, T_out, C_out, x, x_T, C_in, k, dil, pad, skip = (xlen_zero, x_T, pad) if xlen_zero is not None else None)
		return _unpack(packed, k, C_out, C_in, transposed = False)
	packed = ops.conv1d_wgrad(x, x_T, C_in, dy, T_out, C_out, k, -dil, -pad, skip = (xlen_zero, x_T, 0) if xlen_zero is not None else None)
	return _unpack(packed, k, C_out, C_in, transposed = True)


def _unpack(packed, K, Co, Ci, transposed):
	grad = torch.empty(Co, Ci, K, dtype = torch.float32, device = packed.device)
	rc = _lib.load().cab_unpack_wgrad(ops._p(packed), K, Co, Ci, packed.shape[2], int(transposed), ops._p(grad), 0, ops._stream())
	_lib.check(rc, 'cab_unpack_wgrad')
	return grad


def _unpack_stride2(packed, L):
	"""inverse of engine.pack_taps_stride2 for the gradient: [taps, Co, 2*ci_alloc] -> [Co, Ci, K]"""
	dp_min = (0 - L.pad) // 2
	grad = torch.empty(L.C_out, L.C_in, L.k, dtype = torch.float32, device = packed.device)
	for kk in range(L.k):
		j = kk - L.pad
		dp = j // 2
		q = j - 2 * dp
		grad[:, :, kk] = packed[dp - dp_min, :, q * L.ci_alloc:q * L.ci_alloc + L.C_in]
	return grad


def forward_training(model, feats_f32, xlen):
	"""normalised-feature input (fp32 [B, C, F]) -> (logits tuple, log_probs list) with autograd
	through the native kernels"""
	layers = getattr(model, '_train_layers', None)
	if layers is None:
		layers = model._train_layers = _layers(model)
	B, C, Fr = feats_f32.shape
	stride = layers[0].stride
	F_pad = Fr + (Fr % 2) if stride == 2 else Fr
	nf = model.normalize_features
	norm_xlen = xlen if (nf is not None and nf.temporal_mask) else None
	hi, _, _ = ops.instnorm_pack(feats_f32, norm_xlen, nf.eps if nf is not None else -1.0, F_pad = F_pad, C_pad = engine._ceil_to(C, 64), normalize = nf is not None)
	params = []
	for L in layers:
		params += [L.conv.weight, L.bn.weight, L.bn.bias]
	dec = model.decoder[0]
	params += [dec.weight] + ([dec.bias] if dec.bias is not None else [])
	seed = getattr(model, '_dropout_seed', None)
	if seed is None or seed.device != hi.device:
		# device-resident dropout counter, initialised from torch's seed so torch.manual_seed controls it
		seed = model._dropout_seed = torch.full((1, ), torch.initial_seed() & 0x7FFFFFFFFFFF, dtype = torch.int64, device = hi.device)
	holder = dict(model = model, layers = layers, params = params, n_frames = Fr, seed = seed)
	logits, log_probs, argmax = NativeStack.apply(holder, hi, xlen, *params)
	log_probs._convasr_argmax = argmax
	return (logits, ), [log_probs]


class GraphedTrainStep:
	"""One whole training step -- zero_grad, forward, CTC loss, backward, optimizer.step -- captured
	once into a CUDA graph and replayed: ~190 kernel launches per step stop paying host launch
	latency and inter-kernel gaps.  Inputs are copied into static buffers; the per-utterance loss of
	the step is returned (a clone).  Shapes are fixed at construction.  Data-parallel replicas
	(parallel.attach_grad_sync) capture their NCCL all-reduces inside the same graph, on the process
	group's communication stream, so they still overlap the backward."""

	def __init__(self, model, optimizer, x, xlen, y, ylen, warmup = 3, max_grad_norm = None):
		self.model, self.optimizer, self.max_grad_norm = model, optimizer, max_grad_norm
		self.params = [p for g in optimizer.param_groups for p in g['params']]
		self.static = [t.clone() for t in (x, xlen, y, ylen)]
		side = torch.cuda.Stream(device = x.device)
		side.wait_stream(torch.cuda.current_stream())
		with torch.cuda.stream(side):
			for _ in range(warmup):
				self._step()
		torch.cuda.current_stream().wait_stream(side)
		torch.cuda.synchronize()
		self.graph = torch.cuda.CUDAGraph()
		# thread_local: the NCCL watchdog thread may query events while this thread captures
		with torch.cuda.graph(self.graph, capture_error_mode = 'thread_local'):
			self.loss = self._step()

	def _step(self):
		sx, sxlen, sy, sylen = self.static
		self.optimizer.zero_grad(set_to_none = True)
		out = self.model(sx, sxlen, y = sy, ylen = sylen)
		loss = (out['loss'] * sylen[:, 0]).mean()  # train.py:754-755
		loss.backward()
		from . import optimizers
		if isinstance(self.optimizer, optimizers._FusedOptimizer):
			self.optimizer.step(max_grad_norm = self.max_grad_norm)  # clipping folded into the native step
		else:
			if self.max_grad_norm is not None:  # train.py:776-779
				torch.nn.utils.clip_grad_norm_(self.params, self.max_grad_norm, error_if_nonfinite = False)
			self.optimizer.step()
		return out['loss'].detach()

	def __call__(self, x, xlen, y, ylen):
		for dst, src in zip(self.static, (x, xlen, y, ylen)):
			dst.copy_(src, non_blocking = True)
		self.graph.replay()
		return self.loss.clone()
