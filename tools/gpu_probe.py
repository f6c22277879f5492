"""Quick on-GPU sanity probe of each kernel against stock torch CUDA ops (development aid;
the parity tests proper live in tests/ and compare against the oracle).

usage: python tools/gpu_probe.py {conv|frontend|ctc|all}
"""
import os
import sys
import traceback

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from convasr_b200 import _lib, ops

dev = 'cuda'


def pack_weight(w, cin_pad = None):
	# [Cout, Cin, k] fp32 -> bf16 [k, Cout, Cin_pad]
	Cout, Cin, k = w.shape
	cin_pad = cin_pad or (Cin + 63) // 64 * 64
	out = torch.zeros(k, Cout, cin_pad, dtype = torch.bfloat16, device = w.device)
	out[:, :, :Cin] = w.permute(2, 0, 1).to(torch.bfloat16)
	return out.contiguous()


def report(name, got, ref, tol):
	err = (got.float() - ref.float()).abs().max().item()
	scale = ref.float().abs().max().item()
	rel = ((got.float() - ref.float()).norm() / (ref.float().norm() + 1e-30)).item()
	ok = rel <= tol
	print(f'[{"OK" if ok else "FAIL"}] {name}: max_abs_err={err:.4g} ref_max={scale:.4g} rel_fro={rel:.3g} (tol {tol})', flush = True)
	return ok


def conv_case(name, B, T_in, Cin, Cout, k, dil, pad, act, mask, block_n = 0, seed = 0):
	g = torch.Generator(device = 'cpu').manual_seed(seed)
	x = torch.randn(B, Cin, T_in, generator = g).to(dev)
	w = (torch.randn(Cout, Cin, k, generator = g) / (Cin * k)**0.5).to(dev)
	bias = torch.randn(Cout, generator = g).to(dev)
	xlen = torch.rand(B, generator = g).mul(0.5).add(0.5).to(dev) if mask else None
	if xlen is not None:
		xlen[0] = 1.0
	xb = x.to(torch.bfloat16)
	wb = w.to(torch.bfloat16)
	ref = F.conv1d(xb.float(), wb.float(), bias, padding = pad, dilation = dil)
	T_out = ref.shape[-1]
	if act == _lib.ACT_HARDTANH:
		ref = ref.clamp(0, 20)
	elif act == _lib.ACT_RELU:
		ref = ref.relu()
	if xlen is not None:
		ln = (xlen * T_out).ceil().long()
		ref = ref * (torch.arange(T_out, device = dev)[None, None, :] < ln[:, None, None])
	a = xb.permute(0, 2, 1).contiguous()  # [B, T, C]
	out = torch.full((B, T_out, Cout), float('nan'), dtype = torch.bfloat16, device = dev)
	src = ops.Source(a, pack_weight(w), Cin, k, dil, pad)
	ops.conv1d_fused([src], B, T_out, Cout, bias = bias, act = act, act_a = 0.0, act_b = 20.0, xlen = xlen, out_hi = out, block_n = block_n)
	torch.cuda.synchronize()
	return report(name, out.permute(0, 2, 1), ref, 1e-2)


def probe_conv():
	ok = True
	ok &= conv_case('conv 1x1 64->64 T=300', 2, 300, 64, 64, 1, 1, 0, _lib.ACT_NONE, False)
	ok &= conv_case('conv 1x1 128->256 T=128', 1, 128, 128, 256, 1, 1, 0, _lib.ACT_NONE, False)
	ok &= conv_case('conv k11 256->256 T=753 hardtanh+mask', 3, 753, 256, 256, 11, 1, 5, _lib.ACT_HARDTANH, True)
	ok &= conv_case('conv k11 384->512 T=503 relu', 2, 503, 384, 512, 11, 1, 5, _lib.ACT_RELU, True)
	ok &= conv_case('conv k29 d2 768->896 T=751->753', 2, 751, 768, 896, 29, 2, 29, _lib.ACT_HARDTANH, True)
	ok &= conv_case('conv k11 640->640 bn=160', 2, 400, 640, 640, 11, 1, 5, _lib.ACT_HARDTANH, True, block_n = 160)
	ok &= conv_case('conv many tiles 256->256 B=40', 40, 753, 256, 256, 11, 1, 5, _lib.ACT_HARDTANH, True)

	# residual: two sources accumulate into the same tile
	g = torch.Generator().manual_seed(1)
	B, T, C1, C2, Co = 2, 333, 256, 128, 384
	x1 = torch.randn(B, C1, T, generator = g).to(dev).to(torch.bfloat16)
	x2 = torch.randn(B, C2, T, generator = g).to(dev).to(torch.bfloat16)
	w1 = (torch.randn(Co, C1, 13, generator = g) / (C1 * 13)**0.5).to(dev)
	w2 = (torch.randn(Co, C2, 1, generator = g) / C2**0.5).to(dev)
	bias = torch.randn(Co, generator = g).to(dev)
	ref = (F.conv1d(x1.float(), w1.to(torch.bfloat16).float(), None, padding = 6) + F.conv1d(x2.float(), w2.to(torch.bfloat16).float(), None) + bias[None, :, None]).relu()
	out = torch.empty(B, T, Co, dtype = torch.bfloat16, device = dev)
	ops.conv1d_fused([ops.Source(x1.permute(0, 2, 1).contiguous(), pack_weight(w1), C1, 13, 1, 6), ops.Source(x2.permute(0, 2, 1).contiguous(), pack_weight(w2), C2, 1, 1, 0)], B, T, Co, bias = bias, act = _lib.ACT_RELU, out_hi = out)
	torch.cuda.synchronize()
	ok &= report('conv residual 2 sources', out.permute(0, 2, 1), ref, 1e-2)

	# stride 2 through the frame-pair view: k=11, pad=5, 64 -> 256
	F_ = 1501
	x = torch.randn(B, 64, F_, generator = g).to(dev).to(torch.bfloat16)
	w = (torch.randn(256, 64, 11, generator = g) / (64 * 11)**0.5).to(dev)
	ref = F.conv1d(x.float(), w.to(torch.bfloat16).float(), bias[:256], stride = 2, padding = 5).clamp(0, 20)
	T_out = ref.shape[-1]
	F_pad = F_ + (F_ % 2)
	a = torch.zeros(B, F_pad, 64, dtype = torch.bfloat16, device = dev)
	a[:, :F_] = x.permute(0, 2, 1)
	a_pairs = a.view(B, F_pad // 2, 128)
	# tap dp in [-3, 2], parity q: original tap k = 2*dp + q + 5
	wp = torch.zeros(6, 256, 128, dtype = torch.bfloat16, device = dev)
	for dp in range(-3, 3):
		for q in range(2):
			k = 2 * dp + q + 5
			if 0 <= k < 11:
				wp[dp + 3, :, q * 64:(q + 1) * 64] = w[:, :, k].to(torch.bfloat16)
	out = torch.empty(B, T_out, 256, dtype = torch.bfloat16, device = dev)
	ops.conv1d_fused([ops.Source(a_pairs, wp, 128, 6, 1, 3)], B, T_out, 256, bias = bias[:256].contiguous(), act = _lib.ACT_HARDTANH, act_b = 20.0, out_hi = out)
	torch.cuda.synchronize()
	ok &= report('conv stride-2 pair view', out.permute(0, 2, 1), ref, 1e-2)

	# decoder epilogue: 1024 -> 38 + log_softmax + argmax
	B, T, Ci, C = 3, 753, 1024, 38
	x = torch.randn(B, Ci, T, generator = g).to(dev).to(torch.bfloat16)
	w = (torch.randn(C, Ci, 1, generator = g) / Ci**0.5 * 3).to(dev)
	bias = torch.randn(C, generator = g).to(dev)
	ref = F.conv1d(x.float(), w.to(torch.bfloat16).float(), bias)
	logits = torch.empty(B, C, T, device = dev)
	lp = torch.empty(B, C, T, device = dev)
	am = torch.empty(B, T, dtype = torch.int32, device = dev)
	ops.conv1d_fused([ops.Source(x.permute(0, 2, 1).contiguous(), pack_weight(w), Ci, 1, 1, 0)], B, T, C, bias = bias, logits = logits, log_probs = lp, argmax = am, epilogue = _lib.EPI_LOGSOFTMAX)
	torch.cuda.synchronize()
	ok &= report('decoder logits', logits, ref, 1e-3)
	ok &= report('decoder log_probs', lp, ref.log_softmax(1), 1e-3)
	agree = (am.long() == logits.argmax(1)).float().mean().item()
	print(f'[{"OK" if agree == 1.0 else "FAIL"}] decoder argmax agreement with own logits: {agree}')
	ok &= agree == 1.0

	# split-bf16 (hi/lo) fp32 tier: 3 sources
	B, T, Ci, Co = 2, 300, 256, 256
	x = torch.randn(B, Ci, T, generator = g).to(dev)
	w = (torch.randn(Co, Ci, 11, generator = g) / (Ci * 11)**0.5).to(dev)
	ref = F.conv1d(x.double(), w.double(), None, padding = 5).float()
	xh = x.to(torch.bfloat16); xl = (x - xh.float()).to(torch.bfloat16)
	wh = w.to(torch.bfloat16); wl = (w - wh.float()).to(torch.bfloat16)
	ah = xh.permute(0, 2, 1).contiguous(); al = xl.permute(0, 2, 1).contiguous()
	pwh = pack_weight(wh.float()); pwl = pack_weight(wl.float())
	oh = torch.empty(B, T, Co, dtype = torch.bfloat16, device = dev); ol = torch.empty_like(oh)
	ops.conv1d_fused([ops.Source(ah, pwh, Ci, 11, 1, 5), ops.Source(ah, pwl, Ci, 11, 1, 5), ops.Source(al, pwh, Ci, 11, 1, 5)], B, T, Co, out_hi = oh, out_lo = ol)
	torch.cuda.synchronize()
	ok &= report('conv split-bf16 fp32 tier', (oh.float() + ol.float()).permute(0, 2, 1), ref, 2e-5)
	return ok


def probe_frontend():
	import math
	import torchaudio
	ok = True
	g = torch.Generator().manual_seed(0)
	B, T = 3, 24000
	x = (torch.randn(B, T, generator = g) * 3000).round().clamp(-32767, 32767)
	xlen = torch.tensor([1.0, 0.63, 0.81])
	win = torch.hann_window(160, periodic = True)
	mel = torchaudio.functional.melscale_fbanks(129, 0., 4000., 64, 8000, norm = 'slaney', mel_scale = 'slaney').T.contiguous()

	def ref_frontend(sig, xlen):
		sig = sig.float()
		s = sig / (sig.abs().max(dim = -1, keepdim = True).values + 1e-5)
		s = torch.cat([s[..., :1], s[..., 1:] - 0.97 * s[..., :-1]], dim = -1)
		ln = (xlen * T).ceil().long()
		s = s * (torch.arange(T)[None] < ln[:, None])
		p = F.pad(s.unsqueeze(1), (128, 0), mode = 'reflect').squeeze(1)
		p = F.pad(p, (0, 128))
		st = torch.view_as_real(p.stft(256, hop_length = 80, win_length = 160, window = win, center = False, return_complex = True))
		pw = (st * st).sum(-1)
		return (torch.einsum('mk,bkf->bmf', mel, pw) + 2.0**-14).log()

	ref = ref_frontend(x, xlen)
	tw = ops.make_twiddle(256, dev)
	melb = ops.make_mel_band(mel.to(dev))
	for name, sig in [('fp32', x.to(dev)), ('int16', x.to(torch.int16).to(dev))]:
		got = ops.frontend_logmel(sig, xlen.to(dev), win.to(dev), mel.to(dev), melb, tw, 80, 256)
		torch.cuda.synchronize()
		err = (got.cpu() - ref).abs().max().item()
		print(f'[{"OK" if err < 2e-4 else "FAIL"}] frontend logmel {name}: max_abs_err={err:.3g} shape={tuple(got.shape)}', flush = True)
		ok &= err < 2e-4
	# instance norm + pack
	F_ = ref.shape[-1]
	ln = (xlen * F_).ceil().long()
	m = (torch.arange(F_)[None, None] < ln[:, None, None])
	n = m.int().sum(-1, keepdim = True)
	mean = (ref * m).sum(-1, keepdim = True) / n
	z = m * (ref - mean)
	std = ((z * z).sum(-1, keepdim = True) / n).add(2.0**-14).sqrt()
	refn = z / std
	hi, lo, f32 = ops.instnorm_pack(ref.to(dev), xlen.to(dev), 2.0**-14, F_pad = F_ + 1, C_pad = 64, want_lo = True, want_f32 = True)
	torch.cuda.synchronize()
	e1 = (f32.cpu() - refn).abs().max().item()
	e2 = ((hi.float() + lo.float())[:, :F_].permute(0, 2, 1).cpu() - refn).abs().max().item()
	padz = hi[:, F_:].float().abs().max().item()
	print(f'[{"OK" if e1 < 1e-5 and e2 < 1e-4 and padz == 0 else "FAIL"}] instnorm: f32 err={e1:.3g} hi+lo err={e2:.3g} pad={padz}', flush = True)
	ok &= e1 < 1e-5 and e2 < 1e-4 and padz == 0
	return ok


def probe_ctc():
	ok = True
	g = torch.Generator().manual_seed(0)
	B, C, T, L = 5, 38, 203, 40
	logits = torch.randn(B, C, T, generator = g).to(dev)
	lp, am = ops.log_softmax_argmax(logits)
	ref_lp = logits.log_softmax(1)
	ok &= report('log_softmax', lp, ref_lp, 1e-6)
	print('[%s] argmax' % ('OK' if bool((am.long() == logits.argmax(1)).all()) else 'FAIL'))
	y = torch.randint(0, C - 1, (B, L), generator = g).to(dev)
	ylen = torch.tensor([40, 33, 1, 0, 20]).to(dev)
	olen = torch.tensor([203, 150, 203, 77, 60]).to(dev)
	lpt = ref_lp.permute(2, 0, 1).detach().requires_grad_(True)
	ref_loss = F.ctc_loss(lpt, y, olen, ylen, blank = C - 1, reduction = 'none')
	w = torch.rand(B, generator = g).to(dev) + 0.5
	(ref_loss * w).sum().backward()
	lpt2 = ref_lp.permute(2, 0, 1).detach().requires_grad_(True)
	loss = ops.ctc_loss(lpt2, y, olen, ylen, blank = C - 1)
	(loss * w).sum().backward()
	torch.cuda.synchronize()
	ok &= report('ctc loss', loss, ref_loss, 1e-5)
	ok &= report('ctc grad', lpt2.grad, lpt.grad, 1e-4)
	print('loss', loss.tolist(), 'ref', ref_loss.tolist())
	# infeasible
	loss_inf = ops.ctc_loss(lpt2.detach()[:5], y, torch.full((B, ), 5, device = dev), ylen, blank = C - 1)
	print('infeasible losses', loss_inf.tolist())
	e, wme = ops.entropy(lp, olen)
	ee = -(lp.exp() * lp).sum(1)
	mask = torch.arange(T, device = dev)[None] < olen[:, None]
	ref_e = (ee * mask).sum(-1) / (1e-9 + olen.float())
	wts = (1 - lp.exp()[:, -1]) * mask
	ref_w = (ee * wts).sum(-1) / (1e-9 + wts.sum(-1))
	ok &= report('entropy', e, ref_e, 1e-5)
	ok &= report('weighted entropy', wme, ref_w, 1e-5)
	tk = ops.topk_ids(lp, 3)
	ref_tk = lp.topk(3, dim = 1).indices
	print('[%s] topk' % ('OK' if bool((tk.long() == ref_tk).all()) else 'FAIL(ties?)'))
	al = ops.ctc_alignment(lp.permute(2, 0, 1), y, olen, ylen, blank = C - 1)
	print('alignment sample', al[0, :10].tolist(), al[2, :3].tolist())
	return ok


def probe_wgrad():
	ok = True
	g = torch.Generator().manual_seed(0)
	for (B, T_in, Ci, Co, k, dil, pad, splits) in [(2, 200, 64, 128, 1, 1, 0, 1), (3, 333, 256, 256, 11, 1, 5, 0), (2, 751, 768, 896, 29, 2, 29, 0), (4, 300, 128, 384, 3, 1, 1, 3), (2, 150, 1024, 38, 1, 1, 0, 0)]:
		x = torch.randn(B, Ci, T_in, generator = g).to(dev).to(torch.bfloat16)
		w = torch.zeros(Co, Ci, k, device = dev, requires_grad = True)
		y = F.conv1d(x.float(), w, None, padding = pad, dilation = dil)
		dy = torch.randn(y.shape, generator = g).to(dev).to(torch.bfloat16)
		y.backward(dy.float())
		ref = w.grad  # [Co, Ci, k]
		T_out = y.shape[-1]
		co_ld, ci_ld = (Co + 63) // 64 * 64, (Ci + 63) // 64 * 64
		a = torch.zeros(B, T_out, co_ld, dtype = torch.bfloat16, device = dev); a[:, :, :Co] = dy.permute(0, 2, 1)
		bx = torch.zeros(B, T_in, ci_ld, dtype = torch.bfloat16, device = dev); bx[:, :, :Ci] = x.permute(0, 2, 1)
		if Co >= Ci or Co >= 128:
			out = ops.conv1d_wgrad(a, T_out, Co, bx, T_in, Ci, k, dil, pad, n_splits = splits)
			got = out[:, :, :Ci].permute(1, 2, 0)
		else:  # swapped: wide side on M; shift applies to the B operand (here dy) with the opposite sign
			out = ops.conv1d_wgrad(bx, T_in, Ci, a, T_out, Co, k, dil, -pad + 0, n_splits = splits) if k == 1 else None
			got = out[:, :, :Co].permute(2, 1, 0)
		torch.cuda.synchronize()
		ok &= report(f'wgrad B{B} T{T_in} {Ci}->{Co} k{k} d{dil} splits={splits}', got, ref, 2e-3)
	return ok


if __name__ == '__main__':
	what = sys.argv[1] if len(sys.argv) > 1 else 'all'
	print('device', torch.cuda.get_device_name(0), 'lib', _lib.lib_path(), flush = True)
	res = {}
	for name, fn in [('conv', probe_conv), ('frontend', probe_frontend), ('ctc', probe_ctc), ('wgrad', probe_wgrad)]:
		if what in (name, 'all'):
			try:
				res[name] = fn()
			except Exception:
				traceback.print_exc()
				res[name] = False
	print('RESULT', res, 'launches', _lib.launch_count())
	sys.exit(0 if all(res.values()) else 1)
