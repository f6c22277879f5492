"""Drop-in replacement for the reference's `ctc` module (ctc.py:6-75): forced alignment on the
GPU with one CTA per utterance (convasr_b200/csrc/ctc.cu: ctc_align_kernel).

All of the reference's behaviours are reproduced, including the batch-coupled ones (recursion
over all T frames of the padded batch, terminal state read after the last global frame,
back-trace from input_length-1, finite "zero", stay-preferred argmax ties) -- SURVEY.md A12 -- and
its fp16 mode: on fp16 log_probs every operation of the recursion rounds to fp16 ("zero" = finfo(float16).min).
`pack_backpointers` only changes the reference's memory format, never its result; the native
kernel keeps one byte per back-pointer in a caller-owned workspace and ignores the flag.
"""
import torch

from . import ops


def alignment(
	log_probs, targets, input_lengths, target_lengths, blank: int = 0, pack_backpointers: bool = False,
	finfo_min_fp32: float = torch.finfo(torch.float32).min, finfo_min_fp16: float = torch.finfo(torch.float16).min
):
	"""log_probs [T, B, C] (any strides), targets [B, L] -> int64 [B, L]: for every target label the
	last frame index at which the best path sits in it (zeros past target_length)."""
	if log_probs.dtype not in (torch.float32, torch.float16):
		raise NotImplementedError(f'convasr_b200.ctc.alignment: {log_probs.dtype} log_probs (the reference handles float32 and float16, ctc.py:29)')
	return ops.ctc_alignment(log_probs, targets, input_lengths, target_lengths, blank = blank)


def ctc_loss(log_probs, targets, input_lengths, target_lengths, blank = 0, reduction = 'none'):
	"""torch.nn.functional.ctc_loss(..., zero_infinity=False) as called at models.py:323."""
	nll = ops.ctc_loss(log_probs, targets, input_lengths, target_lengths, blank = blank)
	if reduction == 'none':
		return nll
	if reduction == 'sum':
		return nll.sum()
	if reduction == 'mean':
		tl = torch.as_tensor(target_lengths, device = nll.device).clamp(min = 1).to(nll.dtype)
		return (nll / tl).mean()
	raise ValueError(reduction)
