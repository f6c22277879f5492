"""Drop-in replacement for the reference's `decoders` module (decoders.py:5-16).

GreedyDecoder: per-frame top-K class ids truncated to the output length -- NO blank/repeat
collapse (that lives in transcript_generators.GreedyCTCGenerator).  K=1 is the fused argmax
(ties -> lowest id, as torch.argmax); K>1 is the native top-K kernel (K <= 8).
BeamSearchDecoder wraps the external ctcdecode C++ library and is out of scope.
"""
import torch

from . import ops


class GreedyDecoder:
	def decode(self, log_probs, output_lengths = None, K = 1):
		B, C, T = log_probs.shape
		lens = [T] * B if output_lengths is None else torch.as_tensor(output_lengths).tolist()
		if K == 1:
			ids = getattr(log_probs, '_convasr_argmax', None)
			if ids is None or ids.shape != (B, T):
				_, ids = ops.log_softmax_argmax(log_probs, want_log_probs = False)
			rows = ids.cpu().tolist()
			return [row[:o] for row, o in zip(rows, lens)]
		ids = ops.topk_ids(log_probs, K).cpu()
		return [ids[b, :, :o].tolist() for b, o in enumerate(lens)]


class BeamSearchDecoder:
	def __init__(self, *args, **kwargs):
		raise NotImplementedError('convasr_b200: BeamSearchDecoder (external ctcdecode + KenLM) is outside the hot path; use GreedyDecoder')
