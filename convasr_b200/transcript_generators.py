"""Drop-in replacement for the reference's `transcript_generators` module
(transcript_generators.py:8-93).

The reference does `log_probs.argmax(dim=1).cpu().tolist()` and then walks every frame in
Python (B*t iterations).  Here the argmax comes fused from the decoder epilogue (or one native
kernel), the blank/repeat collapse state machine with the 10-blanks->space rule runs on the GPU
(a segmented warp scan, one utterance per warp, csrc/ctc.cu: greedy_collapse_kernel), and the host only touches the
EMITTED tokens to build segments and text.
"""
import torch

from . import ops

try:  # the reference's own segment containers, when its modules are importable next to this one
	import transcripts as _transcripts
except Exception:
	_transcripts = None


class Segment(dict):
	pass


class Transcript(list):
	pass


def _containers():
	if _transcripts is not None and hasattr(_transcripts, 'Segment'):
		return _transcripts.Segment, _transcripts.Transcript
	return Segment, Transcript


class GreedyCTCGenerator:
	def __init__(self, blank_amount_to_space: int = 10):
		self.blank_amount_to_space = blank_amount_to_space
		self._tables = {}

	def _token_tables(self, tokenizer, C, device):
		key = (id(tokenizer), C, str(device))
		if key not in self._tables:
			sil = torch.zeros(C, dtype = torch.uint8)
			for i in tokenizer.silence_tokens_ids:
				if 0 <= i < C:
					sil[i] = 1
			ws = torch.tensor([1 if tokenizer.is_start_word_token(i) else 0 for i in range(C)], dtype = torch.uint8)
			self._tables[key] = (sil.to(device), ws.to(device), ws.tolist())
		return self._tables[key]

	def generate(self, tokenizer, log_probs, begin, end, output_lengths = None, time_stamps = None, segment_text_key: str = 'hyp', segment_extra_info = None):
		SegmentT, TranscriptT = _containers()
		B, C, T = log_probs.shape
		ids = getattr(log_probs, '_convasr_argmax', None)
		if ids is None or ids.shape != (B, T):
			_, ids = ops.log_softmax_argmax(log_probs, want_log_probs = False)
		sil, ws_dev, ws_host = self._token_tables(tokenizer, C, log_probs.device)
		lens = None
		if output_lengths is not None:
			lens = torch.as_tensor(output_lengths, device = log_probs.device).to(torch.int32)
		tok, frm, cnt = ops.greedy_collapse(ids, lens, C, tokenizer.eps_id, tokenizer.space_id, sil, ws_dev, self.blank_amount_to_space)
		# one D2H copy of the emitted tokens (<= T per utterance, typically ~L)
		cnt_h = cnt.cpu()
		n_max = int(cnt_h.clamp(min = 0).max()) if B > 0 else 0
		tok_h = tok[:, :n_max].cpu().tolist()
		frm_h = frm[:, :n_max].cpu().tolist()
		cnt_h = cnt_h.tolist()
		ts = time_stamps.cpu().tolist() if time_stamps is not None else None
		begin_h = torch.clamp(begin, min = 0.0).cpu().tolist() if time_stamps is not None else begin.cpu().tolist()
		end_h = end.cpu().tolist()
		first_frames = None
		if ts is not None:
			# time_begin uses the first non-silence frame of the FULL row (transcript_generators.py:38-46)
			nonsil = (sil[ids.long().clamp(0, C - 1)] == 0)
			first_frames = torch.where(nonsil.any(dim = 1), nonsil.float().argmax(dim = 1), torch.full((B, ), -1, device = ids.device)).cpu().tolist()

		out = []
		for i in range(B):
			transcript = TranscriptT()
			n = cnt_h[i]
			if n < 0:  # nothing but silence in the row
				out.append([transcript])
				continue
			sample_ts = ts[i] if ts is not None else None
			tokens = []
			time_begin = begin_h[i] + sample_ts[first_frames[i]] if sample_ts is not None else begin_h[i]
			time_end = end_h[i]

			def emit():
				seg = SegmentT(begin = time_begin, end = time_end, **{segment_text_key: tokenizer.decode([tokens])[0]})
				if segment_extra_info is not None:
					seg.update(segment_extra_info[i])
				transcript.append(seg)

			for x, f in zip(tok_h[i][:n], frm_h[i][:n]):
				if f < 0:  # space synthesised from a run of blanks: appended without touching the timers
					tokens.append(x)
					continue
				if sample_ts is not None and ws_host[x]:
					emit()  # a word-start token closes the running segment (transcript_generators.py:68-76)
					tokens = [x]
					time_begin = begin_h[i] + sample_ts[f]
				tokens.append(x)
				time_end = begin_h[i] + sample_ts[f] if sample_ts is not None else end_h[i]
			if len(tokens) > 0:
				emit()
			out.append([transcript])
		return out
