"""Batch feed: collate variable-length utterances into one padded batch in PINNED host memory and
upload it on a copy stream while the previous step computes -- SURVEY.md section 8(f) "next" #2.

Reference sites: `AudioTextDataset.collate_fn` datasets.py:305-332 (pageable zero-filled tensors, one
slice assignment per utterance) and the upload in the train loop, train.py:745 (`x.to(device,
non_blocking = True)` from pageable memory, i.e. a synchronous staged copy).  At the native step's rate
(~5e4 audio-s/s per GPU for training, 2e5 for inference) the int16 PCM stream is 0.8-3.2 GB/s per GPU:
`collate` keeps PCM as int16 (the frontend runs on the GPU and never rescales it, models.py:568) and
writes straight into pinned buffers that are recycled, `DeviceFeeder` overlaps the H2D copy of batch
i+1 with the compute of batch i.  Same output contract as the reference's collate_fn:
(meta, s, x, xlen, y, ylen).
"""
import math

import torch

SPEAKER_MISSING = 0  # transcripts.py:13


class PinnedPool:
	"""Recycles pinned host buffers by (dtype, numel bucket): cudaHostAlloc costs milliseconds, a batch
	of 80 x 15 s int16 PCM is 19 MB every ~25 ms."""

	def __init__(self, pin = None):
		self.pin = torch.cuda.is_available() if pin is None else pin
		self.free = {}

	def take(self, shape, dtype, fill = 0):
		numel = 1
		for d in shape:
			numel *= int(d)
		bucket = max(1, 1 << max(0, (numel - 1).bit_length()))
		stack = self.free.setdefault((dtype, bucket), [])
		base = stack.pop() if stack else torch.empty(bucket, dtype = dtype, pin_memory = self.pin)
		out = base[:numel].view(*shape) if numel > 0 else base[:0].view(*shape)
		out._convasr_pool_base = base
		if fill is not None:
			out.fill_(fill)
		return out

	def give(self, *tensors):
		for t in tensors:
			base = getattr(t, '_convasr_pool_base', None)
			if base is not None:
				self.free.setdefault((base.dtype, base.numel()), []).append(base)


def collate(batch, time_padding_multiple = 1, pool = None, batch_mode = False):
	"""datasets.py:305-332.  batch: list of items `[meta, speakers (int64 [S]), x ([C, T] features or [1, T] PCM),
	*targets (int64 [L] per text pipeline)]` as `AudioTextDataset.__getitem__` returns them (`batch_mode`: items are
	already lists per field and get zipped first, datasets.py:307-308).  Returns `(meta, s, x, xlen, y, ylen)`:
	x `[B, C, T_max]` in the items' dtype, zero padded to a multiple of `time_padding_multiple`; xlen fp32 fractions
	`T_k / T_max` (1.0 for an empty batch row of width 0); y int64 `[B, pipelines, L_max]`; ylen int64
	`[B, pipelines]`; s int64 `[B, S_max]` padded with transcripts.speaker_missing."""
	if batch_mode:
		batch = list(zip(*batch))
	pool = pool or PinnedPool(pin = False)
	_, sample_s, sample_x, *sample_y = batch[0]
	multiples = [1, 1, time_padding_multiple] + [time_padding_multiple] * len(sample_y)
	smax, xmax, *ymax = [
		int(math.ceil(max(b[k].shape[-1] for b in batch) / multiples[k])) * multiples[k] for k in range(1, len(batch[0]))
	]
	B = len(batch)
	meta = [b[0] for b in batch]
	x = pool.take((B, len(sample_x), xmax), sample_x.dtype)
	y = pool.take((B, len(sample_y), max(ymax) if ymax else 0), torch.int64)
	s = pool.take((B, smax), torch.int64, fill = SPEAKER_MISSING)
	xlen = pool.take((B, ), torch.float32)
	ylen = pool.take((B, len(sample_y)), torch.int64)
	for k, (_, item_s, item_x, *item_y) in enumerate(batch):
		xlen[k] = item_x.shape[-1] / x.shape[-1] if x.shape[-1] > 0 else 1.0
		x[k, ..., :item_x.shape[-1]] = item_x
		s[k, :item_s.shape[-1]] = item_s
		for j, t in enumerate(item_y):
			y[k, j, :t.shape[-1]] = t
			ylen[k, j] = len(t)
	return meta, s, x, xlen, y, ylen


class DeviceFeeder:
	"""Iterates host batches `(meta, s, x, xlen, y, ylen)` and yields them with x, xlen, y, ylen on the GPU
	(train.py:745), uploading batch i+1 on a dedicated copy stream while the caller computes on batch i.

	Device memory is a ring of `depth + 1` slots that are reused (no allocator traffic in steady state): the upload
	into a slot waits, on the copy stream, for the event the consumer's stream recorded when it asked for the batch
	AFTER the one that last lived in that slot; the consumer's stream waits for the copy's event.  No host
	synchronisation on the data path; pinned host buffers go back to `pool` once their copy has completed."""

	def __init__(self, batches, device, pool = None, depth = 2):
		self.it = iter(batches)
		self.device = torch.device(device)
		self.pool = pool
		self.depth = max(1, depth)
		self.copy_stream = torch.cuda.Stream(device = self.device)
		self.slots = [dict(buffers = [None] * 4, free = None) for _ in range(self.depth + 1)]
		self.n_uploaded = 0
		self.queue = []
		self.bytes_uploaded = 0

	def _slot_view(self, slot, k, host):
		nbytes = host.numel() * host.element_size()
		buf = slot['buffers'][k]
		if buf is None or buf.numel() < nbytes:
			buf = slot['buffers'][k] = torch.empty(max(256, 1 << (nbytes - 1).bit_length()) if nbytes > 0 else 256, dtype = torch.uint8, device = self.device)
		return buf[:nbytes].view(host.dtype).view(host.shape)

	def _upload_next(self):
		try:
			meta, s, x, xlen, y, ylen = next(self.it)
		except StopIteration:
			return False
		host = (x, xlen, y, ylen)
		slot = self.slots[self.n_uploaded % len(self.slots)]
		self.n_uploaded += 1
		with torch.cuda.stream(self.copy_stream):
			if slot['free'] is not None:
				self.copy_stream.wait_event(slot['free'])  # the consumer is done with the batch that lived here
			dev = tuple(self._slot_view(slot, k, t) for k, t in enumerate(host))
			for d, h in zip(dev, host):
				d.copy_(h, non_blocking = True)
			done = torch.cuda.Event()
			done.record(self.copy_stream)
		self.bytes_uploaded += sum(t.numel() * t.element_size() for t in host)
		self.queue.append((meta, s, dev, host, done, slot))
		return True

	def __iter__(self):
		while len(self.queue) < self.depth and self._upload_next():
			pass
		previous = None
		while self.queue:
			meta, s, dev, host, done, slot = self.queue.pop(0)
			consumer = torch.cuda.current_stream(self.device)
			if previous is not None:
				# everything the caller enqueued for the previous batch precedes this point of its stream
				previous['free'] = torch.cuda.Event()
				previous['free'].record(consumer)
			consumer.wait_event(done)
			for t in dev:
				t.record_stream(consumer)  # the slot's memory was allocated on the copy stream; matters only if a slot is ever re-grown
			self._upload_next()  # goes out while the caller works on `dev`
			if self.pool is not None:
				done.synchronize()  # complete in steady state: this batch was uploaded `depth` steps ago
				self.pool.give(*host)
			previous = slot
			yield (meta, s) + dev
