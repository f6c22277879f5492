"""Multi-GPU plumbing: one process per GPU, utterance-sharded replicas for inference (no
data-path collective) and, for training, the gradient all-reduce (NCCL over NVLink) launched layer
by layer from inside the native backward so it overlaps the remaining dgrad / wgrad kernels --
SURVEY.md section 8(e), reference sites train.py:852-874 (init), models.py:755-765 (DDP wrap),
utils.py:193-211 (gather of variable-length results).
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend = None, device = None):
	"""torchrun-style bring-up: RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT.
	Returns (rank, world, local_rank).  NCCL when CUDA is present, gloo otherwise (train.py:855-866)."""
	rank = int(os.environ.get('RANK', 0))
	world = int(os.environ.get('WORLD_SIZE', 1))
	local_rank = int(os.environ.get('LOCAL_RANK', 0))
	if world > 1 and not dist.is_initialized():
		use_cuda = torch.cuda.is_available() if device is None else str(device).startswith('cuda')
		backend = backend or ('nccl' if use_cuda else 'gloo')
		if use_cuda:
			torch.cuda.set_device(local_rank)
			dist.init_process_group(backend, device_id = torch.device('cuda', local_rank))
		else:
			dist.init_process_group(backend)
	return rank, world, local_rank


class GradSync:
	"""Gradient averaging across ranks for the native training step (what DistributedDataParallel's
	reducer does for the reference, models.py:755-765).  The native backward is ONE autograd node that
	walks the layers itself, so it hands every finished weight gradient to `reduce()` right away: the
	all-reduce runs on the process group's communication stream while the compute stream continues with
	the next dgrad / wgrad; `finish()` makes the compute stream wait for all of them.  No bucket copies
	(gradients are reduced in place), one collective per conv layer plus one for all the small tensors.
	Everything is stream-ordered (no host synchronisation), so the step can be captured in a CUDA graph."""

	def __init__(self, group = None):
		assert dist.is_initialized(), 'torch.distributed is not initialised'
		self.group = group
		self.world = dist.get_world_size(group)
		self.native_avg = dist.get_backend(group) == 'nccl'  # gloo has no AVG: sum, then scale
		self.pending = []
		self.n_collectives = 0

	def reduce(self, tensor):
		if self.world == 1:
			return
		assert tensor.is_contiguous()
		op = dist.ReduceOp.AVG if self.native_avg else dist.ReduceOp.SUM
		self.pending.append((dist.all_reduce(tensor, op = op, group = self.group, async_op = True), tensor))
		self.n_collectives += 1

	def finish(self):
		for work, tensor in self.pending:
			work.wait()  # stream-level wait on CUDA, host wait on CPU (gloo)
			if not self.native_avg:
				tensor.mul_(1.0 / self.world)
		self.pending.clear()


def attach_grad_sync(model, group = None):
	"""Make `model` a data-parallel replica: parameters and buffers are broadcast from rank 0 (as the
	DDP constructor does) and the native backward averages gradients through a GradSync."""
	with torch.no_grad():
		for t in list(model.parameters()) + list(model.buffers()):
			dist.broadcast(t, src = dist.get_global_rank(group, 0) if group is not None else 0, group = group)
	model._grad_sync = GradSync(group)
	return model


def shard_bounds(n_items, rank, world):
	"""contiguous, balanced [lo, hi) range of rank: sizes differ by at most one"""
	base, rem = divmod(n_items, world)
	lo = rank * base + min(rank, rem)
	return lo, lo + base + (1 if rank < rem else 0)


def shard_by_length(lengths, world):
	"""Length-bucketed sharding: sort utterances by length (longest first) and deal them out
	round-robin so every rank gets the same number of utterances (+-1) and a similar amount of
	audio; within a rank, indices stay sorted by length so micro-batches pad little
	(the job BucketingBatchSampler does in the reference, datasets.py:357-395)."""
	order = sorted(range(len(lengths)), key = lambda i: (-float(lengths[i]), i))
	return [order[r::world] for r in range(world)]


def gather_in_order(local_items, local_indices, world = None):
	"""Every rank contributes (items, their global indices); every rank gets the full list in global
	order.  Host objects only (token lists, strings): nothing on the data path."""
	if not dist.is_initialized() or dist.get_world_size() == 1:
		pairs = list(zip(local_indices, local_items))
	else:
		bucket = [None] * dist.get_world_size()
		dist.all_gather_object(bucket, list(zip(local_indices, local_items)))
		pairs = [p for part in bucket for p in part]
	pairs.sort(key = lambda p: p[0])
	return [item for _, item in pairs]


def max_over_ranks(value, device = None):
	"""device-timed durations are reported as the max over ranks (never wall clock)"""
	if not dist.is_initialized() or dist.get_world_size() == 1:
		return float(value)
	t = torch.tensor([float(value)], dtype = torch.float64, device = device or ('cuda' if dist.get_backend() == 'nccl' else 'cpu'))
	dist.all_reduce(t, op = dist.ReduceOp.MAX)
	return float(t.item())


def sum_over_ranks(value, device = None):
	if not dist.is_initialized() or dist.get_world_size() == 1:
		return float(value)
	t = torch.tensor([float(value)], dtype = torch.float64, device = device or ('cuda' if dist.get_backend() == 'nccl' else 'cpu'))
	dist.all_reduce(t, op = dist.ReduceOp.SUM)
	return float(t.item())


@torch.no_grad()
def transcribe_sharded(model, generator, tokenizer, signals, xlen, micro_batch = 256):
	"""Inference over this rank's shard in micro-batches; returns per-utterance hypothesis strings
	for the shard (the caller gathers them with gather_in_order).  signals: [N_local, T] on the GPU."""
	out = []
	for lo in range(0, signals.shape[0], micro_batch):
		x, xl = signals[lo:lo + micro_batch], xlen[lo:lo + micro_batch]
		res = model(x, xl)
		lp, olen = res['log_probs'][0], res['olen'][0]
		n = x.shape[0]
		tr = generator.generate(tokenizer, lp, begin = torch.zeros(n, device = x.device), end = torch.ones(n, device = x.device), output_lengths = olen)
		out += [' '.join(seg['hyp'] for seg in t[0]) for t in tr]
	return out
