"""Drop-in replacement for the reference's `optimizers` module (section 8f "next" #3).

NovoGrad (optimizers.py:66-90) and SGD-with-momentum (torch.optim.SGD as configured at
train.py:657-662) as ONE multi-tensor native step for all parameters, with the gradient-norm
clipping of train.py:776-779 folded in: three launches per step instead of ~8 per tensor.
The learning-rate schedulers are host-side bookkeeping.
"""
import ctypes

import torch

from . import _lib, ops

_CHUNK = 65536


class _FusedOptimizer(torch.optim.Optimizer):
	_MODE = 0

	def __init__(self, params, defaults):
		super().__init__(params, defaults)
		self._groups = {}
		self._state_epoch = 0  # bumped by load_state_dict: cached tables are rebuilt from the restored state
		self.total_grad_norm = None  # fp32 [1] on the device after a step that computed it

	def load_state_dict(self, state_dict):
		super().load_state_dict(state_dict)
		self._groups = {}
		self._state_epoch += 1

	def _tables(self, gi, params, with_state = True):
		"""Device tables of one parameter group (gi = 'all': every group, gradient norm only).  Rebuilt completely when
		the set of parameters with a gradient, their momentum buffers or the optimizer state epoch change; only the
		pointer table is re-published when just the gradient tensors were re-allocated (zero_grad(set_to_none))."""
		dev = params[0].device
		if with_state:
			for p in params:
				st = self.state[p]
				if 'momentum_buffer' not in st:
					st['momentum_buffer'] = torch.zeros_like(p, memory_format = torch.preserve_format)
					st['_fresh'] = True
		struct_key = (self._state_epoch, len(params), tuple(p.data_ptr() for p in params), tuple(self.state[p]['momentum_buffer'].data_ptr() for p in params) if with_state else ())
		grad_key = tuple(p.grad.data_ptr() for p in params)
		t = self._groups.get(gi)
		if t is not None and t['struct_key'] == struct_key and t['grad_key'] == grad_key:
			return t
		if t is None or t['struct_key'] != struct_key:
			n = len(params)
			chunk_tensor, chunk_off = [], []
			for i, p in enumerate(params):
				for off in range(0, p.numel(), _CHUNK):
					chunk_tensor.append(i)
					chunk_off.append(off)
			ema, resumed = torch.zeros(n, dtype = torch.float32, device = dev), False
			if with_state:
				resumed = any(not self.state[p].pop('_fresh', False) for p in params)  # momentum restored from a checkpoint / earlier steps
				for p in params:
					self.state[p].pop('_fresh', None)
				if self._MODE == 1:
					for i, p in enumerate(params):
						if '_grads_ema' in self.state[p]:  # restored from a checkpoint or carried over from the previous tables
							ema[i] = self.state[p]['_grads_ema']
						self.state[p]['_grads_ema'] = ema[i]
			old_step = t['step'] if t is not None else None
			t = dict(
				n = n, n_chunks = len(chunk_tensor), ema = ema,
				chunk_tensor = torch.tensor(chunk_tensor, dtype = torch.int32, device = dev), chunk_off = torch.tensor(chunk_off, dtype = torch.int64, device = dev),
				numels = torch.tensor([p.numel() for p in params], dtype = torch.int64, device = dev),
				sumsq = torch.empty(n + len(chunk_tensor), dtype = torch.float32, device = dev), scale = torch.empty(n, dtype = torch.float32, device = dev),
				step = torch.zeros(1, dtype = torch.int64, device = dev), first = torch.zeros(1, dtype = torch.int32, device = dev),
				lr = torch.zeros(1, dtype = torch.float32, device = dev), lr_host = None, norm = torch.zeros(1, dtype = torch.float32, device = dev),
				# two pinned staging tables used alternately, each guarded by an event: the H2D copy of step k may still be
				# queued when step k + 1 rewrites the host side
				ptrs_pinned = [torch.empty(3, n, dtype = torch.int64).pin_memory() for _ in range(2)], ptrs_event = [None, None], ptrs_turn = 0,
				ptrs_pinned_graph = torch.empty(3, n, dtype = torch.int64).pin_memory(),
				ptrs = torch.empty(3, n, dtype = torch.int64, device = dev)
			)
			if old_step is not None:
				t['step'].copy_(old_step)  # same optimizer, a different set of tensors with gradients: the step count carries over
			elif resumed:
				t['step'].fill_(1)  # momentum / ema hold history: this is not a first step
			self._groups[gi] = t
		capturing = torch.cuda.is_current_stream_capturing()
		# (a captured copy node re-reads its pinned source on every replay: the graph gets a staging table of its own,
		# allocated with the tables because host allocations are not allowed while capturing)
		turn = t['ptrs_turn']
		if not capturing and t['ptrs_event'][turn] is not None:
			t['ptrs_event'][turn].synchronize()
		host = t['ptrs_pinned_graph'] if capturing else t['ptrs_pinned'][turn]
		for i, p in enumerate(params):
			host[0, i] = p.data_ptr()
			host[1, i] = p.grad.data_ptr()
			host[2, i] = self.state[p]['momentum_buffer'].data_ptr() if with_state else 0
		t['ptrs'].copy_(host, non_blocking = True)
		if not capturing:
			ev = torch.cuda.Event()
			ev.record()
			t['ptrs_event'][turn] = ev
			t['ptrs_turn'] = turn ^ 1
		t['struct_key'], t['grad_key'] = struct_key, grad_key
		return t

	@torch.no_grad()
	def step(self, closure = None, max_grad_norm = None):
		"""max_grad_norm: fold torch.nn.utils.clip_grad_norm_(all parameters, max_grad_norm) (train.py:776-779) into the
		step -- the norm is global over every parameter group; the pre-clip norm lands in `self.total_grad_norm`."""
		loss = None
		if closure is not None:
			with torch.enable_grad():
				loss = closure()
		lib = _lib.load()
		capturing = torch.cuda.is_current_stream_capturing()
		groups = [(gi, group, [p for p in group['params'] if p.grad is not None]) for gi, group in enumerate(self.param_groups)]
		groups = [g for g in groups if g[2]]
		for _, _, params in groups:
			for p in params:
				if not (p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
					raise RuntimeError('convasr_b200.optimizers: parameters and gradients must be contiguous fp32 CUDA tensors')
		ext_norm = None
		if max_grad_norm and len(groups) > 1:
			every = [p for _, _, params in groups for p in params]
			ta = self._tables('all', every, with_state = False)
			rc = lib.cab_optimizer_step(2, ta['n'], None, ops._p(ta['ptrs'][1]), None, ops._p(ta['numels']), ta['n_chunks'], ops._p(ta['chunk_tensor']), ops._p(ta['chunk_off']), _CHUNK,
										ops._p(ta['sumsq']), None, None, None, None, None, 0.0, 0.0, 0.0, 0.0, 0.0, 0, 0.0, ops._p(ta['norm']), None, ops._stream())
			_lib.check(rc, 'cab_optimizer_step (gradient norm)')
			ext_norm = ta['norm']
		for gi, group, params in groups:
			t = self._tables(gi, params)
			if not capturing and t['lr_host'] != group['lr']:
				t['lr'].fill_(group['lr'])
				t['lr_host'] = group['lr']
			momentum, beta2, eps, dampening, nesterov = self._hyper(group)
			rc = lib.cab_optimizer_step(
				self._MODE, t['n'], ops._p(t['ptrs'][0]), ops._p(t['ptrs'][1]), ops._p(t['ptrs'][2]), ops._p(t['numels']), t['n_chunks'],
				ops._p(t['chunk_tensor']), ops._p(t['chunk_off']), _CHUNK, ops._p(t['sumsq']), ops._p(t['ema']), ops._p(t['scale']), ops._p(t['step']),
				ops._p(t['first']), ops._p(t['lr']), momentum, beta2, eps, float(group['weight_decay']), dampening, int(nesterov),
				float(max_grad_norm) if max_grad_norm else 0.0, ops._p(t['norm']), ops._p(ext_norm), ops._stream()
			)
			_lib.check(rc, 'cab_optimizer_step')
			self.total_grad_norm = ext_norm if ext_norm is not None else t['norm']
			# the kernels wrote the parameters through raw pointers: tell autograd / the cached eval plans
			for p in params:
				torch.autograd.graph.increment_version(p)
		return loss

	def set_lr(self, lr, group = 0):
		"""update the device-resident learning rate (use between CUDA-graph replays)"""
		self.param_groups[group]['lr'] = lr
		t = self._groups.get(group)
		if t is not None:
			t['lr'].fill_(lr)
			t['lr_host'] = lr


class SGD(_FusedOptimizer):
	"""torch.optim.SGD(lr, momentum, dampening, weight_decay, nesterov) semantics, multi-tensor native step."""
	_MODE = 0

	def __init__(self, params, lr = 1e-2, momentum = 0.9, dampening = 0.0, weight_decay = 0.0, nesterov = False):
		super().__init__(params, dict(lr = lr, momentum = momentum, dampening = dampening, weight_decay = weight_decay, nesterov = nesterov))

	def _hyper(self, group):
		return float(group['momentum']), 0.0, 0.0, float(group['dampening']), bool(group['nesterov'])


class NovoGrad(_FusedOptimizer):
	"""optimizers.py:66-90: per-tensor second moment (EMA of the squared gradient norm), normalised
	gradient + weight decay, momentum; state keys `_grads_ema` and `momentum_buffer` as in the reference."""
	_MODE = 1

	def __init__(self, params, lr = 1.0, betas = (0.95, 0.98), eps = 1e-8, weight_decay = 0.0, dampening = False):
		super().__init__(params, dict(lr = lr, betas = betas, eps = eps, weight_decay = weight_decay, dampening = dampening))

	def _hyper(self, group):
		return float(group['betas'][0]), float(group['betas'][1]), float(group['eps']), 1.0 if group['dampening'] else 0.0, False


# ------------------------------------------------------------------------------------------
# host-side schedules (optimizers.py:4-63)
# ------------------------------------------------------------------------------------------
def reset_options(optimizer):
	for group in optimizer.param_groups:
		group.update(optimizer.defaults)


class LRScheduler:
	def __init__(self, optimizer):
		self.optimizer = optimizer

	def step(self, step):
		for gi, (group, lr) in enumerate(zip(self.optimizer.param_groups, self.get_lr(step))):
			if isinstance(self.optimizer, _FusedOptimizer):
				self.optimizer.set_lr(lr, gi)  # also refreshes the device-resident cell a CUDA-graph replay reads
			else:
				group['lr'] = lr


class NoopLR(LRScheduler):
	def get_lr(self, step):
		return [group['lr'] for group in self.optimizer.param_groups]


class MultiStepLR(LRScheduler):
	def __init__(self, optimizer, gamma, milestones):
		self.init_lr = [group['lr'] for group in optimizer.param_groups]
		self.gamma, self.milestones = gamma, milestones
		super().__init__(optimizer)

	def get_lr(self, step):
		passed = sum(1 for m in self.milestones if step >= m)
		return [lr * self.gamma**passed for lr in self.init_lr]


class PolynomialDecayLR(LRScheduler):
	"""linear warm-up, then polynomial decay to end_lr (the reference's version raises NameError at
	optimizers.py:60; this is the schedule it describes)"""

	def __init__(self, optimizer, decay_steps, power = 1.0, begin_decay_at = 0, end_lr = 0.0, warmup_steps = 0):
		self.decay_steps, self.power, self.begin_decay_at, self.end_lr, self.warmup_steps = decay_steps, power, begin_decay_at, end_lr, warmup_steps
		self.init_lr = [group['lr'] for group in optimizer.param_groups]
		super().__init__(optimizer)

	def get_lr(self, step):
		lrs = [(lr * step / self.warmup_steps) if self.warmup_steps > 0 and step < self.warmup_steps else lr for lr in self.init_lr]
		if step >= self.begin_decay_at:
			s = min(step - self.begin_decay_at, self.decay_steps)
			lrs = [self.end_lr + (lr - self.end_lr) * ((self.decay_steps - s) / self.decay_steps)**self.power if s < self.decay_steps else self.end_lr for lr in lrs]
		return lrs


@torch.no_grad()
def larc_(param_groups, larc_mode = 'clip', eps = 1e-7, min_update = 1e-7, larc_eta = 0.1):
	"""optimizers.py:93-106 (no caller in the reference): layer-wise adaptive rate clipping / scaling of the gradients in place --
	every gradient is multiplied by eta * |p| / (|g| + eps), in 'clip' mode relative to the group's lr and capped at 1"""
	for group in param_groups:
		for p in group['params']:
			if p.grad is None:
				continue
			ratio = larc_eta * p.norm() / (p.grad.norm() + eps)
			if larc_mode == 'clip':
				ratio = torch.clamp(ratio / group['lr'], min = min_update, max = 1)
			else:
				ratio = torch.clamp(ratio, min = min_update)
			p.grad.mul_(ratio)
