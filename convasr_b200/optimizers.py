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
		self.total_grad_norm = None  # fp32 [1] on the device after a step that computed it

	def _tables(self, gi, params):
		dev = params[0].device
		key = tuple(p.data_ptr() for p in params) + tuple(p.grad.data_ptr() for p in params)
		t = self._groups.get(gi)
		if t is not None and t['key'] == key:
			return t
		if t is None:
			n = len(params)
			resumed = any('momentum_buffer' in self.state[p] for p in params)  # state restored from a checkpoint
			chunk_tensor, chunk_off = [], []
			for i, p in enumerate(params):
				for off in range(0, p.numel(), _CHUNK):
					chunk_tensor.append(i)
					chunk_off.append(off)
			for p in params:
				st = self.state[p]
				if 'momentum_buffer' not in st:
					st['momentum_buffer'] = torch.zeros_like(p, memory_format = torch.preserve_format)
			ema = torch.zeros(n, dtype = torch.float32, device = dev)
			if self._MODE == 1:
				for i, p in enumerate(params):
					if '_grads_ema' in self.state[p]:  # restored from a checkpoint
						ema[i] = self.state[p]['_grads_ema']
					self.state[p]['_grads_ema'] = ema[i]
			t = dict(
				n = n, n_chunks = len(chunk_tensor), ema = ema,
				chunk_tensor = torch.tensor(chunk_tensor, dtype = torch.int32, device = dev), chunk_off = torch.tensor(chunk_off, dtype = torch.int64, device = dev),
				numels = torch.tensor([p.numel() for p in params], dtype = torch.int64, device = dev),
				sumsq = torch.empty(n, dtype = torch.float32, device = dev), scale = torch.empty(n, dtype = torch.float32, device = dev),
				step = torch.zeros(1, dtype = torch.int64, device = dev), first = torch.zeros(1, dtype = torch.int32, device = dev),
				lr = torch.zeros(1, dtype = torch.float32, device = dev), lr_host = None, norm = torch.zeros(1, dtype = torch.float32, device = dev),
				ptrs_pinned = torch.empty(3, n, dtype = torch.int64).pin_memory(), ptrs = torch.empty(3, n, dtype = torch.int64, device = dev)
			)
			if resumed:
				t['step'].fill_(1)
			self._groups[gi] = t
		# (re)publish the pointer tables: gradients are re-allocated by zero_grad(set_to_none=True)
		host = t['ptrs_pinned']
		for i, p in enumerate(params):
			host[0, i] = p.data_ptr()
			host[1, i] = p.grad.data_ptr()
			host[2, i] = self.state[p]['momentum_buffer'].data_ptr()
		t['ptrs'].copy_(host, non_blocking = True)
		t['key'] = key
		return t

	@torch.no_grad()
	def step(self, closure = None, max_grad_norm = None):
		"""max_grad_norm: fold torch.nn.utils.clip_grad_norm_(params, max_grad_norm) into the step
		(per parameter group); the pre-clip norm lands in `self.total_grad_norm`."""
		loss = None
		if closure is not None:
			with torch.enable_grad():
				loss = closure()
		lib = _lib.load()
		capturing = torch.cuda.is_current_stream_capturing()
		for gi, group in enumerate(self.param_groups):
			params = [p for p in group['params'] if p.grad is not None]
			if not params:
				continue
			for p in params:
				if not (p.is_cuda and p.dtype == torch.float32 and p.grad.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()):
					raise RuntimeError('convasr_b200.optimizers: parameters and gradients must be contiguous fp32 CUDA tensors')
			t = self._tables(gi, params)
			if not capturing and t['lr_host'] != group['lr']:
				t['lr'].fill_(group['lr'])
				t['lr_host'] = group['lr']
			momentum, beta2, eps, dampening, nesterov = self._hyper(group)
			rc = lib.cab_optimizer_step(
				self._MODE, t['n'], ops._p(t['ptrs'][0]), ops._p(t['ptrs'][1]), ops._p(t['ptrs'][2]), ops._p(t['numels']), t['n_chunks'],
				ops._p(t['chunk_tensor']), ops._p(t['chunk_off']), _CHUNK, ops._p(t['sumsq']), ops._p(t['ema']), ops._p(t['scale']), ops._p(t['step']),
				ops._p(t['first']), ops._p(t['lr']), momentum, beta2, eps, float(group['weight_decay']), dampening, int(nesterov),
				float(max_grad_norm) if max_grad_norm else 0.0, ops._p(t['norm']), ops._stream()
			)
			_lib.check(rc, 'cab_optimizer_step')
			self.total_grad_norm = t['norm']
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
