"""Native multi-tensor optimizers (section 8f "next" #3) against torch.optim.SGD on CPU and the
oracle's restatement of the reference's NovoGrad (optimizers.py:72-90), with and without the folded
gradient-norm clipping of train.py:776-779."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle as O


def _tensors(seed):
	g = torch.Generator().manual_seed(seed)
	shapes = [(256, 64, 11), (256, ), (300000, ), (38, 1024, 1), (7, ), (1, )]
	return [torch.randn(s, generator = g) for s in shapes], g


def test_sgd_matches_torch_sgd_with_clipping():
	from convasr_b200 import optimizers
	dev = torch.device('cuda:0')
	for nesterov, wd, max_norm in ((False, 1e-3, 100.0), (True, 0.0, None), (False, 0.0, 5.0)):
		ps, g = _tensors(1)
		ref = [p.clone().requires_grad_(True) for p in ps]
		mine = [p.clone().to(dev).requires_grad_(True) for p in ps]
		opt_ref = torch.optim.SGD(ref, lr = 0.05, momentum = 0.9, weight_decay = wd, nesterov = nesterov)
		opt = optimizers.SGD(mine, lr = 0.05, momentum = 0.9, weight_decay = wd, nesterov = nesterov)
		for step in range(4):
			grads = [torch.randn(p.shape, generator = g) * (3.0 if step == 1 else 0.3) for p in ps]
			for p, q, gr in zip(ref, mine, grads):
				p.grad = gr.clone()
				q.grad = gr.clone().to(dev)
			if max_norm is not None:
				total = torch.nn.utils.clip_grad_norm_(ref, max_norm)
			opt_ref.step()
			opt.step(max_grad_norm = max_norm)
			if max_norm is not None:
				assert torch.allclose(opt.total_grad_norm.cpu(), total.reshape(1), rtol = 1e-5)
			for p, q in zip(ref, mine):
				assert torch.allclose(q.detach().cpu(), p.detach(), rtol = 1e-5, atol = 1e-6), (step, p.shape)


def test_novograd_matches_reference_restatement():
	from convasr_b200 import optimizers
	dev = torch.device('cuda:0')
	for wd, dampening, max_norm in ((1e-3, False, None), (0.0, True, 2.0)):
		ps, g = _tensors(2)
		ref = [p.clone() for p in ps]
		state = [dict() for _ in ps]
		mine = [p.clone().to(dev).requires_grad_(True) for p in ps]
		opt = optimizers.NovoGrad(mine, lr = 0.01, betas = (0.95, 0.98), weight_decay = wd, dampening = dampening)
		for step in range(4):
			grads = [torch.randn(p.shape, generator = g) for p in ps]
			for q, gr in zip(mine, grads):
				q.grad = gr.clone().to(dev)
			use = grads
			if max_norm is not None:
				use, total = O.clip_grad_norm(grads, max_norm)
			O.novograd_step(ref, use, state, lr = 0.01, betas = (0.95, 0.98), weight_decay = wd, dampening = dampening)
			opt.step(max_grad_norm = max_norm)
			for p, q in zip(ref, mine):
				assert torch.allclose(q.detach().cpu(), p, rtol = 2e-5, atol = 1e-6), (step, p.shape)
		# state keys as in the reference (checkpoint compatibility, train.py:324-341)
		st = opt.state[mine[0]]
		assert set(st) == {'momentum_buffer', '_grads_ema'}
		assert torch.allclose(st['_grads_ema'].cpu(), state[0]['_grads_ema'], rtol = 1e-4)


def test_schedulers_host_logic():
	from convasr_b200 import optimizers
	p = [torch.zeros(1, device = 'cuda', requires_grad = True)]
	opt = optimizers.SGD(p, lr = 1.0)
	sch = optimizers.MultiStepLR(opt, gamma = 0.5, milestones = [10, 20])
	got = []
	for s in (0, 9, 10, 19, 20, 100):
		sch.step(s)
		got.append(opt.param_groups[0]['lr'])
	assert got == [1.0, 1.0, 0.5, 0.5, 0.25, 0.25]
	opt2 = optimizers.SGD(p, lr = 1.0)
	poly = optimizers.PolynomialDecayLR(opt2, decay_steps = 10, power = 1.0, end_lr = 0.0, warmup_steps = 2)
	vals = []
	for s in (0, 1, 2, 5, 10, 11):
		poly.step(s)
		vals.append(round(opt2.param_groups[0]['lr'], 6))
	assert vals == [0.0, 0.45, 0.8, 0.5, 0.0, 0.0]


def test_novograd_matches_the_reference_optimizer_golden(golden):
	"""NovoGrad against parameter trajectories minted by the REFERENCE's own optimizers.NovoGrad (optimizers.py:66-90,
	oracle/make_golden.py: golden_misc): weight decay, dampening, three steps with growing gradients"""
	from convasr_b200 import optimizers
	dev = torch.device('cuda:0')
	for c in golden('misc')['novograd']:
		mine = [p.clone().to(dev).requires_grad_(True) for p in c['p0']]
		opt = optimizers.NovoGrad(mine, lr = c['lr'], betas = c['betas'], weight_decay = c['weight_decay'], dampening = c['dampening'])
		for st in c['steps']:
			for q, gr in zip(mine, st['grads']):
				q.grad = gr.clone().to(dev)
			opt.step()
			for q, p in zip(mine, st['params']):
				assert torch.allclose(q.detach().cpu(), p, rtol = 2e-5, atol = 1e-6), (c['weight_decay'], c['dampening'])


def test_clip_norm_is_global_over_parameter_groups_and_survives_changes():
	"""clip_grad_norm_ is one norm over ALL parameters (train.py:776-779), also with several param groups; the
	cached tables follow a changing set of tensors with gradients and a load_state_dict after the first step"""
	from convasr_b200 import optimizers
	dev = torch.device('cuda:0')
	ps, g = _tensors(3)
	ref = [p.clone().requires_grad_(True) for p in ps]
	mine = [p.clone().to(dev).requires_grad_(True) for p in ps]
	groups = lambda t: [dict(params = t[:3], lr = 0.05), dict(params = t[3:], lr = 0.01, weight_decay = 1e-2)]
	opt_ref = torch.optim.SGD(groups(ref), lr = 0.05, momentum = 0.9)
	opt = optimizers.SGD(groups(mine), lr = 0.05, momentum = 0.9)
	for step in range(5):
		grads = [torch.randn(p.shape, generator = g) * 2.0 for p in ps]
		for i, (p, q, gr) in enumerate(zip(ref, mine, grads)):
			skip = step in (2, 3) and i == 1  # a tensor without a gradient for two steps, then back
			p.grad = None if skip else gr.clone()
			q.grad = None if skip else gr.clone().to(dev)
		total = torch.nn.utils.clip_grad_norm_(ref, 3.0)
		opt_ref.step()
		opt.step(max_grad_norm = 3.0)
		assert torch.allclose(opt.total_grad_norm.cpu(), total.reshape(1), rtol = 1e-5), step
		for p, q in zip(ref, mine):
			assert torch.allclose(q.detach().cpu(), p.detach(), rtol = 1e-5, atol = 1e-6), (step, p.shape)
		if step == 1:  # checkpoint round trip in the middle: momentum must carry over
			sd = opt.state_dict()
			opt = optimizers.SGD(groups(mine), lr = 0.05, momentum = 0.9)
			opt.load_state_dict(sd)
