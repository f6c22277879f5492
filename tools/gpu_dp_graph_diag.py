"""2-GPU diagnostic: CUDA-graph capture of the data-parallel training step under different histories (VARIANT env)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch.distributed as dist
from convasr_b200 import models, optimizers, parallel, training
import test_gpu_training as T

variant = os.environ.get('VARIANT', 'A')
rank, world, local = parallel.init_from_env()
dev = torch.device('cuda', local)
C = 38
sig, xlen, y, ylen = [t.to(dev) for t in T._batch(C, seed = 3 + rank)]
kw = dict(base_width = 32, num_blocks = 1) if variant != 'D' else dict(base_width = 128)
m, sd = T._model(dev, kw, precision = 'bf16')
net, _ = models.distributed_data_parallel_and_autocast(m, rank, opt_level = 'O2')
opt = optimizers.SGD([p for p in m.parameters() if p.requires_grad], lr = 1e-3, momentum = 0.9)
if variant in ('B', 'C', 'B2', 'B4'):
	out = net(sig, xlen, y = y, ylen = ylen)
	(out['loss'] * ylen[:, 0]).mean().backward()
	torch.cuda.synchronize()
if variant == 'B3':
	side = torch.cuda.Stream()
	side.wait_stream(torch.cuda.current_stream())
	with torch.cuda.stream(side):
		out = net(sig, xlen, y = y, ylen = ylen)
		(out['loss'] * ylen[:, 0]).mean().backward()
	torch.cuda.synchronize()
if variant == 'C':
	m.zero_grad(set_to_none = True)
if variant == 'B2':
	del out
	import gc; gc.collect()
if variant == 'B4':
	opt.step(max_grad_norm = 100.0); opt.zero_grad(set_to_none = True); del out
	torch.cuda.synchronize()
try:
	step = training.GraphedTrainStep(net, opt, sig, xlen, y, ylen, warmup = 3 if variant == 'E' else 2, max_grad_norm = 100.0)
	for _ in range(3):
		loss = step(sig, xlen, y, ylen)
	torch.cuda.synchronize()
	print(f'variant {variant} rank {rank}: capture + replay ok, loss finite {bool(torch.isfinite(loss).all())}', flush = True)
except Exception as e:
	print(f'variant {variant} rank {rank}: FAILED {type(e).__name__}: {str(e)[:200]}', flush = True)
os._exit(0)
