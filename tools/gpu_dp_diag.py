"""2-GPU diagnostic: where do the data-parallel gradients differ from the mean of per-rank gradients?"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import torch.distributed as dist
from convasr_b200 import models, parallel
import test_gpu_training as T

rank, world, local = parallel.init_from_env()
dev = torch.device('cuda', local)
C = 38
batches = [[t.to(dev) for t in T._batch(C, seed = 3 + k)] for k in range(world)]
for precision in ('fp32', 'bf16'):
	m, sd = T._model(dev, dict(base_width = 32, num_blocks = 1), precision = precision)
	def grads_of(k):
		m.load_state_dict(sd, strict = False); m.zero_grad(set_to_none = True)
		sig, xlen, y, ylen = batches[k]
		out = m(sig, xlen, y = y, ylen = ylen)
		(out['loss'] * ylen[:, 0]).mean().backward()
		return {n: p.grad.clone() for n, p in m.named_parameters() if p.grad is not None}, out['logits'][0].detach().clone()
	a, la = grads_of(rank)
	b, lb = grads_of(rank)
	same = max(T.rel(a[n], b[n]) for n in a)
	# my own gradient of my batch vs the other rank's gradient of the same batch
	mine = torch.cat([a[n].flatten() for n in sorted(a)])
	both = [torch.empty_like(mine) for _ in range(world)]
	other_first = {}
	g0, l0 = grads_of(0)
	flat0 = torch.cat([g0[n].flatten() for n in sorted(g0)])
	dist.all_gather(both, flat0)
	cross = T.rel(both[0], both[1])
	lg = [torch.empty_like(l0) for _ in range(world)]
	dist.all_gather(lg, l0)
	print(f'rank {rank} {precision}: same GPU twice {same:.3e}; batch 0 on GPU0 vs GPU1: grads {cross:.3e} logits {T.rel(lg[0], lg[1]):.3e}', flush = True)
	if rank == 0:
		off = 0
		for n in sorted(g0):
			k = g0[n].numel()
			e = T.rel(both[0][off:off + k], both[1][off:off + k])
			if e > 1e-5: print('   ', n, f'{e:.3e}')
			off += k
dist.barrier(); dist.destroy_process_group()
