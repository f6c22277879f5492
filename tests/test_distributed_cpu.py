"""N > 1 host logic on CPU: world_size 2, gloo backend, 127.0.0.1 rendezvous."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from convasr_b200 import parallel


def _free_port():
	s = socket.socket()
	s.bind(('127.0.0.1', 0))
	port = s.getsockname()[1]
	s.close()
	return port


def test_shard_bounds_partition():
	for n in (0, 1, 7, 8, 8192, 8191):
		for world in (1, 2, 4, 8):
			ranges = [parallel.shard_bounds(n, r, world) for r in range(world)]
			assert ranges[0][0] == 0 and ranges[-1][1] == n
			assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
			sizes = [hi - lo for lo, hi in ranges]
			assert max(sizes) - min(sizes) <= 1


def test_shard_by_length_balances_audio():
	g = torch.Generator().manual_seed(0)
	lengths = (torch.rand(1001, generator = g) * 0.5 + 0.5).tolist()
	shards = parallel.shard_by_length(lengths, 8)
	assert sorted(i for s in shards for i in s) == list(range(1001))
	audio = [sum(lengths[i] for i in s) for s in shards]
	assert (max(audio) - min(audio)) / max(audio) < 0.02
	for s in shards:  # sorted by length inside a shard: little padding per micro-batch
		assert all(lengths[a] >= lengths[b] for a, b in zip(s, s[1:]))


def _worker(rank, world, port, tmp):
	os.environ.update(RANK = str(rank), WORLD_SIZE = str(world), LOCAL_RANK = str(rank), MASTER_ADDR = '127.0.0.1', MASTER_PORT = str(port))
	r, w, lr = parallel.init_from_env(device = 'cpu')
	assert (r, w) == (rank, world) and dist.get_backend() == 'gloo'
	# (1) utterance sharding + ordered gather of variable-length host results
	n = 11
	lo, hi = parallel.shard_bounds(n, rank, world)
	local = [[i] * (i % 4 + 1) for i in range(lo, hi)]
	full = parallel.gather_in_order(local, list(range(lo, hi)))
	assert full == [[i] * (i % 4 + 1) for i in range(n)]
	# (2) timing reduction used by bench.py: max over ranks
	assert parallel.max_over_ranks(1.0 + rank, device = 'cpu') == float(world)
	assert parallel.sum_over_ranks(1.0, device = 'cpu') == float(world)
	# (3) the drop-in wrapper (models.py:755-765) returns the module itself with a GradSync attached and rank 0's
	#     parameters everywhere; averaging a "gradient" per parameter through it gives the mean over ranks
	from convasr_b200 import models
	torch.manual_seed(rank)
	block = models.ConvBn1d(num_channels = (4, 6), kernel_size = 3, repeat = 2, nonlinearity = ('hardtanh', 0, 20))
	ddp, _ = models.distributed_data_parallel_and_autocast(block, rank)
	assert ddp is block and isinstance(block._grad_sync, parallel.GradSync)
	flat = torch.cat([p.detach().flatten() for p in block.parameters()])
	both = [torch.empty_like(flat) for _ in range(world)]
	dist.all_gather(both, flat)
	assert all(torch.equal(both[0], b) for b in both[1:])
	grads = [torch.full_like(p, float(rank + 1)) for p in block.parameters()]
	for g in grads:
		block._grad_sync.reduce(g)
	block._grad_sync.finish()
	assert all(torch.allclose(g, torch.full_like(g, sum(range(1, world + 1)) / world)) for g in grads)
	try:
		models.distributed_data_parallel_and_autocast(block, rank, synchronize_bn = True)
		raise AssertionError('synchronize_bn must be refused')
	except NotImplementedError:
		pass
	# (4) the native training step's gradient exchange (parallel.GradSync): asynchronous per-layer
	#     all-reduces issued in backward order, one flat buffer for the small tensors, then finish()
	lin = torch.nn.Linear(3, 2)
	with torch.no_grad():
		lin.weight.fill_(float(rank + 1))
	parallel.attach_grad_sync(lin)
	assert float(lin.weight[0, 0]) == 1.0  # rank 0's parameters everywhere, as the DDP constructor does
	sync = lin._grad_sync
	layers = [torch.full((5, 7), float(rank + 1 + k)) for k in range(3)]
	small = torch.arange(10, dtype = torch.float32) * (rank + 1)
	for t in layers:
		sync.reduce(t)
	sync.reduce(small)
	sync.finish()
	mean_rank = sum(range(1, world + 1)) / world
	for k, t in enumerate(layers):
		assert torch.allclose(t, torch.full((5, 7), mean_rank + k))
	assert torch.allclose(small, torch.arange(10, dtype = torch.float32) * mean_rank)
	assert sync.n_collectives == 4 and not sync.pending
	dist.barrier()
	dist.destroy_process_group()
	open(os.path.join(tmp, f'ok{rank}'), 'w').write('ok')


def test_world_size_2_gloo(tmp_path):
	port = _free_port()
	mp.spawn(_worker, args = (2, port, str(tmp_path)), nprocs = 2, join = True)
	assert sorted(os.listdir(tmp_path)) == ['ok0', 'ok1']
