"""Times the activation-sized training kernels at the C2 shape (B=80, T=751) for a few widths."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convasr_b200 import _lib, ops, training
lib = _lib.load(); dev = 'cuda'
B, T = 80, 751
flush = torch.empty(256 << 20, dtype = torch.uint8, device = dev)
def timeit(fn, n = 10):
	for _ in range(2): fn()
	ts = []
	for _ in range(n):
		flush.zero_()
		e0, e1 = torch.cuda.Event(enable_timing = True), torch.cuda.Event(enable_timing = True)
		e0.record(); fn(); e1.record(); torch.cuda.synchronize()
		ts.append(e0.elapsed_time(e1) * 1e3)
	ts.sort(); return ts[len(ts) // 2]
for C in (256, 384, 512, 768, 1024):
	y = torch.randn(B, T, C, device = dev).to(torch.bfloat16)
	g = torch.randn(B, T, C, device = dev).to(torch.bfloat16)
	out = torch.empty_like(y); dy = torch.empty_like(y)
	gamma = torch.ones(C, device = dev); beta = torch.zeros(C, device = dev)
	rm = torch.zeros(C, device = dev); rv = torch.ones(C, device = dev)
	ws = torch.empty(2, C, device = dev); ss = torch.empty(4, C, device = dev); sums = torch.empty(2, C, device = dev); part = torch.empty(8, 2, C, device = dev)
	xlen = (torch.rand(B, device = dev) * 0.5 + 0.5)
	mb = y.numel() * 2 / 1e6
	t1 = timeit(lambda: lib.cab_bn_batch_stats(ops._p(y), B, T, C, C, ops._p(gamma), ops._p(beta), 1e-5, 0.1, ops._p(rm), ops._p(rv), ops._p(ws), ops._p(ss), ops._stream()))
	t2 = timeit(lambda: lib.cab_bn_act_mask_fwd(ops._p(y), ops._p(ss), B, T, C, C, 2, 0.0, 20.0, ops._p(xlen), ops._p(out), 0.0, None, 0, ops._stream()))
	t3 = timeit(lambda: lib.cab_bn_act_mask_bwd(ops._p(y), ops._p(g), ops._p(ss), B, T, C, C, 2, 0.0, 20.0, ops._p(xlen), ops._p(sums), ops._p(dy), 0.0, None, 0, ops._p(part), ops._stream()))
	w = torch.randn(C, C, 11, device = dev)
	t4 = timeit(lambda: training._pack(w, C, C, True))
	print(f'C={C:5d} act {mb:6.1f} MB | stats {t1:7.1f} us ({mb/t1*1e-3:.2f} TB/s) | fwd {t2:7.1f} us ({2*mb/t2*1e-3:.2f} TB/s) | bwd(reduce+apply) {t3:7.1f} us ({5*mb/t3*1e-3:.2f} TB/s) | pack k11 {t4:7.1f} us')
