"""Context number for DESIGN.md: the SAME module tree and training step through stock PyTorch CUDA kernels
(cuDNN convolutions, ATen BatchNorm / hardtanh / ctc_loss, torch.optim.SGD) on the same B200 -- what the
reference's own GPU path does -- next to the native step.  Not part of bench.py (library kernels)."""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from convasr_b200 import models
from oracle import oracle as O

dev = torch.device('cuda:0')
B, seconds, C = 80, 15.0, 38
sig, xlen, y, ylen = [t.to(dev) for t in bench.synth_batch(B, seconds, C, seed = 1000, lengths = 'full')]
flush = torch.empty(256 << 20, dtype = torch.uint8, device = dev)


def build():
	m = models.Wav2Letter(64, [C], frontend = models.LogFilterBankFrontend(64, 8000, .02, .01, 'hann_window'), dropout = 0., check_time_dim_padded = False)
	shapes = {k: tuple(v.shape) for k, v in m.state_dict().items() if not k.startswith('frontend.')}
	m.load_state_dict(O.synth_state_dict(shapes, seed = 0), strict = False)
	return m.to(dev).train()


def run(name, native, autocast, steps = 8):
	m = build()
	m.native_training = native
	opt = torch.optim.SGD(m.parameters(), lr = 1e-6, momentum = 0.9, weight_decay = 1e-3)

	def step():
		opt.zero_grad(set_to_none = True)
		with torch.autocast('cuda', dtype = torch.bfloat16, enabled = autocast):
			out = m(sig, xlen, y = y, ylen = ylen)
		(out['loss'] * ylen[:, 0]).mean().backward()
		torch.nn.utils.clip_grad_norm_(m.parameters(), 100.0)
		opt.step()

	for _ in range(3):
		step()
	torch.cuda.synchronize()
	ts = []
	for _ in range(steps):
		flush.zero_()
		e0, e1 = torch.cuda.Event(enable_timing = True), torch.cuda.Event(enable_timing = True)
		e0.record(); step(); e1.record(); torch.cuda.synchronize()
		ts.append(e0.elapsed_time(e1))
	ms = sum(ts) / len(ts)
	print(f'{name:58s} {ms:8.2f} ms/step  {B * seconds / ms * 1e3:10.0f} audio-s/s', flush = True)
	del m, opt
	torch.cuda.empty_cache()


torch.backends.cudnn.benchmark = True
run('native kernels (eager, torch.optim.SGD + clip)', True, False)
run('stock PyTorch: cuDNN/ATen, bf16 autocast', False, True)
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
run('stock PyTorch: cuDNN/ATen, fp32 with TF32 convolutions', False, False)
