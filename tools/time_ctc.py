import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from convasr_b200 import ops, _lib
dev='cuda'
g=torch.Generator().manual_seed(0)
B,C,T,L=80,38,753,169
lp=torch.randn(B,C,T,generator=g).log_softmax(1).to(dev)
y=torch.randint(0,C-1,(B,L),generator=g).to(dev)
ylen=torch.randint(40,L+1,(B,),generator=g).to(dev)
olen=torch.randint(380,T+1,(B,),generator=g).to(dev); olen[0]=T
def run(grad):
    l=lp.clone().requires_grad_(grad)
    nll=ops.ctc_loss(l.permute(2,0,1),y,olen,ylen,blank=C-1)
    if grad: nll.sum().backward()
for grad in (False, True):
    for _ in range(3): run(grad)
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): run(grad)
    e1.record(); torch.cuda.synchronize()
    print(os.environ.get('CONVASR_B200_LIB','default'), 'grad' if grad else 'fwd ', f'{e0.elapsed_time(e1)/10*1e3:.1f} us per call')
