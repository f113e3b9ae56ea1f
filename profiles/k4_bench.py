'''K4 timing (DDIM, fp32 eps, 1024 samples), L2 flushed between launches.'''
import sys, os; sys.path.insert(0,'/root/repo')
import torch
from flexdiffuse_b200 import _native
dev=torch.device('cuda:0')
n=1024*4*64*64
u,c,x,xo=(torch.randn(n,device=dev) for _ in range(4))
k=_native.SchedCoeffs(); k.guidance,k.use_cfg,k.a,k.b=7.5,1,0.98,-0.1; k.w[0]=1.0
flush=torch.empty(96*1024*1024,dtype=torch.float32,device=dev)
f=lambda: _native.cfg_sched_step(u.view(1024,4,64,64),c.view(1024,4,64,64),x.view(1024,4,64,64),k,xo.view(1024,4,64,64))
f(); torch.cuda.synchronize()
ts=[]
for _ in range(15):
    flush.zero_()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); f(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
ts.sort()
print(os.environ.get('FD_LIB_PATH','default').split('/')[-1], 'min %.2f med %.2f us -> %.0f GB/s'%(ts[0],ts[7],16*n/ts[7]/1e3))
