'''K8 (residual add + LayerNorm) graph-replayed launch time per UNet token-matrix shape.'''
import sys, os; sys.path.insert(0,'/root/repo')
import torch
from flexdiffuse_b200 import _native
dev=torch.device('cuda:0')
res=[]
for M,C in [(8192,320),(2048,640),(512,1280),(128,1280)]:
    x=torch.randn(M,C,device=dev).bfloat16(); y=torch.randn(M,C,device=dev).bfloat16()
    w=torch.ones(C,device=dev).bfloat16(); b=torch.zeros(C,device=dev).bfloat16()
    f=lambda: _native.add_layernorm(x,y,w,b,1e-5)
    f(); torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): f()
    g.replay(); torch.cuda.synchronize()
    ts=[]
    for _ in range(7):
        a,e=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); e.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(e)*1e3/20)
    ts.sort(); res.append('%dx%d %.2f'%(M,C,ts[3]))
print(os.environ.get('FD_LIB_PATH','default').split('/')[-1], ' | '.join(res))
