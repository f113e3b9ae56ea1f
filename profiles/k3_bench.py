'''Graph-replayed K3 timing per attn2 shape (8 and 2 samples): min / median of 7 replays of 20 launches.'''
import sys, os; sys.path.insert(0,'/root/repo')
import torch
from flexdiffuse_b200 import _native
dev=torch.device('cuda:0')
res=[]
for (S,nq,C) in [(8,4096,320),(8,1024,640),(8,256,1280),(8,64,1280),(2,4096,320),(2,1024,640),(2,256,1280)]:
    heads=8
    q=torch.randn(S,nq,C,device=dev).bfloat16()
    kv=torch.randn(2*80,2*C,device=dev).bfloat16()
    idx=torch.zeros(S,dtype=torch.int32,device=dev)
    for _ in range(3): _native.cross_attn(q,kv,0,C,idx,heads,77,80,(C//heads)**-0.5)
    torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20): _native.cross_attn(q,kv,0,C,idx,heads,77,80,(C//heads)**-0.5)
    g.replay(); torch.cuda.synchronize()
    ts=[]
    for _ in range(7):
        e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1)*1e3/20)
    ts.sort(); res.append('%.2f/%.2f'%(ts[0],ts[3]))
w=lambda a,b,c,d:(5*a+5*b+5*c+d)/16
print(os.environ.get('FD_LIB_PATH','default').split('/')[-1].ljust(24),' '.join(res))
