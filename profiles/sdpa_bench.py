'''Which SDPA backend is fastest for the UNet's self-attention shapes (B=1 CFG: 2 samples x 8 heads)?'''
import sys; sys.path.insert(0,'/root/repo')
import torch, torch.nn.functional as F
from torch.nn.attention import sdpa_kernel, SDPBackend
dev=torch.device('cuda:0')
def timeit(fn,n=20):
    fn(); torch.cuda.synchronize()
    g=torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record(); g.replay(); b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b)*1e3/n
for (N,d) in [(4096,40),(1024,80),(256,160),(64,160)]:
    q,k,v=(torch.randn(2,8,N,d,device=dev).bfloat16() for _ in range(3))
    res=[]
    for name,be in [('cudnn',SDPBackend.CUDNN_ATTENTION),('flash',SDPBackend.FLASH_ATTENTION),('efficient',SDPBackend.EFFICIENT_ATTENTION)]:
        try:
            with sdpa_kernel(be):
                t=timeit(lambda: F.scaled_dot_product_attention(q,k,v))
            res.append('%s %.1f us'%(name,t))
        except Exception as e:
            res.append('%s n/a (%s)'%(name,str(e)[:40]))
    # padded head dim 64 for d=40
    if d==40:
        qp,kp,vp=(F.pad(t_,(0,24)) for t_ in (q,k,v))
        for name,be in [('cudnn-pad64',SDPBackend.CUDNN_ATTENTION),('flash-pad64',SDPBackend.FLASH_ATTENTION)]:
            try:
                with sdpa_kernel(be):
                    t=timeit(lambda: F.scaled_dot_product_attention(qp,kp,vp,scale=d**-0.5))
                res.append('%s %.1f us'%(name,t))
            except Exception as e:
                res.append('%s n/a'%name)
    t=timeit(lambda: F.scaled_dot_product_attention(q,k,v))
    res.append('default %.1f us'%t)
    print((N,d),' | '.join(res))
