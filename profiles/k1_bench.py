'''K1 kernel time for 1024 prompts x 1 guide (default parameters and the two no-reuse modes).'''
import sys, os; sys.path.insert(0,'/root/repo')
import torch
from flexdiffuse_b200 import _native
dev=torch.device('cuda:0')
res=[]
for nb,mode,reuse in [(256,1,1),(1024,1,1),(2048,1,1),(4096,1,1),(1024,1,0),(1024,0,0),(8,1,1),(1,1,1)]:
    txt=torch.randn(nb,77,768,device=dev); img=torch.randn(1,257,768,device=dev)
    prm=_native.TweenParams(); prm.threshold_floor=prm.threshold_mult=prm.max_guidance=0.5
    prm.clustered,prm.header_max,prm.align_mode,prm.mapping_reuse=0.0,0.15,mode,reuse
    lin=torch.linspace(0.0,0.5,77)[None].to(dev)
    for _ in range(2): _native.sim_blend(txt,img,[prm],lin)
    torch.cuda.synchronize()
    ts=[]
    for _ in range(7):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); _native.sim_blend(txt,img,[prm],lin); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b)*1e3)
    ts.sort(); res.append('%d/m%d/r%d: %.0f us'%(nb,mode,reuse,ts[3]))
print(os.environ.get("FD_LIB_PATH","default").split("/")[-1], ' | '.join(res))
