import sys; sys.path.insert(0,'/root/repo')
import torch, bench
from flexdiffuse_b200 import _native
dev=torch.device('cuda:0')
txt_h,img_h=bench._planted_pair(1024); txt,img=txt_h.to(dev),img_h.to(dev)
prm=_native.TweenParams(); prm.threshold_floor=prm.threshold_mult=prm.max_guidance=prm.clustered=0.5
prm.header_max,prm.align_mode,prm.mapping_reuse=0.15,1,1
lin=torch.linspace(0.0,0.5,77)[None].to(dev)
for _ in range(2): _native.sim_blend(txt,img,[prm],lin)
torch.cuda.synchronize(); torch.cuda.profiler.start()
_native.sim_blend(txt,img,[prm],lin); torch.cuda.synchronize(); torch.cuda.profiler.stop()
