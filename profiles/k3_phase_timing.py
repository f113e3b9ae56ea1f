import sys, ctypes; sys.path.insert(0,'/root/repo')
import torch
from flexdiffuse_b200 import _native, factory
dev=torch.device('cuda:0')
lib=_native.lib()
buf=torch.zeros(8,dtype=torch.int64,device=dev)
lib.fd_debug_set_k3_timing.argtypes=[ctypes.c_void_p]
for (S,nq,C) in [(2,4096,320),(8,4096,320),(2,256,1280),(2,1024,640)]:
    heads=8
    q=torch.randn(S,nq,C,device=dev).bfloat16()
    kv=torch.randn(2*80,2*C,device=dev).bfloat16()
    idx=torch.zeros(S,dtype=torch.int32,device=dev)
    for _ in range(3): _native.cross_attn(q,kv,0,C,idx,heads,77,80,(C//heads)**-0.5)
    torch.cuda.synchronize()
    lib.fd_debug_set_k3_timing(buf.data_ptr())
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); _native.cross_attn(q,kv,0,C,idx,heads,77,80,(C//heads)**-0.5); e1.record()
    torch.cuda.synchronize()
    lib.fd_debug_set_k3_timing(None)
    t=buf.cpu().tolist()
    print((S,nq,C),'event us',round(e0.elapsed_time(e1)*1e3,1),'phases ns',[t[i]-t[0] for i in range(8)])
    # 10 back-to-back launches
    e0.record()
    for _ in range(10): _native.cross_attn(q,kv,0,C,idx,heads,77,80,(C//heads)**-0.5)
    e1.record(); torch.cuda.synchronize()
    print('   10 launches avg us', round(e0.elapsed_time(e1)*100,1))
