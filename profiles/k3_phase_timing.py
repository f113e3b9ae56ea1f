'''K3 trace (%globaltimer): per-tile events of one CTA in a single launch + start/end/SM of every CTA.'''
import sys, ctypes, collections; sys.path.insert(0,'/root/repo')
import torch
from flexdiffuse_b200 import _native
dev=torch.device('cuda:0')
lib=_native.lib()
buf=torch.zeros(160+3*8192,dtype=torch.int64,device=dev)
lib.fd_debug_set_k3_timing.argtypes=[ctypes.c_void_p]
ev=['waitQ','Sissue','waitS','Sseen','Pwr','PVissue','Oseen','epiEnd']
shapes=[(8,4096,320),(2,4096,320),(8,1024,640),(8,256,1280)]
if len(sys.argv)>1: shapes=shapes[:int(sys.argv[1])]
for (S,nq,C) in shapes:
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
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print((S,nq,C),'graph-replayed avg us per launch', round(e0.elapsed_time(e1)*1e3/20,2))
    for cta in (0, 200):
        buf.zero_()
        lib.fd_debug_set_k3_trace_cta(cta)
        lib.fd_debug_set_k3_timing(buf.data_ptr())
        _native.cross_attn(q,kv,0,C,idx,heads,77,80,(C//heads)**-0.5)
        torch.cuda.synchronize()
        lib.fd_debug_set_k3_timing(None)
        t=buf.cpu()
        tr=t[160:].view(-1,3); tr=tr[tr[:,0]>0]
        per_sm=collections.Counter(tr[:,2].tolist())
        if cta>=len(tr): continue
        t0=int(t[160+3*cta])
        print('  CTA',cta,'on SM',int(t[160+3*cta+2]),'co-resident CTAs',per_sm[int(t[160+3*cta+2])],'lifetime us',(int(t[160+3*cta+1])-t0)/1e3)
        for tile in range(16):
            row=t[32+8*tile:32+8*tile+8].tolist()
            if not any(row): break
            print('    tile',tile,' '.join('%s=%5.2f'%(n,(v-t0)/1e3) if v else '%s=  -  '%n for n,v in zip(ev,row)))
    lib.fd_debug_set_k3_trace_cta(0)
    t0=int(tr[:,0].min())
    st=(tr[:,0]-t0).float()/1e3; en=(tr[:,1]-t0).float()/1e3
    print('  CTAs',len(tr),'SMs used',len(per_sm),'CTAs/SM hist',sorted(collections.Counter(per_sm.values()).items()))
    print('  start us: min %.2f med %.2f max %.2f | end us: min %.2f med %.2f max %.2f | dur med %.2f max %.2f'%(st.min(),st.median(),st.max(),en.min(),en.median(),en.max(),(en-st).median(),(en-st).max()))
    for k in (1,2):
        d=[(e-s_) for s_,e,m in zip(st.tolist(),en.tolist(),tr[:,2].tolist()) if per_sm[m]==k]
        if d: print('  %d CTA(s) on the SM: n=%d mean lifetime %.2f us'%(k,len(d),sum(d)/len(d)))
