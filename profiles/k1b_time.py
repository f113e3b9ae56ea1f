'''K1B kernel time at bench.py's blend workload (planted pair, Linear + Clustered 0.5 + Threshold), median of 9.
FD_LIB_PATH selects an A/B variant built by profiles/build_variants.py.'''
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from flexdiffuse_b200 import _native
dev = torch.device('cuda:0')
res = []
for nb in (296, 592, 1024, 4096):
    txt_h, img_h = bench._planted_pair(nb)
    txt, img = txt_h.to(dev), img_h.to(dev)
    prm = _native.TweenParams()
    prm.threshold_floor = prm.threshold_mult = prm.max_guidance = prm.clustered = 0.5
    prm.header_max, prm.align_mode, prm.mapping_reuse = 0.15, 1, 1
    lin = torch.linspace(0.0, 0.5, 77)[None].to(dev)
    for _ in range(3):
        _native.sim_blend(txt, img, [prm], lin, want_maps=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(9):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _native.sim_blend(txt, img, [prm], lin, want_maps=False); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    res.append('%d prompts: %.0f us (min %.0f)' % (nb, ts[4], ts[0]))
print(os.environ.get('FD_LIB_PATH', 'default').split('/')[-1], ' | '.join(res))
