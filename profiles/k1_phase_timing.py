'''Development aid: phase timestamps of K1's CTA (0,0): [0] setup done, [1] GEMM done,
[2] softmax + P^T done, [3] mapping done, [4] weights done, [5] blend done (ns from [0]).'''
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flexdiffuse_b200 import _native
dev = torch.device('cuda:0')
lib = _native.lib()
lib.fd_debug_set_k1_timing.argtypes = [ctypes.c_void_p]
buf = torch.zeros(256, dtype=torch.int64, device=dev)
for nb, mode, reuse in [(1024, 1, 1)]:
    txt = torch.randn(nb, 77, 768, device=dev)
    img = torch.randn(1, 257, 768, device=dev)
    prm = _native.TweenParams()
    prm.threshold_floor = prm.threshold_mult = prm.max_guidance = 0.5
    prm.header_max, prm.align_mode, prm.mapping_reuse = 0.15, mode, reuse
    lin = torch.linspace(0.0, 0.5, 77)[None].to(dev)
    for _ in range(2):
        _native.sim_blend(txt, img, [prm], lin)
    torch.cuda.synchronize()
    lib.fd_debug_set_k1_timing(buf.data_ptr())
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); _native.sim_blend(txt, img, [prm], lin); b.record()
    torch.cuda.synchronize()
    lib.fd_debug_set_k1_timing(None)
    t = buf.cpu().tolist()
    print(f'prompts={nb} mode={mode} reuse={reuse}: kernel {a.elapsed_time(b)*1e3:.1f} us; '
          f'phases ns {[t[i] - t[0] for i in range(6)]} softmax-internal {[t[6]-t[0], t[7]-t[0]]}')
    if t[64 + 9]:  # batched kernel: per-batch events of CTA 0, ns from the first MMA commit's batch start
        t0 = min(x for x in t[64:64 + 16 * 8] if x)
        for bl in range(4):
            ev = t[64 + 16 * bl: 64 + 16 * bl + 10]
            if ev[9]:
                print('   batch', bl, {k: (ev[i] - t0 if ev[i] else None) for k, i in
                                       (('feed_done', 8), ('mma_done', 9), ('norms', 0), ('acc', 1), ('drained', 2),
                                        ('combined', 3), ('weights', 4), ('blend', 5))})
        if t[128]:
            c0 = t[128]
            print('   chunks of batch 1 (ns from the first guide request): guide_req, raw_seen, raw_req, feed_done, mma_full, mma_issued')
            for c in range(16):
                print('    ', c, [t[128 + 6 * c + e] - c0 if t[128 + 6 * c + e] else None for e in range(6)])
        print('   isolated MMA chunk: issue end -> commit visible (ns):', [t[225 + 2 * k] - t[224 + 2 * k] for k in range(4)])
    buf.zero_()
