'''(10 back-to-back launches per graph replay: the first sees a cold L2, the rest a warm one)
K2 variants (1 = v3 cta_group::1 + multicast pairs, 2 = v4 cta_group::2, 3 = v5 six-CTA clusters with the weight tile multicast
over three pairs when M > 512) against torch.matmul (cuBLAS) on
the K/V projection shapes: parity, the bounded-wait record, graph-replayed time with a cold L2.'''
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flexdiffuse_b200 import _native
dev = torch.device('cuda:0')
lib = _native.lib()
flush = torch.empty(96 * 1024 * 1024, dtype=torch.float32, device=dev)
def timeit(fn, n=10):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        for _ in range(10): fn()
    ts = []
    for _ in range(n):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b) * 1e2)
    ts.sort(); return ts[len(ts) // 2]
def timeit_cold(fn, n=15):
    '''one launch per replay, L2 flushed before each (bench.py's K2 timing)'''
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): fn()
    torch.cuda.current_stream().wait_stream(s)
    with torch.cuda.graph(g):
        fn()
    ts = []
    for _ in range(n):
        flush.add_(1.0)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b) * 1e3)
    ts.sort(); return ts[len(ts) // 2]
torch.manual_seed(0)
N, K = 24960, 768
w = (torch.randn(N, K, device=dev) * K ** -0.5).bfloat16()
for n_ctx in (9, 17, 1, 2, 3):
    M = n_ctx * 80
    x = torch.randn(M, K, device=dev).bfloat16()
    want = x.float() @ w.float().t()
    row = [f'M={M}']
    for v in (2, 3, 1):
        lib.fd_debug_set_k2_variant(v)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        _native.kv_project(x, w, out=out); torch.cuda.synchronize()
        err = ((out.float() - want).norm() / want.norm()).item()
        err = max(err, float((out.float() - want).abs().max() > 0.25))   # any single wrong element shows as >= 1
        flag = lib.fd_debug_k2_flag()
        t = timeit(lambda: _native.kv_project(x, w, out=out))
        row.append(f'v{v}: {t:.1f} us (cold {timeit_cold(lambda: _native.kv_project(x, w, out=out)):.1f}) err {err:.1e} flag {flag}')
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    row.append(f'cuBLAS: {timeit(lambda: torch.matmul(x, w.t(), out=out)):.1f} us (cold {timeit_cold(lambda: torch.matmul(x, w.t(), out=out)):.1f})')
    print(' | '.join(row), flush=True)
lib.fd_debug_set_k2_variant(2)
