'''K11 (fd_linear_x3) at the CLIP tower shapes: split and GEMM time per call (graph-replayed, 20 calls per replay),
next to torch fp32 / TF32 matmul and the fp32 SDPA the towers use.  Development aid.'''
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from flexdiffuse_b200 import _native
dev = torch.device('cuda:0')

def timeit(fn, reps=20, n=7):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s): fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps): fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); b.synchronize(); ts.append(a.elapsed_time(b) * 1e3 / reps)
    ts.sort(); return ts[len(ts) // 2]

torch.manual_seed(0)
for M, N, K in [(257, 1024, 1024), (257, 3072, 1024), (257, 4096, 1024), (257, 1024, 4096), (77, 768, 768), (77, 3072, 768), (77, 768, 3072)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) * 0.03; b = torch.randn(N, device=dev)
    op = _native.x3_split(x)
    t_split = timeit(lambda: _native.x3_split(x))
    row = [f'M={M} N={N} K={K}: split {t_split:.1f} us']
    for sk in (1, 2, 4, 8):
        if K // 64 // sk < 2: continue
        row.append(f'gemm sk={sk} {timeit(lambda: _native.linear_x3(x, w, b, operand=op, split_k=sk)):.1f}')
    row.append(f'auto {timeit(lambda: _native.linear_x3(x, w, b, operand=op)):.1f}')
    torch.backends.cuda.matmul.allow_tf32 = False
    row.append(f'| torch fp32 {timeit(lambda: torch.addmm(b, x, w.t())):.1f}')
    torch.backends.cuda.matmul.allow_tf32 = True
    row.append(f'tf32 {timeit(lambda: torch.addmm(b, x, w.t())):.1f}')
    torch.backends.cuda.matmul.allow_tf32 = False
    print(' '.join(row), flush=True)
from torch.nn.functional import scaled_dot_product_attention as sdpa, layer_norm
q = torch.randn(1, 16, 257, 64, device=dev)
print('sdpa fp32 [1,16,257,64]: %.1f us' % timeit(lambda: sdpa(q, q, q)))
q = torch.randn(1, 12, 77, 64, device=dev)
print('sdpa fp32 causal [1,12,77,64]: %.1f us' % timeit(lambda: sdpa(q, q, q, is_causal=True)))
h = torch.randn(1, 257, 1024, device=dev); g = torch.randn(1024, device=dev)
print('layer_norm [257,1024]: %.1f us; add: %.1f us' % (timeit(lambda: layer_norm(h, (1024,), g, g)), timeit(lambda: h + h)))
