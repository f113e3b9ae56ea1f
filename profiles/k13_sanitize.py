'''K13 at small shapes (tail rows, the resident-x run at K = 320, several tiles per CTA pair), for
   compute-sanitizer --tool memcheck|racecheck python profiles/k13_sanitize.py'''
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexdiffuse_b200 import _native  # noqa: E402

dev = torch.device('cuda:0')
for M, C in ((700, 320), (300, 640), (40000, 320)):
    x = torch.randn(M, C, device=dev).bfloat16()
    w = (torch.randn(8 * C, C, device=dev) * C ** -0.5).bfloat16()
    b = (torch.randn(8 * C, device=dev) * 0.1).bfloat16()
    got = _native.ff_geglu(x, w, b)
    v, g = (x.float() @ w.float().t() + b.float()).chunk(2, dim=-1)
    err = (got.float() - v * F.gelu(g)).abs().max().item()
    print(f'M={M} C={C}: max abs err {err:.3e}, flag {_native.lib().fd_debug_k13_flag()}')
