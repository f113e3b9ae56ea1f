'''One K13 launch per feed-forward width between cudaProfilerStart / Stop (16 samples), for
ncu --set full --import-source on --profile-from-start off -k regex:k13_ python profiles/k13_one.py'''
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexdiffuse_b200 import _native  # noqa: E402

dev = torch.device('cuda:0')
cases = []
for C, N in ((320, 4096), (640, 1024), (1280, 256)):
    M = 16 * N
    cases.append(((torch.randn(M, C, device=dev)).bfloat16(), (torch.randn(8 * C, C, device=dev) * C ** -0.5).bfloat16(),
                  (torch.randn(8 * C, device=dev) * 0.1).bfloat16()))
for x, w, b in cases:
    _native.ff_geglu(x, w, b)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for x, w, b in cases:
    _native.ff_geglu(x, w, b)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
