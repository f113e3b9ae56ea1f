'''One launch of each glue kernel (K5 cluster / streaming paths, K7, K8) at the shapes of a B = 8 and a B = 1 CFG forward, between
cudaProfilerStart / Stop:  ncu --set full --import-source on --profile-from-start off -k regex:'k[5-8]_' python profiles/glue_one.py'''
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexdiffuse_b200 import _native  # noqa: E402

dev = torch.device('cuda:0')
cl = lambda t: t.contiguous(memory_format=torch.channels_last)
cases = []
for N in (16, 2):
    for C, H in ((320, 64), (640, 32), (1280, 16)):
        x = cl(torch.randn(N, C, H, H, device=dev).bfloat16())
        g, b = torch.ones(C, device=dev).bfloat16(), torch.zeros(C, device=dev).bfloat16()
        tb = torch.randn(N, C, device=dev).bfloat16()
        cases.append(lambda x=x, g=g, b=b, tb=tb: _native.groupnorm_act(x, g, b, 32, 1e-5, True, tb))
        h = cl(torch.randn(N, C, H, H, device=dev).bfloat16())
        cases.append(lambda x=x, h=h, b=b: _native.add_bias_residual(x, h, b))
        xl, yl = torch.randn(N * H * H, C, device=dev).bfloat16(), torch.randn(N * H * H, C, device=dev).bfloat16()
        cases.append(lambda xl=xl, yl=yl, g=g, b=b: _native.add_layernorm(xl, yl, g, b, 1e-5))
for f in cases:
    f()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for f in cases:
    f()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
