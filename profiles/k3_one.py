'''One K3 launch per attn2 shape (8 samples) for `ncu --set full -k regex:k3_cross`.'''
import sys; sys.path.insert(0, '/root/repo')
import torch
from flexdiffuse_b200 import _native
dev = torch.device('cuda:0')
shapes = [(8, 4096, 320), (8, 1024, 640), (8, 256, 1280)]
if len(sys.argv) > 1:
    shapes = shapes[:int(sys.argv[1])]
for (S, nq, C) in shapes:
    q = torch.randn(S, nq, C, device=dev).bfloat16()
    kv = torch.randn(2 * 80, 2 * C, device=dev).bfloat16()
    idx = torch.zeros(S, dtype=torch.int32, device=dev)
    for _ in range(2):
        out = _native.cross_attn(q, kv, 0, C, idx, 8, 77, 80, (C // 8)**-0.5)
    torch.cuda.synchronize()
    print(S, nq, C, float(out.float().abs().mean()))
