'''Development aid: K13 `fd_ff_geglu` against the path it replaces (cuBLAS bf16 Linear + K6), per UNet feed-forward shape
(graph-replayed, 10 launches per replay), and the captured UNet forward with K13 on / off.  Run under gpurun.'''
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexdiffuse_b200 import _native, factory, unet as unet_mod  # noqa: E402
from flexdiffuse_b200.pipeline.guide import SimpleGuide  # noqa: E402


def graph_time(fn, reps=10, replays=20):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(replays):
        g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * replays) * 1e3


class Enc:
    def __init__(self, u):
        self.u = u

    def prompt(self, p):
        return self.u


def main():
    dev = torch.device('cuda:0')
    for samples in (2, 16, 32):
        for C, N in ((320, 4096), (640, 1024), (1280, 256)):
            M = samples * N
            x = torch.randn(M, C, device=dev).bfloat16()
            w = (torch.randn(8 * C, C, device=dev) * C ** -0.5).bfloat16()
            b = (torch.randn(8 * C, device=dev) * 0.1).bfloat16()
            t_new = graph_time(lambda: _native.ff_geglu(x, w, b))
            t_old = graph_time(lambda: _native.geglu(F.linear(x, w, b)))
            t_gemm = graph_time(lambda: F.linear(x, w, b))
            fl = 2.0 * M * C * 8 * C
            print(f'samples={samples:2d} C={C:4d} M={M:6d}: K13 {t_new:7.1f} us ({fl / t_new / 1e6:6.0f} TFLOP/s)  '
                  f'cuBLAS+K6 {t_old:7.1f} us  (cuBLAS alone {t_gemm:7.1f} us)  flag={_native.lib().fd_debug_k13_flag()}',
                  flush=True)
    unet = factory.build_unet(dev, torch.bfloat16, seed=0)
    for B in (1, 8):
        uncond = torch.randn(1, 77, 768, device=dev)
        embeds = torch.randn(B, 77, 768, device=dev)
        x = torch.randn(B, 4, 64, 64, device=dev)
        for fused in (False, True):
            unet_mod.FF_GEGLU_FUSED = fused
            unet.invalidate_derived()   # the captured graphs bake in the dispatch
            guide = SimpleGuide(Enc(uncond), unet, 7.5, 50, embeds, use_cuda_graph=True)
            buf = guide.model_input_buffer(x)
            buf.copy_(x)
            for _ in range(3):
                guide.noise_pred_pair(buf, 481)
            torch.cuda.synchronize()
            a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                guide.noise_pred_pair(buf, 481)
            e.record()
            torch.cuda.synchronize()
            print(f'UNet forward graph B={B} K13={fused}: {a.elapsed_time(e) / 20:.3f} ms', flush=True)
    unet_mod.FF_GEGLU_FUSED = True


if __name__ == '__main__':
    main()
