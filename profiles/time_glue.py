'''Development aid: CUDA-graph-replayed timings of the glue kernels (K5/K6/K7/K8) per UNet shape.'''
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexdiffuse_b200 import _native  # noqa: E402


def graph_time(fn, reps=20):
    fn()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
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
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) * 1e3 / reps


def main():
    dev = torch.device('cuda:0')
    for (N, C, H) in [(2, 320, 64), (2, 640, 64), (2, 960, 64), (2, 640, 32), (2, 1280, 32),
                      (2, 1920, 32), (2, 1280, 16), (2, 2560, 16), (2, 1280, 8), (2, 2560, 8),
                      (16, 320, 64), (16, 640, 64), (16, 960, 64), (16, 640, 32), (16, 1280, 32), (16, 1280, 16), (16, 2560, 16),
                      (32, 320, 64), (1, 128, 512)]:
        x = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(
            memory_format=torch.channels_last)
        w = torch.ones(C, device=dev).bfloat16()
        b = torch.zeros(C, device=dev).bfloat16()
        t = graph_time(lambda: _native.groupnorm_act(x, w, b, 32, 1e-5, True))
        gn = torch.nn.GroupNorm(32, C).to(dev).bfloat16()
        t_ref = graph_time(lambda: F.silu(gn(x)))
        mb = 2 * x.numel() * 2 / 1e6
        print(f'K5 N={N} C={C} HxW={H}x{H}: {t:7.2f} us  ({mb / t * 1e6 / 1e6:6.2f} TB/s r+w)   '
              f'torch GroupNorm+SiLU {t_ref:7.2f} us')


def cat_main():
    '''K15 against torch.cat on the up blocks' skip concats (channels-last bf16), N = 2 and 16 samples.'''
    dev = torch.device('cuda:0')
    for N in (2, 16):
        for Ca, Cb, H in ((1280, 1280, 8), (1280, 1280, 16), (1280, 640, 16), (1280, 640, 32), (640, 640, 32), (640, 320, 32),
                          (640, 320, 64), (320, 320, 64)):
            a = torch.randn(N, Ca, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
            b = torch.randn(N, Cb, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
            t = graph_time(lambda: _native.concat_channels(a, b))
            t_ref = graph_time(lambda: torch.cat([a, b], dim=1))
            mb = 2 * (a.numel() + b.numel()) * 2 / 1e6
            print(f'K15 N={N} {Ca}+{Cb} {H}x{H}: {t:7.2f} us ({mb / t * 1e6 / 1e6:5.2f} TB/s r+w)   torch.cat {t_ref:7.2f} us')


def up_main():
    dev = torch.device('cuda:0')
    for N in (2, 16):
        for C, H in ((1280, 8), (1280, 16), (640, 32)):
            x = torch.randn(N, C, H, H, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
            t = graph_time(lambda: _native.upsample_nearest2x(x))
            t_ref = graph_time(lambda: F.interpolate(x, scale_factor=2.0, mode='nearest'))
            print(f'K15 upsample N={N} C={C} {H}x{H}: {t:7.2f} us ({5 * x.numel() * 2 / t / 1e6:5.2f} TB/s r+w)   F.interpolate {t_ref:7.2f} us')


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == 'up':
        up_main()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == 'cat':
        cat_main()
        sys.exit(0)
    main()
