'''Development aid: where does one denoising step go?  Times the captured UNet graph replay,
the eager forward, K4, and the VAE decode with CUDA events (run under gpurun).'''
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexdiffuse_b200 import _native, factory, schedulers  # noqa: E402
from flexdiffuse_b200.pipeline.flex import FlexPipeline  # noqa: E402
from flexdiffuse_b200.pipeline.guide import SimpleGuide  # noqa: E402


class Enc:
    def __init__(self, u):
        self.u = u

    def prompt(self, p):
        return self.u


def ev_time(fn, n):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    dev = torch.device('cuda:0')
    torch.backends.cudnn.benchmark = True
    unet = factory.build_unet(dev, torch.bfloat16, seed=0)
    vae = factory.build_vae(dev, torch.bfloat16, seed=1)
    uncond = torch.randn(1, 77, 768, device=dev)
    embeds = torch.randn(1, 77, 768, device=dev)
    x = torch.randn(1, 4, 64, 64, device=dev)
    for graph in (False, True):
        guide = SimpleGuide(Enc(uncond), unet, 7.5, 50, embeds, use_cuda_graph=graph)
        buf = guide.model_input_buffer(x)
        buf.copy_(x)
        for _ in range(3):
            guide.noise_pred_pair(buf, 481)
        t = ev_time(lambda: guide.noise_pred_pair(buf, 481), 20)
        t0 = time.perf_counter()
        for _ in range(20):
            guide.noise_pred_pair(buf, 481)
        host = (time.perf_counter() - t0) / 20 * 1e3
        torch.cuda.synchronize()
        print(f'UNet forward (B=1, CFG => 2 samples) graph={graph}: {t:.3f} ms GPU-timed, '
              f'{host:.3f} ms host issue time')
    sched = schedulers.DDIMScheduler()
    sched.set_timesteps(50)
    u, c = guide.noise_pred_pair(buf, 481)
    lat = x.clone()
    t = ev_time(lambda: sched.fused_step(u, c, 7.5, True, 481, lat, out=lat, scaled_out=buf), 50)
    print(f'K4 fused step incl. host: {t * 1e3:.1f} us')
    z = torch.randn(1, 4, 64, 64, device=dev)
    for _ in range(2):
        vae.decode(z)
    t = ev_time(lambda: vae.decode(z), 5)
    print(f'VAE decode 512x512: {t:.2f} ms')
    pipe = FlexPipeline(vae, None, None, unet, schedulers.DDIMScheduler())
    gen = torch.Generator(device=dev)
    def full():
        g = SimpleGuide(Enc(uncond), unet, 7.5, 50, embeds, use_cuda_graph=True)
        return pipe(g, generator=gen, output_type='pt', return_dict=False)
    full()
    t = ev_time(full, 3)
    print(f'full image (50 steps + decode): {t:.1f} ms')
    # kernel census of one eager forward
    from torch.profiler import profile, ProfilerActivity
    guide = SimpleGuide(Enc(uncond), unet, 7.5, 50, embeds, use_cuda_graph=False)
    guide.noise_pred_pair(buf, 481)
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        guide.noise_pred_pair(buf, 481)
        torch.cuda.synchronize()
    ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    tot = sum(e.device_time for e in ev)
    print(f'eager forward: {len(ev)} kernels, {tot / 1e3:.3f} ms summed kernel time')
    agg = {}
    for e in ev:
        k = e.name[:70]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += e.device_time
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        print(f'{t:9.1f} us n={n:4d} {k}')


if __name__ == '__main__':
    main()
