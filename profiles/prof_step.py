'''Profiling target for ncu (run under gpurun, never a bench number):

  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
      --csv --log-file gpurun_out/launches.csv python profiles/prof_step.py step
  ncu --set full --clock-control none --import-source on --profile-from-start off \
      -k regex:'k[0-9]+f?_' -o gpurun_out/kernels python profiles/prof_step.py kernels

`step`    : ONE denoising step of bench.py's workload (SD v1.5 UNet 512^2, B=1, CFG -> 2
            sample-forwards, eager so every kernel is named) + the fused K4 DDIM update.
`kernels` : the hand-written kernels at the sizes bench.py's roofline section uses.
'''
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from flexdiffuse_b200 import _native, factory, schedulers  # noqa: E402
from flexdiffuse_b200.pipeline.guide import SimpleGuide  # noqa: E402


class Enc:
    def __init__(self, u):
        self.u = u

    def prompt(self, p):
        return self.u


def main(mode: str):
    dev = torch.device('cuda:0')
    torch.backends.cudnn.benchmark = True
    unet = factory.build_unet(dev, torch.bfloat16, seed=0)
    g = torch.Generator(device=dev).manual_seed(0)
    uncond = torch.randn(1, 77, 768, device=dev, generator=g)
    B = int(os.environ.get('FD_PROF_BATCH', '1'))   # images per step (FD_PROF_BATCH=8: the batched regime of configs[2..4])
    embeds = torch.randn(B, 77, 768, device=dev, generator=g)
    if mode == 'step':
        guide = SimpleGuide(Enc(uncond), unet, 7.5, 50, embeds, use_cuda_graph=False)
        sched = schedulers.DDIMScheduler()
        sched.set_timesteps(50)
        x = torch.randn(B, 4, 64, 64, device=dev, generator=g)
        buf = guide.model_input_buffer(x)
        buf.copy_(x)
        ts = [int(t) for t in sched.timesteps]
        for i in range(3):  # warm-up: cuDNN autotune, allocator
            u, c = guide.noise_pred_pair(buf, ts[i])
            x = sched.fused_step(u, c, 7.5, True, ts[i], x, scaled_out=buf).prev_sample
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        u, c = guide.noise_pred_pair(buf, ts[3])
        x = sched.fused_step(u, c, 7.5, True, ts[3], x, scaled_out=buf).prev_sample
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return
    # ---- kernels
    S = 8
    ctx = torch.randn(2, 77, 768, device=dev, generator=g)
    ctx9 = torch.randn(9, 77, 768, device=dev, generator=g)
    kv = unet.build_kv_cache(ctx)
    idx = torch.tensor([0] * 4 + [1] * 4, dtype=torch.int32, device=dev)
    mods = unet.cross_attentions()
    picks = [mods[0], mods[2], mods[4]]  # C = 320 / 640 / 1280
    qs = [torch.randn(S, {320: 4096, 640: 1024, 1280: 256}[m.dim], m.dim, device=dev,
                      generator=g).bfloat16() for m in picks]
    n = 1024 * 4 * 64 * 64
    u, c, x = (torch.randn(n, device=dev, generator=g) for _ in range(3))
    xo = torch.empty_like(x)
    k = _native.SchedCoeffs()
    k.guidance, k.use_cfg, k.a, k.b = 7.5, 1, 1.01, -0.05
    k.w[0] = 1.0
    sys.path.insert(0, ROOT)
    import bench  # the same planted K1 inputs as bench.py's roofline section
    txt_h, img_h = bench._planted_pair(1024)
    txt, img = txt_h.to(dev), img_h.to(dev)
    prm = _native.TweenParams()
    prm.threshold_floor = prm.threshold_mult = prm.max_guidance = prm.clustered = 0.5
    prm.header_max, prm.align_mode, prm.mapping_reuse = 0.15, 1, 1
    lin = torch.linspace(0.0, 0.5, 77)[None].to(dev)
    # K3F at bench.py's sample count (32 = configs[3]'s batch 16 with CFG), one site per width
    S3 = 32
    idx3 = torch.tensor([0] * 16 + [1] * 16, dtype=torch.int32, device=dev)
    xs = [torch.randn(S3, {320: 4096, 640: 1024, 1280: 256}[m.dim], m.dim, device=dev,
                      generator=g).bfloat16() for m in picks]
    outs = [torch.empty_like(t) for t in xs]
    scr = [torch.empty_like(t) if m.dim != 320 else None for m, t in zip(picks, xs)]
    dec = torch.randn(1, 3, 512, 512, device=dev, generator=g).bfloat16().contiguous(
        memory_format=torch.channels_last)

    tx = torch.randn(257, 1024, device=dev, generator=g)
    tg, tb = torch.ones(1024, device=dev), torch.zeros(1024, device=dev)
    w_qkv, b_qkv = torch.randn(3072, 1024, device=dev, generator=g) * 0.03, torch.zeros(3072, device=dev)
    w_o, b_o = torch.randn(1024, 1024, device=dev, generator=g) * 0.03, torch.zeros(1024, device=dev)
    w_fc1, b_fc1 = torch.randn(4096, 1024, device=dev, generator=g) * 0.03, torch.zeros(4096, device=dev)
    w_fc2 = torch.randn(1024, 4096, device=dev, generator=g) * 0.03

    # K13 at bench.py's sample count, one GEGLU projection per width
    ffs = []
    for m in picks:
        C = m.dim
        ffs.append((torch.randn(S3 * {320: 4096, 640: 1024, 1280: 256}[C], C, device=dev, generator=g).bfloat16(),
                    (torch.randn(8 * C, C, device=dev, generator=g) * C ** -0.5).bfloat16(),
                    (torch.randn(8 * C, device=dev, generator=g) * 0.1).bfloat16()))

    def run():
        for xf, wf, bf in ffs:
            _native.ff_geglu(xf, wf, bf)
        for m, q in zip(picks, qs):
            _native.cross_attn(q, kv.kv, m.k_col_off, m.v_col_off, idx, m.heads, 77, 80,
                               m.scale)
        for m, t, o, sc in zip(picks, xs, outs, scr):
            lin_o = m.to_out[0]
            _native.cross_attn_fused(t, m.to_q.weight, kv.kv, m.k_col_off, m.v_col_off, idx3,
                                     lin_o.weight, lin_o.bias, m.heads, 77, 80, m.scale, attn=sc,
                                     out=o, want_attn=False)
        _native.cfg_sched_step(u, c, x, k, xo)
        unet.build_kv_cache(ctx9)
        _native.sim_blend(txt, img, [prm], lin)
        _native.image_tail_u8(dec)
        # K11 / K12 at the ViT-L/14 shapes (M = 257): LN-fused split, q|k|v GEMM, attention core, out_proj (split-K
        # cluster + residual), fc1 + quick_gelu, plain split, fc2 (split-K 4)
        op = _native.x3_split_ln(tx, tg, tb, 1e-5)
        qkv = _native.linear_x3(None, w_qkv, b_qkv, operand=op, rows=257).view(1, 257, 3072)
        o = _native.attention_f32(qkv[..., :1024], qkv[..., 1024:2048], qkv[..., 2048:], 16, 0.125, False)
        h1 = _native.linear_x3(o.view(257, 1024), w_o, b_o, residual=tx)
        h2 = _native.linear_x3(h1, w_fc1, b_fc1, act=_native.LINEAR_ACT_QUICK_GELU)
        _native.linear_x3(h2, w_fc2, b_o, residual=h1)

    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else 'step')
