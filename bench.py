#!/usr/bin/env python
'''bench.py -- FlexDiffuse hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config {1,2,3,4}]
    torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Metric (BASELINE.json): SD1.5 512^2 50-step images/s (+ guided-embed blends/s, % roofline).
Headline workload at every N: BASELINE.json configs[1] -- SD v1.5 UNet 512x512, 50-step DDIM,
CFG 7.5, text-only conditioning, batch 1 per GPU, random-init weights, synthetic prompt embeddings.

One bench "step" = one pass of the hot path over one batch: K2 K/V cache build, 50 denoising steps
(UNet forward over cached K/V with K3F / K3, then the fused CFG+scheduler K4), VAE decode (+ K10).

  value : whole-job images/s with every input already resident in HBM (device-timed)
  e2e   : same metric through the public API (SimpleGuide + FlexPipeline) with HOST inputs:
          per step the prompt / uncond embeddings and the initial noise are copied from pinned
          host memory and the decoded uint8 image is read back (those bytes are reported)
  roofline     : K3F, the fused attn2 layer (to_q + attention + to_out), all 16 sites of a UNet
                 forward at configs[3]'s sample count, timed live with CUDA events; K3 / K4 / K2 /
                 K1 / K5 alongside
  gpu_torch_baseline : the same UNet (same weights) through plain torch bf16 eager ops
                 (cuDNN / cuBLAS / SDPA), eager and CUDA-graphed: the kernel to beat on this box
  other_configs: configs[2] (guided, all three modes, batch 8), configs[3] (img2img 0.6, PNDM,
                 batch 16), configs[4] (parameter x seed sweep, micro-batch 16, strong scaling)
  cpu_baseline : the oracle port of the reference path (fp32, all host cores) on a bounded
                 sample of the same workload, extrapolated as stated in `sample`
  --impl reference : only that CPU arm, as its own JSON line
  --config N       : make configs[N] the headline line (full size; configs[4] = 1024 samples)
'''
from __future__ import annotations

import argparse
import collections
import json
import os
import subprocess
import sys
import threading
import time

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STEPS_PER_IMAGE = 50
GUIDANCE = 7.5
HW = 512
METRIC = 'sd15_512x512_50step_ddim_cfg7.5_images_per_s'
UNIT = 'images/s'
WORKLOAD = ('SD v1.5 UNet 512x512, 50-step DDIM, CFG 7.5, text-only conditioning, '
            'batch 1 per GPU, random-init weights (BASELINE.json configs[1])')
CONFIG_NAMES = {
    1: WORKLOAD,
    2: ('SD v1.5 512x512, image guidance (TEXT, ALIGN and DIRECT modes) blended into the 77-token '
        'context by K1, batch 8 per mode, cached cross-attn K/V, 50-step DDIM, CFG 7.5 '
        '(BASELINE.json configs[2])'),
    3: ('Img2Img strength 0.6 + image guidance, PNDM 50 steps (31 UNet evaluations), batch 16, bf16 '
        '(BASELINE.json configs[3])'),
    4: ('guidance-parameter x seed sweep (8 prompts x 8 Tweener parameter sets x seeds), micro-batch '
        '16, sharded over the ranks, one NCCL all-gather of the output latents '
        '(BASELINE.json configs[4])'),
}


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tensor=p['bf16_tflops'],
                    tensor_sustained=p.get('bf16_tflops_sustained'),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm=6650.0, tensor=1590.0, tensor_sustained=1400.0,
                source='fallback (B200_PROFILING.md)')


class ClockSampler:
    '''nvidia-smi clocks / throttle reasons sampled DURING the timed region.'''
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                 '--format=csv,noheader,nounits', '-lms', '200'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown',
                 'sw_power_cap']
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=mx or None,
                    samples=len(sm), reasons=sorted(reasons))


# ------------------------------------------------------------------ reference / CPU arm
def cpu_reference_arm(steps: int, warmup: int):
    '''The reference's path on the host cores: the oracle port (fp32 restatement of
    pipeline/flex.py + pipeline/guide.py over diffusers' UNet arithmetic).  Each bench step is
    a BOUNDED sample: 1 of the 50 DDIM steps at B=1 (2 UNet sample-forwards + CFG + scheduler
    update); images/s = 1 / (50 * t_step).  VAE decode (~1.5 % of the FLOPs) excluded.'''
    from flexdiffuse_b200.unet import UNet2DConditionModel
    from oracle import loop_oracle as lo
    from oracle import unet_oracle as U
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(0)
    with torch.no_grad():
        sd = {k: v.float() for k, v in UNet2DConditionModel().state_dict().items()}
        g = torch.Generator().manual_seed(0)
        uncond = torch.randn(1, 77, 768, generator=g)
        embeds = torch.randn(1, 77, 768, generator=g)
        x = torch.randn(1, 4, HW // 8, HW // 8, generator=g)
        sched = lo.DDIMScheduler()
        sched.set_timesteps(STEPS_PER_IMAGE)
        ts = [int(t) for t in sched.timesteps]

        def one(i):
            eps = lo.noise_pred(lambda l, t, c: U.unet_forward(sd, l, t, c), uncond,
                                embeds, GUIDANCE, x, ts[i % len(ts)])
            return sched.step(eps, ts[i % len(ts)], x).prev_sample

        for i in range(warmup):
            one(i)
        t0 = time.perf_counter()
        for i in range(steps):
            one(i)
        dt = (time.perf_counter() - t0) / max(steps, 1)
    value = 1.0 / (STEPS_PER_IMAGE * dt)
    return dict(value=value, unit=UNIT, cores=cores, kind='port',
                sample=(f'{steps} x [1 of 50 DDIM steps at B=1: 2 fp32 UNet sample-forwards '
                        f'+ CFG + scheduler step] = {dt:.2f} s/step on {cores} threads; '
                        'images/s = 1/(50*t_step); VAE decode excluded')), dt


def _planted_pair(n_text: int, seed: int = 0, D: int = 768):
    '''Synthetic (prompts, guide) embeddings: 12 guide tokens planted on NON-adjacent text positions
    so arg-max, Threshold and Clustered (peaks / valleys) all fire without the adjacent-peak
    ZeroDivisionError of guidance.py:111-112.'''
    g = torch.Generator().manual_seed(seed)
    txt = torch.randn(n_text, 77, D, generator=g)
    img = torch.randn(1, 257, D, generator=g)
    for k in range(12):
        tok, row = 3 + 6 * k, 5 + 20 * k
        for b in range(n_text):
            txt[b, tok] = img[0, row] * (1.0 + 0.05 * k) + 0.3 * torch.randn(D, generator=g)
    return txt, img


def cpu_blend_baseline(n: int = 8):
    '''The reference blend (guidance.py Tweener.tween: Linear + Clustered + Threshold, reference
    defaults) through the oracle port on the host, blends/s.  The port is vectorised numpy / torch;
    the reference's own per-token Python loops ran 3-10 blends/s on 8 cores,
    so this baseline is the faster of the two CPU statements (conservative for the GPU ratio).'''
    from oracle import guidance_oracle as orc
    txt, img = _planted_pair(1)
    prm = orc.TweenParams()  # clustered 0.5, threshold (0.5, 0.5), linear (0, 0.5)
    orc.tween(txt, img, prm)
    t0 = time.perf_counter()
    for _ in range(n):
        orc.tween(txt, img, prm)
    return n / (time.perf_counter() - t0)


# dram__bytes_read.sum + dram__bytes_write.sum per launch, from `ncu --set full --clock-control none`
# captures of exactly these microbenchmark launches.  Each entry records the file under profiles/
# and the commit the kernel was captured at (NCU_CAPTURED_AT); output still resident in the 126 MB L2
# when a kernel ends is not counted by the DRAM counters.
NCU_CAPTURED_AT = 'see profiles/r02/SUMMARY.md'
NCU_TRAFFIC_BYTES = {'k3f': None, 'k3': None, 'k4': None, 'k2': None, 'k1': None}
_traffic_file = os.path.join(ROOT, 'profiles', 'r02', 'ncu_traffic.json')
if os.path.exists(_traffic_file):
    _t = json.load(open(_traffic_file))
    NCU_CAPTURED_AT = _t.get('captured_at', NCU_CAPTURED_AT)
    NCU_TRAFFIC_BYTES.update(_t.get('bytes_per_launch', {}))


# ------------------------------------------------------------------ kernel microbenches
def _time_cuda(fn, iters, flush=None):
    st = torch.cuda.current_stream()
    times = []
    for _ in range(iters):
        if flush is not None:
            flush.add_(1.0)  # > L2: evict everything between timed launches
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        fn()
        b.record(st)
        b.synchronize()
        times.append(a.elapsed_time(b) * 1e-3)
    return sum(times) / len(times)


def _graphed(fn, reps=1):
    '''Capture `reps` calls of fn in a CUDA graph (host launch latency is not part of a kernel time).'''
    fn()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    return g


def kernel_rooflines(dev, unet, peaks):
    '''Live CUDA-event timings of the hand-written kernels against their rooflines.'''
    from flexdiffuse_b200 import _native
    out = {}
    flush = torch.empty(96 * 1024 * 1024, dtype=torch.float32, device=dev)  # 384 MB
    ctx = torch.randn(2, 77, 768, device=dev)
    kv = unet.build_kv_cache(ctx)
    site_nq = {320: 4096, 640: 1024, 1280: 256}

    def attn2_sites(S):
        idx = torch.tensor([0] * (S // 2) + [1] * (S // 2), dtype=torch.int32, device=dev)
        sites = []
        for m in unet.cross_attentions():
            sites.append([m, torch.randn(S, site_nq[m.dim], m.dim, device=dev).bfloat16()])
        sites[6][1] = sites[6][1][:, :64].contiguous()  # mid block: 8x8 latent
        return idx, sites

    # ---- K3F: the fused attn2 layer (to_q + softmax(QK^T)V + to_out + bias), all 16 sites of one
    # UNet forward at 32 samples (configs[3]: batch 16 with CFG).  Tensor-bound (fused
    # variant): algorithmic FLOPs = 4 N C^2 (to_q, to_out) + 4 N 77 C (attention core) per sample.
    S = 32
    idx, sites = attn2_sites(S)
    outs = [torch.empty_like(x) for _, x in sites]
    scratch = [torch.empty_like(x) if m.dim != 320 else None for m, x in sites]
    flops = sum(S * (4 * x.shape[1] * m.dim * m.dim + 4 * x.shape[1] * 77 * m.dim) for m, x in sites)
    core_flops = sum(S * 4 * x.shape[1] * 77 * m.dim for m, x in sites)
    io_bytes = sum(2 * x.numel() * 2 + 2 * m.dim * m.dim * 2 + 2 * 80 * m.dim * 2 * 2 for m, x in sites)

    def run_fused(which=None):
        for (m, x), o, sc in zip(sites, outs, scratch):
            if which is None or m.dim == which:
                lin = m.to_out[0]
                _native.cross_attn_fused(x, m.to_q.weight, kv.kv, m.k_col_off, m.v_col_off, idx,
                                         lin.weight, lin.bias, m.heads, 77, 80, m.scale,
                                         attn=sc, out=o, want_attn=False)

    def run_unfused(which=None):
        for (m, x), o in zip(sites, outs):
            if which is None or m.dim == which:
                q = F.linear(x, m.to_q.weight)
                a = _native.cross_attn(q, kv.kv, m.k_col_off, m.v_col_off, idx, m.heads, 77, 80,
                                       m.scale)
                F.linear(a, m.to_out[0].weight, m.to_out[0].bias)

    t = _time_cuda(_graphed(run_fused).replay, 10, flush)
    t_unf = _time_cuda(_graphed(run_unfused).replay, 10, flush)
    per_c = {}
    for c in (320, 640, 1280):
        n_l = sum(1 for m, _ in sites if m.dim == c)
        tf = _time_cuda(_graphed(lambda c=c: run_fused(c)).replay, 5, flush)
        tu = _time_cuda(_graphed(lambda c=c: run_unfused(c)).replay, 5, flush)
        per_c[str(c)] = dict(launches=n_l, fused_us_per_site=tf / n_l * 1e6,
                             cublas_k3_cublas_us_per_site=tu / n_l * 1e6)
    assert _native.k3f_status() == [0, 0, 0, 0], 'K3F watchdog fired'
    out['k3f'] = dict(
        bound='tensor', achieved=flops / t / 1e12, peak=peaks['tensor'], unit='TFLOP/s',
        frac=flops / t / 1e12 / peaks['tensor'],
        frac_of_sustained=(flops / t / 1e12 / peaks['tensor_sustained']
                           if peaks.get('tensor_sustained') else None),
        traffic=NCU_TRAFFIC_BYTES['k3f'], traffic_captured_at=NCU_CAPTURED_AT,
        algorithmic_flops_per_launch=flops / 16, algorithmic_bytes_per_launch=io_bytes / 16,
        kernel=f'k3f_kernel<40|80|160> (fd_cross_attn_fused: 16 attn2 sites, {S} samples, one launch '
               'per site replacing cuBLAS to_q + K3 + cuBLAS to_out)',
        launches=16, avg_launch_us=t / 16 * 1e6, samples=S,
        attention_core_tflops=core_flops / t / 1e12,
        same_work_cublas_k3_cublas_us=t_unf * 1e6, fused_us=t * 1e6, per_width=per_c,
        peak_of=peaks['source'] + ' (burst bf16: the kernel is timed alone)')
    del sites, outs, scratch

    # ---- K3 alone (round 1's roofline kernel; still the attention of the non-fused dispatch): HBM
    S = 8
    idx, sites = attn2_sites(S)
    outs = [torch.empty_like(q) for _, q in sites]
    alg_bytes = sum(2 * q.numel() * 2 + 2 * 80 * m.dim * 2 * 2 for m, q in sites)
    alg_flops = sum(4 * q.shape[0] * q.shape[1] * 77 * m.dim for m, q in sites)

    def run_k3():
        for (m, q), o in zip(sites, outs):
            _native.cross_attn(q, kv.kv, m.k_col_off, m.v_col_off, idx, m.heads, 77, 80,
                               m.scale, out=o)

    t = _time_cuda(_graphed(run_k3).replay, 10, flush)
    out['k3'] = dict(bound='hbm', achieved=alg_bytes / t / 1e9, peak=peaks['hbm'],
                     unit='GB/s', frac=alg_bytes / t / 1e9 / peaks['hbm'],
                     traffic=NCU_TRAFFIC_BYTES['k3'], traffic_captured_at=NCU_CAPTURED_AT,
                     algorithmic_bytes_per_launch=alg_bytes / 16,
                     kernel='k3_cross_attn_kernel (16 attn2 sites, 8 samples)',
                     launches=16, avg_launch_us=t / 16 * 1e6,
                     tflops=alg_flops / t / 1e12, peak_of=peaks['source'])
    del sites, outs
    # ---- K4: fused CFG + DDIM step, 1024 samples fp32 (268 MB algorithmic traffic > L2); the launch
    # is graph-replayed like the others, so no host launch gap sits between the two events
    B = 1024
    n = B * 4 * 64 * 64
    u, c, x = (torch.randn(n, device=dev) for _ in range(3))
    xo = torch.empty_like(x)
    k = _native.SchedCoeffs()
    k.guidance, k.use_cfg, k.a, k.b = 7.5, 1, 1.01, -0.05
    k.w[0] = 1.0
    t = _time_cuda(_graphed(lambda: _native.cfg_sched_step(u, c, x, k, xo)).replay, 20)
    # ... and at the batch sizes the loop really runs: launch-bound
    real_b = {}
    for rb in (1, 8, 16):
        m = rb * 4 * 64 * 64
        g4 = _graphed(lambda: _native.cfg_sched_step(u[:m], c[:m], x[:m], k, xo[:m]), reps=20)
        real_b[str(rb)] = round(_time_cuda(g4.replay, 5) / 20 * 1e6, 2)
    out['k4'] = dict(bound='hbm', achieved=16 * n / t / 1e9, peak=peaks['hbm'], unit='GB/s',
                     us_per_step_at_batch=real_b,
                     frac=16 * n / t / 1e9 / peaks['hbm'], traffic=NCU_TRAFFIC_BYTES['k4'],
                     traffic_captured_at=NCU_CAPTURED_AT,
                     algorithmic_bytes_per_launch=16 * n,
                     kernel='k4_cfg_sched_kernel (DDIM, fp32, 1024 samples; graph-replayed, inputs 268 MB > L2)',
                     avg_launch_us=t * 1e6, peak_of=peaks['source'])
    del u, c, x, xo
    # ---- K2: K/V projection of 9 contexts (config 3), tensor bound
    ctx9 = torch.zeros(9 * 80, 768, device=dev, dtype=torch.bfloat16)
    ctx9.view(9, 80, 768)[:, :77] = torch.randn(9, 77, 768, device=dev).bfloat16()
    kv9 = torch.empty(9 * 80, unet._kv_weight.shape[0], device=dev, dtype=torch.bfloat16)
    t = _time_cuda(_graphed(lambda: _native.kv_project(ctx9, unet._kv_weight, out=kv9)).replay, 10, flush)
    fl = 2 * 9 * 80 * 768 * 24960
    # context only (never on the product path): the same GEMM through torch.matmul (cuBLAS)
    t_cublas = _time_cuda(_graphed(lambda: torch.matmul(ctx9, unet._kv_weight.t(), out=kv9)).replay, 10, flush)
    out['k2'] = dict(bound='tensor', achieved=fl / t / 1e12, peak=peaks['tensor'],
                     unit='TFLOP/s', frac=fl / t / 1e12 / peaks['tensor'], traffic=NCU_TRAFFIC_BYTES['k2'],
                     traffic_captured_at=NCU_CAPTURED_AT,
                     kernel='k2_gemm_kernel (M=720,N=24960,K=768; 9 contexts)',
                     avg_launch_us=t * 1e6, cublas_same_shape_us=t_cublas * 1e6, peak_of=peaks['source'])
    # ---- K13: the GEGLU projections of all 16 feed-forward blocks of one UNet forward at 32 samples (batch 16, CFG), tensor
    # bound: one launch per site replacing cuBLAS nn.Linear(C, 8C) + K6; algorithmic flops = 2 M C 8C
    sites13 = []
    for blk in (m for m in unet.modules() if m.__class__.__name__ == 'GEGLU'):
        C = blk.proj.weight.shape[1]
        sites13.append((torch.randn(32 * site_nq[C], C, device=dev).bfloat16(), blk.proj.weight, blk.proj.bias))
    sites13[6] = (sites13[6][0][:32 * 64].contiguous(),) + sites13[6][1:]  # mid block: 8x8 latent
    fl13 = sum(2.0 * x.shape[0] * x.shape[1] * w.shape[0] for x, w, _ in sites13)
    by13 = sum(2.0 * (x.numel() + w.numel() + x.shape[0] * w.shape[0] // 2) for x, w, _ in sites13)
    import torch.nn.functional as F13
    t13 = _time_cuda(_graphed(lambda: [_native.ff_geglu(x, w, b) for x, w, b in sites13]).replay, 5)
    t13_old = _time_cuda(_graphed(lambda: [_native.geglu(F13.linear(x, w, b)) for x, w, b in sites13]).replay, 5)
    t13_gemm = _time_cuda(_graphed(lambda: [F13.linear(x, w, b) for x, w, b in sites13]).replay, 5)
    out['k13'] = dict(bound='tensor', achieved=fl13 / t13 / 1e12, peak=peaks['tensor'], unit='TFLOP/s',
                      frac=fl13 / t13 / 1e12 / peaks['tensor'],
                      frac_of_sustained=(fl13 / t13 / 1e12 / peaks['tensor_sustained']
                                         if peaks.get('tensor_sustained') else None),
                      traffic=NCU_TRAFFIC_BYTES.get('k13'), traffic_captured_at=NCU_CAPTURED_AT,
                      algorithmic_flops_per_launch=fl13 / len(sites13), algorithmic_bytes_per_launch=by13 / len(sites13),
                      kernel='k13_ff_geglu_kernel (fd_ff_geglu: the 16 GEGLU projections of one UNet forward, 32 samples; '
                             'cta_group::2 tcgen05 GEMM with value * gelu(gate) in the epilogue)',
                      launches=len(sites13), avg_launch_us=t13 / len(sites13) * 1e6, fused_us=t13 * 1e6,
                      same_work_cublas_k6_us=t13_old * 1e6, cublas_gemm_alone_us=t13_gemm * 1e6,
                      peak_of=peaks['source'])
    del sites13
    # ---- K1: blends/s -- 1024 prompts x 1 shared guide image, the reference's default parameters
    # (Linear (0, 0.5) + Clustered 0.5 + Threshold (0.5, 0.5): BASELINE.json configs[0]) on planted
    # inputs that do not hit the adjacent-peak ZeroDivisionError
    nb = 1024
    txt_h, img_h = _planted_pair(nb)
    txt, img = txt_h.to(dev), img_h.to(dev)
    prm = _native.TweenParams()
    prm.threshold_floor = prm.threshold_mult = prm.max_guidance = prm.clustered = 0.5
    prm.header_max, prm.align_mode, prm.mapping_reuse = 0.15, 1, 1
    lin = torch.linspace(0.0, 0.5, 77)[None].to(dev)
    res = _native.sim_blend(txt, img, [prm], lin)
    n_flagged = int((res['status'].cpu() != 0).sum())  # ZeroDivision / range outcomes (expected: 0)
    n_guided = int((res['weights'][0, 0] != lin[0]).sum())
    f1 = lambda: _native.sim_blend(txt, img, [prm], lin)
    f1()
    t = _time_cuda(f1, 5)
    fl = 3 * 2 * 128 * 256 * 768 * nb   # three fp16 products on 128 x 256 tiles
    by = nb * (2 * 77 * 768 * 4) + 257 * 768 * 4
    out['k1'] = dict(bound='hbm', achieved=by / t / 1e9, peak=peaks['hbm'], unit='GB/s',
                     frac=by / t / 1e9 / peaks['hbm'], traffic=NCU_TRAFFIC_BYTES['k1'],
                     traffic_captured_at=NCU_CAPTURED_AT,
                     algorithmic_bytes_per_launch=by,
                     kernel='k1_prep_guide_kernel + k1_sim_blend_kernel (1024 prompts x 1 guide, Linear + '
                            'Clustered 0.5 + Threshold; two-term fp16 split GEMM)',
                     blends_per_s=nb / t, issued_fp16_tflops=fl / t / 1e12,
                     tokens_reweighted_by_cluster_or_threshold=n_guided, prompts_flagged=n_flagged,
                     avg_launch_us=t * 1e6, peak_of=peaks['source'])
    # the same launch at 4096 prompts: 1024 prompts are 6.9 per persistent CTA, i.e. three two-prompt batches and a single
    # (four batch slots for 3.46 batches of work); 4096 amortise that quantisation
    txt4 = txt.repeat(4, 1, 1)
    f4 = lambda: _native.sim_blend(txt4, img, [prm], lin)
    f4()
    t4 = _time_cuda(f4, 3)
    by4 = 4 * nb * (2 * 77 * 768 * 4) + 257 * 768 * 4
    out['k1']['at_4096_prompts'] = dict(avg_launch_us=t4 * 1e6, blends_per_s=4 * nb / t4, achieved=by4 / t4 / 1e9,
                                        frac=by4 / t4 / 1e9 / peaks['hbm'])
    del txt, img, txt4, f4
    # ---- K5 (GroupNorm + bias + SiLU glue): by time the largest hand-written kernel of a step, so
    # it is reported too: every (N=2, C, H, W) call of one UNet forward, graph-replayed on its own
    # L2-resident input, as in situ where the producer convolution has just written it
    calls = collections.Counter()
    orig = _native.groupnorm_act

    def spy(x, w, b, groups, eps, silu, bias=None):
        calls[(tuple(x.shape), bool(silu), bias is not None)] += 1
        return orig(x, w, b, groups, eps, silu, bias)

    _native.groupnorm_act = spy
    try:
        with torch.no_grad():
            unet(torch.randn(2, 4, 64, 64, device=dev, dtype=torch.bfloat16), 481,
                 encoder_hidden_states=torch.randn(2, 77, 768, device=dev, dtype=torch.bfloat16))
    finally:
        _native.groupnorm_act = orig
    t5, b5 = 0.0, 0
    for (shape, silu, has_bias), cnt in calls.items():
        N, C, H, W = shape
        xx = torch.randn(N, C, H, W, device=dev).bfloat16().contiguous(memory_format=torch.channels_last)
        ww, bb = torch.ones(C, device=dev).bfloat16(), torch.zeros(C, device=dev).bfloat16()
        tb = torch.randn(N, C, device=dev).bfloat16() if has_bias else None
        g5 = _graphed(lambda: orig(xx, ww, bb, 32, 1e-5, silu, tb), reps=10)
        t5 += cnt * _time_cuda(g5.replay, 3) / 10
        b5 += cnt * 2 * xx.numel() * 2
    out['k5'] = dict(bound='hbm', achieved=b5 / t5 / 1e9, peak=peaks['hbm'], unit='GB/s',
                     frac=b5 / t5 / 1e9 / peaks['hbm'], traffic=None,
                     algorithmic_bytes_per_launch=b5 / sum(calls.values()),
                     kernel='k5_gn_cluster / k5_gn_vec kernels (all %d GroupNorm calls of one B=1 CFG UNet forward; '
                            'inputs L2-resident as in situ, so HBM is the wrong ceiling: fixed latency binds)'
                            % sum(calls.values()),
                     launches=sum(calls.values()), avg_launch_us=t5 / sum(calls.values()) * 1e6,
                     us_per_unet_forward=t5 * 1e6, peak_of=peaks['source'])
    return out


def guide_embeds_e2e(dev, n=10, tf32=False, clip=None):
    '''`Guide.embeds(prompt, image)` end to end on the GPU (BASELINE.json configs[0] on the B200):
    random-init CLIP ViT-L/14 towers (CUDA-graphed PyTorch) + K1, one prompt x one 512x512 guide
    image, the reference's default Linear + Clustered + Threshold parameters.  Returns calls per
    second (the survey measured 0.67 s per call for the reference on 8 CPU cores).'''
    import numpy as np
    from PIL import Image
    from flexdiffuse_b200 import factory
    from flexdiffuse_b200.guidance import Guide
    clip = clip if clip is not None else factory.build_clip(dev)
    guide = Guide(clip, factory.FakeTokenizer(), device=str(dev), tf32_towers=tf32)
    img = Image.fromarray((np.random.RandomState(0).rand(512, 512, 3) * 255).astype('uint8'))
    prompt = 'a photograph of an astronaut riding a horse'
    kw = {}
    with torch.no_grad():
        try:
            guide.embeds(prompt, img)
        except ZeroDivisionError:  # adjacent similarity peaks (guidance.py:111-112): drop the Clustered term
            kw = dict(guide_clustered=0.0)
        for _ in range(3):
            out = guide.embeds(prompt, img, **kw)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n):
            out = guide.embeds(prompt, img, **kw)
        torch.cuda.synchronize()
    assert tuple(out.shape) == (1, 77, 768)
    rate = n / (time.perf_counter() - t0)
    with torch.no_grad():
        gi = guide.encoder.image(img)
    return rate, gi


def _guide_rates(dev):
    from flexdiffuse_b200 import factory
    clip = factory.build_clip(dev)
    r32, g32 = guide_embeds_e2e(dev, clip=clip)
    rtf, gtf = guide_embeds_e2e(dev, tf32=True, clip=clip)
    return {'gpu_guide_embeds_calls_per_s': r32,
            'gpu_guide_embeds_calls_per_s_tf32_towers': rtf,
            'tf32_towers_image_embedding_rel_l2_vs_default': ((gtf - g32).norm() / g32.norm()).item(),
            'guide_embeds_note': 'default towers: every Linear on K11 (fp32-accurate three-product fp16 split on '
                                 'tcgen05, 2e-5 of transformers\' fp32 output); tf32_towers=True is the cuBLAS TF32 '
                                 'comparison point'}


# ------------------------------------------------------------------ plain-torch GPU baseline
def _tb_resnet(m, x, temb):
    gn = lambda n, t: F.group_norm(t, n.num_groups, n.weight, n.bias, n.eps)
    h = F.conv2d(F.silu(gn(m.norm1, x)), m.conv1.weight, m.conv1.bias, padding=1)
    h = h + F.linear(temb, m.time_emb_proj.weight, m.time_emb_proj.bias)[:, :, None, None]
    h = F.conv2d(F.silu(gn(m.norm2, h)), m.conv2.weight, m.conv2.bias, padding=1)
    if m.conv_shortcut is not None:
        x = F.conv2d(x, m.conv_shortcut.weight, m.conv_shortcut.bias)
    return x + h


def _tb_attn(a, x, ctx):
    B, N, C = x.shape
    sp = lambda t: t.view(B, -1, a.heads, C // a.heads).transpose(1, 2)
    o = F.scaled_dot_product_attention(sp(F.linear(x, a.to_q.weight)), sp(F.linear(ctx, a.to_k.weight)),
                                       sp(F.linear(ctx, a.to_v.weight)))
    return F.linear(o.transpose(1, 2).reshape(B, N, C), a.to_out[0].weight, a.to_out[0].bias)


def _tb_transformer(m, x, ctx):
    B, C, H, W = x.shape
    h = F.group_norm(x, m.norm.num_groups, m.norm.weight, m.norm.bias, m.norm.eps)
    h = F.conv2d(h, m.proj_in.weight, m.proj_in.bias).permute(0, 2, 3, 1).reshape(B, H * W, C)
    b = m.transformer_blocks[0]
    ln = lambda n, t: F.layer_norm(t, (C,), n.weight, n.bias, n.eps)
    n1 = ln(b.norm1, h)
    h = _tb_attn(b.attn1, n1, n1) + h
    h = _tb_attn(b.attn2, ln(b.norm2, h), ctx) + h
    a_, gate = F.linear(ln(b.norm3, h), b.ff.net[0].proj.weight, b.ff.net[0].proj.bias).chunk(2, dim=-1)
    h = F.linear(a_ * F.gelu(gate), b.ff.net[2].weight, b.ff.net[2].bias) + h
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return F.conv2d(h, m.proj_out.weight, m.proj_out.bias) + x


def torch_unet_forward(unet, x, temb_sin, ctx):
    '''The reference's UNet call (pipeline/guide.py:56-58 -> diffusers UNet2DConditionModel) written
    with plain torch ops on the product module's weights: cuDNN convolutions, cuBLAS linears, ATen
    GroupNorm / LayerNorm / GELU, SDPA attention, to_k / to_v recomputed every call as diffusers does.
    No flexdiffuse_b200 kernel is involved: this is the GPU "kernel to beat" of BASELINE.md section 5.'''
    te = unet.time_embedding
    temb = F.silu(te['linear_2'](F.silu(te['linear_1'](temb_sin.to(x.dtype).expand(x.shape[0], -1)))))
    h = unet.conv_in(x)
    skips = [h]
    for blk in unet.down_blocks:
        for i, res in enumerate(blk.resnets):
            h = _tb_resnet(res, h, temb)
            if blk.attentions is not None:
                h = _tb_transformer(blk.attentions[i], h, ctx)
            skips.append(h)
        if blk.downsamplers is not None:
            h = blk.downsamplers[0](h)
            skips.append(h)
    h = _tb_resnet(unet.mid_block.resnets[0], h, temb)
    h = _tb_transformer(unet.mid_block.attentions[0], h, ctx)
    h = _tb_resnet(unet.mid_block.resnets[1], h, temb)
    for blk in unet.up_blocks:
        for i, res in enumerate(blk.resnets):
            h = _tb_resnet(res, torch.cat([h, skips.pop()], dim=1), temb)
            if blk.attentions is not None:
                h = _tb_transformer(blk.attentions[i], h, ctx)
        if blk.upsamplers is not None:
            h = blk.upsamplers[0](h)
    n = unet.conv_norm_out
    return unet.conv_out(F.silu(F.group_norm(h, n.num_groups, n.weight, n.bias, n.eps)))


def gpu_torch_baseline(dev, unet, vae, uncond, embeds, images=2):
    '''configs[1] through plain torch: 50 DDIM steps of [cat -> UNet -> chunk -> CFG -> scheduler
    update] exactly as pipeline/guide.py:46-64 + flex.py:262-287 do it, eager and with the UNet
    forward captured in a CUDA graph.  Returns images/s for both (VAE decode through the same module
    as the product: it is 2-3 % of an image).'''
    from flexdiffuse_b200 import schedulers
    from flexdiffuse_b200.unet import timestep_embedding
    sched = schedulers.DDIMScheduler()
    sched.set_timesteps(STEPS_PER_IMAGE)
    ts = [int(t) for t in sched.timesteps]
    coef = [sched.coefficients(t) for t in ts]
    tembs = [timestep_embedding(torch.tensor([float(t)]), 320).to(dev) for t in ts]
    ctx = torch.cat([uncond, embeds]).to(unet.conv_in.weight.dtype)
    out = {}
    with torch.no_grad():
        static_x = torch.zeros(2, 4, HW // 8, HW // 8, device=dev, dtype=ctx.dtype).contiguous(
            memory_format=torch.channels_last)
        static_t = torch.zeros(1, 320, device=dev)

        def fwd():
            return torch_unet_forward(unet, static_x, static_t, ctx)

        for _ in range(2):
            fwd()
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            fwd()
        torch.cuda.current_stream().wait_stream(side)
        with torch.cuda.graph(graph):
            static_eps = fwd()

        def image(graphed):
            lat = torch.randn(1, 4, HW // 8, HW // 8, device=dev)
            for i in range(STEPS_PER_IMAGE):
                static_x.copy_(torch.cat([lat, lat]))
                static_t.copy_(tembs[i])
                eps = static_eps if graphed else fwd()
                if graphed:
                    graph.replay()
                u, c = eps.float().chunk(2)
                e = u + GUIDANCE * (c - u)
                lat = coef[i][0] * lat + coef[i][1] * e
            return vae.decode(lat / 0.18215).sample

        for name, graphed in (('eager', False), ('cuda_graph', True)):
            image(graphed)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(images):
                image(graphed)
            b.record()
            b.synchronize()
            out[name + '_images_per_s'] = images / (a.elapsed_time(b) * 1e-3)
    out['what'] = ('torch 2.11 bf16, cuDNN convolutions / cuBLAS linears / ATen norms / SDPA attention, K and V '
                   're-projected every step, CFG + DDIM update as separate elementwise ops; same weights, B=1')
    return out


# ------------------------------------------------------------------ the other BASELINE configs
class _Enc:
    '''Stand-in for CLIPEncoder on the synthetic-embedding workloads: the CLIP towers stay in
    PyTorch and are not part of configs[1..4]; prompt embeddings are synthetic.'''
    def __init__(self, uncond):
        self.uncond = uncond

    def prompt(self, p):
        return self.uncond


def _timed_steps(fn, steps, warmup, barrier=None):
    for _ in range(warmup):
        fn()
    (barrier or torch.cuda.synchronize)()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    (barrier or torch.cuda.synchronize)()
    return a.elapsed_time(b) / steps


def run_config2(dev, pipe_factory, unet, steps=1, warmup=1):
    '''configs[2]: K1 blends 8 prompts with one guide image in each of the three guide orders, K2
    caches K/V of the 8 + 1 contexts, 50 DDIM steps (K3F / K3, K4) at batch 8, decode -> uint8.
    One step = the three modes = 24 images.  Inputs start on the host (end to end).'''
    from flexdiffuse_b200 import guidance as G
    from flexdiffuse_b200 import schedulers
    from flexdiffuse_b200.pipeline.guide import SimpleGuide
    B = 8
    txt_h, img_h = _planted_pair(B, seed=7)
    txt_h, img_h = txt_h.pin_memory(), img_h.pin_memory()
    unc_h = torch.randn(1, 77, 768, generator=torch.Generator().manual_seed(8)).pin_memory()
    pipe = pipe_factory(schedulers.DDIMScheduler())
    gen = torch.Generator()

    def step():
        txt, img = txt_h.to(dev, non_blocking=True), img_h.to(dev, non_blocking=True)
        unc = unc_h.to(dev, non_blocking=True)
        n = 0
        for mode in (G.GUIDE_ORDER_TEXT, G.GUIDE_ORDER_ALIGN, G.GUIDE_ORDER_DIRECT):
            tw = G.Tweener((0.3, 0.5), (0.1, 0.5), 0.0, 0.35, 0.15, mode, True)
            ctx, _ = tw.tween_batch(txt, img, check=False)
            guide = SimpleGuide(_Enc(unc), unet, GUIDANCE, STEPS_PER_IMAGE, ctx, use_cuda_graph=True)
            gen.manual_seed(100 + mode)
            imgs = pipe(guide, init_size=(HW, HW), generator=gen, output_type='uint8', return_dict=False)[0]
            n += imgs.shape[0]
        return n

    ms = _timed_steps(step, steps, warmup)
    return dict(workload=CONFIG_NAMES[2], images_per_step=3 * B, ms_per_step=ms,
                images_per_s=3 * B / (ms * 1e-3), steps=steps, warmup=warmup, e2e=True,
                h2d_bytes_per_step=int((txt_h.numel() + img_h.numel() + unc_h.numel() + 3 * B * 4 * 64 * 64) * 4),
                d2h_bytes_per_step=3 * B * HW * HW * 3)


def run_config3(dev, pipe_factory, unet, steps=1, warmup=1):
    '''configs[3]: img2img strength 0.6 (t_start = 20 -> 31 PNDM evaluations), image-guided context,
    PNDM 50 steps, batch 16, bf16; init image and embeddings start on the host, uint8 images come back.'''
    from flexdiffuse_b200 import guidance as G
    from flexdiffuse_b200 import schedulers
    from flexdiffuse_b200.pipeline.guide import SimpleGuide
    B = 16
    txt_h, img_h = _planted_pair(B, seed=9)
    txt_h, img_h = txt_h.pin_memory(), img_h.pin_memory()
    unc_h = torch.randn(1, 77, 768, generator=torch.Generator().manual_seed(10)).pin_memory()
    init_h = (torch.rand(1, 3, HW, HW, generator=torch.Generator().manual_seed(11)) * 2 - 1).pin_memory()
    pipe = pipe_factory(schedulers.PNDMScheduler())
    gen = torch.Generator(device=dev)

    def step():
        txt, img = txt_h.to(dev, non_blocking=True), img_h.to(dev, non_blocking=True)
        unc = unc_h.to(dev, non_blocking=True)
        init = init_h.to(dev, non_blocking=True).to(unet.conv_in.weight.dtype)
        ctx, _ = G.Tweener((0.3, 0.5), (0.1, 0.5), 0.0, 0.35).tween_batch(txt, img, check=False)
        guide = SimpleGuide(_Enc(unc), unet, GUIDANCE, STEPS_PER_IMAGE, ctx, use_cuda_graph=True)
        gen.manual_seed(12)
        return pipe(guide, init_image=init, strength=0.6, generator=gen, output_type='uint8',
                    return_dict=False)[0]

    ms = _timed_steps(step, steps, warmup)
    return dict(workload=CONFIG_NAMES[3], images_per_step=B, ms_per_step=ms,
                images_per_s=B / (ms * 1e-3), unet_evaluations_per_image=31, steps=steps, warmup=warmup,
                e2e=True,
                h2d_bytes_per_step=int((txt_h.numel() + img_h.numel() + unc_h.numel() + init_h.numel()) * 4),
                d2h_bytes_per_step=B * HW * HW * 3)


def run_config4(dev, pipe_factory, unet, n_samples, world, barrier, steps=1, warmup=0):
    '''configs[4]: STRONG scaling -- a fixed grid of `n_samples` = 8 prompts x 8 Tweener parameter
    sets x seeds, contexts from ONE K1 launch (n_params = 8), sharded contiguously over the ranks,
    micro-batch 16, 50-step DDIM, then one all-gather of the [n_samples,4,64,64] fp32 latents.'''
    import torch.distributed as dist
    from flexdiffuse_b200 import _native, schedulers, sweep
    n_prompts, n_params = 8, 8
    txt_h, img_h = _planted_pair(n_prompts, seed=13)
    txt, img = txt_h.to(dev), img_h.to(dev)
    prms, lins = [], []
    for p in range(n_params):
        q = _native.TweenParams()
        q.threshold_floor, q.threshold_mult = 0.3, 0.25 + 0.05 * p
        q.clustered, q.max_guidance, q.header_max = 0.0, 0.3 + 0.05 * p, 0.15
        q.align_mode, q.mapping_reuse = 1, 1
        prms.append(q)
        lins.append(torch.linspace(0.0, 0.2 + 0.05 * p, 77))
    res = _native.sim_blend(txt, img, prms, torch.stack(lins).to(dev))
    ctx64 = res['out'].reshape(n_prompts * n_params, 77, 768)
    reps = (n_samples + ctx64.shape[0] - 1) // ctx64.shape[0]
    contexts = ctx64.repeat(reps, 1, 1)[:n_samples].contiguous()
    seeds = [5000 + i for i in range(n_samples)]
    uncond = torch.randn(1, 77, 768, device=dev, generator=torch.Generator(device=dev).manual_seed(14))
    pipe = pipe_factory(schedulers.DDIMScheduler())
    den = sweep.make_denoiser(pipe, _Enc(uncond), unet, contexts, seeds, GUIDANCE, STEPS_PER_IMAGE,
                              init_size=(HW, HW), use_cuda_graph=True)
    gather_ms = []

    def step():
        local = sweep.run_sweep(n_samples, den, micro_batch=16, gather=False)
        if world > 1:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            width = (n_samples + world - 1) // world
            pad = torch.zeros((width,) + tuple(local.shape[1:]), device=dev)
            pad[:local.shape[0]] = local
            buf = torch.empty((world * width,) + tuple(local.shape[1:]), device=dev)
            a.record()
            dist.all_gather_into_tensor(buf, pad)
            b.record()
            b.synchronize()
            gather_ms.append(a.elapsed_time(b))
        return local

    ms = _timed_steps(step, steps, warmup, barrier)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return dict(workload=CONFIG_NAMES[4], scaling='strong', n_samples=n_samples, n_gpus=world,
                micro_batch=16, ms_per_sweep=ms, samples_per_s=n_samples / (ms * 1e-3),
                all_gather_bytes=n_samples * 4 * 64 * 64 * 4,
                all_gather_ms=(sum(gather_ms) / len(gather_ms)) if gather_ms else 0.0,
                steps=steps, warmup=warmup)


# ------------------------------------------------------------------ main arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=4)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--config', type=int, default=1, choices=[1, 2, 3, 4],
                    help='BASELINE.json configs[N] as the headline line (default 1)')
    ap.add_argument('--no-kernels', action='store_true', help='skip kernel microbenches / baselines / other configs')
    ap.add_argument('--batch', type=int, default=1,
                    help='images per step per GPU (default 1 = BASELINE.json configs[1])')
    ap.add_argument('--kernels-only', action='store_true',
                    help='development aid: only the kernel roofline section')
    ap.add_argument('--sweep-samples', type=int, default=0,
                    help='configs[4] grid size (default: 1024 with --config 4, a bounded 128 inside the default run)')
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))

    if args.impl == 'reference':
        if rank != 0:
            return
        base, dt = cpu_reference_arm(max(args.steps, 1), min(args.warmup, 1))
        print(json.dumps({
            'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': UNIT,
            'n_gpus': args.gpus, 'steps': args.steps, 'warmup': min(args.warmup, 1),
            'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': WORKLOAD,
                       'note': 'one bench step = a bounded sample (1 of 50 DDIM steps)'},
            'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': UNIT, 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0}}))
        return

    from flexdiffuse_b200 import _native, factory, schedulers
    from flexdiffuse_b200.pipeline.flex import FlexPipeline
    from flexdiffuse_b200.pipeline.guide import SimpleGuide
    _native.lib()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    _native.require_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    torch.backends.cudnn.benchmark = True
    peaks = _peaks()

    unet = factory.build_unet(dev, torch.bfloat16, seed=0)
    if args.kernels_only:
        print(json.dumps(kernel_rooflines(dev, unet, peaks)))
        return
    vae = factory.build_vae(dev, torch.bfloat16, seed=1)
    pipe_factory = lambda sched: FlexPipeline(vae, None, None, unet, sched)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    common = {'higher_is_better': True, 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
              'n_gpus': world}

    # ------------------------------------------------ headline = one of configs[2..4]
    if args.config in (2, 3):
        fn = run_config2 if args.config == 2 else run_config3
        before = _native.LAUNCHES
        with ClockSampler(local) as cs:
            r = fn(dev, pipe_factory, unet, steps=max(args.steps, 1), warmup=max(args.warmup, 1))
        if rank == 0:
            print(json.dumps({'metric': METRIC.replace('ddim_cfg7.5', 'guided' if args.config == 2 else 'pndm_img2img'),
                              'value': world * r['images_per_s'], 'unit': UNIT, 'steps': r['steps'],
                              'warmup': r['warmup'], 'ms_per_step': r['ms_per_step'], 'scaling': 'weak',
                              'config': {'workload': r['workload']},
                              'e2e': {'value': world * r['images_per_s'], 'unit': UNIT,
                                      'h2d_bytes_per_step': r['h2d_bytes_per_step'],
                                      'd2h_bytes_per_step': r['d2h_bytes_per_step']},
                              'gpu_launches': _native.LAUNCHES - before, 'clocks': cs.summary(), **common}))
        return
    if args.config == 4:
        n = args.sweep_samples or 1024
        before = _native.LAUNCHES
        with ClockSampler(local) as cs:
            r = run_config4(dev, pipe_factory, unet, n, world, barrier, steps=max(min(args.steps, 2), 1), warmup=0)
        if rank == 0:
            print(json.dumps({'metric': 'sd15_512x512_50step_sweep_samples_per_s', 'value': r['samples_per_s'],
                              'unit': 'samples/s', 'steps': r['steps'], 'warmup': 0,
                              'ms_per_step': r['ms_per_sweep'], 'scaling': 'strong',
                              'config': {'workload': r['workload'], 'n_samples': n, 'micro_batch': 16},
                              'all_gather': {'bytes': r['all_gather_bytes'], 'ms': r['all_gather_ms']},
                              'gpu_launches': _native.LAUNCHES - before, 'clocks': cs.summary(), **common}))
        return

    # ------------------------------------------------ headline = configs[1]
    B = args.batch
    g = torch.Generator().manual_seed(1234 + rank)
    h_uncond = torch.randn(1, 77, 768, generator=g).pin_memory()
    h_embeds = torch.randn(B, 77, 768, generator=g).pin_memory()
    d_uncond, d_embeds = h_uncond.to(dev), h_embeds.to(dev)
    noise_elems = B * 4 * (HW // 8) * (HW // 8)
    pipe = pipe_factory(schedulers.DDIMScheduler())
    d_gen = torch.Generator(device=dev)
    h_gen = torch.Generator()
    gathered = (torch.empty((world * B, 4, HW // 8, HW // 8), device=dev)
                if world > 1 else None)

    def collect(lat):
        # the only collective of the path: gather the output latents of the sharded samples
        if world > 1:
            dist.all_gather_into_tensor(gathered, lat.contiguous())

    def run_resident():
        # every input already in HBM, result left in HBM
        d_gen.manual_seed(99 + rank)
        guide = SimpleGuide(_Enc(d_uncond), unet, GUIDANCE, STEPS_PER_IMAGE, d_embeds,
                            use_cuda_graph=True)
        lat = pipe(guide, init_size=(HW, HW), generator=d_gen, output_type='latent',
                   return_dict=False)
        collect(lat)
        return pipe.decode(lat, 'pt')

    def run_e2e():
        # the call a user makes, from HOST buffers to a HOST uint8 image (what PIL wraps)
        h_gen.manual_seed(99 + rank)
        uncond = h_uncond.to(dev, non_blocking=True)
        embeds = h_embeds.to(dev, non_blocking=True)
        guide = SimpleGuide(_Enc(uncond), unet, GUIDANCE, STEPS_PER_IMAGE, embeds,
                            use_cuda_graph=True)
        if world == 1:
            return pipe(guide, init_size=(HW, HW), generator=h_gen, output_type='uint8',
                        return_dict=False)[0]
        lat = pipe(guide, init_size=(HW, HW), generator=h_gen, output_type='latent',
                   return_dict=False)
        collect(lat)
        return pipe.decode(lat, 'uint8')

    def timed(fn, steps, warmup, sample_clocks=False):
        for _ in range(warmup):
            fn()
        barrier()
        before = _native.LAUNCHES
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sampler = ClockSampler(local) if sample_clocks else None
        if sampler:
            sampler.__enter__()
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        if sampler:
            sampler.__exit__()
        ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms / steps, _native.LAUNCHES - before, (sampler.summary() if sampler else None)

    W = max(args.warmup, 3)
    ms_res, launches, clocks = timed(run_resident, args.steps, W, sample_clocks=True)
    ms_e2e, _, _ = timed(run_e2e, args.steps, 1)

    value = world * B / (ms_res * 1e-3)
    e2e = world * B / (ms_e2e * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'steps': args.steps,
        'warmup': W, 'ms_per_step': ms_res, 'scaling': 'weak',
        'config': {'workload': WORKLOAD if B == 1 else WORKLOAD.replace('batch 1 per GPU', f'batch {B} per GPU'),
                   'images_per_step_per_gpu': B,
                   'parallelism': f'replicas x{world}: independent samples sharded over ranks, no '
                                  'collective in the loop, one NCCL all-gather of the output '
                                  'latents per step',
                   'l2': 'working set (1.7 GB bf16 UNet weights streamed every denoising '
                         'step) exceeds the 126 MB L2; no explicit flush',
                   'cuda_graph': 'UNet forward captured once, replayed 50x per image',
                   'attn2_dispatch': 'auto: K3F at the C=320 sites (and C=640 for 4..16 samples), cuBLAS + K3 + '
                                     'cuBLAS elsewhere (profiles/r02/SUMMARY.md)'},
        'e2e': {'value': e2e, 'unit': UNIT, 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': int(h_uncond.numel() + h_embeds.numel() +
                                          noise_elems) * 4,
                'd2h_bytes_per_step': B * 3 * HW * HW},
        'gpu_launches': launches,
        'clocks': clocks,
        **common,
    }
    extras = not args.no_kernels and B == 1
    # configs[4]: strong scaling of a fixed sample grid, measured at every N (all ranks take part)
    if extras:
        n4 = args.sweep_samples or 128
        line['config4_sweep'] = run_config4(dev, pipe_factory, unet, n4, world, barrier, steps=1, warmup=0)
        line['config4_sweep']['note'] = (f'bounded grid of {n4} samples inside the default run (samples/s does not '
                                         'depend on the grid size beyond a few micro-batches); '
                                         '`--config 4` runs the full 1024')
    if world > 1:
        dist.destroy_process_group()  # all collective work is done; rank 0 alone runs the rest
    if rank == 0 and extras:
        # kernel rooflines: rank 0's GPU, after the timed region, at every N
        kr = kernel_rooflines(dev, unet, peaks)
        line['roofline'] = kr['k3f']
        line['roofline_k3'] = kr['k3']
        line['roofline_k4'] = kr['k4']
        line['roofline_k2'] = kr['k2']
        line['roofline_k1'] = kr['k1']
        line['roofline_k5'] = kr['k5']
        line['roofline_k13'] = kr['k13']
        # the microbenchmarks leave several GB of freed blocks in the caching allocator; hand them back so the configs below
        # start from a clean pool (config 3's single timed step varied 14.7 -> 11.5 images/s with the pool's history)
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and extras:
        # the other single-GPU BASELINE configs, the torch GPU baseline and the CPU baselines: N = 1 only
        line['other_configs'] = {'config2': run_config2(dev, pipe_factory, unet, steps=1, warmup=2),
                                 'config3': run_config3(dev, pipe_factory, unet, steps=1, warmup=2)}
        line['gpu_torch_baseline'] = gpu_torch_baseline(dev, unet, vae, d_uncond, d_embeds)
        line['gpu_torch_baseline']['speedup_resident_vs_eager'] = value / line['gpu_torch_baseline']['eager_images_per_s']
        line['gpu_torch_baseline']['speedup_resident_vs_cuda_graph'] = (
            value / line['gpu_torch_baseline']['cuda_graph_images_per_s'])
        base, _ = cpu_reference_arm(2, 1)
        line['cpu_baseline'] = base
        cb = cpu_blend_baseline()
        line['blends'] = {'gpu_blends_per_s': kr['k1']['blends_per_s'],
                          **_guide_rates(dev),
                          'cpu_blends_per_s': cb, 'cpu_kind': 'port (vectorised oracle; the reference\'s own '
                                                              'Tweener.tween loops: 3-10 blends/s on 8 cores, BASELINE.md section 2)',
                          'cpu_cores': os.cpu_count(),
                          'workload': '1 prompt [77,768] x 1 guide image [257,768], Linear + Clustered + Threshold '
                                      'defaults (BASELINE.json configs[0]); GPU batch 1024 prompts'}
    if rank == 0:
        print(json.dumps(line))


if __name__ == '__main__':
    main()
