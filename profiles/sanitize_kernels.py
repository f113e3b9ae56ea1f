'''Development aid: one small launch of every hand-written kernel, for
   compute-sanitizer --tool memcheck|racecheck|initcheck python profiles/sanitize_kernels.py'''
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flexdiffuse_b200 import _native  # noqa: E402

dev = torch.device('cuda:0')
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s: torch.randn(*s, device=dev, generator=g)

# K4
u, c, x = rn(2, 4, 16, 16), rn(2, 4, 16, 16), rn(2, 4, 16, 16)
k = _native.SchedCoeffs()
k.guidance, k.use_cfg, k.a, k.b = 7.5, 1, 1.0, -0.1
k.w[0] = 1.0
_native.cfg_sched_step(u.bfloat16(), c.bfloat16(), x, k, torch.empty_like(x), eps_out=torch.empty_like(x),
                       scaled_out=torch.empty(2, 4, 16, 16, device=dev, dtype=torch.bfloat16))
# K2 (one full tile + edges)
ctx = rn(160, 128).bfloat16()
w = rn(328, 128).bfloat16()
kv = _native.kv_project(ctx, w)
# K3 (d = 40, 80, 160; partial query tile; 48 samples -> one CTA walks 6 tiles with both warpgroups,
# the Q ring wraps and the in-place output staging is reused)
for C, nq, S in ((320, 200, 2), (1280, 64, 2), (320, 768, 48), (640, 640, 48)):
    q = rn(S, nq, C).bfloat16()
    kvc = rn(160, 2 * C).bfloat16()
    _native.cross_attn(q, kvc, 0, C, (torch.arange(S, dtype=torch.int32, device=dev) + 1) % 2, 8, 77, 80,
                       (C // 8)**-0.5)
# K1 (257 guide tokens: TMA + remainder row; all modes)
txt, img = rn(2, 77, 64), rn(1, 257, 64)
prms = []
for mode in (0, 1, 2):
    for reuse in (0, 1):
        p = _native.TweenParams()
        p.threshold_floor = p.threshold_mult = p.max_guidance = 0.5
        p.clustered, p.header_max, p.align_mode, p.mapping_reuse = 0.0, 0.15, mode, reuse
        prms.append(p)
lin = torch.linspace(0, 0.5, 77)[None].repeat(len(prms), 1).to(dev)
_native.sim_blend(txt, img, prms, lin, want_sim=True)
# K5 (cluster and streaming paths), K6, K7, K8, K9
xs = rn(2, 320, 16, 16).bfloat16().contiguous(memory_format=torch.channels_last)
gam, bet = torch.ones(320, device=dev).bfloat16(), torch.zeros(320, device=dev).bfloat16()
_native.groupnorm_act(xs, gam, bet, 32, 1e-5, True, bias=rn(2, 320).bfloat16())
os.environ['FD_GN_CLUSTER_MAX_BYTES'] = '0'
xl = rn(3, 128, 40, 40).bfloat16().contiguous(memory_format=torch.channels_last)
_native.groupnorm_act(xl, torch.ones(128, device=dev).bfloat16(), torch.zeros(128, device=dev).bfloat16(), 32,
                      1e-6, False)
_native.geglu(rn(77, 2560).bfloat16())
_native.add_bias_residual(xs, xs.clone(), gam)
_native.add_layernorm(rn(133, 640).bfloat16(), rn(133, 640).bfloat16(), torch.ones(640, device=dev).bfloat16(),
                      torch.zeros(640, device=dev).bfloat16(), 1e-5)
b = _native.EntityBox()
b.ox, b.oy, b.sx, b.sy, b.blend = 2, 3, 8, 20, 0.8
_native.composite_eps(rn(3, 4, 16, 16).bfloat16(), [b])
# K1B (batched path: 5 prompts -> a CTA with one full and one single-prompt batch), K1P, K3F, K10, K11 (split-K cluster, LN split), K12
_native.lib().fd_debug_set_k1_fast(1, 1)
p = _native.TweenParams()
p.threshold_floor = p.threshold_mult = p.max_guidance = p.clustered = 0.5
p.header_max, p.align_mode, p.mapping_reuse = 0.15, 1, 1
_native.sim_blend(rn(5, 77, 768), rn(1, 257, 768), [p, p], torch.linspace(0, 0.5, 77)[None].repeat(2, 1).to(dev))
_native.lib().fd_debug_set_k1_fast(1, 64)
_native.visual_projection(rn(130, 192), rn(260, 192) * 0.03)
for C, nq in ((320, 200), (640, 130)):
    d_kv = rn(160, 2 * C).bfloat16()
    _native.cross_attn_fused(rn(2, nq, C).bfloat16(), (rn(C, C) * C ** -0.5).bfloat16(), d_kv, 0, C,
                             torch.tensor([0, 1], dtype=torch.int32, device=dev), (rn(C, C) * C ** -0.5).bfloat16(),
                             rn(C).bfloat16(), 8, 77, 80, (C // 8) ** -0.5)
_native.image_tail_u8(rn(1, 3, 24, 24))
xa, wa, ba = rn(130, 256), rn(260, 256) * 0.05, rn(260)
op = _native.x3_split_ln(xa, torch.ones(256, device=dev), torch.zeros(256, device=dev), 1e-5)
for sk in (1, 2, 4):
    _native.linear_x3(None, wa, ba, act=1, operand=op, rows=130, residual=rn(130, 260), split_k=sk)
_native.linear_x3(xa, wa, ba)
qkv = rn(2, 77, 3 * 128)
_native.attention_f32(qkv[..., :128], qkv[..., 128:256], qkv[..., 256:], 2, 0.125, True)
qkv = rn(1, 257, 3 * 64)
_native.attention_f32(qkv[..., :64], qkv[..., 64:128], qkv[..., 128:], 4, 0.25, False)
torch.cuda.synchronize()
print('all kernels launched once')
