'''ORACLE (test infrastructure): fp32 restatement of the reference's denoising loop

    /root/reference/pipeline/guide.py:46-64    SimpleGuide.noise_pred (CFG batch + combine)
    /root/reference/pipeline/flex.py:170-310   FlexPipeline.__call__   (latent init, t_start,
                                               per-step scheduler update, decode)

for machines where /root/reference is not mounted (the GPU box).  In the build
container `tests/test_loop_oracle_pinning.py` runs the UNMODIFIED reference
`pipeline/flex.py` + `pipeline/guide.py` on top of oracle/diffusers_shim and checks this
restatement against it bit for bit, so the control flow is pinned to the reference; the
diffusers arithmetic underneath both is this repo's restatement (parity unpinned there).
'''
from __future__ import annotations

import os
import sys
from typing import Callable, List, Optional

import torch

_SHIM = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'diffusers_shim')
if _SHIM not in sys.path:
    sys.path.insert(0, _SHIM)

from diffusers.schedulers import (DDIMScheduler, LMSDiscreteScheduler,  # noqa: E402
                                  PNDMScheduler)

__all__ = ['DDIMScheduler', 'LMSDiscreteScheduler', 'PNDMScheduler',
           'noise_pred', 'denoise', 'composite_noise_pred']


def noise_pred(unet_fn: Callable, uncond: torch.Tensor, embeds: torch.Tensor,
               guidance: float, latents: torch.Tensor, step):
    '''guide.py:46-64.  unet_fn(latents, t, ctx) -> eps.'''
    B = embeds.shape[0]
    cfg = guidance > 1.0
    ctx = embeds
    if cfg:
        ctx = torch.cat(([uncond] * B) + [embeds])
        latents = torch.cat([latents] * 2)
    eps = unet_fn(latents, step, ctx)
    if cfg:
        u, c = eps.chunk(2)
        eps = u + guidance * (c - u)
    return eps


def denoise(unet_fn: Callable, scheduler, uncond: torch.Tensor,
            embeds: torch.Tensor, guidance: float, steps: int,
            init_latents: Optional[torch.Tensor] = None,
            init_size=(512, 512), strength: float = 0.6, eta: float = 0.0,
            generator: Optional[torch.Generator] = None,
            device='cpu', trace: Optional[List] = None) -> torch.Tensor:
    '''flex.py:170-287 up to (not including) the decode.

    init_latents: VAE-encoded, 0.18215-scaled init image latents [1,4,h,w]
    (flex.py:189-192) or None for txt2img.  `trace` collects (eps, latents) per step.'''
    if strength < 0 or strength > 1:
        raise ValueError(
            f'The value of strength should in [0.0, 1.0] but is {strength}')
    B = embeds.shape[0]
    is_lms = isinstance(scheduler, LMSDiscreteScheduler)
    scheduler.set_timesteps(steps)
    if init_latents is not None:
        lat = torch.cat([init_latents] * B)
        offset = scheduler.config.get('steps_offset', 0)
        init_timestep = min(int(steps * strength) + offset, steps)
        if is_lms:
            ts = torch.tensor([steps - init_timestep] * B, dtype=torch.long)
        else:
            ts = torch.tensor([int(scheduler.timesteps[-init_timestep])] * B,
                              dtype=torch.long)
        noise = torch.randn(lat.shape, generator=generator, device=device)
        lat = scheduler.add_noise(lat, noise, ts)
        t_start = max(steps - init_timestep + offset, 0)
    else:
        h, w = init_size
        lat = torch.randn((B, 4, h // 8, w // 8), generator=generator,
                          device=device)
        scheduler.set_timesteps(steps)
        if is_lms:
            lat = lat * scheduler.sigmas[0]
        t_start = 0
    extra = {'eta': eta} if isinstance(scheduler, DDIMScheduler) else {}
    for i, t in enumerate(scheduler.timesteps[t_start:]):
        t_index = t
        model_in = lat
        if is_lms:
            t_index = t_start + i
            sigma = scheduler.sigmas[t_index]
            model_in = model_in / ((sigma**2 + 1)**0.5)
        eps = noise_pred(unet_fn, uncond, embeds, guidance, model_in, t)
        lat = scheduler.step(eps, t_index, lat, **extra).prev_sample
        if trace is not None:
            trace.append((eps, lat))
    return lat


def composite_noise_pred(unet_fn: Callable, uncond: torch.Tensor, background: torch.Tensor,
                         entities, guidance: float, latents: torch.Tensor, step):
    '''/root/reference/composition/guide.py:58-95 `CompositeGuide._guide_latents` (batch 1):
    one UNet call on (uncond, background, entity embeds) over the same latents, rectangular
    lerp of each entity's prediction into the background's, then CFG.

    entities: list of (embed [1,77,768], (ox, oy) blocks, (sx, sy) blocks, blend).'''
    cfg = guidance > 1.0
    conds = [background] + [e[0] for e in entities]
    ctx = torch.cat(([uncond] if cfg else []) + conds)
    eps = unet_fn(torch.cat([latents] * ctx.shape[0]), step, ctx)
    stack = eps[1:] if cfg else eps
    out = stack[:1].clone()
    ent = stack[1:]
    for ei, (_, (ow, oh), (sw, sh), blend) in enumerate(entities):
        bw, bh = ow + sw, oh + sh
        bgs = out[:, :, oh:bh, ow:bw]
        out[:, :, oh:bh, ow:bw] = bgs + blend * (ent[ei:ei + 1, :, oh:bh, ow:bw] - bgs)
    if cfg:
        u = eps[:1]
        out = u + guidance * (out - u)
    return out
