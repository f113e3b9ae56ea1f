'''Host-side scheduler state machines for the fused CFG + scheduler kernel (K4).

The reference drives diffusers 0.3.0 schedulers from its loop
(/root/reference/pipeline/flex.py:177, 197-218, 233-238, 270-285): `set_format`,
`config`, `set_timesteps`, `timesteps`, `step(...).prev_sample`, `add_noise`, `sigmas`.
These classes keep exactly that surface, but `step` does no tensor arithmetic on the
host: each scheduler reduces its update to one row of scalar coefficients (computed in
float64 from the float32 beta schedule diffusers uses)

        e' = w0*eps + w1*h1 + w2*h2 + w3*h3 ,   x' = a*x + b*e' + c_noise*noise

and launches `fd_cfg_sched_step` once.  `fused_step` additionally folds the
classifier-free-guidance combine of pipeline/guide.py:61-63 into the same launch.

Restated from diffusers 0.3.0 (third-party, not vendored in the reference; no golden
vectors exist for it => parity at this boundary is pinned only against this repo's own
fp32 restatement in oracle/diffusers_shim, see DESIGN.md "parity unpinned").
'''
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np
import torch

from . import _native
from .unet import FrozenConfig


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor


@dataclass
class _Plan:
    '''One K4 launch.'''
    w: List[float]
    a: float
    b: float
    x_src: torch.Tensor
    hist: List[torch.Tensor] = field(default_factory=list)
    c_noise: float = 0.0
    keep_eps: bool = False
    in_scale: float = 1.0


def _betas(num_train_timesteps, beta_start, beta_end, beta_schedule):
    if beta_schedule == 'linear':
        return np.linspace(beta_start, beta_end, num_train_timesteps,
                           dtype=np.float32)
    if beta_schedule == 'scaled_linear':
        return np.linspace(beta_start**0.5, beta_end**0.5, num_train_timesteps,
                           dtype=np.float32)**2
    raise NotImplementedError(f'{beta_schedule} is not implemented')


class _SchedulerBase:
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085,
                 beta_end=0.012, beta_schedule='scaled_linear', **extra):
        self.config = FrozenConfig(num_train_timesteps=num_train_timesteps,
                                   beta_start=beta_start, beta_end=beta_end,
                                   beta_schedule=beta_schedule, **extra)
        self.betas = _betas(num_train_timesteps, beta_start, beta_end,
                            beta_schedule)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = np.cumprod(self.alphas, axis=0)  # float32, as diffusers
        self.num_inference_steps: Optional[int] = None
        self.timesteps: Optional[torch.Tensor] = None
        self._eps_ring: List[torch.Tensor] = []
        self._ring_pos = 0

    def set_format(self, tensor_format='pt'):
        '''flex.py:55 -- tensors are always torch here.'''
        return self

    # -- helpers ---------------------------------------------------------------
    def _ac(self, t: int) -> float:
        return float(self.alphas_cumprod[t])

    def _new_eps_slot(self, like: torch.Tensor) -> torch.Tensor:
        '''fp32 history ring (4 deep) for the multistep schedulers.'''
        if (not self._eps_ring or self._eps_ring[0].shape != like.shape
                or self._eps_ring[0].device != like.device):
            self._eps_ring = [
                torch.empty(like.shape, dtype=torch.float32, device=like.device)
                for _ in range(5)
            ]
            self._ring_pos = 0
        slot = self._eps_ring[self._ring_pos % 5]
        self._ring_pos += 1
        return slot

    def _launch(self, plan: _Plan, eps_uncond, eps_cond, guidance, use_cfg,
                noise, out, scaled_out, eps_slot):
        k = _native.SchedCoeffs()
        k.guidance = float(guidance)
        k.use_cfg = int(bool(use_cfg))
        for i in range(4):
            k.w[i] = float(plan.w[i]) if i < len(plan.w) else 0.0
        k.a, k.b = float(plan.a), float(plan.b)
        k.c_noise = float(plan.c_noise)
        k.in_scale = float(plan.in_scale)
        _native.cfg_sched_step(eps_uncond if use_cfg else None, eps_cond,
                               plan.x_src, k, out, hist=plan.hist, noise=noise,
                               eps_out=eps_slot, scaled_out=scaled_out)

    def _run(self, plan_fn, eps_uncond, eps_cond, guidance, use_cfg, sample,
             out, scaled_out, noise_fn=None):
        if sample.dtype != torch.float32:
            raise _native.NativeError('latents must be float32 '
                                      '(flex.py keeps them fp32)')
        sample = sample.contiguous()
        eps_cond = eps_cond.contiguous()
        if eps_uncond is not None:
            eps_uncond = eps_uncond.contiguous()
        plan, commit = plan_fn(sample)
        noise = noise_fn(plan) if noise_fn is not None else None
        if out is None:
            out = torch.empty_like(sample)
        eps_slot = self._new_eps_slot(sample) if plan.keep_eps else None
        self._launch(plan, eps_uncond, eps_cond, guidance, use_cfg, noise, out,
                     scaled_out, eps_slot)
        commit(eps_slot)
        return SchedulerOutput(prev_sample=out)

    def add_noise(self, original_samples, noise, timesteps):
        '''flex.py:215-218 (DDIM / PNDM form). One-off, outside the loop.'''
        t = torch.as_tensor(timesteps).reshape(-1).cpu().numpy()
        ac = torch.from_numpy(self.alphas_cumprod[t]).to(original_samples.device)
        shape = (-1,) + (1,) * (original_samples.dim() - 1)
        return (ac**0.5).view(shape) * original_samples + (
            (1 - ac)**0.5).view(shape) * noise


class DDIMScheduler(_SchedulerBase):
    '''diffusers 0.3.0 DDIMScheduler (SD config: clip_sample=False,
    set_alpha_to_one=False).'''
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085,
                 beta_end=0.012, beta_schedule='scaled_linear',
                 clip_sample=False, set_alpha_to_one=False):
        super().__init__(num_train_timesteps, beta_start, beta_end,
                         beta_schedule, clip_sample=clip_sample,
                         set_alpha_to_one=set_alpha_to_one)
        if clip_sample:
            raise NotImplementedError(
                'clip_sample=True is not linear in (x, eps); the SD v1 DDIM '
                'config uses clip_sample=False')
        self.final_alpha_cumprod = 1.0 if set_alpha_to_one else self._ac(0)
        self.set_timesteps(num_train_timesteps)

    def set_timesteps(self, num_inference_steps: int, offset: int = 0):
        self.num_inference_steps = num_inference_steps
        n = self.config.num_train_timesteps
        ts = np.arange(0, n, n // num_inference_steps)[::-1].copy() + offset
        self.timesteps = torch.from_numpy(ts.astype(np.int64))

    def coefficients(self, t: int, eta: float = 0.0):
        '''(a, b, c_noise) with x' = a x + b eps + c_noise z.'''
        n = self.config.num_train_timesteps
        prev_t = t - n // self.num_inference_steps
        ac_t = self._ac(t)
        ac_p = self._ac(prev_t) if prev_t >= 0 else self.final_alpha_cumprod
        var = (1 - ac_p) / (1 - ac_t) * (1 - ac_t / ac_p)
        sigma = eta * var**0.5
        a = (ac_p / ac_t)**0.5
        b = (1 - ac_p - sigma**2)**0.5 - ac_p**0.5 * (1 - ac_t)**0.5 / ac_t**0.5
        return a, b, sigma

    def _plan(self, t, eta):
        a, b, sigma = self.coefficients(int(t), eta)

        def plan_fn(sample):
            return _Plan(w=[1.0], a=a, b=b, x_src=sample,
                         c_noise=sigma), (lambda slot: None)

        return plan_fn

    def fused_step(self, eps_uncond, eps_cond, guidance, use_cfg, timestep,
                   sample, eta: float = 0.0, generator=None, out=None,
                   scaled_out=None):
        def noise_fn(plan):
            if plan.c_noise == 0.0:
                return None
            # diffusers 0.3.0 draws the eta noise on the host (`step` never sees the device);
            # a CUDA generator draws in place, anything else on its own device then moves
            gdev = generator.device if generator is not None else torch.device('cpu')
            return torch.randn(sample.shape, generator=generator, device=gdev,
                               dtype=torch.float32).to(sample.device)

        return self._run(self._plan(timestep, eta), eps_uncond, eps_cond,
                         guidance, use_cfg, sample, out, scaled_out, noise_fn)

    def step(self, model_output, timestep, sample, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None):
        return self.fused_step(None, model_output, 1.0, False, timestep,
                               sample, eta, generator)


class PNDMScheduler(_SchedulerBase):
    '''diffusers 0.3.0 PNDMScheduler, PLMS branch (SD config: skip_prk_steps=True).
    Quirks kept on purpose: steps+1 model evaluations (repeated second
    timestep) and the warm-up restarting wherever an img2img slice begins.'''
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085,
                 beta_end=0.012, beta_schedule='scaled_linear',
                 skip_prk_steps=True):
        super().__init__(num_train_timesteps, beta_start, beta_end,
                         beta_schedule, skip_prk_steps=skip_prk_steps)
        if not skip_prk_steps:
            raise NotImplementedError('Runge-Kutta warm-up (skip_prk_steps='
                                      'False) is not used by the SD v1 config')
        self.ets: List[torch.Tensor] = []
        self.counter = 0
        self.cur_sample = None
        self._offset = 0
        self.set_timesteps(num_train_timesteps)

    def set_timesteps(self, num_inference_steps: int, offset: int = 0):
        self.num_inference_steps = num_inference_steps
        n = self.config.num_train_timesteps
        base = np.array(list(range(0, n, n // num_inference_steps))) + offset
        self._offset = offset
        plms = np.concatenate([base[:-1], base[-2:-1], base[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.ets, self.counter, self.cur_sample = [], 0, None

    def prev_sample_coefficients(self, t: int, t_prev: int):
        ac_t = self._ac(t + 1 - self._offset)
        ac_p = self._ac(t_prev + 1 - self._offset)
        a = (ac_p / ac_t)**0.5
        denom = ac_t * (1 - ac_p)**0.5 + (ac_t * (1 - ac_t) * ac_p)**0.5
        b = -(ac_p - ac_t) / denom
        return a, b

    def _plan(self, timestep):
        t = int(timestep)
        ratio = self.config.num_train_timesteps // self.num_inference_steps

        def plan_fn(sample):
            prev_t = max(t - ratio, 0)
            tt = t
            keep = self.counter != 1
            n_ets = len(self.ets) + (1 if keep else 0)
            x_src = sample
            if not keep:
                prev_t, tt = t, t + ratio
            if n_ets == 1 and self.counter == 0:
                w, hist = [1.0], []
            elif n_ets == 1 and self.counter == 1:
                w, hist = [0.5, 0.5], [self.ets[-1]]
                x_src = self.cur_sample
            elif n_ets == 2:
                w, hist = [1.5, -0.5], [self.ets[-1]]
            elif n_ets == 3:
                w = [23 / 12, -16 / 12, 5 / 12]
                hist = [self.ets[-1], self.ets[-2]]
            else:
                w = [55 / 24, -59 / 24, 37 / 24, -9 / 24]
                hist = [self.ets[-1], self.ets[-2], self.ets[-3]]
            a, b = self.prev_sample_coefficients(tt, prev_t)

            def commit(slot):
                if keep:
                    self.ets = (self.ets + [slot])[-4:]
                if self.counter == 0:
                    self.cur_sample = sample
                elif self.counter == 1:
                    self.cur_sample = None
                self.counter += 1

            return _Plan(w=w, a=a, b=b, x_src=x_src, hist=hist,
                         keep_eps=keep), commit

        return plan_fn

    def fused_step(self, eps_uncond, eps_cond, guidance, use_cfg, timestep,
                   sample, out=None, scaled_out=None, **_):
        if out is not None and self.counter == 0 and out is sample:
            raise ValueError('PLMS step 0 must not overwrite its input '
                             '(cur_sample is reused by step 1)')
        return self._run(self._plan(timestep), eps_uncond, eps_cond, guidance,
                         use_cfg, sample, out, scaled_out)

    def step(self, model_output, timestep, sample):
        return self.fused_step(None, model_output, 1.0, False, timestep, sample)


class LMSDiscreteScheduler(_SchedulerBase):
    '''diffusers 0.3.0 LMSDiscreteScheduler (order 4).  `step` takes the step INDEX
    (flex.py:271,280-284).  Order 1 is the Euler update in sigma space.'''
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085,
                 beta_end=0.012, beta_schedule='scaled_linear', order=4):
        super().__init__(num_train_timesteps, beta_start, beta_end,
                         beta_schedule)
        self.order = order
        self.derivatives: List[torch.Tensor] = []
        self._all_sigmas = ((1 - self.alphas_cumprod) /
                            self.alphas_cumprod)**0.5
        self.sigmas = torch.from_numpy(
            np.concatenate([self._all_sigmas[::-1], [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(
            np.arange(0, num_train_timesteps)[::-1].copy().astype(np.float64))

    def set_timesteps(self, num_inference_steps: int):
        self.num_inference_steps = num_inference_steps
        n = self.config.num_train_timesteps
        ts = np.linspace(n - 1, 0, num_inference_steps, dtype=float)
        low, high = np.floor(ts).astype(int), np.ceil(ts).astype(int)
        frac = np.mod(ts, 1.0)
        sig = np.array(self._all_sigmas)
        sig = (1 - frac) * sig[low] + frac * sig[high]
        self.sigmas = torch.from_numpy(
            np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.derivatives = []

    def lms_coefficient(self, order: int, t: int, current_order: int) -> float:
        from scipy import integrate
        sig = self.sigmas.numpy().astype(np.float64)

        def deriv(tau):
            prod = 1.0
            for k in range(order):
                if k == current_order:
                    continue
                prod *= (tau - sig[t - k]) / (sig[t - current_order] -
                                              sig[t - k])
            return prod

        return integrate.quad(deriv, sig[t], sig[t + 1], epsrel=1e-4)[0]

    def input_scale(self, t_index: int) -> float:
        '''flex.py:272-274: model input = latents / sqrt(sigma^2 + 1).'''
        s = float(self.sigmas[t_index])
        return 1.0 / (s * s + 1.0)**0.5

    def _plan(self, t_index, order):
        t = int(t_index)

        def plan_fn(sample):
            o = min(t + 1, order)
            coeffs = [self.lms_coefficient(o, t, c) for c in range(o)]
            # zip(coeffs, reversed(derivatives)) with the current derivative (= eps) first
            hist = list(reversed(self.derivatives))[:min(o, 4) - 1]
            w = coeffs[:1 + len(hist)]
            nxt = t + 1
            scale = self.input_scale(nxt) if nxt < len(self.sigmas) else 1.0

            def commit(slot):
                self.derivatives = (self.derivatives + [slot])[-order:]

            return _Plan(w=w, a=1.0, b=1.0, x_src=sample, hist=hist,
                         keep_eps=True, in_scale=scale), commit

        return plan_fn

    def fused_step(self, eps_uncond, eps_cond, guidance, use_cfg, timestep,
                   sample, order: int = 4, out=None, scaled_out=None, **_):
        if order > 4:
            raise NotImplementedError('K4 mixes at most 4 derivatives')
        return self._run(self._plan(timestep, order), eps_uncond, eps_cond,
                         guidance, use_cfg, sample, out, scaled_out)

    def step(self, model_output, timestep, sample, order: int = 4):
        return self.fused_step(None, model_output, 1.0, False, timestep,
                               sample, order)

    def add_noise(self, original_samples, noise, timesteps):
        t = torch.as_tensor(timesteps).reshape(-1).cpu().long()
        sig = self.sigmas[t].to(original_samples.device)
        shape = (-1,) + (1,) * (original_samples.dim() - 1)
        return original_samples + noise * sig.view(shape)


class EulerDiscreteScheduler(LMSDiscreteScheduler):
    '''Euler update in sigma space: x' = x + (sigma[i+1] - sigma[i]) * eps.  EXTENSION -- the
    reference pins diffusers 0.3.0, which has no Euler class; BASELINE.json's north_star names
    one, and it is exactly the order-1 case of the LMS scheduler the reference drives at
    /root/reference/pipeline/flex.py:271-284 (same sigma grid, same input scaling, same
    step-index convention), so it is checked against LMSDiscreteScheduler(order=1): one K4
    launch per step, no derivative history, the coefficient in closed form.'''
    def lms_coefficient(self, order: int, t: int, current_order: int) -> float:
        assert order == 1 and current_order == 0
        sig = self.sigmas.numpy().astype(np.float64)
        return float(sig[t + 1] - sig[t])

    def fused_step(self, eps_uncond, eps_cond, guidance, use_cfg, timestep,
                   sample, order: int = 1, out=None, scaled_out=None, **_):
        return super().fused_step(eps_uncond, eps_cond, guidance, use_cfg,
                                  timestep, sample, 1, out, scaled_out)

    def step(self, model_output, timestep, sample, order: int = 1):
        return self.fused_step(None, model_output, 1.0, False, timestep, sample)
