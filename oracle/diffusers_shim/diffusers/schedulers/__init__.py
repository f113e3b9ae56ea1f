'''ORACLE restatement of the diffusers 0.3.0 schedulers the reference drives from
/root/reference/pipeline/flex.py:177,197-218,233-238,270-285: DDIMScheduler,
PNDMScheduler (PLMS branch, skip_prk_steps) and LMSDiscreteScheduler.

Transcribed formula by formula in diffusers' own form (float32 beta schedule, numpy
scalar x fp32 tensor products, several elementwise ops per step) -- deliberately NOT
the single coefficient row the product's K4 planner uses, so the two can be compared.
Restated from the published 0.3.0 algorithms; diffusers is not installable here =>
PARITY UNPINNED at this boundary.'''
from dataclasses import dataclass

import numpy as np
import torch
from scipy import integrate

from ..configuration_utils import FrozenDict


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor


def _betas(n, start, end, schedule):
    if schedule == 'linear':
        return np.linspace(start, end, n, dtype=np.float32)
    if schedule == 'scaled_linear':
        return np.linspace(start**0.5, end**0.5, n, dtype=np.float32)**2
    raise NotImplementedError(schedule)


class _Mixin:
    def _setup(self, **cfg):
        self.config = FrozenDict(cfg)
        self.betas = _betas(cfg['num_train_timesteps'], cfg['beta_start'],
                            cfg['beta_end'], cfg['beta_schedule'])
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = np.cumprod(self.alphas, axis=0)
        self.num_inference_steps = None

    def set_format(self, tensor_format='pt'):
        return self

    def add_noise(self, original_samples, noise, timesteps):
        ac = torch.from_numpy(self.alphas_cumprod)[timesteps.cpu()].to(
            original_samples.device)
        sa = (ac**0.5).view(-1, 1, 1, 1)
        sb = ((1 - ac)**0.5).view(-1, 1, 1, 1)
        return sa * original_samples + sb * noise


class DDIMScheduler(_Mixin):
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085,
                 beta_end=0.012, beta_schedule='scaled_linear',
                 clip_sample=False, set_alpha_to_one=False):
        self._setup(num_train_timesteps=num_train_timesteps,
                    beta_start=beta_start, beta_end=beta_end,
                    beta_schedule=beta_schedule, clip_sample=clip_sample,
                    set_alpha_to_one=set_alpha_to_one)
        self.final_alpha_cumprod = (np.array(1.0, dtype=np.float32)
                                    if set_alpha_to_one else
                                    self.alphas_cumprod[0])
        self.timesteps = torch.from_numpy(
            np.arange(0, num_train_timesteps)[::-1].copy())

    def set_timesteps(self, num_inference_steps, offset=0):
        self.num_inference_steps = num_inference_steps
        n = self.config['num_train_timesteps']
        ts = np.arange(0, n, n // num_inference_steps)[::-1].copy()
        self.timesteps = torch.from_numpy(ts + offset)

    def _get_variance(self, t, prev_t):
        a_t = self.alphas_cumprod[t]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        return (1 - a_p) / (1 - a_t) * (1 - a_t / a_p)

    def step(self, model_output, timestep, sample, eta=0.0,
             use_clipped_model_output=False, generator=None):
        timestep = int(timestep)
        prev_t = timestep - self.config['num_train_timesteps'] // self.num_inference_steps
        a_t = self.alphas_cumprod[timestep]
        a_p = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        b_t = 1 - a_t
        pred_x0 = (sample - b_t**0.5 * model_output) / a_t**0.5
        if self.config['clip_sample']:
            pred_x0 = torch.clamp(pred_x0, -1, 1)
        std = eta * self._get_variance(timestep, prev_t)**0.5
        direction = (1 - a_p - std**2)**0.5 * model_output
        prev = a_p**0.5 * pred_x0 + direction
        if eta > 0:
            noise = torch.randn(model_output.shape, generator=generator,
                                device=model_output.device)
            prev = prev + std * noise
        return SchedulerOutput(prev)


class PNDMScheduler(_Mixin):
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085,
                 beta_end=0.012, beta_schedule='scaled_linear',
                 skip_prk_steps=True):
        self._setup(num_train_timesteps=num_train_timesteps,
                    beta_start=beta_start, beta_end=beta_end,
                    beta_schedule=beta_schedule, skip_prk_steps=skip_prk_steps)
        assert skip_prk_steps, 'only the PLMS branch is restated (SD v1 config)'
        self.cur_sample = None
        self.ets = []
        self.counter = 0
        self._offset = 0
        self.timesteps = torch.from_numpy(
            np.arange(0, num_train_timesteps)[::-1].copy())

    def set_timesteps(self, num_inference_steps, offset=0):
        self.num_inference_steps = num_inference_steps
        n = self.config['num_train_timesteps']
        self._timesteps = list(range(0, n, n // num_inference_steps))
        self._offset = offset
        self._timesteps = np.array([t + offset for t in self._timesteps])
        plms = np.concatenate([self._timesteps[:-1], self._timesteps[-2:-1],
                               self._timesteps[-1:]])[::-1].copy()
        self.timesteps = torch.from_numpy(plms.astype(np.int64))
        self.ets = []
        self.counter = 0

    def step(self, model_output, timestep, sample):
        timestep = int(timestep)
        ratio = self.config['num_train_timesteps'] // self.num_inference_steps
        prev_timestep = max(timestep - ratio, 0)
        if self.counter != 1:
            self.ets.append(model_output)
        else:
            prev_timestep = timestep
            timestep = timestep + ratio
        if len(self.ets) == 1 and self.counter == 0:
            self.cur_sample = sample
        elif len(self.ets) == 1 and self.counter == 1:
            model_output = (model_output + self.ets[-1]) / 2
            sample = self.cur_sample
            self.cur_sample = None
        elif len(self.ets) == 2:
            model_output = (3 * self.ets[-1] - self.ets[-2]) / 2
        elif len(self.ets) == 3:
            model_output = (23 * self.ets[-1] - 16 * self.ets[-2] + 5 * self.ets[-3]) / 12
        else:
            model_output = (1 / 24) * (55 * self.ets[-1] - 59 * self.ets[-2] +
                                       37 * self.ets[-3] - 9 * self.ets[-4])
        prev = self._get_prev_sample(sample, timestep, prev_timestep, model_output)
        self.counter += 1
        return SchedulerOutput(prev)

    def _get_prev_sample(self, sample, timestep, timestep_prev, model_output):
        a_t = self.alphas_cumprod[timestep + 1 - self._offset]
        a_p = self.alphas_cumprod[timestep_prev + 1 - self._offset]
        b_t, b_p = 1 - a_t, 1 - a_p
        sample_coeff = (a_p / a_t)**0.5
        denom = a_t * b_p**0.5 + (a_t * b_t * a_p)**0.5
        return sample_coeff * sample - (a_p - a_t) * model_output / denom


class LMSDiscreteScheduler(_Mixin):
    def __init__(self, num_train_timesteps=1000, beta_start=0.00085,
                 beta_end=0.012, beta_schedule='scaled_linear'):
        self._setup(num_train_timesteps=num_train_timesteps,
                    beta_start=beta_start, beta_end=beta_end,
                    beta_schedule=beta_schedule)
        self._sig_all = ((1 - self.alphas_cumprod) / self.alphas_cumprod)**0.5
        self.sigmas = torch.from_numpy(self._sig_all.copy())
        self.timesteps = torch.from_numpy(
            np.arange(0, num_train_timesteps)[::-1].copy())
        self.derivatives = []

    def get_lms_coefficient(self, order, t, current_order):
        sig = self.sigmas.numpy()

        def lms_derivative(tau):
            prod = 1.0
            for k in range(order):
                if current_order == k:
                    continue
                prod *= (tau - sig[t - k]) / (sig[t - current_order] - sig[t - k])
            return prod

        return integrate.quad(lms_derivative, sig[t], sig[t + 1], epsrel=1e-4)[0]

    def set_timesteps(self, num_inference_steps):
        self.num_inference_steps = num_inference_steps
        n = self.config['num_train_timesteps']
        ts = np.linspace(n - 1, 0, num_inference_steps, dtype=float)
        low, high = np.floor(ts).astype(int), np.ceil(ts).astype(int)
        frac = np.mod(ts, 1.0)
        sig = np.array(self._sig_all)
        sig = (1 - frac) * sig[low] + frac * sig[high]
        self.sigmas = torch.from_numpy(
            np.concatenate([sig, [0.0]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts)
        self.derivatives = []

    def step(self, model_output, timestep, sample, order=4):
        timestep = int(timestep)
        sigma = self.sigmas[timestep]
        pred_x0 = sample - sigma * model_output
        derivative = (sample - pred_x0) / sigma
        self.derivatives.append(derivative)
        if len(self.derivatives) > order:
            self.derivatives.pop(0)
        order = min(timestep + 1, order)
        coeffs = [self.get_lms_coefficient(order, timestep, c) for c in range(order)]
        prev = sample + sum(c * d for c, d in zip(coeffs, reversed(self.derivatives)))
        return SchedulerOutput(prev)

    def add_noise(self, original_samples, noise, timesteps):
        sig = self.sigmas[timesteps.cpu()].to(original_samples.device)
        return original_samples + noise * sig.view(-1, 1, 1, 1)
