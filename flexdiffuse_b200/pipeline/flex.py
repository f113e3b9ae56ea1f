'''Denoising pipeline -- API mirror of /root/reference/pipeline/flex.py (`FlexPipeline`,
itself a modification of diffusers' Img2Img pipeline).

Same constructor, `__call__` signature, latent initialisation, `t_start` arithmetic and
decode as flex.py:46-83,126-310.  What changes is the hot loop (flex.py:262-287):

  reference, per step : cat -> UNet (re-projects K/V in 16 layers) -> chunk -> 3 elementwise
                        CFG ops -> scheduler.step (5-15 tiny elementwise ops + host scalar math)
  here, per step      : [graph replay of] UNet over cached K/V (K3)  ->  ONE K4 launch doing
                        CFG + scheduler update (+ the LMS model-input pre-scale, flex.py:270-274)

The fast path is taken when the guide exposes `noise_pred_pair` and the scheduler
`fused_step` (flexdiffuse_b200's own classes); any other GuideBase / scheduler object
goes through the reference's call sequence unchanged.
'''
from __future__ import annotations

import inspect
import warnings
from dataclasses import dataclass
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from .. import _native
from ..encode.clip import preprocess
from ..schedulers import LMSDiscreteScheduler
from ..unet import FrozenConfig
from .guide import GuideBase


class StableDiffusionPipelineOutput(dict):
    '''images + nsfw flags; attribute and key access like diffusers' BaseOutput
    (flex.py:308-310).  `output['sample']` is kept as an alias of `images`.'''
    def __init__(self, images, nsfw_content_detected):
        super().__init__(images=images,
                         nsfw_content_detected=nsfw_content_detected)
        self.images = images
        self.nsfw_content_detected = nsfw_content_detected

    def __getitem__(self, k):
        return super().__getitem__('images' if k == 'sample' else k)


class FlexPipeline():
    r'''Text / image guided image generation with Stable Diffusion (see module doc).
    Args: vae (AutoencoderKL), clip (CLIPModel), tokenizer (CLIPTokenizer),
    unet (UNet2DConditionModel), scheduler (DDIM / PNDM / LMSDiscrete).'''
    def __init__(self, vae, clip, tokenizer, unet, scheduler):
        scheduler = scheduler.set_format('pt')
        if (hasattr(scheduler.config, 'steps_offset')
                and scheduler.config['steps_offset'] != 1):
            warnings.warn(
                f'The configuration file of this scheduler: {scheduler} is '
                'outdated. `steps_offset` should be set to 1 instead of '
                f'{scheduler.config["steps_offset"]}.', DeprecationWarning)
            new_config = dict(scheduler.config)
            new_config['steps_offset'] = 1
            scheduler.config = FrozenConfig(new_config)
        self.register_modules(vae=vae, clip=clip, tokenizer=tokenizer,
                              unet=unet, scheduler=scheduler)

    # -- the slice of diffusers.DiffusionPipeline the reference relies on ---------
    def register_modules(self, **modules):
        self._modules = list(modules)
        for k, v in modules.items():
            setattr(self, k, v)

    @property
    def device(self) -> torch.device:
        for name in self._modules:
            m = getattr(self, name)
            if isinstance(m, torch.nn.Module):
                return next(m.parameters()).device
        return torch.device('cpu')

    def to(self, device):
        for name in self._modules:
            m = getattr(self, name)
            if isinstance(m, torch.nn.Module):
                m.to(device)
        return self

    def progress_bar(self, iterable):
        return iterable

    @staticmethod
    def numpy_to_pil(images: np.ndarray):
        from PIL import Image
        if images.ndim == 3:
            images = images[None, ...]
        images = (images * 255).round().astype('uint8')
        return [Image.fromarray(image) for image in images]

    def enable_attention_slicing(self, slice_size: Optional[Union[str, int]] = 'auto'):
        '''Memory knob of the reference (flex.py:85-102); a no-op for K3, which never
        materialises the score tensor.'''
        if slice_size == 'auto':
            slice_size = self.unet.config['attention_head_dim'] // 2
        self.unet.set_attention_slice(slice_size)

    def disable_attention_slicing(self):
        self.enable_attention_slicing(None)

    def _latents_to_image(self, latents: torch.Tensor, pil: bool = True):
        # flex.py:112-124
        latents = 1 / 0.18215 * latents
        image = self.vae.decode(latents).sample
        if pil and image.is_cuda:
            # K10: scale / clamp / * 255 / round / uint8 NHWC in one pass; 3 bytes per pixel cross
            # PCIe instead of 12 (same bytes as numpy_to_pil makes from the fp32 array)
            from PIL import Image
            u8 = _native.image_tail_u8(image).cpu().numpy()
            return [Image.fromarray(a) for a in u8]
        image = (image / 2 + 0.5).clamp(0, 1)
        image = image.float().cpu().permute(0, 2, 3, 1).numpy()
        if pil:
            return self.numpy_to_pil(image)
        return image

    @torch.no_grad()
    def __call__(self,
                 guide: GuideBase,
                 init_image=None,
                 init_size: Tuple[int, int] = (512, 512),
                 strength: float = 0.6,
                 eta: float = 0.0,
                 generator: Optional[torch.Generator] = None,
                 output_type: str = 'pil',
                 return_dict: bool = True,
                 debug: bool = False,
                 latents: Optional[torch.Tensor] = None):
        '''Arguments as flex.py:127-168, plus `latents`: pre-drawn initial noise [B,4,h,w]
        (txt2img only; used by the sweep driver for shard-invariant per-sample seeds).  `output_type='latent'` additionally returns the
        final latents without decoding (sweep driver), `'pt'` the decoded images as a
        device tensor and `'uint8'` the [B,H,W,3] uint8 host array PIL would wrap.'''
        if strength < 0 or strength > 1:
            raise ValueError(
                f'The value of strength should in [0.0, 1.0] but is {strength}')
        batch_size = guide.batch_size
        sched = self.scheduler
        is_lms = isinstance(sched, LMSDiscreteScheduler) or (
            type(sched).__name__ == 'LMSDiscreteScheduler')
        sched.set_timesteps(guide.steps)
        assert sched.timesteps is not None

        if init_image is not None and not (torch.is_tensor(init_image)
                                           and init_image.numel() == 0):
            if not torch.is_tensor(init_image):
                init_image = preprocess(init_image)
            init_image = init_image.to(self.device)
            init_latents = self.vae.encode(init_image).latent_dist.sample(
                generator=generator)
            init_latents = (0.18215 * init_latents).float()
            init_latents = torch.cat([init_latents] * batch_size)
            offset = sched.config.get('steps_offset', 0)
            init_timestep = min(int(guide.steps * strength) + offset, guide.steps)
            if is_lms:
                timesteps = torch.tensor([guide.steps - init_timestep] *
                                         batch_size, dtype=torch.long)
            else:
                timesteps = torch.tensor(
                    [int(sched.timesteps[-init_timestep])] * batch_size,
                    dtype=torch.long)
            noise = torch.randn(init_latents.shape, generator=generator,
                                device=self.device)
            init_latents = sched.add_noise(init_latents, noise, timesteps)
            t_start = max(guide.steps - init_timestep + offset, 0)
        else:
            height, width = init_size
            shape = (batch_size, self.unet.in_channels, height // 8, width // 8)
            if latents is not None:
                if tuple(latents.shape) != shape:
                    raise ValueError(f'latents must be {shape}, got {tuple(latents.shape)}')
                # never write into the caller's tensor: it becomes a ping-pong buffer of the loop
                init_latents = latents.to(self.device, torch.float32).clone()
            elif generator is not None and generator.device.type == 'cpu':
                # host-side RNG (reproducible across devices): draw on the host, copy once
                init_latents = torch.randn(shape, generator=generator).pin_memory().to(
                    self.device, non_blocking=True)
            else:
                init_latents = torch.randn(shape, generator=generator,
                                           device=self.device)
            sched.set_timesteps(guide.steps)
            if is_lms:
                init_latents = init_latents * sched.sigmas[0].to(self.device)
            t_start = 0

        accepts_eta = 'eta' in set(
            inspect.signature(sched.step).parameters.keys())
        extra = {'eta': eta} if accepts_eta else {}

        latents = init_latents.float().contiguous()
        all_latents = [latents] if debug else None
        steps_ts = sched.timesteps[t_start:]
        fused = hasattr(guide, 'noise_pred_pair') and hasattr(sched, 'fused_step')
        if fused:
            latents = self._fused_loop(guide, sched, latents, steps_ts, t_start,
                                       is_lms, extra, generator, all_latents)
        else:
            for i, t in enumerate(self.progress_bar(steps_ts)):
                t_index = t
                model_in = latents
                if is_lms:
                    t_index = t_start + i
                    sigma = sched.sigmas[t_index]
                    model_in = model_in / ((sigma**2 + 1)**0.5)
                noise_pred = guide.noise_pred(model_in, t)
                latents = sched.step(noise_pred, t_index, latents,
                                     **extra).prev_sample
                if all_latents:
                    all_latents.append(latents)

        if output_type == 'latent':
            return latents if not return_dict else StableDiffusionPipelineOutput(
                latents, [False] * latents.shape[0])
        if all_latents:
            batches = [self.decode(l, output_type) for l in all_latents]
            if isinstance(batches[0], list):
                batch_images = [im for b in batches for im in b]
            elif torch.is_tensor(batches[0]):
                batch_images = torch.cat(batches)
            else:
                batch_images = np.concatenate(batches, axis=0)
        else:
            batch_images = self.decode(latents, output_type)
        if not return_dict:
            return (batch_images, False)
        return StableDiffusionPipelineOutput(
            images=batch_images,
            nsfw_content_detected=[False for _ in batch_images])

    @torch.no_grad()
    def decode(self, latents: torch.Tensor, output_type: str = 'pil'):
        '''flex.py:112-124 for 'pil' / 'np'; 'pt' leaves the [B,3,H,W] images in [0,1] on the
        device (no host copy).'''
        if output_type == 'pt':
            image = self.vae.decode(1 / 0.18215 * latents).sample
            return (image / 2 + 0.5).clamp(0, 1)
        if output_type == 'uint8':  # the array PIL would wrap, [B,H,W,3] uint8 on the host
            image = self.vae.decode(1 / 0.18215 * latents).sample
            return _native.image_tail_u8(image).cpu().numpy()
        return self._latents_to_image(latents, output_type == 'pil')

    def _fused_loop(self, guide, sched, latents, steps_ts, t_start, is_lms,
                    extra, generator, all_latents):
        '''flex.py:262-287 with CFG + scheduler update (+ next model input) in one K4
        launch per step.  Two fp32 latent buffers ping-pong; the bf16 model input lives
        in the guide's static buffer.'''
        use_cfg = guide.guidance > 1.0
        model_in = guide.model_input_buffer(latents)
        scale0 = sched.input_scale(t_start) if is_lms else 1.0
        model_in.copy_(latents * scale0 if scale0 != 1.0 else latents)
        bufs = [latents, torch.empty_like(latents)]
        ts = [float(t) if is_lms else int(t) for t in steps_ts.tolist()]
        for i, t in enumerate(self.progress_bar(ts)):
            t_index = t_start + i if is_lms else t
            u, c = guide.noise_pred_pair(model_in, t)
            kw = dict(extra)
            if 'eta' in kw:
                kw['generator'] = generator
            cur = bufs[i % 2] if not all_latents else latents
            out = bufs[(i + 1) % 2] if not all_latents else None
            latents = sched.fused_step(u, c, guide.guidance, use_cfg, t_index,
                                       cur, out=out, scaled_out=model_in,
                                       **kw).prev_sample
            if all_latents:
                all_latents.append(latents)
        return latents
