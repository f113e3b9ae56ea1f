'''Random-init model builders for benchmarks, smoke and tests.

There are no checkpoints or tokenizer files in the image (no network), so every
benchmark in this repo runs on seeded random-init weights of the reference's
architectures: SD-v1 UNet + VAE (what `StableDiffusionPipeline.from_pretrained` yields at
/root/reference/utils.py:65-66) and CLIP ViT-L/14 (utils.py:61-63), plus a stand-in
tokenizer with the call surface /root/reference/encode/clip.py:57-63 uses.
'''
from __future__ import annotations

import hashlib
from types import SimpleNamespace
from typing import List

import torch

from .unet import UNet2DConditionModel
from .vae import AutoencoderKL


def _seeded_init(module: torch.nn.Module, seed: int, device) -> torch.nn.Module:
    '''Deterministic init on `device` without a host round trip of the fp32 weights.'''
    gen = torch.Generator(device=device).manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if p.dim() > 1:
                fan_in = p[0].numel()
                bound = (1.0 / fan_in)**0.5
                p.copy_((torch.rand(p.shape, generator=gen, device=device) * 2 - 1)
                        * bound * 1.7)
            elif name.endswith('weight'):
                p.fill_(1.0)  # norm scales
            else:
                p.copy_((torch.rand(p.shape, generator=gen, device=device) * 2 - 1)
                        * 0.02)
    return module


def build_unet(device, dtype=torch.bfloat16, seed: int = 0,
               channels_last: bool = True) -> UNet2DConditionModel:
    with torch.device(device):
        unet = UNet2DConditionModel()
    _seeded_init(unet, seed, device)
    unet = unet.to(dtype).eval().requires_grad_(False)
    if channels_last:
        unet = unet.to(memory_format=torch.channels_last)
    unet.refresh_kv_weight()
    return unet


def build_vae(device, dtype=torch.bfloat16, seed: int = 1) -> AutoencoderKL:
    with torch.device(device):
        vae = AutoencoderKL()
    _seeded_init(vae, seed, device)
    return vae.to(dtype).eval().requires_grad_(False).to(
        memory_format=torch.channels_last)


class FakeTokenizer:
    '''Deterministic stand-in for CLIPTokenizer (no vocab files on disk): BOS 49406,
    hashed word ids, EOS/pad 49407, padded / truncated to model_max_length = 77.'''
    model_max_length = 77
    bos, eos = 49406, 49407

    def _ids(self, text: str) -> List[int]:
        ids = [self.bos]
        for w in text.lower().split():
            h = int.from_bytes(hashlib.sha256(w.encode()).digest()[:4], 'little')
            ids.append(1000 + h % 48000)
        ids = ids[:self.model_max_length - 1] + [self.eos]
        return ids + [self.eos] * (self.model_max_length - len(ids))

    def __call__(self, prompt, padding='max_length', max_length=77,
                 truncation=True, return_tensors='pt'):
        if isinstance(prompt, str):
            prompt = [prompt]
        return SimpleNamespace(
            input_ids=torch.tensor([self._ids(p) for p in prompt]))


def build_clip(device, seed: int = 2, dtype=torch.float32):
    '''Random-init CLIP ViT-L/14 (text 768/12L, vision 1024/24L/patch 14, projection 768).'''
    from transformers import CLIPConfig, CLIPModel
    cfg = CLIPConfig(
        text_config=dict(hidden_size=768, intermediate_size=3072,
                         num_hidden_layers=12, num_attention_heads=12,
                         vocab_size=49408, max_position_embeddings=77),
        vision_config=dict(hidden_size=1024, intermediate_size=4096,
                           num_hidden_layers=24, num_attention_heads=16,
                           image_size=224, patch_size=14),
        projection_dim=768)
    torch.manual_seed(seed)
    clip = CLIPModel(cfg).eval().requires_grad_(False)
    return clip.to(device=device, dtype=dtype)
