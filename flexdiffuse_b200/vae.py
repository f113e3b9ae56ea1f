'''SD-v1 `AutoencoderKL` (diffusers 0.3.0), the `vae` the reference pipeline decodes
with at /root/reference/pipeline/flex.py:112-124 and encodes the init image with at
flex.py:189-192.  Convolutions and the mid-block attention stay in PyTorch (cuDNN / SDPA) --
BASELINE.json's north_star keeps the VAE out of the four hand-written subsystems; it sits
either side of the hot loop.  Its GroupNorm+SiLU pairs go through K5.  Parameter names follow diffusers so a `state_dict` is shared
with the oracle restatement.
'''
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native


def _gn(x: torch.Tensor, norm: nn.GroupNorm, silu: bool) -> torch.Tensor:
    '''GroupNorm (+ SiLU) through K5 (`fd_groupnorm_act`).  CUDA bf16 only, like the UNet: there
    is no PyTorch / CPU fallback (the native call raises).'''
    return _native.groupnorm_act(x, norm.weight, norm.bias, norm.num_groups, norm.eps, silu)


class VaeResnet(nn.Module):
    def __init__(self, cin, cout, groups=32):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(_gn(x, self.norm1, True))
        h = self.conv2(_gn(h, self.norm2, True))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class AttentionBlock(nn.Module):
    '''single-head spatial self-attention of the VAE mid block.'''
    def __init__(self, ch, groups=32):
        super().__init__()
        self.group_norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.query = nn.Linear(ch, ch)
        self.key = nn.Linear(ch, ch)
        self.value = nn.Linear(ch, ch)
        self.proj_attn = nn.Linear(ch, ch)

    def forward(self, x):
        B, C, H, W = x.shape
        h = _gn(x, self.group_norm, False).permute(0, 2, 3, 1).reshape(B, H * W, C)
        q, k, v = self.query(h), self.key(h), self.value(h)
        o = F.scaled_dot_product_attention(q[:, None], k[:, None], v[:, None])[:, 0]
        o = self.proj_attn(o).reshape(B, H, W, C).permute(0, 3, 1, 2)
        return o + x


class VaeMid(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.resnets = nn.ModuleList([VaeResnet(ch, ch), VaeResnet(ch, ch)])
        self.attentions = nn.ModuleList([AttentionBlock(ch)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class _Down(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))


class _Up(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        if (os.environ.get('FD_ATEN_GLUE', '0') != '1' and x.is_cuda and x.dtype == torch.bfloat16
                and x.shape[1] % 8 == 0):
            return self.conv(_native.upsample_nearest2x(x))   # K15 (ATen's nhwc nearest kernel moves < 1 TB/s)
        return self.conv(F.interpolate(x, scale_factor=2.0, mode='nearest'))


class _EncBlock(nn.Module):
    def __init__(self, cin, cout, down):
        super().__init__()
        self.resnets = nn.ModuleList(
            [VaeResnet(cin, cout), VaeResnet(cout, cout)])
        self.downsamplers = nn.ModuleList([_Down(cout)]) if down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.downsamplers[0](x) if self.downsamplers else x


class _DecBlock(nn.Module):
    def __init__(self, cin, cout, up):
        super().__init__()
        self.resnets = nn.ModuleList([
            VaeResnet(cin if i == 0 else cout, cout) for i in range(3)
        ])
        self.upsamplers = nn.ModuleList([_Up(cout)]) if up else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.upsamplers[0](x) if self.upsamplers else x


class Encoder(nn.Module):
    def __init__(self, chs=(128, 256, 512, 512), latent=4):
        super().__init__()
        self.conv_in = nn.Conv2d(3, chs[0], 3, padding=1)
        blocks, c = [], chs[0]
        for i, ch in enumerate(chs):
            blocks.append(_EncBlock(c, ch, down=i != len(chs) - 1))
            c = ch
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = VaeMid(c)
        self.conv_norm_out = nn.GroupNorm(32, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 2 * latent, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(_gn(x, self.conv_norm_out, True))


class Decoder(nn.Module):
    def __init__(self, chs=(128, 256, 512, 512), latent=4):
        super().__init__()
        rev = tuple(reversed(chs))
        self.conv_in = nn.Conv2d(latent, rev[0], 3, padding=1)
        self.mid_block = VaeMid(rev[0])
        blocks, c = [], rev[0]
        for i, ch in enumerate(rev):
            blocks.append(_DecBlock(c, ch, up=i != len(chs) - 1))
            c = ch
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(32, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 3, 3, padding=1)

    def forward(self, z):
        x = self.mid_block(self.conv_in(z))
        for b in self.up_blocks:
            x = b(x)
        return self.conv_out(_gn(x, self.conv_norm_out, True))


class DiagonalGaussianDistribution:
    def __init__(self, moments: torch.Tensor):
        self.mean, logvar = moments.chunk(2, dim=1)
        self.logvar = logvar.clamp(-30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def sample(self, generator: Optional[torch.Generator] = None):
        noise = torch.randn(self.mean.shape, generator=generator,
                            device=self.mean.device, dtype=torch.float32)
        return self.mean + self.std * noise.to(self.mean.dtype)

    def mode(self):
        return self.mean


@dataclass
class EncoderOutput:
    latent_dist: DiagonalGaussianDistribution


@dataclass
class DecoderOutput:
    sample: torch.Tensor


class AutoencoderKL(nn.Module):
    def __init__(self):
        super().__init__()
        self.encoder = Encoder()
        self.decoder = Decoder()
        self.quant_conv = nn.Conv2d(8, 8, 1)
        self.post_quant_conv = nn.Conv2d(4, 4, 1)

    def encode(self, x: torch.Tensor) -> EncoderOutput:
        dt = self.quant_conv.weight.dtype
        return EncoderOutput(
            DiagonalGaussianDistribution(
                self.quant_conv(self.encoder(x.to(dt))).float()))

    def decode(self, z: torch.Tensor) -> DecoderOutput:
        dt = self.quant_conv.weight.dtype
        return DecoderOutput(self.decoder(self.post_quant_conv(z.to(dt))))
