'''ORACLE (test infrastructure, never on the product path).

fp32 restatement of the third-party arithmetic the reference's denoising loop calls
into -- diffusers==0.3.0 (requirements.txt:1), NOT vendored in /root/reference and not
installable here (no network, not in the wheelhouse):

  * `UNet2DConditionModel.forward` with the SD-v1 config, incl. `CrossAttention`
    (to_q / to_k / to_v / softmax(QK^T d^-1/2)V / to_out) -- call site
    /root/reference/pipeline/guide.py:56-58
  * `AutoencoderKL.encode / .decode` -- call sites pipeline/flex.py:118,189-191

written as plain functions over a `state_dict` (different structure from the product's
nn.Module in flexdiffuse_b200/unet.py, same parameter names so weights are shared).
Every cross-attention here RECOMPUTES to_k(context) / to_v(context) and materialises the
score matrix, exactly what the reference's stack does at every step.

PARITY UNPINNED: the reference has no tests / golden vectors at the diffusers boundary
and diffusers 0.3.0 itself cannot be run here, so this restatement (from the published
architecture) is this repo's own oracle for K2/K3 and the loop; see DESIGN.md.
'''
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]

BLOCK_OUT = (320, 640, 1280, 1280)
HEADS = 8
GROUPS = 32


def _sub(sd: SD, prefix: str) -> SD:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def _lin(sd: SD, name: str, x):
    return F.linear(x, sd[name + '.weight'], sd.get(name + '.bias'))


def _conv(sd: SD, name: str, x, stride=1, padding=1):
    return F.conv2d(x, sd[name + '.weight'], sd[name + '.bias'], stride=stride,
                    padding=padding)


def _gn(sd: SD, name: str, x, eps):
    return F.group_norm(x, GROUPS, sd[name + '.weight'], sd[name + '.bias'], eps)


def _ln(sd: SD, name: str, x):
    return F.layer_norm(x, x.shape[-1:], sd[name + '.weight'],
                        sd[name + '.bias'], 1e-5)


def sinusoid(t: torch.Tensor, dim: int) -> torch.Tensor:
    half = dim // 2
    e = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32,
                                                  device=t.device) / half)
    a = t[:, None].float() * e[None, :]
    return torch.cat([torch.sin(a), torch.cos(a)], -1).roll(half, -1)  # flip_sin_to_cos


def resnet(sd: SD, x, temb, eps=1e-5):
    h = _conv(sd, 'conv1', F.silu(_gn(sd, 'norm1', x, eps)))
    if temb is not None:
        h = h + _lin(sd, 'time_emb_proj', F.silu(temb))[:, :, None, None]
    h = _conv(sd, 'conv2', F.silu(_gn(sd, 'norm2', h, eps)))
    if 'conv_shortcut.weight' in sd:
        x = _conv(sd, 'conv_shortcut', x, padding=0)
    return x + h


def attention(sd: SD, x, context):
    '''diffusers CrossAttention: context=None -> self-attention.'''
    ctx = x if context is None else context
    B, N, C = x.shape
    d = C // HEADS

    def heads(t):
        return t.view(B, -1, HEADS, d).permute(0, 2, 1, 3)

    q, k, v = (heads(_lin(sd, 'to_q', x)), heads(_lin(sd, 'to_k', ctx)),
               heads(_lin(sd, 'to_v', ctx)))
    scores = torch.einsum('bhid,bhjd->bhij', q, k) * d**-0.5
    o = torch.einsum('bhij,bhjd->bhid', scores.softmax(dim=-1), v)
    return _lin(sd, 'to_out.0', o.permute(0, 2, 1, 3).reshape(B, N, C))


def transformer(sd: SD, x, context):
    B, C, H, W = x.shape
    h = _conv(sd, 'proj_in', _gn(sd, 'norm', x, 1e-6), padding=0)
    h = h.permute(0, 2, 3, 1).reshape(B, H * W, C)
    b = _sub(sd, 'transformer_blocks.0.')
    h = attention(_sub(b, 'attn1.'), _ln(b, 'norm1', h), None) + h
    h = attention(_sub(b, 'attn2.'), _ln(b, 'norm2', h), context) + h
    g = _lin(b, 'ff.net.0.proj', _ln(b, 'norm3', h))
    a, gate = g.chunk(2, dim=-1)
    h = _lin(b, 'ff.net.2', a * F.gelu(gate)) + h
    h = h.reshape(B, H, W, C).permute(0, 3, 1, 2)
    return _conv(sd, 'proj_out', h, padding=0) + x


def unet_forward(sd: SD, sample, timestep, context):
    '''UNet2DConditionModel(sample, timestep, encoder_hidden_states=context).sample'''
    B = sample.shape[0]
    t = torch.as_tensor(timestep, device=sample.device).reshape(-1).expand(B)
    temb = _lin(sd, 'time_embedding.linear_2',
                F.silu(_lin(sd, 'time_embedding.linear_1',
                            sinusoid(t, BLOCK_OUT[0]).to(sample.dtype))))
    x = _conv(sd, 'conv_in', sample)
    skips: List[torch.Tensor] = [x]
    for i in range(4):
        blk = _sub(sd, f'down_blocks.{i}.')
        for j in range(2):
            x = resnet(_sub(blk, f'resnets.{j}.'), x, temb)
            if i < 3:
                x = transformer(_sub(blk, f'attentions.{j}.'), x, context)
            skips.append(x)
        if i < 3:
            x = _conv(blk, 'downsamplers.0.conv', x, stride=2)
            skips.append(x)
    mid = _sub(sd, 'mid_block.')
    x = resnet(_sub(mid, 'resnets.0.'), x, temb)
    x = transformer(_sub(mid, 'attentions.0.'), x, context)
    x = resnet(_sub(mid, 'resnets.1.'), x, temb)
    for i in range(4):
        blk = _sub(sd, f'up_blocks.{i}.')
        for j in range(3):
            x = resnet(_sub(blk, f'resnets.{j}.'),
                       torch.cat([x, skips.pop()], dim=1), temb)
            if i > 0:
                x = transformer(_sub(blk, f'attentions.{j}.'), x, context)
        if i < 3:
            x = F.interpolate(x, scale_factor=2.0, mode='nearest')
            x = _conv(blk, 'upsamplers.0.conv', x)
    return _conv(sd, 'conv_out', F.silu(_gn(sd, 'conv_norm_out', x, 1e-5)))


# ------------------------------------------------------------------------- VAE
def _vae_attn(sd: SD, x):
    B, C, H, W = x.shape
    h = _gn(sd, 'group_norm', x, 1e-6).view(B, C, H * W).transpose(1, 2)
    q, k, v = _lin(sd, 'query', h), _lin(sd, 'key', h), _lin(sd, 'value', h)
    w = torch.softmax(q @ k.transpose(1, 2) * C**-0.5, dim=-1)
    o = _lin(sd, 'proj_attn', w @ v).transpose(1, 2).reshape(B, C, H, W)
    return o + x


def _vae_mid(sd: SD, x):
    x = resnet(_sub(sd, 'resnets.0.'), x, None, 1e-6)
    x = _vae_attn(_sub(sd, 'attentions.0.'), x)
    return resnet(_sub(sd, 'resnets.1.'), x, None, 1e-6)


def vae_decode(sd: SD, z):
    z = _conv(sd, 'post_quant_conv', z, padding=0)
    d = _sub(sd, 'decoder.')
    x = _vae_mid(_sub(d, 'mid_block.'), _conv(d, 'conv_in', z))
    for i in range(4):
        blk = _sub(d, f'up_blocks.{i}.')
        for j in range(3):
            x = resnet(_sub(blk, f'resnets.{j}.'), x, None, 1e-6)
        if i < 3:
            x = F.interpolate(x, scale_factor=2.0, mode='nearest')
            x = _conv(blk, 'upsamplers.0.conv', x)
    return _conv(d, 'conv_out', F.silu(_gn(d, 'conv_norm_out', x, 1e-6)))


def vae_encode_moments(sd: SD, img):
    e = _sub(sd, 'encoder.')
    x = _conv(e, 'conv_in', img)
    for i in range(4):
        blk = _sub(e, f'down_blocks.{i}.')
        for j in range(2):
            x = resnet(_sub(blk, f'resnets.{j}.'), x, None, 1e-6)
        if i < 3:
            x = _conv(blk, 'downsamplers.0.conv', F.pad(x, (0, 1, 0, 1)),
                      stride=2, padding=0)
    x = _vae_mid(_sub(e, 'mid_block.'), x)
    x = _conv(e, 'conv_out', F.silu(_gn(e, 'conv_norm_out', x, 1e-6)))
    return _conv(sd, 'quant_conv', x, padding=0)
