'''SD-v1 UNet (`UNet2DConditionModel` of diffusers 0.3.0, the `unet` object the
reference hands to `SimpleGuide`, /root/reference/pipeline/guide.py:9,56-58) with its
16 cross-attention (`attn2`) sites served by the sm_100a kernels:

  * K2 `fd_kv_project`: to_k / to_v of the fixed 77-token context for ALL 16 layers in
    one tcgen05 GEMM, once per guide (`build_kv_cache`) instead of 32 Linears per step;
  * K3F `fd_cross_attn_fused`: to_q, softmax(Q K^T) V over that cache and to_out (+ bias) in
    one launch per site (K3 `fd_cross_attn` + two cuBLAS GEMMs when `FUSED_ATTN2` is off).

Convolutions, GroupNorm, self-attention (SDPA) and the feed-forward stay in PyTorch
(cuDNN / cuBLAS), as BASELINE.json's north_star prescribes.  Parameter names follow
diffusers 0.3.0 so a `state_dict` is interchangeable with the oracle restatement
(oracle/unet_oracle.py).  The module is used in bf16 on a B200; it has no CPU path for
attn2 (the native call raises).
'''
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native

T_VALID = 77
T_PAD = 80


@dataclass
class UNetOutput:
    sample: torch.Tensor


@dataclass
class KVCache:
    '''K2 output for a set of contexts: kv[n_ctx * T_PAD, n_kv] bf16.'''
    kv: torch.Tensor
    n_ctx: int


class FrozenConfig(dict):
    '''dict with attribute access, like diffusers' FrozenDict (flex.py:57-70, 101).'''
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


SD_V1_CONFIG = dict(in_channels=4, out_channels=4,
                    block_out_channels=(320, 640, 1280, 1280),
                    layers_per_block=2, attention_head_dim=8,
                    cross_attention_dim=768, norm_num_groups=32)


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    '''diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0).'''
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) *
                      torch.arange(half, dtype=torch.float32, device=t.device) /
                      half)
    args = t.float()[:, None] * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(x: torch.Tensor, norm: nn.GroupNorm, silu: bool, bias=None) -> torch.Tensor:
    '''GroupNorm (+ optional pre-bias, + optional SiLU) through K5 (`fd_groupnorm_act`).'''
    return _native.groupnorm_act(x, norm.weight, norm.bias, norm.num_groups, norm.eps,
                                 silu, bias)


class TembBank:
    '''All resnets' `time_emb_proj(silu(temb)) + conv1.bias` rows from one GEMM
    ([B,1280] x [sum C_out,1280]^T) instead of 22 tiny ones per forward.'''
    def __init__(self, rows: torch.Tensor, slices):
        self.rows, self.slices = rows, slices

    def take(self, resnet) -> torch.Tensor:
        a, b = self.slices[id(resnet)]
        return self.rows[:, a:b]


def _conv1x1(x: torch.Tensor, conv: nn.Conv2d) -> torch.Tensor:
    '''1x1 convolution of a channels-last tensor as one cuBLAS GEMM with the bias in its
    epilogue: the NHWC memory *is* the [N*H*W, C] row-major matrix.'''
    if not x.is_contiguous(memory_format=torch.channels_last):
        x = x.contiguous(memory_format=torch.channels_last)
    y = F.linear(x.permute(0, 2, 3, 1), conv.weight.reshape(conv.out_channels, -1),
                 conv.bias)
    return y.permute(0, 3, 1, 2)


class ResnetBlock2D(nn.Module):
    def __init__(self, cin: int, cout: int, temb: int, groups: int):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps=1e-5)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb, cout)
        self.norm2 = nn.GroupNorm(groups, cout, eps=1e-5)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def _temb_conv1_bias(self) -> torch.Tensor:
        '''time_emb_proj.bias + conv1.bias, cached until either parameter changes.'''
        a, b = self.time_emb_proj.bias, self.conv1.bias
        key = (a.data_ptr(), a._version, b.data_ptr(), b._version)
        cached = self.__dict__.get('_tb_cache')
        if cached is None or cached[0] != key:
            cached = (key, (a.detach() + b.detach()))
            self.__dict__['_tb_cache'] = cached
        return cached[1]

    def forward(self, x, temb_act, next_norm: Optional[nn.GroupNorm] = None, next_silu: bool = False):
        '''`next_norm`: the GroupNorm that consumes this block's output (a SpatialTransformer's `norm`, the next resnet's
        `norm1`): the residual add (K7) is then folded into that GroupNorm's launch and the call returns (y, act(GN(y))).'''
        # K5: GroupNorm + SiLU in one NHWC pass.  cuDNN adds a convolution bias with a separate
        # broadcast kernel, so conv1's bias rides along with the time embedding into norm2 (K5's
        # per-(n,c) bias) and conv2's bias is folded into the residual add (K7).
        h = F.conv2d(_gn(x, self.norm1, silu=True), self.conv1.weight, None, padding=1)
        if isinstance(temb_act, TembBank):
            tb = temb_act.take(self)   # slice of ONE batched projection for all 22 resnets
        else:
            tb = F.linear(temb_act, self.time_emb_proj.weight, self._temb_conv1_bias())
        h = F.conv2d(_gn(h, self.norm2, silu=True, bias=tb), self.conv2.weight, None,
                     padding=1)
        if self.conv_shortcut is not None:
            x = _conv1x1(x, self.conv_shortcut)
        if next_norm is not None:
            if FUSED_ADD_GN:
                return _native.add_groupnorm_act(x, h, self.conv2.bias, next_norm.weight, next_norm.bias, next_norm.num_groups,
                                                 next_norm.eps, next_silu)
            y = _native.add_bias_residual(x, h, self.conv2.bias)
            return y, _gn(y, next_norm, silu=next_silu)
        return _native.add_bias_residual(x, h, self.conv2.bias)


class SelfAttention(nn.Module):
    '''attn1: stays in PyTorch (SDPA).'''
    def __init__(self, dim: int, heads: int):
        super().__init__()
        self.heads = heads
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(dim, dim, bias=False)
        self.to_v = nn.Linear(dim, dim, bias=False)
        # diffusers 0.3.0: to_out = Sequential(Linear, Dropout) -> checkpoint keys `to_out.0.*`
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])

    def _qkv_weight(self) -> torch.Tensor:
        '''to_q | to_k | to_v stacked into one [3C, C] matrix (one GEMM instead of three), cached
        until any of the three parameters changes.'''
        ws = (self.to_q.weight, self.to_k.weight, self.to_v.weight)
        key = tuple((w.data_ptr(), w._version) for w in ws)
        cached = self.__dict__.get('_qkv_cache')
        if cached is None or cached[0] != key:
            cached = (key, torch.cat([w.detach() for w in ws]).contiguous())
            self.__dict__['_qkv_cache'] = cached
        return cached[1]

    def forward(self, x):
        B, N, C = x.shape
        h = self.heads
        qkv = F.linear(x, self._qkv_weight()).view(B, N, 3, h, C // h)
        q, k, v = (qkv[:, :, i].transpose(1, 2) for i in range(3))
        o = F.scaled_dot_product_attention(q, k, v)
        return self.to_out[0](o.transpose(1, 2).reshape(B, N, C))


# attn2 as ONE launch of K3F (to_q + attention + to_out + bias) or as round 1's cuBLAS to_q -> K3 ->
# cuBLAS to_out sequence.  'auto' picks per site from the same-box A/B in profiles/r02/SUMMARY.md:
# K3F wins wherever its grid fills the GPU with work per CTA that its serial phases amortise -- all
# C = 320 sites, C = 640 between 4 and 16 samples -- and loses on the deep sites (C = 1280: 2-16 CTAs
# at B = 1, each walking 2 x 20 k-chunks alone).  True / False force one path (tests, A/B).
FUSED_ATTN2 = 'auto'


def _use_fused_attn2(dim: int, n_samples: int) -> bool:
    if FUSED_ATTN2 is True or FUSED_ATTN2 is False:
        return FUSED_ATTN2
    return dim == 320 or (dim == 640 and 4 <= n_samples <= 16)


class CrossAttention(nn.Module):
    '''attn2: K/V from the K2 cache; to_q, attention and to_out in K3F (`fd_cross_attn_fused`).'''
    def __init__(self, dim: int, ctx_dim: int, heads: int):
        super().__init__()
        self.heads = heads
        self.dim = dim
        self.scale = (dim // heads)**-0.5
        self.to_q = nn.Linear(dim, dim, bias=False)
        self.to_k = nn.Linear(ctx_dim, dim, bias=False)
        self.to_v = nn.Linear(ctx_dim, dim, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(dim, dim), nn.Dropout(0.0)])
        self.k_col_off = -1  # set by UNet2DConditionModel.refresh_kv_weight
        self.v_col_off = -1

    def forward(self, x, kv: KVCache, ctx_index: torch.Tensor):
        if self.dim % 320 == 0 and self.dim // self.heads in (40, 80, 160) \
                and _use_fused_attn2(self.dim, x.shape[0]):
            if not x.is_contiguous():
                x = x.contiguous()
            lin = self.to_out[0]
            out, _ = _native.cross_attn_fused(x, self.to_q.weight, kv.kv, self.k_col_off,
                                              self.v_col_off, ctx_index, lin.weight, lin.bias,
                                              self.heads, T_VALID, T_PAD, self.scale,
                                              want_attn=False)
            return out
        q = self.to_q(x)
        if not q.is_contiguous():
            q = q.contiguous()
        o = _native.cross_attn(q, kv.kv, self.k_col_off, self.v_col_off,
                               ctx_index, self.heads, T_VALID, T_PAD,
                               self.scale)
        return self.to_out[0](o)


FF_GEGLU_FUSED = True   # K13 on / off (off = cuBLAS projection + K6); profiles/time_unet.py A/Bs the two
FUSED_ADD_GN = os.environ.get('FD_FUSED_ADD_GN', '1') != '0'   # a resnet's residual add inside the GroupNorm launch that follows it
MERGED_OUT_GEMM = os.environ.get('FD_MERGED_OUT', '1') != '0'   # ff.net[2] + residual + proj_out as one K = 5C GEMM (needs K13)


class GEGLU(nn.Module):
    def __init__(self, dim: int, inner: int):
        super().__init__()
        self.proj = nn.Linear(dim, inner * 2)

    def forward(self, x):
        w = self.proj.weight
        if (FF_GEGLU_FUSED and x.is_cuda and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
                and self.proj.bias is not None and _native.ff_geglu_supported(w.shape[1], w.shape[0] // 2)):
            # K13: the projection GEMM with value * gelu(gate) in its epilogue (the [M, 8 C] tensor is never written)
            return _native.ff_geglu(x, w, self.proj.bias)
        return _native.geglu(self.proj(x))  # cuBLAS + K6: x * gelu(gate) in one pass


class FeedForward(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.net = nn.ModuleList(
            [GEGLU(dim, dim * 4),
             nn.Dropout(0.0),
             nn.Linear(dim * 4, dim)])

    def forward(self, x):
        return self.net[2](self.net[0](x))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim: int, heads: int, ctx_dim: int):
        super().__init__()
        self.attn1 = SelfAttention(dim, heads)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, ctx_dim, heads)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)

    def forward(self, x, kv, ctx_index, merged_out=None):
        # K8: each residual add is fused with the LayerNorm that feeds the next branch
        ln = _native.add_layernorm
        if not x.is_contiguous():
            x = x.contiguous()
        _, n = ln(x, None, self.norm1.weight, self.norm1.bias, self.norm1.eps)
        x, n = ln(x, self.attn1(n), self.norm2.weight, self.norm2.bias, self.norm2.eps)
        if merged_out is None:
            x, n = ln(x, self.attn2(n, kv, ctx_index), self.norm3.weight, self.norm3.bias, self.norm3.eps)
            return self.ff(n) + x
        # Merged output GEMM (see SpatialTransformer._merged_out): the block's tail
        #     y = x3 + W2 h + b2 ,  p = Wo y + bo            (ff.net[2], residual add, proj_out)
        # is the single GEMM  p = [h | x3] [Wo W2 | Wo]^T + (Wo b2 + bo)  over K = 5C: K13 writes h = GEGLU(n3) into the left
        # 4C columns of one [M, 5C] buffer and K8 the residual stream x3 into the right C columns, so neither the second
        # feed-forward GEMM nor the add exists (2 launches fewer per block).
        w_m, b_m = merged_out
        B, N, C = x.shape
        buf = torch.empty((B * N, 5 * C), dtype=x.dtype, device=x.device)
        _, n = ln(x, self.attn2(n, kv, ctx_index), self.norm3.weight, self.norm3.bias, self.norm3.eps,
                  sum_out=buf[:, 4 * C:])
        proj = self.ff.net[0].proj
        _native.ff_geglu(n.reshape(B * N, C), proj.weight, proj.bias, out=buf[:, :4 * C])
        return F.linear(buf, w_m, b_m).view(B, N, C)


class SpatialTransformer(nn.Module):
    def __init__(self, ch: int, heads: int, ctx_dim: int, groups: int):
        super().__init__()
        self.norm = nn.GroupNorm(groups, ch, eps=1e-6)
        self.proj_in = nn.Conv2d(ch, ch, 1)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(ch, heads, ctx_dim)])
        self.proj_out = nn.Conv2d(ch, ch, 1)

    def _merged_out(self):
        '''[C, 5C] weight [Wo W2 | Wo] and bias Wo b2 + bo of the merged output GEMM (products in fp64, stored in the model
        dtype), cached until ff.net[2] or proj_out changes.  Only for a single transformer block whose GEGLU runs on K13.'''
        blk = self.transformer_blocks[0]
        lin2, po = blk.ff.net[2], self.proj_out
        ps = (lin2.weight, lin2.bias, po.weight, po.bias)
        key = tuple((p.data_ptr(), p._version) for p in ps)
        cached = self.__dict__.get('_merged_cache')
        if cached is None or cached[0] != key:
            C = po.out_channels
            wo = po.weight.detach().reshape(C, C).double()   # float64: independent of the TF32 matmul switch
            w_m = torch.cat([wo @ lin2.weight.detach().double(), wo], dim=1).to(po.weight.dtype).contiguous()
            b_m = (wo @ lin2.bias.detach().double() + po.bias.detach().double()).to(po.weight.dtype)
            cached = (key, (w_m, b_m), ps)
            self.__dict__['_merged_cache'] = cached
        return cached[1]

    def forward(self, x, kv, ctx_index, normed=None):
        '''`normed`: GroupNorm(x) when the producer of x already computed it (ResnetBlock2D(next_norm=self.norm)).'''
        B, C, H, W = x.shape
        # proj_in / proj_out are 1x1 convolutions: run them as GEMMs on the token matrix
        h = (normed if normed is not None else _gn(x, self.norm, silu=False)).permute(0, 2, 3, 1).reshape(B, H * W, C)
        h = F.linear(h, self.proj_in.weight.reshape(C, C), self.proj_in.bias)
        blk0 = self.transformer_blocks[0]
        if (MERGED_OUT_GEMM and FF_GEGLU_FUSED and len(self.transformer_blocks) == 1 and h.is_cuda
                and h.dtype == torch.bfloat16 and _native.ff_geglu_supported(C, 4 * C)
                and blk0.ff.net[0].proj.bias is not None and blk0.ff.net[2].bias is not None
                and self.proj_out.bias is not None):
            h = blk0(h, kv, ctx_index, merged_out=self._merged_out())   # already proj_out's output
        else:
            for blk in self.transformer_blocks:
                h = blk(h, kv, ctx_index)
            h = F.linear(h, self.proj_out.weight.reshape(C, C), self.proj_out.bias)
        return h.reshape(B, H, W, C).permute(0, 3, 1, 2) + x


# development A/B switch: FD_ATEN_GLUE=1 sends the convolution-bias adds, the skip concats and the nearest upsamples back to ATen
NATIVE_GLUE = os.environ.get('FD_ATEN_GLUE', '0') != '1'


def _conv_bias(x, conv: nn.Conv2d, stride: int = 1):
    '''3x3 convolution with its bias added by K7 (h = None) instead of ATen's broadcasting add kernel.'''
    if (not NATIVE_GLUE or conv.bias is None or not x.is_cuda or x.dtype != torch.bfloat16 or conv.out_channels % 8
            or conv.bias.dtype != torch.bfloat16):
        return conv(x)
    h = F.conv2d(x, conv.weight, None, stride=stride, padding=conv.padding)
    return _native.add_bias_residual(h, None, conv.bias, inplace=True)


def _cat_skip(x, skip):
    '''torch.cat([x, skip], dim=1) through K15 on channels-last bf16 activations.'''
    if (NATIVE_GLUE and x.is_cuda and x.dtype == torch.bfloat16 and skip.dtype == torch.bfloat16 and x.shape[1] % 8 == 0
            and skip.shape[1] % 8 == 0):
        return _native.concat_channels(x, skip)
    return torch.cat([x, skip], dim=1)


class Downsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return _conv_bias(x, self.conv, stride=2)


class Upsample2D(nn.Module):
    def __init__(self, ch: int):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        if NATIVE_GLUE and x.is_cuda and x.dtype == torch.bfloat16 and x.shape[1] % 8 == 0:
            return _conv_bias(_native.upsample_nearest2x(x), self.conv)   # K15
        return _conv_bias(F.interpolate(x, scale_factor=2.0, mode='nearest'), self.conv)


class DownBlock(nn.Module):
    def __init__(self, cin, cout, temb, groups, n_layers, heads, ctx_dim,
                 cross: bool, downsample: bool):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(cin if i == 0 else cout, cout, temb, groups)
            for i in range(n_layers)
        ])
        self.attentions = nn.ModuleList([
            SpatialTransformer(cout, heads, ctx_dim, groups)
            for _ in range(n_layers)
        ]) if cross else None
        self.downsamplers = nn.ModuleList([Downsample2D(cout)
                                           ]) if downsample else None

    def forward(self, x, temb, kv, ctx_index):
        outs = []
        for i, res in enumerate(self.resnets):
            if self.attentions is not None:
                x, z = res(x, temb, next_norm=self.attentions[i].norm)
                x = self.attentions[i](x, kv, ctx_index, normed=z)
            else:
                x = res(x, temb)
            outs.append(x)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
            outs.append(x)
        return x, outs


class MidBlock(nn.Module):
    def __init__(self, ch, temb, groups, heads, ctx_dim):
        super().__init__()
        self.resnets = nn.ModuleList([
            ResnetBlock2D(ch, ch, temb, groups),
            ResnetBlock2D(ch, ch, temb, groups)
        ])
        self.attentions = nn.ModuleList(
            [SpatialTransformer(ch, heads, ctx_dim, groups)])

    def forward(self, x, temb, kv, ctx_index):
        x, z = self.resnets[0](x, temb, next_norm=self.attentions[0].norm)
        x = self.attentions[0](x, kv, ctx_index, normed=z)
        return self.resnets[1](x, temb)


class UpBlock(nn.Module):
    def __init__(self, cin, cout, cprev, temb, groups, n_layers, heads,
                 ctx_dim, cross: bool, upsample: bool):
        super().__init__()
        res = []
        for i in range(n_layers):
            skip = cin if i == n_layers - 1 else cout
            rin = cprev if i == 0 else cout
            res.append(ResnetBlock2D(rin + skip, cout, temb, groups))
        self.resnets = nn.ModuleList(res)
        self.attentions = nn.ModuleList([
            SpatialTransformer(cout, heads, ctx_dim, groups)
            for _ in range(n_layers)
        ]) if cross else None
        self.upsamplers = nn.ModuleList([Upsample2D(cout)
                                         ]) if upsample else None

    def forward(self, x, skips: List[torch.Tensor], temb, kv, ctx_index):
        for i, res in enumerate(self.resnets):
            if self.attentions is not None:
                x, z = res(_cat_skip(x, skips.pop()), temb, next_norm=self.attentions[i].norm)
                x = self.attentions[i](x, kv, ctx_index, normed=z)
            else:
                x = res(_cat_skip(x, skips.pop()), temb)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class UNet2DConditionModel(nn.Module):
    '''SD-v1 conditional UNet; see module docstring.'''
    def __init__(self, **overrides):
        super().__init__()
        cfg = dict(SD_V1_CONFIG)
        cfg.update(overrides)
        self.config = FrozenConfig(cfg)
        self.in_channels = cfg['in_channels']  # flex.py:224
        chs: Tuple[int, ...] = tuple(cfg['block_out_channels'])
        heads, ctx, groups = (cfg['attention_head_dim'],
                              cfg['cross_attention_dim'],
                              cfg['norm_num_groups'])
        n_layers = cfg['layers_per_block']
        temb = chs[0] * 4
        self.time_embedding = nn.ModuleDict(
            dict(linear_1=nn.Linear(chs[0], temb),
                 linear_2=nn.Linear(temb, temb)))
        self.conv_in = nn.Conv2d(cfg['in_channels'], chs[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        cout = chs[0]
        for i, ch in enumerate(chs):
            cin, cout = cout, ch
            last = i == len(chs) - 1
            self.down_blocks.append(
                DownBlock(cin, cout, temb, groups, n_layers, heads, ctx,
                          cross=not last, downsample=not last))
        self.mid_block = MidBlock(chs[-1], temb, groups, heads, ctx)
        self.up_blocks = nn.ModuleList()
        rev = tuple(reversed(chs))
        cout = rev[0]
        for i, ch in enumerate(rev):
            cprev, cout = cout, ch
            cin = rev[min(i + 1, len(chs) - 1)]
            last = i == len(chs) - 1
            self.up_blocks.append(
                UpBlock(cin, cout, cprev, temb, groups, n_layers + 1, heads,
                        ctx, cross=i != 0, upsample=not last))
        self.conv_norm_out = nn.GroupNorm(groups, chs[0], eps=1e-5)
        self.conv_out = nn.Conv2d(chs[0], cfg['out_channels'], 3, padding=1)
        self._kv_weight: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------ K2 plumbing
    def cross_attentions(self) -> List[CrossAttention]:
        '''The 16 attn2 modules in execution order.'''
        return [m for m in self.modules() if isinstance(m, CrossAttention)]

    @torch.no_grad()
    def refresh_kv_weight(self) -> torch.Tensor:
        '''Pack every to_k / to_v weight into one [sum 2*C_l, 768] bf16 matrix (the B operand
        of K2) and record each layer's column offsets in the cache rows.'''
        rows, off = [], 0
        self.__dict__['_derived_at'] = self._derived_key()
        for m in self.cross_attentions():
            m.k_col_off, m.v_col_off = off, off + m.dim
            rows += [m.to_k.weight, m.to_v.weight]
            off += 2 * m.dim
        self._kv_weight = torch.cat(rows).to(torch.bfloat16).contiguous()
        # packed time-embedding projections of every resnet (+ their conv1 bias)
        resnets = [m for m in self.modules() if isinstance(m, ResnetBlock2D)]
        self._tb_slices, off = {}, 0
        for m in resnets:
            c = m.time_emb_proj.out_features
            self._tb_slices[id(m)] = (off, off + c)
            off += c
        self._tb_weight = torch.cat([m.time_emb_proj.weight for m in resnets]).contiguous()
        self._tb_bias = torch.cat([m.time_emb_proj.bias + m.conv1.bias
                                   for m in resnets]).contiguous()
        return self._kv_weight

    @torch.no_grad()
    def build_kv_cache(self, contexts: torch.Tensor) -> KVCache:
        '''contexts [n_ctx, 77, 768] (any float dtype) -> K/V of all 16 layers, one GEMM.'''
        if self._kv_weight is None or self._kv_weight.device != contexts.device \
                or self.__dict__.get('_derived_at') != self._derived_key():
            self.invalidate_derived()
            self.refresh_kv_weight()
        n_ctx, t, d = contexts.shape
        if t != T_VALID:
            raise ValueError(f'context must have {T_VALID} tokens, got {t}')
        ctx = torch.zeros((n_ctx, T_PAD, d), dtype=torch.bfloat16,
                          device=contexts.device)
        ctx[:, :t] = contexts.to(torch.bfloat16)
        kv = _native.kv_project(ctx.view(n_ctx * T_PAD, d), self._kv_weight)
        return KVCache(kv=kv, n_ctx=n_ctx)

    def _derived_key(self):
        '''(data_ptr, version) of every parameter the packed K/V and time-embedding weights are
        built from: in-place edits and reloads both change it.'''
        ps = []
        for m in self.cross_attentions():
            ps += [m.to_k.weight, m.to_v.weight]
        for m in self.modules():
            if isinstance(m, ResnetBlock2D):
                ps += [m.time_emb_proj.weight, m.time_emb_proj.bias, m.conv1.bias]
            elif isinstance(m, SpatialTransformer):   # the merged output GEMM's weight is derived from these
                lin2 = m.transformer_blocks[0].ff.net[2]
                ps += [lin2.weight, lin2.bias, m.proj_out.weight, m.proj_out.bias]
        return tuple((p.data_ptr(), p._version) for p in ps)

    def invalidate_derived(self):
        '''Drop everything computed from the parameters: packed weights and captured graphs
        (a graph has the packed weights and its K/V buffer baked in by address).'''
        self._kv_weight = None
        self._tb_weight = None
        for r in self.__dict__.get('_graph_runners', {}).values():
            r.owner = None
        self.__dict__['_graph_runners'] = {}

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self.invalidate_derived()
        return out

    def set_attention_slice(self, slice_size):  # flex.py:102; memory knob only
        self._attention_slice = slice_size

    def graph_runner(self, latent_shape, n_ctx: int, repeat: int) -> 'UNetGraphRunner':
        '''CUDA-graph runner for one (latent shape, #contexts, #copies of the latents) signature
        (CFG evaluates the latents twice, a composite guide 2 + #entities times), shared by every
        guide with that signature (capturing costs ~100 ms, replaying ~nothing).'''
        key = (tuple(latent_shape), n_ctx, int(repeat))
        runners = self.__dict__.setdefault('_graph_runners', {})
        if key not in runners:
            runners[key] = UNetGraphRunner(self, latent_shape, n_ctx, int(repeat))
        return runners[key]

    # ------------------------------------------------------------------ forward
    def forward(self,
                sample: torch.Tensor,
                timestep,
                encoder_hidden_states: Optional[torch.Tensor] = None,
                *,
                kv_cache: Optional[KVCache] = None,
                ctx_index: Optional[torch.Tensor] = None,
                temb_sin: Optional[torch.Tensor] = None) -> UNetOutput:
        dtype = self.conv_in.weight.dtype
        B = sample.shape[0]
        if kv_cache is None:
            # reference call form unet(latents, t, encoder_hidden_states=ctx): project now
            if encoder_hidden_states is None:
                raise ValueError('need encoder_hidden_states or kv_cache')
            kv_cache = self.build_kv_cache(encoder_hidden_states)
            ctx_index = torch.arange(B, dtype=torch.int32, device=sample.device)
        if temb_sin is None:
            if not torch.is_tensor(timestep):
                timestep = torch.tensor([timestep], device=sample.device)
            timestep = timestep.reshape(-1).to(sample.device)
            temb_sin = timestep_embedding(timestep, self.conv_in.out_channels)
        temb_sin = temb_sin.to(dtype).expand(B, -1)
        emb = self.time_embedding['linear_2'](F.silu(
            self.time_embedding['linear_1'](temb_sin)))
        emb = F.silu(emb)  # every resnet applies silu(temb) first
        if getattr(self, '_tb_weight', None) is not None and \
                self._tb_weight.device == emb.device and self._tb_weight.dtype == emb.dtype:
            emb = TembBank(F.linear(emb, self._tb_weight, self._tb_bias), self._tb_slices)

        x = _conv_bias(sample.to(dtype), self.conv_in)
        skips = [x]
        for blk in self.down_blocks:
            x, outs = blk(x, emb, kv_cache, ctx_index)
            skips += outs
        x = self.mid_block(x, emb, kv_cache, ctx_index)
        for blk in self.up_blocks:
            x = blk(x, skips, emb, kv_cache, ctx_index)
        x = self.conv_out(_gn(x, self.conv_norm_out, silu=True))
        return UNetOutput(sample=x)


class UNetGraphRunner:
    '''One captured UNet forward (B=1 is launch-bound: ~700 small kernels per forward).

    Static buffers: bf16 model input [B,4,h,w] (K4 writes the next step's input straight into
    it), the sinusoidal timestep row, the K2 cache and the sample->context index.  A guide
    "owns" the runner while its cache is loaded; switching guides costs one D2D copy.'''
    def __init__(self, unet: UNet2DConditionModel, latent_shape, n_ctx: int, repeat: int):
        dev = unet.conv_in.weight.device
        dt = unet.conv_in.weight.dtype
        self.unet, self.repeat = unet, repeat
        self.static_in = torch.zeros(tuple(latent_shape), dtype=dt, device=dev)
        self.static_temb = torch.zeros((1, unet.conv_in.out_channels),
                                       dtype=torch.float32, device=dev)
        n_kv = unet._kv_weight.shape[0]
        self.kv = KVCache(torch.zeros((n_ctx * T_PAD, n_kv), dtype=torch.bfloat16,
                                      device=dev), n_ctx)
        B = latent_shape[0]
        self.ctx_index = torch.zeros((B * repeat,), dtype=torch.int32, device=dev)
        self.owner = None
        self.graph = None
        self.static_out = None
        self.native_launches = 0

    def _forward(self):
        x = torch.cat([self.static_in] * self.repeat) if self.repeat > 1 else self.static_in
        return self.unet(x, None, kv_cache=self.kv, ctx_index=self.ctx_index,
                         temb_sin=self.static_temb).sample

    def load(self, owner, kv: KVCache, ctx_index: torch.Tensor):
        # keyed on the guide AND on the cache tensor it hands over: a guide rebuilds its cache
        # when its embeddings / CFG layout change
        key = (kv.kv.data_ptr(), kv.kv._version, ctx_index.data_ptr())
        if self.owner is not owner or self.__dict__.get('_loaded') != key:
            self.kv.kv.copy_(kv.kv)
            self.ctx_index.copy_(ctx_index)
            self.owner = owner
            self._loaded = key

    @torch.no_grad()
    def run(self, temb: torch.Tensor) -> torch.Tensor:
        if self.graph is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):  # cuDNN / cuBLAS heuristics settle off-graph
                    self._forward()
            torch.cuda.current_stream().wait_stream(side)
            before = _native.LAUNCHES
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self.static_out = self._forward()
            self.native_launches = _native.LAUNCHES - before
            _native.count_launch(-self.native_launches)  # capturing launches nothing
        self.static_temb.copy_(temb)
        self.graph.replay()
        _native.count_launch(self.native_launches)
        return self.static_out
