'''Composite (regional) noise guide -- API mirror of /root/reference/composition/guide.py.

`CompositeGuide` evaluates the UNet once per step on (uncond, background, entity_0..E-1)
contexts over the SAME latents, lerps each entity's noise prediction into the background
prediction inside its rectangle (composition/guide.py:66-87) and applies classifier-free
guidance (:89-93).  Here:

  * the K/V projections of all 2+E contexts are cached once by K2, attention runs in K3,
    the whole forward can be replayed from a CUDA graph;
  * the rectangular lerps run in K9 (`fd_composite_eps`), CFG (+ the scheduler update, when
    driven by flexdiffuse_b200's FlexPipeline) in K4.

The reference concatenates `[latents] * n_contexts` against `[uncond] * batch + [bg, entities]`
(:45-62), which only has matching batch sizes for batch_size == 1; other values raise here
instead of failing inside the UNet.  The style tween computed in the reference's `noise_pred`
(:112-121) is dead code there ("TODO") and is not reproduced.
'''
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .. import _native
from ..encode.clip import CLIPEncoder
from ..pipeline.guide import GuideBase
from ..unet import timestep_embedding
from .embeds import encode_schema
from .schema import Schema

MIN_DIM = 64  # 64 * 8 = 512 px, where SD generates best


class CompositeGuide(GuideBase):
    def __init__(self, encoder: CLIPEncoder, unet, guidance: float, schema: Schema,
                 steps: int, batch_size: int = 1, use_cuda_graph: bool = False):
        GuideBase.__init__(self, encoder, unet, guidance, steps)
        if batch_size != 1:
            raise ValueError('CompositeGuide: the reference only forms consistent UNet batches '
                             f'for batch_size == 1 (got {batch_size})')
        self.schema = schema
        self.embeds = encode_schema(schema, encoder)
        self.batch_size = batch_size
        self.classifier_free_guidance = self.guidance > 1.0
        conds = [self.embeds.background_embed, *[e.embed for e in self.embeds.entities]]
        if self.classifier_free_guidance:
            self.embed_tensor = torch.cat([self.uncond_embeds] * batch_size + conds)
        else:
            self.embed_tensor = torch.cat(conds)
        self.use_cuda_graph = use_cuda_graph
        self._kv = None
        self._ctx_index = None
        self._plain_in = None
        self._temb_cache = {}
        self._boxes = []
        for e in self.embeds.entities:
            b = _native.EntityBox()
            b.ox, b.oy = int(e.offset_blocks[0]), int(e.offset_blocks[1])
            b.sx, b.sy = int(e.size_blocks[0]), int(e.size_blocks[1])
            b.blend = float(e.blend)
            self._boxes.append(b)

    # ------------------------------------------------------------------ plumbing (as SimpleGuide)
    def _contexts(self) -> torch.Tensor:
        '''[uncond, background, entities...]; without CFG the uncond slot repeats the background
        so K9's sample layout stays the same.'''
        if self.classifier_free_guidance:
            return self.embed_tensor
        return torch.cat([self.embed_tensor[:1], self.embed_tensor])

    def _ensure_cache(self):
        key = (self.embed_tensor.data_ptr(), self.embed_tensor._version, self.guidance > 1.0)
        if self._kv is not None and self.__dict__.get('_kv_key') != key:
            self._kv = None
        if self._kv is None:
            self._kv_key = key
            if not hasattr(self.unet, 'build_kv_cache'):
                raise _native.NativeError('CompositeGuide needs flexdiffuse_b200.unet.'
                                          'UNet2DConditionModel (K2/K3 cross-attention)')
            ctx = self._contexts()
            with torch.no_grad():
                self._kv = self.unet.build_kv_cache(ctx)
            self._ctx_index = torch.arange(ctx.shape[0], dtype=torch.int32, device=ctx.device)

    def _temb(self, step, device):
        key = float(step)
        t = self._temb_cache.get(key)
        if t is None:
            t = timestep_embedding(torch.tensor([key]), self.unet.conv_in.out_channels).to(device)
            self._temb_cache[key] = t
        return t

    def model_input_buffer(self, latents: torch.Tensor) -> torch.Tensor:
        self._ensure_cache()
        if self.use_cuda_graph:
            return self.unet.graph_runner(latents.shape, self._kv.n_ctx,
                                          self._kv.n_ctx).static_in
        dt = self.unet.conv_in.weight.dtype
        if self._plain_in is None or self._plain_in.shape != latents.shape:
            self._plain_in = torch.empty(latents.shape, dtype=dt, device=latents.device)
        return self._plain_in

    @torch.no_grad()
    def noise_pred_pair(self, latents: torch.Tensor,
                        step) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        '''(eps_uncond, composite eps_cond), both fp32 [1,4,h,w]; eps_uncond None without CFG.'''
        self._ensure_cache()
        if latents.shape[0] != 1:
            raise ValueError('CompositeGuide works on one latent at a time')
        buf = self.model_input_buffer(latents)
        if latents.data_ptr() != buf.data_ptr():
            buf.copy_(latents)
        temb = self._temb(step, latents.device)
        n = self._kv.n_ctx
        if self.use_cuda_graph:
            runner = self.unet.graph_runner(latents.shape, n, n)
            runner.load(self, self._kv, self._ctx_index)
            eps = runner.run(temb)
        else:
            eps = self.unet(torch.cat([buf] * n), None, kv_cache=self._kv,
                            ctx_index=self._ctx_index, temb_sin=temb).sample
        u, c = _native.composite_eps(eps.contiguous(), self._boxes)
        return (u if self.classifier_free_guidance else None), c

    def noise_pred(self, latents: torch.Tensor, step) -> torch.FloatTensor:
        '''Reference form (composition/guide.py:97-139): the composed, CFG-combined prediction.'''
        u, c = self.noise_pred_pair(latents, step)
        if u is None:
            return c
        k = _native.SchedCoeffs()
        k.guidance, k.use_cfg = float(self.guidance), 1
        k.w[0], k.a, k.b = 1.0, 0.0, 1.0  # x' = eps : combine only
        out = torch.zeros_like(c)
        _native.cfg_sched_step(u, c, out, k, out)
        return out
