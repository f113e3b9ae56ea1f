'''Noise guides -- API mirror of /root/reference/pipeline/guide.py.

`SimpleGuide` keeps the reference's constructor and `noise_pred(latents, step)`
(guide.py:39-64) but the context never changes during a run (guide.py:30,43,49-53), so:

  * its K/V projections for all 16 cross-attention layers are computed ONCE here by
    K2 (`unet.build_kv_cache`) -- the uncond context is one cache row shared by the
    whole batch -- and every step's UNet call attends over that cache (K3);
  * the CFG combine `u + g (c - u)` (guide.py:61-63) runs in K4; `noise_pred_pair`
    hands the un-combined halves to the pipeline so CFG and the scheduler update fuse
    into one launch;
  * optionally the whole UNet forward is captured in a CUDA graph (B=1 is launch-bound).
'''
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch

from .. import _native
from ..encode.clip import CLIPEncoder
from ..unet import timestep_embedding


class GuideBase():
    def __init__(self, encoder: CLIPEncoder, unet, guidance: float,
                 steps: int) -> None:
        '''encoder: text encoder; unet: noise predictor; guidance: CFG scale `w`
        (enabled when > 1); steps: number of denoising steps (guide.py:9-33).'''
        self.encoder = encoder
        self.unet = unet
        self.uncond_embeds = encoder.prompt('')
        self.batch_size = 1
        self.guidance = guidance
        self.steps = steps

    def noise_pred(self, latents: torch.Tensor, step: int) -> torch.FloatTensor:
        raise NotImplementedError('noise_pred must be implemented.')


class SimpleGuide(GuideBase):
    def __init__(self, encoder: CLIPEncoder, unet, guidance: float, steps: int,
                 clip_embeds: torch.Tensor, use_cuda_graph: bool = False):
        GuideBase.__init__(self, encoder, unet, guidance, steps)
        self.embeds = clip_embeds
        self.batch_size = self.embeds.shape[0]
        self.use_cuda_graph = use_cuda_graph
        self._kv = None
        self._ctx_index = None
        self._plain_in = None
        self._temb_cache: Dict[float, torch.Tensor] = {}

    # ------------------------------------------------------------------ K2 cache
    @property
    def classifier_free_guidance(self) -> bool:
        return self.guidance > 1.0  # guide.py:47

    def _cache_key(self):
        e, u = self.embeds, self.uncond_embeds
        return (e.data_ptr(), e._version, tuple(e.shape), u.data_ptr(), u._version,
                self.classifier_free_guidance)

    def invalidate(self):
        '''Forget the K/V cache (it is also rebuilt when `embeds`, `uncond_embeds` or the CFG
        on/off state change: the reference reads these attributes on every step).'''
        self._kv = None

    def _ensure_cache(self):
        key = self._cache_key()
        if self._kv is not None and self.__dict__.get('_kv_key') == key:
            return
        self._kv_key = key
        self.batch_size = self.embeds.shape[0]
        if not hasattr(self.unet, 'build_kv_cache'):
            raise _native.NativeError(
                'SimpleGuide needs flexdiffuse_b200.unet.UNet2DConditionModel '
                '(K2/K3 cross-attention); got ' + type(self.unet).__name__)
        B = self.batch_size
        dev = self.embeds.device
        with torch.no_grad():
            if self.classifier_free_guidance:
                ctx = torch.cat([self.uncond_embeds.to(dev)[:1], self.embeds])
                index = [0] * B + list(range(1, B + 1))
            else:
                ctx = self.embeds
                index = list(range(B))
            self._kv = self.unet.build_kv_cache(ctx)
            self._ctx_index = torch.tensor(index, dtype=torch.int32, device=dev)

    def _temb(self, step, device) -> torch.Tensor:
        key = float(step)
        t = self._temb_cache.get(key)
        if t is None:
            t = timestep_embedding(torch.tensor([key]),
                                   self.unet.conv_in.out_channels).to(device)
            self._temb_cache[key] = t
        return t

    # ------------------------------------------------------------------ UNet call
    def _runner(self, latents: torch.Tensor):
        self._ensure_cache()
        return self.unet.graph_runner(latents.shape, self._kv.n_ctx,
                                      2 if self.classifier_free_guidance else 1)

    def model_input_buffer(self, latents: torch.Tensor) -> torch.Tensor:
        '''bf16 [B,4,h,w] buffer the UNet reads; K4 can write the next step's model input
        straight into it (`scaled_out`).  With CUDA graphs it is the runner's static input.'''
        if self.use_cuda_graph:
            return self._runner(latents).static_in
        dt = self.unet.conv_in.weight.dtype
        if self._plain_in is None or self._plain_in.shape != latents.shape:
            self._plain_in = torch.empty(latents.shape, dtype=dt, device=latents.device)
        return self._plain_in

    @torch.no_grad()
    def noise_pred_pair(self, latents: torch.Tensor,
                        step) -> Tuple[Optional[torch.Tensor], torch.Tensor]:
        '''(eps_uncond, eps_cond) before the CFG combine; eps_uncond is None without CFG.'''
        self._ensure_cache()
        buf = self.model_input_buffer(latents)
        if latents.data_ptr() != buf.data_ptr():
            buf.copy_(latents)
        temb = self._temb(step, latents.device)
        if self.use_cuda_graph:
            runner = self._runner(latents)
            runner.load(self, self._kv, self._ctx_index)
            out = runner.run(temb)
        else:
            x = torch.cat([buf, buf]) if self.classifier_free_guidance else buf
            out = self.unet(x, None, kv_cache=self._kv, ctx_index=self._ctx_index,
                            temb_sin=temb).sample
        if self.classifier_free_guidance:
            u, c = out.chunk(2)
            return u, c
        return None, out

    def noise_pred(self, latents: torch.Tensor, step) -> torch.FloatTensor:
        '''Reference form (guide.py:46-64): returns the CFG-combined prediction.  Must not
        modify `latents`.'''
        u, c = self.noise_pred_pair(latents, step)
        if u is None:
            # with CUDA graphs `c` is the runner's static output, overwritten by the next replay:
            # a caller keeping a history of predictions must get its own tensor
            return c.clone() if self.use_cuda_graph else c
        k = _native.SchedCoeffs()
        k.guidance, k.use_cfg = float(self.guidance), 1
        k.w[0], k.a, k.b = 1.0, 0.0, 1.0  # x' = eps : combine only
        # a = 0: the sample operand only has to be a finite fp32 tensor; use the output
        out = torch.zeros(c.shape, dtype=torch.float32, device=c.device)
        _native.cfg_sched_step(u.contiguous(), c.contiguous(), out, k, out)
        return out


class PromptGuide(SimpleGuide):
    def __init__(self, encoder: CLIPEncoder, unet, guidance: float, steps: int,
                 prompt: str | List[str], use_cuda_graph: bool = False):
        SimpleGuide.__init__(self, encoder, unet, guidance, steps,
                             encoder.prompt(prompt), use_cuda_graph)
        self.prompt = prompt
