'''CLIP encoders -- API mirror of /root/reference/encode/clip.py.

`preprocess` (clip.py:15-39) and `CLIPEncoder.prompt / .image` (clip.py:47-100) keep the
reference's signatures and numerics.  The weights, embeddings, layer norms and the attention
core stay transformers' / PyTorch's; on an sm_100 device every Linear of the two encoders
(q / k / v / out_proj, fc1 + activation, fc2 -- 99 % of the towers' arithmetic) runs on K11
`fd_linear_x3`, an fp32-accurate tcgen05 GEMM, and `visual_projection` on
K1P; the 257- / 77-token attention core is K12 (exact fp32).  Each tower forward is captured once per
input shape and replayed as a CUDA graph.
'''
from __future__ import annotations

from typing import Any, List

import numpy as np
import torch

CLIP_IMAGE_SIZE = 224
MAX_SINGLE_DIM = 512  # Stable Diffusion's native side length

_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _resized_rgb(image: Any) -> np.ndarray:
    '''The LANCZOS resize of clip.py:24-36: the longer side becomes 512, the shorter one is scaled with
    it and floored to a multiple of 64; uint8 [h, w, 3].'''
    from PIL.Image import LANCZOS
    w, h = image.size
    if h == w:
        w = h = MAX_SINGLE_DIM
    elif h > w:
        w, h = (int(w / (h / MAX_SINGLE_DIM)) // 64) * 64, MAX_SINGLE_DIM
    else:
        h, w = (int(h / (w / MAX_SINGLE_DIM)) // 64) * 64, MAX_SINGLE_DIM
    return np.array(image.resize((w, h), resample=LANCZOS).convert('RGB'))


# clip.py:37-39 computes `2.0 * (uint8.astype(float32) / 255.0) - 1.0` (numpy division, torch multiply / subtract) over the
# whole image in three passes; the same three float32 operations on the 256 possible byte values, then one gather, give
# the identical bits with one pass over the pixels (`Guide.embeds` is bounded by this host work + the vision tower)
_PIXEL_LUT = (2.0 * torch.from_numpy(np.arange(256).astype(np.float32) / 255.0) - 1.0).numpy()


def preprocess(image: Any) -> torch.Tensor:
    '''PIL image -> [1,3,h,w] float32 in [-1,1]; the longer side becomes 512, the
    shorter one is scaled with it and floored to a multiple of 64 (clip.py:24-39).'''
    return torch.from_numpy(np.take(_PIXEL_LUT, _resized_rgb(image))[None].transpose(0, 3, 1, 2))


class _TowerGraph:
    '''One CUDA-graph capture of a tower forward for a fixed input shape: the towers are ~300 (text)
    and ~600 (vision) small eager launches at batch 1, i.e. launch-bound; replaying them as a graph
    is what `Guide.embeds` calls per second are bounded by.  Same kernels, same
    order, same numerics as the eager call.'''
    def __init__(self, fn, example: torch.Tensor):
        self.static_in = example.clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(2):
                fn(self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.static_out = fn(self.static_in)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        self.static_in.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.static_out.clone()  # callers keep / edit the result (guidance.py:449, 467-472)


class _matmul_precision:
    '''Scoped `torch.backends.cuda.matmul.allow_tf32` (cuBLAS reads it at dispatch time, so it is baked
    into a graph at capture).'''
    def __init__(self, tf32: bool):
        self.tf32 = tf32

    def __enter__(self):
        self.old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = self.tf32

    def __exit__(self, *a):
        torch.backends.cuda.matmul.allow_tf32 = self.old


_X3_ACTS = {'quick_gelu': 1, 'gelu': 2}


def _x3_ok(encoder, hidden: torch.Tensor) -> bool:
    '''K11 serves fp32 CUDA towers whose widths are multiples of 64 (every CLIP variant).'''
    if not (hidden.is_cuda and hidden.dtype == torch.float32) or not len(encoder.layers):
        return False
    l0 = encoder.layers[0]
    return (l0.mlp.fc1.in_features % 64 == 0 and l0.mlp.fc1.out_features % 64 == 0
            and l0.mlp.fc1.weight.dtype == torch.float32)


_qkv_cache = {}   # (q, k, v weight storage + version) -> (concatenated weight, bias, the three source weights)


def _qkv_weight(at):
    '''[3C, C] weight and [3C] bias of one attention block, so q, k and v are ONE K11 GEMM; rebuilt when any
    of the three projections changes (the cache keeps the sources alive: no address reuse under a stale entry).'''
    ws = (at.q_proj.weight, at.k_proj.weight, at.v_proj.weight)
    bs = (at.q_proj.bias, at.k_proj.bias, at.v_proj.bias)
    key = tuple((w.data_ptr(), w._version) for w in ws) + tuple((b.data_ptr(), b._version) for b in bs)
    ent = _qkv_cache.get(key)
    if ent is None:
        if len(_qkv_cache) > 256:
            _qkv_cache.clear()
        ent = (torch.cat([w.detach() for w in ws]).contiguous(), torch.cat([b.detach() for b in bs]).contiguous(), ws, bs)
        _qkv_cache[key] = ent
    return ent[0], ent[1]


def _x3_encoder(encoder, hidden: torch.Tensor, causal: bool) -> torch.Tensor:
    '''transformers' CLIPEncoder.forward (pre-LN blocks: `x + attn(ln1(x))`, `x + mlp(ln2(x))`) with every
    Linear on K11 and the attention core on K12.  The two LayerNorms are fused into the operand split of the
    GEMM they feed, q, k and v are one GEMM against the concatenated weights, fc1's bias and activation and both
    residual adds ride in GEMM epilogues: 7 launches per block.  The text tower is causal, as CLIPTextTransformer
    builds its mask.'''
    from .. import _native
    B, T, C = hidden.shape
    hidden = hidden.reshape(B * T, C)
    M = B * T

    def ln_operand(ln, x):   # operand buffer of ln(x): fused LayerNorm + split where K11 has the kernel
        if C % 64 == 0 and C <= 8192 and ln.elementwise_affine and ln.bias is not None:
            return _native.x3_split_ln(x, ln.weight, ln.bias, ln.eps)
        return _native.x3_split(ln(x))

    for layer in encoder.layers:
        at, mlp = layer.self_attn, layer.mlp
        H = at.num_heads
        op = ln_operand(layer.layer_norm1, hidden)
        if at.q_proj.bias is not None and at.k_proj.bias is not None and at.v_proj.bias is not None:
            w_qkv, b_qkv = _qkv_weight(at)
            qkv = _native.linear_x3(None, w_qkv, b_qkv, operand=op, rows=M).view(B, T, 3 * C)
            q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
        else:
            q, k, v = (_native.linear_x3(None, p.weight, p.bias, operand=op, rows=M).view(B, T, C)
                       for p in (at.q_proj, at.k_proj, at.v_proj))
        if _native.attention_f32_supported(T, C // H):
            o = _native.attention_f32(q, k, v, H, at.scale, causal).view(M, C)
        else:  # sequences the K12 kernel does not serve: torch's fp32 SDPA
            from torch.nn.functional import scaled_dot_product_attention as sdpa
            q, k, v = (t.reshape(B, T, H, C // H).transpose(1, 2) for t in (q, k, v))
            o = sdpa(q, k, v, is_causal=causal, scale=at.scale).transpose(1, 2).reshape(M, C)
        hidden = _native.linear_x3(o, at.out_proj.weight, at.out_proj.bias, residual=hidden)   # x + attn(ln1(x))
        op = ln_operand(layer.layer_norm2, hidden)
        act = _X3_ACTS.get(getattr(mlp.config, 'hidden_act', None), 0)
        h = _native.linear_x3(None, mlp.fc1.weight, mlp.fc1.bias, act=act, operand=op, rows=M)
        if not act:
            h = mlp.activation_fn(h)
        hidden = _native.linear_x3(h, mlp.fc2.weight, mlp.fc2.bias, residual=hidden)           # x + mlp(ln2(x))
    return hidden.view(B, T, C)


class CLIPEncoder():
    def __init__(self, clip, token, cuda_graph: bool = True, tf32: bool = False, x3: bool = True) -> None:
        '''`cuda_graph` / `tf32` / `x3` are not reference arguments.  cuda_graph: on a CUDA device, replay
        each tower as a captured graph per input shape (call `invalidate()` after changing the CLIP
        weights).  x3 (default): the towers' Linears on K11, fp32-accurate on the tensor cores.
        tf32 (only without x3): cuBLAS TF32 GEMMs, ~3e-4 relative error of the embeddings -- kept as the
        comparison point, OFF by default because the reference towers are exact fp32 and K1's decisions
        sit on near-ties of 100 * cos.'''
        self.clip = clip
        self.token = token
        self.cuda_graph = cuda_graph
        self.tf32 = tf32
        self.x3 = x3
        self._graphs = {}

    def invalidate(self) -> None:
        self._graphs = {}

    def _run(self, kind: str, fn, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            return fn(x)
        with _matmul_precision(self.tf32):
            if not self.cuda_graph or torch.is_grad_enabled() \
                    or torch.cuda.is_current_stream_capturing():
                return fn(x)
            key = (kind, tuple(x.shape), x.dtype, self.tf32, self.x3)
            g = self._graphs.get(key)
            if g is None:
                try:
                    g = _TowerGraph(fn, x)
                except Exception:  # a tower that cannot be captured still runs, un-graphed
                    torch.cuda.synchronize()
                    g = False
                self._graphs[key] = g
            return g(x) if g else fn(x)

    def prompt(self, prompt: str | List[str]) -> torch.Tensor:
        '''Final-layer-norm hidden states of the text tower, NOT projected
        (clip.py:57-65): [B, 77, 768].'''
        ids = self.token(prompt, padding='max_length',
                         max_length=self.token.model_max_length,
                         truncation=True, return_tensors='pt').input_ids
        return self._run('text', self._text, ids.to(self.clip.device))

    def _text(self, ids: torch.Tensor) -> torch.Tensor:
        tm = self.clip.text_model
        if self.x3 and ids.is_cuda:
            hidden = tm.embeddings(input_ids=ids.view(-1, ids.shape[-1]))
            if _x3_ok(tm.encoder, hidden):
                return tm.final_layer_norm(_x3_encoder(tm.encoder, hidden, causal=True))
        return tm(ids)[0]

    def image(self, image) -> torch.Tensor:
        '''All 257 vision tokens through post_layernorm and visual_projection
        (clip.py:76-100): [1, 257, 768].  Note the CLIP mean/std are applied to a
        [-1,1] tensor, exactly as the reference does.'''
        from torchvision.transforms.functional import (InterpolationMode,
                                                       center_crop, normalize,
                                                       resize)
        dev = self.clip.device
        # (crop / antialiased bicubic resize stay on the CPU like the reference: torchvision's CUDA resize differs
        # from its CPU one by ~1e-4 in the pixels, 3e-4 in the embeddings -- measured, not worth 0.4 ms)
        x = preprocess(image)
        side = min(x.shape[-2:])
        x = center_crop(x, [side, side])
        x = resize(x, [CLIP_IMAGE_SIZE, CLIP_IMAGE_SIZE],
                   interpolation=InterpolationMode.BICUBIC, antialias=True)
        x = normalize(x, list(_CLIP_MEAN), list(_CLIP_STD)).to(dev)
        return self._run('image', self._vision, x)

    def _vision(self, x: torch.Tensor) -> torch.Tensor:
        vm = self.clip.vision_model
        hidden = vm.pre_layrnorm(vm.embeddings(x))
        if self.x3 and _x3_ok(vm.encoder, hidden):
            hidden = _x3_encoder(vm.encoder, hidden, causal=False)
        else:
            hidden = vm.encoder(inputs_embeds=hidden, output_attentions=False,
                                output_hidden_states=False, return_dict=True)[0]
        hidden = vm.post_layernorm(hidden)
        proj = self.clip.visual_projection
        if (hidden.is_cuda and hidden.dtype == torch.float32 and proj.bias is None
                and proj.in_features % 64 == 0 and proj.out_features % 4 == 0):
            # K1P: clip.py:100 on tcgen05 with fp32 accuracy -- the guide embeddings K1 consumes are
            # produced by the hand-written path
            from .. import _native
            return _native.visual_projection(hidden, proj.weight)
        return proj(hidden)
