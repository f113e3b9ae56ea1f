'''Prompt / image guided embeddings -- API mirror of /root/reference/guidance.py with the
arithmetic on the B200:

    _map_emb        (guidance.py:23-85)    \
    _clustered_...  (guidance.py:135-172)   >  one launch of K1 `fd_sim_blend`
    _blend_weights  (guidance.py:175-193)  /   (tcgen05 split-fp16 similarity GEMM, per-lane
    Tweener.tween   (guidance.py:215-272) /     softmax, warp-shuffle mapping + weights, lerp)
    ConceptMapper   (guidance.py:275-312)      two more K1 mappings + a row scatter
    Guide           (guidance.py:315-474)      host glue, same signature and defaults

Same names, argument meaning and error behaviour as the reference; there is no CPU
path (embeddings must live on an sm_100 device, otherwise `NativeError`).

Documented divergences:
  * batch > 1: the reference raises IndexError; here each prompt gets the
    solo path against the shared guide, all in one launch (the evident intent of
    guidance.py:439-444).
  * the reference prints tween statistics on every call; here only when
    `guidance.VERBOSE` is true (printing forces a device->host sync).
'''
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native
from .encode.clip import CLIPEncoder

CLIP_IMAGE_SIZE = 224
MAX_SINGLE_DIM = 512

GUIDE_ORDER_TEXT = 0
GUIDE_ORDER_ALIGN = 1
GUIDE_ORDER_DIRECT = 2

VERBOSE = False


def _native_params(threshold_floor, threshold_mult, clustered, max_guidance,
                   header_max, align_mode, mapping_reuse, slerp=False):
    p = _native.TweenParams()
    p.threshold_floor = float(threshold_floor)
    p.threshold_mult = float(threshold_mult)
    p.clustered = float(clustered)
    p.max_guidance = float(max_guidance)
    p.header_max = float(header_max)
    p.align_mode = int(align_mode)
    p.mapping_reuse = int(bool(mapping_reuse))
    p.blend_mode = (_native.BLEND_MODE_SLERP if slerp
                    else _native.BLEND_MODE_LERP)
    return p


def _as_f32_3d(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dim() != 3:
        # the reference indexes [0, i] / shape[1] (guidance.py:48,55,76)
        raise IndexError(f'{name} must be [batch, tokens, dim], got '
                         f'{tuple(t.shape)}')
    return t.float().contiguous()


def _map_emb(alt_emb: torch.Tensor,
             txt_emb: torch.Tensor,
             alt_emb_reuse: bool = True,
             guide_order: int = GUIDE_ORDER_ALIGN) -> np.ndarray:
    '''Map alternate (image / text) embeddings onto text embeddings by highest
    alignment (guidance.py:23-85).  Returns the reference's host table:
    float64 [T, 2] = (alt index, softmax similarity); row r <-> text token r+1.'''
    txt = _as_f32_3d(txt_emb, 'txt_emb')[:1]
    alt = _as_f32_3d(alt_emb, 'alt_emb')[:1]
    T = txt.shape[1]
    # weights are irrelevant for the map; a neutral parameter set keeps the blend trivial
    prm = _native_params(0.0, 0.0, 0.0, 0.0, 1.0, guide_order, alt_emb_reuse)
    lin = torch.zeros((1, T), dtype=torch.float32, device=txt.device)
    res = _native.sim_blend(txt, alt, [prm], lin)
    mapped = np.zeros((T, 2))
    mapped[:, 0] = res['map_idx'][0, 0].cpu().numpy()
    mapped[:, 1] = res['map_s'][0, 0].cpu().numpy().astype(np.float64)
    return mapped


class Tweener():
    def __init__(self,
                 threshold: Tuple[float, float] = (0.5, 0.5),
                 linear: Tuple[float, float] = (0.0, 0.5),
                 clustered: float = 0.5,
                 max_guidance: float = 0.5,
                 header_max: float = 0.15,
                 align_mode: int = GUIDE_ORDER_ALIGN,
                 mapping_reuse: bool = True,
                 slerp: bool = False) -> None:
        # `slerp` is the one argument the reference's Tweener (guidance.py:203-213) does
        # not have: rows it would lerp are spherically interpolated instead (off by
        # default, so the default object behaves exactly like the reference's)
        self.slerp = slerp
        self.threshold_floor = threshold[0]
        self.threshold_mult = threshold[1]
        self.linear_start = linear[0]
        self.linear_end = linear[1]
        self.clustered = clustered
        self.max_guidance = max_guidance
        self.header_max = header_max
        self.align_mode = align_mode
        self.mapping_reuse = mapping_reuse

    def _params(self):
        return _native_params(self.threshold_floor, self.threshold_mult,
                              self.clustered, self.max_guidance,
                              self.header_max, self.align_mode,
                              self.mapping_reuse, self.slerp)

    def _linear(self, steps: int, device) -> torch.Tensor:
        # guidance.py:231-233 -- torch.linspace on the host, exactly as the reference
        return torch.linspace(self.linear_start, self.linear_end,
                              steps=steps).to(device)

    def tween_batch(self, base_emb: torch.Tensor, alt_emb: torch.Tensor,
                    check: bool = True):
        '''[B,T,D] prompts x shared [1,A,D] (or per-prompt [B,A,D]) guide in one
        launch.  Returns (blended [B,T,D], kernel result dict).'''
        base = _as_f32_3d(base_emb, 'base_emb')
        alt = _as_f32_3d(alt_emb, 'alt_emb')
        lin = self._linear(base.shape[1], base.device)[None]
        res = _native.sim_blend(base, alt, [self._params()], lin, want_maps=VERBOSE)
        if check or VERBOSE:
            status = res['status'].cpu()
            if VERBOSE:
                s = res['map_s'][:, 0].double().cpu().numpy()
                for b in range(base.shape[0]):
                    print(f'Tweening with, Avg Similarity: {s[b].mean():.2%}, '
                          f'Threshold: {self.threshold_floor:.2%}, '
                          f'Threshold Multiplier: {self.threshold_mult:.2%}, '
                          f'Clustered: {self.clustered:.2%}, '
                          f'Linear: {self.linear_start:.2%}'
                          f'-{self.linear_end:.2%}, '
                          f'Guidance Max: {self.max_guidance:.2%}')
                    print('Alt Embed Blend Weights:', res['weights'][b, 0].shape,
                          ':', res['weights'][b, 0].cpu())
            if check and bool((status == _native.FD_BLEND_RANGE).any()):
                raise ValueError('text embeddings outside the supported range (|x| < 1023, finite)')
            if check and bool(
                (status == _native.FD_BLEND_ZERO_DIVISION).any()):
                # two adjacent similarity peaks: guidance.py:111-112 divides by zero
                raise ZeroDivisionError('float division by zero')
        return res['out'][:, 0].to(base_emb.dtype), res

    def tween(self, base_emb: torch.Tensor,
              alt_emb: torch.Tensor) -> torch.Tensor:
        '''guidance.py:215-272.  Returns a fresh tensor like `base_emb`.'''
        return self.tween_batch(base_emb, alt_emb)[0]


class ConceptMapper():
    def __init__(self, guide_embeddings: torch.Tensor,
                 concept_embeddings: torch.Tensor) -> None:
        self.guide_embeddings = guide_embeddings
        self.concept_embeddings = concept_embeddings
        self.concept_mappings = _map_emb(guide_embeddings, concept_embeddings,
                                         False, GUIDE_ORDER_TEXT)
        if VERBOSE:
            print('Image Feature and Concept alignment:')
            for txt_i, (img_i, s) in enumerate(self.concept_mappings, 1):
                print(f'ConceptTok {txt_i:>02d} ImgTok '
                      f'{int(img_i):>02d} {100 * s:.2f}%')

    def map(self,
            base_embeddings: torch.Tensor,
            output_embeddings: Optional[torch.Tensor] = None) -> torch.Tensor:
        '''guidance.py:288-312: prompt tokens whose best concept match exceeds 0.9
        are replaced by that concept's guide token.  Edits (and returns)
        `output_embeddings` in place, like the reference.'''
        if output_embeddings is None:
            output_embeddings = base_embeddings.clone()
        concept_text = _map_emb(self.concept_embeddings, base_embeddings, True,
                                GUIDE_ORDER_ALIGN)
        dst, src = [], []
        for txt_i, (concept_i, s) in enumerate(concept_text, 1):
            cmi = int(concept_i) - 1  # mappings start from token 1
            if cmi < 0 or not s > 0.9:
                continue
            dst.append(txt_i)
            src.append(int(self.concept_mappings[cmi, 0]))
        if dst:
            dev = output_embeddings.device
            output_embeddings[0, torch.tensor(dst, device=dev)] = (
                self.guide_embeddings[0, torch.tensor(src, device=dev)].to(
                    output_embeddings.dtype))
        return output_embeddings


class Guide():
    def __init__(self, clip, tokenizer, device: str = 'cuda', tf32_towers: bool = False) -> None:
        '''Context for generating prompt / image embeddings and tweening them
        (guidance.py:316-335).  `tf32_towers` (not a reference argument): cuBLAS TF32 towers instead of the
        default fp32-accurate K11 ones; see CLIPEncoder.'''
        self.clip = clip
        self.tokenizer = tokenizer
        self.device = device
        self.encoder = CLIPEncoder(clip, tokenizer, tf32=tf32_towers, x3=not tf32_towers)
        # header token of this embed is used for direct image guidance
        self.placeholder_embed = self.encoder.prompt('{}')

    def embeds(self,
               prompt: str | List[str] = '',
               guide=None,
               mapping_concepts: str = '',
               guide_threshold_mult: float = 0.5,
               guide_threshold_floor: float = 0.5,
               guide_clustered: float = 0.5,
               guide_linear: Tuple[float, float] = (0.0, 0.5),
               guide_max_guidance: float = 0.5,
               guide_header_max: float = 0.15,
               guide_mode: int = GUIDE_ORDER_ALIGN,
               guide_reuse: bool = True,
               guide_slerp: bool = False) -> torch.Tensor:
        '''Same arguments, defaults and return value as guidance.py:337-474
        (`guide_slerp` is an extension, see Tweener).'''
        if isinstance(prompt, str):
            prompt = prompt.strip()
        elif isinstance(prompt, list):
            prompt = [ss for ss in (s.strip() for s in prompt) if ss]
        else:
            raise ValueError(f'`prompt` has to be of type `str` '
                             f'or `list` but is {type(prompt)}')
        if not prompt and guide is None:
            raise ValueError('No prompt, or guide image provided.')

        text_embeddings = self.encoder.prompt(prompt) if prompt else None
        guide_embeddings = None
        concept_mapper = None
        guide_is_text = isinstance(guide, str)
        if guide is not None:
            if guide_is_text:
                guide = guide.strip()
                if guide:
                    guide_embeddings = self.encoder.prompt(guide)
            else:
                guide_embeddings = self.encoder.image(guide)
                if mapping_concepts:
                    concept_mapper = ConceptMapper(
                        guide_embeddings, self.encoder.prompt(mapping_concepts))
        tweener = Tweener((guide_threshold_floor, guide_threshold_mult),
                          guide_linear, guide_clustered, guide_max_guidance,
                          guide_header_max, guide_mode, guide_reuse,
                          guide_slerp)

        if text_embeddings is not None:
            if guide_embeddings is None:
                return text_embeddings  # the encoder's tensor itself (gd.py:449)
            clip_embeddings, _ = tweener.tween_batch(text_embeddings,
                                                     guide_embeddings)
            if concept_mapper is not None:
                for b in range(clip_embeddings.shape[0]):
                    concept_mapper.map(text_embeddings[b:b + 1],
                                       clip_embeddings[b:b + 1])
            if VERBOSE:
                print('Tweened text and image embeddings:',
                      guide_embeddings.shape, ' text shape:',
                      text_embeddings.shape, ' embed shape:',
                      clip_embeddings.shape)
            return clip_embeddings

        assert guide_embeddings is not None
        if guide_is_text:
            print('Warning: using the guide like prompt.. just use prompt.')
            return guide_embeddings
        print('Warning: trying to guide purely from image, '
              'this will generate weird stuff, enjoy :)\n'
              'If you\'re bored try an image of yourself '
              'and see what the model thinks.')
        # first 77 image tokens, header pulled 85% towards the text header; edits a view of
        # guide_embeddings in place like guidance.py:467-472
        clip_embeddings = guide_embeddings[:, :self.tokenizer.model_max_length, :]
        d_emb = self.placeholder_embed[:, 0, :] - clip_embeddings[:, 0, :]
        clip_embeddings[:, 0, :] += d_emb * 0.85
        return clip_embeddings
