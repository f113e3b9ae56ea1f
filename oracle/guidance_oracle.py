'''ORACLE (test infrastructure, never shipped, never on the product path).

CPU restatement of the reference's image-guided embedding blend:

    /root/reference/guidance.py:23-85    _map_emb
    /root/reference/guidance.py:88-132   _traverse_a_to_b
    /root/reference/guidance.py:135-172  _clustered_guidance
    /root/reference/guidance.py:175-193  _blend_weights
    /root/reference/guidance.py:196-272  Tweener / Tweener.tween
    /root/reference/guidance.py:275-312  ConceptMapper

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this module.

Pinning: `tests/test_oracle_pinning.py` checks this restatement (a) bit-for-bit
against the committed golden vectors in `tests/golden/` that were produced by
importing the unmodified reference (script: `tests/golden/make_golden.py`), and
(b) when `/root/reference` is present, against the reference run live on fresh
random parameters.  The reference has no tests or fixtures of its own
(SURVEY.md section 4), so those self-generated vectors are the pin.

The restatement is structured differently from the reference on purpose: the
similarity matrix is built once, the sort + greedy scan is expressed with numpy
lexsort, and the weight heuristics are closed-form per token.  Arithmetic that
decides the output bits (fp32 row-wise matmul + softmax, fp32 weight algebra,
float64 comparisons) follows the reference's dtype at every step.
'''
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np
import torch

GUIDE_ORDER_TEXT = 0
GUIDE_ORDER_ALIGN = 1
GUIDE_ORDER_DIRECT = 2


def similarity_matrix(alt_emb: torch.Tensor,
                      txt_emb: torch.Tensor,
                      rowwise: bool = True) -> torch.Tensor:
    '''P[i, j] = softmax_j(100 * cos(alt_i, txt_j)), fp32, shape [A, T].

    guidance.py:43-50.  `rowwise=True` issues one [1,D]@[1,D,T] matmul per alt
    token exactly like the reference's loop (bit-identical on the same
    machine); `rowwise=False` is one batched matmul (differs by ~1e-6, Q19).
    '''
    altft = alt_emb / alt_emb.norm(dim=-1, keepdim=True)
    txtft = txt_emb / txt_emb.norm(dim=-1, keepdim=True)
    if rowwise:
        rows = []
        for i in range(altft.shape[1]):
            row = altft[0, i].unsqueeze(0)
            rows.append((100.0 * (row @ txtft.mT)).softmax(dim=-1)[0, 0])
        return torch.stack(rows)
    return (100.0 * (altft[0] @ txtft[0].T)).softmax(dim=-1)


def map_from_similarity(P: np.ndarray, n_txt: int, reuse: bool,
                        order: int) -> np.ndarray:
    '''guidance.py:54-85 on a precomputed similarity matrix.

    P: float64 [A, T] (values are exact fp32).  Returns float64 [T, 2] =
    (alt index, similarity); row r describes text token r+1, the
    last row is always (0, 0).
    '''
    A = P.shape[0]
    S = P[:, 1:]  # header column dropped, re-enumerated from 0
    ncol = S.shape[1]
    mapped = np.zeros((n_txt, 2))
    if order == GUIDE_ORDER_DIRECT:
        # only (alt_i == txt_i) pairs survive the scan
        for r in range(min(A, ncol)):
            mapped[r] = (r, S[r, r])
        return mapped
    alt_i = np.repeat(np.arange(A), ncol)
    txt_i = np.tile(np.arange(ncol), A)
    s = S.reshape(-1)
    if order == GUIDE_ORDER_TEXT:
        # asc text, desc similarity, asc alt
        perm = np.lexsort((alt_i, -s, txt_i))
    else:
        # desc similarity, asc text, asc alt
        perm = np.lexsort((alt_i, txt_i, -s))
    used = np.zeros(A, dtype=bool)
    for k in perm:
        t, i = txt_i[k], alt_i[k]
        if mapped[t, 1] > 0 or used[i]:
            continue
        mapped[t] = (i, s[k])
        if not reuse:
            used[i] = True
    return mapped


def map_emb(alt_emb: torch.Tensor,
            txt_emb: torch.Tensor,
            alt_emb_reuse: bool = True,
            guide_order: int = GUIDE_ORDER_ALIGN,
            rowwise: bool = True) -> np.ndarray:
    '''Restatement of guidance.py:23-85 `_map_emb`.'''
    P = similarity_matrix(alt_emb, txt_emb, rowwise).double().numpy()
    return map_from_similarity(P, txt_emb.shape[1], alt_emb_reuse, guide_order)


def clustered_weights(mapped: np.ndarray, threshold: float,
                      guidance: float) -> Optional[torch.Tensor]:
    '''guidance.py:135-172 (+ :88-132) in closed form per token.

    Raises ZeroDivisionError for two adjacent peaks exactly where the
    reference does.
    '''
    n = mapped.shape[0]
    s = mapped[:, 1]
    peaks = [
        r for r in range(1, n - 1)
        if not (s[r] < threshold) and s[r - 1] <= s[r] >= s[r + 1]
    ]
    if not peaks:
        return None
    for p1, p2 in zip(peaks, peaks[1:]):
        if p2 - p1 == 1:
            # valley = p1 + ceil(1/2) = p2 -> traverse_right(p2, p2): slope / 0
            raise ZeroDivisionError('float division by zero')
    valleys = [0] + [p1 + math.ceil((p2 - p1) / 2)
                     for p1, p2 in zip(peaks, peaks[1:])] + [n - 1]
    w = torch.ones((n,))
    w[0] -= 1.0
    for k, p in enumerate(peaks):
        vl, vr = valleys[k], valleys[k + 1]
        gl = 1.0 / (p - vl)
        for i in range(1, p - vl):
            w[p - i] -= gl * i
        gr = 1.0 / (vr - p)
        for i in range(1, vr - p + 1):
            w[p + i] -= gr * i
    return w * guidance


def blend_weights(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    '''guidance.py:175-193: global-sign switch.'''
    assert a.shape == b.shape, f'Tensor shapes a={a.shape} != b={b.shape}'
    if a.max() >= 0:
        return torch.maximum(a, b) if b.max() >= 0 else a + b
    return torch.minimum(a, b)


@dataclass
class TweenParams:
    '''Constructor arguments of guidance.py:196-213 `Tweener`.'''
    threshold: Tuple[float, float] = (0.5, 0.5)  # (floor, mult)
    linear: Tuple[float, float] = (0.0, 0.5)
    clustered: float = 0.5
    max_guidance: float = 0.5
    header_max: float = 0.15
    align_mode: int = GUIDE_ORDER_ALIGN
    mapping_reuse: bool = True
    slerp: bool = False  # extension, NOT a reference argument (see slerp_rows)


SLERP_DOT_THRESHOLD = 0.9995


def slerp_rows(base_rows: torch.Tensor, alt_rows: torch.Tensor,
               iw: np.ndarray) -> Tuple[torch.Tensor, np.ndarray, np.ndarray]:
    '''EXTENSION -- parity unpinned: the reference has no slerp (its only blend
    is the lerp at guidance.py:271); BASELINE.json's north_star names an optional
    one, so this states the usual per-vector spherical interpolation it is checked
    against: out = sin((1-w)O)/sin(O) * base + sin(wO)/sin(O) * alt with
    O = acos(<base,alt>/(|base||alt|)); rows with |cos O| > 0.9995 keep the lerp.
    Returns (rows [T,D] float64, used_slerp [T] bool, cos [T]).'''
    b = base_rows.double().numpy()
    a = alt_rows.double().numpy()
    den = np.linalg.norm(b, axis=1) * np.linalg.norm(a, axis=1)
    cosv = np.where(den > 0, (b * a).sum(1) / np.where(den > 0, den, 1), 1.0)
    use = np.abs(cosv) <= SLERP_DOT_THRESHOLD
    theta = np.arccos(np.clip(cosv, -1, 1))
    st = np.where(use, np.sin(theta), 1.0)
    ca = np.sin(theta - theta * iw) / st
    cb = np.sin(theta * iw) / st
    out = ca[:, None] * b + cb[:, None] * a
    return torch.from_numpy(out), use, cosv


def tween_weights(mapped: np.ndarray, prm: TweenParams) -> torch.Tensor:
    '''guidance.py:219-254: the final `alt_weights` vector (fp32, [T]).'''
    n = mapped.shape[0]
    avg = mapped[:, 1].mean()
    w = torch.linspace(prm.linear[0], prm.linear[1], steps=n)
    if prm.clustered != 0:
        cw = clustered_weights(mapped, avg, prm.clustered)
        if cw is not None:
            w = blend_weights(w, cw)
    floor, mult = prm.threshold
    if mult != 0:
        th = torch.ones_like(w) * mult
        th[torch.from_numpy(mapped[:, 1] < floor)] = 0
        w = blend_weights(w, th)
    if prm.header_max < 1.0:
        hw = w[0].item()
        w[0] = min(hw, prm.header_max) if hw >= 0 else max(hw, -prm.header_max)
    return w


def tween_select(mapped: np.ndarray, w: torch.Tensor,
                 max_guidance: float) -> Tuple[np.ndarray, np.ndarray]:
    '''guidance.py:259-271 decisions: sel 0 = keep text, 1 = take alt, 2 = lerp;
    iw = the (python float) blend weight.'''
    n = mapped.shape[0]
    sel = np.zeros(n, dtype=np.int64)
    iw = np.zeros(n, dtype=np.float64)
    for r in range(n):
        sd = 1.0 - mapped[r, 1]
        v = min(w[r].item(), max_guidance)
        iw[r] = v
        sel[r] = 0 if v == 0 else (1 if abs(v) >= sd else 2)
    return sel, iw


def tween(base_emb: torch.Tensor,
          alt_emb: torch.Tensor,
          prm: TweenParams,
          rowwise: bool = True,
          return_parts: bool = False):
    '''Restatement of guidance.py:215-272 `Tweener.tween` (solo path:
    base [1,T,D], alt [1,A,D]) -> [1,T,D].'''
    mapped = map_emb(alt_emb, base_emb, prm.mapping_reuse, prm.align_mode,
                     rowwise)
    w = tween_weights(mapped, prm)
    sel, iw = tween_select(mapped, w, prm.max_guidance)
    out = torch.zeros_like(base_emb)
    idx = torch.from_numpy(mapped[:, 0].astype(np.int64))
    alt_rows = alt_emb[0, idx]
    base_rows = base_emb[0]
    lerp = base_rows + (alt_rows - base_rows) * torch.tensor(
        iw, dtype=torch.float64).to(base_emb.dtype)[:, None]
    if prm.slerp:
        sl, use, _ = slerp_rows(base_rows, alt_rows, iw)
        lerp = torch.where(torch.from_numpy(use)[:, None],
                           sl.to(base_emb.dtype), lerp)
    sel_t = torch.from_numpy(sel)[:, None]
    out[0] = torch.where(sel_t == 0, base_rows,
                         torch.where(sel_t == 1, alt_rows, lerp))
    if return_parts:
        return out, mapped, w, sel, iw
    return out


def tween_batch(base_emb: torch.Tensor, alt_emb: torch.Tensor,
                prm: TweenParams, rowwise: bool = True) -> torch.Tensor:
    '''Batched semantics the reference intends but cannot run (SURVEY Q5:
    guidance.py:443-444 raises IndexError): the solo path per prompt against a
    shared (or per-prompt) guide.'''
    outs = []
    for b in range(base_emb.shape[0]):
        g = alt_emb if alt_emb.shape[0] == 1 else alt_emb[b:b + 1]
        outs.append(tween(base_emb[b:b + 1], g, prm, rowwise))
    return torch.cat(outs)


def concept_map(guide_emb: torch.Tensor, concept_emb: torch.Tensor,
                base_emb: torch.Tensor,
                output_emb: Optional[torch.Tensor] = None) -> torch.Tensor:
    '''Restatement of guidance.py:275-312 `ConceptMapper(...).map(...)`.'''
    concept_mappings = map_emb(guide_emb, concept_emb, False, GUIDE_ORDER_TEXT)
    if output_emb is None:
        output_emb = base_emb.clone()
    concept_text = map_emb(concept_emb, base_emb, True, GUIDE_ORDER_ALIGN)
    for txt_i, (concept_i, s) in enumerate(concept_text, 1):
        cmi = int(concept_i) - 1
        if cmi < 0:
            continue
        if s > 0.9:
            output_emb[0, txt_i] = guide_emb[0, int(concept_mappings[cmi, 0])]
    return output_emb


# --------------------------------------------------------------- synthetic data
def synthetic_pair(seed: int, T: int = 77, A: int = 257, D: int = 768,
                   planted: int = 12) -> Tuple[torch.Tensor, torch.Tensor]:
    '''Seeded (text, guide) embeddings with `planted` correlated token pairs so
    the arg-max, peak and threshold branches all fire.
    numpy RandomState => identical on every machine.'''
    rs = np.random.RandomState(seed)
    txt = rs.standard_normal((1, T, D)).astype(np.float32)
    img = rs.standard_normal((1, A, D)).astype(np.float32)
    if planted:
        toks = rs.choice(np.arange(1, T), size=min(planted, T - 1),
                         replace=False)
        rows = rs.choice(np.arange(A), size=len(toks), replace=False)
        for k, i in zip(toks, rows):
            gain = rs.uniform(0.5, 2.0)
            noise = rs.uniform(0.0, 0.6)
            img[0, i] = (txt[0, k] * gain + noise *
                         rs.standard_normal(D)).astype(np.float32)
    return torch.from_numpy(txt), torch.from_numpy(img)


def random_params(rs: np.random.RandomState) -> TweenParams:
    '''Random Tweener parameters covering all modes (incl. negative weights).'''
    return TweenParams(
        threshold=(float(rs.choice([0.0, 0.25, 0.5, 0.75, 0.9])),
                   float(rs.choice([0.0, 0.25, 0.5, -0.3, 1.0]))),
        linear=(float(rs.choice([0.0, 0.1, -0.2, 0.5])),
                float(rs.choice([0.5, 0.0, 1.0, -0.5]))),
        clustered=float(rs.choice([0.0, 0.15, 0.5, 1.0, -0.4])),
        max_guidance=float(rs.choice([0.35, 0.5, 1.0, 0.05])),
        header_max=float(rs.choice([0.0, 0.15, 1.0, 0.5])),
        align_mode=int(rs.choice([0, 1, 2])),
        mapping_reuse=bool(rs.choice([True, False])))
