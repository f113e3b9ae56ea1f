'''Size-independent properties of the hot-path kernels at BASELINE.json's full sizes (no oracle run
needed): what must hold for ANY input, checked on the shapes the benchmark uses.'''
import pytest
import torch

pytestmark = pytest.mark.gpu

T_VALID, T_PAD = 77, 80


def _attn(native, q, kv, C, idx):
    out = native.cross_attn(q, kv, 0, C, idx, 8, T_VALID, T_PAD, (C // 8)**-0.5)
    torch.cuda.synchronize()
    return out


@pytest.mark.parametrize('n_q,C', [(4096, 320), (1024, 640), (256, 1280)])
def test_k3_output_is_a_convex_combination_of_value_rows(native, cuda_dev, n_q, C):
    '''softmax weights are >= 0 and sum to 1, so every output element lies between the smallest
    and largest value of its (head, channel) column over the 77 valid keys -- 8 samples, the
    benchmark's shapes; the 3 padded key rows hold huge values that must never leak in.'''
    S = 8
    g = torch.Generator().manual_seed(n_q)
    q = (torch.randn(S, n_q, C, generator=g) * 2).to(cuda_dev).bfloat16()
    kv = torch.randn(2 * T_PAD, 2 * C, generator=g)
    kv[T_VALID:T_PAD] = 3e4     # padding rows of context 0 (masked keys)
    kv[T_PAD + T_VALID:] = -3e4  # ... and of context 1
    kv = kv.to(cuda_dev).bfloat16()
    idx = (torch.arange(S, dtype=torch.int32) % 2).to(cuda_dev)
    out = _attn(native, q, kv, C, idx).float()
    assert torch.isfinite(out).all()
    for s in range(S):
        v = kv[int(idx[s]) * T_PAD:int(idx[s]) * T_PAD + T_VALID, C:].float()
        lo, hi = v.min(0).values, v.max(0).values
        tol = 2e-2 * (hi - lo) + 1e-2  # bf16 P and output rounding
        assert (out[s] >= lo - tol).all() and (out[s] <= hi + tol).all()


def test_k3_key_permutation_invariance(native, cuda_dev):
    '''Permuting the 77 keys (K and V rows together) only changes the summation order.'''
    S, n_q, C = 2, 4096, 320
    g = torch.Generator().manual_seed(7)
    q = torch.randn(S, n_q, C, generator=g).to(cuda_dev).bfloat16()
    kv = torch.zeros(T_PAD, 2 * C)
    kv[:T_VALID] = torch.randn(T_VALID, 2 * C, generator=g)
    perm = torch.randperm(T_VALID, generator=g)
    kv2 = kv.clone()
    kv2[:T_VALID] = kv[perm]
    idx = torch.zeros(S, dtype=torch.int32, device=cuda_dev)
    a = _attn(native, q, kv.to(cuda_dev).bfloat16(), C, idx).float()
    b = _attn(native, q, kv2.to(cuda_dev).bfloat16(), C, idx).float()
    torch.testing.assert_close(a, b, rtol=2e-2, atol=2e-2)


def test_k2_rows_depend_only_on_their_own_context_row(native, cuda_dev):
    '''The K/V projection is row-wise: equal context rows give bit-equal cache rows wherever they
    sit in the 720-row batch (9 contexts, all 16 layers' weights), zero rows give zeros.'''
    g = torch.Generator().manual_seed(3)
    N, K = 24960, 768
    w = (torch.randn(N, K, generator=g) * 0.05).to(cuda_dev).bfloat16()
    ctx = torch.randn(9 * T_PAD, K, generator=g)
    ctx[5] = ctx[700]
    ctx[129] = ctx[700]
    ctx[300:303] = 0
    ctx = ctx.to(cuda_dev).bfloat16()
    kv = native.kv_project(ctx, w)
    torch.cuda.synchronize()
    assert torch.equal(kv[5], kv[700]) and torch.equal(kv[129], kv[700])
    assert (kv[300:303] == 0).all()
    ref = ctx[700].float() @ w.float().t()
    torch.testing.assert_close(kv[700].float(), ref, rtol=2e-2, atol=2e-2)


def test_k1_every_output_row_is_text_guide_or_between(native, cuda_dev):
    '''1024 prompts x one 257-token guide, default parameters: whatever the similarity values are,
    a blended row is bit-equal to its text row (weight 0), bit-equal to its mapped guide row
    (weight >= 1 - s), or the lerp  base + (alt - base) * w  with 0 < w <= max_guidance -- checked
    with the kernel's own map / weights, for every row of every prompt.'''
    B, T, A, D = 1024, 77, 257, 768
    g = torch.Generator(device=cuda_dev).manual_seed(11)
    txt = torch.randn(B, T, D, device=cuda_dev, generator=g)
    img = torch.randn(1, A, D, device=cuda_dev, generator=g)
    img[0, 40:60] = txt[3, 5:25] + 0.05 * torch.randn(20, D, device=cuda_dev, generator=g)  # some real matches
    prm = native.TweenParams()
    prm.threshold_floor = prm.threshold_mult = prm.max_guidance = 0.5
    prm.clustered, prm.header_max, prm.align_mode, prm.mapping_reuse = 0.0, 0.15, 1, 1
    lin = torch.linspace(0.0, 0.5, T)[None].to(cuda_dev)
    res = native.sim_blend(txt, img, [prm], lin)
    torch.cuda.synchronize()
    assert int(res['status'].max()) == 0
    out, idx, w = res['out'][:, 0], res['map_idx'][:, 0].long(), res['weights'][:, 0]
    assert int(idx.min()) >= 0 and int(idx.max()) < A
    alt = img[0][idx]                                   # [B, T, D] mapped guide rows
    iw = torch.minimum(w, torch.tensor(0.5, device=cuda_dev))[..., None]
    lerp = txt + (alt - txt) * iw                       # same fp32 expression as guidance.py:271
    is_txt = (out == txt).all(-1)
    is_alt = (out == alt).all(-1)
    is_lerp = (out == lerp).all(-1)
    assert bool((is_txt | is_alt | is_lerp).all())
    assert bool((is_txt | (w != 0)).all())              # weight 0 <=> text row kept
    assert int(is_lerp.sum()) > 0 and int(is_txt.sum()) > 0
