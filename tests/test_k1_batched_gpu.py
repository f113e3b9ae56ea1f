'''K1B, the batched kernel of `fd_sim_blend` (many prompts x one shared guide, mappings with reuse or
DIRECT): guide tokens on the TMEM lanes, softmax in registers, redux arg-max, persistent CTAs.

Same two exact statements as tests/test_k1_sim_blend.py: (1) its similarity matrix P is within
SIM_RTOL / SIM_ATOL of the oracle's fp32 P; (2) every decision and every output bit it derives from
ITS P equals what the oracle derives from the same P.  Plus: it agrees with the one-prompt-per-CTA
kernel wherever both kernels' P lead to the same decisions.'''
import numpy as np
import pytest
import torch

from oracle import guidance_oracle as orc
from tests import k1_common as kc

pytestmark = pytest.mark.gpu


@pytest.fixture()
def fast(native):
    '''Force the batched kernel even for a handful of prompts; restore the default afterwards.'''
    lib = native.lib()
    lib.fd_debug_set_k1_fast(1, 1)
    yield lib
    lib.fd_debug_set_k1_fast(1, 64)


def _prompts(n, seed, A=257, D=768, planted=10):
    _, img = orc.synthetic_pair(seed, A=A, D=D, planted=0)
    txts = []
    rs = np.random.RandomState(seed)
    for b in range(n):
        t, _ = orc.synthetic_pair(seed + 1 + b, A=A, D=D, planted=0)
        for k in range(planted):
            tok = 2 + (5 * k + 3 * b) % 74
            row = (11 * k + 7 * b) % A
            t[0, tok] = img[0, row] * float(rs.uniform(0.6, 1.8)) + float(rs.uniform(0, 0.5)) * torch.randn(D)
        txts.append(t)
    return torch.cat(txts), img


@pytest.mark.parametrize('n,A', [(1, 257), (2, 257), (3, 257), (7, 256), (5, 240)])
def test_batched_kernel_exact_on_its_own_P(native, cuda_dev, fast, n, A):
    txt, img = _prompts(n, 100 + n + A, A=A)
    rs = np.random.RandomState(n)
    prms = []
    while len(prms) < 6:
        q = orc.random_params(rs)
        if q.align_mode == 2 or q.mapping_reuse:  # what the batched kernel serves
            prms.append(q)
    prms.append(orc.TweenParams())  # reference defaults (Clustered 0.5)
    res = kc.run_kernel(native, cuda_dev, txt, img, prms)
    seen = set()
    for b in range(n):
        P = orc.similarity_matrix(img, txt[b:b + 1], rowwise=False)
        torch.testing.assert_close(res['sim'][b], P, rtol=kc.SIM_RTOL, atol=kc.SIM_ATOL)
        for pi, prm in enumerate(prms):
            seen.add(kc.check_exact_given_P(res, b, pi, txt[b:b + 1], img, prm))
    assert 'ok' in seen


def test_batched_equals_per_prompt_kernel(native, cuda_dev):
    '''The two kernels compute P in different orders (different MMA orientation), so P differs in the
    last bits; wherever that does not flip a decision the blended rows are bit-identical, and the
    fraction of rows that differ is bounded by the near-tie rate (<= 1 %).'''
    lib = native.lib()
    txt, img = _prompts(9, 321)
    prm = orc.TweenParams(clustered=0.0)
    try:
        lib.fd_debug_set_k1_fast(0, 0)
        slow = kc.run_kernel(native, cuda_dev, txt, img, [prm])
        lib.fd_debug_set_k1_fast(1, 1)
        quick = kc.run_kernel(native, cuda_dev, txt, img, [prm])
    finally:
        lib.fd_debug_set_k1_fast(1, 64)
    torch.testing.assert_close(quick['sim'], slow['sim'], rtol=kc.SIM_RTOL, atol=kc.SIM_ATOL)
    same = (quick['out'] == slow['out']).all(-1).float().mean().item()
    assert same >= 0.99, same
    assert (quick['map_idx'] == slow['map_idx']).float().mean().item() >= 0.99


def test_batched_large_launch_matches_oracle(native, cuda_dev):
    '''300 prompts (default threshold: the batched kernel is chosen without forcing): persistent CTAs,
    odd tail batch, sampled prompts checked exactly.'''
    n = 301
    txt, img = _prompts(n, 777, planted=6)
    prm = orc.TweenParams(clustered=0.0, threshold=(0.3, 0.5), linear=(0.1, 0.5), max_guidance=0.35)
    res = kc.run_kernel(native, cuda_dev, txt, img, [prm])
    for b in (0, 1, 2, 147, 148, 149, 298, 299, 300):
        assert kc.check_exact_given_P(res, b, 0, txt[b:b + 1], img, prm) == 'ok'
    assert torch.isfinite(res['out']).all()
    # no-sim call (the product call) gives the same blend
    lin = torch.linspace(prm.linear[0], prm.linear[1], steps=77)[None].to(cuda_dev)
    out = native.sim_blend(txt.to(cuda_dev), img.to(cuda_dev), [kc.to_native_params(native, prm)], lin)['out']
    assert torch.equal(out.cpu(), res['out'])


def test_batched_zero_division_slerp_and_range_flags(native, cuda_dev, fast):
    # ZeroDivision outcome (Q6) must be flagged exactly where the oracle raises
    hits = 0
    for seed in range(5000, 5030):
        txt, img = orc.synthetic_pair(seed, planted=30)
        prm = orc.TweenParams()
        res = kc.run_kernel(native, cuda_dev, txt, img, [prm])
        hits += kc.check_exact_given_P(res, 0, 0, txt, img, prm) == 'zde'
    assert hits >= 1
    # slerp rows follow the same decisions
    txt, img = _prompts(2, 55)
    prms = [orc.TweenParams(clustered=0.0, slerp=True), orc.TweenParams(clustered=0.0)]
    res = kc.run_kernel(native, cuda_dev, txt, img, prms)
    for b in range(2):
        want = kc.oracle_from_P(res['sim'][b], txt[b:b + 1], img, prms[0])
        assert torch.equal(res['weights'][b, 0], want['w'])
        lerp_rows = torch.from_numpy(want['sel'] != 2)
        assert torch.equal(res['out'][b, 0][lerp_rows], res['out'][b, 1][lerp_rows])
    # an out-of-range prompt is flagged, its neighbour in the same batch is not
    txt2 = txt.clone()
    txt2[1, 5, 17] = 5000.0
    res = kc.run_kernel(native, cuda_dev, txt2, img, [orc.TweenParams(clustered=0.0)])
    assert int(res['status'][0, 0]) == 0 and int(res['status'][1, 0]) == native.FD_BLEND_RANGE
