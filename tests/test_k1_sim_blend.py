'''K1 parity (GPU, through the C ABI) against the oracle restatement of
guidance.py `_map_emb` / `Tweener.tween`.'''
import numpy as np
import pytest
import torch

from oracle import guidance_oracle as orc
from tests import k1_common as kc

pytestmark = pytest.mark.gpu


def _cases(n, seed0, D):
    rs = np.random.RandomState(seed0)
    for i in range(n):
        planted = int(rs.choice([0, 12, 30]))
        yield (orc.synthetic_pair(seed0 * 100 + i, D=D, planted=planted),
               orc.random_params(rs))


@pytest.mark.parametrize('D', [64, 768])
def test_similarity_matrix_close_to_fp32(native, cuda_dev, D):
    '''(1) 3xTF32 tcgen05 GEMM + per-lane softmax vs the oracle's fp32 P.'''
    for (txt, img), prm in _cases(4, 11, D):
        res = kc.run_kernel(native, cuda_dev, txt, img, [prm])
        P = orc.similarity_matrix(img, txt, rowwise=False)
        torch.testing.assert_close(res['sim'][0], P, rtol=kc.SIM_RTOL,
                                   atol=kc.SIM_ATOL)
        assert torch.allclose(res['sim'][0].sum(-1), torch.ones(img.shape[1]),
                              atol=1e-5)


@pytest.mark.parametrize('D', [64, 768])
def test_decisions_and_blend_bit_exact_given_P(native, cuda_dev, D):
    '''(2) all modes x reuse x random parameters, incl. ZeroDivision cases.'''
    seen = {'ok': 0, 'zde': 0}
    for (txt, img), prm in _cases(24, 5, D):
        res = kc.run_kernel(native, cuda_dev, txt, img, [prm])
        seen[kc.check_exact_given_P(res, 0, 0, txt, img, prm)] += 1
    assert seen['ok'] >= 8 and seen['zde'] >= 1, seen


def test_every_mode_reuse_combo(native, cuda_dev):
    txt, img = orc.synthetic_pair(42, planted=12)
    prms = [
        orc.TweenParams(align_mode=m, mapping_reuse=r, clustered=0.0)
        for m in (0, 1, 2) for r in (True, False)
    ]
    res = kc.run_kernel(native, cuda_dev, txt, img, prms)
    for pi, prm in enumerate(prms):
        assert kc.check_exact_given_P(res, 0, pi, txt, img, prm) == 'ok'


def test_end_to_end_vs_oracle_default_params(native, cuda_dev):
    '''(3) against the oracle's own fp32 P: agreeing rows are bit-identical,
    disagreeing rows must be near-ties (relative margin < TIE_REL).'''
    n_rows = n_flip = 0
    for seed in range(6):
        txt, img = orc.synthetic_pair(900 + seed, planted=12)
        prm = orc.TweenParams(clustered=0.0)  # defaults minus the ZDE-prone part
        res = kc.run_kernel(native, cuda_dev, txt, img, [prm])
        out, mapped, w, sel, iw = orc.tween(txt, img, prm, return_parts=True)
        P = orc.similarity_matrix(img, txt).double().numpy()[:, 1:]
        for r in range(txt.shape[1]):
            n_rows += 1
            if torch.equal(res['out'][0, 0, r], out[0, r]):
                continue
            n_flip += 1
            # a flipped row must be explained by a near-tie in its column
            col = np.sort(P[:, r])[::-1] if r < P.shape[1] else np.zeros(2)
            tie = abs(col[0] - col[1]) <= kc.TIE_REL * max(col[0], 1e-30)
            edge = abs(abs(iw[r]) - (1.0 - mapped[r, 1])) <= kc.TIE_REL
            floor = abs(mapped[r, 1] - prm.threshold[0]) <= kc.TIE_REL
            assert tie or edge or floor, (seed, r, col[:2], iw[r], mapped[r])
    assert n_flip <= max(1, n_rows // 100), (n_flip, n_rows)


def test_batch_shared_guide_and_param_sweep(native, cuda_dev):
    '''B prompts x 1 guide x P parameter sets in one launch == per-pair runs
    (the batched semantics the reference intends).'''
    rs = np.random.RandomState(3)
    txts = []
    _, img = orc.synthetic_pair(77, D=64, planted=0)
    for b in range(5):
        t, _ = orc.synthetic_pair(200 + b, D=64, planted=0)
        t[0, 3 + b] = img[0, 10 * b + 1] * 1.5  # plant one match per prompt
        txts.append(t)
    txt = torch.cat(txts)
    prms = [orc.random_params(rs) for _ in range(7)]
    res = kc.run_kernel(native, cuda_dev, txt, img, prms)
    for b in range(5):
        for pi, prm in enumerate(prms):
            kc.check_exact_given_P(res, b, pi, txt[b:b + 1], img, prm)


def test_text_guide_A77_and_per_prompt_guides(native, cuda_dev):
    '''Guide given as text (A = 77) and one guide per prompt (guide_batch = B).'''
    a, _ = orc.synthetic_pair(1, A=77, D=64, planted=0)
    b, _ = orc.synthetic_pair(2, A=77, D=64, planted=0)
    txt = torch.cat([a, b])
    guide = torch.cat([b, a]).clone()
    guide[0, 7] = txt[0, 9]
    prm = orc.TweenParams(clustered=0.0, mapping_reuse=False)
    res = kc.run_kernel(native, cuda_dev, txt, guide, [prm])
    for i in range(2):
        kc.check_exact_given_P(res, i, 0, txt[i:i + 1], guide[i:i + 1], prm)


def test_identical_embeddings_saturate(native, cuda_dev):
    '''Q4/Q6 territory: guide == text gives saturated similarities and exact
    zeros after underflow; the kernel must follow the oracle decision for
    decision, or flag ZeroDivision when the oracle raises.'''
    txt, _ = orc.synthetic_pair(5, A=77, D=64, planted=0)
    for mode in (0, 1, 2):
        for reuse in (True, False):
            prm = orc.TweenParams(align_mode=mode, mapping_reuse=reuse)
            res = kc.run_kernel(native, cuda_dev, txt, txt.clone(), [prm])
            kc.check_exact_given_P(res, 0, 0, txt, txt, prm)


def test_slerp_extension(native, cuda_dev):
    '''FD_BLEND_MODE_SLERP: decisions (map, weights, select) stay bit-exact given the
    kernel's P; rows the reference lerps are spherically interpolated instead.  The reference
    has no slerp (parity unpinned): the check is oracle/guidance_oracle.py:slerp_rows, in
    float64, so the tolerance below is fp32 sin/acos error (<= 2e-5 relative to the row's
    largest element), and rows within 1e-4 of the |cos| = 0.9995 switch may use either form.'''
    txt, img = orc.synthetic_pair(11)
    prms = [orc.TweenParams(clustered=0.0, slerp=True),
            orc.TweenParams(clustered=0.0, slerp=False),
            orc.TweenParams(clustered=0.0, slerp=True, linear=(-0.3, 0.4),
                            threshold=(0.2, 0.3), align_mode=0,
                            mapping_reuse=False)]
    res = kc.run_kernel(native, cuda_dev, txt, img, prms)
    P = res['sim'][0]
    n_slerp = 0
    for p, prm in enumerate(prms):
        want = kc.oracle_from_P(P, txt, img, prm)
        assert not want['zde']
        assert np.array_equal(res['map_idx'][0, p].numpy(),
                              want['mapped'][:, 0].astype(np.int64))
        assert torch.equal(res['weights'][0, p], want['w'])
        got = res['out'][0, p]
        if not prm.slerp:
            assert torch.equal(got, want['out'])
            continue
        idx = torch.from_numpy(want['mapped'][:, 0].astype(np.int64))
        sl, use, cosv = orc.slerp_rows(txt[0], img[0, idx], want['iw'])
        for r in range(txt.shape[1]):
            if want['sel'][r] != 2:
                assert torch.equal(got[r], want['out'][r]), r
                continue
            scale = max(txt[0, r].abs().max().item(),
                        img[0, idx[r]].abs().max().item())
            d_sl = (got[r].double() - sl[r]).abs().max().item()
            d_lerp = (got[r].double() - want['out'][r].double()).abs().max().item()
            near = abs(abs(cosv[r]) - orc.SLERP_DOT_THRESHOLD) < 1e-4
            if use[r] or near:
                ok = d_sl <= 2e-5 * scale
                n_slerp += int(ok and use[r])
            if not use[r] or near:
                ok = (d_lerp == 0.0) or (near and ok)
            assert ok, (p, r, d_sl, d_lerp, cosv[r])
    assert n_slerp >= 5  # the mode really ran


def test_rejects_bad_shapes(native, cuda_dev):
    txt = torch.zeros(1, 77, 100, device=cuda_dev)  # D not a multiple of 32
    img = torch.zeros(1, 257, 100, device=cuda_dev)
    lin = torch.zeros(1, 77, device=cuda_dev)
    with pytest.raises(native.NativeError):
        native.sim_blend(txt, img, [native.TweenParams()], lin)
