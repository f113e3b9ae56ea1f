'''`Guide.embeds` / `Tweener` / `ConceptMapper` / `_map_emb` drop-in behaviour on the GPU
(K1 through the Python mirror of guidance.py), every branch of guidance.py:392-474,
checked against the oracle restatement on the same encoder outputs.'''
import numpy as np
import pytest
import torch

from flexdiffuse_b200 import guidance as G
from oracle import guidance_oracle as orc
from tests import k1_common as kc
from tests.encode_helpers import FakeTok, test_images as _images, tiny_clip

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def guide(native, cuda_dev):
    clip = tiny_clip().to(cuda_dev)
    return G.Guide(clip, FakeTok(), device=str(cuda_dev))


def _oracle_blend(txt, img, **kw):
    prm = orc.TweenParams(**kw)
    try:
        return orc.tween_batch(txt.cpu(), img.cpu(), prm, rowwise=False)
    except ZeroDivisionError:
        return None


def _assert_rows_match(got, want, min_frac=0.97):
    '''bit-identical rows except provable near-tie flips.'''
    same = (got.cpu() == want).all(-1).float().mean().item()
    assert same >= min_frac, same


def test_text_only_returns_encoder_output(guide):
    with torch.no_grad():
        out = guide.embeds('a red fox')
        assert torch.equal(out, guide.encoder.prompt('a red fox'))
        two = guide.embeds(['a red fox', ' ', 'a blue bird'])  # blank entries dropped
    assert tuple(two.shape) == (2, 77, 64)


def test_errors(guide):
    with pytest.raises(ValueError):
        guide.embeds('', None)
    with pytest.raises(ValueError):
        guide.embeds(42)
    with pytest.raises(ValueError):
        guide.embeds(['  '], None)


@pytest.mark.parametrize('mode,reuse', [(1, True), (0, False), (2, True)])
def test_text_plus_image_guide(guide, mode, reuse):
    img = _images()[0]
    kw = dict(guide_clustered=0.0, guide_mode=mode, guide_reuse=reuse,
              guide_threshold_floor=0.02, guide_linear=(0.1, 0.5),
              guide_max_guidance=0.35)
    with torch.no_grad():
        out = guide.embeds('a photo of a cat', img, **kw)
        txt = guide.encoder.prompt('a photo of a cat')
        gi = guide.encoder.image(img)
    want = _oracle_blend(txt, gi, threshold=(0.02, 0.5), linear=(0.1, 0.5),
                         clustered=0.0, max_guidance=0.35, align_mode=mode,
                         mapping_reuse=reuse)
    assert out.shape == txt.shape and out.dtype == txt.dtype
    assert out.data_ptr() != txt.data_ptr()  # a fresh tensor (guidance.py:258)
    _assert_rows_match(out, want)


def test_text_guide_string_and_batch_of_prompts(guide):
    with torch.no_grad():
        out = guide.embeds(['a dog', 'two birds on a wire'], 'an oil painting',
                           guide_clustered=0.0)
        txt = guide.encoder.prompt(['a dog', 'two birds on a wire'])
        gt = guide.encoder.prompt('an oil painting')
    want = _oracle_blend(txt, gt, clustered=0.0)  # solo path per prompt (Q5)
    assert tuple(out.shape) == (2, 77, 64)
    _assert_rows_match(out, want)


def test_zero_division_error_is_raised_like_the_reference(guide, cuda_dev):
    '''adjacent similarity peaks: guidance.py:111-112 divides by zero (Q6).'''
    txt, img = orc.synthetic_pair(5023, A=77, D=64, planted=30)
    found = False
    for seed in range(5000, 5040):
        txt, img = orc.synthetic_pair(seed, D=64, planted=30)
        try:
            orc.tween(txt, img, orc.TweenParams(), rowwise=False)
        except ZeroDivisionError:
            found = True
            break
    assert found
    with pytest.raises(ZeroDivisionError):
        G.Tweener().tween(txt.to(cuda_dev), img.to(cuda_dev))


def test_concept_mapper(guide, cuda_dev):
    seed = 7001
    txt, img = orc.synthetic_pair(seed, D=64, planted=20)
    con, _ = orc.synthetic_pair(seed + 50, D=64, planted=0)
    con[0, 1:6] = txt[0, 3:8] * 1.25
    img[0, 40:45] = con[0, 1:6] * 0.8
    want = orc.concept_map(img, con, txt)
    cm = G.ConceptMapper(img.to(cuda_dev), con.to(cuda_dev))
    base = txt.to(cuda_dev)
    out = cm.map(base)
    assert out.data_ptr() != base.data_ptr()
    assert torch.equal(out.cpu(), want)
    assert not torch.equal(want, txt)
    target = base.clone()
    assert cm.map(base, target) is target  # edits the given tensor in place


def test_map_emb_host_table(guide, cuda_dev):
    txt, img = orc.synthetic_pair(11, D=64, planted=12)
    got = G._map_emb(img.to(cuda_dev), txt.to(cuda_dev), False, G.GUIDE_ORDER_TEXT)
    assert isinstance(got, np.ndarray) and got.shape == (77, 2) and got.dtype == np.float64
    want = orc.map_emb(img, txt, False, orc.GUIDE_ORDER_TEXT, rowwise=False)
    assert (got[:, 0] == want[:, 0]).mean() > 0.97
    np.testing.assert_allclose(got[:, 1], want[:, 1], rtol=2e-4, atol=2e-6)
    with pytest.raises(IndexError):  # 2-D input, as in the reference's broken batch path
        G._map_emb(img.to(cuda_dev), txt[0].to(cuda_dev))


def test_image_only_and_guide_string_only(guide, capsys):
    img = _images()[2]
    with torch.no_grad():
        out = guide.embeds('', img)
        gi = guide.encoder.image(img)
        ph = guide.placeholder_embed
    assert tuple(out.shape) == (1, 77, 64)
    want0 = gi[:, 0] + (ph[:, 0] - gi[:, 0]) * 0.85  # header pulled 85 % to the text header
    torch.testing.assert_close(out[:, 0], want0)
    torch.testing.assert_close(out[:, 1:], gi[:, 1:77])
    assert 'guide purely from image' in capsys.readouterr().out
    with torch.no_grad():
        s = guide.embeds('', 'just a string')
        assert torch.equal(s, guide.encoder.prompt('just a string'))
    assert 'just use prompt' in capsys.readouterr().out
