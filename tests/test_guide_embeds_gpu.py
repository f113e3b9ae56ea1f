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


def native_mod():
    from flexdiffuse_b200 import _native
    return _native


@pytest.fixture(scope='module')
def guide(native, cuda_dev):
    clip = tiny_clip().to(cuda_dev)
    return G.Guide(clip, FakeTok(), device=str(cuda_dev))


def _assert_exact(native, dev, got, txt, alt, **kw):
    '''Every row bit-identical to what the oracle derives from the kernel's own similarity matrix
    (itself within tolerance of the oracle's): no fraction-of-rows allowance.'''
    want = kc.expected_from_kernel_P(native, dev, txt, alt, orc.TweenParams(**kw))
    assert want is not None, 'oracle raises ZeroDivisionError on this case'
    assert torch.equal(got.cpu(), want)
    return want


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
def test_text_plus_image_guide(guide, native, cuda_dev, mode, reuse):
    img = _images()[0]
    kw = dict(guide_clustered=0.0, guide_mode=mode, guide_reuse=reuse,
              guide_threshold_floor=0.02, guide_linear=(0.1, 0.5),
              guide_max_guidance=0.35)
    with torch.no_grad():
        out = guide.embeds('a photo of a cat', img, **kw)
        txt = guide.encoder.prompt('a photo of a cat')
        gi = guide.encoder.image(img)
    assert out.shape == txt.shape and out.dtype == txt.dtype
    assert out.data_ptr() != txt.data_ptr()  # a fresh tensor (guidance.py:258)
    want = _assert_exact(native, cuda_dev, out, txt, gi, threshold=(0.02, 0.5), linear=(0.1, 0.5),
                         clustered=0.0, max_guidance=0.35, align_mode=mode,
                         mapping_reuse=reuse)
    assert not torch.equal(want, txt.cpu())  # the guide really changed the context


def test_text_guide_string_and_batch_of_prompts(guide, native, cuda_dev):
    with torch.no_grad():
        out = guide.embeds(['a dog', 'two birds on a wire'], 'an oil painting',
                           guide_clustered=0.0)
        txt = guide.encoder.prompt(['a dog', 'two birds on a wire'])
        gt = guide.encoder.prompt('an oil painting')
    assert tuple(out.shape) == (2, 77, 64)
    _assert_exact(native, cuda_dev, out, txt, gt, clustered=0.0)  # solo path per prompt (Q5)


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
    # the table the oracle derives from the kernel's own similarity matrix: exact, every row
    res = kc.run_kernel(native_mod(), cuda_dev, txt, img,
                        [orc.TweenParams(clustered=0.0, mapping_reuse=False, align_mode=0)])
    torch.testing.assert_close(res['sim'][0], orc.similarity_matrix(img, txt, rowwise=False),
                               rtol=kc.SIM_RTOL, atol=kc.SIM_ATOL)
    want = orc.map_from_similarity(res['sim'][0].double().numpy(), 77, False, orc.GUIDE_ORDER_TEXT)
    np.testing.assert_array_equal(got, want)
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


# ---------------------------------------------------------------------------------------------
# D = 768 (the real CLIP ViT-L/14 widths: text 768, vision 1024 -> projection 768) with the
# Clustered term ON -- BASELINE.json configs[0] is Linear + Clustered + Threshold.
@pytest.fixture(scope='module')
def guide768(native, cuda_dev):
    from transformers import CLIPConfig, CLIPModel
    cfg = CLIPConfig(
        text_config=dict(hidden_size=768, intermediate_size=1024, num_hidden_layers=1,
                         num_attention_heads=12, vocab_size=1000, max_position_embeddings=77),
        vision_config=dict(hidden_size=1024, intermediate_size=1024, num_hidden_layers=1,
                           num_attention_heads=16, image_size=224, patch_size=14),
        projection_dim=768)
    torch.manual_seed(99)
    clip = CLIPModel(cfg).eval().requires_grad_(False).to(cuda_dev)
    return G.Guide(clip, FakeTok(), device=str(cuda_dev))


@pytest.mark.parametrize('clustered', [0.15, 0.5, 1.0])
def test_guide_embeds_d768_with_clustered(guide768, native, cuda_dev, clustered):
    img = _images()[1]
    prompt = 'a watercolor painting of a lighthouse at dusk with seagulls overhead'
    # Linear and Threshold kept below the weakest Clustered strength so the cluster weights decide
    kw = dict(guide_clustered=clustered, guide_threshold_floor=0.02, guide_threshold_mult=0.1,
              guide_linear=(0.0, 0.1), guide_max_guidance=0.35)
    with torch.no_grad():
        out = guide768.embeds(prompt, img, **kw)
        txt = guide768.encoder.prompt(prompt)
        gi = guide768.encoder.image(img)
    assert tuple(out.shape) == (1, 77, 768) and tuple(gi.shape) == (1, 257, 768)
    prm = dict(threshold=(0.02, 0.1), linear=(0.0, 0.1), max_guidance=0.35)
    want = _assert_exact(native, cuda_dev, out, txt, gi, clustered=clustered, **prm)
    # the Clustered term took part: the same call without it gives another context
    flat = kc.expected_from_kernel_P(native, cuda_dev, txt, gi, orc.TweenParams(clustered=0.0, **prm))
    assert not torch.equal(want, flat)


def test_tweener_d768_clustered_zero_division_and_ok(native, cuda_dev):
    '''Planted (saturating) matches at D = 768 with every Clustered strength the UI offers: where the
    oracle raises ZeroDivisionError (two adjacent peaks, Q6) so does Tweener.tween; everywhere else
    the blend is bit-identical.  Both outcomes must occur.'''
    seen = dict(ok=0, zde=0)
    for seed in range(40, 52):
        txt, img = orc.synthetic_pair(seed, D=768, planted=10 + seed % 17)
        for clustered in (0.15, 0.5, 1.0):
            prm = orc.TweenParams(clustered=clustered, threshold=(0.5, 0.5))
            want = kc.expected_from_kernel_P(native, cuda_dev, txt, img, prm)
            tw = G.Tweener((0.5, 0.5), (0.0, 0.5), clustered)
            if want is None:
                with pytest.raises(ZeroDivisionError):
                    tw.tween(txt.to(cuda_dev), img.to(cuda_dev))
                seen['zde'] += 1
            else:
                got = tw.tween(txt.to(cuda_dev), img.to(cuda_dev))
                assert torch.equal(got.cpu(), want[0:1])
                seen['ok'] += 1
    assert seen['ok'] >= 3 and seen['zde'] >= 3, seen
