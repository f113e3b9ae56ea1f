'''Encoder / composition rows of the path: flexdiffuse_b200.encode.clip and
composition.{schema,embeds} against golden vectors produced by the unmodified reference
(tests/golden/make_encode_golden.py).  These stay in PyTorch, so they run on CPU here.'''
import hashlib
import os

import numpy as np
import pytest
import torch

from flexdiffuse_b200.composition.embeds import encode_schema, px_to_block
from flexdiffuse_b200.composition.schema import EntitySchema, Schema
from flexdiffuse_b200.encode.clip import CLIPEncoder, preprocess
from tests.encode_helpers import FakeTok, test_images as _images, tiny_clip

GOLDEN = os.path.join(os.path.dirname(__file__), 'golden', 'encode_golden.npz')


@pytest.fixture(scope='module')
def setup():
    torch.set_num_threads(1)
    return np.load(GOLDEN), CLIPEncoder(tiny_clip(), FakeTok())


def test_preprocess_matches_reference(setup):
    gold, _ = setup
    for i, img in enumerate(_images()):
        pre = preprocess(img).numpy()
        assert tuple(pre.shape) == tuple(gold[f'pre{i}_shape'])
        assert max(pre.shape[-2:]) == 512 and min(pre.shape[-2:]) % 64 == 0
        sha = np.frombuffer(hashlib.sha256(pre.tobytes()).digest(), np.uint8)
        np.testing.assert_array_equal(sha, gold[f'pre{i}_sha'])
        np.testing.assert_array_equal(pre[0, :, :4, :4], gold[f'pre{i}_corner'])


def test_image_encoder_all_tokens_projected(setup):
    gold, enc = setup
    for i, img in enumerate(_images()):
        with torch.no_grad():
            out = enc.image(img)
        assert tuple(out.shape) == (1, 257, 64)  # all 257 tokens, projected
        np.testing.assert_allclose(out.numpy(), gold[f'img{i}'], rtol=1e-5,
                                   atol=1e-6)


def test_prompt_encoder_unprojected_hidden_states(setup):
    gold, enc = setup
    with torch.no_grad():
        a = enc.prompt('a photo of a cat')
        b = enc.prompt(['a dog', 'two birds on a wire'])
    assert tuple(a.shape) == (1, 77, 64) and tuple(b.shape) == (2, 77, 64)
    np.testing.assert_allclose(a.numpy(), gold['txt1'], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(b.numpy(), gold['txt2'], rtol=1e-5, atol=1e-6)


def test_schema_and_embeds(setup):
    gold, enc = setup
    s = Schema('bg prompt', 'style a', 'style b', (0.1, 0.9),
               [EntitySchema('thing one', (64, 128), (256, 200), 0.7)])
    assert s.json().encode() == gold['schema_json'].tobytes()
    with torch.no_grad():
        e = encode_schema(s, enc)
    np.testing.assert_allclose(e.background_embed.numpy(), gold['schema_bg'],
                               rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(e.entities[0].embed.numpy(), gold['schema_ent'],
                               rtol=1e-5, atol=1e-6)
    assert list(e.entities[0].offset_blocks) + list(
        e.entities[0].size_blocks) == list(gold['schema_blocks'])
    assert e.entities[0].blend == 0.7 and e.style_blend == (0.1, 0.9)
    assert list(px_to_block((512, 257, 7))) == list(gold['px']) == [64, 32, 0]
    assert EntitySchema('x', (0, 0), (8, 8)).blend == 0.8
