'''Golden vectors for the encoder row of the path: produced by importing the
UNMODIFIED reference `encode/clip.py` and `composition/{schema,embeds}.py` in the build
container.  Inputs are seeded; a tiny random-init CLIP keeps the fixture small.

    python tests/golden/make_encode_golden.py  ->  tests/golden/encode_golden.npz
'''
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

import encode.clip as ref_clip  # noqa: E402  (the reference, unmodified)
from composition.schema import EntitySchema, Schema  # noqa: E402
from composition.embeds import encode_schema, px_to_block  # noqa: E402

from tests.encode_helpers import tiny_clip, test_images, FakeTok  # noqa: E402


def main():
    torch.set_num_threads(1)
    store = {}
    clip = tiny_clip()
    enc = ref_clip.CLIPEncoder(clip, FakeTok())
    for i, img in enumerate(test_images()):
        pre = ref_clip.preprocess(img).numpy()
        store[f'pre{i}_shape'] = np.array(pre.shape)
        store[f'pre{i}_sha'] = np.frombuffer(
            hashlib.sha256(pre.tobytes()).digest(), dtype=np.uint8)
        store[f'pre{i}_corner'] = pre[0, :, :4, :4]
        with torch.no_grad():
            store[f'img{i}'] = enc.image(img).numpy()
    with torch.no_grad():
        store['txt1'] = enc.prompt('a photo of a cat').numpy()
        store['txt2'] = enc.prompt(['a dog', 'two birds on a wire']).numpy()
        s = Schema('bg prompt', 'style a', 'style b', (0.1, 0.9),
                   [EntitySchema('thing one', (64, 128), (256, 200), 0.7)])
        e = encode_schema(s, enc)
    store['schema_json'] = np.frombuffer(s.json().encode(), dtype=np.uint8)
    store['schema_bg'] = e.background_embed.numpy()
    store['schema_ent'] = e.entities[0].embed.numpy()
    store['schema_blocks'] = np.array(
        list(e.entities[0].offset_blocks) + list(e.entities[0].size_blocks))
    store['px'] = np.array(px_to_block((512, 257, 7)))
    path = os.path.join(HERE, 'encode_golden.npz')
    np.savez_compressed(path, **store)
    print('wrote', path, os.path.getsize(path) / 1e6, 'MB')


if __name__ == '__main__':
    main()
