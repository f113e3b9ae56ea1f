'''Seeded tiny CLIP + images + tokenizer shared by the encoder golden generator and tests.'''
import numpy as np
import torch


class FakeTok:
    model_max_length = 77

    def __call__(self, prompt, padding='max_length', max_length=77,
                 truncation=True, return_tensors='pt'):
        from types import SimpleNamespace
        if isinstance(prompt, str):
            prompt = [prompt]
        rows = []
        for p in prompt:
            ids = [998] + [3 + (sum(map(ord, w)) % 990) for w in p.split()]
            ids = ids[:76] + [999]
            rows.append(ids + [999] * (77 - len(ids)))
        return SimpleNamespace(input_ids=torch.tensor(rows))


def tiny_clip():
    from transformers import CLIPConfig, CLIPModel
    cfg = CLIPConfig(
        text_config=dict(hidden_size=64, intermediate_size=128,
                         num_hidden_layers=2, num_attention_heads=4,
                         vocab_size=1000, max_position_embeddings=77),
        vision_config=dict(hidden_size=96, intermediate_size=128,
                           num_hidden_layers=2, num_attention_heads=4,
                           image_size=224, patch_size=14),
        projection_dim=64)
    torch.manual_seed(4321)
    return CLIPModel(cfg).eval().requires_grad_(False)


def test_images():
    from PIL import Image
    rs = np.random.RandomState(9)
    out = []
    for w, h in ((300, 200), (200, 333), (128, 128)):
        out.append(Image.fromarray((rs.rand(h, w, 3) * 255).astype('uint8')))
    return out
