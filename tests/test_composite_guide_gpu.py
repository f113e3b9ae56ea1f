'''CompositeGuide (composition/guide.py mirror: K2 cache over uncond + background + entity
contexts, K3, regional blend in K9, CFG / scheduler in K4) against the oracle restatement
(pinned to the reference in tests/test_loop_oracle_pinning.py) on the same random-init
weights.  bf16 tolerance as for the plain guide.'''
import pytest
import torch

from flexdiffuse_b200 import schedulers as prod
from flexdiffuse_b200.composition.guide import CompositeGuide
from flexdiffuse_b200.composition.schema import EntitySchema, Schema
from flexdiffuse_b200.pipeline.flex import FlexPipeline
from oracle import loop_oracle as lo
from oracle import unet_oracle as U
from tests.model_helpers import models, rel_l2

pytestmark = pytest.mark.gpu


class Enc:
    '''prompt -> seeded embedding (stands in for CLIPEncoder.prompt).'''
    def __init__(self, dev):
        self.dev, self.cache = dev, {}

    def prompt(self, p):
        if isinstance(p, (list, tuple)):  # encode_schema batches every prompt into one call
            return torch.cat([self.prompt(q) for q in p])
        if p not in self.cache:
            g = torch.Generator(device=self.dev).manual_seed(len(self.cache) + 3)
            self.cache[p] = torch.randn(1, 77, 768, device=self.dev, generator=g)
        return self.cache[p]


SCHEMA = Schema('a meadow', 'oil', 'ink', (0.0, 1.0), [
    EntitySchema('a bear', (64, 32), (96, 128), 0.8),
    EntitySchema('a hat', (96, 0), (64, 64), 0.5),
    EntitySchema('off canvas', (224, 224), (128, 128), 0.3),  # clipped at the border
])


def test_k9_matches_reference_arithmetic(native, cuda_dev):
    g = torch.Generator(device=cuda_dev).manual_seed(0)
    eps = torch.randn(5, 4, 32, 32, device=cuda_dev, generator=g)
    boxes = []
    for (ox, oy, sx, sy, bl) in [(8, 4, 12, 16, 0.8), (12, 0, 8, 8, 0.5), (28, 28, 16, 16, 0.3)]:
        b = native.EntityBox()
        b.ox, b.oy, b.sx, b.sy, b.blend = ox, oy, sx, sy, bl
        boxes.append(b)
    for dt in (torch.float32, torch.bfloat16):
        e = eps.to(dt)
        u, c = native.composite_eps(e, boxes)
        want = e[1:2].float().clone()
        for i, b in enumerate(boxes):
            sl = (slice(None), slice(None), slice(b.oy, b.oy + b.sy), slice(b.ox, b.ox + b.sx))
            want[sl] = want[sl] + b.blend * (e[2 + i:3 + i].float()[sl] - want[sl])
        torch.testing.assert_close(c, want, rtol=1e-6, atol=1e-6)
        assert torch.equal(u, e[:1].float())


@pytest.mark.parametrize('graph', [False, True])
@pytest.mark.parametrize('guidance', [7.5, 1.0])
def test_composite_noise_pred_and_pipeline(native, cuda_dev, graph, guidance):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    unet, vae, usd, _ = models(str(cuda_dev))
    enc = Enc(cuda_dev)
    guide = CompositeGuide(enc, unet, guidance, SCHEMA, 4, use_cuda_graph=graph)
    x = torch.randn(1, 4, 32, 32, device=cuda_dev)
    keep = x.clone()
    got = guide.noise_pred(x, 481)
    assert torch.equal(x, keep)
    f = lambda l, t, c: U.unet_forward(usd, l.bfloat16().float(), t, c.bfloat16().float())
    ents = [(enc.prompt(e.prompt), tuple(v // 8 for v in e.offset),
             tuple(v // 8 for v in e.size), e.blend) for e in SCHEMA.entities]
    want = lo.composite_noise_pred(f, enc.prompt(''), enc.prompt('a meadow'), ents, guidance,
                                   x, 481)
    assert rel_l2(got, want) < (1e-1 if guidance > 1 else 3e-2)
    # and through the pipeline (fused K4 path), 4 DDIM steps
    pipe = FlexPipeline(vae, None, None, unet, prod.DDIMScheduler())
    lat = pipe(guide, init_size=(256, 256),
               generator=torch.Generator(device=cuda_dev).manual_seed(3),
               output_type='latent', return_dict=False)
    assert tuple(lat.shape) == (1, 4, 32, 32) and torch.isfinite(lat).all()


def test_batch_size_other_than_one_is_rejected(native, cuda_dev):
    unet, _, _, _ = models(str(cuda_dev))
    with pytest.raises(ValueError):
        CompositeGuide(Enc(cuda_dev), unet, 7.5, SCHEMA, 4, batch_size=2)
