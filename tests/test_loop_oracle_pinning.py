'''Pin oracle/loop_oracle.py (the restated flex.py / guide.py control flow) to the
UNMODIFIED reference loop running on oracle/diffusers_shim.  CPU only; needs
/root/reference, i.e. runs in the build container and is skipped on the GPU box.'''
import os
import sys

import pytest
import torch

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(REF),
                                reason='reference only mounted in the build container')


class ToyUNet(torch.nn.Module):
    '''Cheap deterministic stand-in with the attribute surface guide.py / flex.py use.'''
    in_channels = 4
    config = {'attention_head_dim': 8}

    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(1)
        self.w = torch.nn.Parameter(torch.randn(4, 4, 3, 3, generator=g) * 0.2)
        self.c = torch.nn.Parameter(torch.randn(768, 4, generator=g) * 0.05)

    def forward(self, x, t, encoder_hidden_states):
        from types import SimpleNamespace
        bias = (encoder_hidden_states.mean(1) @ self.c)[:, :, None, None]
        tt = torch.as_tensor(t, dtype=torch.float32) / 1000.0
        y = torch.nn.functional.conv2d(x, self.w, padding=1) + bias * (1 + tt)
        return SimpleNamespace(sample=torch.tanh(y))


class FakeEncoder:
    def __init__(self):
        g = torch.Generator().manual_seed(2)
        self.u = torch.randn(1, 77, 768, generator=g)

    def prompt(self, p):
        return self.u.clone()


class ToyVAE(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))

    def decode(self, z):
        from types import SimpleNamespace
        return SimpleNamespace(sample=torch.nn.functional.interpolate(
            z[:, :3], scale_factor=8.0))

    def encode(self, x):
        from types import SimpleNamespace
        lat = torch.nn.functional.avg_pool2d(x, 8)
        lat = torch.cat([lat, lat[:, :1]], 1)

        class D:
            def sample(self, generator=None):
                return lat + 0.1 * torch.randn(lat.shape, generator=generator)

        return SimpleNamespace(latent_dist=D())


@pytest.fixture(scope='module')
def ref_mods():
    from oracle import loop_oracle  # puts the shim on sys.path
    sys.path.insert(0, REF)
    try:
        import pipeline.flex as rflex
        import pipeline.guide as rguide
    finally:
        sys.path.remove(REF)
    return rflex, rguide, loop_oracle


@pytest.mark.parametrize('sched_name', ['DDIMScheduler', 'PNDMScheduler',
                                        'LMSDiscreteScheduler'])
@pytest.mark.parametrize('img2img', [False, True])
def test_restated_loop_equals_reference_loop(ref_mods, sched_name, img2img):
    rflex, rguide, lo = ref_mods
    torch.manual_seed(0)
    unet, enc, vae = ToyUNet(), FakeEncoder(), ToyVAE()
    embeds = torch.randn(2, 77, 768, generator=torch.Generator().manual_seed(3))
    steps, g = 12, 7.5
    # reference
    sched = getattr(lo, sched_name)()
    pipe = rflex.FlexPipeline(vae, None, None, unet, sched)
    guide = rguide.SimpleGuide(enc, unet, g, steps, embeds)
    gen = torch.Generator().manual_seed(11)
    init = None
    if img2img:
        init = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(5)) * 2 - 1
    # capture the reference's final latents by wrapping the decode
    got = {}
    orig = pipe._latents_to_image
    pipe._latents_to_image = lambda l, pil=True: (got.setdefault('lat', l), orig(l, False))[1]
    kw = dict(init_size=(64, 64), generator=gen, output_type='np', eta=0.3)
    if img2img:
        # flex.py:181 `if init_image:` needs a truthy object; a PIL image would be resized
        # to 512, so hand it a 1-element-truthiness-safe wrapper
        class Img(torch.Tensor):
            def __bool__(self):
                return True
        kw['init_image'] = init.as_subclass(Img)
    torch.manual_seed(123)  # DDIM eta>0 draws its variance noise from the global RNG
    with torch.no_grad():
        pipe(guide, **kw)
    # restatement
    sched2 = getattr(lo, sched_name)()
    gen2 = torch.Generator().manual_seed(11)
    init_lat = None
    if img2img:
        init_lat = 0.18215 * vae.encode(init).latent_dist.sample(generator=gen2)
    torch.manual_seed(123)
    with torch.no_grad():
        lat = lo.denoise(lambda x, t, c: unet(x, t, c).sample, sched2, enc.u,
                         embeds, g, steps, init_latents=init_lat,
                         init_size=(64, 64), strength=0.6, eta=0.3,
                         generator=gen2)
    assert torch.equal(lat, got['lat'])


def test_composite_restatement_equals_reference(ref_mods):
    '''oracle composite_noise_pred == the UNMODIFIED reference CompositeGuide.noise_pred.'''
    rflex, rguide, lo = ref_mods
    sys.path.insert(0, REF)
    try:
        import composition.guide as rcomp
        from composition.schema import EntitySchema, Schema
    finally:
        sys.path.remove(REF)

    class Enc:
        def __init__(self):
            self.g = torch.Generator().manual_seed(9)
            self.cache = {}

        def prompt(self, p):
            if p not in self.cache:
                self.cache[p] = torch.randn(1, 77, 768, generator=self.g)
            return self.cache[p]

    enc, unet = Enc(), ToyUNet()
    schema = Schema('a meadow', 'oil', 'ink', (0.0, 1.0), [
        EntitySchema('a bear', (16, 8), (24, 32), 0.8),
        EntitySchema('a hat', (24, 0), (16, 16), 0.5),
    ])
    x = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(1))
    for g in (7.5, 1.0):
        guide = rcomp.CompositeGuide(enc, unet, g, schema, 20)
        with torch.no_grad():
            want = guide.noise_pred(x.clone(), torch.tensor(481))
            ents = [(enc.prompt(e.prompt), tuple(v // 8 for v in e.offset),
                     tuple(v // 8 for v in e.size), e.blend) for e in schema.entities]
            got = lo.composite_noise_pred(lambda l, t, c: unet(l, t, c).sample, enc.prompt(''),
                                          enc.prompt('a meadow'), ents, g, x.clone(),
                                          torch.tensor(481))
        assert torch.equal(got, want), g
