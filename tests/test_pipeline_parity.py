'''End-to-end parity (GPU): flexdiffuse_b200's FlexPipeline / SimpleGuide (bf16 UNet, cached
K/V, K3, fused K4, optional CUDA graph) against the oracle loop (oracle/loop_oracle.py,
pinned to the reference's flex.py/guide.py control flow) with the fp32 oracle UNet/VAE on
the same random-init weights, same initial latents.

Stated tolerances: final latents relative L2 <= 5e-2; decoded images PSNR >= 30 dB.'''
import pytest
import torch

from flexdiffuse_b200 import schedulers as prod
from flexdiffuse_b200.pipeline.flex import FlexPipeline
from flexdiffuse_b200.pipeline.guide import SimpleGuide
from oracle import loop_oracle as lo
from oracle import unet_oracle as U
from tests.model_helpers import models, psnr, rel_l2

pytestmark = pytest.mark.gpu

LATENT_TOL = 5e-2
PSNR_MIN = 30.0


class _Enc:
    def __init__(self, uncond):
        self.uncond = uncond

    def prompt(self, p):
        return self.uncond


def _setup(dev, B, seed):
    unet, vae, usd, vsd = models(str(dev))
    g = torch.Generator(device=dev).manual_seed(seed)
    uncond = torch.randn(1, 77, 768, device=dev, generator=g)
    embeds = torch.randn(B, 77, 768, device=dev, generator=g)
    return unet, vae, usd, vsd, uncond, embeds


@pytest.mark.parametrize('name,steps,graph', [('DDIMScheduler', 10, False),
                                              ('DDIMScheduler', 10, True),
                                              ('PNDMScheduler', 8, False),
                                              ('LMSDiscreteScheduler', 8, True)])
def test_txt2img_latents_and_images(native, cuda_dev, name, steps, graph):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    B, hw = 2, 256
    unet, vae, usd, vsd, uncond, embeds = _setup(cuda_dev, B, 5)
    pipe = FlexPipeline(vae, None, None, unet, getattr(prod, name)())
    guide = SimpleGuide(_Enc(uncond), unet, 7.5, steps, embeds,
                        use_cuda_graph=graph)
    gen = torch.Generator(device=cuda_dev).manual_seed(77)
    lat = pipe(guide, init_size=(hw, hw), generator=gen, output_type='latent',
               return_dict=False)
    gen2 = torch.Generator(device=cuda_dev).manual_seed(77)
    want = lo.denoise(
        lambda x, t, c: U.unet_forward(usd, x, t, c.bfloat16().float()),
        getattr(lo, name)(), uncond, embeds, 7.5, steps, init_size=(hw, hw),
        generator=gen2, device=cuda_dev)
    err = rel_l2(lat, want)
    assert err < LATENT_TOL, err
    img = (vae.decode(lat / 0.18215).sample.float() / 2 + 0.5).clamp(0, 1)
    ref = (U.vae_decode(vsd, want / 0.18215) / 2 + 0.5).clamp(0, 1)
    assert psnr(img, ref) >= PSNR_MIN, psnr(img, ref)


def test_reference_call_form_noise_pred(native, cuda_dev):
    '''guide.noise_pred(latents, t) (guide.py:46-64 form) returns the CFG-combined eps and
    leaves `latents` untouched.  The un-combined halves are held to the per-step bf16
    tolerance (3e-2); the combine u + 7.5 (c - u) amplifies their difference, so the combined
    prediction is checked (a) bit-for-bit as K4's CFG of the halves and (b) within 1e-1.'''
    unet, vae, usd, _, uncond, embeds = _setup(cuda_dev, 1, 6)
    guide = SimpleGuide(_Enc(uncond), unet, 7.5, 10, embeds)
    x = torch.randn(1, 4, 32, 32, device=cuda_dev)
    keep = x.clone()
    eps = guide.noise_pred(x, 481)
    assert torch.equal(x, keep)
    u, c = guide.noise_pred_pair(x, 481)
    f = lambda l, t, cc: U.unet_forward(usd, l.bfloat16().float(), t,
                                        cc.bfloat16().float())
    assert rel_l2(u, f(x, 481, uncond)) < 3e-2
    assert rel_l2(c, f(x, 481, embeds)) < 3e-2
    torch.testing.assert_close(eps, u.float() + 7.5 * (c.float() - u.float()),
                               rtol=1e-5, atol=1e-5)
    want = lo.noise_pred(f, uncond, embeds, 7.5, x, 481)
    assert rel_l2(eps, want) < 1e-1


def test_img2img_strength_and_pil_output(native, cuda_dev):
    '''init image path: VAE encode -> add_noise -> t_start slice (flex.py:181-221).'''
    from PIL import Image
    import numpy as np
    unet, vae, _, _, uncond, embeds = _setup(cuda_dev, 1, 8)
    pipe = FlexPipeline(vae, None, None, unet, prod.PNDMScheduler())
    guide = SimpleGuide(_Enc(uncond), unet, 7.5, 10, embeds)
    rs = np.random.RandomState(0)
    init = Image.fromarray((rs.rand(256, 256, 3) * 255).astype('uint8'))
    calls = []
    orig = guide.noise_pred_pair
    guide.noise_pred_pair = lambda l, t: (calls.append(t), orig(l, t))[1]
    out = pipe(guide, init_image=init, strength=0.6,
               generator=torch.Generator(device=cuda_dev).manual_seed(1))
    # steps=10, strength .6 -> init_timestep 6, t_start 4; PNDM list has 11 entries -> 7 evals
    assert len(calls) == 7
    assert len(out.images) == 1 and out.images[0].size == (512, 512)
    assert out['sample'] is out.images
    with pytest.raises(ValueError):
        pipe(guide, strength=1.5)
