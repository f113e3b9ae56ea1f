'''BASELINE.json north_star parity bar: "50-step decoded images within a stated PSNR".

configs[1] at FULL size -- SD v1.5 UNet 512x512, 50-step DDIM, CFG 7.5, batch 1 -- through the
product pipeline (bf16 UNet, K2 cache, K3F / K3, fused K4, CUDA graph, K10 tail) against the fp32
oracle loop (oracle/loop_oracle.py on oracle/unet_oracle.py, TF32 off) on the same random-init
weights, same prompt embeddings, same initial noise.

Stated tolerances: final latents relative L2 <= 3e-2; decoded 512x512 images PSNR >= 38 dB
(measured on B200: 1.3e-2 and 46.8 dB)
(50 sequential bf16 UNet evaluations, each amplified 7.5x by CFG; the 10-step 256^2 cases in
test_pipeline_parity.py hold 5e-2 / 30 dB).'''
import pytest
import torch

from flexdiffuse_b200 import schedulers as prod
from flexdiffuse_b200.pipeline.flex import FlexPipeline
from flexdiffuse_b200.pipeline.guide import SimpleGuide
from oracle import loop_oracle as lo
from oracle import unet_oracle as U
from tests.model_helpers import models, psnr, rel_l2

pytestmark = pytest.mark.gpu

LATENT_TOL = 3e-2
PSNR_MIN = 38.0


class _Enc:
    def __init__(self, uncond):
        self.uncond = uncond

    def prompt(self, p):
        return self.uncond


def test_50_step_512_ddim_images_match_fp32_oracle(native, cuda_dev):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    unet, vae, usd, vsd = models(str(cuda_dev))
    g = torch.Generator(device=cuda_dev).manual_seed(2024)
    uncond = torch.randn(1, 77, 768, device=cuda_dev, generator=g)
    embeds = torch.randn(1, 77, 768, device=cuda_dev, generator=g)
    pipe = FlexPipeline(vae, None, None, unet, prod.DDIMScheduler())
    guide = SimpleGuide(_Enc(uncond), unet, 7.5, 50, embeds, use_cuda_graph=True)
    lat = pipe(guide, init_size=(512, 512), generator=torch.Generator(device=cuda_dev).manual_seed(50),
               output_type='latent', return_dict=False)
    with torch.no_grad():
        want = lo.denoise(lambda x, t, c: U.unet_forward(usd, x, t, c.bfloat16().float()),
                          lo.DDIMScheduler(), uncond, embeds, 7.5, 50, init_size=(512, 512),
                          generator=torch.Generator(device=cuda_dev).manual_seed(50), device=cuda_dev)
        img = pipe.decode(lat, 'pt').float()
        img_want = (U.vae_decode(vsd, want / 0.18215) / 2 + 0.5).clamp(0, 1)
    assert tuple(img.shape) == (1, 3, 512, 512)
    e, p = rel_l2(lat, want), psnr(img, img_want)
    print(f'50-step 512^2 DDIM: latents rel-L2 {e:.3e}, decoded PSNR {p:.1f} dB')
    assert torch.isfinite(lat).all() and want.abs().mean() > 1e-2
    assert e < LATENT_TOL, e
    assert p > PSNR_MIN, p
