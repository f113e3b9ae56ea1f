'''Shared random-init models for the GPU parity tests (built once per session).'''
import functools

import torch


@functools.lru_cache(maxsize=None)
def models(device_str: str):
    from flexdiffuse_b200 import factory
    dev = torch.device(device_str)
    unet = factory.build_unet(dev, torch.bfloat16, seed=0)
    vae = factory.build_vae(dev, torch.bfloat16, seed=1)
    # the oracle computes in fp32 on the SAME (bf16-representable) weights
    unet_sd = {k: v.float().contiguous() for k, v in unet.state_dict().items()}
    vae_sd = {k: v.float().contiguous() for k, v in vae.state_dict().items()}
    return unet, vae, unet_sd, vae_sd


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    return ((a.float() - b.float()).norm() / b.float().norm()).item()


def psnr(a: torch.Tensor, b: torch.Tensor) -> float:
    mse = ((a.float() - b.float())**2).mean().item()
    return 99.0 if mse == 0 else 10 * torch.log10(torch.tensor(1.0 / mse)).item()
