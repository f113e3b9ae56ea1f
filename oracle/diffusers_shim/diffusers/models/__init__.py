'''fp32 nn.Module faces of the functional restatement in oracle/unet_oracle.py.'''
from dataclasses import dataclass

import torch

from oracle import unet_oracle as U


@dataclass
class _Out:
    sample: torch.Tensor


class UNet2DConditionModel(torch.nn.Module):
    '''Holds a diffusers-named state_dict as buffers; forward = unet_oracle.unet_forward.'''
    def __init__(self, state_dict):
        super().__init__()
        self.in_channels = 4
        self.config = {'attention_head_dim': 8}
        self._keys = list(state_dict)
        for k, v in state_dict.items():
            self.register_buffer(k.replace('.', '__'), v.detach().clone().float())

    def sd(self):
        return {k: getattr(self, k.replace('.', '__')) for k in self._keys}

    def set_attention_slice(self, slice_size):
        pass

    def forward(self, sample, timestep, encoder_hidden_states):
        return _Out(U.unet_forward(self.sd(), sample, timestep,
                                   encoder_hidden_states))


class _Dist:
    def __init__(self, moments):
        self.mean, logvar = torch.chunk(moments, 2, dim=1)
        self.std = torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0))

    def sample(self, generator=None):
        return self.mean + self.std * torch.randn(
            self.mean.shape, generator=generator, device=self.mean.device)


@dataclass
class _Enc:
    latent_dist: _Dist


class AutoencoderKL(torch.nn.Module):
    def __init__(self, state_dict):
        super().__init__()
        self._keys = list(state_dict)
        for k, v in state_dict.items():
            self.register_buffer(k.replace('.', '__'), v.detach().clone().float())

    def sd(self):
        return {k: getattr(self, k.replace('.', '__')) for k in self._keys}

    def decode(self, z):
        return _Out(U.vae_decode(self.sd(), z))

    def encode(self, x):
        return _Enc(_Dist(U.vae_encode_moments(self.sd(), x)))
