'''Multi-GPU sweep driver: independent (prompt x seed x guidance-parameter) samples are
partitioned across ranks -- one process per GPU, no collective inside the denoising loop --
and the output latents are collected with ONE all-gather at the end.

The reference has no distributed code at all (single process, single device string,
/root/reference/utils.py:54); this is the replica-sharding BASELINE.json's north_star
asks for.  Noise is drawn per SAMPLE (generator seeded `base_seed + global_index`) so the
result does not depend on how the grid is sharded or micro-batched; this equals the
reference's single `randn((B,4,h,w), generator)` (flex.py:226-230) only for B = 1.
'''
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    '''Contiguous slice [lo, hi) of `n` samples for `rank`; sizes differ by at most 1.'''
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sample_noise(seed: int, shape: Sequence[int], device) -> torch.Tensor:
    '''Initial latents of ONE sample: shard- and batch-invariant.'''
    g = torch.Generator(device=device).manual_seed(int(seed))
    return torch.randn(tuple(shape), generator=g, device=device)


def run_sweep(n_samples: int,
              denoise: Callable[[int, int], torch.Tensor],
              micro_batch: int = 16,
              gather: bool = True) -> torch.Tensor:
    '''denoise(lo, hi) -> latents [hi-lo, ...] for global sample indices [lo, hi).
    Returns all `n_samples` latents (on every rank) when `gather`, else the local shard.'''
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank() if distributed else 0
    world = dist.get_world_size() if distributed else 1
    lo, hi = shard_range(n_samples, rank, world)
    outs: List[torch.Tensor] = []
    for s in range(lo, hi, micro_batch):
        outs.append(denoise(s, min(s + micro_batch, hi)))
    local = torch.cat(outs) if outs else None
    if not (gather and distributed and world > 1):
        return local
    # shards differ by at most one sample: pad to the largest, gather once, trim
    counts = [shard_range(n_samples, r, world) for r in range(world)]
    width = max(h - l for l, h in counts)
    probe = local if local is not None else denoise(0, 0)
    pad = torch.zeros((width,) + tuple(probe.shape[1:]), dtype=probe.dtype,
                      device=probe.device)
    if local is not None:
        pad[:local.shape[0]] = local
    buf = torch.empty((world * width,) + tuple(pad.shape[1:]), dtype=pad.dtype,
                      device=pad.device)
    dist.all_gather_into_tensor(buf, pad)
    parts = [buf[r * width:r * width + (h - l)] for r, (l, h) in enumerate(counts)]
    return torch.cat(parts)


def make_denoiser(pipe, encoder, unet, contexts: torch.Tensor, seeds: Sequence[int],
                  guidance: float, steps: int, init_size=(512, 512),
                  use_cuda_graph: bool = True) -> Callable[[int, int], torch.Tensor]:
    '''denoise(lo, hi) for `run_sweep`: sample i uses context `contexts[i]` ([N,77,768], e.g. the
    K1 blends of a prompt x guidance-parameter grid) and noise seeded `seeds[i]`.'''
    from .pipeline.guide import SimpleGuide
    h, w = init_size
    dev = contexts.device

    def denoise(lo: int, hi: int) -> torch.Tensor:
        if hi <= lo:
            return torch.zeros((0, unet.in_channels, h // 8, w // 8), device=dev)
        guide = SimpleGuide(encoder, unet, guidance, steps, contexts[lo:hi],
                            use_cuda_graph=use_cuda_graph)
        noise = torch.stack([
            sample_noise(seeds[i], (unet.in_channels, h // 8, w // 8), dev)
            for i in range(lo, hi)
        ])
        return pipe(guide, init_size=init_size, latents=noise, output_type='latent',
                    return_dict=False)

    return denoise
