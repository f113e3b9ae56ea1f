'''Text embeddings of a composition schema.  API-compatible with
/root/reference/composition/embeds.py:10-44 (`EntityEmbeds`, `Embeds`, `px_to_block`,
`encode_entity`, `encode_schema`): every prompt of the schema becomes a [1, 77, 768]
`CLIPEncoder.prompt` embedding and pixel boxes become 8x8 latent blocks.

Unlike the reference, `encode_schema` runs the text tower ONCE over all prompts of the schema
(background, both style prompts, every entity) instead of once per prompt; the rows are then
handed out as [1, 77, 768] views, so the values are identical.
'''
from __future__ import annotations

import dataclasses
from typing import List, Sequence, Tuple

import torch

from ..encode.clip import CLIPEncoder
from .schema import EntitySchema, Schema

LATENT_BLOCK_PX = 8  # one latent cell covers 8 x 8 image pixels


@dataclasses.dataclass
class EntityEmbeds():
    embed: torch.Tensor
    offset_blocks: Tuple[int, int]
    size_blocks: Tuple[int, int]
    blend: float


@dataclasses.dataclass
class Embeds():
    background_embed: torch.Tensor
    style_start_embed: torch.Tensor
    style_end_embed: torch.Tensor
    style_blend: Tuple[float, float]
    entities: List[EntityEmbeds]


def px_to_block(px_shape: Sequence[int]) -> Tuple[int, ...]:
    '''Image pixels -> latent blocks, floor division like the reference.'''
    return tuple(int(v) // LATENT_BLOCK_PX for v in px_shape)


def _entity(e: EntitySchema, embed: torch.Tensor) -> EntityEmbeds:
    return EntityEmbeds(embed, px_to_block(e.offset), px_to_block(e.size), e.blend)


def encode_entity(e: EntitySchema, encode: CLIPEncoder) -> EntityEmbeds:
    return _entity(e, encode.prompt(e.prompt))


def encode_schema(s: Schema, encode: CLIPEncoder) -> Embeds:
    prompts = [s.background_prompt, s.style_start_prompt, s.style_end_prompt]
    prompts += [e.prompt for e in s.entities]
    rows = encode.prompt(prompts)  # one batched text-tower call: [3 + E, 77, 768]
    if rows.shape[0] != len(prompts):
        # an encoder that cannot batch (e.g. a stub returning one row): fall back to per-prompt calls
        rows = torch.cat([encode.prompt(p) for p in prompts])
    each = [rows[i:i + 1] for i in range(len(prompts))]
    return Embeds(each[0], each[1], each[2], s.style_blend,
                  [_entity(e, each[3 + i]) for i, e in enumerate(s.entities)])
