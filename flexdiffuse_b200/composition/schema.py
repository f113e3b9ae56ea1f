'''Composition schema: the plain-data description of a regional composition that
`CompositeGuide` consumes.  API-compatible with /root/reference/composition/schema.py:6-25
(same class names, field order, defaults and `Schema.json()` output); pixel boxes are
(x, y) offsets and (width, height) sizes in image pixels, converted to 8x8 latent blocks by
`composition.embeds.px_to_block`.

Beyond the reference: fields are normalised to tuples on construction, and `Schema.from_json`
round-trips what `json()` writes (handy for the sweep driver's job files).
'''
from __future__ import annotations

import dataclasses
import json
from typing import Any, Dict, List, Tuple

_Pair = Tuple[int, int]


def _pair(value, kind=int) -> tuple:
    a, b = value
    return (kind(a), kind(b))


@dataclasses.dataclass
class EntitySchema():
    '''One prompted region: `prompt` guides the box at `offset` (x, y) of `size` (w, h) pixels,
    lerped into the background prediction with weight `blend`.'''
    prompt: str
    offset: _Pair
    size: _Pair
    blend: float = 0.8

    def __post_init__(self):
        self.offset = _pair(self.offset)
        self.size = _pair(self.size)

    def as_dict(self) -> Dict[str, Any]:
        return {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}


@dataclasses.dataclass
class Schema():
    '''Background prompt, a (currently unused by the guide) style tween, and the entity list.'''
    background_prompt: str
    style_start_prompt: str
    style_end_prompt: str
    style_blend: Tuple[float, float]
    entities: List[EntitySchema]

    def __post_init__(self):
        self.style_blend = _pair(self.style_blend, float)
        self.entities = list(self.entities)

    def json(self) -> str:
        '''Same text as the reference's `Schema.json()` (tuples become JSON arrays).'''
        body = {f.name: getattr(self, f.name) for f in dataclasses.fields(self)}
        body['entities'] = [e.as_dict() for e in self.entities]
        return json.dumps(body)

    @classmethod
    def from_json(cls, text: str) -> 'Schema':
        raw = json.loads(text)
        raw['entities'] = [EntitySchema(**e) for e in raw['entities']]
        return cls(**raw)
