'''Composition schema -- API mirror of /root/reference/composition/schema.py:6-25
(pure data; kept because composition/embeds.py consumes it).'''
from dataclasses import asdict, dataclass
import json
from typing import List, Tuple


@dataclass
class EntitySchema():
    prompt: str
    offset: Tuple[int, int]
    size: Tuple[int, int]
    blend: float = 0.8


@dataclass
class Schema():
    background_prompt: str
    style_start_prompt: str
    style_end_prompt: str
    style_blend: Tuple[float, float]
    entities: List[EntitySchema]

    def json(self) -> str:
        return json.dumps(asdict(self))
