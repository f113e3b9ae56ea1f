'''StableDiffusionPipelineOutput (/root/reference/pipeline/flex.py:20,308).'''
from dataclasses import dataclass
from typing import Any, List


@dataclass
class StableDiffusionPipelineOutput:
    images: Any
    nsfw_content_detected: List[bool]

    def __getitem__(self, k):
        return self.images if k in ('sample', 'images') else getattr(self, k)


class StableDiffusionPipeline:  # imported by the reference's utils.py only
    pass
