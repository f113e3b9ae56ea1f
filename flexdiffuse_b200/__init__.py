'''flexdiffuse_b200 -- B200-native (sm_100a) hot path of tim-speed/flexdiffuse.

Mirrors the reference's module-level exports (/root/reference/__init__.py:7-14) for the path
this package replaces: image-guided conditioning and the denoising loop.  `Runner` and
`image_grid` (host orchestration, PNG grids, gradio UI) are out of scope.
'''
from . import _native  # noqa: F401  (ctypes binding of include/flexdiffuse_b200.h; loads lazily)
from . import guidance
from .composition.guide import CompositeGuide
from .encode import clip as encode
from .pipeline import flex
from .pipeline.guide import GuideBase, PromptGuide, SimpleGuide

CLIPEncoder = encode.CLIPEncoder
GUIDE_ORDER_TEXT = guidance.GUIDE_ORDER_TEXT
GUIDE_ORDER_ALIGN = guidance.GUIDE_ORDER_ALIGN
GUIDE_ORDER_DIRECT = guidance.GUIDE_ORDER_DIRECT
Guide = guidance.Guide
preprocess = encode.preprocess
FlexPipeline = flex.FlexPipeline

__all__ = ['CLIPEncoder', 'GUIDE_ORDER_TEXT', 'GUIDE_ORDER_ALIGN', 'GUIDE_ORDER_DIRECT',
           'Guide', 'preprocess', 'FlexPipeline', 'GuideBase', 'SimpleGuide', 'PromptGuide',
           'CompositeGuide']
