'''flexdiffuse_b200 -- B200-native (sm_100a) hot path of tim-speed/flexdiffuse.

Mirrors the reference's module-level exports (/root/reference/__init__.py:7-14)
for the path this package replaces: image-guided conditioning and the denoising
loop.  `Runner` / `image_grid` (host orchestration, UI) are out of scope.
'''
from . import _native  # noqa: F401  (ctypes binding; loads lazily)

__all__ = ['_native']
