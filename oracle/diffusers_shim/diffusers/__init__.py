'''ORACLE shim (test infrastructure): the minimal `diffusers==0.3.0`-compatible surface
that /root/reference/pipeline/flex.py:15-20, pipeline/guide.py:3 and
composition/guide.py:5 import, so that the reference's own loop runs UNCHANGED on top of
this repo's fp32 restatement of the diffusers arithmetic (oracle/unet_oracle.py and the
schedulers here).  diffusers itself is a third-party dependency that is neither
vendored in the reference nor installable in this image (no network).
Put `oracle/diffusers_shim` on sys.path *before* importing the reference modules.'''
__version__ = '0.3.0+oracle-shim'
