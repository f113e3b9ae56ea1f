'''The C-ABI library builds for sm_100a, loads without a GPU and exports every
symbol include/flexdiffuse_b200.h declares.  No compute calls here.'''
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_all_exported(native):
    header = open(os.path.join(ROOT, 'include', 'flexdiffuse_b200.h')).read()
    declared = set(re.findall(r'\b(fd_[a-z0-9_]+)\s*\(', header))
    assert declared == set(native.ABI_SYMBOLS), declared ^ set(native.ABI_SYMBOLS)
    lib = ctypes.CDLL(str(native.LIB_PATH))
    for sym in declared:
        assert hasattr(lib, sym), f'{sym} not exported'
    assert lib.fd_version() == native.FD_ABI_VERSION


def test_struct_layouts_match_header(native):
    assert ctypes.sizeof(native.SchedCoeffs) == 40
    assert ctypes.sizeof(native.TweenParams) == 56


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only behaviour')
def test_no_cpu_fallback(native):
    '''Without a GPU the product path fails loudly instead of falling back.'''
    with pytest.raises(native.NativeError):
        native.require_device(0)
    x = torch.zeros(8)
    with pytest.raises(native.NativeError):
        native.cfg_sched_step(None, x, x, native.SchedCoeffs(), x.clone())


def test_sass_is_blackwell_native(native):
    '''tcgen05 / TMA instructions are present in the built cubin.'''
    import shutil
    import subprocess
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    if not os.path.exists(cuobjdump):
        pytest.skip('cuobjdump not available')
    sass = subprocess.run([cuobjdump, '-sass', str(native.LIB_PATH)],
                          capture_output=True, text=True).stdout
    assert 'UTCHMMA' in sass      # tcgen05.mma kind::f16 / tf32
    assert 'UTMALDG' in sass      # cp.async.bulk.tensor
    assert 'LDTM' in sass and 'STTM' in sass
    assert 'sm_100a' in sass or 'sm_100' in sass
