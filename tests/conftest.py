import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a real B200 (sm_100) GPU')


@pytest.fixture(scope='session')
def native():
    '''The built C-ABI library; building it here if needed (nvcc only).'''
    from flexdiffuse_b200 import build as fd_build
    fd_build.build()
    from flexdiffuse_b200 import _native
    _native.lib()
    return _native


@pytest.fixture(scope='session')
def cuda_dev(native):
    import torch
    if not torch.cuda.is_available():
        pytest.fail('gpu-marked test running without a CUDA device')
    native.require_device(0)
    return torch.device('cuda:0')
