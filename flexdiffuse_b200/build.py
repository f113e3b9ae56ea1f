'''In-tree build of the C-ABI CUDA library (sm_100a only).

`python -m flexdiffuse_b200.build` (or `__graft_entry__.build()`) compiles every
`csrc/*.cu` with nvcc for `compute_100a/sm_100a` and links
`flexdiffuse_b200/libflexdiffuse_b200.so`.  nvcc cross-compiles without a GPU.
The library is git-ignored but travels with the repo snapshot to the GPU box.
'''
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / 'csrc'
LIB_PATH = PKG_DIR / 'libflexdiffuse_b200.so'
BUILD_DIR = PKG_DIR / 'build'

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-std=c++17',
    '-lineinfo', '-Xcompiler', '-fPIC', '-Xptxas', '-v', '--expt-relaxed-constexpr'
]


def _nvcc() -> str:
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; cannot build flexdiffuse_b200')
    return nvcc


def _sources():
    return sorted(CSRC.glob('*.cu'))


def _stamp() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob('*')) +
                    [PKG_DIR.parent / 'include' / 'flexdiffuse_b200.h']):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    '''Compile csrc/*.cu -> libflexdiffuse_b200.so (skipped when up to date).'''
    stamp_file = BUILD_DIR / 'stamp'
    stamp = _stamp()
    if (not force and LIB_PATH.exists() and stamp_file.exists()
            and stamp_file.read_text() == stamp):
        return LIB_PATH
    BUILD_DIR.mkdir(exist_ok=True)
    nvcc = _nvcc()
    objs = []

    def compile_one(src: Path):
        obj = BUILD_DIR / (src.stem + '.o')
        cmd = [nvcc, *NVCC_FLAGS, '-c', str(src), '-o', str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        (BUILD_DIR / (src.stem + '.ptxas.log')).write_text(res.stderr)
        if res.returncode != 0:
            raise RuntimeError(f'nvcc failed for {src.name}:\n{res.stdout}\n'
                               f'{res.stderr}')
        if verbose:
            print(res.stderr, file=sys.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    link = [
        nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o',
        str(LIB_PATH), *map(str, objs), '-cudart', 'static'
    ]
    res = subprocess.run(link, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f'link failed:\n{res.stdout}\n{res.stderr}')
    stamp_file.write_text(stamp)
    return LIB_PATH


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='-v' in sys.argv)
    print(path)
