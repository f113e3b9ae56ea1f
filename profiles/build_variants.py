'''Build A/B variants of the native library: `python profiles/build_variants.py name:-DK3_X=1,-DK3_Y=0 ...`
-> profiles/variants/libfd_<name>.so (git-ignored; loaded with FD_LIB_PATH).  Development aid.'''
import subprocess, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from flexdiffuse_b200 import build as B

out = Path(__file__).resolve().parent / 'variants'
out.mkdir(exist_ok=True)
for spec in sys.argv[1:]:
    name, _, defs = spec.partition(':')
    defs = [d for d in defs.split(',') if d]
    objs = []
    for src in B._sources():
        obj = out / f'{name}_{src.stem}.o'
        subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *defs, '-c', str(src), '-o', str(obj)], check=True,
                       capture_output=True)
        objs.append(str(obj))
    lib = out / f'libfd_{name}.so'
    subprocess.run([B._nvcc(), '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', str(lib), *objs,
                    '-cudart', 'static'], check=True, capture_output=True)
    print('built', lib)
