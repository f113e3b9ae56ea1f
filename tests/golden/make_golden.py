'''Generate the golden vectors that pin oracle/guidance_oracle.py to the reference.

Runs ONLY in the build container, where /root/reference is mounted: it imports the
UNMODIFIED reference `guidance.py` and records, for seeded synthetic inputs
(oracle.guidance_oracle.synthetic_pair -- numpy RandomState, machine independent):

  * `_map_emb` output (77 x 2 float64)                      guidance.py:23-85
  * `Tweener.tween` output ([1,77,D] fp32), or the fact that it raised
    ZeroDivisionError                           guidance.py:215-272
  * `ConceptMapper(...).map(...)` output                    guidance.py:275-312

Small cases (D=64) store the full output tensor; full-size cases (D=768) store a
sha256 of the output bytes plus three sampled rows.  The file is committed;
nothing at test time reads /root/reference.

    python tests/golden/make_golden.py   ->  tests/golden/guidance_golden.npz
'''
import contextlib
import hashlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

import guidance as ref  # noqa: E402  (the reference, unmodified)
from oracle import guidance_oracle as orc  # noqa: E402  (only for inputs/params)

N_SMALL, N_FULL, N_CONCEPT = 36, 6, 3


def run_ref(txt, img, prm):
    tw = ref.Tweener(prm.threshold, prm.linear, prm.clustered, prm.max_guidance,
                     prm.header_max, prm.align_mode, prm.mapping_reuse)
    with contextlib.redirect_stdout(io.StringIO()):
        mapped = ref._map_emb(img, txt, prm.mapping_reuse, prm.align_mode)
        try:
            out = tw.tween(txt, img)
        except ZeroDivisionError:
            out = None
    return mapped, out


def prm_row(p):
    return np.array([
        p.threshold[0], p.threshold[1], p.linear[0], p.linear[1], p.clustered,
        p.max_guidance, p.header_max, p.align_mode,
        float(p.mapping_reuse)
    ])


def main():
    torch.set_num_threads(1)  # bit-reproducible reductions
    rs = np.random.RandomState(20260101)
    store = {}
    meta = []
    case = 0
    for kind, n, D in (('small', N_SMALL, 64), ('full', N_FULL, 768)):
        for _ in range(n):
            seed = 5000 + case
            planted = int(rs.choice([0, 12, 30]))
            A = int(rs.choice([257, 77])) if kind == 'small' else 257
            prm = orc.random_params(rs)
            if case % 6 == 0:  # make sure the defaults are represented
                prm = orc.TweenParams(align_mode=prm.align_mode,
                                      mapping_reuse=prm.mapping_reuse)
            txt, img = orc.synthetic_pair(seed, A=A, D=D, planted=planted)
            mapped, out = run_ref(txt, img, prm)
            key = f'c{case:03d}'
            store[key + '_prm'] = prm_row(prm)
            store[key + '_mapped'] = mapped
            store[key + '_zde'] = np.array(out is None)
            if out is not None:
                o = out.numpy()
                store[key + '_sha'] = np.frombuffer(
                    hashlib.sha256(o.tobytes()).digest(), dtype=np.uint8)
                if kind == 'small':
                    store[key + '_out'] = o
                else:
                    store[key + '_rows'] = o[0, [0, 5, 76]]
            meta.append((case, seed, A, D, planted))
            case += 1
    # ConceptMapper
    for j in range(N_CONCEPT):
        seed = 7000 + j
        txt, img = orc.synthetic_pair(seed, D=64, planted=20)
        con, _ = orc.synthetic_pair(seed + 50, D=64, planted=0)
        con[0, 1:6] = txt[0, 3:8] * 1.25  # concepts aligned with prompt tokens
        img[0, 40:45] = con[0, 1:6] * 0.8  # and present in the image
        with contextlib.redirect_stdout(io.StringIO()):
            cm = ref.ConceptMapper(img, con)
            out = cm.map(txt)
        store[f'm{j}_out'] = out.numpy()
        store[f'm{j}_seed'] = np.array(seed)
    store['meta'] = np.array(meta, dtype=np.int64)
    path = os.path.join(HERE, 'guidance_golden.npz')
    np.savez_compressed(path, **store)
    n_zde = sum(bool(store[f'c{c:03d}_zde']) for c in range(case))
    print(f'wrote {path}: {case} tween cases ({n_zde} ZeroDivisionError), '
          f'{N_CONCEPT} concept cases, {os.path.getsize(path) / 1e6:.2f} MB')


if __name__ == '__main__':
    main()
