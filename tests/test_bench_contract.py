'''bench.py contract checks that need no GPU: the reference (CPU) arm prints one JSON line with
the keys the driver reads, and the B200 arm refuses to run without a GPU instead of falling
back to anything.'''
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(*args, timeout=600):
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), *args],
                          capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_arm_json_line():
    res = _run('--impl', 'reference', '--steps', '1', '--warmup', '0')
    assert res.returncode == 0, res.stderr[-2000:]
    line = json.loads(res.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['unit'] == 'images/s'
    assert line['metric'] == 'sd15_512x512_50step_ddim_cfg7.5_images_per_s'
    assert line['higher_is_better'] is True and line['value'] > 0
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1
    assert 'sample' in line['cpu_baseline']
    assert line['e2e'] == {'value': line['value'], 'unit': 'images/s',
                           'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert 'workload' in line['config'] and 'model' not in line['config']


@pytest.mark.skipif(torch.cuda.is_available(), reason='CPU-only behaviour')
def test_b200_arm_fails_loudly_without_a_gpu():
    res = _run('--steps', '1', '--warmup', '1', timeout=300)
    assert res.returncode != 0
    assert 'NativeError' in res.stderr or 'CUDA' in res.stderr or 'cuda' in res.stderr
