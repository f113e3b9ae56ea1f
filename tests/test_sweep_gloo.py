'''Multi-rank sweep logic on CPU: world_size 2 over gloo (the N>1 path of bench.py),
plus shard arithmetic.  No GPU needed.'''
import os
import subprocess
import sys

import torch

from flexdiffuse_b200 import sweep

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from flexdiffuse_b200 import sweep
dist.init_process_group('gloo')
def denoise(lo, hi):
    # deterministic per-sample "latents" from per-sample seeds (shard invariant)
    return torch.stack([sweep.sample_noise(1000 + i, (4, 8, 8), 'cpu') * (i + 1)
                        for i in range(lo, hi)]) if hi > lo else torch.zeros(0, 4, 8, 8)
out = sweep.run_sweep(int(sys.argv[2]), denoise, micro_batch=3)
want = denoise(0, int(sys.argv[2]))
assert torch.equal(out, want), (dist.get_rank(), out.shape)
lo, hi = sweep.shard_range(int(sys.argv[2]), dist.get_rank(), dist.get_world_size())
print(f'rank {dist.get_rank()} ok shard [{lo},{hi})')
dist.destroy_process_group()
'''


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 16, 1024):
        for world in (1, 2, 3, 8):
            spans = [sweep.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [h - l for l, h in spans]
            assert max(sizes) - min(sizes) <= 1


def test_single_process_sweep_is_shard_invariant():
    f = lambda lo, hi: torch.stack(
        [sweep.sample_noise(5 + i, (4, 8, 8), 'cpu') for i in range(lo, hi)])
    a = sweep.run_sweep(10, f, micro_batch=4)
    b = sweep.run_sweep(10, f, micro_batch=1)
    assert torch.equal(a, b)


def test_two_rank_gloo_sweep(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER)
    for n in (8, 7):
        res = subprocess.run([
            sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
            '--nproc-per-node=2', '--master-addr', '127.0.0.1', '--master-port',
            str(29500 + n), str(script), ROOT, str(n)
        ], capture_output=True, text=True, timeout=240)
        assert res.returncode == 0, res.stdout + res.stderr
        assert res.stdout.count(' ok shard') == 2
