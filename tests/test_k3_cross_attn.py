'''K3 parity: tcgen05 cross-attention over the cached K/V vs fp32 torch
softmax(QK^T*scale)V on the same bf16-rounded operands (the arithmetic of
diffusers CrossAttention._attention).  Stated bf16 tolerance: 2e-2 abs on
O(1) outputs (P is rounded to bf16 before the second MMA).'''
import pytest
import torch

pytestmark = pytest.mark.gpu

T_VALID, T_PAD = 77, 80


def _ref(q, kv, koff, voff, ctx_index, heads, scale):
    S, Nq, C = q.shape
    d = C // heads
    outs = []
    for s in range(S):
        rows = kv[ctx_index[s] * T_PAD:ctx_index[s] * T_PAD + T_VALID].float()
        k = rows[:, koff:koff + C].view(T_VALID, heads, d).permute(1, 0, 2)
        v = rows[:, voff:voff + C].view(T_VALID, heads, d).permute(1, 0, 2)
        qq = q[s].float().view(Nq, heads, d).permute(1, 0, 2)
        p = torch.softmax(qq @ k.transpose(1, 2) * scale, dim=-1)
        outs.append((p @ v).permute(1, 0, 2).reshape(Nq, C))
    return torch.stack(outs)


@pytest.mark.parametrize('n_q,C', [(4096, 320), (1024, 640), (256, 1280),
                                   (64, 1280), (200, 320)])
def test_k3_matches_torch(native, cuda_dev, n_q, C):
    heads, S, n_ctx = 8, 3, 2
    g = torch.Generator().manual_seed(n_q + C)
    q = torch.randn(S, n_q, C, generator=g).to(cuda_dev).bfloat16()
    width = 2 * C + 64  # K | pad | V inside a wider cache row
    kv = torch.randn(n_ctx * T_PAD, width, generator=g).to(cuda_dev).bfloat16()
    koff, voff = 0, C + 64
    ctx_index = torch.tensor([1, 0, 1], dtype=torch.int32, device=cuda_dev)
    scale = (C // heads)**-0.5
    out = native.cross_attn(q, kv, koff, voff, ctx_index, heads, T_VALID,
                            T_PAD, scale)
    torch.cuda.synchronize()
    ref = _ref(q, kv, koff, voff, ctx_index.tolist(), heads, scale)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize('S', [1, 2, 16, 40])
@pytest.mark.parametrize('n_q,C', [(4096, 320), (1024, 640), (1024, 1280),
                                   (300, 640)])
def test_k3_tile_walk(native, cuda_dev, S, n_q, C):
    '''The sample count changes how many query tiles one CTA walks (1 ... 16, odd and even,
    ragged last tile), which is what the two alternating softmax warpgroups, the Q ring and the
    O-accumulator hand-off depend on.  Values differ per tile so a swapped tile is caught.'''
    heads = 8
    g = torch.Generator().manual_seed(S * n_q + C)
    q = torch.randn(S, n_q, C, generator=g).to(cuda_dev).bfloat16()
    kv = torch.randn(3 * T_PAD, 2 * C, generator=g).to(cuda_dev).bfloat16()
    ctx_index = (torch.arange(S, dtype=torch.int32) % 3).to(cuda_dev)
    scale = (C // heads)**-0.5
    out = native.cross_attn(q, kv, 0, C, ctx_index, heads, T_VALID, T_PAD, scale)
    torch.cuda.synchronize()
    ref = _ref(q, kv, 0, C, ctx_index.tolist(), heads, scale)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)


@pytest.mark.parametrize('C', [320, 640, 1280])
def test_k3_every_key_position(native, cuda_dev, C):
    '''Each query row is aligned with one key (row r -> key r % 77) at three sharpness levels, so
    every key column carries the weight for some row: part of the columns take 2^x from the FMA-pipe
    polynomial and part from MUFU.EX2, and both must stay inside the same 2e-2 tolerance.  The soft
    rows (gain 0.5, many keys contribute) would expose a systematic bias between the two paths.'''
    heads, n_q = 8, 3 * 77
    d = C // heads
    g = torch.Generator().manual_seed(C)
    kv = torch.randn(T_PAD, 2 * C, generator=g).bfloat16()
    k = kv[:T_VALID, :C].float().view(T_VALID, heads, d)
    q = torch.empty(1, n_q, C)
    for r in range(n_q):
        gain = (0.5, 2.0, 8.0)[r // 77]
        q[0, r] = (gain * k[r % 77]).reshape(C)
    q = q.to(cuda_dev).bfloat16()
    kv = kv.to(cuda_dev)
    ctx_index = torch.zeros(1, dtype=torch.int32, device=cuda_dev)
    scale = d**-0.5
    out = native.cross_attn(q, kv, 0, C, ctx_index, heads, T_VALID, T_PAD, scale)
    torch.cuda.synchronize()
    ref = _ref(q, kv, 0, C, [0], heads, scale)
    torch.testing.assert_close(out.float(), ref, rtol=2e-2, atol=2e-2)
    # the sharp rows reproduce "their" value row
    v = kv[:T_VALID, C:].float()
    sharp = out[0, 2 * 77:].float()
    assert (sharp - v).abs().max() < 5e-2


def test_k3_peaked_softmax(native, cuda_dev):
    '''Large logits: softmax must stay finite and pick the dominant key.'''
    heads, C, n_q = 8, 320, 128
    q = torch.zeros(1, n_q, C, device=cuda_dev, dtype=torch.bfloat16)
    q[..., 0] = 30.0
    kv = torch.zeros(T_PAD, 2 * C, device=cuda_dev, dtype=torch.bfloat16)
    kv[5, 0] = 30.0       # key 5 matches head 0 strongly
    kv[5, C:C + 40] = 2.0  # its value
    kv[78, 0] = 100.0     # padded key (>= 77) must be masked
    kv[78, C:C + 40] = -7.0
    ctx_index = torch.zeros(1, dtype=torch.int32, device=cuda_dev)
    out = native.cross_attn(q, kv, 0, C, ctx_index, heads, T_VALID, T_PAD,
                            40**-0.5)
    torch.cuda.synchronize()
    assert torch.isfinite(out.float()).all()
    torch.testing.assert_close(out[0, :, :40].float(),
                               torch.full((n_q, 40), 2.0, device=cuda_dev),
                               rtol=1e-2, atol=1e-2)
