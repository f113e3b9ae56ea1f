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
