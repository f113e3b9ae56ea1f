'''K3F `fd_cross_attn_fused` (GPU, through the C ABI): the whole attn2 layer -- to_q GEMM, softmax
attention over the K2 cache, to_out GEMM + bias -- against fp32 torch on the same bf16 operands.

Stated tolerance (bf16 operands, fp32 accumulation, bf16 Q / P / attention output as under the
reference's autocast): relative L2 <= 1.5e-2 on both the attention output and the layer output;
the watchdog record must stay clear (no bounded wait timed out).'''
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1.5e-2
HEADS = 8


def _reference(x, wq, kv, k_off, v_off, idx, wo, bo, scale):
    S, N, C = x.shape
    d = C // HEADS
    q = (x.float() @ wq.float().t()).bfloat16().float()
    attn = torch.empty(S, N, C, device=x.device)
    for s in range(S):
        rows = kv[idx[s] * 80: idx[s] * 80 + 77].float()
        k = rows[:, k_off:k_off + C].view(77, HEADS, d)
        v = rows[:, v_off:v_off + C].view(77, HEADS, d)
        p = torch.softmax(torch.einsum('nhd,thd->hnt', q[s].view(N, HEADS, d), k) * scale, dim=-1)
        attn[s] = torch.einsum('hnt,thd->nhd', p, v).reshape(N, C)
    return attn, attn.bfloat16().float() @ wo.float().t() + bo.float()


def _case(dev, C, N, S, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    r = lambda *s: torch.randn(*s, device=dev, generator=g)
    x = r(S, N, C).bfloat16()
    wq, wo = (r(C, C) * C ** -0.5).bfloat16(), (r(C, C) * C ** -0.5).bfloat16()
    bo = (r(C) * 0.1).bfloat16()
    n_ctx, stride = 3, 2 * C + 64
    kv = r(n_ctx * 80, stride).bfloat16()
    kv.view(n_ctx, 80, stride)[:, 77:] = 0   # K2 rows of the padded context tokens are zero
    idx = torch.tensor([(2 * i + 1) % n_ctx for i in range(S)], dtype=torch.int32, device=dev)
    return x, wq, kv, 32, 32 + C, idx, wo, bo, (C // HEADS) ** -0.5


def _rel(a, b):
    return ((a.float() - b).norm() / b.norm()).item()


# the four (C, n_q) geometries of the SD-v1 UNet at 512^2, a ragged n_q (clipped tile) and a
# single partial tile; samples 1..5 exercise the shared / per-sample context rows
@pytest.mark.parametrize('C,N,S', [(320, 4096, 2), (640, 1024, 2), (1280, 256, 2), (1280, 64, 2),
                                   (320, 1000, 3), (640, 200, 1), (1280, 130, 5), (320, 77, 1)])
def test_fused_layer_matches_fp32(native, cuda_dev, C, N, S):
    args = _case(cuda_dev, C, N, S, seed=C + N + S)
    x, wq, kv, k_off, v_off, idx, wo, bo, scale = args
    want_attn, want_out = _reference(*args)
    out, attn = native.cross_attn_fused(x, wq, kv, k_off, v_off, idx, wo, bo, HEADS, 77, 80, scale)
    torch.cuda.synchronize()
    assert native.k3f_status() == [0, 0, 0, 0]
    assert torch.isfinite(out.float()).all() and torch.isfinite(attn.float()).all()
    assert _rel(attn, want_attn) < TOL
    assert _rel(out, want_out) < TOL


def test_fused_equals_unfused_sequence(native, cuda_dev):
    '''K3F against the round-1 sequence it replaces (cuBLAS to_q -> K3 -> cuBLAS to_out) on the same
    inputs: both round Q and the attention output to bf16, so they agree far inside the oracle bar.'''
    x, wq, kv, k_off, v_off, idx, wo, bo, scale = _case(cuda_dev, 640, 1024, 4, seed=11)
    out, attn = native.cross_attn_fused(x, wq, kv, k_off, v_off, idx, wo, bo, HEADS, 77, 80, scale)
    q = torch.nn.functional.linear(x, wq)
    a3 = native.cross_attn(q, kv, k_off, v_off, idx, HEADS, 77, 80, scale)
    o3 = torch.nn.functional.linear(a3, wo, bo)
    assert _rel(attn, a3.float()) < 5e-3
    assert _rel(out, o3.float()) < 5e-3


def test_fused_is_deterministic_and_context_indexed(native, cuda_dev):
    x, wq, kv, k_off, v_off, idx, wo, bo, scale = _case(cuda_dev, 320, 512, 4, seed=5)
    a = native.cross_attn_fused(x, wq, kv, k_off, v_off, idx, wo, bo, HEADS, 77, 80, scale)[0].clone()
    b = native.cross_attn_fused(x, wq, kv, k_off, v_off, idx, wo, bo, HEADS, 77, 80, scale)[0]
    assert torch.equal(a, b)
    # two samples with the same hidden states and the same context row give identical rows
    x2 = x.clone()
    x2[2] = x2[0]
    idx2 = idx.clone()
    idx2[2] = idx2[0]
    c = native.cross_attn_fused(x2, wq, kv, k_off, v_off, idx2, wo, bo, HEADS, 77, 80, scale)[0]
    assert torch.equal(c[0], c[2])
    assert not torch.equal(c[0], c[1])


def test_fused_argument_errors(native, cuda_dev):
    x, wq, kv, k_off, v_off, idx, wo, bo, scale = _case(cuda_dev, 320, 128, 1, seed=1)
    with pytest.raises(native.NativeError):   # C = 8 * 24 is not a multiple of 320
        native.cross_attn_fused(x[..., :192].contiguous(), wq[:192, :192].contiguous(), kv, k_off, v_off,
                                idx, wo[:192, :192].contiguous(), bo[:192].contiguous(), HEADS, 77, 80, scale)
    with pytest.raises(native.NativeError):   # t_pad must be 80
        native.cross_attn_fused(x, wq, kv, k_off, v_off, idx, wo, bo, HEADS, 77, 96, scale)
    with pytest.raises(native.NativeError):   # K slice beyond the cache row
        native.cross_attn_fused(x, wq, kv, kv.shape[1] - 8, v_off, idx, wo, bo, HEADS, 77, 80, scale)
