'''K2 parity: tcgen05 K/V projection GEMM vs fp32 torch matmul on the same
bf16-rounded operands.  Tolerance = bf16 output rounding (2^-8 relative).'''
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('M,N,K', [(80, 256, 768), (160, 640, 768),
                                   (720, 24960, 768), (77, 328, 64),
                                   (300, 136, 128)])
def test_k2_matches_torch(native, cuda_dev, M, N, K):
    g = torch.Generator().manual_seed(M + N)
    ctx = torch.randn(M, K, generator=g).to(cuda_dev).bfloat16()
    w = (torch.randn(N, K, generator=g) / K**0.5).to(cuda_dev).bfloat16()
    out = native.kv_project(ctx, w)
    torch.cuda.synchronize()
    ref = ctx.float() @ w.float().T
    err = (out.float() - ref).abs().max().item()
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2), err


def test_k2_rejects_bad_k(native, cuda_dev):
    ctx = torch.zeros(16, 40, device=cuda_dev, dtype=torch.bfloat16)
    w = torch.zeros(16, 40, device=cuda_dev, dtype=torch.bfloat16)
    with pytest.raises(native.NativeError):
        native.kv_project(ctx, w)
