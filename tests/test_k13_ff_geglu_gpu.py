'''K13 `fd_ff_geglu`: the GEGLU projection of diffusers' FeedForward (nn.Linear(C, 8C) then hidden * F.gelu(gate)), reached
from the UNet call at pipeline/guide.py:56-58, against fp32 torch on the same bf16 operands.  Tolerance 2e-2 (bf16 output
of a K <= 1280 contraction; the unfused path it replaces -- cuBLAS bf16 Linear + K6 -- is checked against the same bar).'''
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('M,C', [(8192, 320), (2048, 640), (512, 1280), (128, 1280), (1000, 320), (300, 640),
                                  (40000, 320)])   # (40000, 320): resident-x feature runs with a ragged last row block
def test_ff_geglu_matches_fp32(native, cuda_dev, M, C):
    g = torch.Generator(device=cuda_dev).manual_seed(M + C)
    x = torch.randn(M, C, device=cuda_dev, generator=g).bfloat16()
    w = (torch.randn(8 * C, C, device=cuda_dev, generator=g) * C ** -0.5).bfloat16()
    b = (torch.randn(8 * C, device=cuda_dev, generator=g) * 0.1).bfloat16()
    got = native.ff_geglu(x, w, b)
    h = x.float() @ w.float().t() + b.float()
    val, gate = h.chunk(2, dim=-1)
    want = val * F.gelu(gate)
    assert tuple(got.shape) == (M, 4 * C)
    torch.testing.assert_close(got.float(), want, rtol=2e-2, atol=2e-2)
    # the path it replaces (cuBLAS bf16 Linear + K6) rounds the projection to bf16 first: K13 does the same, so the two agree
    # to bf16 rounding of the result
    old = native.geglu(F.linear(x, w, b))
    assert (got.float() - old.float()).abs().max().item() <= 2e-2 * max(1.0, want.abs().max().item())
    assert native.lib().fd_debug_k13_flag() == 0
    # leading batch dims
    got3 = native.ff_geglu(x.view(2, M // 2, C), w, b)
    assert torch.equal(got3.view(M, 4 * C), got)
    # output as the left column block of a wider matrix (the merged output GEMM's operand): same bits, the rest untouched
    wide = torch.full((M, 5 * C), 3.0, device=cuda_dev).bfloat16()
    assert native.ff_geglu(x, w, b, out=wide[:, :4 * C]).data_ptr() == wide.data_ptr()
    assert torch.equal(wide[:, :4 * C], got) and (wide[:, 4 * C:] == 3.0).all()


def test_ff_geglu_gate_range(native, cuda_dev):
    '''The epilogue's gelu is `g * 2^p(|g|)` (a fit of log2 Phi(-a) on [0, 6.72], clamped at 11): drive the gate through
    [-60, 60], zeros and tiny values with an identity-like projection and compare with F.gelu on the same bf16 gates.
    Tolerance: one bf16 ulp of the result (2^-8 relative) plus 1e-6 absolute for the far negative tail.'''
    C = 320
    F_ = 4 * C
    w = torch.zeros(2 * F_, C, device=cuda_dev)
    idx = torch.arange(F_, device=cuda_dev)
    w[idx, idx % C] = 1.0                 # value[f] = x[f % C]
    w[F_ + idx, (idx + 1) % C] = 1.0      # gate[f]  = x[(f + 1) % C]
    gates = torch.cat([torch.linspace(-60, 60, 257), torch.tensor([0.0, -0.0, 1e-30, -1e-30, 1e-4, -1e-4]),
                       torch.linspace(-8, 8, 57)]).to(cuda_dev)
    x = torch.empty(256, C, device=cuda_dev)
    for r in range(256):
        x[r] = gates.roll(r)
    x = x.bfloat16()
    got = native.ff_geglu(x, w.bfloat16(), torch.zeros(2 * F_, device=cuda_dev).bfloat16()).float()
    xf = x.float()
    want = xf[:, idx % C] * F.gelu(xf[:, (idx + 1) % C].double()).float()
    assert torch.isfinite(got).all()
    err = (got - want).abs()
    assert (err <= want.abs() * 2.0 ** -8 + 1e-6).all(), float((err - want.abs() * 2.0 ** -8).max())
    assert native.lib().fd_debug_k13_flag() == 0


def test_ff_geglu_rejects_bad_shapes(native, cuda_dev):
    x = torch.randn(64, 96, device=cuda_dev).bfloat16()
    with pytest.raises(native.NativeError):
        native.ff_geglu(x, torch.randn(256, 96, device=cuda_dev).bfloat16(), torch.randn(256, device=cuda_dev).bfloat16())
    with pytest.raises(native.NativeError):
        native.ff_geglu(x.float(), torch.randn(256, 96, device=cuda_dev).bfloat16(), torch.randn(256, device=cuda_dev).bfloat16())
