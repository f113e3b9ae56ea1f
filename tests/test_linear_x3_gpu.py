'''K11 `fd_linear_x3` (the CLIP towers' Linears, encode/clip.py:57-65, 86-100 through transformers' CLIPModel)
against a float64 matmul: the two-term fp16 split with per-row power-of-two scales must be as accurate as the
fp32 Linear it replaces.  Stated bar: max |error| <= 4e-6 of the row scale |x| |w| (torch's own fp32 GEMM sits
at ~1e-6 on the same inputs), for any operand magnitude (no range assumption inside the towers).'''
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(x, w, b, act):
    y = x.double() @ w.double().t()
    if b is not None:
        y = y + b.double()
    if act == 1:
        y = y * torch.sigmoid(1.702 * y)
    elif act == 2:
        y = torch.nn.functional.gelu(y)
    return y


@pytest.mark.parametrize('M,N,K,act,split', [
    (257, 1024, 1024, 0, None), (257, 4096, 1024, 1, None), (257, 1024, 4096, 0, None), (77, 768, 768, 0, None),
    (77, 3072, 768, 1, 1), (77, 768, 3072, 0, 4), (514, 1024, 1024, 2, 2), (1, 64, 64, 0, None), (130, 260, 192, 0, 2)])
def test_linear_x3_matches_float64(native, cuda_dev, M, N, K, act, split):
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator(device=cuda_dev).manual_seed(M + N + K)
        x = torch.randn(M, K, device=cuda_dev, generator=g) * 1.7
        x[0, :6] = torch.tensor([3.0e4, -4.5e3, 1e-4, -1e-7, 0.0, 7.25], device=cuda_dev)   # huge and tiny in one row
        if M > 2:
            x[2] *= 1e-6                                                                     # a tiny row
        w = torch.randn(N, K, device=cuda_dev, generator=g) * 0.03
        w[0] *= 300.0                                                                        # a large weight row
        b = torch.randn(N, device=cuda_dev, generator=g)
        got = native.linear_x3(x, w, b, act=act, split_k=split)  # split-K: in-kernel ordered reduce
        lin = x.double() @ w.double().t() + b.double()
        scale = x.double().norm(dim=1, keepdim=True) * w.double().norm(dim=1)[None] + b.double().abs()[None]
        if act == 0:
            err = ((got.double() - lin).abs() / scale).max().item()
            ref_err = (((x @ w.t() + b).double() - lin).abs() / scale).max().item()
            assert err <= 4e-6, (err, ref_err)
        else:
            want = _ref(x, w, b, act)
            torch.testing.assert_close(got.double(), want, rtol=2e-5, atol=4e-6 * scale.max().item())
        assert tuple(got.shape) == (M, N)
        assert native.lib().fd_linear_x3_flag() == 0
        # no bias, batched leading dims, a shared activation operand
        op = native.x3_split(x)
        g1 = native.linear_x3(x.view(1, M, K), w, None, operand=op)
        g2 = native.linear_x3(x, w * 2, None, operand=op)
        torch.testing.assert_close(g2, g1[0] * 2, rtol=1e-6, atol=1e-30)
        err = ((g1[0].double() - (lin - b.double())).abs() / scale).max().item()
        assert err <= 4e-6, err
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def test_linear_x3_flags_non_finite(native, cuda_dev):
    x = torch.randn(8, 64, device=cuda_dev)
    w = torch.randn(16, 64, device=cuda_dev)
    native.linear_x3(x, w)
    assert native.lib().fd_linear_x3_flag() == 0
    x[3, 5] = float('inf')
    native.linear_x3(x, w)
    assert native.lib().fd_linear_x3_flag() == 1


def test_linear_x3_rejects_cpu_and_bad_shapes(native, cuda_dev):
    with pytest.raises(native.NativeError):
        native.linear_x3(torch.randn(4, 64), torch.randn(8, 64, device=cuda_dev))
    with pytest.raises(native.NativeError):
        native.linear_x3(torch.randn(4, 60, device=cuda_dev), torch.randn(8, 60, device=cuda_dev))


@pytest.mark.parametrize('B,T,H,d,causal', [(1, 257, 16, 64, False), (1, 77, 12, 64, True), (2, 50, 4, 16, True),
                                            (1, 288, 4, 64, False), (1, 100, 2, 96, True), (1, 90, 2, 128, False), (3, 33, 3, 20, False), (1, 1, 2, 8, True)])
def test_attention_f32_matches_float64(native, cuda_dev, B, T, H, d, causal):
    '''K12 `fd_attention_f32` (the attention core of the CLIP towers) against a float64 softmax(QK^T)V, reading q / k / v
    as the column blocks of one fused projection output (token stride 3C).  Bar: 2e-6 absolute on outputs of O(1).'''
    g = torch.Generator(device=cuda_dev).manual_seed(B * 1000 + T)
    C = H * d
    qkv = torch.randn(B, T, 3 * C, device=cuda_dev, generator=g)
    q, k, v = qkv[..., :C], qkv[..., C:2 * C], qkv[..., 2 * C:]
    scale = d ** -0.5
    got = native.attention_f32(q, k, v, H, scale, causal)
    qd, kd, vd = (t.double().reshape(B, T, H, d).transpose(1, 2) for t in (q, k, v))
    s = qd @ kd.transpose(-1, -2) * scale
    if causal:
        s = s.masked_fill(torch.ones(T, T, dtype=torch.bool, device=cuda_dev).triu(1), float('-inf'))
    want = (torch.softmax(s, -1) @ vd).transpose(1, 2).reshape(B, T, C)
    assert tuple(got.shape) == (B, T, C)
    assert (got.double() - want).abs().max().item() <= 2e-6 * max(1.0, want.abs().max().item())
    if T > 1:
        with pytest.raises(native.NativeError):
            native.attention_f32(q, k, v.contiguous(), H, scale, causal)   # v no longer shares q's token stride


@pytest.mark.parametrize('M,K,N', [(257, 1024, 3072), (77, 768, 768), (9, 128, 64), (130, 2048, 132)])
def test_ln_split_and_residual_epilogue(native, cuda_dev, M, K, N):
    '''`fd_linear_x3_split_ln` (LayerNorm fused into the operand split) and the residual epilogue of `fd_linear_x3`:
    residual + Linear(LayerNorm(x)) against the float64 chain, with and without split-K (the K slices of a tile are a
    cluster; they are added in a fixed order over DSMEM: two runs are bit-identical).'''
    g = torch.Generator(device=cuda_dev).manual_seed(M + K)
    x = torch.randn(M, K, device=cuda_dev, generator=g) * 3 + 0.5
    gam = torch.randn(K, device=cuda_dev, generator=g) * 0.3 + 1
    bet = torch.randn(K, device=cuda_dev, generator=g) * 0.2
    w = torch.randn(N, K, device=cuda_dev, generator=g) * 0.05
    b = torch.randn(N, device=cuda_dev, generator=g)
    res = torch.randn(M, N, device=cuda_dev, generator=g)
    y = torch.nn.functional.layer_norm(x.double(), (K,), gam.double(), bet.double(), 1e-5)
    want = res.double() + y @ w.double().t() + b.double()
    scale = y.norm(dim=1, keepdim=True) * w.double().norm(dim=1)[None] + 1.0
    op = native.x3_split_ln(x, gam, bet, 1e-5)
    for sk in (1, 2, 4):
        if K // 64 < sk:
            continue
        got = native.linear_x3(None, w, b, operand=op, rows=M, residual=res, split_k=sk)
        again = native.linear_x3(None, w, b, operand=op, rows=M, residual=res, split_k=sk)
        assert ((got.double() - want).abs() / scale).max().item() <= 6e-6
        assert torch.equal(got, again)
    assert native.lib().fd_linear_x3_flag() == 0
