'''K5 (fused NHWC GroupNorm + time-embedding bias + SiLU) and K6 (GEGLU) vs plain fp32
PyTorch on the same bf16-rounded inputs.  Stated tolerance: bf16 output rounding,
1e-2 relative / 1e-2 absolute on O(1) values.'''
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('N,C,H,W', [(2, 320, 64, 64), (2, 960, 32, 32),
                                     (4, 1280, 8, 8), (1, 2560, 16, 16),
                                     (3, 640, 5, 7), (2, 1920, 32, 32),
                                     # cluster path with 1 / 2 / 4 groups per cluster (16-byte strips
                                     # of 16..240 B), the VAE widths, and shapes that must stream
                                     (1, 128, 64, 64), (1, 256, 32, 32), (2, 512, 16, 16),
                                     (2, 640, 64, 64), (2, 960, 64, 64), (8, 320, 64, 64),
                                     (2, 2560, 8, 8), (1, 512, 61, 3)])
@pytest.mark.parametrize('silu,with_bias', [(True, False), (True, True),
                                            (False, False)])
def test_groupnorm_act_matches_torch(native, cuda_dev, N, C, H, W, silu, with_bias):
    g = torch.Generator(device=cuda_dev).manual_seed(C + H)
    x = (torch.randn(N, C, H, W, device=cuda_dev, generator=g) * 1.7 + 0.4)
    x = x.bfloat16().contiguous(memory_format=torch.channels_last)
    gamma = (torch.randn(C, device=cuda_dev, generator=g) * 0.3 + 1).bfloat16()
    beta = (torch.randn(C, device=cuda_dev, generator=g) * 0.2).bfloat16()
    bias = (torch.randn(N, C, device=cuda_dev, generator=g) * 0.5).bfloat16() \
        if with_bias else None
    y = native.groupnorm_act(x, gamma, beta, 32, 1e-5, silu, bias)
    torch.cuda.synchronize()
    assert y.is_contiguous(memory_format=torch.channels_last) and y.shape == x.shape
    xf = x.float()
    if bias is not None:
        xf = xf + bias.float()[:, :, None, None]
    ref = F.group_norm(xf, 32, gamma.float(), beta.float(), 1e-5)
    if silu:
        ref = F.silu(ref)
    torch.testing.assert_close(y.float(), ref, rtol=1e-2, atol=1e-2)


def test_groupnorm_accepts_nchw_input_and_rejects_fp32(native, cuda_dev):
    x = torch.randn(1, 320, 8, 8, device=cuda_dev).bfloat16()  # NCHW strides
    w = torch.ones(320, device=cuda_dev).bfloat16()
    b = torch.zeros(320, device=cuda_dev).bfloat16()
    y = native.groupnorm_act(x, w, b, 32, 1e-6, False)
    ref = F.group_norm(x.float(), 32, w.float(), b.float(), 1e-6)
    torch.testing.assert_close(y.float(), ref, rtol=1e-2, atol=1e-2)
    with pytest.raises(native.NativeError):
        native.groupnorm_act(x.float(), w, b, 32, 1e-6, False)


@pytest.mark.parametrize('M,Fdim', [(2 * 4096, 1280), (2 * 64, 5120), (77, 2560)])
def test_geglu_matches_torch(native, cuda_dev, M, Fdim):
    g = torch.Generator(device=cuda_dev).manual_seed(M)
    x = (torch.randn(M, 2 * Fdim, device=cuda_dev, generator=g) * 2).bfloat16()
    out = native.geglu(x)
    torch.cuda.synchronize()
    a, gate = x.float().chunk(2, dim=-1)
    torch.testing.assert_close(out.float(), a * F.gelu(gate), rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('N,C,H,W', [(2, 320, 64, 64), (4, 1280, 8, 8)])
def test_add_bias_residual_matches_torch(native, cuda_dev, N, C, H, W):
    g = torch.Generator(device=cuda_dev).manual_seed(C)
    mk = lambda: torch.randn(N, C, H, W, device=cuda_dev, generator=g).bfloat16() \
        .contiguous(memory_format=torch.channels_last)
    x, h = mk(), mk()
    b = torch.randn(C, device=cuda_dev, generator=g).bfloat16()
    y = native.add_bias_residual(x, h, b)
    ref = x.float() + (h.float() + b.float()[None, :, None, None])
    assert y.is_contiguous(memory_format=torch.channels_last)
    torch.testing.assert_close(y.float(), ref, rtol=1e-2, atol=1e-2)
    # h = None: the plain convolution-bias add (conv_in / Downsample2D / Upsample2D), bit-equal to torch's bf16 add, also in place
    want = x + b[None, :, None, None]
    assert torch.equal(native.add_bias_residual(x, None, b), want)
    x2 = x.clone(memory_format=torch.preserve_format)
    assert native.add_bias_residual(x2, None, b, inplace=True).data_ptr() == x2.data_ptr()
    assert torch.equal(x2, want)


@pytest.mark.parametrize('N,Ca,Cb,H', [(2, 640, 320, 64), (2, 1280, 1280, 8), (3, 320, 320, 17), (16, 640, 320, 64)])
def test_concat_channels_equals_torch_cat(native, cuda_dev, N, Ca, Cb, H):
    g = torch.Generator(device=cuda_dev).manual_seed(Ca + Cb)
    a = torch.randn(N, Ca, H, H, device=cuda_dev, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
    b = torch.randn(N, Cb, H, H, device=cuda_dev, generator=g).bfloat16()   # NCHW-contiguous: converted on the way in
    y = native.concat_channels(a, b)
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(y, torch.cat([a, b], dim=1))


@pytest.mark.parametrize('N,C,H,W', [(2, 1280, 8, 8), (2, 640, 32, 32), (3, 320, 5, 9), (16, 640, 32, 32)])
def test_upsample_nearest2x_equals_interpolate(native, cuda_dev, N, C, H, W):
    x = torch.randn(N, C, H, W, device=cuda_dev).bfloat16().contiguous(memory_format=torch.channels_last)
    y = native.upsample_nearest2x(x)
    assert y.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(y, F.interpolate(x, scale_factor=2.0, mode='nearest'))


@pytest.mark.parametrize('N,C,H,silu', [(2, 320, 64, False), (2, 640, 32, True), (2, 1280, 8, False), (4, 960, 16, True),
                                        (16, 320, 64, False)])   # the last one takes the streaming path (K7 + three launches)
def test_add_groupnorm_equals_k7_then_k5(native, cuda_dev, N, C, H, silu):
    '''fd_add_groupnorm_act == fd_add_bias_residual followed by fd_groupnorm_act, bit for bit (sum and normalised output).'''
    g = torch.Generator(device=cuda_dev).manual_seed(N * C + H)
    mk = lambda: (torch.randn(N, C, H, H, device=cuda_dev, generator=g) * 1.5).bfloat16().contiguous(memory_format=torch.channels_last)
    x, h = mk(), mk()
    rb = (torch.randn(C, device=cuda_dev, generator=g) * 0.3).bfloat16()
    gam = (torch.randn(C, device=cuda_dev, generator=g) * 0.3 + 1).bfloat16()
    bet = (torch.randn(C, device=cuda_dev, generator=g) * 0.2).bfloat16()
    y_ref = native.add_bias_residual(x, h, rb)
    z_ref = native.groupnorm_act(y_ref, gam, bet, 32, 1e-6, silu)
    y, z = native.add_groupnorm_act(x, h, rb, gam, bet, 32, 1e-6, silu)
    assert torch.equal(y, y_ref)
    assert torch.equal(z, z_ref)


def test_groupnorm_is_bit_reproducible(native, cuda_dev):
    x = torch.randn(2, 640, 32, 32, device=cuda_dev).bfloat16() \
        .contiguous(memory_format=torch.channels_last)
    w = torch.randn(640, device=cuda_dev).bfloat16()
    b = torch.randn(640, device=cuda_dev).bfloat16()
    a = native.groupnorm_act(x, w, b, 32, 1e-5, True)
    for _ in range(3):
        assert torch.equal(a, native.groupnorm_act(x, w, b, 32, 1e-5, True))


@pytest.mark.parametrize('M,C', [(8192, 320), (2048, 640), (133, 1280)])
@pytest.mark.parametrize('with_y', [True, False])
def test_add_layernorm_matches_torch(native, cuda_dev, M, C, with_y):
    g = torch.Generator(device=cuda_dev).manual_seed(M + C)
    x = (torch.randn(M, C, device=cuda_dev, generator=g) * 2 + 0.3).bfloat16()
    y = torch.randn(M, C, device=cuda_dev, generator=g).bfloat16() if with_y else None
    w = (torch.randn(C, device=cuda_dev, generator=g) * 0.3 + 1).bfloat16()
    b = (torch.randn(C, device=cuda_dev, generator=g) * 0.2).bfloat16()
    total, norm = native.add_layernorm(x, y, w, b, 1e-5)
    s = (x.float() + y.float()).bfloat16() if with_y else x
    assert torch.equal(total, s)
    ref = F.layer_norm(s.float(), (C,), w.float(), b.float(), 1e-5)
    torch.testing.assert_close(norm.float(), ref, rtol=1e-2, atol=1e-2)
    if with_y:   # the sum written into a column block of a wider matrix: same bits, nothing else touched
        wide = torch.full((M, 5 * C), 7.0, device=cuda_dev).bfloat16()
        total2, norm2 = native.add_layernorm(x, y, w, b, 1e-5, sum_out=wide[:, 4 * C:])
        assert total2.data_ptr() == wide[:, 4 * C:].data_ptr()
        assert torch.equal(wide[:, 4 * C:], s) and torch.equal(norm2, norm)
        assert (wide[:, :4 * C] == 7.0).all()
