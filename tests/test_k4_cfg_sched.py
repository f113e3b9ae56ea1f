'''K4 parity: fused CFG + scheduler update vs the plain fp32 PyTorch expressions
(pipeline/guide.py:61-63 + the linear form of diffusers' scheduler.step).
Floating point kernel => fp32 torch reference, tolerance stated below.'''
import pytest
import torch

pytestmark = pytest.mark.gpu

RTOL, ATOL = 2e-6, 2e-6  # fp32 elementwise; FMA contraction only


def _ref(u, c, x, hist, noise, k):
    u, c = u.float(), c.float()
    eps = u + k.guidance * (c - u) if k.use_cfg else c
    e = k.w[0] * eps
    for i, h in enumerate(hist):
        e = e + k.w[i + 1] * h
    xn = k.a * x + k.b * e
    if noise is not None:
        xn = xn + k.c_noise * noise
    return xn, eps


@pytest.mark.parametrize('eps_dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('nh', [0, 1, 2, 3])
@pytest.mark.parametrize('batch', [1, 16])
def test_k4_matches_torch(native, cuda_dev, eps_dtype, nh, batch):
    g = torch.Generator(device='cpu').manual_seed(nh * 10 + batch)
    shape = (batch, 4, 64, 64)
    mk = lambda: torch.randn(shape, generator=g).to(cuda_dev)
    u, c = mk().to(eps_dtype), mk().to(eps_dtype)
    x = mk()
    hist = [mk() for _ in range(nh)]
    noise = mk() if nh == 0 else None
    k = native.SchedCoeffs()
    k.guidance, k.use_cfg = 7.5, 1
    w = [55 / 24, -59 / 24, 37 / 24, -9 / 24]
    for i in range(4):
        k.w[i] = w[i] if i <= nh else 0.0
    if nh == 0:
        k.w[0] = 1.0
    k.a, k.b, k.c_noise, k.in_scale = 1.0123, -0.0456, 0.3, 0.87
    x_out = torch.empty_like(x)
    eps_out = torch.empty_like(x)
    scaled = torch.empty(shape, dtype=torch.bfloat16, device=cuda_dev)
    native.cfg_sched_step(u, c, x, k, x_out, hist=hist, noise=noise,
                          eps_out=eps_out, scaled_out=scaled)
    torch.cuda.synchronize()
    xr, er = _ref(u, c, x, hist, noise, k)
    torch.testing.assert_close(x_out, xr, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(eps_out, er, rtol=RTOL, atol=ATOL)
    torch.testing.assert_close(scaled.float(), (xr * k.in_scale), rtol=1e-2,
                               atol=1e-2)


def test_k4_no_cfg_inplace(native, cuda_dev):
    x = torch.randn(2, 4, 64, 64, device=cuda_dev)
    c = torch.randn_like(x)
    k = native.SchedCoeffs()
    k.guidance, k.use_cfg = 1.0, 0
    k.w[0] = 1.0
    k.a, k.b = 0.99, 0.02
    want = k.a * x + k.b * c
    native.cfg_sched_step(None, c, x, k, x)  # in place
    torch.cuda.synchronize()
    torch.testing.assert_close(x, want, rtol=RTOL, atol=ATOL)


def test_k4_rejects_bad_args(native, cuda_dev):
    x = torch.randn(6, device=cuda_dev)  # not a multiple of 4
    k = native.SchedCoeffs()
    with pytest.raises(native.NativeError):
        native.cfg_sched_step(None, x, x, k, torch.empty_like(x))
    with pytest.raises(native.NativeError):  # CPU tensors: no fallback
        native.cfg_sched_step(None, torch.randn(8), torch.randn(8), k,
                              torch.empty(8))
