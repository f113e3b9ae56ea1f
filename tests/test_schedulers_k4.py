'''Scheduler parity (GPU): the product's host state machines + ONE K4 launch per step
(flexdiffuse_b200/schedulers.py) against the oracle's formula-by-formula restatement of
the diffusers 0.3.0 schedulers (oracle/diffusers_shim), on identical eps sequences.
fp32 throughout; stated tolerance: 2e-5 relative to the latent scale after up to 51
chained steps.'''
import pytest
import torch

from flexdiffuse_b200 import schedulers as prod
from oracle import loop_oracle as lo

pytestmark = pytest.mark.gpu


def _run(name, steps, t_start, dev, cfg, eta=0.0):
    ps, os_ = getattr(prod, name)(), getattr(lo, name)()
    ps.set_timesteps(steps)
    os_.set_timesteps(steps)
    assert torch.equal(ps.timesteps.double(), os_.timesteps.double())
    is_lms = name == 'LMSDiscreteScheduler'
    g = torch.Generator().manual_seed(steps + t_start)
    x = torch.randn(2, 4, 64, 64, generator=g)
    if is_lms:
        x = x * os_.sigmas[0]
        assert torch.allclose(ps.sigmas, os_.sigmas)
    xp, xo = x.to(dev), x.clone()
    worst = 0.0
    for i, t in enumerate(os_.timesteps[t_start:]):
        u = torch.randn(2, 4, 64, 64, generator=g)
        c = torch.randn(2, 4, 64, 64, generator=g)
        t_index = t_start + i if is_lms else int(t)
        eps = u + 7.5 * (c - u) if cfg else c
        kw = {'eta': eta} if name == 'DDIMScheduler' else {}
        if eta:
            torch.manual_seed(1000 + i)
        xo = os_.step(eps, t_index, xo, **kw).prev_sample
        if eta:
            torch.manual_seed(1000 + i)
            kw['generator'] = None
        if cfg:
            xp = ps.fused_step(u.to(dev), c.to(dev), 7.5, True, t_index, xp,
                               **kw).prev_sample
        else:
            kw.pop('generator', None)
            xp = ps.step(c.to(dev), t_index, xp, **kw).prev_sample
        err = ((xp.cpu() - xo).abs().max() / xo.abs().max()).item()
        worst = max(worst, err)
    return worst


@pytest.mark.parametrize('name', ['DDIMScheduler', 'PNDMScheduler',
                                  'LMSDiscreteScheduler'])
@pytest.mark.parametrize('steps,t_start', [(50, 0), (50, 20), (30, 12), (7, 0)])
@pytest.mark.parametrize('cfg', [True, False])
def test_scheduler_chain_matches_oracle(native, cuda_dev, name, steps, t_start, cfg):
    assert _run(name, steps, t_start, cuda_dev, cfg) < 2e-5


def test_ddim_eta_draws_noise(native, cuda_dev):
    '''eta > 0: the variance-noise coefficient matches (noise itself comes from
    different RNG streams on CPU / GPU, so compare the deterministic part only).'''
    ps, os_ = prod.DDIMScheduler(), lo.DDIMScheduler()
    ps.set_timesteps(50)
    os_.set_timesteps(50)
    for t in (981, 501, 1):
        a, b, sigma = ps.coefficients(t, eta=0.7)
        prev_t = t - 20
        assert abs(sigma - 0.7 * float(os_._get_variance(t, prev_t))**0.5) < 1e-7


def test_lms_scaled_model_input(native, cuda_dev):
    '''K4's `scaled_out` is the next step's  latents / sqrt(sigma^2 + 1)  (flex.py:272-274).'''
    ps = prod.LMSDiscreteScheduler()
    ps.set_timesteps(20)
    x = torch.randn(1, 4, 64, 64, device=cuda_dev) * ps.sigmas[0]
    eps = torch.randn(1, 4, 64, 64, device=cuda_dev)
    scaled = torch.empty(1, 4, 64, 64, device=cuda_dev, dtype=torch.bfloat16)
    out = ps.fused_step(None, eps, 1.0, False, 0, x, scaled_out=scaled).prev_sample
    want = out / (ps.sigmas[1].item()**2 + 1)**0.5
    torch.testing.assert_close(scaled.float(), want, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('steps,t_start', [(50, 0), (30, 12)])
def test_euler_is_lms_order_1(native, cuda_dev, steps, t_start):
    '''EulerDiscreteScheduler (extension; diffusers 0.3.0 has none) == the oracle's LMS
    restatement stepped with order=1, same 2e-5 chained tolerance.'''
    ps, os_ = prod.EulerDiscreteScheduler(), lo.LMSDiscreteScheduler()
    ps.set_timesteps(steps)
    os_.set_timesteps(steps)
    g = torch.Generator().manual_seed(steps)
    x = torch.randn(2, 4, 64, 64, generator=g) * os_.sigmas[0]
    xp, xo = x.to(cuda_dev), x.clone()
    for i in range(t_start, steps):
        u = torch.randn(2, 4, 64, 64, generator=g)
        c = torch.randn(2, 4, 64, 64, generator=g)
        xo = os_.step(u + 7.5 * (c - u), i, xo, order=1).prev_sample
        xp = ps.fused_step(u.to(cuda_dev), c.to(cuda_dev), 7.5, True, i,
                           xp).prev_sample
        assert ((xp.cpu() - xo).abs().max() / xo.abs().max()).item() < 2e-5
