'''Host logic of flexdiffuse_b200/schedulers.py without a GPU: the planners (coefficients, history
ring, PLMS warm-up, LMS quadrature, step-index conventions of flex.py:262-287) are exercised with the
ONE K4 launch per step replaced by the kernel's documented formula in torch,

    eps = u + g (c - u);  e = sum_i w[i] * {eps, hist...};  x' = a x + b e (+ c_noise n);  scaled = x' * in_scale

and compared with the oracle's restatement of the diffusers 0.3.0 schedulers on identical eps
sequences.  (The CUDA kernel itself is checked against the same formula in tests/test_k4_cfg_sched.py.)
Tolerance: 2e-5 relative to the latent scale after up to 51 chained steps, as on the GPU.'''
import pytest
import torch

from flexdiffuse_b200 import schedulers as prod
from oracle import loop_oracle as lo


def _torch_launch(self, plan, eps_uncond, eps_cond, guidance, use_cfg, noise, out, scaled_out,
                  eps_slot):
    eps = eps_cond.float()
    if use_cfg:
        u = eps_uncond.float()
        eps = u + float(guidance) * (eps - u)
    terms = [eps] + [h.float() for h in plan.hist]
    assert len(plan.w) == len(terms) <= 4
    e = sum(float(w) * t for w, t in zip(plan.w, terms))
    x = float(plan.a) * plan.x_src + float(plan.b) * e
    if noise is not None:
        x = x + float(plan.c_noise) * noise
    out.copy_(x)
    if eps_slot is not None:
        eps_slot.copy_(eps)
    if scaled_out is not None:
        scaled_out.copy_((x * float(plan.in_scale)).to(scaled_out.dtype))


@pytest.fixture(autouse=True)
def _no_gpu_launch(monkeypatch):
    monkeypatch.setattr(prod._SchedulerBase, '_launch', _torch_launch)


def _chain(name, steps, t_start, cfg):
    ps, os_ = getattr(prod, name)(), getattr(lo, name)()
    ps.set_timesteps(steps)
    os_.set_timesteps(steps)
    assert torch.equal(ps.timesteps.double(), os_.timesteps.double())
    is_lms = name == 'LMSDiscreteScheduler'
    g = torch.Generator().manual_seed(100 * steps + t_start)
    x = torch.randn(2, 4, 16, 16, generator=g)
    if is_lms:
        assert torch.allclose(ps.sigmas, os_.sigmas)
        x = x * os_.sigmas[0]
    xp, xo, worst = x.clone(), x.clone(), 0.0
    for i, t in enumerate(os_.timesteps[t_start:]):
        u = torch.randn(2, 4, 16, 16, generator=g)
        c = torch.randn(2, 4, 16, 16, generator=g)
        t_index = t_start + i if is_lms else int(t)
        xo = os_.step(u + 7.5 * (c - u) if cfg else c, t_index, xo).prev_sample
        if cfg:
            xp = ps.fused_step(u, c, 7.5, True, t_index, xp).prev_sample
        else:
            xp = ps.step(c, t_index, xp).prev_sample
        worst = max(worst, ((xp - xo).abs().max() / xo.abs().max()).item())
    return worst


@pytest.mark.parametrize('name', ['DDIMScheduler', 'PNDMScheduler', 'LMSDiscreteScheduler'])
@pytest.mark.parametrize('steps,t_start', [(50, 0), (50, 20), (7, 0)])
@pytest.mark.parametrize('cfg', [True, False])
def test_planners_match_oracle_chain(name, steps, t_start, cfg):
    assert _chain(name, steps, t_start, cfg) < 2e-5


def test_euler_is_lms_order_1_and_scales_next_input():
    ps, os_ = prod.EulerDiscreteScheduler(), lo.LMSDiscreteScheduler()
    ps.set_timesteps(20)
    os_.set_timesteps(20)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, 4, 8, 8, generator=g) * os_.sigmas[0]
    xp, xo = x.clone(), x.clone()
    scaled = torch.empty(1, 4, 8, 8)
    for i in range(20):
        e = torch.randn(1, 4, 8, 8, generator=g)
        xo = os_.step(e, i, xo, order=1).prev_sample
        xp = ps.fused_step(None, e, 1.0, False, i, xp, scaled_out=scaled).prev_sample
        assert ((xp - xo).abs().max() / xo.abs().max()).item() < 2e-5
        nxt = ps.sigmas[i + 1].item()
        torch.testing.assert_close(scaled, xp / (nxt * nxt + 1)**0.5, rtol=1e-6, atol=1e-6)  # flex.py:272-274


def test_ddim_eta_noise_coefficient_and_fp32_requirement():
    ps, os_ = prod.DDIMScheduler(), lo.DDIMScheduler()
    ps.set_timesteps(50)
    os_.set_timesteps(50)
    for t in (981, 501, 1):
        _, _, sigma = ps.coefficients(t, eta=0.7)
        assert abs(sigma - 0.7 * float(os_._get_variance(t, t - 20))**0.5) < 1e-7
    with pytest.raises(Exception):  # latents must stay fp32 (flex.py keeps them fp32)
        ps.step(torch.zeros(1, 4, 8, 8), 981, torch.zeros(1, 4, 8, 8, dtype=torch.bfloat16))
