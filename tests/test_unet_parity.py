'''UNet parity (GPU): the product UNet (bf16, cuDNN/cuBLAS + K2 cached K/V + K3 tcgen05
cross-attention) against the oracle's fp32 restatement of diffusers' UNet2DConditionModel
(oracle/unet_oracle.py) on the same random-init weights and inputs.

Stated bf16 tolerance for a per-step noise prediction: relative L2 error <= 3e-2
(bf16 has 8 mantissa bits; ~60 sequential layers).  The attn2-only comparison isolates
K2 + K3 and is held to 1.5e-2.'''
import pytest
import torch

from oracle import unet_oracle as U
from tests.model_helpers import models, rel_l2

pytestmark = pytest.mark.gpu

EPS_TOL = 3e-2
ATTN_TOL = 1.5e-2


@pytest.fixture(scope='module')
def fp32_math():
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old


@pytest.mark.parametrize('fused', [True, False])
def test_attn2_layers_match_oracle(native, cuda_dev, fp32_math, fused, monkeypatch):
    '''Each of the 16 cross-attention sites: K2 cache + K3F (one launch: to_q, attention, to_out)
    -- or K3 between two cuBLAS GEMMs -- vs fp32 to_q/to_k/to_v/softmax/to_out.'''
    from flexdiffuse_b200 import unet as unet_mod
    monkeypatch.setattr(unet_mod, 'FUSED_ATTN2', fused)
    unet, _, sd, _ = models(str(cuda_dev))
    g = torch.Generator(device=cuda_dev).manual_seed(3)
    ctx = torch.randn(3, 77, 768, device=cuda_dev, generator=g)
    kv = unet.build_kv_cache(ctx)
    idx = torch.tensor([0, 2, 1, 2], dtype=torch.int32, device=cuda_dev)
    names = [n for n, m in unet.named_modules() if n.endswith('attn2')]
    assert len(names) == 16
    for name, mod in zip(names, unet.cross_attentions()):
        n_q = {320: 1024, 640: 256, 1280: 64}[mod.dim]
        x = torch.randn(4, n_q, mod.dim, device=cuda_dev, generator=g)
        before = native.LAUNCHES
        got = mod(x.bfloat16(), kv, idx)
        assert native.LAUNCHES - before == 1  # K3F or K3: one native launch either way
        assert native.k3f_status() == [0, 0, 0, 0]
        want = U.attention(U._sub(sd, name + '.'), x.bfloat16().float(),
                           ctx.bfloat16().float()[idx.long()])
        assert rel_l2(got, want) < ATTN_TOL, name


@pytest.mark.parametrize('hw,merged', [(32, True), (64, True), (32, False)])
def test_noise_prediction_matches_oracle(native, cuda_dev, fp32_math, hw, merged, monkeypatch):
    '''`merged`: ff.net[2] + residual + proj_out as the one K = 5C GEMM (the default) or as the separate launches.'''
    from flexdiffuse_b200 import unet as unet_mod
    monkeypatch.setattr(unet_mod, 'MERGED_OUT_GEMM', merged)
    unet, _, sd, _ = models(str(cuda_dev))
    g = torch.Generator(device=cuda_dev).manual_seed(hw)
    x = torch.randn(2, 4, hw, hw, device=cuda_dev, generator=g)
    ctx = torch.randn(2, 77, 768, device=cuda_dev, generator=g)
    for t in (981, 401, 1):
        got = unet(x, t, encoder_hidden_states=ctx).sample
        want = U.unet_forward(sd, x.bfloat16().float(), t, ctx.bfloat16().float())
        assert torch.isfinite(got.float()).all()
        assert want.abs().mean() > 1e-3  # a degenerate (all ~0) output would prove nothing
        assert rel_l2(got, want) < EPS_TOL, (t, rel_l2(got, want))


def test_merged_output_gemm_equals_separate_launches(native, cuda_dev, fp32_math, monkeypatch):
    '''Every SpatialTransformer: [h | x3] [Wo W2 | Wo]^T + (Wo b2 + bo) against ff.net[2] -> add -> proj_out on the same
    input, both against the fp32 evaluation of the same tail; the merged form must not be further from it than 1.5x the
    separate one (it skips two bf16 roundings and adds the rounding of Wo W2).'''
    from flexdiffuse_b200 import unet as unet_mod
    unet, _, _, _ = models(str(cuda_dev))
    g = torch.Generator(device=cuda_dev).manual_seed(5)
    ctx = torch.randn(2, 77, 768, device=cuda_dev, generator=g)
    kv = unet.build_kv_cache(ctx)
    idx = torch.tensor([0, 1], dtype=torch.int32, device=cuda_dev)
    sts = [m for m in unet.modules() if isinstance(m, unet_mod.SpatialTransformer)]
    assert len(sts) == 16
    for st in sts:
        C = st.proj_out.out_channels
        hw = {320: 32, 640: 16, 1280: 8}[C]
        x = torch.randn(2, C, hw, hw, device=cuda_dev, generator=g).bfloat16().contiguous(memory_format=torch.channels_last)
        outs = {}
        for merged in (True, False):
            monkeypatch.setattr(unet_mod, 'MERGED_OUT_GEMM', merged)
            before = native.LAUNCHES
            outs[merged] = st(x, kv, idx).float()
            outs[merged, 'launches'] = native.LAUNCHES - before
        assert rel_l2(outs[True], outs[False]) < 1e-2, C
        w_m, b_m = st._merged_out()
        lin2 = st.transformer_blocks[0].ff.net[2]
        wo = st.proj_out.weight.reshape(C, C).double()
        torch.testing.assert_close(w_m.double(), torch.cat([wo @ lin2.weight.double(), wo], 1), rtol=1e-2, atol=1e-3)
        torch.testing.assert_close(b_m.double(), wo @ lin2.bias.double() + st.proj_out.bias.double(), rtol=1e-2, atol=1e-3)


def test_cached_context_equals_inline_projection(native, cuda_dev):
    '''unet(x, t, encoder_hidden_states=ctx) == unet(x, t, kv_cache=K2(ctx)) bit for bit, and
    a shared uncond row (ctx_index) equals a duplicated context.'''
    unet, _, _, _ = models(str(cuda_dev))
    g = torch.Generator(device=cuda_dev).manual_seed(9)
    x = torch.randn(2, 4, 32, 32, device=cuda_dev, generator=g)
    ctx = torch.randn(1, 77, 768, device=cuda_dev, generator=g)
    a = unet(x, 500, encoder_hidden_states=torch.cat([ctx, ctx])).sample
    kv = unet.build_kv_cache(ctx)
    idx = torch.zeros(2, dtype=torch.int32, device=cuda_dev)
    b = unet(x, 500, kv_cache=kv, ctx_index=idx).sample
    assert torch.equal(a, b)
