'''K10 `fd_image_tail_u8` (the uint8 NHWC tail of the VAE decode, pipeline/flex.py:119-124 +
numpy_to_pil) and the CUDA-graphed CLIP towers of `CLIPEncoder` (encode/clip.py:47-100).

K10 is byte work: bit-exact against the reference's expression evaluated by torch / numpy on the same
decoder output.  A graph replay runs the same kernels in the same order as the eager tower: equal to
1e-6.'''
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _reference_bytes(image: torch.Tensor) -> np.ndarray:
    x = (image / 2 + 0.5).clamp(0, 1)                       # flex.py:119
    x = x.float().cpu().permute(0, 2, 3, 1).numpy()         # flex.py:121-122 (fp32 under autocast too)
    return (x * 255).round().astype('uint8')                # diffusers numpy_to_pil


@pytest.mark.parametrize('dtype', [torch.bfloat16, torch.float32])
@pytest.mark.parametrize('shape', [(2, 3, 64, 48), (1, 3, 512, 512), (3, 3, 5, 7)])
def test_image_tail_is_bit_exact(native, cuda_dev, dtype, shape):
    g = torch.Generator(device=cuda_dev).manual_seed(sum(shape))
    x = (torch.randn(shape, device=cuda_dev, generator=g) * 1.3).to(dtype)
    x.view(-1)[:6] = torch.tensor([-1.0, 1.0, 0.0, 3.0, -3.0, 0.00390625], device=cuda_dev).to(dtype)
    x = x.contiguous(memory_format=torch.channels_last)
    got = native.image_tail_u8(x)
    assert got.dtype == torch.uint8 and tuple(got.shape) == (shape[0], shape[2], shape[3], 3)
    want = _reference_bytes(x)
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    assert want.min() == 0 and want.max() == 255  # both clamps exercised
    # NCHW-strided input is re-laid out, not misread
    np.testing.assert_array_equal(native.image_tail_u8(x.contiguous()).cpu().numpy(), want)


def test_pipeline_pil_and_uint8_outputs_match_reference_tail(native, cuda_dev):
    from flexdiffuse_b200 import schedulers as prod
    from flexdiffuse_b200.pipeline.flex import FlexPipeline
    from tests.model_helpers import models
    unet, vae, _, _ = models(str(cuda_dev))
    pipe = FlexPipeline(vae, None, None, unet, prod.DDIMScheduler())
    lat = torch.randn(2, 4, 16, 16, device=cuda_dev, generator=torch.Generator(device=cuda_dev).manual_seed(3))
    want = _reference_bytes(vae.decode(1 / 0.18215 * lat).sample)
    u8 = pipe.decode(lat, 'uint8')
    assert isinstance(u8, np.ndarray) and u8.dtype == np.uint8 and u8.shape == (2, 128, 128, 3)
    np.testing.assert_array_equal(u8, want)
    pil = pipe.decode(lat, 'pil')
    assert len(pil) == 2 and pil[0].size == (128, 128)
    np.testing.assert_array_equal(np.asarray(pil[1]), want[1])
    f = pipe.decode(lat, 'np')  # the float path is unchanged
    np.testing.assert_array_equal((f * 255).round().astype('uint8'), want)
    with pytest.raises(native.NativeError):
        native.image_tail_u8(torch.zeros(1, 3, 8, 8))  # CPU tensor: no fallback


def test_clip_towers_graph_replay_equals_eager(native, cuda_dev):
    from flexdiffuse_b200.encode.clip import CLIPEncoder
    from tests.encode_helpers import FakeTok, test_images, tiny_clip
    clip = tiny_clip().to(cuda_dev)
    eager = CLIPEncoder(clip, FakeTok(), cuda_graph=False)
    graphed = CLIPEncoder(clip, FakeTok(), cuda_graph=True)
    img = test_images()[0]
    with torch.no_grad():
        for prompt in ('a red fox', ['a dog', 'two birds on a wire'], 'a red fox'):
            a, b = eager.prompt(prompt), graphed.prompt(prompt)
            torch.testing.assert_close(b, a, rtol=1e-6, atol=1e-6)
        for _ in range(2):  # second call replays
            a, b = eager.image(img), graphed.image(img)
            torch.testing.assert_close(b, a, rtol=1e-6, atol=1e-6)
        first = graphed.prompt('a red fox')
        second = graphed.prompt('a blue bird')
    assert any(g for g in graphed._graphs.values()), 'no tower was captured'
    assert not torch.equal(first, second)           # a replay does not alias the previous result
    assert first.data_ptr() != second.data_ptr()


@pytest.mark.parametrize('M,N,K', [(257, 768, 1024), (514, 768, 1024), (100, 64, 128), (129, 260, 192)])
def test_visual_projection_is_fp32_accurate(native, cuda_dev, M, N, K):
    '''K1P `fd_visual_projection` (encode/clip.py:100) against a float64 matmul: the two-term fp16
    split must be as accurate as the fp32 Linear it replaces (stated bar: max error <= 4e-6 of the
    row scale |h| |w|; torch's own fp32 GEMM sits at ~1e-6 on the same inputs).'''
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        g = torch.Generator(device=cuda_dev).manual_seed(M + N + K)
        h = torch.randn(M, K, device=cuda_dev, generator=g) * 1.7
        h[0, :7] = torch.tensor([30.0, -45.0, 1e-4, -1e-5, 0.0, 7.25, 500.0], device=cuda_dev)
        w = torch.randn(N, K, device=cuda_dev, generator=g) * 0.03
        got = native.visual_projection(h, w)
        want = h.double() @ w.double().t()
        scale = h.double().norm(dim=1, keepdim=True) * w.double().norm(dim=1)[None]
        err = ((got.double() - want).abs() / scale).max().item()
        ref_err = (((h @ w.t()).double() - want).abs() / scale).max().item()
        assert tuple(got.shape) == (M, N)
        assert err <= 4e-6, (err, ref_err)
        assert native.lib().fd_visual_projection_range_flag() == 0
        # batched leading dims, and a changed weight tensor invalidates the cached planes
        got3 = native.visual_projection(h.view(1, M, K), w * 2)
        torch.testing.assert_close(got3[0], got * 2, rtol=1e-5, atol=1e-6)
        # out-of-range activations are flagged, not silently wrong
        h2 = h.clone()
        h2[1, 3] = 5000.0
        native.visual_projection(h2, w * 2)
        assert native.lib().fd_visual_projection_range_flag() == 1
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def test_clip_encoder_image_uses_projection_kernel(native, cuda_dev):
    '''CLIP ViT-L/14 widths (1024 -> 768): CLIPEncoder.image goes through K1P and matches the plain
    fp32 Linear to 1e-5 of the output scale.'''
    from transformers import CLIPConfig, CLIPModel
    from flexdiffuse_b200.encode.clip import CLIPEncoder
    from tests.encode_helpers import FakeTok, test_images
    cfg = CLIPConfig(
        text_config=dict(hidden_size=64, intermediate_size=128, num_hidden_layers=1, num_attention_heads=4,
                         vocab_size=1000, max_position_embeddings=77),
        vision_config=dict(hidden_size=1024, intermediate_size=512, num_hidden_layers=1, num_attention_heads=16,
                           image_size=224, patch_size=14),
        projection_dim=768)
    torch.manual_seed(5)
    clip = CLIPModel(cfg).eval().requires_grad_(False).to(cuda_dev)
    enc = CLIPEncoder(clip, FakeTok(), cuda_graph=False)
    img = test_images()[0]
    before = native.LAUNCHES
    with torch.no_grad():
        got = enc.image(img)
        assert native.LAUNCHES > before
        # the same tower with the stock Linear
        from torchvision.transforms.functional import InterpolationMode, center_crop, normalize, resize
        from flexdiffuse_b200.encode.clip import _CLIP_MEAN, _CLIP_STD, preprocess
        x = preprocess(img)
        side = min(x.shape[-2:])
        x = normalize(resize(center_crop(x, [side, side]), [224, 224], interpolation=InterpolationMode.BICUBIC,
                             antialias=True), list(_CLIP_MEAN), list(_CLIP_STD)).to(cuda_dev)
        vm = clip.vision_model
        hid = vm.encoder(inputs_embeds=vm.pre_layrnorm(vm.embeddings(x)), return_dict=True)[0]
        want = clip.visual_projection(vm.post_layernorm(hid))
    assert tuple(got.shape) == (1, 257, 768)
    # both are fp32-accurate products of the same operands: they differ by rounding-order noise,
    # ~1e-6 of the output scale (elementwise rtol would be meaningless for outputs near zero)
    assert (got - want).abs().max().item() <= 1e-5 * want.abs().max().item()


def test_tf32_towers_are_opt_in_and_close(native, cuda_dev):
    '''tf32=True runs the towers' GEMMs in TF32 (stated: embeddings within 5e-3 relative L2 of the
    fp32 towers); the default stays exact fp32 and leaves the global matmul flag untouched.'''
    from flexdiffuse_b200.encode.clip import CLIPEncoder
    from tests.encode_helpers import FakeTok, test_images, tiny_clip
    clip = tiny_clip().to(cuda_dev)
    exact = CLIPEncoder(clip, FakeTok(), cuda_graph=False, x3=False)
    fast = CLIPEncoder(clip, FakeTok(), tf32=True, x3=False)
    flag = torch.backends.cuda.matmul.allow_tf32
    with torch.no_grad():
        a, b = exact.image(test_images()[0]), fast.image(test_images()[0])
        c, d = exact.prompt('a red fox'), fast.prompt('a red fox')
    assert torch.backends.cuda.matmul.allow_tf32 == flag
    assert ((a - b).norm() / a.norm()).item() < 5e-3
    assert ((c - d).norm() / c.norm()).item() < 5e-3


def _towers_vs_transformers(native, cuda_dev, clip, tol):
    from flexdiffuse_b200.encode.clip import CLIPEncoder
    from tests.encode_helpers import FakeTok, test_images
    ours = CLIPEncoder(clip, FakeTok(), cuda_graph=True)                 # K11 towers (default), graphed
    stock = CLIPEncoder(clip, FakeTok(), cuda_graph=False, x3=False)     # transformers' own fp32 forward
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            before = native.LAUNCHES
            for img in test_images()[:2]:
                a, b = ours.image(img), stock.image(img)
                assert (a - b).abs().max().item() <= tol * b.abs().max().item(), ((a - b).abs().max().item(), b.abs().max().item())
            for prompt in ('a red fox', 'an astronaut riding a horse on the moon, oil painting'):
                c, d = ours.prompt(prompt), stock.prompt(prompt)
                assert (c - d).abs().max().item() <= tol * d.abs().max().item()
            assert native.LAUNCHES > before
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old
    assert native.lib().fd_linear_x3_flag() == 0
    assert any(g for g in ours._graphs.values()), 'no tower was captured'


def test_x3_towers_match_transformers_fp32_tiny(native, cuda_dev):
    '''CLIPEncoder's default towers (every Linear on K11, attention on SDPA) against transformers' own fp32
    forward of the same weights (encode/clip.py:57-65, 86-100): fp32-level agreement, 2e-5 of the output
    scale (both sides are fp32-accurate; they differ by summation order).'''
    from tests.encode_helpers import tiny_clip
    _towers_vs_transformers(native, cuda_dev, tiny_clip().to(cuda_dev), 2e-5)


def test_x3_towers_match_transformers_fp32_vit_l14_widths(native, cuda_dev):
    '''Same at CLIP ViT-L/14 widths (vision 1024 / 4096 / 16 heads, text 768 / 3072 / 12 heads) with 3 layers
    each: the shapes K11 runs in production (q / k / v / o 1024x1024, fc1 4096x1024, fc2 1024x4096 split-K).'''
    from transformers import CLIPConfig, CLIPModel
    cfg = CLIPConfig(
        text_config=dict(hidden_size=768, intermediate_size=3072, num_hidden_layers=3, num_attention_heads=12,
                         vocab_size=49408, max_position_embeddings=77),
        vision_config=dict(hidden_size=1024, intermediate_size=4096, num_hidden_layers=3, num_attention_heads=16,
                           image_size=224, patch_size=14),
        projection_dim=768)
    torch.manual_seed(11)
    clip = CLIPModel(cfg).eval().requires_grad_(False).to(cuda_dev)
    _towers_vs_transformers(native, cuda_dev, clip, 2e-5)
