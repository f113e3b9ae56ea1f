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
