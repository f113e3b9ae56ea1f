'''The reference's error conventions on the path, checked where they are raised BEFORE any
device work, so they hold without a GPU: guidance.py:397-401 (bad `prompt` type, neither prompt nor
guide), pipeline/flex.py:170-172 (strength outside [0,1]), pipeline/guide.py:36 (base `noise_pred`),
composition/guide.py batch restriction, and this repo's own rule that nothing falls back to the CPU.'''
import pytest
import torch

from flexdiffuse_b200 import _native, schedulers
from flexdiffuse_b200.guidance import Guide, Tweener
from flexdiffuse_b200.pipeline.flex import FlexPipeline
from flexdiffuse_b200.pipeline.guide import GuideBase, SimpleGuide
from tests.encode_helpers import FakeTok, tiny_clip


class _Enc:
    def prompt(self, p):
        n = 1 if isinstance(p, str) else len(p)
        return torch.zeros(n, 77, 768)


@pytest.fixture(scope='module')
def guide():
    torch.set_num_threads(1)
    return Guide(tiny_clip(), FakeTok(), device='cpu')


def test_embeds_rejects_bad_prompt_type(guide):
    with pytest.raises(ValueError, match='`prompt` has to be of type'):
        guide.embeds(prompt=123)


def test_embeds_needs_prompt_or_guide(guide):
    with pytest.raises(ValueError, match='No prompt, or guide image provided'):
        guide.embeds(prompt='   ', guide=None)
    with pytest.raises(ValueError):
        guide.embeds(prompt=['', ' '], guide=None)


def test_text_only_embeds_is_the_encoder_output(guide):
    # guidance.py:449: no guide -> the text tower's tensor itself, no kernel involved
    out = guide.embeds('a photograph of an astronaut')
    assert tuple(out.shape[:2]) == (1, 77) and out.device.type == 'cpu'


def test_base_guide_noise_pred_not_implemented():
    g = GuideBase(_Enc(), unet=None, guidance=7.5, steps=50)
    assert g.batch_size == 1 and tuple(g.uncond_embeds.shape) == (1, 77, 768)
    with pytest.raises(NotImplementedError):
        g.noise_pred(torch.zeros(1, 4, 8, 8), 1)


def test_pipeline_rejects_strength_outside_unit_interval():
    class _Mod(torch.nn.Module):
        in_channels = 4

    pipe = FlexPipeline(_Mod(), _Mod(), None, _Mod(), schedulers.DDIMScheduler())
    g = SimpleGuide(_Enc(), _Mod(), 7.5, 10, torch.zeros(1, 77, 768))
    for bad in (-0.1, 1.5):
        with pytest.raises(ValueError, match='strength'):
            pipe(g, strength=bad)


def test_blend_has_no_cpu_fallback():
    '''A CPU tensor (or a missing device) must fail loudly in the product path.'''
    tw = Tweener()
    with pytest.raises((_native.NativeError, RuntimeError)):
        tw.tween_batch(torch.zeros(1, 77, 768), torch.zeros(1, 257, 768))
