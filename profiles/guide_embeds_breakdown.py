'''Where one `Guide.embeds(prompt, image)` call spends its time on the B200 box (host + device, ms).'''
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from flexdiffuse_b200 import factory, _native
from flexdiffuse_b200.guidance import Guide, Tweener
from flexdiffuse_b200.encode.clip import preprocess, CLIPEncoder, _CLIP_MEAN, _CLIP_STD
from torchvision.transforms.functional import InterpolationMode, center_crop, normalize, resize
dev = torch.device('cuda:0')
clip = factory.build_clip(dev)
guide = Guide(clip, factory.FakeTokenizer(), device=str(dev))
img = Image.fromarray((np.random.RandomState(0).rand(512, 512, 3) * 255).astype('uint8'))
prompt = 'a photograph of an astronaut riding a horse'
def t(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): r = fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
with torch.no_grad():
    print('embeds total          %.2f ms' % t(lambda: guide.embeds(prompt, img)))
    print('encoder.prompt        %.2f ms' % t(lambda: guide.encoder.prompt(prompt)))
    print('encoder.image         %.2f ms' % t(lambda: guide.encoder.image(img)))
    print('  preprocess (PIL)    %.2f ms' % t(lambda: preprocess(img)))
    x = preprocess(img)
    def host_rest():
        y = center_crop(x, [512, 512]); y = resize(y, [224, 224], interpolation=InterpolationMode.BICUBIC, antialias=True)
        return normalize(y, list(_CLIP_MEAN), list(_CLIP_STD))
    print('  crop/resize/norm    %.2f ms' % t(host_rest))
    y = host_rest()
    print('  H2D                 %.2f ms' % t(lambda: y.to(dev)))
    yd = y.to(dev)
    print('  vision tower graph  %.2f ms' % t(lambda: guide.encoder._run('image', guide.encoder._vision, yd)))
    txt = guide.encoder.prompt(prompt); gi = guide.encoder.image(img)
    tw = Tweener()
    print('tween_batch (check)   %.2f ms' % t(lambda: tw.tween_batch(txt, gi)))
    print('tween_batch (nocheck) %.2f ms' % t(lambda: tw.tween_batch(txt, gi, check=False)))
