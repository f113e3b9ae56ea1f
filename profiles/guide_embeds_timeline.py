'''Host timeline of one `Guide.embeds(prompt, image)` call (ms from entry, no extra synchronisation inside the call):
where the host blocks and what the GPU is doing meanwhile.'''
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from flexdiffuse_b200 import factory, guidance as G
from flexdiffuse_b200.encode import clip as C
dev = torch.device('cuda:0')
clipm = factory.build_clip(dev)
guide = G.Guide(clipm, factory.FakeTokenizer(), device=str(dev))
img = Image.fromarray((np.random.RandomState(0).rand(512, 512, 3) * 255).astype('uint8'))
prompt = 'a photograph of an astronaut riding a horse'
marks = []
def wrap(obj, name, label):
    f = getattr(obj, name)
    def g(*a, **k):
        marks.append((label + ' >', time.perf_counter()))
        r = f(*a, **k)
        marks.append((label + ' <', time.perf_counter()))
        return r
    setattr(obj, name, g)
wrap(guide.encoder, 'prompt', 'encoder.prompt')
wrap(guide.encoder, 'image', 'encoder.image')
wrap(C, 'preprocess', 'preprocess')
wrap(G.Tweener, 'tween_batch', 'tween_batch')
with torch.no_grad():
    for _ in range(5):
        guide.embeds(prompt, img)
    torch.cuda.synchronize()
    tot = []
    for _ in range(10):
        marks.clear()
        t0 = time.perf_counter()
        guide.embeds(prompt, img)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        tot.append((t1 - t0, t2 - t0, list(marks), t0))
    a, b, mk, t0 = sorted(tot, key=lambda r: r[1])[len(tot) // 2]
    print(f'embeds returned after {a * 1e3:.2f} ms, GPU idle after {b * 1e3:.2f} ms')
    for label, t in mk:
        print(f'  {label:22s} {(t - t0) * 1e3:6.2f} ms')
