'''Profiling target for ncu: one eager `CLIPEncoder.image` + `.prompt` at ViT-L/14 size (every kernel named).
  ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file ... python profiles/prof_towers.py'''
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from PIL import Image
from flexdiffuse_b200 import factory
from flexdiffuse_b200.encode.clip import CLIPEncoder
dev = torch.device('cuda:0')
clip = factory.build_clip(dev)
enc = CLIPEncoder(clip, factory.FakeTokenizer(), cuda_graph=False)
img = Image.fromarray((np.random.RandomState(0).rand(512, 512, 3) * 255).astype('uint8'))
with torch.no_grad():
    for _ in range(2):
        enc.image(img); enc.prompt('a photograph of an astronaut riding a horse')
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    enc.image(img)
    enc.prompt('a photograph of an astronaut riding a horse')
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
