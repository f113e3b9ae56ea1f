'''The slice of diffusers.DiffusionPipeline the reference relies on
(/root/reference/pipeline/flex.py:26,54,77-83,123,186,263).'''
import torch


class DiffusionPipeline:
    def register_modules(self, **kwargs):
        self._names = list(kwargs)
        for k, v in kwargs.items():
            setattr(self, k, v)

    @property
    def device(self):
        for n in getattr(self, '_names', []):
            m = getattr(self, n)
            if isinstance(m, torch.nn.Module):
                for t in list(m.parameters()) + list(m.buffers()):
                    return t.device
        return torch.device('cpu')

    def to(self, device):
        for n in self._names:
            m = getattr(self, n)
            if isinstance(m, torch.nn.Module):
                m.to(device)
        return self

    def progress_bar(self, iterable):
        return iterable

    @staticmethod
    def numpy_to_pil(images):
        from PIL import Image
        if images.ndim == 3:
            images = images[None, ...]
        images = (images * 255).round().astype('uint8')
        return [Image.fromarray(i) for i in images]

    from_pretrained = classmethod(lambda cls, *a, **k: None)
