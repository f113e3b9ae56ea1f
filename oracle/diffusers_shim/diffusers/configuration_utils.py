'''FrozenDict as used at /root/reference/pipeline/flex.py:15,70.'''


class FrozenDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        for k, v in self.items():
            object.__setattr__(self, k, v)

    def __setitem__(self, k, v):
        raise Exception('FrozenDict is immutable')
