import yaml


class DotDict(dict):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        for k, v in list(self.items()):
            dict.__setitem__(self, k, self._w(v))

    @classmethod
    def _w(cls, v):
        if isinstance(v, dict) and not isinstance(v, DotDict):
            return cls(v)
        if isinstance(v, list):
            return [cls._w(i) for i in v]
        return v

    @classmethod
    def from_file(cls, path):
        with open(path) as f:
            return cls(yaml.safe_load(f))

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = self._w(v)

    def __setitem__(self, k, v):
        dict.__setitem__(self, k, self._w(v))
