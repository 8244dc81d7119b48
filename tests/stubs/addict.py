"""Minimal stand-in for `addict.Dict` (test infrastructure, see README.md): a dict with attribute access whose nested dicts are
converted recursively and whose missing keys create empty children (subclasses may override __missing__, as the reference's
ForceKeyErrorDict does, utils/io_util.py:194-196)."""
import copy


class Dict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__()
        for a in args:
            if not a:
                continue
            for k, v in (a.items() if isinstance(a, dict) else a):
                self[k] = self._hook(v)
        for k, v in kwargs.items():
            self[k] = self._hook(v)

    @classmethod
    def _hook(cls, v):
        if isinstance(v, dict) and not isinstance(v, Dict):
            return cls(v)
        if isinstance(v, (list, tuple)):
            return type(v)(cls._hook(x) for x in v)
        return v

    def __setitem__(self, k, v):
        super().__setitem__(k, self._hook(v))

    def __setattr__(self, k, v):
        self[k] = v

    def __getattr__(self, k):
        if k.startswith('__'):
            raise AttributeError(k)
        return self[k]

    def __missing__(self, k):
        v = type(self)()
        super().__setitem__(k, v)
        return v

    def __delattr__(self, k):
        del self[k]

    def update(self, *args, **kwargs):
        other = dict(*args, **kwargs)
        for k, v in other.items():
            if isinstance(v, dict) and isinstance(self.get(k), dict) and k in self:
                self[k].update(v)
            else:
                self[k] = v

    def setdefault(self, k, default=None):
        if k in self:
            return self[k]
        self[k] = default
        return self[k]

    def to_dict(self):
        out = {}
        for k, v in self.items():
            if isinstance(v, Dict):
                out[k] = v.to_dict()
            elif isinstance(v, (list, tuple)):
                out[k] = type(v)(x.to_dict() if isinstance(x, Dict) else x for x in v)
            else:
                out[k] = v
        return out

    def copy(self):
        return copy.copy(self)

    def __deepcopy__(self, memo):
        new = type(self)()
        for k, v in self.items():
            dict.__setitem__(new, copy.deepcopy(k, memo), copy.deepcopy(v, memo))
        return new

    def __getstate__(self):
        return dict(self)

    def __setstate__(self, state):
        for k, v in state.items():
            dict.__setitem__(self, k, v)
