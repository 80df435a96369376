"""Minimal attribute-dict with addict.Dict's behaviour as models/CtrlHair/shape_branch/config.py:10-60 uses it:
attribute get/set, missing attributes create nested Dicts, `in`, iteration -- addict itself is not installed here."""


class Dict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__()
        for k, v in dict(*args, **kwargs).items():
            self[k] = self._wrap(v)

    @classmethod
    def _wrap(cls, v):
        if isinstance(v, dict) and not isinstance(v, Dict):
            return cls(v)
        return v

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        if name not in self:
            self[name] = Dict()
        return self[name]

    def __setattr__(self, name, value):
        self[name] = self._wrap(value)

    def __delattr__(self, name):
        del self[name]

    def to_dict(self):
        return {k: (v.to_dict() if isinstance(v, Dict) else v) for k, v in self.items()}
