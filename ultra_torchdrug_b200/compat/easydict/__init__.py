"""`easydict` stand-in (reference util.py:8): attribute-style dict."""


class EasyDict(dict):
    def __init__(self, d=None, **kwargs):
        super(EasyDict, self).__init__()
        d = dict(d or {}, **kwargs)
        for key, value in d.items():
            self[key] = value

    @classmethod
    def _wrap(cls, value):
        if isinstance(value, dict) and not isinstance(value, EasyDict):
            return cls(value)
        if isinstance(value, (list, tuple)):
            return type(value)(cls._wrap(v) for v in value)
        return value

    def __setitem__(self, key, value):
        super(EasyDict, self).__setitem__(key, self._wrap(value))

    __setattr__ = __setitem__

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError:
            raise AttributeError(key)
