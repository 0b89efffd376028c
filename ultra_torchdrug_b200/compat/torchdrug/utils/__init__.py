"""`torchdrug.utils` stand-in: `cached` (reference model.py:101, rel_model.py:351) and `comm`."""
import ast
import inspect

import torch
from decorator import decorator

from . import comm  # noqa: F401


def _same(x, y):
    if isinstance(x, torch.Tensor) or isinstance(y, torch.Tensor):
        if not (isinstance(x, torch.Tensor) and isinstance(y, torch.Tensor)):
            return False
        return x.shape == y.shape and x.dtype == y.dtype and bool((x == y).all())
    return x is y or x == y


#: set False to make `cached` a pass-through (benchmarks that replay one batch, SURVEY.md 8d "Trap")
CACHE_ENABLED = True


@decorator
def cached(forward, self, *args, **kwargs):
    """Last-call cache, active only in eval mode; pass-through while training."""
    if self.training or not CACHE_ENABLED:
        return forward(self, *args, **kwargs)
    bound = inspect.signature(forward).bind(self, *args, **kwargs)
    bound.apply_defaults()
    arguments = dict(list(bound.arguments.items())[1:])
    store = getattr(self, "_forward_cache", None)
    if store is not None and store["func"] is forward.__name__ and store["keys"] == list(arguments):
        try:
            hit = all(_same(arguments[k], store["arguments"][k]) for k in arguments)
        except Exception:
            hit = False
        if hit:
            return store["result"]
    result = forward(self, *args, **kwargs)
    object.__setattr__(self, "_forward_cache",
                       {"func": forward.__name__, "keys": list(arguments), "arguments": arguments,
                        "result": result})
    return result


def literal_eval(string):
    try:
        return ast.literal_eval(string)
    except (ValueError, SyntaxError):
        return string


def cuda(obj, *args, **kwargs):
    if hasattr(obj, "cuda"):
        return obj.cuda(*args, **kwargs)
    if isinstance(obj, dict):
        return {k: cuda(v, *args, **kwargs) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(cuda(v, *args, **kwargs) for v in obj)
    return obj
