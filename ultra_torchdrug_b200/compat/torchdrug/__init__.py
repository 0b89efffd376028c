"""Minimal stand-in for the `torchdrug` import surface used by the reference's
`ultra/layer.py`, `ultra/model.py`, `ultra/rel_model.py`.

Written from scratch; only the attributes those files touch exist.
"""
import collections
import collections.abc

import torch
from torch import nn

if not hasattr(collections, "Sequence"):
    collections.Sequence = collections.abc.Sequence

__version__ = "0.2.1+b200shim"


def _module_device(self):
    """torchdrug patches `nn.Module` with a `.device` property (used at reference model.py:91,108)."""
    for tensor in self.parameters():
        return tensor.device
    for tensor in self.buffers():
        return tensor.device
    return torch.device("cpu")


if not isinstance(getattr(nn.Module, "device", None), property):
    nn.Module.device = property(_module_device)

from . import core, data, layers, utils, tasks  # noqa: E402,F401
