"""`torchdrug.layers.functional` stand-in.  `generalized_rspmm` is the drop-in boundary
(SURVEY.md section 8b): it resolves to the B200 operator."""
import torch

from ultra_torchdrug_b200.functional import generalized_rspmm  # noqa: F401  (the product)
from ultra_torchdrug_b200 import functional as _b200_functional

for _name in dir(_b200_functional):
    if _name.startswith("RSPMM"):
        globals()[_name] = getattr(_b200_functional, _name)


def as_mask(indexes, length):
    """Index tensor -> boolean mask of the given length (reference model.py:73)."""
    mask = torch.zeros(length, dtype=torch.bool, device=indexes.device)
    mask[indexes] = True
    return mask


def variadic_sample(input, size, num_sample):
    """`num_sample` draws with replacement from each of the variadic segments of `input` (segment lengths `size`)
    [ext-recall of torchdrug.layers.functional.variadic_sample: uniform `torch.rand`, scaled by the segment length and
    truncated, offset by the segment start].  Reference call sites: ultra/task.py:108,113."""
    rand = torch.rand(len(size), num_sample, device=size.device)
    index = (rand * size.unsqueeze(-1)).long()
    index = index + (size.cumsum(0) - size).unsqueeze(-1)
    return input[index]
