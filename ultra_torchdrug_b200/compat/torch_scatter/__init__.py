"""`torch_scatter` stand-in over torch.scatter_reduce / index_add (reference layer.py:8, model.py:8).

Only used by the reference's *fallback* message()+aggregate() path, i.e. as part of the parity
specification, never by the product path.
"""
import torch


def _prepare(src, index, dim, dim_size):
    dim = dim % src.dim()
    if dim_size is None:
        dim_size = int(index.max()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    view = [1] * src.dim()
    view[dim] = -1
    if index.dim() == 1:
        index = index.view(view).expand_as(src)
    return dim, shape, index


def scatter_sum(src, index, dim=-1, out=None, dim_size=None):
    dim, shape, index = _prepare(src, index, dim, dim_size)
    if out is None:
        out = torch.zeros(shape, dtype=src.dtype, device=src.device)
    return out.scatter_add(dim, index, src)


scatter_add = scatter_sum


def scatter_mean(src, index, dim=-1, out=None, dim_size=None):
    dim, shape, index = _prepare(src, index, dim, dim_size)
    total = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(dim, index, src)
    count = torch.zeros(shape, dtype=src.dtype, device=src.device).scatter_add(dim, index, torch.ones_like(src))
    return total / count.clamp(min=1)


def _scatter_extreme(src, index, dim, dim_size, reduce):
    dim, shape, index = _prepare(src, index, dim, dim_size)
    out = torch.zeros(shape, dtype=src.dtype, device=src.device)
    out = out.scatter_reduce(dim, index, src, reduce=reduce, include_self=False)
    # arg index: lowest source position attaining the extremum (sentinel = src.size(dim) for empty rows)
    hit = src == out.gather(dim, index)
    position = torch.arange(src.shape[dim], device=src.device)
    view = [1] * src.dim()
    view[dim] = -1
    position = position.view(view).expand_as(src)
    sentinel = src.shape[dim]
    candidate = torch.where(hit, position, torch.full_like(position, sentinel))
    arg = torch.full(shape, sentinel, dtype=torch.long, device=src.device)
    arg = arg.scatter_reduce(dim, index, candidate, reduce="amin", include_self=True)
    return out, arg


def scatter_max(src, index, dim=-1, out=None, dim_size=None):
    return _scatter_extreme(src, index, dim, dim_size, "amax")


def scatter_min(src, index, dim=-1, out=None, dim_size=None):
    return _scatter_extreme(src, index, dim, dim_size, "amin")


def scatter(src, index, dim=-1, out=None, dim_size=None, reduce="sum"):
    table = {"sum": scatter_sum, "add": scatter_sum, "mean": scatter_mean}
    if reduce in table:
        return table[reduce](src, index, dim, out, dim_size)
    if reduce == "max":
        return scatter_max(src, index, dim, out, dim_size)[0]
    if reduce == "min":
        return scatter_min(src, index, dim, out, dim_size)[0]
    raise ValueError("Unknown reduce `%s`" % reduce)
