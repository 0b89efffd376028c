"""`torchdrug.utils.comm` stand-in over torch.distributed (reference rel_model.py:13, util.py:106-123)."""
import os

import torch
from torch import distributed as dist


def get_rank():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank()
    return int(os.environ.get("RANK", 0))


def get_world_size():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size()
    return int(os.environ.get("WORLD_SIZE", 1))


def init_process_group(backend, init_method=None, **kwargs):
    dist.init_process_group(backend, init_method, **kwargs)


def synchronize():
    if get_world_size() > 1 and dist.is_initialized():
        dist.barrier()


def reduce(obj, op="sum"):
    if get_world_size() == 1 or not dist.is_initialized():
        return obj
    if isinstance(obj, dict):
        return {k: reduce(v, op) for k, v in obj.items()}
    tensor = obj.clone()
    dist.all_reduce(tensor, op=dist.ReduceOp.SUM if op in ("sum", "mean") else getattr(dist.ReduceOp, op.upper()))
    if op == "mean":
        tensor = tensor / get_world_size()
    return tensor


def cat(obj):
    if get_world_size() == 1 or not dist.is_initialized():
        return obj
    if isinstance(obj, (tuple, list)):
        return type(obj)(cat(o) for o in obj)
    sizes = [torch.zeros(1, dtype=torch.long, device=obj.device) for _ in range(get_world_size())]
    dist.all_gather(sizes, torch.tensor([len(obj)], device=obj.device))
    longest = int(max(s.item() for s in sizes))
    padded = torch.zeros(longest, *obj.shape[1:], dtype=obj.dtype, device=obj.device)
    padded[:len(obj)] = obj
    parts = [torch.zeros_like(padded) for _ in range(get_world_size())]
    dist.all_gather(parts, padded)
    return torch.cat([p[:int(s.item())] for p, s in zip(parts, sizes)])
