"""Query-batch sharding of the hot path across the GPUs of one box (one process per GPU).

The reference data-parallelises over triples with a DistributedSampler + DDP (reference ultra/engine.py:48-60):
every rank holds the full graph and its own slice of the batch.  Queries are independent column blocks of the
folded feature axis (feature = b * 64 + c, reference ultra/layer.py:118,306), so message passing needs no
exchange: rank g owns queries [g*B/G, (g+1)*B/G) and runs rspmm on its (N, B/G * d) slab with the full graph
index.  Collectives appear only around it: one gradient all-reduce per fine-tuning step (~0.78 MB of parameters,
engine.py:55-60) and a gather of per-query results at evaluation (engine.py:148-150).
"""
import torch
import torch.distributed as dist


def query_slab(num_query, rank, world_size):
    """[start, stop) of the contiguous block of queries owned by `rank` (sizes differ by at most one)."""
    if not 0 <= rank < world_size:
        raise ValueError("rank %d outside world of size %d" % (rank, world_size))
    base, extra = divmod(num_query, world_size)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_queries(tensor, rank, world_size, dim=0):
    """Slice a per-query tensor (queries along `dim`) to this rank's slab."""
    start, stop = query_slab(tensor.shape[dim], rank, world_size)
    return tensor.narrow(dim, start, stop - start)


def feature_slab(folded, num_query, rank, world_size):
    """Columns of a folded (rows, num_query * d) operand that belong to this rank's queries."""
    if folded.shape[1] % num_query:
        raise ValueError("feature width %d is not a multiple of the %d queries" % (folded.shape[1], num_query))
    width = folded.shape[1] // num_query
    start, stop = query_slab(num_query, rank, world_size)
    return folded[:, start * width:stop * width]


def all_reduce_gradients(parameters, group=None, average=True):
    """One flat all-reduce over every gradient (the ~168k-parameter model fits a single 0.78 MB bucket).
    Parameters without a gradient contribute zeros, like DDP with find_unused_parameters (engine.py:57-58)."""
    parameters = [p for p in parameters if p.requires_grad]
    if not parameters or not (dist.is_available() and dist.is_initialized()):
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in parameters])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    if average:
        flat /= dist.get_world_size(group)
    offset = 0
    for p in parameters:
        count = p.numel()
        p.grad = flat[offset:offset + count].view_as(p).clone()
        offset += count


def gather_queries(local, num_query, group=None, dim=0):
    """All-gather per-query results (ranks hold slabs of possibly different sizes) back into batch order."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world_size = dist.get_world_size(group)
    sizes = [query_slab(num_query, r, world_size) for r in range(world_size)]
    widest = max(stop - start for start, stop in sizes)
    moved = local.movedim(dim, 0).contiguous()
    padded = torch.zeros((widest,) + tuple(moved.shape[1:]), dtype=moved.dtype, device=moved.device)
    padded[:moved.shape[0]] = moved
    pieces = [torch.empty_like(padded) for _ in range(world_size)]
    dist.all_gather(pieces, padded, group=group)
    joined = torch.cat([piece[:stop - start] for piece, (start, stop) in zip(pieces, sizes)])
    return joined.movedim(0, dim)
