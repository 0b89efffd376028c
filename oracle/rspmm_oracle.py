"""CPU oracle for `generalized_rspmm` forward/backward.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this package; the product (`ultra_torchdrug_b200`) never does and has no CPU fallback.

What it restates
----------------
The arithmetic of the hot path lives in an un-vendored dependency, `torchdrug>=0.2.1`
(reference `requirements.txt:3`; files `torchdrug/layers/functional/spmm.py` and
`torchdrug/layers/functional/extension/rspmm.{h,cpp,cu}` - not under /root/reference, not installed).
This module therefore restates the *published algorithm* of that operator and anchors it on the
reference's own in-repo statement of the same math:

* message:   reference `ultra/layer.py:52-79` / `:232-266`  (`relation_input[relation] (x) input[node_in]`,
             distmult = mul, transe = add)
* aggregate: reference `ultra/layer.py:81-109` / `:268-296` (`scatter_*(message * edge_weight, node_out)`)
* which part is operator vs post-op: reference `ultra/layer.py:154-178` / `:356-380`
* operand layout: reference `ultra/layer.py:118-127` / `:306-328` (`adjacency.transpose(0, 1)`,
  rows = node_out, cols = node_in, layers = relation; features = batch*dim flattened)

Conventions pinned here (SURVEY.md section 8c "Oracle definition"):
* the sparse operand is *coalesced* first: lexicographic sort by (row, col, layer), duplicates merged by
  summing their values (torch `coalesce()` semantics; the callee coalesces, layer.py:127 passes it raw);
* empty rows yield the reduction identity: 0 (add), lowest finite (max), largest finite (min);
* arg-index = lowest coalesced edge position attaining the extremum, -1 for empty rows;
* max/min backward uses the all-ties rule (`out == y` gate): every edge whose recomputed message equals
  the saved output receives the full gradient.

PARITY PIN: the reference ships no tests and no golden vectors for this path ("parity unpinned" by the
reference itself).  The pin used instead: outputs of the reference's own fallback path
(`message()`+`aggregate()`, reached by `graph.requires_grad`, layer.py:112/299) run in the authoring
container from /root/reference through import shims - committed as `tests/golden/*.npz` together with
the generator `tests/golden/make_golden.py`.  `tests/test_oracle.py` checks this oracle against them.
"""
import numpy as np

SUM_OPS = ("add", "min", "max")
MUL_OPS = ("mul", "add")


def _check_ops(sum, mul):
    if sum not in SUM_OPS:
        raise ValueError("Unknown sum op `%s`" % sum)
    if mul not in MUL_OPS:
        raise ValueError("Unknown mul op `%s`" % mul)


def identity(sum, dtype):
    """Reduction identity per op (torchdrug NaryAdd/NaryMin/NaryMax `zero`)."""
    info = np.finfo(dtype)
    return {"add": 0.0, "max": info.min, "min": info.max}[sum]


def coalesce(indices, values, shape):
    """Sort COO entries by (row, col, layer) and merge duplicates by summing values.

    indices: (3, E) integer array [row=node_out, col=node_in, layer=relation]; values: (E,).
    Returns (indices (3, M) int64, values (M,), first_position (M,) of each merged entry in the input).
    """
    indices = np.asarray(indices, dtype=np.int64).reshape(3, -1)
    values = np.asarray(values)
    n_out, n_in, n_rel = (int(s) for s in shape)
    if indices.shape[1] == 0:
        return indices.copy(), values.copy(), np.zeros(0, dtype=np.int64)
    if indices.min() < 0 or (indices.max(axis=1) >= np.array([n_out, n_in, n_rel])).any():
        raise ValueError("sparse index out of range for shape %s" % (shape,))
    key = (indices[0] * n_in + indices[1]) * n_rel + indices[2]
    order = np.argsort(key, kind="stable")
    key = key[order]
    first = np.ones(len(key), dtype=bool)
    first[1:] = key[1:] != key[:-1]
    starts = np.flatnonzero(first)
    # duplicates are summed sequentially in stable (original) order, in fp64, then rounded once -
    # np.add.at is unbuffered and in-order (np.add.reduceat may re-associate), the CUDA index build does the same
    merged = np.zeros(len(starts), dtype=np.float64)
    np.add.at(merged, np.cumsum(first) - 1, values[order].astype(np.float64))
    merged = merged.astype(values.dtype)
    return indices[:, order[starts]], merged, order[starts]


def _combine(mul, rel_rows, in_rows):
    return rel_rows * in_rows if mul == "mul" else rel_rows + in_rows


def _row_segments(rows, n_out):
    """CSR pointers for sorted `rows`."""
    ptr = np.zeros(n_out + 1, dtype=np.int64)
    np.add.at(ptr, rows + 1, 1)
    return np.cumsum(ptr)


def _column_blocks(num_edge, dim, budget_bytes=256 << 20):
    step = max(1, int(budget_bytes // max(1, num_edge * 8)))
    for start in range(0, dim, step):
        yield start, min(dim, start + step)


def rspmm_forward(indices, values, shape, relation, input, sum="add", mul="mul", dtype=None):
    """out[i, :] = (+)_{(i,j,k)} w_ijk * (relation[k, :] (x) input[j, :]) on the coalesced operand.

    `dtype=None` computes in the dtype of `input` (bit-faithful messages for max/min);
    `dtype=np.float64` up-casts first (the "true value" used for tolerance checks of sums).
    Returns (out (n_out, D), argidx (n_out, D) int64 coalesced edge position or -1).
    """
    _check_ops(sum, mul)
    relation = np.asarray(relation)
    input = np.asarray(input)
    n_out, n_in, n_rel = (int(s) for s in shape)
    if relation.ndim != 2 or input.ndim != 2:
        raise ValueError("`relation` and `input` must be 2-D")
    if relation.shape[0] != n_rel or input.shape[0] != n_in or relation.shape[1] != input.shape[1]:
        raise ValueError("shape mismatch: sparse %s, relation %s, input %s" % (shape, relation.shape, input.shape))
    dtype = np.dtype(dtype or input.dtype)
    index, weight, _ = coalesce(indices, values, shape)
    row, col, layer = index
    weight = weight.astype(dtype)
    dim = input.shape[1]
    out = np.full((n_out, dim), identity(sum, dtype), dtype=dtype)
    arg = np.full((n_out, dim), -1, dtype=np.int64)
    if len(row) == 0 or dim == 0:
        return out, arg
    ptr = _row_segments(row, n_out)
    nonempty = np.flatnonzero(ptr[1:] > ptr[:-1])
    starts = ptr[:-1][nonempty]
    position = np.arange(len(row), dtype=np.int64)[:, None]
    for lo, hi in _column_blocks(len(row), dim):
        message = weight[:, None] * _combine(mul, relation[layer, lo:hi].astype(dtype), input[col, lo:hi].astype(dtype))
        if sum == "add":
            out[nonempty, lo:hi] = np.add.reduceat(message, starts, axis=0)
        else:
            ufunc = np.maximum if sum == "max" else np.minimum
            best = ufunc.reduceat(message, starts, axis=0)
            out[nonempty, lo:hi] = best
            expand = np.repeat(np.arange(len(nonempty)), (ptr[1:] - ptr[:-1])[nonempty])
            hit = message == best[expand]
            candidate = np.where(hit, position, np.iinfo(np.int64).max)
            arg[nonempty, lo:hi] = np.minimum.reduceat(candidate, starts, axis=0)
    return out, arg


def rspmm_backward(indices, values, shape, relation, input, output, grad_output, sum="add", mul="mul", dtype=None):
    """Gradients w.r.t. `relation` and `input` (torchdrug `rspmm_backward_out_*`).

    mul:  d rel[k] += g[i] * w * in[j],  d in[j] += g[i] * w * rel[k]
    add:  d rel[k] += g[i] * w,          d in[j] += g[i] * w
    max/min: each term additionally gated by (output[i] == w * (rel[k] (x) in[j]))  - all-ties rule.
    `output` is only read for max/min.  Returns (grad_relation, grad_input).
    """
    _check_ops(sum, mul)
    relation = np.asarray(relation)
    input = np.asarray(input)
    grad_output = np.asarray(grad_output)
    n_out, n_in, n_rel = (int(s) for s in shape)
    dtype = np.dtype(dtype or input.dtype)
    index, weight, _ = coalesce(indices, values, shape)
    row, col, layer = index
    weight = weight.astype(dtype)
    dim = input.shape[1]
    grad_relation = np.zeros((n_rel, dim), dtype=dtype)
    grad_input = np.zeros((n_in, dim), dtype=dtype)
    if len(row) == 0 or dim == 0:
        return grad_relation, grad_input
    for lo, hi in _column_blocks(len(row), dim, budget_bytes=128 << 20):
        rel_rows = relation[layer, lo:hi].astype(dtype)
        in_rows = input[col, lo:hi].astype(dtype)
        upstream = grad_output[row, lo:hi].astype(dtype) * weight[:, None]
        if sum != "add":
            # the gate compares the saved output with the message recomputed in the *forward* dtype
            fwd = np.dtype(np.asarray(output).dtype)
            message = weight.astype(fwd)[:, None] * _combine(mul, relation[layer, lo:hi].astype(fwd),
                                                            input[col, lo:hi].astype(fwd))
            upstream = np.where(np.asarray(output)[row, lo:hi] == message, upstream, 0)
        if mul == "mul":
            to_relation, to_input = upstream * in_rows, upstream * rel_rows
        else:
            to_relation, to_input = upstream, upstream
        np.add.at(grad_relation[:, lo:hi], layer, to_relation)
        np.add.at(grad_input[:, lo:hi], col, to_input)
    return grad_relation, grad_input


def rspmm_argidx_backward(indices, values, shape, relation, input, argidx, grad_output, mul="mul", dtype=None):
    """Single-winner variant (gradient only to the saved arg-index edge).  NOT the reference rule
    (SURVEY.md hard-part 4); kept to document the difference between the two conventions."""
    relation = np.asarray(relation)
    input = np.asarray(input)
    dtype = np.dtype(dtype or input.dtype)
    n_out, n_in, n_rel = (int(s) for s in shape)
    index, weight, _ = coalesce(indices, values, shape)
    _, col, layer = index
    dim = input.shape[1]
    grad_relation = np.zeros((n_rel, dim), dtype=dtype)
    grad_input = np.zeros((n_in, dim), dtype=dtype)
    rows, feats = np.nonzero(np.asarray(argidx) >= 0)
    edge = np.asarray(argidx)[rows, feats]
    upstream = np.asarray(grad_output)[rows, feats].astype(dtype) * weight[edge].astype(dtype)
    if mul == "mul":
        np.add.at(grad_relation, (layer[edge], feats), upstream * input[col[edge], feats].astype(dtype))
        np.add.at(grad_input, (col[edge], feats), upstream * relation[layer[edge], feats].astype(dtype))
    else:
        np.add.at(grad_relation, (layer[edge], feats), upstream)
        np.add.at(grad_input, (col[edge], feats), upstream)
    return grad_relation, grad_input


# --------------------------------------------------------------------------- torch-facing helpers
def generalized_rspmm_oracle(sparse, relation, input, sum="add", mul="mul"):
    """Same call signature as the operator, evaluated by this oracle with autograd support.
    Tests inject it in place of the CUDA operator to run model-level code on CPU."""
    import torch

    class _OracleFunction(torch.autograd.Function):
        @staticmethod
        def forward(ctx, values, relation, input):
            indices = sparse._indices().cpu().numpy()
            out, _ = rspmm_forward(indices, values.detach().cpu().numpy(), tuple(sparse.shape),
                                   relation.detach().cpu().numpy(), input.detach().cpu().numpy(), sum, mul)
            out = torch.from_numpy(out).to(input.device)
            ctx.save_for_backward(values, relation, input, out)
            return out

        @staticmethod
        def backward(ctx, grad_output):
            values, relation, input, out = ctx.saved_tensors
            indices = sparse._indices().cpu().numpy()
            g_rel, g_in = rspmm_backward(indices, values.detach().cpu().numpy(), tuple(sparse.shape),
                                         relation.detach().cpu().numpy(), input.detach().cpu().numpy(),
                                         out.cpu().numpy(), grad_output.contiguous().cpu().numpy(), sum, mul)
            return None, torch.from_numpy(g_rel).to(relation.device), torch.from_numpy(g_in).to(input.device)

    _check_ops(sum, mul)
    return _OracleFunction.apply(sparse._values(), relation.contiguous(), input.contiguous())


def dense_edge_reference(edge_list, edge_weight, num_node, relation, input, sum="add", mul="mul"):
    """Pure-PyTorch dense-edge evaluation following the reference fallback literally
    (`message` layer.py:52-79, `aggregate` layer.py:81-98, without the boundary self-loop).
    Differentiable through torch autograd for `sum="add"`; used to cross-check `rspmm_backward`."""
    import torch

    node_in, node_out, rel = edge_list.t()
    message = relation[rel] * input[node_in] if mul == "mul" else relation[rel] + input[node_in]
    message = message * edge_weight.unsqueeze(-1)
    index = node_out.unsqueeze(-1).expand_as(message)
    if sum == "add":
        out = torch.zeros(num_node, input.shape[1], dtype=input.dtype).scatter_add(0, index, message)
    else:
        fill = identity(sum, np.dtype(str(input.dtype).replace("torch.", "")))
        out = torch.full((num_node, input.shape[1]), fill, dtype=input.dtype)
        out = out.scatter_reduce(0, index, message, reduce="amax" if sum == "max" else "amin", include_self=True)
    return out
