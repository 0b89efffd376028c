"""Shared helpers for the parity tests: seeded random operands and oracle calls."""
import numpy as np
import torch

from oracle import rspmm_oracle

OPS = [(s, m) for s in ("add", "min", "max") for m in ("mul", "add")]


def random_coo(n_out, n_in, n_rel, nnz, seed=0, duplicates=0, weights="unit", skew=False, dtype=np.float32):
    """Un-coalesced COO (3, nnz + duplicates) + values.  `weights`: "unit" | "random"."""
    rng = np.random.default_rng(seed)
    if nnz == 0 or n_out == 0 or n_in == 0 or n_rel == 0:
        return np.zeros((3, 0), dtype=np.int64), np.zeros(0, dtype=dtype)
    if skew:
        p = 1.0 / np.arange(1, n_out + 1)
        row = rng.choice(n_out, size=nnz, p=p / p.sum())
    else:
        row = rng.integers(0, n_out, nnz)
    indices = np.stack([row, rng.integers(0, n_in, nnz), rng.integers(0, n_rel, nnz)]).astype(np.int64)
    if duplicates:
        pick = rng.integers(0, nnz, duplicates)
        indices = np.concatenate([indices, indices[:, pick]], axis=1)
    indices = indices[:, rng.permutation(indices.shape[1])]
    if weights == "unit":
        values = np.ones(indices.shape[1], dtype=dtype)
    else:
        values = rng.uniform(0.25, 2.0, indices.shape[1]).astype(dtype)
    return indices, values


def random_dense(rows, dim, seed, dtype=np.float32, ties=False):
    rng = np.random.default_rng(seed)
    if ties:  # few distinct values => many exact ties under min/max (the common case in NBF states)
        return rng.integers(-2, 3, (rows, dim)).astype(dtype)
    return rng.standard_normal((rows, dim)).astype(dtype)


def to_sparse(indices, values, shape, device):
    return torch.sparse_coo_tensor(torch.from_numpy(indices).to(device), torch.from_numpy(values).to(device),
                                   tuple(shape), check_invariants=False)


def oracle_forward(indices, values, shape, relation, input, sum, mul, dtype=None):
    return rspmm_oracle.rspmm_forward(indices, values, shape, relation, input, sum, mul, dtype=dtype)


def oracle_backward(indices, values, shape, relation, input, output, grad_output, sum, mul, dtype=None):
    return rspmm_oracle.rspmm_backward(indices, values, shape, relation, input, output, grad_output, sum, mul,
                                       dtype=dtype)
