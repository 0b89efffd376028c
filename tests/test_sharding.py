"""Multi-process host logic on CPU (gloo, world_size 2): query-batch sharding, gradient all-reduce, result gather.
The per-rank operator is evaluated by the oracle here (no GPU); the property under test is that sharding the
query batch over ranks and re-assembling reproduces the full-batch result exactly."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rspmm_oracle
from tests import util
from ultra_torchdrug_b200 import sharding


def test_query_slab_partition():
    for num_query in (1, 7, 64, 65):
        for world in (1, 2, 3, 8):
            slabs = [sharding.query_slab(num_query, r, world) for r in range(world)]
            assert slabs[0][0] == 0 and slabs[-1][1] == num_query
            assert all(a[1] == b[0] for a, b in zip(slabs, slabs[1:]))
            sizes = [b - a for a, b in slabs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.query_slab(4, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, results):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        num_query, width = 5, 4       # 5 queries over 2 ranks: slabs of 3 and 2
        indices, values = util.random_coo(30, 30, 4, 200, seed=1, duplicates=10, weights="random")
        shape = (30, 30, 4)
        relation = util.random_dense(4, num_query * width, 2)
        input = util.random_dense(30, num_query * width, 3)
        grad = util.random_dense(30, num_query * width, 4)
        full, _ = rspmm_oracle.rspmm_forward(indices, values, shape, relation, input, "add", "mul")
        full_rel, full_in = rspmm_oracle.rspmm_backward(indices, values, shape, relation, input, None, grad, "add", "mul")

        my = lambda array: np.ascontiguousarray(  # noqa: E731
            sharding.feature_slab(torch.from_numpy(array), num_query, rank, world).numpy())
        out, _ = rspmm_oracle.rspmm_forward(indices, values, shape, my(relation), my(input), "add", "mul")
        g_rel, g_in = rspmm_oracle.rspmm_backward(indices, values, shape, my(relation), my(input), None, my(grad), "add", "mul")

        def assemble(local, rows):
            local = torch.from_numpy(local).view(rows, -1, width)             # (rows, my queries, width)
            return sharding.gather_queries(local, num_query, dim=1).reshape(rows, -1).numpy()

        assert np.array_equal(assemble(out, 30), full)
        assert np.array_equal(assemble(g_in, 30), full_in)
        assert np.array_equal(assemble(g_rel, 4), full_rel)

        # gradient all-reduce: DDP semantics (mean over ranks), unused parameters contribute zeros
        used = torch.nn.Parameter(torch.zeros(3, 2))
        unused = torch.nn.Parameter(torch.zeros(4))
        used.grad = torch.full((3, 2), float(rank + 1))
        if rank == 0:
            unused.grad = torch.ones(4)
        sharding.all_reduce_gradients([used, unused])
        assert torch.equal(used.grad, torch.full((3, 2), 1.5))
        assert torch.equal(unused.grad, torch.full((4,), 0.5))
        results[rank] = True
    finally:
        dist.destroy_process_group()


def test_sharded_operator_and_collectives_gloo():
    world = 2
    port = _free_port()
    manager = mp.get_context("spawn").Manager()
    results = manager.dict()
    mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
    assert all(results.get(r) for r in range(world))
