"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Needs a B200: `-m gpu`.

Bar (BASELINE.json north_star): min/max values and arg-indices bit-exact; fp32 sums and gradients within
rtol 1e-5 / atol 1e-6 of the oracle.  Sums are compared against the oracle evaluated in float64 (the
"true" value) because the kernel's summation order legitimately differs from any sequential order; the
absolute tolerance is scaled by the row's sum of |terms| as usual for floating-point sums.
"""
import ctypes
import os

import numpy as np
import pytest
import torch

from tests import util

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 1e-6


def _assert_sum_close(actual, expected64, scale64, what):
    """|actual - expected| <= atol + rtol * sum_of_abs_terms  (per element)."""
    err = np.abs(actual.astype(np.float64) - expected64)
    bound = ATOL + RTOL * scale64
    worst = np.unravel_index(np.argmax(err - bound), err.shape) if err.size else None
    assert (err <= bound).all(), "%s: error %g > bound %g at %s" % (what, err[worst], bound[worst], worst)


def _run_case(cuda, n_out, n_in, n_rel, nnz, dim, sum, mul, seed=0, duplicates=0, weights="unit", skew=False,
              ties=False, dtype=np.float32, check_backward=True):
    from ultra_torchdrug_b200 import functional as F
    indices, values = util.random_coo(n_out, n_in, n_rel, nnz, seed, duplicates, weights, skew, dtype)
    shape = (n_out, n_in, n_rel)
    relation = util.random_dense(n_rel, dim, seed + 1, dtype, ties)
    input = util.random_dense(n_in, dim, seed + 2, dtype, ties)
    grad_output = util.random_dense(n_out, dim, seed + 3, dtype)

    index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), shape)
    d_rel, d_in = torch.from_numpy(relation).to(cuda), torch.from_numpy(input).to(cuda)
    out, arg = index.forward(d_rel, d_in, sum, mul, return_argidx=True)
    out_np = out.cpu().numpy()

    exp, exp_arg = util.oracle_forward(indices, values, shape, relation, input, sum, mul)
    if sum == "add":
        exp64, _ = util.oracle_forward(indices, values, shape, relation, input, sum, mul, dtype=np.float64)
        scale, _ = util.oracle_forward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), sum, mul,
                                       dtype=np.float64)
        _assert_sum_close(out_np, exp64, scale, "forward %s/%s" % (sum, mul))
    else:
        assert np.array_equal(out_np, exp), "forward %s/%s values not bit-exact" % (sum, mul)
        assert np.array_equal(arg.cpu().numpy().astype(np.int64), exp_arg), "arg-index mismatch"
    if not check_backward:
        return
    g_rel, g_in = index.backward(d_rel, d_in, out, torch.from_numpy(grad_output).to(cuda), sum, mul)
    e_rel, e_in = util.oracle_backward(indices, values, shape, relation, input, out_np, grad_output, sum, mul,
                                       dtype=np.float64)
    s_rel, s_in = util.oracle_backward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), out_np,
                                       np.abs(grad_output), "add", mul, dtype=np.float64)
    _assert_sum_close(g_rel.cpu().numpy(), e_rel, s_rel, "grad_relation %s/%s" % (sum, mul))
    _assert_sum_close(g_in.cpu().numpy(), e_in, s_in, "grad_input %s/%s" % (sum, mul))


@pytest.mark.parametrize("sum,mul", util.OPS)
@pytest.mark.parametrize("dim", [1, 3, 64, 100, 128, 260])
def test_parity_small(cuda, sum, mul, dim):
    # 37 destination rows over 23 sources: duplicates, empty rows (n_out > distinct rows), random weights
    _run_case(cuda, 37, 23, 5, 150, dim, sum, mul, seed=dim, duplicates=20, weights="random")


@pytest.mark.parametrize("sum,mul", util.OPS)
def test_parity_unit_weights_with_ties(cuda, sum, mul):
    # integer-valued operands: every min/max has many exact ties (all-ties backward rule, lowest arg-index)
    _run_case(cuda, 64, 64, 7, 900, 192, sum, mul, seed=7, duplicates=0, weights="unit", ties=True)


@pytest.mark.parametrize("sum,mul", util.OPS)
def test_parity_split_rows(cuda, sum, mul):
    # hub rows: with chunk = 8 most segments are split into partial rows + combine pass
    from ultra_torchdrug_b200 import _lib
    lib = _lib.lib()
    lib.ultra_rspmm_set_tuning(8, 0, 0)
    try:
        _run_case(cuda, 50, 40, 3, 2000, 132, sum, mul, seed=11, duplicates=50, weights="random", skew=True, ties=True)
    finally:
        lib.ultra_rspmm_set_tuning(256, 0, 0)


@pytest.mark.parametrize("sum,mul", [("add", "mul"), ("max", "mul"), ("min", "add")])
def test_parity_medium_default_chunk(cuda, sum, mul):
    # Zipf destinations: the hub row has > 256 edges => split under the default chunk
    _run_case(cuda, 300, 300, 11, 6000, 256, sum, mul, seed=3, duplicates=10, skew=True)


@pytest.mark.parametrize("sum,mul", util.OPS)
def test_parity_float64(cuda, sum, mul):
    _run_case(cuda, 20, 20, 4, 90, 6, sum, mul, seed=5, duplicates=5, weights="random", dtype=np.float64)
    _run_case(cuda, 20, 20, 4, 90, 7, sum, mul, seed=6, duplicates=5, weights="random", dtype=np.float64)


def test_empty_and_degenerate(cuda):
    for sum, mul in util.OPS:
        _run_case(cuda, 5, 4, 3, 0, 8, sum, mul)            # no edges at all: identity rows, zero grads
        _run_case(cuda, 1, 1, 1, 1, 1, sum, mul)            # single edge, single feature
    from ultra_torchdrug_b200 import functional as F
    empty = F.GraphIndex(torch.zeros(3, 0, dtype=torch.long, device=cuda), torch.zeros(0, device=cuda), (0, 3, 2))
    out = empty.forward(torch.zeros(2, 4, device=cuda), torch.zeros(3, 4, device=cuda))
    assert out.shape == (0, 4)
    out = F.GraphIndex(torch.zeros(3, 0, dtype=torch.long, device=cuda), torch.zeros(0, device=cuda),
                       (3, 3, 2)).forward(torch.zeros(2, 0, device=cuda), torch.zeros(3, 0, device=cuda))
    assert out.shape == (3, 0)


def test_index_matches_coalesce(cuda):
    """The index holds exactly the `coalesce()`d edges (sorted, duplicates merged by sum), numbered by their position
    in coalesce() order (`eid`), in three orders: CSR (dst, rel, src), CSC (src, rel, dst), by relation (rel, dst, src)."""
    from oracle import rspmm_oracle
    from ultra_torchdrug_b200 import functional as F
    indices, values = util.random_coo(40, 30, 6, 500, seed=2, duplicates=80, weights="random")
    shape = (40, 30, 6)
    index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), shape)
    exp_index, exp_w, _ = rspmm_oracle.coalesce(indices, values, shape)
    row, col, rel = exp_index
    m = exp_index.shape[1]
    assert index.nnz == m and index.c.nnz_raw == indices.shape[1] and index.c.unit_weight == 0

    # read the device arrays back through torch views of the index buffer
    base = index.buffer.data_ptr()

    def view(ptr, count, dtype):
        offset = ptr - base
        return index.buffer[offset:offset + count * torch.tensor([], dtype=dtype).element_size()].view(dtype).cpu().numpy()

    def check(order_c, order, seg, first, second, n_seg):
        assert np.array_equal(view(order_c.eid, m, torch.int32), order)
        edge = view(order_c.edge, 2 * m, torch.int32).reshape(m, 2)
        assert np.array_equal(edge[:, 0], first[order]) and np.array_equal(edge[:, 1], second[order])
        np.testing.assert_allclose(view(order_c.w, m, torch.float32), exp_w[order], rtol=1e-6)
        assert np.array_equal(view(order_c.ptr, n_seg + 1, torch.int32), np.searchsorted(seg[order], np.arange(n_seg + 1)))

    csr_order = np.lexsort((col, rel, row))
    check(index.c.csr, csr_order, row, col, rel, shape[0])
    check(index.c.csc, np.lexsort((row, rel, col)), col, row, rel, shape[1])
    check(index.c.rel, csr_order[np.argsort(rel[csr_order], kind="stable")], rel, row, col, shape[2])
    # tasks cover every segment exactly once, longest first; ids are also stored packed; non-unit tasks are flagged
    for order_c, n_seg in ((index.c.csr, shape[0]), (index.c.csc, shape[1]), (index.c.rel, shape[2])):
        task = view(order_c.task, 4 * order_c.n_task, torch.int32).reshape(-1, 4)
        edges = view(order_c.edge, 2 * m, torch.int32).reshape(m, 2).astype(np.int64)
        weights = view(order_c.w, m, torch.float32)
        assert order_c.pack_shift > 0
        packed = view(order_c.packed, m, torch.int32).astype(np.int64) & 0xffffffff
        assert np.array_equal(packed, edges[:, 0] | (edges[:, 1] << order_c.pack_shift))
        for seg, begin, end, encoded in task:
            assert bool(encoded & 0x40000000) == bool((weights[begin:end] != 1).any())
            assert (encoded & 0x3fffffff) - 1 == -1   # no split segments at this size
        seg_ptr = view(order_c.ptr, n_seg + 1, torch.int32)
        length = task[:, 2] - task[:, 1]
        assert (np.diff(length) <= 0).all()
        covered = np.zeros(m, dtype=np.int64)
        for seg, begin, end, slot in task:
            assert seg_ptr[seg] <= begin <= end <= seg_ptr[seg + 1]
            covered[begin:end] += 1
        assert (covered == 1).all() and set(task[:, 0]) == set(range(n_seg))


def test_out_of_range_index_raises(cuda):
    from ultra_torchdrug_b200 import functional as F, _lib
    indices = torch.tensor([[0, 5], [0, 1], [0, 0]], device=cuda)
    with pytest.raises(_lib.RspmmError) as info:
        F.GraphIndex(indices, torch.ones(2, device=cuda), (3, 3, 1))
    assert info.value.status == _lib.ERR_INDEX


def test_deterministic(cuda):
    from ultra_torchdrug_b200 import functional as F
    indices, values = util.random_coo(500, 500, 9, 30000, seed=9, skew=True)
    relation, input = util.random_dense(9, 512, 1), util.random_dense(500, 512, 2)
    grad = util.random_dense(500, 512, 3)
    index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), (500, 500, 9))
    d = [torch.from_numpy(x).to(cuda) for x in (relation, input, grad)]
    first = None
    for _ in range(3):
        out = index.forward(d[0], d[1])
        g_rel, g_in = index.backward(d[0], d[1], out, d[2])
        now = [t.clone() for t in (out, g_rel, g_in)]
        if first is None:
            first = now
        else:
            assert all(torch.equal(a, b) for a, b in zip(first, now)), "results differ between runs"


def test_operator_autograd_and_cache(cuda):
    """The public operator: autograd wiring, index sharing across fresh `transpose` objects, error behaviour."""
    from ultra_torchdrug_b200 import functional as F
    indices, values = util.random_coo(60, 60, 6, 700, seed=4, duplicates=30)
    shape = (60, 60, 6)
    relation = torch.from_numpy(util.random_dense(6, 128, 1)).to(cuda).requires_grad_()
    input = torch.from_numpy(util.random_dense(60, 128, 2)).to(cuda).requires_grad_()
    F.clear_index_cache()
    before = dict(F.cache_stats)
    outs = []
    for _ in range(3):  # a fresh sparse tensor object with equal content per "layer"
        sparse = util.to_sparse(indices, values, shape, cuda)
        outs.append(F.generalized_rspmm(sparse, relation, input, sum="add", mul="mul"))
    assert F.cache_stats["built"] - before["built"] == 1
    assert F.cache_stats["fingerprint_hit"] - before["fingerprint_hit"] == 2
    out = F.generalized_rspmm(sparse, relation, input)  # same object again: attached index
    assert F.cache_stats["attached"] - before["attached"] == 1
    loss = (out * torch.arange(128, device=cuda)).sum()
    loss.backward()
    exp, _ = util.oracle_forward(indices, values, shape, relation.detach().cpu().numpy(), input.detach().cpu().numpy(),
                                 "add", "mul", dtype=np.float64)
    np.testing.assert_allclose(out.detach().cpu().numpy(), exp, rtol=1e-4, atol=1e-4)
    g = np.broadcast_to(np.arange(128, dtype=np.float32), (60, 128))
    e_rel, e_in = util.oracle_backward(indices, values, shape, relation.detach().cpu().numpy(),
                                       input.detach().cpu().numpy(), None, g, "add", "mul", dtype=np.float64)
    np.testing.assert_allclose(relation.grad.cpu().numpy(), e_rel, rtol=1e-4, atol=1e-2)
    np.testing.assert_allclose(input.grad.cpu().numpy(), e_in, rtol=1e-4, atol=1e-2)
    # no_grad: nothing saved
    with torch.no_grad():
        assert not F.generalized_rspmm(sparse, relation, input).requires_grad
    # error behaviour (torchdrug: ValueError for unknown ops, RuntimeError from TORCH_CHECK)
    with pytest.raises(ValueError):
        F.generalized_rspmm(sparse, relation, input, sum="mean")
    with pytest.raises(RuntimeError):
        F.generalized_rspmm(sparse, relation[:, :64], input)
    with pytest.raises(RuntimeError):
        F.generalized_rspmm(sparse, relation.double(), input)
    with pytest.raises(RuntimeError):
        F.generalized_rspmm(sparse, relation.cpu(), input.cpu())
    # non-contiguous operands are accepted (callee makes them contiguous)
    wide = torch.randn(60, 256, device=cuda)
    assert torch.equal(F.generalized_rspmm(sparse, relation.detach(), wide[:, ::2]),
                       F.generalized_rspmm(sparse, relation.detach(), wide[:, ::2].contiguous()))


@pytest.mark.parametrize("sum,mul", util.OPS)
def test_gradcheck_float64(cuda, sum, mul):
    from ultra_torchdrug_b200 import functional as F
    indices, values = util.random_coo(7, 6, 3, 25, seed=8, duplicates=3, weights="random", dtype=np.float64)
    sparse = util.to_sparse(indices, values, (7, 6, 3), cuda)
    relation = torch.randn(3, 5, dtype=torch.float64, device=cuda, requires_grad=True)
    input = torch.randn(6, 5, dtype=torch.float64, device=cuda, requires_grad=True)
    assert torch.autograd.gradcheck(lambda r, x: F.generalized_rspmm(sparse, r, x, sum=sum, mul=mul),
                                    (relation, input), eps=1e-6, atol=1e-6)


@pytest.mark.parametrize("dim", [96, 1160])
def test_host_buffer_ctx(cuda, dim):
    """The torch-free host-buffer entry points (what a non-Python host binds; bench.py's e2e path).
    dim = 1160 spans three 512-column chunks of the upload / compute / download pipeline (ragged last chunk)."""
    from ultra_torchdrug_b200 import _lib
    lib = _lib.lib()
    indices, values = util.random_coo(80, 70, 5, 900, seed=12, duplicates=40, weights="random")
    shape = (80, 70, 5)
    relation, input, grad = util.random_dense(5, dim, 1), util.random_dense(70, dim, 2), util.random_dense(80, dim, 3)
    ctx = ctypes.c_void_p()
    _lib.check(lib.ultra_rspmm_ctx_create(ctypes.byref(ctx), 0), "ctx_create")
    try:
        indices = np.ascontiguousarray(indices)
        _lib.check(lib.ultra_rspmm_ctx_set_graph(ctx, indices.ctypes.data, values.ctypes.data, indices.shape[1],
                                                 80, 70, 5, _lib.F32), "ctx_set_graph")
        assert lib.ultra_rspmm_ctx_nnz(ctx) == len(np.unique(indices, axis=1).T)
        out = np.empty((80, dim), dtype=np.float32)
        g_rel, g_in = np.empty_like(relation), np.empty_like(input)
        for _ in range(2):   # second call reuses the buffer sets
            _lib.check(lib.ultra_rspmm_ctx_forward_backward(
                ctx, relation.ctypes.data, input.ctypes.data, grad.ctypes.data, out.ctypes.data, g_rel.ctypes.data,
                g_in.ctypes.data, dim, 0, 0), "ctx_forward_backward")
        assert lib.ultra_rspmm_ctx_last_kernel_ms(ctx) > 0
        exp, _ = util.oracle_forward(indices, values, shape, relation, input, "add", "mul", dtype=np.float64)
        e_rel, e_in = util.oracle_backward(indices, values, shape, relation, input, None, grad, "add", "mul",
                                           dtype=np.float64)
        np.testing.assert_allclose(out, exp, rtol=1e-4, atol=1e-4)
        np.testing.assert_allclose(g_rel, e_rel, rtol=1e-4, atol=1e-3)
        np.testing.assert_allclose(g_in, e_in, rtol=1e-4, atol=1e-3)
        out2 = np.empty_like(out)
        _lib.check(lib.ultra_rspmm_ctx_forward(ctx, relation.ctypes.data, input.ctypes.data, out2.ctypes.data, dim, 2, 1),
                   "ctx_forward")
        exp2, _ = util.oracle_forward(indices, values, shape, relation, input, "max", "add")
        assert np.array_equal(out2, exp2)
    finally:
        lib.ultra_rspmm_ctx_destroy(ctx)


@pytest.mark.parametrize("dim", [4, 16, 64, 128])
def test_layer_epilogue_matches_torch(cuda, dim):
    """Fused relu(layer_norm(x) * g + b) + residual vs the reference's separate PyTorch ops (layer.py:386-392)."""
    from ultra_torchdrug_b200 import functional as F
    torch.manual_seed(dim)
    x = torch.randn(1000 + dim, 3, dim, device=cuda) * 3 + 1
    residual = torch.randn_like(x)
    weight, bias = torch.randn(dim, device=cuda), torch.randn(dim, device=cuda)
    shift = torch.randn(dim, device=cuda)
    want = torch.relu(torch.nn.functional.layer_norm(x + shift, (dim,), weight, bias, 1e-5)) + residual
    got = F.layer_norm_relu_residual(x, weight, bias, residual, 1e-5, relu=True, linear_bias=shift)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=2e-6)
    want = torch.nn.functional.layer_norm(x, (dim,), None, None, 1e-5)
    torch.testing.assert_close(F.layer_norm_relu_residual(x, None, None, None, 1e-5, relu=False), want, rtol=1e-5, atol=2e-6)
    assert not F.layer_epilogue_supported(x, 48) and not F.layer_epilogue_supported(x.cpu(), dim)
    # fused backward vs autograd through the separate PyTorch ops
    leaves = [t.clone().requires_grad_() for t in (x, shift, weight, bias, residual)]
    ours = [t.clone().requires_grad_() for t in (x, shift, weight, bias, residual)]
    upstream = torch.randn_like(x)
    (torch.relu(torch.nn.functional.layer_norm(leaves[0] + leaves[1], (dim,), leaves[2], leaves[3], 1e-5)) + leaves[4]).backward(upstream)
    F.layer_norm_relu_residual(ours[0], ours[2], ours[3], ours[4], 1e-5, relu=True, linear_bias=ours[1]).backward(upstream)
    for name, a, b in zip(("x", "linear_bias", "weight", "bias", "residual"), ours, leaves):
        scale = float(b.grad.abs().max()) + 1e-6
        torch.testing.assert_close(a.grad, b.grad, rtol=1e-4, atol=1e-5 * scale, msg=lambda m: "%s: %s" % (name, m))
    again = [t.clone().requires_grad_() for t in (x, shift, weight, bias, residual)]
    F.layer_norm_relu_residual(again[0], again[2], again[3], again[4], 1e-5, relu=True, linear_bias=again[1]).backward(upstream)
    assert all(torch.equal(a.grad, b.grad) for a, b in zip(ours, again)), "fused backward is not deterministic"


def test_forward_with_boundary_addend(cuda):
    """`update + boundary` fused into the kernel epilogue (direct rows and split rows) == the separate addition."""
    from ultra_torchdrug_b200 import _lib, functional as F
    lib = _lib.lib()
    lib.ultra_rspmm_set_tuning(16, 0, 0)
    try:
        indices, values = util.random_coo(90, 90, 6, 3000, seed=21, duplicates=30, skew=True)
        sparse = util.to_sparse(indices, values, (90, 90, 6), cuda)
        relation = torch.from_numpy(util.random_dense(6, 200, 1)).to(cuda)
        input = torch.from_numpy(util.random_dense(90, 200, 2)).to(cuda)
        boundary = torch.from_numpy(util.random_dense(90, 200, 3)).to(cuda)
        F.clear_index_cache()
        with torch.no_grad():
            for mul in ("mul", "add"):
                want = F.generalized_rspmm(sparse, relation, input, sum="add", mul=mul) + boundary
                assert torch.equal(F.rspmm_add_boundary(sparse, relation, input, boundary, mul), want)
        assert F.graph_index(sparse).c.csr.n_split > 0
        with pytest.raises(RuntimeError):
            F.graph_index(sparse).forward(relation, input, "max", "mul", addend=boundary)
        # under autograd: same values, gradients of the separate formulation (d boundary = upstream gradient)
        upstream = torch.from_numpy(util.random_dense(90, 200, 4)).to(cuda)
        for mul in ("mul", "add"):
            fused = [t.clone().requires_grad_() for t in (relation, input, boundary)]
            plain = [t.clone().requires_grad_() for t in (relation, input, boundary)]
            out = F.rspmm_add_boundary(sparse, fused[0], fused[1], fused[2], mul)
            want = F.generalized_rspmm(sparse, plain[0], plain[1], sum="add", mul=mul) + plain[2]
            assert torch.equal(out, want)
            out.backward(upstream)
            want.backward(upstream)
            assert all(torch.equal(a.grad, b.grad) for a, b in zip(fused, plain))
        with pytest.raises(RuntimeError):
            F.rspmm_add_boundary(sparse, relation, input, boundary[:, :-1])
        # one-hot boundary (model.py:106-109) in sparse form: values and gradients of the dense formulation
        batch, width = 8, 25
        node_index = torch.randint(90, (batch,), device=cuda)
        query = torch.from_numpy(util.random_dense(batch, width, 5)).to(cuda)
        for mul in ("mul", "add"):
            fused = [t.clone().requires_grad_() for t in (relation, input, query)]
            plain = [t.clone().requires_grad_() for t in (relation, input, query)]
            out = F.rspmm_add_one_hot(sparse, fused[0], fused[1], node_index, fused[2], mul)
            dense = torch.zeros(90, batch, width, device=cuda)
            dense = dense.index_put((node_index, torch.arange(batch, device=cuda)), plain[2], accumulate=True)
            want = F.generalized_rspmm(sparse, plain[0], plain[1], sum="add", mul=mul) + dense.flatten(1)
            assert torch.equal(out, want)
            out.backward(upstream)
            want.backward(upstream)
            assert all(torch.equal(a.grad, b.grad) for a, b in zip(fused, plain))
        with pytest.raises(RuntimeError):
            F.rspmm_add_one_hot(sparse, relation, input, node_index[:-1], query)
    finally:
        lib.ultra_rspmm_set_tuning(256, 0, 0)
        F.clear_index_cache()


@pytest.mark.parametrize("mul", ["mul", "add"])
@pytest.mark.parametrize("block,batch,chunk,nnz", [(64, 3, 256, 4000), (16, 5, 16, 3000), (4, 7, 256, 150), (128, 2, 16, 2500)])
def test_forward_blocked_equals_plain_forward(cuda, mul, block, batch, chunk, nnz):
    """The cat-free layout (operands embedded in (rows, B, stride) buffers, reference layer.py:387) is bit-identical to
    the plain operator + addend on direct rows, split rows (chunk 16) and grouped short rows (nnz 150)."""
    from ultra_torchdrug_b200 import _lib, functional as F
    lib = _lib.lib()
    lib.ultra_rspmm_set_tuning(chunk, 0, 0)
    try:
        n, n_rel, dim = 120, 5, block * batch
        indices, values = util.random_coo(n, n, n_rel, nnz, seed=block + batch, duplicates=10, skew=True)
        index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), (n, n, n_rel))
        relation = torch.from_numpy(util.random_dense(n_rel, dim, 1)).to(cuda)
        input = torch.from_numpy(util.random_dense(n, dim, 2)).to(cuda)
        boundary = torch.from_numpy(util.random_dense(n, dim, 3)).to(cuda)
        want = index.forward(relation, input, "add", mul, addend=boundary)
        # one buffer: input in the left halves, result into the right halves
        buffer = torch.full((n, batch, 2 * block), float("nan"), device=cuda)
        buffer[..., :block] = input.view(n, batch, block)
        index.forward_blocked(relation, buffer, buffer, block, 0, block, mul, addend=boundary)
        assert torch.equal(buffer[..., block:], want.view(n, batch, block)), "blocked result differs"
        assert torch.equal(buffer[..., :block], input.view(n, batch, block)), "input halves were modified"
        # separate buffers, input at an offset, no addend, wider output stride
        source = torch.full((n, batch, block + 8), float("nan"), device=cuda)
        source[..., 8:] = input.view(n, batch, block)
        target = torch.full((n, batch, 3 * block), -7.0, device=cuda)
        index.forward_blocked(relation, source, target, block, 8, 2 * block, mul)
        assert torch.equal(target[..., 2 * block:], index.forward(relation, input, "add", mul).view(n, batch, block))
        assert bool((target[..., :2 * block] == -7.0).all()), "bytes outside the output blocks were written"
        with pytest.raises(RuntimeError):
            index.forward_blocked(relation, buffer[:, :, :-1], buffer, block, 0, block, mul)
    finally:
        lib.ultra_rspmm_set_tuning(256, 0, 0)


def test_layer_epilogue_strided_equals_contiguous(cuda):
    """The strided epilogue (residual read from / result written into halves of wider buffers) == the contiguous one."""
    from ultra_torchdrug_b200 import functional as F
    torch.manual_seed(5)
    dim = 64
    x = torch.randn(700, 3, dim, device=cuda) * 2
    weight, bias, shift = (torch.randn(dim, device=cuda) for _ in range(3))
    current = torch.randn(700, 3, 2 * dim, device=cuda)
    following = torch.full((700, 3, 2 * dim), 9.0, device=cuda)
    want = F.layer_norm_relu_residual(x, weight, bias, current[..., :dim].contiguous(), 1e-5, relu=True, linear_bias=shift)
    F.layer_norm_relu_residual_into(x, following[..., :dim], weight, bias, current[..., :dim], 1e-5, True, shift)
    assert torch.equal(following[..., :dim], want)
    assert bool((following[..., dim:] == 9.0).all())
    want = F.layer_norm_relu_residual(x, None, None, None, 1e-5, relu=False)
    F.layer_norm_relu_residual_into(x, following[..., dim:], relu=False)
    assert torch.equal(following[..., dim:], want)
    with pytest.raises(RuntimeError):
        F.layer_norm_relu_residual_into(x, following[..., ::2])


@pytest.fixture(params=["mma.sync", "tcgen05"])
def linear_kernel(request):
    from ultra_torchdrug_b200 import _lib
    _lib.check(_lib.lib().ultra_layer_linear_set_kernel(1 if request.param == "mma.sync" else 2), "set_kernel")
    yield request.param
    _lib.lib().ultra_layer_linear_set_kernel(0)


@pytest.mark.timeout(120)
@pytest.mark.parametrize("dim,rows", [(64, 1000), (64, 128 * 150 + 37), (32, 777), (64, 5), (64, 128 * 148 * 5)])
def test_fused_linear_epilogue_has_fp32_accuracy(cuda, linear_kernel, dim, rows):
    """Fused Linear + LayerNorm + ReLU + short-cut (3xTF32 split on the tensor cores) vs a float64 evaluation of
    reference layer.py:386-392 + model.py:126-127: its error must be at the level of the fp32 cuBLAS path, far below
    what a plain TF32 product gives."""
    from ultra_torchdrug_b200 import functional as F
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(rows)
    buffer = torch.randn(rows, 1, 2 * dim, device=cuda) * 2 + 0.3
    linear = torch.nn.Linear(2 * dim, dim).to(cuda)
    norm = torch.nn.LayerNorm(dim).to(cuda)
    with torch.no_grad():
        norm.weight.uniform_(0.5, 1.5)
        norm.bias.normal_()

        def evaluate(dtype, tf32=False):
            x = buffer.to(dtype)
            if tf32:                                     # what the reference switches off: 10-bit mantissa operands
                x = (x.view(torch.int32) & ~0x1fff).view(torch.float32)
                w = (linear.weight.view(torch.int32) & ~0x1fff).view(torch.float32)
            else:
                w = linear.weight.to(dtype)
            hidden = torch.nn.functional.linear(x, w, linear.bias.to(dtype))
            hidden = torch.nn.functional.layer_norm(hidden, (dim,), norm.weight.to(dtype), norm.bias.to(dtype), norm.eps)
            return torch.relu(hidden) + buffer.to(dtype)[..., :dim]

        exact = evaluate(torch.float64)
        cublas_error = float((evaluate(torch.float32).double() - exact).abs().max())
        tf32_error = float((evaluate(torch.float32, tf32=True).double() - exact).abs().max())
        following = torch.full((rows, 1, 2 * dim), 3.0, device=cuda)
        F.linear_norm_relu_residual_into(buffer, linear.weight, following[..., :dim], linear.bias, norm.weight, norm.bias,
                                         norm.eps, relu=True, shortcut=True)
        fused_error = float((following[..., :dim].double() - exact).abs().max())
        assert fused_error <= 4 * cublas_error + 1e-6, (fused_error, cublas_error)
        assert fused_error < tf32_error / 50, (fused_error, tf32_error)
        assert bool((following[..., dim:] == 3.0).all()), "columns outside the output half were written"
        torch.testing.assert_close(following[..., :dim], evaluate(torch.float32), rtol=1e-5, atol=1e-5)
        # no affine, no ReLU, no short-cut, contiguous output
        plain = torch.empty(rows, 1, dim, device=cuda)
        F.linear_norm_relu_residual_into(buffer, linear.weight, plain, None, None, None, norm.eps, relu=False, shortcut=False)
        want = torch.nn.functional.layer_norm(torch.nn.functional.linear(buffer, linear.weight), (dim,), None, None, norm.eps)
        torch.testing.assert_close(plain, want, rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError):
        F.linear_norm_relu_residual_into(buffer[..., :-4], linear.weight, plain)


@pytest.mark.parametrize("dim,nodes,batch", [(64, 300, 5), (32, 77, 3), (64, 1, 7)])
def test_score_head_linear_has_fp32_accuracy(cuda, dim, nodes, batch):
    """Fused scoring head (K = d GEMM on 3xTF32 + per-query bias + ReLU + 1-row Linear) vs a float64 evaluation of the
    reference's MLP over cat([hidden, query]) (model.py:141-143, 177-193)."""
    from ultra_torchdrug_b200 import functional as F
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(dim + nodes)
    feature = torch.randn(nodes, batch, 2 * dim, device=cuda)
    query = torch.randn(batch, dim, device=cuda)
    first, second = torch.nn.Linear(2 * dim, 2 * dim).to(cuda), torch.nn.Linear(2 * dim, 1).to(cuda)
    with torch.no_grad():
        def evaluate(dtype):
            joined = torch.cat([feature[..., :dim], query.expand(nodes, -1, -1)], dim=-1).to(dtype)
            hidden = torch.relu(torch.nn.functional.linear(joined, first.weight.to(dtype), first.bias.to(dtype)))
            return torch.nn.functional.linear(hidden, second.weight.to(dtype), second.bias.to(dtype)).squeeze(-1)

        exact = evaluate(torch.float64)
        cublas_error = float((evaluate(torch.float32).double() - exact).abs().max())
        query_bias = torch.nn.functional.linear(query, first.weight[:, dim:], first.bias)
        got = F.score_head_linear(feature, dim, first.weight, query_bias, second.weight, second.bias)
        assert got.shape == (nodes, batch)
        assert float((got.double() - exact).abs().max()) <= 4 * cublas_error + 1e-6
        torch.testing.assert_close(got, evaluate(torch.float32), rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError):
        F.score_head_linear(feature, dim, first.weight, query_bias[:-1], second.weight, second.bias)


@pytest.mark.parametrize("dim", [4, 32, 128])
def test_score_head_matches_torch(cuda, dim):
    """Fused `relu(z + query_bias) . w + b` vs the separate PyTorch ops of the scoring MLP (model.py:177-193)."""
    from ultra_torchdrug_b200 import functional as F
    torch.manual_seed(dim)
    z = torch.randn(333, 5, dim, device=cuda)
    query_bias, weight, bias = torch.randn(5, dim, device=cuda), torch.randn(1, dim, device=cuda), torch.randn(1, device=cuda)
    want = torch.nn.functional.linear(torch.relu(z + query_bias), weight, bias).squeeze(-1)
    torch.testing.assert_close(F.score_head(z, query_bias, weight, bias), want, rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(F.score_head(z, query_bias, weight), want - bias, rtol=1e-5, atol=1e-5)
    with pytest.raises(RuntimeError):
        F.score_head(z, query_bias[:4], weight, bias)


@pytest.mark.parametrize("seed", range(int(os.environ.get("ULTRA_TEST_SEEDS", "24"))))
def test_parity_randomized(cuda, seed):
    """Seeded random shapes: rectangular operands, self-loops, isolated rows, Zipf destinations, odd feature widths,
    tiny / huge chunk sizes, unit and merged weights, fp32 and fp64 (SURVEY.md section 8c test (4))."""
    from ultra_torchdrug_b200 import _lib
    rng = np.random.default_rng(1000 + seed)
    n_out, n_in = int(rng.integers(1, 120)), int(rng.integers(1, 120))
    n_rel = int(rng.choice([1, 2, 4, 9, 40]))
    nnz = int(rng.integers(0, 6 * max(n_out, n_in)))
    dim = int(rng.choice([1, 2, 5, 8, 31, 64, 96, 129, 256, 300]))
    sum, mul = util.OPS[seed % len(util.OPS)]
    chunk = int(rng.choice([4, 32, 256]))
    dtype = np.float64 if seed % 5 == 4 else np.float32
    lib = _lib.lib()
    lib.ultra_rspmm_set_tuning(chunk, int(rng.integers(0, 3)), 0)
    try:
        _run_case(cuda, n_out, n_in, n_rel, nnz, dim, sum, mul, seed=seed, duplicates=int(rng.integers(0, 20)) if nnz else 0,
                  weights="random" if seed % 3 == 0 else "unit", skew=bool(seed % 2), ties=bool(seed % 4 == 1), dtype=dtype)
    finally:
        lib.ultra_rspmm_set_tuning(256, 0, 0)


@pytest.mark.parametrize("mul", ["mul", "add"])
@pytest.mark.parametrize("weights", ["unit", "random"])
def test_fused_pna_equals_four_calls(cuda, mul, weights):
    """One-pass PNA aggregates == the reference's four operator calls: max / min bit for bit, the two sums to fp32
    rounding (the one-pass kernel rounds each message before adding it, the add kernel fuses multiply and add)."""
    from ultra_torchdrug_b200 import _lib, functional as F
    lib = _lib.lib()
    lib.ultra_rspmm_set_tuning(32, 0, 0)      # some split rows
    try:
        indices, values = util.random_coo(70, 60, 5, 2500, seed=31, duplicates=40, weights=weights, skew=True)
        sparse = util.to_sparse(indices, values, (70, 60, 5), cuda)
        relation = torch.from_numpy(util.random_dense(5, 264, 1)).to(cuda)
        input = torch.from_numpy(util.random_dense(60, 264, 2)).to(cuda)
        F.clear_index_cache()
        with torch.no_grad():
            got = F.rspmm_pna(sparse, relation, input, mul)
            want = (F.generalized_rspmm(sparse, relation, input, sum="add", mul=mul),
                    F.generalized_rspmm(sparse, relation ** 2, input ** 2, sum="add", mul=mul),
                    F.generalized_rspmm(sparse, relation, input, sum="max", mul=mul),
                    F.generalized_rspmm(sparse, relation, input, sum="min", mul=mul))
        for name, a, b in zip(("sum", "square sum", "max", "min"), got, want):
            if name in ("max", "min"):
                assert torch.equal(a, b), name
            else:
                torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-4, msg=lambda m: "%s: %s" % (name, m))
        with pytest.raises(RuntimeError):
            F.rspmm_pna(sparse, relation.requires_grad_(), input, mul)
    finally:
        lib.ultra_rspmm_set_tuning(256, 0, 0)
        F.clear_index_cache()


def test_grouped_task_list_covers_every_segment_once(cuda):
    """Low-degree operand: the grouped task list (short rows share a task) covers every segment and every edge exactly
    once, group tasks hold at most 16 consecutive short segments, and results equal the plain path bit for bit."""
    from ultra_torchdrug_b200 import _lib, functional as F
    rng = np.random.default_rng(5)
    n, r, nnz = 600, 7, 2400                                   # average degree 4, some empty rows, one hub
    indices = np.stack([rng.integers(0, n, nnz), rng.integers(0, n, nnz), rng.integers(0, r, nnz)]).astype(np.int64)
    indices[0, :300] = 17                                      # a long row (300 edges > chunk / 4)
    values = np.ones(nnz, dtype=np.float32)
    index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), (n, n, r))
    base = index.buffer.data_ptr()

    def view(ptr, count, dtype):
        offset = ptr - base
        return index.buffer[offset:offset + count * 4].view(dtype).cpu().numpy()

    for order in (index.c.csr, index.c.csc):
        assert order.n_gtask > 0 and order.n_gtask < order.n_task and order.group_edges == 64
        ptr = view(order.ptr, n + 1, torch.int32)
        task = view(order.gtask, 4 * order.n_gtask, torch.int32).reshape(-1, 4)
        seg_hits, edge_hits = np.zeros(n, dtype=int), np.zeros(index.nnz, dtype=int)
        for seg, begin, end, encoded in task:
            rows = ((encoded >> 24) & 15) + 1 if encoded & 0x20000000 else 1
            if encoded & 0x20000000:
                assert begin == ptr[seg] and end == ptr[seg + rows] and rows <= 16
                assert (np.diff(ptr[seg:seg + rows + 1]) <= 64).all()
                seg_hits[seg:seg + rows] += 1
            else:
                assert ptr[seg] <= begin <= end <= ptr[seg + 1]
                seg_hits[seg] += (begin == ptr[seg])
            edge_hits[begin:end] += 1
        assert (seg_hits == 1).all() and (edge_hits == 1).all()
    relation, input = util.random_dense(r, 256, 1), util.random_dense(n, 256, 2)
    grad = util.random_dense(n, 256, 3)
    d = [torch.from_numpy(x).to(cuda) for x in (relation, input, grad)]
    out = index.forward(d[0], d[1], "add", "mul", addend=d[2])
    g_rel, g_in = index.backward(d[0], d[1], out, d[2], "add", "mul")
    exp, _ = util.oracle_forward(indices, values, (n, n, r), relation, input, "add", "mul", dtype=np.float64)
    np.testing.assert_allclose(out.cpu().numpy(), exp + grad, rtol=1e-5, atol=1e-5)
    e_rel, e_in = util.oracle_backward(indices, values, (n, n, r), relation, input, None, grad, "add", "mul", dtype=np.float64)
    np.testing.assert_allclose(g_in.cpu().numpy(), e_in, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(g_rel.cpu().numpy(), e_rel, rtol=1e-5, atol=2e-3)
    mx, arg = index.forward(d[0], d[1], "max", "mul", return_argidx=True)       # arg-index path keeps the plain list
    exp_max, exp_arg = util.oracle_forward(indices, values, (n, n, r), relation, input, "max", "mul")
    assert np.array_equal(mx.cpu().numpy(), exp_max) and np.array_equal(arg.cpu().numpy().astype(np.int64), exp_arg)
    assert np.array_equal(index.forward(d[0], d[1], "max", "mul").cpu().numpy(), exp_max)   # grouped, no arg-index


@pytest.mark.parametrize("sum,mul", [("add", "mul"), ("max", "mul"), ("add", "add")])
def test_parity_moderate_power_law(cuda, sum, mul):
    """4,000 nodes, 45,000 edges, Zipf destinations: hub rows split into partial rows, thousands of short rows walked in
    group tasks, empty rows - all in one index; against the oracle."""
    _run_case(cuda, 4000, 4000, 12, 45000, 260, sum, mul, seed=41, duplicates=500, weights="random", skew=True)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_derived_index_equals_rebuilt_index(cuda, dtype):
    """`GraphIndex.derive(values)` (same structure, new values, no sort) gives bit for bit what an index built from
    (indices, values) gives - forward in all three reductions, both gradients - incl. duplicates and split rows."""
    from ultra_torchdrug_b200 import _lib, functional as F
    lib = _lib.lib()
    lib.ultra_rspmm_set_tuning(16, 0, 0)
    try:
        n, n_rel, dim = 70, 4, 36
        indices, values = util.random_coo(n, n, n_rel, 1500, seed=9, duplicates=60, weights="random", skew=True, dtype=dtype)
        d_indices = torch.from_numpy(indices).to(cuda)
        base = F.GraphIndex(d_indices, torch.ones(indices.shape[1], dtype=torch.from_numpy(values).dtype, device=cuda), (n, n, n_rel))
        assert base.c.unit_weight == 0 or base.c.nnz == base.c.nnz_raw     # duplicates merged to weight 2 make it non-unit
        masked = values.copy()
        masked[::7] = 0
        for new_values in (values, masked, np.ones_like(values)):
            d_values = torch.from_numpy(new_values).to(cuda)
            derived, rebuilt = base.derive(d_values), F.GraphIndex(d_indices, d_values, (n, n, n_rel))
            relation = torch.from_numpy(util.random_dense(n_rel, dim, 1, dtype)).to(cuda)
            input = torch.from_numpy(util.random_dense(n, dim, 2, dtype)).to(cuda)
            grad = torch.from_numpy(util.random_dense(n, dim, 3, dtype)).to(cuda)
            for sum in ("add", "max", "min"):
                got, got_arg = derived.forward(relation, input, sum, "mul", return_argidx=True)
                want, want_arg = rebuilt.forward(relation, input, sum, "mul", return_argidx=True)
                assert torch.equal(got, want), sum
                assert sum == "add" or torch.equal(got_arg, want_arg), sum
                for a, b in zip(derived.backward(relation, input, got, grad, sum, "mul"),
                                rebuilt.backward(relation, input, want, grad, sum, "mul")):
                    assert torch.equal(a, b), sum
        with pytest.raises(RuntimeError):
            base.derive(torch.ones(3, device=cuda))
    finally:
        lib.ultra_rspmm_set_tuning(256, 0, 0)


# ---- rows-in-shared-memory kernel (csrc/rspmm_staged.cu) --------------------------------------------------------------
@pytest.fixture
def staged_always():
    from ultra_torchdrug_b200 import _lib
    _lib.check(_lib.lib().ultra_rspmm_set_staged(2), "ultra_rspmm_set_staged")
    yield _lib
    _lib.check(_lib.lib().ultra_rspmm_set_staged(1), "ultra_rspmm_set_staged")


@pytest.mark.parametrize("mul", ["mul", "add"])
@pytest.mark.parametrize("dim", [4, 60, 64, 200, 256, 1028])
@pytest.mark.parametrize("weights,chunk", [("unit", 0), ("random", 0), ("random", 16)])
def test_staged_kernel_parity(cuda, staged_always, mul, dim, weights, chunk):
    """Few-row operands through the rows-in-shared-memory kernel (forced on): ragged widths, merged duplicates with
    non-unit weights, split rows (chunk 16), empty rows; forward and grad_input are served by it, grad_relation (two
    gathered operands) by the generic kernel."""
    lib = staged_always.lib()
    if chunk:
        lib.ultra_rspmm_set_tuning(chunk, 0, 0)
    try:
        from ultra_torchdrug_b200 import functional as F
        n, n_rel, nnz = 90, 4, 5000
        indices, values = util.random_coo(n, n - 7, n_rel, nnz, seed=dim, duplicates=200, weights=weights, skew=True)
        shape = (n, n - 7, n_rel)
        relation, input = util.random_dense(n_rel, dim, 1), util.random_dense(n - 7, dim, 2)
        grad = util.random_dense(n, dim, 3)
        index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), shape)
        d_rel, d_in, d_grad = (torch.from_numpy(x).to(cuda) for x in (relation, input, grad))
        out = index.forward(d_rel, d_in, "add", mul)
        assert staged_always.pass_info(staged_always.PASS_FORWARD)["kernel_name"] == "rows_in_smem"
        g_rel, g_in = index.backward(d_rel, d_in, out, d_grad, "add", mul)
        assert staged_always.pass_info(staged_always.PASS_GRAD_INPUT)["kernel_name"] == "rows_in_smem"
        exp, _ = util.oracle_forward(indices, values, shape, relation, input, "add", mul, dtype=np.float64)
        scale, _ = util.oracle_forward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), "add", mul, dtype=np.float64)
        _assert_sum_close(out.cpu().numpy(), exp, scale, "staged forward")
        e_rel, e_in = util.oracle_backward(indices, values, shape, relation, input, None, grad, "add", mul, dtype=np.float64)
        s_rel, s_in = util.oracle_backward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), None, np.abs(grad),
                                           "add", mul, dtype=np.float64)
        _assert_sum_close(g_rel.cpu().numpy(), e_rel, s_rel, "grad_relation")
        _assert_sum_close(g_in.cpu().numpy(), e_in, s_in, "staged grad_input")
        assert torch.equal(out, index.forward(d_rel, d_in, "add", mul)), "two runs differ"
        # + boundary in the epilogue
        with_addend = index.forward(d_rel, d_in, "add", mul, addend=d_grad)
        _assert_sum_close(with_addend.cpu().numpy(), exp + grad, scale + np.abs(grad), "staged forward + addend")
    finally:
        if chunk:
            lib.ultra_rspmm_set_tuning(256, 0, 0)


def test_staged_kernel_blocked_layout(cuda, staged_always):
    """The cat-free layer buffers (block 64, stride 128) through the rows-in-shared-memory kernel equal the plain call."""
    from ultra_torchdrug_b200 import functional as F
    n, n_rel, batch, width = 70, 4, 5, 64
    indices, values = util.random_coo(n, n, n_rel, 4000, seed=5, duplicates=50)
    index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), (n, n, n_rel))
    generator = torch.Generator(device=cuda).manual_seed(3)
    buffer = torch.randn(n, batch, 2 * width, device=cuda, generator=generator)
    relation = torch.randn(n_rel, batch * width, device=cuda, generator=generator)
    addend = torch.randn(n, batch * width, device=cuda, generator=generator)
    plain = index.forward(relation, buffer[..., :width].reshape(n, -1).contiguous(), "add", "mul", addend=addend)
    left = buffer[..., :width].clone()
    index.forward_blocked(relation, buffer, buffer, width, 0, width, "mul", addend=addend)
    assert staged_always.pass_info(staged_always.PASS_FORWARD)["kernel_name"] == "rows_in_smem"
    assert torch.equal(buffer[..., width:].reshape(n, -1), plain)
    assert torch.equal(buffer[..., :width], left), "the input halves were modified"


def test_raw_calls_check_their_operands(cuda):
    """The raw `GraphIndex` methods (used by the buffered inference loop without `generalized_rspmm`'s checks) reject
    operands the kernels would read out of bounds or misinterpret - as the reference's TORCH_CHECKs do."""
    from ultra_torchdrug_b200 import functional as F
    indices, values = util.random_coo(30, 30, 6, 300, seed=1)
    index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), (30, 30, 6))
    relation, input = torch.randn(6, 128, device=cuda), torch.randn(30, 128, device=cuda)
    grad = torch.randn(30, 128, device=cuda)
    index.forward(relation, input)
    for bad_relation in (relation[:5], relation.double(), relation[:, :64], relation.cpu(), relation.t().contiguous().t()):
        with pytest.raises(RuntimeError):
            index.forward(bad_relation, input)
        with pytest.raises(RuntimeError):
            index.forward_pna(bad_relation, input)
        with pytest.raises(RuntimeError):
            index.backward(bad_relation, input, None, grad)
    with pytest.raises(RuntimeError):
        index.forward(relation, input[:29])
    with pytest.raises(RuntimeError):
        index.backward(relation, input, None, grad[:, :64])
    with pytest.raises(RuntimeError):
        index.backward(relation, input, None, grad, "max", "mul")       # min / max need the saved output
    with pytest.raises(ValueError):
        index.forward(relation, input, "mean", "mul")
    buffer = torch.randn(30, 2, 128, device=cuda)
    index.forward_blocked(relation, buffer, buffer, 64, 0, 64)
    for bad_relation in (relation[:5], relation.double(), relation[:, :64]):
        with pytest.raises(RuntimeError):
            index.forward_blocked(bad_relation, buffer, buffer, 64, 0, 64)
    with pytest.raises(RuntimeError):
        index.forward_blocked(relation, buffer, buffer, 64, 0, 96)       # output block leaves the stride
    with pytest.raises(RuntimeError):
        index.forward_blocked(relation, buffer, buffer, 64, 96, 64)      # input block leaves the stride


def test_graph_index_lookup_refuses_stream_capture(cuda):
    """An operand without an attached index needs a fingerprint read-back: inside CUDA-graph capture that is an error
    with instructions, not a silent synchronisation."""
    from ultra_torchdrug_b200 import functional as F
    indices, values = util.random_coo(20, 20, 3, 100, seed=2)
    sparse = util.to_sparse(indices, values, (20, 20, 3), cuda)
    relation, input = torch.randn(3, 64, device=cuda), torch.randn(20, 64, device=cuda)
    stream = torch.cuda.Stream(device=cuda)
    graph = torch.cuda.CUDAGraph()
    with torch.no_grad():
        with pytest.raises(RuntimeError, match="before capturing"):
            with torch.cuda.stream(stream), torch.cuda.graph(graph, stream=stream):
                F.generalized_rspmm(sparse, relation, input)
    torch.cuda.synchronize()
    with torch.no_grad():
        expected = F.generalized_rspmm(sparse, relation, input)      # attaches the index to `sparse`
        captured = torch.cuda.CUDAGraph()
        with torch.cuda.graph(captured):
            out = F.generalized_rspmm(sparse, relation, input)
        captured.replay()
        torch.cuda.synchronize()
    assert torch.equal(out, expected)


# ---- index extensions: pair lists (graph of relations) and the destination-block table ----------------------------------
@pytest.fixture
def extensions_always():
    from ultra_torchdrug_b200 import _lib
    _lib.check(_lib.lib().ultra_rspmm_set_extensions(2, 2), "ultra_rspmm_set_extensions")
    yield _lib
    _lib.check(_lib.lib().ultra_rspmm_set_extensions(1, 1), "ultra_rspmm_set_extensions")


def _relation_like_graph(n, density, seed):
    """(3, E) COO of a 4-relation graph over n nodes: every (dst, src) pair present with probability `density`, carrying a
    random non-empty subset of the 4 relations - the structure `construct_relation_graph` produces."""
    rng = np.random.default_rng(seed)
    present = rng.random((n, n)) < density
    masks = rng.integers(1, 16, (n, n)) * present
    masks[n - 1] = 0                                                   # an empty destination row
    rows, cols, rels = [], [], []
    for k in range(4):
        r, c = np.nonzero(masks & (1 << k))
        rows.append(r); cols.append(c); rels.append(np.full(len(r), k))
    indices = np.stack([np.concatenate(rows), np.concatenate(cols), np.concatenate(rels)]).astype(np.int64)
    return indices[:, rng.permutation(indices.shape[1])]


@pytest.mark.parametrize("mul", ["mul", "add"])
@pytest.mark.parametrize("n,density,dim", [(50, 1.0, 256), (97, 0.3, 64), (200, 0.05, 132), (864, 0.02, 60)])
def test_pair_kernel_parity(cuda, extensions_always, mul, n, density, dim):
    """The pair kernel (<= 4 relation types, unit weights, <= 864 nodes: the graph of relations) against the oracle:
    forward and grad_input are served by it, every (node, node) pair read once for up to 4 edges."""
    from oracle import rspmm_oracle
    from ultra_torchdrug_b200 import functional as F
    lib = extensions_always
    indices = _relation_like_graph(n, density, seed=n)
    values = np.ones(indices.shape[1], dtype=np.float32)
    shape = (n, n, 4)
    index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), shape)
    assert index.c.pairs[0].n_pair > 0 and index.c.pairs[1].n_pair > 0
    coalesced, _, _ = rspmm_oracle.coalesce(indices, values, shape)
    assert index.c.pairs[0].n_pair == len(np.unique(coalesced[0] * n + coalesced[1]))
    relation, input, grad = util.random_dense(4, dim, 1), util.random_dense(n, dim, 2), util.random_dense(n, dim, 3)
    d_rel, d_in, d_grad = (torch.from_numpy(x).to(cuda) for x in (relation, input, grad))
    out = index.forward(d_rel, d_in, "add", mul)
    assert lib.pass_info(lib.PASS_FORWARD)["kernel_name"] == "pairs_in_smem"
    g_rel, g_in = index.backward(d_rel, d_in, out, d_grad, "add", mul)
    assert lib.pass_info(lib.PASS_GRAD_INPUT)["kernel_name"] == "pairs_in_smem"
    exp, _ = util.oracle_forward(indices, values, shape, relation, input, "add", mul, dtype=np.float64)
    scale, _ = util.oracle_forward(indices, values, shape, np.abs(relation), np.abs(input), "add", mul, dtype=np.float64)
    _assert_sum_close(out.cpu().numpy(), exp, scale, "pair forward")
    e_rel, e_in = util.oracle_backward(indices, values, shape, relation, input, None, grad, "add", mul, dtype=np.float64)
    s_rel, s_in = util.oracle_backward(indices, values, shape, np.abs(relation), np.abs(input), None, np.abs(grad), "add", mul,
                                       dtype=np.float64)
    _assert_sum_close(g_in.cpu().numpy(), e_in, s_in, "pair grad_input")
    _assert_sum_close(g_rel.cpu().numpy(), e_rel, s_rel, "grad_relation")
    assert torch.equal(out, index.forward(d_rel, d_in, "add", mul)), "two runs differ"
    with_addend = index.forward(d_rel, d_in, "add", mul, addend=d_grad)
    _assert_sum_close(with_addend.cpu().numpy(), exp + grad, scale + np.abs(grad), "pair forward + addend")
    # min / max and non-unit weights keep the edge kernels
    index.forward(d_rel, d_in, "max", mul)
    assert lib.pass_info(lib.PASS_FORWARD)["kernel_name"] == "seg_reduce"


def test_pair_kernel_blocked_layout_and_weighted_fallback(cuda, extensions_always):
    from ultra_torchdrug_b200 import functional as F
    lib = extensions_always
    n, batch, width = 60, 3, 64
    indices = _relation_like_graph(n, 0.5, seed=4)
    values = np.ones(indices.shape[1], dtype=np.float32)
    index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), (n, n, 4))
    generator = torch.Generator(device=cuda).manual_seed(3)
    buffer = torch.randn(n, batch, 2 * width, device=cuda, generator=generator)
    relation = torch.randn(4, batch * width, device=cuda, generator=generator)
    plain = index.forward(relation, buffer[..., :width].reshape(n, -1).contiguous(), "add", "mul")
    index.forward_blocked(relation, buffer, buffer, width, 0, width, "mul")
    assert lib.pass_info(lib.PASS_FORWARD)["kernel_name"] == "pairs_in_smem"
    assert torch.equal(buffer[..., width:].reshape(n, -1), plain)
    # duplicates merge into weight 2: no pair lists, the edge kernels serve the graph
    doubled = np.concatenate([indices, indices[:, :10]], axis=1)
    weighted = F.GraphIndex(torch.from_numpy(doubled).to(cuda), torch.ones(doubled.shape[1], device=cuda), (n, n, 4))
    assert weighted.c.pairs[0].n_pair == 0
    weighted.forward(relation, buffer[..., :width].reshape(n, -1).contiguous(), "add", "mul")
    assert lib.pass_info(lib.PASS_FORWARD)["kernel_name"] in ("rows_in_smem", "seg_reduce")


@pytest.mark.parametrize("mul", ["mul", "add"])
@pytest.mark.parametrize("weights", ["unit", "random"])
@pytest.mark.parametrize("n,n_rel,nnz,dim", [(2000, 7, 30000, 192), (1700, 40, 9000, 64), (800, 3, 5000, 100)])
def test_destination_blocked_grad_relation_parity(cuda, extensions_always, mul, weights, n, n_rel, nnz, dim):
    """grad_relation through the destination-blocked kernel (forced on): grad_output rows of a 768-row block staged in
    shared memory, input rows gathered, one partial row per (relation, block), folded in block order."""
    from ultra_torchdrug_b200 import functional as F
    lib = extensions_always
    indices, values = util.random_coo(n, n - 50, n_rel, nnz, seed=nnz, duplicates=100, weights=weights, skew=True)
    shape = (n, n - 50, n_rel)
    index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), shape)
    assert index.c.n_block == (n + 767) // 768 and index.c.block_rows == 768
    relation, input, grad = util.random_dense(n_rel, dim, 1), util.random_dense(n - 50, dim, 2), util.random_dense(n, dim, 3)
    d_rel, d_in, d_grad = (torch.from_numpy(x).to(cuda) for x in (relation, input, grad))
    g_rel, g_in = index.backward(d_rel, d_in, None, d_grad, "add", mul)
    assert lib.pass_info(lib.PASS_GRAD_RELATION)["kernel_name"] == "dst_blocked"
    e_rel, e_in = util.oracle_backward(indices, values, shape, relation, input, None, grad, "add", mul, dtype=np.float64)
    s_rel, s_in = util.oracle_backward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), None, np.abs(grad),
                                       "add", mul, dtype=np.float64)
    _assert_sum_close(g_rel.cpu().numpy(), e_rel, s_rel, "blocked grad_relation")
    _assert_sum_close(g_in.cpu().numpy(), e_in, s_in, "grad_input")
    again, _ = index.backward(d_rel, d_in, None, d_grad, "add", mul)
    assert torch.equal(g_rel, again), "two runs differ"
    # a derived index (same structure, other weights) inherits the block table
    new_values = torch.from_numpy(values).to(cuda) * 0.5
    derived = index.derive(new_values)
    h_rel, _ = derived.backward(d_rel, d_in, None, d_grad, "add", mul)
    assert lib.pass_info(lib.PASS_GRAD_RELATION)["kernel_name"] == "dst_blocked"
    _assert_sum_close(h_rel.cpu().numpy(), 0.5 * e_rel, 0.5 * s_rel, "blocked grad_relation of a derived index")


@pytest.mark.parametrize("sum", ["max", "min"])
@pytest.mark.parametrize("mul", ["mul", "add"])
@pytest.mark.parametrize("weights,ties", [("unit", True), ("random", False)])
@pytest.mark.parametrize("n,n_rel,nnz,dim", [(2000, 7, 30000, 192), (1700, 40, 9000, 64), (500, 3, 5000, 100)])
def test_destination_blocked_gated_grad_relation_parity(cuda, extensions_always, sum, mul, weights, ties, n, n_rel, nnz, dim):
    """grad_relation of min / max through the destination-blocked gated kernel (forced on): grad_output AND output rows of
    half a block staged in shared memory, input rows gathered, the all-ties gate `output[dst] == w (x (x) relation)` in the
    same expressions as the generic gated kernel - so the sums agree with the oracle, and with integer-valued operands
    (every extremum has exact ties, every sum is exact) bit for bit with the generic kernel."""
    from ultra_torchdrug_b200 import functional as F
    lib = extensions_always
    indices, values = util.random_coo(n, n - 50, n_rel, nnz, seed=nnz + 1, duplicates=100, weights=weights, skew=True)
    shape = (n, n - 50, n_rel)
    index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), shape)
    relation, input = util.random_dense(n_rel, dim, 1, np.float32, ties), util.random_dense(n - 50, dim, 2, np.float32, ties)
    grad = util.random_dense(n, dim, 3, np.float32, ties)
    d_rel, d_in, d_grad = (torch.from_numpy(x).to(cuda) for x in (relation, input, grad))
    out = index.forward(d_rel, d_in, sum, mul)
    g_rel, g_in = index.backward(d_rel, d_in, out, d_grad, sum, mul)
    assert lib.pass_info(lib.PASS_GRAD_RELATION)["kernel_name"] == "dst_blocked_gated"
    out_np = out.cpu().numpy()
    e_rel, e_in = util.oracle_backward(indices, values, shape, relation, input, out_np, grad, sum, mul, dtype=np.float64)
    s_rel, s_in = util.oracle_backward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), out_np, np.abs(grad),
                                       "add", mul, dtype=np.float64)
    _assert_sum_close(g_rel.cpu().numpy(), e_rel, s_rel, "blocked gated grad_relation %s/%s" % (sum, mul))
    _assert_sum_close(g_in.cpu().numpy(), e_in, s_in, "grad_input %s/%s" % (sum, mul))
    again, _ = index.backward(d_rel, d_in, out, d_grad, sum, mul)
    assert torch.equal(g_rel, again), "two runs differ"
    lib.check(lib.lib().ultra_rspmm_set_extensions(2, 0), "ultra_rspmm_set_extensions")       # the generic gated kernel
    generic, _ = index.backward(d_rel, d_in, out, d_grad, sum, mul)
    assert lib.pass_info(lib.PASS_GRAD_RELATION)["kernel_name"] == "seg_gated"
    if ties:
        assert torch.equal(g_rel, generic), "integer-valued operands: the two kernels must agree exactly"
    else:
        _assert_sum_close(g_rel.cpu().numpy(), generic.cpu().numpy().astype(np.float64), s_rel, "blocked vs generic")


@pytest.mark.parametrize("sum", ["max", "min"])
def test_minmax_gradient_all_ties_rule_ext_recall(cuda, sum):
    """Hand-computed known answer for the min / max backward: EVERY edge whose message equals the extremum receives the
    full upstream gradient (torchdrug `NaryMax::backward(out, y) = (out == y)`).  [ext-recall]: the rule follows this
    repo's recollection of the un-vendored torchdrug source; the reference's own fallback path
    cannot pin it (torch_scatter routes the gradient to one arg-index), and tests/test_torchdrug_probe.py compares with
    the real operator whenever torchdrug is importable."""
    from ultra_torchdrug_b200 import functional as F
    # destination 0 receives three edges with equal messages (sources 0, 1, 2 hold the same value), destination 1 one edge
    indices = torch.tensor([[0, 0, 0, 1], [0, 1, 2, 3], [0, 0, 1, 0]], device=cuda)
    values = torch.ones(4, device=cuda)
    index = F.GraphIndex(indices, values, (2, 4, 2))
    relation = torch.tensor([[2.0, 2.0], [2.0, 2.0]], device=cuda)
    input = torch.tensor([[3.0, -1.0], [3.0, -1.0], [3.0, -1.0], [5.0, 4.0]], device=cuda)
    grad = torch.tensor([[1.0, 10.0], [100.0, 1000.0]], device=cuda)
    out, arg = index.forward(relation, input, sum, "mul", return_argidx=True)
    assert out.tolist() == [[6.0, -2.0], [10.0, 8.0]]
    assert arg.tolist() == [[0, 0], [3, 3]]                     # the first tied edge in coalesce() order
    g_rel, g_in = index.backward(relation, input, out, grad, sum, "mul")
    # all three tied edges of destination 0 get g * rel = (2, 20); a single-winner rule would give (2, 20), 0, 0
    assert g_in.tolist() == [[2.0, 20.0], [2.0, 20.0], [2.0, 20.0], [200.0, 2000.0]]
    # relation 0: edges (0<-0), (0<-1), (1<-3); relation 1: edge (0<-2)
    assert g_rel.tolist() == [[3.0 + 3.0 + 500.0, -10.0 - 10.0 + 4000.0], [3.0, -10.0]]


@pytest.mark.parametrize("sub", [2, 4])
@pytest.mark.parametrize("mul", ["mul", "add"])
@pytest.mark.parametrize("weights,chunk,dim", [("unit", 0, 256), ("random", 0, 100), ("random", 16, 132), ("unit", 0, 36)])
def test_subwarp_rows_kernel_parity(cuda, sub, mul, weights, chunk, dim):
    """The sub-warp rows kernel (2 or 4 tasks per warp, 256- / 128-byte slabs; production use: slabs beyond L2), forced
    on a small graph: forward, grad_input and grad_relation (two gathered operands) against the oracle, ragged widths,
    merged duplicates with non-unit weights, split rows (chunk 16) and empty rows included."""
    from ultra_torchdrug_b200 import functional as F, _lib
    lib = _lib.lib()
    _lib.check(lib.ultra_rspmm_set_narrow(1, sub), "ultra_rspmm_set_narrow")
    if chunk:
        lib.ultra_rspmm_set_tuning(chunk, 0, 0)
    try:
        n, n_rel, nnz = 1500, 9, 20000
        indices, values = util.random_coo(n, n - 11, n_rel, nnz, seed=dim + sub, duplicates=300, weights=weights, skew=True)
        shape = (n, n - 11, n_rel)
        relation, input, grad = util.random_dense(n_rel, dim, 1), util.random_dense(n - 11, dim, 2), util.random_dense(n, dim, 3)
        index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), shape)
        d_rel, d_in, d_grad = (torch.from_numpy(x).to(cuda) for x in (relation, input, grad))
        out = index.forward(d_rel, d_in, "add", mul)
        info = _lib.pass_info(_lib.PASS_FORWARD)
        assert info["kernel_name"] == "subwarp_rows" and info["n_slab"] == -(-dim // (128 // sub)), info
        g_rel, g_in = index.backward(d_rel, d_in, out, d_grad, "add", mul)
        assert _lib.pass_info(_lib.PASS_GRAD_INPUT)["kernel_name"] == "subwarp_rows"
        assert _lib.pass_info(_lib.PASS_GRAD_RELATION)["kernel_name"] == "subwarp_rows"
        exp, _ = util.oracle_forward(indices, values, shape, relation, input, "add", mul, dtype=np.float64)
        scale, _ = util.oracle_forward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), "add", mul, dtype=np.float64)
        _assert_sum_close(out.cpu().numpy(), exp, scale, "sub-warp forward")
        e_rel, e_in = util.oracle_backward(indices, values, shape, relation, input, None, grad, "add", mul, dtype=np.float64)
        s_rel, s_in = util.oracle_backward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), None, np.abs(grad),
                                           "add", mul, dtype=np.float64)
        _assert_sum_close(g_rel.cpu().numpy(), e_rel, s_rel, "sub-warp grad_relation")
        _assert_sum_close(g_in.cpu().numpy(), e_in, s_in, "sub-warp grad_input")
        assert torch.equal(out, index.forward(d_rel, d_in, "add", mul)), "two runs differ"
        with_addend = index.forward(d_rel, d_in, "add", mul, addend=d_grad)
        _assert_sum_close(with_addend.cpu().numpy(), exp + grad, scale + np.abs(grad), "sub-warp forward + addend")
        index.forward(d_rel, d_in, "max", mul)                      # min / max keep the generic kernel
        assert _lib.pass_info(_lib.PASS_FORWARD)["kernel_name"] == "seg_reduce"
    finally:
        _lib.check(lib.ultra_rspmm_set_narrow(0, 0), "ultra_rspmm_set_narrow")
        if chunk:
            lib.ultra_rspmm_set_tuning(256, 0, 0)


@pytest.mark.parametrize("rows", [1, 127, 128, 1000, 40000])
def test_combine_linear_forward_and_backward_have_fp32_accuracy(cuda, rows):
    """`combine_linear` (tcgen05 + TMA product, mma.sync weight gradient; no cat, no cuBLAS SIMT SGEMM) against a float64
    evaluation of `cat([input, update]) @ W^T` (reference layer.py:386-388) and its three gradients: errors at the level of
    the fp32 cuBLAS path, far below what TF32 operands give; deterministic."""
    from ultra_torchdrug_b200 import functional as F
    torch.backends.cuda.matmul.allow_tf32 = False
    generator = torch.Generator(device=cuda).manual_seed(rows)
    input = (torch.randn(rows, 64, device=cuda, generator=generator) * 2 + 0.3).requires_grad_()
    update = (torch.randn(rows, 64, device=cuda, generator=generator) * 3 - 0.1).requires_grad_()
    weight = (torch.randn(64, 128, device=cuda, generator=generator) / 8).requires_grad_()
    grad = torch.randn(rows, 64, device=cuda, generator=generator)

    def evaluate(dtype, tf32=False):
        a, b, w = (t.detach().to(dtype).requires_grad_() for t in (input, update, weight))
        joined = torch.cat([a, b], dim=-1)
        if tf32:
            cut = lambda t: (t.detach().view(torch.int32) & ~0x1fff).view(torch.float32)
            out = cut(joined) @ cut(w).t()
            return out, None
        out = joined @ w.t()
        out.backward(grad.to(dtype))
        return out.detach(), (a.grad, b.grad, w.grad)

    exact, exact_grads = evaluate(torch.float64)
    plain, plain_grads = evaluate(torch.float32)
    tf32_out, _ = evaluate(torch.float32, tf32=True)
    out = F.combine_linear(input, update, weight)
    out.backward(grad)
    got_grads = (input.grad, update.grad, weight.grad)
    error = lambda value, reference: float((value.double() - reference).abs().max())
    assert error(out.detach(), exact) <= 4 * error(plain, exact) + 1e-6
    assert error(out.detach(), exact) < max(error(tf32_out, exact) / 20, 1e-6)
    for got, want32, want64, name in zip(got_grads, plain_grads, exact_grads, ("input", "update", "weight")):
        scale = float(want64.abs().max()) + 1e-6
        assert error(got, want64) <= 4 * error(want32, want64) + 1e-6 * scale, (name, error(got, want64), error(want32, want64))
    first = [g.clone() for g in got_grads]
    input.grad = update.grad = weight.grad = None
    F.combine_linear(input, update, weight).backward(grad)
    assert all(torch.equal(a, b) for a, b in zip(first, (input.grad, update.grad, weight.grad))), "two runs differ"
    # (N, B, 64) operands, as the layers pass them
    shaped = F.combine_linear(input.detach().view(rows, 1, 64), update.detach().view(rows, 1, 64), weight.detach())
    assert shaped.shape == (rows, 1, 64) and torch.equal(shaped.view(rows, 64), out.detach())
    with pytest.raises(RuntimeError):
        F.combine_linear(input[:, :32], update[:, :32], weight)


@pytest.mark.parametrize("nodes,batch,relations,mul,shortcut", [(300, 3, 7, "mul", True), (1500, 2, 40, "add", True),
                                                                  (90, 5, 4, "mul", False)])
def test_single_node_layer_equals_three_node_layer(cuda, nodes, batch, relations, mul, shortcut):
    """`nbf_layer` (operator + Linear + LayerNorm/ReLU/short-cut as ONE autograd node whose backward folds the three
    gradients of the layer input into the kernels) against the composition of the three separate nodes
    (`rspmm_add_one_hot`, `combine_linear`, `layer_norm_relu_residual`): same arithmetic up to summation order, forward and
    all seven gradients, bit-reproducible; and `backward(input_addend=...)` against backward + add."""
    from ultra_torchdrug_b200 import functional as F, synthetic
    generator = torch.Generator().manual_seed(nodes)
    edges = torch.stack([torch.randint(nodes, (nodes * 6,), generator=generator), torch.randint(nodes, (nodes * 6,), generator=generator),
                         torch.randint(relations, (nodes * 6,), generator=generator)], dim=1)
    sparse = synthetic.operator_operand(edges, nodes, relations, cuda)
    device_generator = torch.Generator(device=cuda).manual_seed(batch)
    leaf = lambda *shape, scale=1.0: (torch.randn(*shape, device=cuda, generator=device_generator) * scale).requires_grad_()
    x, relation, query = leaf(nodes, batch, 64), leaf(relations, batch * 64), leaf(batch, 64)
    weight, linear_bias, gamma, beta = leaf(64, 128, scale=0.125), leaf(64), leaf(64), leaf(64)
    node_index = torch.randint(nodes, (batch,), device=cuda, generator=device_generator)
    grad = torch.randn(nodes, batch, 64, device=cuda, generator=device_generator)
    leaves = (x, relation, query, weight, linear_bias, gamma, beta)

    def three_nodes():
        update = F.rspmm_add_one_hot(sparse, relation, x.flatten(1), node_index, query, mul).view(nodes, batch, 64)
        return F.layer_norm_relu_residual(F.combine_linear(x, update, weight), gamma, beta, x if shortcut else None, 1e-5,
                                          relu=True, linear_bias=linear_bias)

    def grads(fn):
        for t in leaves:
            t.grad = None
        out = fn()
        out.backward(grad)
        return out.detach(), [t.grad.clone() for t in leaves]

    want, want_grads = grads(three_nodes)
    got, got_grads = grads(lambda: F.nbf_layer(sparse, relation, x, node_index, query, weight, linear_bias, gamma, beta, mul, 1e-5,
                                               True, shortcut))
    # the single node's forward is ONE kernel (Linear + LayerNorm + ReLU + short-cut, also returning the Linear's output); the
    # three nodes run a GEMM and an epilogue kernel whose row statistics are summed in another order
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    for name, a, b in zip(("x", "relation", "query", "weight", "linear_bias", "gamma", "beta"), got_grads, want_grads):
        torch.testing.assert_close(a, b, rtol=2e-5, atol=2e-5 * float(b.abs().max()), msg=lambda m, name=name: name + ": " + m)
    again, again_grads = grads(lambda: F.nbf_layer(sparse, relation, x, node_index, query, weight, linear_bias, gamma, beta, mul,
                                                   1e-5, True, shortcut))
    assert torch.equal(again, got) and all(torch.equal(a, b) for a, b in zip(again_grads, got_grads)), "two runs differ"
    # the operator's backward with an addend
    index = F.graph_index(sparse)
    flat_x, flat_grad = x.detach().flatten(1), grad.flatten(1)
    addend = torch.randn_like(flat_x)
    g_rel, g_in = index.backward(relation.detach(), flat_x, None, flat_grad, "add", mul)
    a_rel, a_in = index.backward(relation.detach(), flat_x, None, flat_grad, "add", mul, input_addend=addend)
    assert torch.equal(a_rel, g_rel)
    torch.testing.assert_close(a_in, g_in + addend, rtol=1e-6, atol=1e-6 * float(g_in.abs().max()))
    with pytest.raises(RuntimeError):
        index.backward(relation.detach(), flat_x, None, flat_grad, "add", mul, need_input=False, input_addend=addend)


@pytest.mark.parametrize("seed", range(int(os.environ.get("ULTRA_TEST_SEEDS", "24"))))
def test_parity_randomized_specialised_kernels(cuda, seed):
    """Seeded random shapes through the kernels that production selects only on particular graphs, all forced on here:
    rows-in-shared-memory / pair kernels (few rows; pairs need <= 4 relation types and unit weights), the
    destination-blocked grad_relation (sum, and its gated min / max form), the sub-warp rows kernel.  Both message functions,
    ragged widths, duplicates, split rows."""
    from ultra_torchdrug_b200 import functional as F, _lib
    rng = np.random.default_rng(5000 + seed)
    family = ("pairs", "staged", "blocked", "subwarp")[seed % 4]
    n = int(rng.integers(2, 300)) if family in ("pairs", "staged") else int(rng.integers(2, 2500))
    n_in = n if family == "pairs" else max(1, n - int(rng.integers(0, min(n, 20))))
    n_rel = int(rng.integers(1, 5)) if family == "pairs" else int(rng.choice([1, 3, 4, 17, 60]))
    nnz = int(rng.integers(0, 12 * n))
    dim = int(rng.choice([4, 8, 36, 64, 100, 128, 192, 260]))
    mul = ("mul", "add")[(seed // 4) % 2]
    weights = "unit" if family == "pairs" or seed % 3 else "random"
    duplicates = 0 if family == "pairs" else int(rng.integers(0, 30)) if nnz else 0
    chunk = int(rng.choice([8, 64, 256]))
    lib = _lib.lib()
    lib.ultra_rspmm_set_tuning(chunk, 0, 0)
    _lib.check(lib.ultra_rspmm_set_extensions(2 if family == "pairs" else 0, 2 if family == "blocked" else 0), "set_extensions")
    _lib.check(lib.ultra_rspmm_set_staged(2 if family in ("pairs", "staged") else 0), "set_staged")
    _lib.check(lib.ultra_rspmm_set_narrow(1 if family == "subwarp" else 0, int(rng.choice([2, 4]))), "set_narrow")
    try:
        indices, values = util.random_coo(n, n_in, n_rel, nnz, seed, duplicates, weights, bool(seed % 2))
        if family == "pairs" and indices.shape[1]:          # pair lists need coalesced unit weights: drop duplicate triples
            indices = np.unique(indices, axis=1)
            values = np.ones(indices.shape[1], dtype=np.float32)
        shape = (n, n_in, n_rel)
        relation, input, grad = util.random_dense(n_rel, dim, seed + 1), util.random_dense(n_in, dim, seed + 2), util.random_dense(n, dim, seed + 3)
        index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), shape)
        d_rel, d_in, d_grad = (torch.from_numpy(x).to(cuda) for x in (relation, input, grad))
        out = index.forward(d_rel, d_in, "add", mul)
        served = _lib.pass_info(_lib.PASS_FORWARD)["kernel_name"] if indices.shape[1] and dim else None
        g_rel, g_in = index.backward(d_rel, d_in, out, d_grad, "add", mul)
        if indices.shape[1]:
            expected = {"pairs": ("pairs_in_smem",), "staged": ("rows_in_smem",), "blocked": ("seg_reduce",), "subwarp": ("subwarp_rows",)}
            if dim % 4 == 0:
                assert served in expected[family], (family, served)
                if family == "blocked":
                    assert _lib.pass_info(_lib.PASS_GRAD_RELATION)["kernel_name"] == "dst_blocked"
        exp, _ = util.oracle_forward(indices, values, shape, relation, input, "add", mul, dtype=np.float64)
        scale, _ = util.oracle_forward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), "add", mul, dtype=np.float64)
        _assert_sum_close(out.cpu().numpy(), exp, scale, "%s forward" % family)
        e_rel, e_in = util.oracle_backward(indices, values, shape, relation, input, None, grad, "add", mul, dtype=np.float64)
        s_rel, s_in = util.oracle_backward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), None, np.abs(grad),
                                           "add", mul, dtype=np.float64)
        _assert_sum_close(g_rel.cpu().numpy(), e_rel, s_rel, "%s grad_relation" % family)
        _assert_sum_close(g_in.cpu().numpy(), e_in, s_in, "%s grad_input" % family)
        if family == "blocked" and indices.shape[1] and dim % 4 == 0:
            # the gated (min / max) form of the destination-blocked pass on the same graph
            extremum = ("max", "min")[(seed // 8) % 2]
            m_out = index.forward(d_rel, d_in, extremum, mul)
            m_rel, m_in = index.backward(d_rel, d_in, m_out, d_grad, extremum, mul)
            assert _lib.pass_info(_lib.PASS_GRAD_RELATION)["kernel_name"] == "dst_blocked_gated"
            m_np = m_out.cpu().numpy()
            e_rel, e_in = util.oracle_backward(indices, values, shape, relation, input, m_np, grad, extremum, mul, dtype=np.float64)
            s_rel, s_in = util.oracle_backward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), m_np, np.abs(grad),
                                               "add", mul, dtype=np.float64)
            _assert_sum_close(m_rel.cpu().numpy(), e_rel, s_rel, "blocked gated grad_relation (%s)" % extremum)
            _assert_sum_close(m_in.cpu().numpy(), e_in, s_in, "gated grad_input (%s)" % extremum)
    finally:
        lib.ultra_rspmm_set_tuning(256, 0, 0)
        _lib.check(lib.ultra_rspmm_set_extensions(1, 1), "set_extensions")
        _lib.check(lib.ultra_rspmm_set_staged(1), "set_staged")
        _lib.check(lib.ultra_rspmm_set_narrow(0, 0), "set_narrow")
