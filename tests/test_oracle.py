"""CPU tests of the oracle itself (`-m "not gpu"`): the numpy restatement and the C restatement are pinned
against (1) the golden vectors produced by the reference's own message()+aggregate() code
(tests/golden/make_golden.py), (2) hand-computed known answers, (3) a literal PyTorch dense-edge evaluation
with autograd, and against each other."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import cpu_ref, rspmm_oracle
from tests import util

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "layer_fallback_*.npz")))


def test_golden_files_present():
    assert len(GOLDEN) == 8


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[15:-4] for p in GOLDEN])
def test_oracle_matches_reference_fallback(path):
    g = np.load(path)
    mul = "mul" if "distmult" in path else "add"
    shape = tuple(g["shape"])
    out, _ = rspmm_oracle.rspmm_forward(g["indices"], g["values"], shape, g["relation"], g["input"], "add", mul,
                                        dtype=np.float64)
    np.testing.assert_allclose(out, g["out_add"], rtol=1e-5, atol=1e-5)
    g_rel, g_in = rspmm_oracle.rspmm_backward(g["indices"], g["values"], shape, g["relation"], g["input"], None,
                                              g["grad_output"], "add", mul, dtype=np.float64)
    np.testing.assert_allclose(g_in, g["grad_input_add"], rtol=1e-5, atol=1e-5)
    if "grad_relation_add" in g:
        np.testing.assert_allclose(g_rel, g["grad_relation_add"], rtol=1e-5, atol=1e-5)
    else:  # shared (R, d) embedding: the operator gradient summed over the query batch
        dim = g["grad_relation_weight_add"].shape[1]
        folded = g_rel.reshape(g_rel.shape[0], -1, dim).sum(axis=1)
        np.testing.assert_allclose(folded, g["grad_relation_weight_add"], rtol=1e-5, atol=1e-5)
    if "out_max" in g:  # min/max values are bit-exact with the reference's scatter_max / scatter_min
        for sum in ("max", "min"):
            out, arg = rspmm_oracle.rspmm_forward(g["indices"], g["values"], shape, g["relation"], g["input"], sum, mul)
            assert np.array_equal(out, g["out_" + sum])
            assert ((arg >= 0) == (out != rspmm_oracle.identity(sum, np.float32))).all()


@pytest.mark.parametrize("path", GOLDEN[:4], ids=[os.path.basename(p)[15:-4] for p in GOLDEN[:4]])
def test_c_port_matches_reference_fallback(path):
    g = np.load(path)
    mul = "mul" if "distmult" in path else "add"
    csr = cpu_ref.CsrOperand(g["indices"], g["values"], tuple(g["shape"]))
    out = cpu_ref.forward(csr, g["relation"], g["input"], "add", mul)
    np.testing.assert_allclose(out, g["out_add"], rtol=1e-5, atol=1e-5)
    if "out_max" in g:
        for sum in ("max", "min"):
            assert np.array_equal(cpu_ref.forward(csr, g["relation"], g["input"], sum, mul), g["out_" + sum])


def test_known_answer():
    """3 destination rows, 2 relations, D = 2; row 2 is empty, edge (0, 1, 0) appears twice (merged weight 2)."""
    indices = np.array([[0, 0, 1, 0], [1, 1, 0, 2], [0, 0, 1, 1]])
    values = np.ones(4, dtype=np.float32)
    shape = (3, 3, 2)
    relation = np.array([[2.0, -1.0], [0.5, 3.0]], dtype=np.float32)
    input = np.array([[1.0, 2.0], [3.0, -4.0], [-5.0, 6.0]], dtype=np.float32)
    lo, hi = np.finfo(np.float32).min, np.finfo(np.float32).max
    expected = {
        # row 0: 2 * (rel0 (x) in1) and 1 * (rel1 (x) in2); row 1: 1 * (rel1 (x) in0)
        ("add", "mul"): [[2 * 6.0 + -2.5, 2 * 4.0 + 18.0], [0.5, 6.0], [0, 0]],
        ("add", "add"): [[2 * 5.0 + -4.5, 2 * -5.0 + 9.0], [1.5, 5.0], [0, 0]],
        ("max", "mul"): [[12.0, 18.0], [0.5, 6.0], [lo, lo]],
        ("min", "mul"): [[-2.5, 8.0], [0.5, 6.0], [hi, hi]],
        ("max", "add"): [[10.0, 9.0], [1.5, 5.0], [lo, lo]],
        ("min", "add"): [[-4.5, -10.0], [1.5, 5.0], [hi, hi]],
    }
    expected_arg = {("max", "mul"): [[0, 1], [2, 2], [-1, -1]], ("min", "mul"): [[1, 0], [2, 2], [-1, -1]]}
    csr = cpu_ref.CsrOperand(indices, values, shape)
    assert list(csr.row_ptr) == [0, 2, 3, 3] and list(csr.val) == [2.0, 1.0, 1.0]
    for (sum, mul), want in expected.items():
        out, arg = rspmm_oracle.rspmm_forward(indices, values, shape, relation, input, sum, mul)
        assert np.array_equal(out, np.array(want, dtype=np.float32)), (sum, mul)
        assert np.array_equal(cpu_ref.forward(csr, relation, input, sum, mul), out)
        if (sum, mul) in expected_arg:
            assert np.array_equal(arg, np.array(expected_arg[(sum, mul)]))
    # backward, max x mul with the all-ties rule: g = ones
    out, _ = rspmm_oracle.rspmm_forward(indices, values, shape, relation, input, "max", "mul")
    g_rel, g_in = rspmm_oracle.rspmm_backward(indices, values, shape, relation, input, out, np.ones((3, 2), np.float32),
                                              "max", "mul")
    # feature 0: row 0 winner is edge 0 (w=2, rel0, in1), row 1 winner edge 2 (rel1, in0)
    # feature 1: row 0 winner is edge 1 (w=1, rel1, in2), row 1 winner edge 2
    assert np.array_equal(g_rel, np.array([[2 * 3.0, 0.0], [1.0, 6.0 + 2.0]], dtype=np.float32))
    assert np.array_equal(g_in, np.array([[0.5, 3.0], [2 * 2.0, 0.0], [0.0, 3.0]], dtype=np.float32))


def test_all_ties_rule():
    """Two edges of one row produce the same message: both receive the full gradient (reference `out == y` gate)."""
    indices = np.array([[0, 0], [0, 1], [0, 0]])
    values = np.ones(2, dtype=np.float32)
    relation = np.ones((1, 1), dtype=np.float32)
    input = np.zeros((2, 1), dtype=np.float32)
    out, arg = rspmm_oracle.rspmm_forward(indices, values, (1, 2, 1), relation, input, "max", "mul")
    assert out[0, 0] == 0 and arg[0, 0] == 0   # lowest coalesced position wins the arg-index
    g_rel, g_in = rspmm_oracle.rspmm_backward(indices, values, (1, 2, 1), relation, input, out, np.ones((1, 1), np.float32),
                                              "max", "mul")
    assert np.array_equal(g_in, np.ones((2, 1), dtype=np.float32))
    one_rel, one_in = rspmm_oracle.rspmm_argidx_backward(indices, values, (1, 2, 1), relation, input, arg,
                                                         np.ones((1, 1), np.float32), "mul")
    assert np.array_equal(one_in, np.array([[1.0], [0.0]], dtype=np.float32))   # single-winner convention differs
    csr = cpu_ref.CsrOperand(indices, values, (1, 2, 1))
    c_rel, c_in = cpu_ref.backward(csr, relation, input, out, np.ones((1, 1), np.float32), "max", "mul")
    assert np.array_equal(c_in, g_in) and np.array_equal(c_rel, g_rel)


@pytest.mark.parametrize("sum,mul", util.OPS)
def test_numpy_and_c_restatements_agree(sum, mul):
    indices, values = util.random_coo(60, 50, 7, 800, seed=3, duplicates=60, weights="random", skew=True)
    shape = (60, 50, 7)
    relation, input = util.random_dense(7, 40, 1, ties=sum != "add"), util.random_dense(50, 40, 2, ties=sum != "add")
    grad = util.random_dense(60, 40, 3)
    out, _ = rspmm_oracle.rspmm_forward(indices, values, shape, relation, input, sum, mul)
    csr = cpu_ref.CsrOperand(indices, values, shape)
    c_out = cpu_ref.forward(csr, relation, input, sum, mul)
    if sum == "add":
        np.testing.assert_allclose(c_out, out, rtol=1e-5, atol=1e-5)
    else:
        assert np.array_equal(c_out, out)
    g_rel, g_in = rspmm_oracle.rspmm_backward(indices, values, shape, relation, input, c_out, grad, sum, mul,
                                              dtype=np.float64)
    c_rel, c_in = cpu_ref.backward(csr, relation, input, c_out, grad, sum, mul)
    np.testing.assert_allclose(c_rel, g_rel, rtol=1e-4, atol=1e-3)
    np.testing.assert_allclose(c_in, g_in, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("mul", ["mul", "add"])
def test_oracle_backward_matches_autograd_of_dense_edges(mul):
    """sum="add": gradients of the literal gather / scatter_add formulation (reference layer.py:52-98) via autograd."""
    indices, values = util.random_coo(30, 30, 5, 300, seed=5, duplicates=20, weights="random", dtype=np.float64)
    relation = torch.randn(5, 12, dtype=torch.float64, requires_grad=True)
    input = torch.randn(30, 12, dtype=torch.float64, requires_grad=True)
    grad = torch.randn(30, 12, dtype=torch.float64)
    edge_list = torch.from_numpy(indices[[1, 0, 2]].T.copy())
    out = rspmm_oracle.dense_edge_reference(edge_list, torch.from_numpy(values), 30, relation, input, "add", mul)
    out.backward(grad)
    o_out, _ = rspmm_oracle.rspmm_forward(indices, values, (30, 30, 5), relation.detach().numpy(), input.detach().numpy(),
                                          "add", mul)
    np.testing.assert_allclose(o_out, out.detach().numpy(), rtol=1e-12, atol=1e-12)
    g_rel, g_in = rspmm_oracle.rspmm_backward(indices, values, (30, 30, 5), relation.detach().numpy(),
                                              input.detach().numpy(), None, grad.numpy(), "add", mul)
    np.testing.assert_allclose(g_rel, relation.grad.numpy(), rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(g_in, input.grad.numpy(), rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("sum", ["max", "min"])
def test_oracle_extrema_match_dense_edges_without_duplicates(sum):
    indices, values = util.random_coo(30, 30, 5, 200, seed=6, duplicates=0)
    indices = np.unique(indices, axis=1)
    values = np.ones(indices.shape[1], dtype=np.float32)
    relation, input = util.random_dense(5, 9, 1), util.random_dense(30, 9, 2)
    edge_list = torch.from_numpy(indices[[1, 0, 2]].T.copy())
    want = rspmm_oracle.dense_edge_reference(edge_list, torch.from_numpy(values), 30, torch.from_numpy(relation),
                                             torch.from_numpy(input), sum, "mul").numpy()
    got, _ = rspmm_oracle.rspmm_forward(indices, values, (30, 30, 5), relation, input, sum, "mul")
    assert np.array_equal(got, want)


def test_coalesce_semantics_match_torch():
    indices, values = util.random_coo(20, 20, 3, 150, seed=8, duplicates=40, weights="random")
    merged_index, merged_value, _ = rspmm_oracle.coalesce(indices, values, (20, 20, 3))
    sparse = torch.sparse_coo_tensor(torch.from_numpy(indices), torch.from_numpy(values), (20, 20, 3)).coalesce()
    assert np.array_equal(merged_index, sparse.indices().numpy())
    np.testing.assert_allclose(merged_value, sparse.values().numpy(), rtol=1e-6)


def test_invalid_arguments():
    with pytest.raises(ValueError):
        rspmm_oracle.rspmm_forward(np.zeros((3, 0)), np.zeros(0), (1, 1, 1), np.zeros((1, 1)), np.zeros((1, 1)), "mean", "mul")
    with pytest.raises(ValueError):
        rspmm_oracle.rspmm_forward(np.array([[0], [5], [0]]), np.ones(1), (1, 1, 1), np.zeros((1, 1)), np.zeros((1, 1)))
    with pytest.raises(ValueError):
        rspmm_oracle.rspmm_forward(np.zeros((3, 0)), np.zeros(0), (1, 2, 1), np.zeros((1, 1)), np.zeros((1, 1)))


@pytest.mark.parametrize("sum,mul", util.OPS)
def test_c_port_exact_variants_match_numpy_oracle(sum, mul):
    """The double-accumulating / arg-index entry points of the C restatement (what the full-size GPU parity tests
    compare with) against the numpy oracle, on a graph with duplicates, empty rows, random weights and exact ties."""
    indices, values = util.random_coo(70, 50, 6, 900, seed=21, duplicates=60, weights="random", skew=True)
    shape = (70, 50, 6)
    relation, input = util.random_dense(6, 20, 1, ties=True), util.random_dense(50, 20, 2, ties=True)
    grad = util.random_dense(70, 20, 3)
    csr = cpu_ref.CsrOperand(indices, values, shape)
    if sum == "add":
        want, _ = rspmm_oracle.rspmm_forward(indices, values, shape, relation, input, "add", mul, dtype=np.float64)
        np.testing.assert_allclose(cpu_ref.forward_f64(csr, relation, input, mul), want, rtol=1e-12, atol=1e-12)
        scale, _ = rspmm_oracle.rspmm_forward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), "add", mul,
                                              dtype=np.float64)
        np.testing.assert_allclose(cpu_ref.forward_f64(csr, relation, input, mul, absolute=True), scale, rtol=1e-12)
        out = want.astype(np.float32)
    else:
        out, arg = cpu_ref.forward_arg(csr, relation, input, sum, mul)
        want, want_arg = rspmm_oracle.rspmm_forward(indices, values, shape, relation, input, sum, mul)
        assert np.array_equal(out, want) and np.array_equal(arg, want_arg)
    got = cpu_ref.backward_f64(csr, relation, input, out, grad, sum, mul)
    want = rspmm_oracle.rspmm_backward(indices, values, shape, relation, input, out, grad, sum, mul, dtype=np.float64)
    for a, b in zip(got, want):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-10)
    got = cpu_ref.backward_f64(csr, relation, input, out, grad, sum, mul, absolute=True)
    want = rspmm_oracle.rspmm_backward(indices, np.abs(values), shape, np.abs(relation), np.abs(input), None, np.abs(grad),
                                       "add", mul, dtype=np.float64)
    for a, b in zip(got, want):
        np.testing.assert_allclose(a, b, rtol=1e-10, atol=1e-10)
