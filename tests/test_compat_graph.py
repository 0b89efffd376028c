"""The `torchdrug.data.Graph` stand-in (ultra_torchdrug_b200/compat): the operations the reference modules and the
mirror rely on - undirected(add_inverse), match with wildcards, edge_mask, adjacency / degree caches, attribute scopes."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from oracle import cpu_ref, rspmm_oracle
from ultra_torchdrug_b200.compat.torchdrug import data


def _graph(seed=0, num_node=12, num_relation=3, num_edge=40):
    generator = torch.Generator().manual_seed(seed)
    edge_list = torch.stack([torch.randint(num_node, (num_edge,), generator=generator),
                             torch.randint(num_node, (num_edge,), generator=generator),
                             torch.randint(num_relation, (num_edge,), generator=generator)], dim=-1)
    return data.Graph(edge_list, num_node=num_node, num_relation=num_relation)


def test_undirected_layout_and_memoisation():
    graph = _graph()
    both = graph.undirected(add_inverse=True)
    assert both.num_relation == 6 and both.num_edge == 2 * graph.num_edge
    assert torch.equal(both.edge_list[0::2], graph.edge_list)                       # interleaved: edge, inverse, edge, ...
    assert torch.equal(both.edge_list[1::2, 0], graph.edge_list[:, 1])
    assert torch.equal(both.edge_list[1::2, 2], graph.edge_list[:, 2] + 3)
    again = graph.undirected(add_inverse=True)
    assert again is not both and again.edge_list is both.edge_list                  # fresh object, shared tensors
    assert again.adjacency.transpose(0, 1) is both.adjacency.transpose(0, 1)        # one adjacency / transpose per edge set
    with both.graph():
        both.query = torch.ones(1)
    assert not hasattr(again, "query")                                              # per-call attributes do not leak


def test_match_against_brute_force():
    graph = _graph(seed=3)
    pattern = torch.tensor([[1, -1, 0], [-1, 4, -1], [2, 3, 1], [-1, -1, 2], [11, 11, 2]])
    index, count = graph.match(pattern)
    offset = 0
    for row, n in zip(pattern.tolist(), count.tolist()):
        want = [e for e, edge in enumerate(graph.edge_list.tolist())
                if all(p < 0 or p == v for p, v in zip(row, edge))]
        assert sorted(index[offset:offset + n].tolist()) == want
        offset += n
    assert offset == len(index)


def test_edge_mask_degree_and_adjacency():
    graph = _graph(seed=5)
    keep = torch.arange(graph.num_edge) % 3 != 0
    sub = graph.edge_mask(keep)
    assert torch.equal(sub.edge_list, graph.edge_list[keep]) and sub.num_node == graph.num_node
    degree = torch.zeros(graph.num_node).index_add_(0, graph.edge_list[:, 1], torch.ones(graph.num_edge))
    assert torch.equal(graph.degree_out, degree)
    adjacency = graph.adjacency
    assert adjacency.shape == (12, 12, 3) and torch.equal(adjacency._indices(), graph.edge_list.t())
    transposed = adjacency.transpose(0, 1)
    assert torch.equal(transposed._indices()[0], graph.edge_list[:, 1]) and torch.equal(transposed._indices()[1], graph.edge_list[:, 0])


@settings(max_examples=25, deadline=None)
@given(st.integers(1, 20), st.integers(1, 20), st.integers(1, 5), st.integers(0, 120), st.integers(1, 9),
       st.sampled_from(["add", "min", "max"]), st.sampled_from(["mul", "add"]), st.integers(0, 10 ** 6))
def test_oracle_restatements_agree_on_random_graphs(n_out, n_in, n_rel, nnz, dim, sum, mul, seed):
    """hypothesis-generated operands: the numpy restatement and the C restatement of the reference algorithm agree
    (values bit-exact for min/max, sums to fp32 rounding), including empty operands and duplicates."""
    rng = np.random.default_rng(seed)
    indices = np.stack([rng.integers(0, n_out, nnz), rng.integers(0, n_in, nnz), rng.integers(0, n_rel, nnz)]).astype(np.int64)
    values = rng.integers(1, 3, nnz).astype(np.float32)
    relation = rng.integers(-3, 4, (n_rel, dim)).astype(np.float32)
    input = rng.integers(-3, 4, (n_in, dim)).astype(np.float32)
    out, _ = rspmm_oracle.rspmm_forward(indices, values, (n_out, n_in, n_rel), relation, input, sum, mul)
    csr = cpu_ref.CsrOperand(indices, values, (n_out, n_in, n_rel))
    assert np.array_equal(cpu_ref.forward(csr, relation, input, sum, mul), out)     # small integers: sums are exact too
    grad = rng.integers(-2, 3, (n_out, dim)).astype(np.float32)
    g_rel, g_in = rspmm_oracle.rspmm_backward(indices, values, (n_out, n_in, n_rel), relation, input, out, grad, sum, mul)
    c_rel, c_in = cpu_ref.backward(csr, relation, input, out, grad, sum, mul)
    assert np.array_equal(c_rel, g_rel) and np.array_equal(c_in, g_in)
