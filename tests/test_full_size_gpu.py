"""Size-independent properties at BASELINE.json's full size (configs[1]: FB15k-237 shape, batch 64, D = 4096), where the
CPU oracle would take minutes: checksums against a float64 gather evaluation on the GPU, linearity, the adjoint
(dot-product) identity that ties forward and both backward passes together, permutation invariance of the COO
input, sampled bounds for max, and run-to-run determinism."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2(cuda):
    from ultra_torchdrug_b200 import functional as F, synthetic
    edge_list, n, r = synthetic.named_graph("fb15k237")
    sparse = synthetic.operator_operand(edge_list, n, r, cuda)
    generator = torch.Generator(device=cuda).manual_seed(11)
    d = 64 * 64
    relation = torch.randn(r, d, device=cuda, generator=generator)
    input = torch.randn(n, d, device=cuda, generator=generator)
    grad = torch.randn(n, d, device=cuda, generator=generator)
    index = F.graph_index(sparse)
    return {"F": F, "sparse": sparse, "n": n, "r": r, "d": d, "relation": relation, "input": input, "grad": grad,
            "index": index, "edge_list": edge_list.to(cuda)}


def _column_checksum(c2, relation, input, lo, hi):
    """sum_i out[i, lo:hi] evaluated edge by edge in float64 (duplicates counted separately = merged weights)."""
    node_in, node_out, rel = c2["edge_list"].t()
    return (relation[rel, lo:hi].double() * input[node_in, lo:hi].double()).sum(dim=0)


def test_forward_checksum_and_row_samples(c2):
    out = c2["index"].forward(c2["relation"], c2["input"], "add", "mul")
    assert out.shape == (c2["n"], c2["d"]) and torch.isfinite(out).all()
    for lo in range(0, c2["d"], 1024):
        want = _column_checksum(c2, c2["relation"], c2["input"], lo, lo + 256)
        got = out[:, lo:lo + 256].double().sum(dim=0)
        torch.testing.assert_close(got, want, rtol=1e-6, atol=1e-3)      # sums of ~5e5 terms of size ~1
    # a few complete rows against a float64 evaluation of their own edges
    node_in, node_out, rel = c2["edge_list"].t()
    for row in (0, 1234, c2["n"] - 1):
        mask = node_out == row
        want = (c2["relation"][rel[mask]].double() * c2["input"][node_in[mask]].double()).sum(dim=0)
        torch.testing.assert_close(out[row].double(), want, rtol=1e-5, atol=1e-5)


def test_linearity(c2):
    index, relation, x = c2["index"], c2["relation"], c2["input"]
    y = torch.roll(x, 7, dims=0)
    combined = index.forward(relation, 2.0 * x - 0.5 * y, "add", "mul")
    separate = 2.0 * index.forward(relation, x, "add", "mul") - 0.5 * index.forward(relation, y, "add", "mul")
    torch.testing.assert_close(combined, separate, rtol=1e-4, atol=2e-4)


def test_adjoint_identity_links_forward_and_backward(c2):
    """DistMult x add is bilinear: <out, g> = <input, grad_input> = <relation, grad_relation> (float64 dot products)."""
    index, relation, x, g = c2["index"], c2["relation"], c2["input"], c2["grad"]
    out = index.forward(relation, x, "add", "mul")
    grad_relation, grad_input = index.backward(relation, x, out, g, "add", "mul")
    lhs = (out.double() * g.double()).sum()
    torch.testing.assert_close((x.double() * grad_input.double()).sum(), lhs, rtol=1e-6, atol=1e-2)
    torch.testing.assert_close((relation.double() * grad_relation.double()).sum(), lhs, rtol=1e-6, atol=1e-2)
    # TransE: out = sum w (rel + in)  =>  grad_input[j] = sum_i w g[i],  checksum: sum_j grad_input = sum_e w g[dst_e]
    _, grad_input = index.backward(relation, x, None, g, "add", "add", need_relation=False)
    node_out = c2["edge_list"][:, 1]
    want = g[:, :128].double()[node_out].sum(dim=0)
    torch.testing.assert_close(grad_input[:, :128].double().sum(dim=0), want, rtol=1e-6, atol=1e-3)


def test_permutation_invariance_and_determinism(c2):
    F = c2["F"]
    indices, values = c2["sparse"]._indices(), c2["sparse"]._values()
    perm = torch.randperm(indices.shape[1], device=indices.device, generator=torch.Generator(device=indices.device).manual_seed(3))
    shuffled = torch.sparse_coo_tensor(indices[:, perm], values[perm], c2["sparse"].shape, check_invariants=False)
    with torch.no_grad():
        for sum in ("add", "max"):
            first = F.generalized_rspmm(c2["sparse"], c2["relation"], c2["input"], sum=sum)
            again = F.generalized_rspmm(c2["sparse"], c2["relation"], c2["input"], sum=sum)
            other = F.generalized_rspmm(shuffled, c2["relation"], c2["input"], sum=sum)
            assert torch.equal(first, again), "two runs differ"
            assert torch.equal(first, other), "result depends on the order of the COO entries"


def test_max_bounds_sampled_edges(c2):
    index, relation, x = c2["index"], c2["relation"], c2["input"]
    out, arg = index.forward(relation, x, "max", "mul", return_argidx=True)
    node_in, node_out, rel = c2["edge_list"][::97].t()            # every 97th edge
    message = relation[rel, :512] * x[node_in, :512]
    assert (out[node_out, :512] >= message).all()                  # the maximum dominates every message of its row
    assert (arg >= 0).all() and (arg < index.nnz).all()            # no empty rows on this graph
    back = index.forward(relation, x, "min", "mul")
    assert (back <= out).all()
