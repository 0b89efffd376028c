"""Layer-level parity: `message_and_aggregate` + `combine` of the mirrored layers against the reference layers'
own fallback path (message + aggregate, boundary as a self-loop message) for sum / mean / max / pna and both message
functions (tests/golden/make_golden.py `layer_case`; SURVEY.md section 8c test (2))."""
import glob
import os

import numpy as np
import pytest
import torch

from ultra_torchdrug_b200 import nbf
from ultra_torchdrug_b200.compat.torchdrug import data

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "layer_full_*.npz")))


def _run(path, device):
    g = np.load(path)
    kind, message_func, aggregate_func = os.path.basename(path)[len("layer_full_"):-4].split("_")
    num_node, num_relation, batch, dim = (int(x) for x in g["shape"])
    if kind == "nbf":
        layer = nbf.GeneralizedRelationalConvNBF(dim, dim, num_relation, dim, message_func, aggregate_func, dependent=True)
    else:
        layer = nbf.GeneralizedRelationalConvNBFMod(dim, dim, num_relation, dim, message_func, aggregate_func, project=True)
    state = {k[len("state/"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("state/")}
    layer.load_state_dict(state, strict=True)
    layer = layer.to(device)
    if kind == "nbfmod":
        layer.relation = torch.from_numpy(g["relation"]).to(device)
    graph = data.Graph(torch.from_numpy(g["edge_list"]), num_node=num_node, num_relation=num_relation).to(device)
    with graph.graph():
        graph.query = torch.from_numpy(g["query"]).to(device)
    with graph.node():
        graph.boundary = torch.from_numpy(g["boundary"]).to(device)
    input = torch.from_numpy(g["input"]).to(device)
    with torch.no_grad():
        update = layer.message_and_aggregate(graph, input)
        output = layer(graph, input)
    return g, update.cpu().numpy(), output.cpu().numpy()


def _compare(path, g, update, output, atol):
    # always: the reference's own fast path (run on the oracle operator when the goldens were made)
    np.testing.assert_allclose(update, g["fast_update"], rtol=1e-5, atol=atol)
    np.testing.assert_allclose(output, g["fast_output"], rtol=1e-4, atol=atol)
    # and the reference's fallback path - except pna x transe, where the reference's two paths disagree with each
    # other (sum of squared operands vs squared messages, layer.py:142/165/344/367 vs :92/279)
    if not path.endswith("transe_pna.npz"):
        np.testing.assert_allclose(update, g["update"], rtol=1e-4, atol=10 * atol)
        np.testing.assert_allclose(output, g["output"], rtol=1e-4, atol=10 * atol)


def test_layer_goldens_present():
    assert len(GOLDEN) == 16


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[11:-4] for p in GOLDEN])
def test_mirror_layer_on_oracle(path, monkeypatch):
    from oracle.rspmm_oracle import generalized_rspmm_oracle
    monkeypatch.setattr(nbf, "generalized_rspmm", generalized_rspmm_oracle)
    g, update, output = _run(path, torch.device("cpu"))
    _compare(path, g, update, output, 1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[11:-4] for p in GOLDEN])
def test_mirror_layer_on_cuda(cuda, path):
    torch.backends.cuda.matmul.allow_tf32 = False
    g, update, output = _run(path, cuda)
    _compare(path, g, update, output, 2e-5)


@pytest.mark.parametrize("seed", range(5))
def test_easy_edge_mask_equals_graph_match(seed):
    """`nbf.easy_edge_mask` (sort + binary search, no host sync) marks exactly the edges `graph.match` finds for the
    (h, t, r) rows of a training batch (reference model.py:57-74), duplicates and absent triples included."""
    generator = torch.Generator().manual_seed(seed)
    num_node, num_relation = 12, 3
    edge_list = torch.stack([torch.randint(num_node, (90,), generator=generator), torch.randint(num_node, (90,), generator=generator),
                             torch.randint(num_relation, (90,), generator=generator)], dim=1)
    graph = data.Graph(edge_list, num_node=num_node, num_relation=num_relation)
    h_index = torch.randint(num_node, (6, 4), generator=generator)
    t_index = torch.randint(num_node, (6, 4), generator=generator)
    r_index = torch.randint(num_relation, (6, 4), generator=generator)
    h_index[0], t_index[0], r_index[0] = edge_list[:4, 0], edge_list[:4, 1], edge_list[:4, 2]   # some certain matches
    pattern = torch.stack([h_index, t_index, r_index], dim=-1).flatten(0, -2)
    want = torch.zeros(graph.num_edge, dtype=torch.bool)
    want[graph.match(pattern)[0]] = True
    assert want.any()
    assert torch.equal(nbf.easy_edge_mask(graph, h_index, t_index, r_index), want)
    empty = torch.zeros(0, 3, dtype=torch.long)
    assert not nbf.easy_edge_mask(graph, empty[:, 0], empty[:, 1], empty[:, 2]).any()


@pytest.mark.gpu
def test_layer_uses_the_boundary_of_the_call_not_a_stale_one_hot(cuda):
    """the sparse form of the boundary travels with the call that built it.  After a bellmanford pass
    over a persistent graph object, a later layer call on the same graph with another `graph.boundary` must aggregate
    with that boundary."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    num_node, num_relation, dim, batch = 30, 4, 32, 3
    generator = torch.Generator().manual_seed(1)
    edge_list = torch.stack([torch.randint(num_node, (200,), generator=generator), torch.randint(num_node, (200,), generator=generator),
                             torch.randint(num_relation, (200,), generator=generator)], dim=1)
    graph = data.Graph(edge_list, num_node=num_node, num_relation=num_relation).to(cuda)
    model = nbf.CustomNBFNetFull(dim, [dim, dim], num_relation=num_relation, aggregate_func="sum", layer_norm=True,
                                 short_cut=True).to(cuda)
    h_index = torch.tensor([0, 5, 7], device=cuda)
    first = model(graph, h_index)                 # autograd on: the per-layer path, boundary handed down as one-hot
    assert first.requires_grad and not hasattr(graph, "boundary_one_hot")
    layer = model.layers[0]
    other_boundary = torch.randn(num_node, batch, dim, device=cuda)
    with graph.node():
        graph.boundary = other_boundary
    input = torch.randn(num_node, batch, dim, device=cuda)
    with torch.no_grad():
        update = layer.message_and_aggregate(graph, input)
        relation_input = layer.relation_input(graph, batch)
        plain = nbf.rspmm.generalized_rspmm(graph.adjacency.transpose(0, 1), relation_input, input.flatten(1))
    torch.testing.assert_close(update.flatten(1), plain + other_boundary.flatten(1), rtol=1e-5, atol=1e-5)
