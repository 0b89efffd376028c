"""Runs the UNMODIFIED reference modules (`/root/reference/ultra/{layer,model,rel_model}.py`) through the import shims
on the CPU oracle operator and compares them with the host-side mirror on fresh random graphs.  Only possible where
the reference tree exists (the authoring container); skipped elsewhere - the committed golden vectors
(tests/golden/) carry the same evidence to the GPU box."""
import os
import sys

import numpy as np
import pytest
import torch

REFERENCE = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "ultra")), reason="reference tree not present")


@pytest.fixture(scope="module")
def reference():
    from ultra_torchdrug_b200 import compat
    compat.install()
    compat.add_reference_to_path(REFERENCE)
    import torchdrug.layers.functional as td_functional
    from oracle.rspmm_oracle import generalized_rspmm_oracle
    previous = td_functional.generalized_rspmm
    td_functional.generalized_rspmm = generalized_rspmm_oracle
    from ultra import layer, model, rel_model
    yield {"layer": layer, "model": model, "rel_model": rel_model}
    td_functional.generalized_rspmm = previous
    for name in [n for n in sys.modules if n == "ultra" or n.startswith("ultra.")]:
        del sys.modules[name]


@pytest.mark.parametrize("seed", [1, 2])
def test_reference_models_equal_mirror(reference, monkeypatch, seed):
    from oracle.rspmm_oracle import generalized_rspmm_oracle
    from torchdrug import data
    from ultra_torchdrug_b200 import nbf, synthetic
    monkeypatch.setattr(nbf, "generalized_rspmm", generalized_rspmm_oracle)
    torch.manual_seed(seed)
    num_node, num_relation, hidden, layers, batch_size = 40 + seed, 3 + seed, 8, 2, 3
    triples = synthetic.triples(num_node, num_relation, 150, seed=seed)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation)
    ref_model = reference["model"].TransferNBFNet(
        input_dim=hidden, hidden_dims=[hidden] * layers, num_relation=num_relation, message_func="distmult",
        aggregate_func="sum", short_cut=True, layer_norm=True, project=True, mod=True).eval()
    ref_rel = reference["rel_model"].RelNBFNet(input_dim=hidden, hidden=hidden, num_layers=layers, input_type="ones").eval()
    model, rel_model = nbf.ultra_models(num_relation, hidden=hidden, num_layers=layers)
    model.load_state_dict(ref_model.state_dict(), strict=True)
    rel_model.load_state_dict(ref_rel.state_dict(), strict=True)
    ranker = nbf.UltraRanker(model.eval(), rel_model.eval(), graph)
    ref_rel_graph = ref_rel.construct_relation_graph(graph)
    assert {tuple(e) for e in ref_rel_graph.edge_list.tolist()} == {tuple(e) for e in ranker.rel_graph.edge_list.tolist()}
    batch = triples[torch.randperm(len(triples))[:batch_size]]
    pos_h, pos_t, pos_r = batch.t()
    with torch.no_grad():
        rel_input = ref_rel(ref_rel_graph, None, pos_r)["node_feature"]
        candidates = torch.arange(num_node)
        r_index = pos_r.unsqueeze(-1).expand(-1, num_node)
        h_index, t_index = torch.meshgrid(pos_h, candidates, indexing="ij")
        t_pred = ref_model(graph, [rel_input], h_index, t_index, r_index)
        t_index, h_index = torch.meshgrid(pos_t, candidates, indexing="ij")
        h_pred = ref_model(graph, [rel_input], h_index, t_index, r_index)
        want = torch.stack([t_pred, h_pred], dim=1)
        got = ranker.predict(batch)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("aggregate_func", ["sum", "mean", "max", "pna"])
def test_reference_layer_fast_path_equals_its_fallback(reference, aggregate_func):
    """The in-reference cross-check of SURVEY.md section 4: the same layer through the rspmm fast path (oracle operator)
    and through message() + aggregate() (forced by graph.requires_grad, layer.py:299)."""
    from torchdrug import data
    from ultra_torchdrug_b200 import synthetic
    torch.manual_seed(3)
    num_node, num_relation, batch, dim = 25, 4, 2, 8
    triples = torch.unique(synthetic.triples(num_node, num_relation, 120, seed=3), dim=0)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation)
    layer = reference["layer"].GeneralizedRelationalConvNBFMod(dim, dim, num_relation, dim, "distmult", aggregate_func,
                                                               project=True)
    layer.relation = torch.randn(batch, num_relation, dim)
    input = torch.randn(num_node, batch, dim)
    with graph.graph():
        graph.query = torch.randn(batch, dim)
    with graph.node():
        graph.boundary = torch.randn(num_node, batch, dim)
    with torch.no_grad():
        fast = layer.message_and_aggregate(graph, input)
        graph.requires_grad = True
        slow = layer.message_and_aggregate(graph, input)
    np.testing.assert_allclose(fast.numpy(), slow.numpy(), rtol=1e-4, atol=1e-5)


def test_reference_concat_hidden_equals_mirror(reference, monkeypatch):
    """`concat_hidden=True` (reference model.py:49, 134-136: every layer's state feeds the scoring MLP) - unused by the
    shipped configs, mirrored all the same."""
    from oracle.rspmm_oracle import generalized_rspmm_oracle
    from torchdrug import data
    from ultra_torchdrug_b200 import nbf, synthetic
    monkeypatch.setattr(nbf, "generalized_rspmm", generalized_rspmm_oracle)
    torch.manual_seed(5)
    num_node, num_relation, hidden, layers = 35, 4, 8, 3
    triples = synthetic.triples(num_node, num_relation, 140, seed=5)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation)
    arguments = dict(input_dim=hidden, hidden_dims=[hidden] * layers, num_relation=num_relation, message_func="distmult",
                     aggregate_func="sum", short_cut=True, layer_norm=True, project=True, mod=True, concat_hidden=True)
    ref_model = reference["model"].TransferNBFNet(**arguments).eval()
    model = nbf.TransferNBFNet(**arguments).eval()
    model.load_state_dict(ref_model.state_dict(), strict=True)
    assert model.mlp.layers[0].in_features == hidden * layers + hidden
    batch = triples[:4]
    pos_h, pos_t, pos_r = batch.t()
    rel_input = torch.randn(len(batch), 2 * num_relation, hidden)
    candidates = torch.arange(num_node)
    r_index = pos_r.unsqueeze(-1).expand(-1, num_node)
    h_index, t_index = torch.meshgrid(pos_h, candidates, indexing="ij")
    with torch.no_grad():
        want = ref_model(graph, [rel_input], h_index, t_index, r_index)
        got = model(graph, [rel_input], h_index, t_index, r_index)
    torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
