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
