"""Model-level parity: the host-side mirror (`ultra_torchdrug_b200.nbf`) against outputs of the reference's own
modules (tests/golden/make_model_golden.py) - scores, relation representations, relation graph, training
gradients, and ranking metrics.  CPU variant: the mirror runs on the oracle operator (checks the mirror);
GPU variant: the mirror runs on the CUDA operator (checks the product end to end)."""
import os

import numpy as np
import pytest
import torch

from ultra_torchdrug_b200 import nbf
from ultra_torchdrug_b200.compat.torchdrug import data

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "model_ultra_small.npz")


def _load(device):
    g = np.load(GOLDEN)
    num_node, num_relation, hidden, num_layers = (int(x) for x in g["shape"])
    model, rel_model = nbf.ultra_models(num_relation, hidden=hidden, num_layers=num_layers)
    for prefix, module in (("model/", model), ("rel_model/", rel_model)):
        state = {k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}
        module.load_state_dict(state, strict=True)      # same parameter names as the reference modules
    graph = data.Graph(torch.from_numpy(g["edge_list"]), num_node=num_node, num_relation=num_relation).to(device)
    return g, model.to(device), rel_model.to(device), graph


def _check(g, model, rel_model, graph, device, rtol, atol):
    ranker = nbf.UltraRanker(model, rel_model, graph)
    want = {tuple(e) for e in g["rel_graph_edge_list"].tolist()}
    assert {tuple(e) for e in ranker.rel_graph.edge_list.cpu().tolist()} == want
    batch = torch.from_numpy(g["batch"]).to(device)
    model.eval()
    rel_model.eval()
    with torch.no_grad():
        rel_input = rel_model(ranker.rel_graph, batch[:, 2])
        pred = ranker.predict(batch)
    np.testing.assert_allclose(rel_input.cpu().numpy(), g["rel_input"], rtol=rtol, atol=atol)
    np.testing.assert_allclose(pred.cpu().numpy(), g["pred"], rtol=rtol, atol=atol)
    # ranking metrics: identical ranks => MRR / Hits@10 equal to 4 decimals and beyond
    target = torch.stack([batch[:, 1], batch[:, 0]], dim=1)
    mask = ranker.filter_mask(batch)
    rank = ranker.rank(pred, target, mask).cpu()
    rank_ref = ranker.rank(torch.from_numpy(g["pred"]).to(device), target, mask).cpu()
    assert torch.equal(rank, rank_ref)
    got, ref = nbf.metrics(rank), nbf.metrics(rank_ref)
    assert round(got["mrr"], 4) == round(ref["mrr"], 4) and round(got["hits@10"], 4) == round(ref["hits@10"], 4)

    # training branch: remove_easy_edges + backward through both networks
    model.train()
    rel_model.train()
    h, t, r = (torch.from_numpy(g["train_%s_index" % k]).to(device) for k in "htr")
    weight = torch.from_numpy(g["train_weight"]).to(device)
    rel_input = rel_model(ranker.rel_graph, batch[:, 2])
    train_pred = model(graph, [rel_input], h, t, r, remove_easy_edges=True)
    np.testing.assert_allclose(train_pred.detach().cpu().numpy(), g["train_pred"], rtol=rtol, atol=atol)
    (train_pred * weight).sum().backward()
    checked = 0
    for prefix, module in (("grad/model/", model), ("grad/rel_model/", rel_model)):
        for name, parameter in module.named_parameters():
            key = prefix + name
            if key in g.files:
                scale = max(1.0, float(np.abs(g[key]).max()))
                np.testing.assert_allclose(parameter.grad.cpu().numpy(), g[key], rtol=10 * rtol, atol=10 * atol * scale,
                                           err_msg=key)
                checked += 1
            else:
                assert parameter.grad is None or float(parameter.grad.abs().max()) == 0.0, key
    assert checked >= 40


def test_mirror_on_oracle_matches_reference_modules(monkeypatch):
    from oracle.rspmm_oracle import generalized_rspmm_oracle
    monkeypatch.setattr(nbf, "generalized_rspmm", generalized_rspmm_oracle)
    g, model, rel_model, graph = _load(torch.device("cpu"))
    _check(g, model, rel_model, graph, torch.device("cpu"), rtol=1e-5, atol=1e-5)


@pytest.mark.gpu
def test_mirror_on_cuda_matches_reference_modules(cuda):
    from ultra_torchdrug_b200 import functional as F
    torch.backends.cuda.matmul.allow_tf32 = False     # reference script/run_full.py:19-20
    torch.backends.cudnn.allow_tf32 = False
    g, model, rel_model, graph = _load(cuda)
    before = F.launch_count()
    _check(g, model, rel_model, graph, cuda, rtol=2e-4, atol=2e-5)
    assert F.launch_count() > before


# ---- BASELINE.json configs[0]: FB15k237Inductive-v1 shape, shipped architecture (6 + 6 layers x 64-d) ------------------
GOLDEN_C1 = os.path.join(os.path.dirname(__file__), "golden", "model_ultra_c1.npz")


def _check_c1(device, rtol, atol):
    from ultra_torchdrug_b200 import synthetic
    g = np.load(GOLDEN_C1)
    num_node, num_relation, hidden, num_layers = (int(x) for x in g["shape"])
    model, rel_model = nbf.ultra_models(num_relation, hidden=hidden, num_layers=num_layers)
    for prefix, module in (("model/", model), ("rel_model/", rel_model)):
        module.load_state_dict({k[len(prefix):]: torch.from_numpy(g[k]) for k in g.files if k.startswith(prefix)}, strict=True)
    assert sum(p.numel() for p in model.parameters()) + sum(p.numel() for p in rel_model.parameters()) == 117505 + 76608
    triples = synthetic.triples(num_node, num_relation, synthetic.SHAPES["fb15k237_ind_v1"][2], seed=1024)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation).to(device)
    ranker = nbf.UltraRanker(model.to(device).eval(), rel_model.to(device).eval(), graph)
    assert ranker.rel_graph.num_edge == int(g["num_rel_graph_edge"])
    batch = torch.from_numpy(g["batch"]).to(device)
    with torch.no_grad():
        pred = ranker.predict(batch)
    want = torch.from_numpy(g["pred"]).to(device)
    torch.testing.assert_close(pred, want, rtol=rtol, atol=atol)
    target = torch.stack([batch[:, 1], batch[:, 0]], dim=1)
    mask = ranker.filter_mask(batch)
    got, ref = nbf.metrics(ranker.rank(pred, target, mask)), nbf.metrics(ranker.rank(want, target, mask))
    for name in ("mrr", "hits@10", "hits@1", "mr"):      # north_star: MRR / Hits@10 equal to 4 decimals
        assert round(got[name], 4) == round(ref[name], 4), (name, got, ref)


@pytest.mark.gpu
def test_c1_zero_shot_scores_and_mrr_on_cuda(cuda):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    _check_c1(cuda, rtol=1e-3, atol=1e-4)


@pytest.mark.gpu
def test_cuda_graph_predict_equals_eager(cuda):
    """`UltraRanker.capture`: one CUDA graph per batch size, replayed on fresh batches, reproduces eager `predict`."""
    from ultra_torchdrug_b200 import functional as F, synthetic
    torch.manual_seed(5)
    num_node, num_relation = 300, 6
    triples = synthetic.triples(num_node, num_relation, 1500, seed=5)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation).to(cuda)
    model, rel_model = nbf.ultra_models(num_relation, hidden=32, num_layers=3)
    ranker = nbf.UltraRanker(model.to(cuda).eval(), rel_model.to(cuda).eval(), graph)
    run = ranker.capture(batch_size=8)
    launches = F.launch_count()
    for seed in range(3):
        batch = triples[torch.randperm(len(triples), generator=torch.Generator().manual_seed(seed))[:8]].to(cuda)
        with torch.no_grad():
            eager = ranker.predict(batch)
        replayed = run(batch).clone()
        assert torch.equal(replayed, eager), "graph replay differs from eager execution"
    with pytest.raises(ValueError):
        run(batch[:4])
    assert F.launch_count() > launches


@pytest.mark.gpu
def test_copy_free_layer_loops_equal_generic_loop(cuda, monkeypatch):
    """The two cat-free inference loops - plain (N, B, d) planes read through two TMA tensor maps (`_run_layers_planes`, the
    default) and the interleaved (N, B, 2d) buffers (`_run_layers_buffered`) - give the scores of the generic loop (cat +
    Linear + epilogue): same arithmetic, only where the bytes live differs."""
    from ultra_torchdrug_b200 import functional as F, synthetic
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(11)
    num_node, num_relation = 400, 7
    triples = synthetic.triples(num_node, num_relation, 2500, seed=11)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation).to(cuda)
    model, rel_model = nbf.ultra_models(num_relation, hidden=64, num_layers=3)
    ranker = nbf.UltraRanker(model.to(cuda).eval(), rel_model.to(cuda).eval(), graph)
    batch = triples[:6].to(cuda)
    taken = {"planes": 0, "buffered": 0}
    planes, buffered = nbf._run_layers_planes, nbf._run_layers_buffered
    monkeypatch.setattr(nbf, "_run_layers_planes", lambda *a, **k: taken.__setitem__("planes", taken["planes"] + 1) or planes(*a, **k))
    monkeypatch.setattr(nbf, "_run_layers_buffered", lambda *a, **k: taken.__setitem__("buffered", taken["buffered"] + 1) or buffered(*a, **k))
    with torch.no_grad():
        from_planes = ranker.predict(batch)
    assert taken == {"planes": 3, "buffered": 0}, "the plane loop must serve the relation pass and both entity passes"
    monkeypatch.setattr(F, "linear_planes_supported", lambda hidden, out_dim: False)
    with torch.no_grad():
        from_buffers = ranker.predict(batch)
    assert taken == {"planes": 3, "buffered": 3}
    monkeypatch.setattr(nbf, "_buffered_layers_supported", lambda layers, boundary: False)
    with torch.no_grad():
        generic = ranker.predict(batch)
    assert taken == {"planes": 3, "buffered": 3}
    torch.testing.assert_close(from_planes, generic, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(from_buffers, generic, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(from_planes, from_buffers, rtol=1e-6, atol=1e-6)
    with torch.enable_grad():                        # training keeps the autograd path
        ranker.predict(batch).sum().backward()
    assert taken == {"planes": 3, "buffered": 3}


@pytest.mark.gpu
def test_masking_easy_edges_equals_removing_them(cuda, monkeypatch):
    """Training forward with `remove_easy_edges`: weight-0 masking on the full graph's index (GraphIndex.derive) gives
    the scores and parameter gradients of the reference's edge removal (model.py:57-74) - and builds no new index."""
    from ultra_torchdrug_b200 import functional as F, synthetic
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(3)
    num_node, num_relation = 200, 5
    triples = synthetic.triples(num_node, num_relation, 1500, seed=3)
    triples = torch.cat([triples, triples[:40]])                                  # duplicate edges: merged weights
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation).to(cuda)
    model, rel_model = nbf.ultra_models(num_relation, hidden=32, num_layers=3)
    model, rel_model = model.to(cuda).train(), rel_model.to(cuda).train()
    rel_graph = nbf.construct_relation_graph(graph)
    batch = triples[:8].to(cuda)
    negatives = torch.randint(num_node, (8, 4), device=cuda)
    h_index, t_index, r_index = (batch[:, i].unsqueeze(-1).repeat(1, 5) for i in range(3))
    t_index[:4, 1:] = negatives[:4]
    h_index[4:, 1:] = negatives[4:]
    parameters = list(model.parameters()) + list(rel_model.parameters())

    def run():
        for p in parameters:
            p.grad = None
        rel_input = rel_model(rel_graph, batch[:, 2])
        pred = model(graph, [rel_input], h_index, t_index, r_index, remove_easy_edges=True)
        pred.square().sum().backward()
        return pred.detach().clone(), [None if p.grad is None else p.grad.clone() for p in parameters]

    run()                                                                        # builds the index of the full graph
    built = F.cache_stats["built"]
    masked_pred, masked_grads = run()
    assert F.cache_stats["built"] == built, "masking must not build another index"
    monkeypatch.setattr(nbf.TransferNBFNet, "_can_mask_easy_edges", lambda self, graph: False)
    removed_pred, removed_grads = run()
    assert F.cache_stats["built"] > built
    torch.testing.assert_close(masked_pred, removed_pred, rtol=1e-4, atol=1e-5)
    for a, b in zip(masked_grads, removed_grads):
        assert (a is None) == (b is None)
        if a is not None:
            torch.testing.assert_close(a, b, rtol=1e-3, atol=1e-5 * (1 + float(b.abs().max())))
