"""Model-level golden vectors from the REFERENCE's own modules (authoring container only).

    python tests/golden/make_model_golden.py      # needs /root/reference; writes tests/golden/model_ultra_small.npz

Runs the unmodified `ultra/model.py` (TransferNBFNet, mod=True) and `ultra/rel_model.py` (RelNBFNet) through the
import shims with seeded random weights on a small synthetic graph, following the evaluation and training branches
of `KnowledgeGraphCompletionAdapted.predict` (reference ultra/task.py:228-277).  There is no GPU here, so the
operator under the reference modules is the CPU oracle (`oracle.rspmm_oracle.generalized_rspmm_oracle`, itself
pinned against the reference's message()+aggregate() code by make_golden.py).  Stored: the graph, the batch, every
parameter, and the reference's outputs - relation representations, (B, 2, N) scores, the relation graph, and the
gradients of a weighted sum of training scores w.r.t. every parameter.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from ultra_torchdrug_b200 import compat, synthetic  # noqa: E402

compat.install()
compat.add_reference_to_path("/root/reference")

import torchdrug.layers.functional as td_functional  # noqa: E402  (the shim)
from oracle.rspmm_oracle import generalized_rspmm_oracle  # noqa: E402
from torchdrug import data  # noqa: E402

td_functional.generalized_rspmm = generalized_rspmm_oracle   # ultra/layer.py looks the operator up at call time

from ultra import model as ref_model, rel_model as ref_rel_model  # noqa: E402

NUM_NODE, NUM_RELATION, NUM_TRIPLE, HIDDEN, LAYERS, BATCH, NEGATIVE = 50, 4, 220, 16, 3, 4, 6


def main():
    torch.manual_seed(1024)
    triples = synthetic.triples(NUM_NODE, NUM_RELATION, NUM_TRIPLE, seed=1024)
    edge_list = triples[:, [0, 1, 2]]                                     # [h, t, r] = [node_in, node_out, rel]
    graph = data.Graph(edge_list, num_node=NUM_NODE, num_relation=NUM_RELATION)
    model = ref_model.TransferNBFNet(input_dim=HIDDEN, hidden_dims=[HIDDEN] * LAYERS, num_relation=NUM_RELATION,
                                     message_func="distmult", aggregate_func="sum", short_cut=True, layer_norm=True,
                                     project=True, mod=True)
    rel_model = ref_rel_model.RelNBFNet(input_dim=HIDDEN, hidden=HIDDEN, num_layers=LAYERS, input_type="ones")
    rel_graph = rel_model.construct_relation_graph(graph)
    batch = triples[torch.randperm(NUM_TRIPLE)[:BATCH]]
    pos_h, pos_t, pos_r = batch.t()

    record = {"edge_list": edge_list.numpy(), "batch": batch.numpy(),
              "shape": np.array([NUM_NODE, NUM_RELATION, HIDDEN, LAYERS]),
              "rel_graph_edge_list": rel_graph.edge_list.numpy()}
    for prefix, module in (("model/", model), ("rel_model/", rel_model)):
        for name, tensor in module.state_dict().items():
            record[prefix + name] = tensor.numpy().copy()

    # ---- evaluation branch (task.py:238-263, full_batch_eval) ----------------------------------
    model.eval()
    rel_model.eval()
    with torch.no_grad():
        rel_input = rel_model(rel_graph, None, pos_r)["node_feature"]
        candidates = torch.arange(NUM_NODE)
        r_index = pos_r.unsqueeze(-1).expand(-1, NUM_NODE)
        h_index, t_index = torch.meshgrid(pos_h, candidates, indexing="ij")
        t_pred = model(graph, [rel_input], h_index, t_index, r_index)
        t_index, h_index = torch.meshgrid(pos_t, candidates, indexing="ij")
        h_pred = model(graph, [rel_input], h_index, t_index, r_index)
    record["rel_input"] = rel_input.numpy().copy()
    record["pred"] = torch.stack([t_pred, h_pred], dim=1).numpy().copy()

    # ---- training branch (task.py:264-275): negatives, remove_easy_edges, backward ----------------
    model.train()
    rel_model.train()
    negative = torch.randint(NUM_NODE, (BATCH, NEGATIVE))
    h_index = pos_h.unsqueeze(-1).repeat(1, NEGATIVE + 1)
    t_index = pos_t.unsqueeze(-1).repeat(1, NEGATIVE + 1)
    r_index = pos_r.unsqueeze(-1).repeat(1, NEGATIVE + 1)
    t_index[:BATCH // 2, 1:] = negative[:BATCH // 2]
    h_index[BATCH // 2:, 1:] = negative[BATCH // 2:]
    weight = torch.randn(BATCH, NEGATIVE + 1)
    rel_input = rel_model(rel_graph, None, pos_r, all_loss=torch.zeros(1), metric={})["node_feature"]
    pred = model(graph, [rel_input], h_index, t_index, r_index, all_loss=torch.zeros(1), metric={})
    (pred * weight).sum().backward()
    record.update({"train_h_index": h_index.numpy(), "train_t_index": t_index.numpy(), "train_r_index": r_index.numpy(),
                   "train_weight": weight.numpy(), "train_pred": pred.detach().numpy().copy()})
    for prefix, module in (("grad/model/", model), ("grad/rel_model/", rel_model)):
        for name, parameter in module.named_parameters():
            if parameter.grad is not None:
                record[prefix + name] = parameter.grad.numpy().copy()
    path = os.path.join(HERE, "model_ultra_small.npz")
    np.savez_compressed(path, **record)
    print("wrote", path, "keys:", len(record), "pred", record["pred"].shape, "grads",
          len([k for k in record if k.startswith("grad/")]))


def c1_case():
    """BASELINE.json configs[0] shape (FB15k237Inductive-v1: 1,594 entities, 180 relations, 4,245 triples) with the
    shipped architecture (6 + 6 layers x 64-d, reference config/transductive/inference.yaml), seeded random weights
    (the td_ultra_4g.pth blob is missing from the reference tree), evaluation branch, 8 test triples = 16 queries."""
    torch.manual_seed(1024)
    num_node, num_relation, num_triple = synthetic.SHAPES["fb15k237_ind_v1"]
    triples = synthetic.triples(num_node, num_relation, num_triple, seed=1024)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation)
    model = ref_model.TransferNBFNet(input_dim=64, hidden_dims=[64] * 6, num_relation=num_relation,
                                     message_func="distmult", aggregate_func="sum", short_cut=True, layer_norm=True,
                                     project=True, mod=True)
    rel_model = ref_rel_model.RelNBFNet(input_dim=64, hidden=64, num_layers=6, input_type="ones")
    rel_graph = rel_model.construct_relation_graph(graph)
    batch = triples[torch.randperm(num_triple)[:8]]
    pos_h, pos_t, pos_r = batch.t()
    model.eval()
    rel_model.eval()
    with torch.no_grad():
        rel_input = rel_model(rel_graph, None, pos_r)["node_feature"]
        candidates = torch.arange(num_node)
        r_index = pos_r.unsqueeze(-1).expand(-1, num_node)
        h_index, t_index = torch.meshgrid(pos_h, candidates, indexing="ij")
        t_pred = model(graph, [rel_input], h_index, t_index, r_index)
        t_index, h_index = torch.meshgrid(pos_t, candidates, indexing="ij")
        h_pred = model(graph, [rel_input], h_index, t_index, r_index)
    record = {"batch": batch.numpy(), "shape": np.array([num_node, num_relation, 64, 6]),
              "num_rel_graph_edge": np.array(rel_graph.num_edge), "pred": torch.stack([t_pred, h_pred], dim=1).numpy().copy()}
    for prefix, module in (("model/", model), ("rel_model/", rel_model)):
        for name, tensor in module.state_dict().items():
            record[prefix + name] = tensor.numpy().copy()
    path = os.path.join(HERE, "model_ultra_c1.npz")
    np.savez_compressed(path, **record)
    print("wrote", path, "pred", record["pred"].shape, "relation graph edges", rel_graph.num_edge)


if __name__ == "__main__":
    main()
    c1_case()
