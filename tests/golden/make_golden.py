"""Generates the golden vectors in this directory from the REFERENCE's own code (authoring container only).

    python tests/golden/make_golden.py            # needs /root/reference; writes tests/golden/*.npz

The reference states the rspmm math a second time, in plain PyTorch: `message()` + `aggregate()` of
`GeneralizedRelationalConvNBF` / `GeneralizedRelationalConvNBFMod` (reference ultra/layer.py:52-109, 232-296),
the fallback `message_and_aggregate` takes when `graph.requires_grad` (layer.py:112, 299).  This script imports the
UNMODIFIED `ultra/layer.py` through the import shims (`ultra_torchdrug_b200.compat`), runs those two methods on
small seeded graphs and stores, per case, the operands exactly as the fast path hands them to
`generalized_rspmm` (layer.py:118-127, 306-328) together with the fallback's results:

  out_add   aggregate_func="sum",  boundary = 0          -> scatter_add(messages + zero self-loop)  == operator "add"
  out_max   aggregate_func="max",  boundary = -FLT_MAX   -> scatter_max(messages, lowest)           == operator "max"
  out_min   aggregate_func="pna",  boundary = +FLT_MAX   -> the `min` feature (scale column 0)      == operator "min"
  grad_*    autograd of the "sum" case for a fixed upstream gradient

Graphs used for max/min have no duplicate triples (the fast path coalesces duplicates, the fallback does not:
SURVEY.md hard-part 2); the "sum" cases keep duplicates (w = 2 is exact either way).  Max/min *gradients* are not
taken from the fallback: the shimmed torch_scatter does not implement torchdrug's all-ties rule.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from ultra_torchdrug_b200 import compat  # noqa: E402

compat.install()
compat.add_reference_to_path("/root/reference")

import torchdrug.layers.functional as td_functional  # noqa: E402  (the shim)
from oracle.rspmm_oracle import generalized_rspmm_oracle  # noqa: E402
from torchdrug import data  # noqa: E402  (the shim)
from ultra import layer as ref_layer  # noqa: E402  (the unmodified reference module)

td_functional.generalized_rspmm = generalized_rspmm_oracle   # the reference's fast path runs on the CPU oracle here

FLT_MAX = float(np.finfo(np.float32).max)


def make_graph(num_node, num_relation, num_edge, seed, duplicates):
    rng = np.random.default_rng(seed)
    triples = set()
    while len(triples) < num_edge:
        triples.add((int(rng.integers(num_node)), int(rng.integers(num_node)), int(rng.integers(num_relation))))
    edge_list = np.array(sorted(triples), dtype=np.int64)
    edge_list = edge_list[rng.permutation(len(edge_list))]
    if duplicates:
        edge_list = np.concatenate([edge_list, edge_list[rng.integers(0, len(edge_list), duplicates)]])
    # leave a few nodes without incoming edges (empty destination rows)
    edge_list = edge_list[edge_list[:, 1] % 7 != 3]
    return torch.from_numpy(edge_list)


def run_fallback(layer, graph, input, boundary, query):
    with graph.graph():
        graph.query = query
    with graph.node():
        graph.boundary = boundary
    message = layer.message(graph, input)       # reference layer.py:52-79 / 232-266
    return layer.aggregate(graph, message)      # reference layer.py:81-109 / 268-296


def one_case(kind, message_func, seed, duplicates):
    torch.manual_seed(seed)
    num_node, num_relation, batch, dim = 40, 6, 3, 8
    edge_list = make_graph(num_node, num_relation, 260, seed, duplicates)
    graph = data.Graph(edge_list, num_node=num_node, num_relation=num_relation)
    query = torch.randn(batch, dim)
    input = torch.randn(num_node, batch, dim, requires_grad=True)
    grad_output = torch.randn(num_node, batch, dim)

    def build(aggregate_func):
        if kind == "nbf":
            layer = ref_layer.GeneralizedRelationalConvNBF(dim, dim, num_relation, dim, message_func, aggregate_func,
                                                           dependent=False)
            with torch.no_grad():
                layer.relation.weight.copy_(relation_weight)
            return layer, layer.relation.weight
        layer = ref_layer.GeneralizedRelationalConvNBFMod(dim, dim, num_relation, dim, message_func, aggregate_func,
                                                          project=False)
        layer.relation = relation_batched
        return layer, relation_batched

    relation_weight = torch.randn(num_relation, dim)
    relation_batched = torch.randn(batch, num_relation, dim, requires_grad=True)

    # operands as the fast path builds them (layer.py:118-127 / 306-328)
    if kind == "nbf":
        relation_input = relation_weight.repeat(1, batch)
    else:
        relation_input = relation_batched.detach().transpose(1, 0).flatten(1)
    adjacency = graph.adjacency.transpose(0, 1)
    record = {
        "indices": adjacency._indices().numpy().copy(), "values": adjacency._values().numpy().copy(),
        "shape": np.array(adjacency.shape, dtype=np.int64), "relation": relation_input.numpy().copy(),
        "input": input.detach().flatten(1).numpy().copy(), "grad_output": grad_output.flatten(1).numpy().copy(),
    }

    layer, parameter = build("sum")
    update = run_fallback(layer, graph, input, torch.zeros(num_node, batch, dim), query)
    record["out_add"] = update.detach().flatten(1).numpy().copy()
    (update * grad_output).sum().backward()
    record["grad_input_add"] = input.grad.flatten(1).numpy().copy()
    if kind == "nbf":
        record["grad_relation_weight_add"] = parameter.grad.numpy().copy()      # (R, d): summed over the batch
    else:
        record["grad_relation_add"] = parameter.grad.transpose(1, 0).flatten(1).numpy().copy()

    if not duplicates:
        with torch.no_grad():
            layer, _ = build("max")
            update = run_fallback(layer, graph, input, torch.full((num_node, batch, dim), -FLT_MAX), query)
            record["out_max"] = update.flatten(1).numpy().copy()
            layer, _ = build("pna")
            update = run_fallback(layer, graph, input, torch.full((num_node, batch, dim), FLT_MAX), query)
            # update[..., (c * 4 + f) * 3 + s]: feature f = 2 is `min`, scale s = 0 is 1 (layer.py:96-104)
            minimum = update.view(num_node, batch, dim, 4, 3)[..., 2, 0]
            record["out_min"] = minimum.flatten(1).numpy().copy()
    return record


def layer_case(kind, message_func, aggregate_func, seed):
    """Full `message_and_aggregate` of the reference layer through its fallback (message + aggregate with the
    boundary as a self-loop message, layer.py:77, 83-84) - what the fast path's operator call + post-ops
    (layer.py:154-180, 356-382) must reproduce.  No duplicate triples (see module docstring)."""
    torch.manual_seed(seed)
    num_node, num_relation, batch, dim = 30, 5, 2, 8
    edge_list = make_graph(num_node, num_relation, 150, seed, duplicates=0)
    graph = data.Graph(edge_list, num_node=num_node, num_relation=num_relation)
    query = torch.randn(batch, dim)
    input = torch.randn(num_node, batch, dim)
    boundary = torch.randn(num_node, batch, dim)
    if kind == "nbf":
        layer = ref_layer.GeneralizedRelationalConvNBF(dim, dim, num_relation, dim, message_func, aggregate_func,
                                                       dependent=True)
    else:
        layer = ref_layer.GeneralizedRelationalConvNBFMod(dim, dim, num_relation, dim, message_func, aggregate_func,
                                                          project=True)
        layer.relation = torch.randn(batch, num_relation, dim)
    with torch.no_grad():
        update = run_fallback(layer, graph, input, boundary, query)
        with graph.graph():
            graph.query = query
        with graph.node():
            graph.boundary = boundary
        output = layer.combine(input, update)
        # the reference's own fast path (layer.py:111-182 / 298-384) on the oracle operator.  For pna x transe it
        # differs from the fallback by construction: it squares the operands (rel^2 + in^2), the fallback squares
        # the message ((rel + in)^2) - a property of the reference, kept as is.
        fast_update = layer.message_and_aggregate(graph, input)
        fast_output = layer.combine(input, fast_update)
    record = {"edge_list": edge_list.numpy(), "shape": np.array([num_node, num_relation, batch, dim]),
              "query": query.numpy(), "input": input.numpy(), "boundary": boundary.numpy(),
              "update": update.numpy().copy(), "output": output.numpy().copy(),
              "fast_update": fast_update.numpy().copy(), "fast_output": fast_output.numpy().copy()}
    if kind == "nbfmod":
        record["relation"] = layer.relation.numpy().copy()
    for name, tensor in layer.state_dict().items():
        record["state/" + name] = tensor.numpy().copy()
    return record


def main():
    for kind in ("nbf", "nbfmod"):
        for message_func in ("distmult", "transe"):
            for aggregate_func in ("sum", "mean", "max", "pna"):
                record = layer_case(kind, message_func, aggregate_func, seed=7)
                name = "layer_full_%s_%s_%s.npz" % (kind, message_func, aggregate_func)
                np.savez_compressed(os.path.join(HERE, name), **record)
                print("wrote", name)
    for kind in ("nbf", "nbfmod"):
        for message_func in ("distmult", "transe"):
            for duplicates in (0, 25):
                record = one_case(kind, message_func, seed=1024 + duplicates, duplicates=duplicates)
                name = "layer_fallback_%s_%s_%s.npz" % (kind, message_func, "dup" if duplicates else "nodup")
                np.savez_compressed(os.path.join(HERE, name), **record)
                print("wrote", name, {k: v.shape for k, v in record.items()})


if __name__ == "__main__":
    main()
