"""Synthetic knowledge graphs of the shapes BASELINE.json names (datasets cannot be downloaded offline).

Follows SURVEY.md section 8d: triples (h, t, r) uniform with `torch.Generator().manual_seed(1024)`
(the reference's default seed, ultra/util.py:77), then the `undirected(add_inverse=True)` layout of
reference ultra/model.py:166 - edges [(h, t, r), (t, h, r + R)] interleaved, E = 2 * triples,
R' = 2R, edge_weight = 1.  Duplicate triples are kept so that coalescing is exercised.
"""
import torch

#: name -> (entities, relations (before inverses), triples)
SHAPES = {
    "fb15k237_ind_v1": (1594, 180, 4245),      # BASELINE.json configs[0]  (C1)
    "fb15k237": (14541, 237, 272115),          # configs[1]                 (C2)
    "codex_l": (77951, 69, 551193),            # configs[2]                 (C3)
    "yago310": (123182, 37, 1079040),          # configs[3]                 (C4)
    "wn18rr": (40943, 11, 86835),              # low-degree shape (average in-degree 4.2), not a BASELINE config
}


def triples(num_node, num_relation, num_triple, seed=1024, skew=None, relation_skew=None):
    """(num_triple, 3) int64 [h, t, r].  `skew`: Zipf exponent for the tail distribution (hub nodes);
    `relation_skew`: Zipf exponent for the relation distribution (a few frequent relation types, as in real KGs)."""
    generator = torch.Generator().manual_seed(seed)
    h = torch.randint(num_node, (num_triple,), generator=generator)
    if skew:
        rank = torch.arange(1, num_node + 1, dtype=torch.float64)
        t = torch.multinomial(rank.pow(-float(skew)), num_triple, replacement=True, generator=generator)
    else:
        t = torch.randint(num_node, (num_triple,), generator=generator)
    if relation_skew:
        rank = torch.arange(1, num_relation + 1, dtype=torch.float64)
        r = torch.multinomial(rank.pow(-float(relation_skew)), num_triple, replacement=True, generator=generator)
    else:
        r = torch.randint(num_relation, (num_triple,), generator=generator)
    return torch.stack([h, t, r], dim=-1)


def undirected_edge_list(triple, num_relation):
    """[node_in, node_out, rel] rows: every triple followed by its inverse with relation r + R."""
    h, t, r = triple.t()
    forward = torch.stack([h, t, r], dim=-1)
    inverse = torch.stack([t, h, r + num_relation], dim=-1)
    return torch.stack([forward, inverse], dim=1).flatten(0, 1)


def operator_operand(edge_list, num_node, num_relation, device=None, dtype=torch.float32):
    """The sparse COO operand the layer hands to generalized_rspmm: `graph.adjacency.transpose(0, 1)`
    (reference layer.py:127,328) = indices [node_out; node_in; rel], un-coalesced, unit values."""
    indices = edge_list[:, [1, 0, 2]].t().contiguous()
    values = torch.ones(indices.shape[1], dtype=dtype)
    if device is not None:
        indices, values = indices.to(device), values.to(device)
    return torch.sparse_coo_tensor(indices, values, (num_node, num_node, num_relation), check_invariants=False)


def named_graph(name, seed=1024, skew=None, relation_skew=None):
    """(edge_list (E, 3), num_node, num_relation incl. inverses) of a BASELINE.json shape."""
    num_node, num_relation, num_triple = SHAPES[name]
    edge_list = undirected_edge_list(triples(num_node, num_relation, num_triple, seed, skew, relation_skew), num_relation)
    return edge_list, num_node, 2 * num_relation
