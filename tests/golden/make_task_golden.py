"""Golden vectors of the reference's strict negative sampling and filter masks (authoring container only).

    python tests/golden/make_task_golden.py      # needs /root/reference; writes tests/golden/task_strict_negative.npz

Runs the unmodified `KnowledgeGraphCompletionBase._strict_negative`, `_calculate_t_mask` and `_calculate_h_mask`
(reference ultra/task.py:65-118) through the import shims on a small synthetic fact graph with duplicate triples.
`functional.variadic_sample` is torchdrug's (un-vendored); the shim restates it [ext-recall] - what is pinned here is
the reference's own candidate construction (masks, `nonzero` order, the tail / head split of the batch) and the way
the uniform numbers are consumed.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from ultra_torchdrug_b200 import compat, synthetic  # noqa: E402

compat.install()
compat.add_reference_to_path("/root/reference")

from torchdrug import data  # noqa: E402
from ultra import task as ref_task  # noqa: E402

NUM_NODE, NUM_RELATION, NUM_TRIPLE, BATCH, NEGATIVE, SEED = 60, 3, 700, 10, 32, 77


def main():
    triples = synthetic.triples(NUM_NODE, NUM_RELATION, NUM_TRIPLE, seed=5)
    triples = torch.cat([triples, triples[:50]])                       # duplicate facts
    graph = data.Graph(triples, num_node=NUM_NODE, num_relation=NUM_RELATION)
    base = ref_task.KnowledgeGraphCompletionBase
    fake = types.SimpleNamespace(fact_graph=graph, num_negative=NEGATIVE, device=torch.device("cpu"), num_entity=NUM_NODE)
    fake._calculate_t_mask = types.MethodType(base._calculate_t_mask, fake)
    fake._calculate_h_mask = types.MethodType(base._calculate_h_mask, fake)
    batch = triples[torch.randperm(NUM_TRIPLE, generator=torch.Generator().manual_seed(1))[:BATCH]]
    pos_h, pos_t, pos_r = batch.t()
    torch.manual_seed(SEED)
    negative = base._strict_negative.__wrapped__(fake, pos_h, pos_t, pos_r) if hasattr(base._strict_negative, "__wrapped__") \
        else base._strict_negative(fake, pos_h, pos_t, pos_r)
    t_mask = fake._calculate_t_mask(graph, pos_h, pos_r)
    h_mask = fake._calculate_h_mask(graph, pos_t, pos_r)
    path = os.path.join(HERE, "task_strict_negative.npz")
    np.savez_compressed(path, triples=triples.numpy(), batch=batch.numpy(), negative=negative.numpy(), t_mask=t_mask.numpy(),
                        h_mask=h_mask.numpy(), shape=np.array([NUM_NODE, NUM_RELATION, NEGATIVE, SEED]))
    print("wrote", path, negative.shape, "answers per row up to", int((~t_mask).sum(1).max()))


if __name__ == "__main__":
    main()
