"""What an UNPATCHED reference layer pays per operator call (development tool).

reference ultra/layer.py:127 / :328 hand `graph.adjacency.transpose(0, 1)` to the operator - a fresh tensor object per
layer call - so the index attached to the previous object is not found and `functional.graph_index` falls back to the
content fingerprint (one small kernel + a 16-byte device-to-host read = one host synchronisation per call).  This
script times a 6-layer loop both ways at a named shape: (a) a fresh transpose per call, as the unmodified layer does,
(b) one operand object reused (what `nbf.py` and `integrate.patch_torchdrug()` arrange).

    python tools/unpatched_overhead.py [--graph fb15k237] [--batch 64]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_torchdrug_b200 import functional as F, synthetic  # noqa: E402
from ultra_torchdrug_b200.compat.torchdrug import data  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--graph", default="fb15k237")
    parser.add_argument("--batch", type=int, default=64)
    args = parser.parse_args()
    device = torch.device("cuda", 0)
    num_node, num_relation, num_triple = synthetic.SHAPES[args.graph]
    graph = data.Graph(synthetic.triples(num_node, num_relation, num_triple), num_node=num_node,
                       num_relation=num_relation).to(device).undirected(add_inverse=True)
    d = args.batch * 64
    relation = torch.randn(graph.num_relation, d, device=device)
    hidden = torch.randn(num_node, d, device=device)

    # torchdrug's `graph.adjacency` is a cached plain sparse tensor: every `.transpose(0, 1)` is a new tensor object (the
    # import shim of this repo memoises it - not used here on purpose)
    adjacency = torch.sparse_coo_tensor(graph.edge_list.t(), graph.edge_weight, (num_node, num_node, graph.num_relation),
                                        check_invariants=False)

    def loop(fresh):
        reused = graph.adjacency.transpose(0, 1)
        with torch.no_grad():
            for _ in range(6):
                operand = adjacency.transpose(0, 1) if fresh else reused
                F.generalized_rspmm(operand, relation, hidden, sum="add", mul="mul")

    for fresh in (True, False):
        loop(fresh)
        torch.cuda.synchronize()
        before = dict(F.cache_stats)
        start = time.perf_counter()
        for _ in range(10):
            loop(fresh)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - start) / 10 * 1e3
        stats = {k: F.cache_stats[k] - before[k] for k in before}
        print("%-44s %7.3f ms per 6-layer loop   lookups: %s"
              % ("fresh transpose per call (unpatched layer.py)" if fresh else "one operand object (mirror / patched)", wall, stats))


if __name__ == "__main__":
    main()
