"""Where a ULTRA `predict` batch spends its time (development tool): relation-model pass vs entity passes, and the
share of rspmm kernels (library launch list) in each."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_torchdrug_b200 import functional as F, nbf, synthetic  # noqa: E402
from ultra_torchdrug_b200.compat.torchdrug import data  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "fb15k237"
batch_size = int(sys.argv[2]) if len(sys.argv) > 2 else 64
device = torch.device("cuda", 0)
torch.backends.cuda.matmul.allow_tf32 = False
num_node, num_relation, num_triple = synthetic.SHAPES[name]
triples = synthetic.triples(num_node, num_relation, num_triple)
graph = data.Graph(triples, num_node=num_node, num_relation=num_relation).to(device)
model, rel_model = nbf.ultra_models(num_relation)
ranker = nbf.UltraRanker(model.to(device).eval(), rel_model.to(device).eval(), graph)
print("relation graph: %d nodes, %d edges" % (ranker.rel_graph.num_node, ranker.rel_graph.num_edge))


def timed(fn, n=5):
    fn(); fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        out = fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n, (time.perf_counter() - t0) / n * 1e3, out


with torch.no_grad():
    batch = triples[torch.randint(num_triple, (batch_size,))].to(device)
    ms, wall, rel_input = timed(lambda: rel_model(ranker.rel_graph, batch[:, 2]))
    print("relation model pass   %8.3f ms device  %8.3f ms wall" % (ms, wall))
    cand = torch.arange(num_node, device=device)
    r_index = batch[:, 2].unsqueeze(-1).expand(-1, num_node)
    h_index, t_index = torch.meshgrid(batch[:, 0], cand, indexing="ij")
    ms, wall, _ = timed(lambda: model(graph, [rel_input], h_index, t_index, r_index))
    print("entity model pass     %8.3f ms device  %8.3f ms wall" % (ms, wall))
    ms, wall, _ = timed(lambda: ranker.predict(batch))
    print("predict (1 rel + 2 entity passes) %8.3f ms device  -> %.0f queries/s" % (ms, 2 * batch_size / ms * 1e3))
    run = ranker.capture(batch_size)
    ms, wall, _ = timed(lambda: run(batch))
    print("predict, CUDA graph replay        %8.3f ms device  %8.3f ms wall -> %.0f queries/s" % (ms, wall, 2 * batch_size / ms * 1e3))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ranker.predict(batch)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
