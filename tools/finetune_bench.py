"""ULTRA fine-tuning step (fwd + bwd + gradient all-reduce + AdamW) with the query batch sharded over the ranks
(BASELINE.json configs[2]: CoDEx-L-shaped synthetic graph, 64 triples per step over all ranks, 128 negatives).

    python tools/finetune_bench.py [--graph codex_l] [--batch 64] [--steps 5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/finetune_bench.py

Every rank holds the full graph and both networks; rank g trains on triples [g*B/G, (g+1)*B/G) of the step's batch
(reference ultra/engine.py:48-60 shards the same way with DistributedSampler + DDP); the only collective is one flat
all-reduce of the ~0.78 MB of gradients (`sharding.all_reduce_gradients`).  Strong scaling: the global batch is fixed.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_torchdrug_b200 import functional as F, nbf, sharding, synthetic  # noqa: E402
from ultra_torchdrug_b200.compat.torchdrug import data  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--graph", default="codex_l")
    parser.add_argument("--batch", type=int, default=64)
    parser.add_argument("--negatives", type=int, default=128)
    parser.add_argument("--steps", type=int, default=5)
    args = parser.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    torch.backends.cuda.matmul.allow_tf32 = False
    num_node, num_relation, num_triple = synthetic.SHAPES[args.graph]
    triples = synthetic.triples(num_node, num_relation, num_triple)
    graph = data.Graph(triples, num_node=num_node, num_relation=num_relation).to(device)
    torch.manual_seed(1024)                       # same initial weights on every rank
    model, rel_model = nbf.ultra_models(num_relation)
    model, rel_model = model.to(device).train(), rel_model.to(device).train()
    rel_graph = nbf.construct_relation_graph(graph)
    parameters = list(model.parameters()) + list(rel_model.parameters())
    optimizer = torch.optim.AdamW(parameters, lr=5e-4)
    generator = torch.Generator().manual_seed(7)   # same batches on every rank, each takes its slab

    def step():
        batch = triples[torch.randint(num_triple, (args.batch,), generator=generator)]
        negative = torch.randint(num_node, (args.batch, args.negatives), generator=generator)
        start, stop = sharding.query_slab(args.batch, rank, world)
        batch, negative = batch[start:stop].to(device), negative[start:stop].to(device)
        pos_h, pos_t, pos_r = batch.t()
        h_index = pos_h.unsqueeze(-1).repeat(1, args.negatives + 1)
        t_index = pos_t.unsqueeze(-1).repeat(1, args.negatives + 1)
        r_index = pos_r.unsqueeze(-1).repeat(1, args.negatives + 1)
        half = (args.batch // 2 - start) if start < args.batch // 2 else 0     # first half of the GLOBAL batch: tail negatives
        half = max(0, min(half, stop - start))
        t_index[:half, 1:] = negative[:half]
        h_index[half:, 1:] = negative[half:]
        rel_input = rel_model(rel_graph, pos_r)
        pred = model(graph, [rel_input], h_index, t_index, r_index, remove_easy_edges=True)
        target = torch.zeros_like(pred)
        target[:, 0] = 1
        loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, target, reduction="mean")
        optimizer.zero_grad(set_to_none=True)
        (loss * (stop - start) / (args.batch / world)).backward()   # mean over the global batch after the all-reduce
        sharding.all_reduce_gradients(parameters)
        optimizer.step()
        return loss

    for _ in range(2):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = F.launch_count()
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    begin.record()
    for _ in range(args.steps):
        loss = step()
    end.record()
    torch.cuda.synchronize()
    ms = torch.tensor([begin.elapsed_time(end) / args.steps], device=device)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"workload": "ULTRA fine-tuning step on %s-shaped graph" % args.graph, "n_gpus": world,
                          "global_batch": args.batch, "negatives": args.negatives, "ms_per_step": ms.item(),
                          "triples_per_s": args.batch / ms.item() * 1e3, "scaling": "strong",
                          "rspmm_launches_per_step": (F.launch_count() - launches) / args.steps,
                          "loss": float(loss)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
