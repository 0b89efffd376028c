"""Development microbenchmark: rspmm forward / backward device time on a named synthetic graph.

    python tools/microbench.py --graph fb15k237 --batch 64 [--sum add --mul mul] [--skew 1.0]

Prints one line per pass with the CUDA-event time, the edge-model GB/s and the
fraction of the measured HBM peak.  Inputs rotate over buffers larger than L2 between iterations.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ultra_torchdrug_b200 import functional as F, synthetic  # noqa: E402


def edge_model_bytes(n, r, e, d, passes="fwd", elem=4, gated=False):
    idx = 12 * e + 4 * (n + 1)
    if gated and passes == "bwd":                     # min/max backward (all-ties rule): gathers out[i] and g[i]
        return elem * d * (2 * e + 2 * n + 2 * r) + idx
    if gated and passes in ("bwd_input", "bwd_relation"):
        return elem * d * (2 * e + (n if passes == "bwd_input" else 0) + r + n) + idx
    if passes in ("fwd+addend", "fwd_blocked"):
        return elem * d * (e + r + 2 * n) + idx
    if passes == "fwd":
        return elem * d * (e + r + n) + idx
    if passes == "bwd_input":
        return elem * d * (e + r + n) + idx           # gather g[i], table rel, write grad_in
    if passes == "bwd_relation":
        return elem * d * (2 * e + r) + idx           # gather g[i] and in[j], write grad_rel
    if passes == "bwd":
        return elem * d * (e + 2 * n + 2 * r) + idx   # SURVEY.md 8d graded figure
    raise ValueError(passes)


def timed(fn, iters, warmup=3):
    for _ in range(warmup):
        fn(0)
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(iters):
        fn(i)
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / iters


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--graph", default="fb15k237")
    parser.add_argument("--batch", type=int, default=64)
    parser.add_argument("--sum", default="add")
    parser.add_argument("--mul", default="mul")
    parser.add_argument("--skew", type=float, default=None)
    parser.add_argument("--relskew", type=float, default=None)
    parser.add_argument("--iters", type=int, default=10)
    parser.add_argument("--chunk", type=int, default=0)
    parser.add_argument("--l2mb", type=int, default=0, help="L2 budget (MiB) used to pick the slab width")
    parser.add_argument("--variant", type=int, default=0, help="0 auto, 1 no L2 eviction hints, 2 hints always")
    parser.add_argument("--relgraph", action="store_true", help="dense 4-relation graph over R' nodes instead")
    parser.add_argument("--staged", type=int, default=-1, help="rows-in-shared-memory kernel: 0 off, 1 auto, 2 always")
    parser.add_argument("--blocked", type=int, default=-1, help="destination-blocked grad_relation: 0 never, 1 auto, 2 always")
    parser.add_argument("--uniform", default=None, help="E:N:R - uniform random graph (BASELINE configs[4] sweep shapes)")
    parser.add_argument("--dim", type=int, default=0, help="feature width (overrides --batch * 64)")
    args = parser.parse_args()
    device = torch.device("cuda", 0)
    peak = 6551.4
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    if args.chunk or args.l2mb or args.variant:
        from ultra_torchdrug_b200 import _lib
        _lib.lib().ultra_rspmm_set_tuning(args.chunk, args.variant, args.l2mb << 20)
    if args.staged >= 0:
        from ultra_torchdrug_b200 import _lib
        _lib.lib().ultra_rspmm_set_staged(args.staged)
    if args.blocked >= 0:
        from ultra_torchdrug_b200 import _lib
        _lib.lib().ultra_rspmm_set_extensions(1, args.blocked)
    if args.uniform:
        e_raw, n, r = (int(v) for v in args.uniform.split(":"))
        generator = torch.Generator().manual_seed(1024)
        edge_list = torch.stack([torch.randint(n, (e_raw,), generator=generator), torch.randint(n, (e_raw,), generator=generator),
                                 torch.randint(r, (e_raw,), generator=generator)], dim=1)
        args.graph = "uniform " + args.uniform
    else:
        edge_list, n, r = synthetic.named_graph(args.graph, skew=args.skew, relation_skew=args.relskew)
    if args.relgraph:
        nodes = r
        grid = torch.cartesian_prod(torch.arange(nodes), torch.arange(nodes), torch.arange(4))
        edge_list, n, r = grid[:, [1, 0, 2]].contiguous(), nodes, 4
    sparse = synthetic.operator_operand(edge_list, n, r, device)
    d = args.dim or args.batch * 64
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    F.graph_index(sparse)  # warm (CUB kernels load lazily)
    F.clear_index_cache()
    sparse = synthetic.operator_operand(edge_list, n, r, device)
    t0.record()
    index = F.graph_index(sparse)
    t1.record()
    torch.cuda.synchronize()
    e = index.nnz
    print("graph %s: N=%d R'=%d E_raw=%d E=%d D=%d  index build %.3f ms  tasks csr/csc/rel = %d/%d/%d  slots %d/%d/%d  max seg %d/%d/%d"
          % (args.graph, n, r, edge_list.shape[0], e, d, t0.elapsed_time(t1), index.c.csr.n_task, index.c.csc.n_task,
             index.c.rel.n_task, index.c.csr.n_slot, index.c.csc.n_slot, index.c.rel.n_slot, index.c.csr.max_seg_nnz,
             index.c.csc.max_seg_nnz, index.c.rel.max_seg_nnz))
    generator = torch.Generator(device=device).manual_seed(1024)
    copies = max(2, int(300e6 // (n * d * 4)) + 1)
    inputs = [torch.randn(n, d, device=device, generator=generator) for _ in range(copies)]
    grads = [torch.randn(n, d, device=device, generator=generator) for _ in range(copies)]
    relation = torch.randn(r, d, device=device, generator=generator)
    out = index.forward(relation, inputs[0], args.sum, args.mul)

    results = {}
    results["fwd"] = timed(lambda i: index.forward(relation, inputs[i % copies], args.sum, args.mul), args.iters)
    results["bwd_input"] = timed(lambda i: index.backward(relation, inputs[i % copies], out, grads[i % copies], args.sum,
                                                          args.mul, need_relation=False), args.iters)
    results["bwd_relation"] = timed(lambda i: index.backward(relation, inputs[i % copies], out, grads[i % copies],
                                                             args.sum, args.mul, need_input=False), args.iters)
    results["bwd"] = timed(lambda i: index.backward(relation, inputs[i % copies], out, grads[i % copies], args.sum,
                                                    args.mul), args.iters)
    if args.sum == "add":        # the two forms the layers actually call: + boundary, and the cat-free blocked layout
        results["fwd+addend"] = timed(lambda i: index.forward(relation, inputs[i % copies], "add", args.mul,
                                                               addend=grads[i % copies]), args.iters)
        if d == args.batch * 64 and n * d < (1 << 30):
            buffers = [torch.randn(n, args.batch, 128, device=device, generator=generator) for _ in range(copies)]
            results["fwd_blocked"] = timed(lambda i: index.forward_blocked(relation, buffers[i % copies], buffers[i % copies],
                                                                           64, 0, 64, args.mul, addend=grads[i % copies]),
                                           args.iters)
    for name, ms in results.items():
        gb = edge_model_bytes(n, r, e, d, name, gated=args.sum != "add") / 1e9
        print("%-13s %8.3f ms   %8.1f GB/s edge-model  (%.1f%% of measured HBM %.0f GB/s)   %.2f G edge-msg/s"
              % (name, ms, gb / ms * 1e3, 100 * gb / ms * 1e3 / peak, peak, e * d / ms / 1e6))
    total = results["fwd"] + results["bwd"]
    gb = (edge_model_bytes(n, r, e, d, "fwd") + edge_model_bytes(n, r, e, d, "bwd", gated=args.sum != "add")) / 1e9
    print("fwd+bwd       %8.3f ms   %8.1f GB/s edge-model  (%.1f%% of measured HBM)" % (total, gb / total * 1e3, 100 * gb / total * 1e3 / peak))


if __name__ == "__main__":
    main()
