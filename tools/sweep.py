"""rspmm microbenchmark sweep (BASELINE.json configs[4]): E in {1, 2, 4, 8, 16} M directed edges, N = E / 32, R' in 64..2000,
D in 256..8192, {add, max} x {mul (DistMult), add (TransE)}, forward and forward+backward vs the HBM roofline.

    python tools/sweep.py [--quick] > profiles/r01_sweep.csv

Random uniform edges (seed 1024), fp32.  GB/s under the edge-traffic model of SURVEY.md section 8d.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_torchdrug_b200 import functional as F  # noqa: E402


def edge_bytes(n, r, e, d, sum, mul, which):
    idx = 12 * e + 4 * (n + 1)
    fwd = 4 * d * (e + r + n) + idx
    if which == "fwd":
        return fwd
    if sum != "add":
        bwd = 4 * d * (2 * e + 2 * n + 2 * r) + idx
    elif mul == "add":
        bwd = 4 * d * (e + n + r) + idx
    else:
        bwd = 4 * d * (e + 2 * n + 2 * r) + idx
    return fwd + bwd


def timed(fn, iters):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--quick", action="store_true")
    args = parser.parse_args()
    device = torch.device("cuda", 0)
    peak = 6551.4
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    edges = [1 << 20, 1 << 22] if args.quick else [1 << 20, 1 << 21, 1 << 22, 1 << 23, 1 << 24]
    relations = [64, 2000] if args.quick else [64, 474, 2000]
    dims = [256, 4096] if args.quick else [256, 1024, 4096, 8192]
    ops = [("add", "mul"), ("max", "mul"), ("add", "add"), ("max", "add")]
    print("E_raw,E,N,R,D,sum,mul,index_ms,fwd_ms,fwd_bwd_ms,fwd_GBps,fwd_pct_hbm,fwd_bwd_GBps,fwd_bwd_pct_hbm,G_edge_msgs_per_s_fwd")
    generator = torch.Generator(device=device).manual_seed(1024)
    for e_raw in edges:
        n = e_raw // 32
        for r in relations:
            indices = torch.stack([torch.randint(n, (e_raw,), device=device, generator=generator),
                                   torch.randint(n, (e_raw,), device=device, generator=generator),
                                   torch.randint(r, (e_raw,), device=device, generator=generator)])
            values = torch.ones(e_raw, device=device)
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            index = F.GraphIndex(indices, values, (n, n, r))
            t1.record()
            torch.cuda.synchronize()
            index_ms = t0.elapsed_time(t1)
            e = index.nnz
            for d in dims:
                relation = torch.randn(r, d, device=device, generator=generator)
                input = torch.randn(n, d, device=device, generator=generator)
                grad = torch.randn(n, d, device=device, generator=generator)
                for sum, mul in ops:
                    iters = 3 if e_raw * d >= (1 << 33) else 10
                    out = index.forward(relation, input, sum, mul)
                    fwd = timed(lambda: index.forward(relation, input, sum, mul), iters)
                    both = timed(lambda: index.backward(relation, input, index.forward(relation, input, sum, mul), grad, sum, mul), iters)
                    bf, bb = edge_bytes(n, r, e, d, sum, mul, "fwd") / 1e9, edge_bytes(n, r, e, d, sum, mul, "both") / 1e9
                    print("%d,%d,%d,%d,%d,%s,%s,%.3f,%.3f,%.3f,%.0f,%.1f,%.0f,%.1f,%.1f" % (
                        e_raw, e, n, r, d, sum, mul, index_ms, fwd, both, bf / fwd * 1e3, 100 * bf / fwd * 1e3 / peak,
                        bb / both * 1e3, 100 * bb / both * 1e3 / peak, e * d / fwd / 1e6), flush=True)
                    del out
                del relation, input, grad
                torch.cuda.empty_cache()
            del index, indices, values
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
