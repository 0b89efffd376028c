"""Development microbenchmark of the fused layer `combine` (csrc/layer_linear.cu) against its unfused form
(cuBLAS fp32 Linear + fused LayerNorm/ReLU/short-cut epilogue) at a named graph shape.

    python tools/linear_bench.py [--graph fb15k237] [--batch 64] [--dim 64]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ultra_torchdrug_b200 import functional as F, synthetic  # noqa: E402


def timed(fn, iters=20, warmup=3, replays=5):
    """ms per call, kernels only: `iters` calls captured into a CUDA graph and replayed (an eager loop of 0.15 ms kernels
    measures the host's launch path instead)."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(iters):
            fn(i)
    graph.replay()
    torch.cuda.synchronize()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(replays):
        graph.replay()
    stop.record()
    torch.cuda.synchronize()
    return start.elapsed_time(stop) / (iters * replays)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--graph", default="fb15k237")
    parser.add_argument("--batch", type=int, default=64)
    parser.add_argument("--dim", type=int, default=64)
    parser.add_argument("--iters", type=int, default=20)
    parser.add_argument("--kernel", type=int, default=0, help="0 library default, 1 mma.sync, 2 tcgen05")
    args = parser.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = False
    from ultra_torchdrug_b200 import _lib
    _lib.check(_lib.lib().ultra_layer_linear_set_kernel(args.kernel), "ultra_layer_linear_set_kernel")
    device = torch.device("cuda", 0)
    num_node = synthetic.SHAPES[args.graph][0]
    dim = args.dim
    rows = num_node * args.batch
    buffers = [torch.randn(num_node, args.batch, 2 * dim, device=device) for _ in range(2)]   # 2 x 476 MB at C2: > L2
    linear = torch.nn.Linear(2 * dim, dim).to(device)
    norm = torch.nn.LayerNorm(dim).to(device)

    def fused(i):
        source, target = buffers[i % 2], buffers[(i + 1) % 2]
        F.linear_norm_relu_residual_into(source, linear.weight, target[..., :dim], linear.bias, norm.weight, norm.bias,
                                         norm.eps, relu=True, shortcut=True)

    def unfused(i):
        source, target = buffers[i % 2], buffers[(i + 1) % 2]
        projected = torch.nn.functional.linear(source.view(rows, 2 * dim), linear.weight)
        F.layer_norm_relu_residual_into(projected.view(num_node, args.batch, dim), target[..., :dim], norm.weight, norm.bias,
                                        source[..., :dim], norm.eps, True, linear.bias)

    with torch.no_grad():
        t_fused, t_unfused = timed(fused, args.iters), timed(unfused, args.iters)
    with torch.no_grad():      # accuracy of the three ways to run the layer, against a float64 evaluation (first 200k rows)
        sample = buffers[0][:max(1, 200000 // args.batch)]

        def evaluate(dtype, tf32=False):
            x, w = sample.to(dtype), linear.weight.to(dtype)
            if tf32:                                   # operands cut to 10 mantissa bits: the mode the reference switches off
                x = (x.view(torch.int32) & ~0x1fff).view(torch.float32)
                w = (w.view(torch.int32) & ~0x1fff).view(torch.float32)
            hidden = torch.nn.functional.linear(x, w, linear.bias.to(dtype))
            hidden = torch.nn.functional.layer_norm(hidden, (dim,), norm.weight.to(dtype), norm.bias.to(dtype), norm.eps)
            return torch.relu(hidden) + sample.to(dtype)[..., :dim]

        exact = evaluate(torch.float64)
        out = torch.empty_like(sample)
        F.linear_norm_relu_residual_into(sample, linear.weight, out[..., :dim], linear.bias, norm.weight, norm.bias, norm.eps,
                                         relu=True, shortcut=True)
        for name, value in (("fused kernel (3xTF32)", out[..., :dim]), ("cuBLAS fp32 + PyTorch ops", evaluate(torch.float32)),
                            ("plain TF32 operands", evaluate(torch.float32, tf32=True))):
            error = (value.double() - exact).abs()
            print("max |error| vs float64  %-28s %.3e   (mean %.3e)" % (name, float(error.max()), float(error.mean())))
    moved = rows * (2 * dim + dim) * 4 / 1e9
    flops = 2.0 * rows * 2 * dim * dim
    print("rows %d  K %d  N %d" % (rows, 2 * dim, dim))
    print("fused    %.3f ms   %.0f GB/s of compulsory bytes (%.2f GB)   %.1f TFLOP/s fp32-equivalent"
          % (t_fused, moved / t_fused * 1e3, moved, flops / t_fused / 1e9))
    print("unfused  %.3f ms   (cuBLAS fp32 Linear + fused epilogue)" % t_unfused)


if __name__ == "__main__":
    main()
