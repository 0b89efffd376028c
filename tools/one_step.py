"""One rspmm forward + backward between cudaProfilerStart / Stop, for ncu captures of exactly one step of a shape:

    ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/traffic_c4.csv python tools/one_step.py --graph yago310 --batch 64

`tools/traffic_from_csv.py` turns the launch lists into profiles/traffic.json (DRAM bytes per forward and per step).
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ultra_torchdrug_b200 import _lib, functional as F, synthetic  # noqa: E402


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("--graph", default="fb15k237")
    parser.add_argument("--batch", type=int, default=64)
    parser.add_argument("--uniform", default=None, help="E:N:R - uniform random graph (BASELINE configs[4] sweep shapes)")
    parser.add_argument("--dim", type=int, default=0)
    parser.add_argument("--sum", default="add")
    parser.add_argument("--mul", default="mul")
    args = parser.parse_args()
    device = torch.device("cuda", 0)
    if args.uniform:
        e_raw, n, r = (int(v) for v in args.uniform.split(":"))
        generator = torch.Generator().manual_seed(1024)
        edge_list = torch.stack([torch.randint(n, (e_raw,), generator=generator), torch.randint(n, (e_raw,), generator=generator),
                                 torch.randint(r, (e_raw,), generator=generator)], dim=1)
    else:
        edge_list, n, r = synthetic.named_graph(args.graph)
    index = F.graph_index(synthetic.operator_operand(edge_list, n, r, device))
    d = args.dim or args.batch * 64
    generator = torch.Generator(device=device).manual_seed(1024)
    relation = torch.randn(r, d, device=device, generator=generator)
    inputs = [torch.randn(n, d, device=device, generator=generator) for _ in range(2)]
    grads = [torch.randn(n, d, device=device, generator=generator) for _ in range(2)]
    out = index.forward(relation, inputs[0], args.sum, args.mul)                   # warm: workspaces, lazy module load
    index.backward(relation, inputs[0], out, grads[0], args.sum, args.mul)
    torch.cuda.synchronize()
    lib = _lib.lib()
    lib.ultra_rspmm_launch_count_reset()
    torch.cuda.cudart().cudaProfilerStart()
    out = index.forward(relation, inputs[1], args.sum, args.mul)
    torch.cuda.synchronize()
    forward_launches = int(lib.ultra_rspmm_launch_count())
    index.backward(relation, inputs[1], out, grads[1], args.sum, args.mul)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("forward_launches %d step_launches %d" % (forward_launches, int(lib.ultra_rspmm_launch_count())))
    print("shape N=%d R'=%d E=%d D=%d: forward %s, grad_input %s, grad_relation %s" % (
        n, r, index.nnz, d, *(_lib.pass_info(which)["kernel_name"] for which in (_lib.PASS_FORWARD, _lib.PASS_GRAD_INPUT, _lib.PASS_GRAD_RELATION))))


if __name__ == "__main__":
    main()
