"""Development: where the warp roles of the tcgen05 + TMA fused Linear wait (barrier wait cycles per CTA, averaged).

    python tools/linear_stalls.py [--graph fb15k237] [--batch 64]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_torchdrug_b200 import _lib, functional as F, synthetic  # noqa: E402

parser = argparse.ArgumentParser()
parser.add_argument("--graph", default="fb15k237")
parser.add_argument("--batch", type=int, default=64)
args = parser.parse_args()
device = torch.device("cuda", 0)
num_node = synthetic.SHAPES[args.graph][0]
buffers = [torch.randn(num_node, args.batch, 128, device=device) for _ in range(2)]
linear, norm = torch.nn.Linear(128, 64).to(device), torch.nn.LayerNorm(64).to(device)
debug = torch.zeros(8 * 256, dtype=torch.int64, device=device)
lib = _lib.lib()
with torch.no_grad():
    for i in range(3):
        F.linear_norm_relu_residual_into(buffers[i % 2], linear.weight, buffers[(i + 1) % 2][..., :64], linear.bias, norm.weight,
                                         norm.bias, norm.eps)
    torch.cuda.synchronize()
    lib.ultra_layer_linear_set_debug(debug.data_ptr())
    F.linear_norm_relu_residual_into(buffers[1], linear.weight, buffers[0][..., :64], linear.bias, norm.weight, norm.bias, norm.eps)
    torch.cuda.synchronize()
    lib.ultra_layer_linear_set_debug(None)
table = debug.view(-1, 8)[:148].double()
names = ["tma waits empty", "split waits landed", "split waits lo_empty", "mma waits tmem_empty", "mma waits full",
         "epilogue waits tmem_full", "total cycles", "tiles"]
for name, column in zip(names, table.t()):
    print("%-26s mean %10.0f   min %10.0f   max %10.0f" % (name, column.mean(), column.min(), column.max()))
total, tiles = table[:, 6].mean(), table[:, 7].mean()
print("cycles per tile %.0f; fraction of the kernel each role waits: tma %.2f  split %.2f + %.2f  mma %.2f + %.2f  epilogue %.2f"
      % (total / tiles, *(float(table[:, i].mean() / total) for i in range(6))))
