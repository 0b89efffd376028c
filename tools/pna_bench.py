"""Development: PNA aggregation - four operator calls vs the fused one-pass kernel (C2 shape)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ultra_torchdrug_b200 import functional as F, synthetic
device = torch.device("cuda", 0)
edge_list, n, r = synthetic.named_graph(sys.argv[1] if len(sys.argv) > 1 else "fb15k237")
sparse = synthetic.operator_operand(edge_list, n, r, device)
d = 64 * 64
relation, input = torch.randn(r, d, device=device), torch.randn(n, d, device=device)
def timed(fn, iters=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters
with torch.no_grad():
    four = timed(lambda: (F.generalized_rspmm(sparse, relation, input, sum="add"), F.generalized_rspmm(sparse, relation ** 2, input ** 2, sum="add"),
                          F.generalized_rspmm(sparse, relation, input, sum="max"), F.generalized_rspmm(sparse, relation, input, sum="min")))
    fused = timed(lambda: F.rspmm_pna(sparse, relation, input))
print("pna aggregates: four calls %.3f ms, fused %.3f ms (%.2fx)" % (four, fused, four / fused))
