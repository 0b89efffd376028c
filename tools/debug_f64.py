import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import util
from ultra_torchdrug_b200 import functional as F
cuda = torch.device("cuda", 0)
for dim in (6, 7):
    indices, values = util.random_coo(20, 20, 4, 90, 5 if dim == 6 else 6, 5, "random", False, np.float64)
    shape = (20, 20, 4)
    seed = 5 if dim == 6 else 6
    relation = util.random_dense(4, dim, seed + 1, np.float64)
    input = util.random_dense(20, dim, seed + 2, np.float64)
    index = F.GraphIndex(torch.from_numpy(indices).to(cuda), torch.from_numpy(values).to(cuda), shape)
    out, arg = index.forward(torch.from_numpy(relation).to(cuda), torch.from_numpy(input).to(cuda), "min", "mul", return_argidx=True)
    exp, exp_arg = util.oracle_forward(indices, values, shape, relation, input, "min", "mul")
    out = out.cpu().numpy(); arg = arg.cpu().numpy()
    bad = np.argwhere(out != exp)
    print("dim", dim, "mismatches", len(bad), "arg mismatches", (arg != exp_arg).sum())
    from oracle import rspmm_oracle
    ci, cw, _ = rspmm_oracle.coalesce(indices, values, shape)
    for r, c in bad[:5]:
        e = exp_arg[r, c]
        print(r, c, repr(out[r, c]), repr(exp[r, c]), "arg", arg[r, c], e, "w", repr(cw[e]), "rel", repr(relation[ci[2, e], c]), "in", repr(input[ci[1, e], c]),
              "w*(r*x)", repr(cw[e] * (relation[ci[2, e], c] * input[ci[1, e], c])), "(w*r)*x", repr((cw[e] * relation[ci[2, e], c]) * input[ci[1, e], c]),
              "(w*x)*r", repr((cw[e] * input[ci[1, e], c]) * relation[ci[2, e], c]))
