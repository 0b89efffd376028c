"""CUDA path vs the golden vectors produced by the reference's own message()+aggregate() code
(tests/golden/make_golden.py; the reference tree is not needed at run time)."""
import glob
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "layer_fallback_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[15:-4] for p in GOLDEN])
def test_cuda_matches_reference_fallback(cuda, path):
    from ultra_torchdrug_b200 import functional as F
    g = np.load(path)
    mul = "mul" if "distmult" in path else "add"
    sparse = torch.sparse_coo_tensor(torch.from_numpy(g["indices"]).to(cuda), torch.from_numpy(g["values"]).to(cuda),
                                     tuple(g["shape"]), check_invariants=False)
    relation = torch.from_numpy(g["relation"]).to(cuda).requires_grad_()
    input = torch.from_numpy(g["input"]).to(cuda).requires_grad_()
    out = F.generalized_rspmm(sparse, relation, input, sum="add", mul=mul)
    out.backward(torch.from_numpy(g["grad_output"]).to(cuda))
    # fp32 sums: rtol 1e-5 / atol 1e-6 of the north star, atol scaled to the magnitude of the sums (~10)
    np.testing.assert_allclose(out.detach().cpu().numpy(), g["out_add"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(input.grad.cpu().numpy(), g["grad_input_add"], rtol=1e-5, atol=1e-5)
    grad_relation = relation.grad.cpu().numpy()
    if "grad_relation_add" in g:
        np.testing.assert_allclose(grad_relation, g["grad_relation_add"], rtol=1e-5, atol=1e-5)
    else:
        dim = g["grad_relation_weight_add"].shape[1]
        folded = grad_relation.reshape(grad_relation.shape[0], -1, dim).sum(axis=1)
        np.testing.assert_allclose(folded, g["grad_relation_weight_add"], rtol=1e-5, atol=2e-5)
    if "out_max" in g:
        for sum in ("max", "min"):
            got = F.generalized_rspmm(sparse, relation.detach(), input.detach(), sum=sum, mul=mul)
            assert np.array_equal(got.cpu().numpy(), g["out_" + sum]), "%s not bit-exact" % sum
