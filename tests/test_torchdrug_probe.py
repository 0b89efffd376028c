"""Probe for the reference's REAL operator: if `torchdrug` is importable on this machine -
it is not in this image: un-vendored, no wheel offline - every op combination of `generalized_rspmm`, INCLUDING the min / max
gradients (the all-ties rule that this repo otherwise pins only on its recollection of torchdrug's `NaryMax::backward`,
SURVEY.md Appendix A [ext-recall]), is compared with it.  Skipped with a reason when torchdrug is absent."""
import os
import sys

import numpy as np
import pytest
import torch

from tests import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _real_torchdrug():
    extra = os.path.join(ROOT, "baseline", "_ref")
    if os.path.isdir(extra) and extra not in sys.path:
        sys.path.append(extra)
    try:
        from torchdrug.layers import functional as td_functional
    except Exception as error:                                   # noqa: BLE001 (any import failure means "absent")
        return None, "torchdrug is not importable here (%s): the all-ties min/max gradient rule stays [ext-recall]" % type(error).__name__
    if "ultra_torchdrug_b200" in (getattr(td_functional, "__file__", "") or ""):
        return None, "`torchdrug` resolves to this repo's import shim, not to the real package"
    return td_functional.generalized_rspmm, ""


def _case(seed, ties):
    indices, values = util.random_coo(60, 50, 5, 700, seed=seed, duplicates=40, weights="random", skew=True)
    relation, input = util.random_dense(5, 96, seed + 1, ties=ties), util.random_dense(50, 96, seed + 2, ties=ties)
    grad = util.random_dense(60, 96, seed + 3)
    return indices, values, (60, 50, 5), relation, input, grad


@pytest.mark.parametrize("sum,mul", util.OPS)
@pytest.mark.parametrize("ties", [False, True])
def test_oracle_equals_real_torchdrug_cpu(sum, mul, ties):
    real, why = _real_torchdrug()
    if real is None:
        pytest.skip(why)
    from oracle import rspmm_oracle
    indices, values, shape, relation, input, grad = _case(3, ties)
    sparse = torch.sparse_coo_tensor(torch.from_numpy(indices), torch.from_numpy(values), shape)
    a, b = torch.from_numpy(relation).requires_grad_(), torch.from_numpy(input).requires_grad_()
    out = real(sparse, a, b, sum=sum, mul=mul)
    out.backward(torch.from_numpy(grad))
    want, _ = rspmm_oracle.rspmm_forward(indices, values, shape, relation, input, sum, mul)
    if sum == "add":
        np.testing.assert_allclose(out.detach().numpy(), want, rtol=1e-5, atol=1e-5)
    else:
        assert np.array_equal(out.detach().numpy(), want)
    g_rel, g_in = rspmm_oracle.rspmm_backward(indices, values, shape, relation, input, out.detach().numpy(), grad, sum, mul)
    np.testing.assert_allclose(a.grad.numpy(), g_rel, rtol=1e-4, atol=1e-4)
    np.testing.assert_allclose(b.grad.numpy(), g_in, rtol=1e-4, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("sum,mul", util.OPS)
@pytest.mark.parametrize("ties", [False, True])
def test_cuda_operator_equals_real_torchdrug(cuda, sum, mul, ties):
    real, why = _real_torchdrug()
    if real is None:
        pytest.skip(why)
    from ultra_torchdrug_b200 import functional as F
    indices, values, shape, relation, input, grad = _case(5, ties)
    sparse = util.to_sparse(indices, values, shape, cuda)
    results = []
    for operator in (real, F.generalized_rspmm):
        a = torch.from_numpy(relation).to(cuda).requires_grad_()
        b = torch.from_numpy(input).to(cuda).requires_grad_()
        out = operator(sparse, a, b, sum=sum, mul=mul)
        out.backward(torch.from_numpy(grad).to(cuda))
        results.append((out.detach(), a.grad, b.grad))
    (want, want_rel, want_in), (got, got_rel, got_in) = results
    if sum == "add":
        torch.testing.assert_close(got, want, rtol=1e-5, atol=1e-5)
    else:
        assert torch.equal(got, want)
    torch.testing.assert_close(got_rel, want_rel, rtol=1e-4, atol=1e-4)      # torchdrug's CUDA backward sums with float atomics
    torch.testing.assert_close(got_in, want_in, rtol=1e-4, atol=1e-4)
