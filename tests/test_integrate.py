"""`integrate.patch_torchdrug` against a package with torchdrug's layout (the import shims stand in for it here)."""
import torch

from ultra_torchdrug_b200 import functional, integrate
from ultra_torchdrug_b200.compat import torchdrug as fake_torchdrug


def test_patch_binds_operator_and_memoises_transpose():
    fake_torchdrug.layers.functional.generalized_rspmm = None
    integrate.patch_torchdrug(fake_torchdrug)
    integrate.patch_torchdrug(fake_torchdrug)     # idempotent
    assert fake_torchdrug.layers.functional.generalized_rspmm is functional.generalized_rspmm
    graph = fake_torchdrug.data.Graph(torch.tensor([[0, 1, 0], [1, 2, 1], [2, 0, 1]]), num_node=3, num_relation=2)
    first = graph.adjacency
    assert graph.adjacency is first
    assert first.transpose(0, 1) is first.transpose(0, 1)
    assert torch.equal(first.transpose(0, 1)._indices()[0], torch.tensor([1, 2, 0]))
    other = fake_torchdrug.data.Graph(torch.tensor([[0, 1, 0]]), num_node=3, num_relation=2)
    assert other.adjacency is not first


def test_memoise_transpose_on_plain_sparse_tensor():
    sparse = torch.sparse_coo_tensor(torch.tensor([[0, 1], [1, 0], [0, 0]]), torch.ones(2), (2, 2, 1), check_invariants=False)
    integrate.memoise_transpose(sparse)
    assert sparse.transpose(0, 1) is sparse.transpose(1, 0)
    assert integrate.memoise_transpose(sparse) is sparse
