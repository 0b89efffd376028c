"""The C-ABI library loads on a CPU-only host and exports every symbol `include/ultra_rspmm.h` declares.
No compute call is made here (no GPU): only argument validation paths that return before touching CUDA."""
import ctypes
import os
import re

from ultra_torchdrug_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    header = open(os.path.join(ROOT, "include", "ultra_rspmm.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    return sorted(set(re.findall(r"\b(ultra_(?:rspmm|layer|score|probe)_[a-z_0-9]+)\s*\(", header)))


def test_library_exports_every_declared_symbol():
    library = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(library, name), "libultra_rspmm.so does not export %s" % name
    assert sorted(_lib.SYMBOLS) == declared, "ctypes binding and header disagree"
    assert os.path.exists(build.LIB_PATH)


def test_struct_layout_matches_header():
    # ultra_rspmm_order_t: 8 x int32 + 8 pointers; ultra_rspmm_index_t: 2 x int64 + 6 x int32 + 3 orders + 2 pointers
    assert ctypes.sizeof(_lib.Order) == 8 * 4 + 8 * 8
    assert ctypes.sizeof(_lib.Pairs) == 2 * 4 + 3 * 8
    assert ctypes.sizeof(_lib.Index) == 2 * 8 + 6 * 4 + 3 * ctypes.sizeof(_lib.Order) + 2 * 8 + 2 * ctypes.sizeof(_lib.Pairs) \
        + 2 * 8 + 2 * 4


def test_version_status_and_argument_validation():
    library = _lib.lib()
    assert library.ultra_rspmm_abi_version() == _lib.ABI_VERSION
    assert library.ultra_rspmm_status_string(_lib.OK) == b"ok"
    assert b"out of range" in library.ultra_rspmm_status_string(_lib.ERR_INDEX)
    index_bytes, scratch_bytes = ctypes.c_size_t(), ctypes.c_size_t()
    assert library.ultra_rspmm_index_bytes(1000, 10, 10, 3, _lib.F32, ctypes.byref(index_bytes),
                                           ctypes.byref(scratch_bytes)) == _lib.OK
    assert index_bytes.value > 1000 * 8 * 3 and scratch_bytes.value > 1000 * 8 * 2
    assert library.ultra_rspmm_index_bytes(-1, 10, 10, 3, _lib.F32, ctypes.byref(index_bytes),
                                           ctypes.byref(scratch_bytes)) == _lib.ERR_ARG
    assert library.ultra_rspmm_index_bytes(10, 10, 10, 3, 7, ctypes.byref(index_bytes),
                                           ctypes.byref(scratch_bytes)) == _lib.ERR_ARG
    assert library.ultra_rspmm_index_bytes(10, 2 ** 31 - 1, 2 ** 31 - 1, 2 ** 20, _lib.F32, ctypes.byref(index_bytes),
                                           ctypes.byref(scratch_bytes)) == _lib.ERR_RANGE
    assert library.ultra_rspmm_forward(None, None, None, None, None, None, 4, _lib.F32, 0, 0, None, 0, None) == _lib.ERR_ARG
    index = _lib.Index()
    index.dtype = _lib.F32
    assert library.ultra_rspmm_forward(ctypes.byref(index), None, None, None, None, None, 4, _lib.F64, 0, 0, None, 0,
                                       None) == _lib.ERR_DTYPE
    assert library.ultra_rspmm_forward(ctypes.byref(index), None, None, None, None, None, 4, _lib.F32, 9, 0, None, 0,
                                       None) == _lib.ERR_ARG
    assert library.ultra_rspmm_set_tuning(-1, 0, 0) == _lib.ERR_ARG
    assert library.ultra_rspmm_launch_count() >= 0


def test_operator_rejects_cpu_tensors_and_unknown_ops():
    import pytest
    import torch
    from ultra_torchdrug_b200 import functional as F
    sparse = torch.sparse_coo_tensor(torch.zeros(3, 0, dtype=torch.long), torch.zeros(0), (2, 2, 2), check_invariants=False)
    with pytest.raises(RuntimeError, match="no CPU implementation"):
        F.generalized_rspmm(sparse, torch.zeros(2, 4), torch.zeros(2, 4))
    with pytest.raises(ValueError):
        F.generalized_rspmm(sparse, torch.zeros(2, 4), torch.zeros(2, 4), sum="mean")
    with pytest.raises(ValueError):
        F.generalized_rspmm(sparse, torch.zeros(2, 4), torch.zeros(2, 4), mul="rotate")
    with pytest.raises(RuntimeError):
        F.generalized_rspmm(sparse, torch.zeros(2, 4), torch.zeros(2, 4, 1))
    # torchdrug's six autograd Functions, plus the two fused `+ boundary` forms of this library
    assert {n for n in dir(F) if n.startswith("RSPMM")} == {
        "RSPMM%s%sFunction" % (s, m) for s in ("Add", "Min", "Max") for m in ("Mul", "Add")} | {"RSPMMAddBoundaryFunction", "RSPMMAddOneHotFunction"}
