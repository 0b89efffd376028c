"""ctypes binding of the C ABI in `include/ultra_rspmm.h` (`libultra_rspmm.so`).

No torch types cross this boundary: callers pass `tensor.data_ptr()` values, sizes and the raw
`cudaStream_t` handle.  There is no CPU fallback: if the library cannot be loaded the import fails.
"""
import ctypes
import os

from . import build as _build

ABI_VERSION = 7

OK, ERR_ARG, ERR_WORKSPACE, ERR_CUDA, ERR_INDEX, ERR_DTYPE, ERR_RANGE = range(7)
SUM_CODE = {"add": 0, "min": 1, "max": 2}
MUL_CODE = {"mul": 0, "add": 1}
F32, F64 = 0, 1

c_void_p, c_int32, c_int64, c_size_t = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_size_t


class Order(ctypes.Structure):
    """`ultra_rspmm_order_t`"""
    _fields_ = [("n_seg", c_int32), ("n_task", c_int32), ("n_slot", c_int32), ("n_split", c_int32),
                ("max_seg_nnz", c_int32), ("pack_shift", c_int32), ("n_gtask", c_int32), ("group_edges", c_int32),
                ("ptr", c_void_p), ("edge", c_void_p), ("w", c_void_p), ("eid", c_void_p), ("packed", c_void_p),
                ("task", c_void_p), ("split", c_void_p), ("gtask", c_void_p)]


class Pairs(ctypes.Structure):
    """`ultra_rspmm_pairs_t`"""
    _fields_ = [("n_pair", c_int32), ("id_bits", c_int32), ("ptr", c_void_p), ("pair", c_void_p), ("rows", c_void_p)]


class Index(ctypes.Structure):
    """`ultra_rspmm_index_t`"""
    _fields_ = [("nnz", c_int64), ("nnz_raw", c_int64), ("n_out", c_int32), ("n_in", c_int32), ("n_rel", c_int32),
                ("dtype", c_int32), ("unit_weight", c_int32), ("chunk", c_int32),
                ("csr", Order), ("csc", Order), ("rel", Order), ("merge_perm", c_void_p), ("merge_start", c_void_p),
                ("pairs", Pairs * 2), ("block_ptr", c_void_p), ("block_split", c_void_p), ("block_rows", c_int32),
                ("n_block", c_int32)]


class PassInfo(ctypes.Structure):
    """`ultra_rspmm_pass_info_t`"""
    _fields_ = [("kernel", c_int32), ("vec", c_int32), ("keep", c_int32), ("grouped", c_int32), ("packed", c_int32),
                ("n_task", c_int32), ("n_slab", c_int32), ("n_split", c_int32)]


PASS_FORWARD, PASS_GRAD_INPUT, PASS_GRAD_RELATION = 0, 1, 2
KERNEL_NAMES = {0: "none", 1: "seg_reduce", 2: "seg_gated", 3: "seg_pna", 4: "rows_in_smem", 5: "dst_blocked",
                6: "pairs_in_smem", 7: "subwarp_rows", 8: "dst_blocked_gated"}

#: every symbol `include/ultra_rspmm.h` declares: name -> (restype, argtypes)
SYMBOLS = {
    "ultra_rspmm_last_pass_info": (ctypes.c_int, [c_int32, ctypes.POINTER(PassInfo)]),
    "ultra_rspmm_set_staged": (ctypes.c_int, [c_int32]),
    "ultra_rspmm_set_extensions": (ctypes.c_int, [c_int32, c_int32]),
    "ultra_rspmm_set_narrow": (ctypes.c_int, [c_int64, c_int32]),
    "ultra_rspmm_index_extend_bytes": (ctypes.c_int, [ctypes.POINTER(Index), ctypes.POINTER(c_size_t)]),
    "ultra_rspmm_index_extend": (ctypes.c_int, [ctypes.POINTER(Index), c_void_p, c_size_t, c_void_p]),
    "ultra_probe_gather": (ctypes.c_int, [c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p, ctypes.POINTER(c_int64),
                                          c_void_p]),
    "ultra_rspmm_abi_version": (ctypes.c_int, []),
    "ultra_rspmm_last_cuda_error": (ctypes.c_int, []),
    "ultra_rspmm_status_string": (ctypes.c_char_p, [ctypes.c_int]),
    "ultra_rspmm_launch_count": (c_int64, []),
    "ultra_rspmm_launch_count_reset": (None, []),
    "ultra_rspmm_set_tuning": (ctypes.c_int, [c_int32, c_int32, c_int64]),
    "ultra_rspmm_index_bytes": (ctypes.c_int, [c_int64, c_int32, c_int32, c_int32, c_int32,
                                               ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t)]),
    "ultra_rspmm_index_build": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32,
                                               c_void_p, c_size_t, c_void_p, c_size_t, ctypes.POINTER(Index), c_void_p]),
    "ultra_rspmm_index_derive_bytes": (ctypes.c_int, [ctypes.POINTER(Index), ctypes.POINTER(c_size_t)]),
    "ultra_rspmm_index_derive": (ctypes.c_int, [ctypes.POINTER(Index), c_void_p, c_void_p, c_size_t, ctypes.POINTER(Index),
                                                c_void_p]),
    "ultra_rspmm_fingerprint": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int32, c_void_p, c_void_p]),
    "ultra_rspmm_workspace_bytes": (ctypes.c_int, [ctypes.POINTER(Index), c_int64, c_int32,
                                                   ctypes.POINTER(c_size_t), ctypes.POINTER(c_size_t)]),
    "ultra_rspmm_forward": (ctypes.c_int, [ctypes.POINTER(Index), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                           c_int32, c_int32, c_int32, c_void_p, c_size_t, c_void_p]),
    "ultra_rspmm_forward_blocked": (ctypes.c_int, [ctypes.POINTER(Index), c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                                   c_int32, c_int32, c_int64, c_int64, c_int64, c_int64, c_void_p, c_size_t,
                                                   c_void_p]),
    "ultra_rspmm_forward_pna": (ctypes.c_int, [ctypes.POINTER(Index), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_void_p, c_int64, c_int32, c_int32, c_void_p, c_size_t, c_void_p]),
    "ultra_rspmm_backward": (ctypes.c_int, [ctypes.POINTER(Index), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                            c_void_p, c_int64, c_int32, c_int32, c_int32, c_void_p, c_size_t, c_void_p]),
    "ultra_rspmm_backward_addend": (ctypes.c_int, [ctypes.POINTER(Index), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                   c_void_p, c_int64, c_int32, c_int32, c_void_p, c_size_t, c_void_p]),
    "ultra_rspmm_ctx_create": (ctypes.c_int, [ctypes.POINTER(c_void_p), c_int32]),
    "ultra_rspmm_ctx_destroy": (ctypes.c_int, [c_void_p]),
    "ultra_rspmm_ctx_set_graph": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_int32, c_int32]),
    "ultra_rspmm_ctx_forward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32]),
    "ultra_rspmm_ctx_forward_backward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                        c_void_p, c_int64, c_int32, c_int32]),
    "ultra_rspmm_ctx_last_kernel_ms": (ctypes.c_float, [c_void_p]),
    "ultra_rspmm_ctx_nnz": (c_int64, [c_void_p]),
    "ultra_rspmm_host_alloc": (ctypes.c_int, [ctypes.POINTER(c_void_p), c_size_t]),
    "ultra_rspmm_host_free": (ctypes.c_int, [c_void_p]),
    "ultra_layer_norm_relu_residual": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                                      ctypes.c_float, c_int32, c_void_p]),
    "ultra_layer_norm_relu_residual_strided": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                              c_int64, c_int32, c_int64, c_int64, ctypes.c_float, c_int32,
                                                              c_void_p]),
    "ultra_layer_norm_relu_residual_backward_bytes": (ctypes.c_int, [c_int32, ctypes.POINTER(c_size_t)]),
    "ultra_layer_norm_relu_residual_backward": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                               c_void_p, c_void_p, c_void_p, c_int64, c_int32, ctypes.c_float,
                                                               c_int32, c_void_p, c_size_t, c_void_p]),
    "ultra_layer_linear_norm_relu_residual": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                                             c_int64, c_int64, c_int32, ctypes.c_float, c_int32, c_int32,
                                                             c_void_p]),
    "ultra_layer_linear_norm_relu_residual_two": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                                                 c_void_p, c_void_p, c_int64, c_int64, c_int32, ctypes.c_float,
                                                                 c_int32, c_int32, c_void_p]),
    "ultra_layer_linear_norm_relu_residual_two_pre": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p,
                                                                     c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                                                     c_int64, c_int32, ctypes.c_float, c_int32, c_int32, c_void_p]),
    "ultra_layer_linear_set_kernel": (ctypes.c_int, [c_int32]),
    "ultra_layer_linear_get_kernel": (ctypes.c_int, []),
    "ultra_layer_linear_set_debug": (ctypes.c_int, [c_void_p]),
    "ultra_layer_rows_gemm": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_int64,
                                             c_void_p, c_int64, c_int64, c_int32, c_int32, c_void_p]),
    "ultra_layer_rows_gemm_weight_bytes": (ctypes.c_int, [ctypes.POINTER(c_size_t)]),
    "ultra_layer_rows_gemm_weight": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p,
                                                    c_void_p, c_size_t, c_void_p]),
    "ultra_score_head_linear": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p,
                                               c_int64, c_int32, c_int32, c_void_p]),
    "ultra_score_head": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p]),
}

_lib = None


class RspmmError(RuntimeError):
    """A C-ABI call returned a non-zero status (reference: TORCH_CHECK failures surface as RuntimeError)."""

    def __init__(self, status, where):
        self.status = status
        detail = lib().ultra_rspmm_status_string(status).decode()
        if status == ERR_CUDA:
            detail += " [cudaError_t %d]" % lib().ultra_rspmm_last_cuda_error()
        super(RspmmError, self).__init__("%s: %s" % (where, detail))


def lib():
    """Load (building first when the .so is absent or stale and nvcc is available) and bind every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("ULTRA_RSPMM_LIB")   # development: load an alternative build of the library
    if not path:
        path = _build.LIB_PATH
        try:
            path = _build.build()
        except Exception as error:
            if not os.path.exists(path):
                raise
            import warnings
            warnings.warn("could not (re)build libultra_rspmm.so (%s); loading the existing %s" % (error, path))
    handle = ctypes.CDLL(path)
    for name, (restype, argtypes) in SYMBOLS.items():
        function = getattr(handle, name)  # AttributeError when the library does not export a declared symbol
        function.restype = restype
        function.argtypes = argtypes
    if handle.ultra_rspmm_abi_version() != ABI_VERSION:
        raise ImportError("libultra_rspmm.so has ABI %d, the binding expects %d - rebuild with "
                          "`python -m ultra_torchdrug_b200.build --force`" % (handle.ultra_rspmm_abi_version(), ABI_VERSION))
    _lib = handle
    return _lib


def pass_info(which):
    """How the last pass of a kind (PASS_*) was launched: dict of the `ultra_rspmm_pass_info_t` fields + kernel name."""
    info = PassInfo()
    check(lib().ultra_rspmm_last_pass_info(which, ctypes.byref(info)), "ultra_rspmm_last_pass_info")
    result = {name: int(getattr(info, name)) for name, _ in PassInfo._fields_}
    result["kernel_name"] = KERNEL_NAMES.get(result["kernel"], "?")
    return result


def check(status, where):
    if status != OK:
        raise RspmmError(status, where)
