"""ctypes loader for the C restatement (`rspmm_cpu_ref.c`).  TEST / BASELINE INFRASTRUCTURE ONLY.

Follows SURVEY.md Appendix A (torchdrug `rspmm_forward_out_cpu` / `rspmm_backward_out_cpu`, un-vendored);
the math it mirrors is reference ultra/layer.py:52-109.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import rspmm_oracle

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "librspmm_cpu_ref.so")
_lib = None

SUM_CODE = {"add": 0, "min": 1, "max": 2}
MUL_CODE = {"mul": 0, "add": 1}


def build(force=False):
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "rspmm_cpu_ref.c")):
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s", "all"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.rspmm_ref_num_threads.restype = ctypes.c_int
    return _lib


def _ptr(array):
    return array.ctypes.data_as(ctypes.c_void_p)


class CsrOperand(object):
    """Coalesced COO -> int64 CSR exactly as `coo2csr3d` does."""

    def __init__(self, indices, values, shape):
        index, weight, _ = rspmm_oracle.coalesce(indices, values, shape)
        self.shape = tuple(int(s) for s in shape)
        self.col = np.ascontiguousarray(index[1])
        self.layer = np.ascontiguousarray(index[2])
        self.val = np.ascontiguousarray(weight, dtype=np.float32)
        self.row_ptr = np.zeros(self.shape[0] + 1, dtype=np.int64)
        row = np.ascontiguousarray(index[0])
        status = lib().rspmm_ref_coo2csr(ctypes.c_int64(self.shape[0]), ctypes.c_int64(len(row)),
                                         _ptr(row), _ptr(self.row_ptr))
        if status:
            raise RuntimeError("rspmm_ref_coo2csr failed with %d" % status)


def forward(csr, relation, input, sum="add", mul="mul"):
    relation = np.ascontiguousarray(relation, dtype=np.float32)
    input = np.ascontiguousarray(input, dtype=np.float32)
    output = np.empty((csr.shape[0], input.shape[1]), dtype=np.float32)
    status = lib().rspmm_ref_forward_f32(
        ctypes.c_int64(csr.shape[0]), ctypes.c_int64(input.shape[1]), _ptr(csr.row_ptr), _ptr(csr.col),
        _ptr(csr.layer), _ptr(csr.val), _ptr(relation), _ptr(input), _ptr(output),
        ctypes.c_int(SUM_CODE[sum]), ctypes.c_int(MUL_CODE[mul]))
    if status:
        raise RuntimeError("rspmm_ref_forward_f32 failed with %d" % status)
    return output


def backward(csr, relation, input, output, grad_output, sum="add", mul="mul"):
    relation = np.ascontiguousarray(relation, dtype=np.float32)
    input = np.ascontiguousarray(input, dtype=np.float32)
    output = np.ascontiguousarray(output, dtype=np.float32)
    grad_output = np.ascontiguousarray(grad_output, dtype=np.float32)
    grad_relation = np.empty_like(relation)
    grad_input = np.empty_like(input)
    status = lib().rspmm_ref_backward_f32(
        ctypes.c_int64(csr.shape[0]), ctypes.c_int64(csr.shape[1]), ctypes.c_int64(csr.shape[2]),
        ctypes.c_int64(input.shape[1]), _ptr(csr.row_ptr), _ptr(csr.col), _ptr(csr.layer), _ptr(csr.val),
        _ptr(relation), _ptr(input), _ptr(output), _ptr(grad_output), _ptr(grad_relation), _ptr(grad_input),
        ctypes.c_int(SUM_CODE[sum]), ctypes.c_int(MUL_CODE[mul]))
    if status:
        raise RuntimeError("rspmm_ref_backward_f32 failed with %d" % status)
    return grad_relation, grad_input


def forward_arg(csr, relation, input, sum="max", mul="mul"):
    """min/max forward with the arg-index (position of the first edge, in coalesced order, attaining the extremum)."""
    relation = np.ascontiguousarray(relation, dtype=np.float32)
    input = np.ascontiguousarray(input, dtype=np.float32)
    output = np.empty((csr.shape[0], input.shape[1]), dtype=np.float32)
    argidx = np.empty((csr.shape[0], input.shape[1]), dtype=np.int64)
    status = lib().rspmm_ref_forward_arg_f32(
        ctypes.c_int64(csr.shape[0]), ctypes.c_int64(input.shape[1]), _ptr(csr.row_ptr), _ptr(csr.col),
        _ptr(csr.layer), _ptr(csr.val), _ptr(relation), _ptr(input), _ptr(output), _ptr(argidx),
        ctypes.c_int(SUM_CODE[sum]), ctypes.c_int(MUL_CODE[mul]))
    if status:
        raise RuntimeError("rspmm_ref_forward_arg_f32 failed with %d" % status)
    return output, argidx


def forward_f64(csr, relation, input, mul="mul", absolute=False):
    """sum aggregation with fp32 operands up-cast to double (the value fp32 sums are compared with);
    `absolute=True` evaluates it on |w|, |relation|, |input|: the sum of |terms| that scales the tolerance."""
    relation = np.ascontiguousarray(relation, dtype=np.float32)
    input = np.ascontiguousarray(input, dtype=np.float32)
    val = csr.val
    if absolute:
        relation, input, val = np.abs(relation), np.abs(input), np.abs(val)
    output = np.empty((csr.shape[0], input.shape[1]), dtype=np.float64)
    status = lib().rspmm_ref_forward_f64(
        ctypes.c_int64(csr.shape[0]), ctypes.c_int64(input.shape[1]), _ptr(csr.row_ptr), _ptr(csr.col),
        _ptr(csr.layer), _ptr(val), _ptr(relation), _ptr(input), _ptr(output), ctypes.c_int(MUL_CODE[mul]))
    if status:
        raise RuntimeError("rspmm_ref_forward_f64 failed with %d" % status)
    return output


def backward_f64(csr, relation, input, output, grad_output, sum="add", mul="mul", absolute=False):
    """Both gradients accumulated in double; the min/max gate is evaluated in fp32 like the reference's.
    `absolute=True`: the ungated sums of |terms| (tolerance scale)."""
    relation = np.ascontiguousarray(relation, dtype=np.float32)
    input = np.ascontiguousarray(input, dtype=np.float32)
    grad_output = np.ascontiguousarray(grad_output, dtype=np.float32)
    val = csr.val
    if absolute:
        relation, input, grad_output, val, sum, output = np.abs(relation), np.abs(input), np.abs(grad_output), \
            np.abs(val), "add", None
    if sum != "add":
        output = np.ascontiguousarray(output, dtype=np.float32)
    grad_relation = np.empty(relation.shape, dtype=np.float64)
    grad_input = np.empty(input.shape, dtype=np.float64)
    status = lib().rspmm_ref_backward_f64(
        ctypes.c_int64(csr.shape[0]), ctypes.c_int64(csr.shape[1]), ctypes.c_int64(csr.shape[2]),
        ctypes.c_int64(input.shape[1]), _ptr(csr.row_ptr), _ptr(csr.col), _ptr(csr.layer), _ptr(val),
        _ptr(relation), _ptr(input), _ptr(output) if sum != "add" else ctypes.c_void_p(0), _ptr(grad_output),
        _ptr(grad_relation), _ptr(grad_input), ctypes.c_int(SUM_CODE[sum]), ctypes.c_int(MUL_CODE[mul]))
    if status:
        raise RuntimeError("rspmm_ref_backward_f64 failed with %d" % status)
    return grad_relation, grad_input


def num_threads():
    return int(lib().rspmm_ref_num_threads())


def use_all_cores():
    """Use every core this process may run on (torchrun exports OMP_NUM_THREADS=1 for its workers)."""
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    lib().rspmm_ref_set_num_threads(ctypes.c_int(cores))
    return num_threads()
